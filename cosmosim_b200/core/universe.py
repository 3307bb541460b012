"""Host-side ``Universe``: drop-in for the reference's README/examples API
(cosmosim/core/universe_old.py:22-87,306-362) with the per-frame physics update executed by
the sm_100a kernels behind the C ABI, plus a headless stepping mode so no UI sits in the
timed loop.

What changed relative to the reference: ``interact()`` no longer rebuilds numpy arrays from
the planet list and walks an N x N Python loop; it enqueues one frame on the GPU.  The planet
list stays the public view: it is refreshed from the device lazily (first attribute read
after stepping), and any host-side mutation is re-uploaded before the next frame.
"""
from __future__ import annotations

import math
import random

import numpy as np

from .. import engine as _engine
from .._abi import F_IMMOBILE, F_MASS_INT
from ..util import functions as F
from .planet import AU, ME, MS, RE, RS, Planet  # noqa: F401  (re-exported constants)

DAYTIME = 86400


def _randint(lo, hi):
    # Python >= 3.12 rejects non-integer bounds; the reference passes floats (universe_old.py:53,55)
    return random.randint(int(lo), int(hi))


class Universe:
    def __init__(self, G=6.674e-11, bounded=True, r_max=100 * AU, precision="fp64", device=None,
                 track_energy=False):
        self.G = G
        self.bounded = bounded
        self.r_max = r_max
        self.planets = []
        self._angular_momentum = 0.0
        self._total_energy = 0.0
        # energies / angular momentum are an extra O(N^2) pass (universe_old.py:181-189); the HUD
        # numbers are only evaluated when asked for
        self.track_energy = track_energy
        self.clicked_planet = None
        self.running = False
        self.paused = False
        self.dragging = False
        self.iterations = 0
        # simulate() settings with the reference's defaults (universe_old.py:306-320)
        self.collisions = True
        self.trail_length = 1000
        self.speed = 1e6
        self.fps = 33
        self.dt_default = self.speed / self.fps
        self.dt = self.dt_default
        # device side
        self.precision = precision
        self.device = device
        self._engine = None
        self._device_ahead = False   # GPU state newer than the Planet objects
        self._host_dirty = True      # Planet objects newer than the GPU state
        self._frames = 0             # frames executed on the device (stamps the merge stream)
        self._events_seen = 0
        self._merge_events = []      # (frame, absorber index, absorbed index) in visitation order
        self.active_planets = []
        self.absorbed_planets = []
        self.bound_planets = []
        self.unbound_planets = []
        self.escaped_planets = []

    @property
    def total_energy(self):
        """sum(K + U) of the last frame (universe_old.py:187); needs track_energy=True."""
        self._pull_if_needed()
        return self._total_energy

    @property
    def angular_momentum(self):
        """sum(m * cross(p, v)) of the last frame (universe_old.py:189); needs track_energy=True."""
        self._pull_if_needed()
        return self._angular_momentum

    @property
    def merge_events(self):
        """Merge stream so far: (frame, absorber index, absorbed index), visitation order."""
        self._pull_if_needed()
        return self._merge_events

    # ------------------------------------------------------------------ world model
    def add_planet(self, planet):
        if self._device_ahead:
            self._pull()
        planet.index = len(self.planets)
        self.planets.append(planet)
        planet.universe = self
        self._host_dirty = True

    def create_planet(self, **kwargs):
        planet = Planet(**kwargs)
        self.add_planet(planet)
        return planet

    def random_planet(self, density=4000, min_mass=0.1 * ME, max_mass=1000 * ME, dmin=0, dmax=50 * AU,
                      min_velocity=0, max_velocity=5e4, name=None, color=None):
        """Random planet; draw order mass, distance, angle, speed, heading (universe_old.py:50-61)."""
        mass = _randint(min_mass, max_mass)
        radius = F.get_radius(mass, density)
        dist = _randint(dmin, dmax)
        angle = random.random() * (2 * math.pi)
        position = F.to_cartesian(dist, angle)
        speed = min_velocity + (max_velocity - min_velocity) * random.random()
        heading = 2 * math.pi * random.random()
        velocity = [speed * math.cos(heading), speed * math.sin(heading)]
        return self.create_planet(mass=mass, radius=radius, position=position, velocity=velocity,
                                  color=color, name=name)

    def get_active_planets(self):
        self._pull_if_needed()
        return [p for p in self.planets if p._alive]

    def get_absorbed_planets(self):
        self._pull_if_needed()
        return [p for p in self.planets if not p._alive]

    def get_bound_planets(self):
        return [p for p in self.get_active_planets() if p._energy < 0]

    def get_unbound_planets(self):
        return [p for p in self.get_active_planets() if p._energy >= 0]

    def get_escaped_planets(self):
        return [p for p in self.get_active_planets() if np.linalg.norm(p._position) > self.r_max]

    def destroy_escaped_planets(self):
        for p in self.escaped_planets:
            p.destroy()

    def update_planet_lists(self):
        self.active_planets = self.get_active_planets()
        self.absorbed_planets = self.get_absorbed_planets()
        self.bound_planets = self.get_bound_planets()
        self.unbound_planets = self.get_unbound_planets()
        self.escaped_planets = self.get_escaped_planets()

    # ------------------------------------------------------------------ host <-> device
    def _pull_if_needed(self):
        if self._device_ahead:
            self._pull()

    def _gather(self):
        """AoS -> SoA over the ALIVE planets in insertion order (universe_old.py:170-172,192)."""
        act = [p for p in self.planets if p._alive]
        n = len(act)
        a = {"pos": np.zeros((n, 2)), "vel": np.zeros((n, 2)), "mass_f64": np.zeros(n),
             "mass_hi": np.zeros(n, dtype=np.uint64), "mass_lo": np.zeros(n, dtype=np.uint64),
             "flags": np.zeros(n, dtype=np.uint8), "radius": np.zeros(n), "volume": np.zeros(n),
             "eaten": np.zeros(n, dtype=np.int64), "id": np.zeros(n, dtype=np.int32)}
        for k, p in enumerate(act):
            f, is_int, hi, lo = _engine.split_mass(p._mass)
            a["pos"][k] = p._position
            a["vel"][k] = p._velocity
            a["mass_f64"][k] = f
            a["mass_hi"][k], a["mass_lo"][k] = hi, lo
            a["flags"][k] = (F_IMMOBILE if p._immobile else 0) | (F_MASS_INT if is_int else 0)
            a["radius"][k] = p._radius
            a["volume"][k] = p._volume
            a["eaten"][k] = p._planets_eaten
            a["id"][k] = p.index
        return a

    def _push(self):
        for k, p in enumerate(self.planets):
            p.index = k
        n_total = len(self.planets)
        if self._engine is None or self._engine.n_total < n_total:
            self._engine = _engine.Engine(max(n_total, 1), device=self.device)
            self._events_seen = 0
        self._engine.upload(self._gather(), frame=self._frames)
        self._host_dirty = False
        self._device_ahead = False

    def _pull(self):
        """Refresh every Planet from the device state (and the merge stream)."""
        self._device_ahead = False  # first: the setters below must not recurse
        eng = self._engine
        d = eng.download()
        if d["status"]["error"]:
            eng.check_errors()
        alive_ids = set()
        ids = d["id"]
        for k in range(ids.shape[0]):
            p = self.planets[int(ids[k])]
            alive_ids.add(int(ids[k]))
            p._position = d["pos"][k].copy()
            p._velocity = d["vel"][k].copy()
            fl = int(d["flags"][k])
            if fl & F_MASS_INT:
                p._mass = (int(d["mass_hi"][k]) << 64) | int(d["mass_lo"][k])
            else:
                p._mass = float(d["mass"][k])
            p._radius = float(d["radius"][k])
            p._volume = float(d["volume"][k])
            p._planets_eaten = int(d["eaten"][k])
            if self.track_energy:
                p._energy = float(d["energy"][k])
        if self.track_energy:
            self._total_energy, self._angular_momentum = d["total_energy"], d["angular_momentum"]
        for idx in np.nonzero(d["rec_death"] >= 0)[0]:
            p = self.planets[int(idx)]
            if p._alive:  # died on the device since the last pull: Planet.destroy() semantics
                fl = int(d["rec_flags"][idx])
                if fl & F_MASS_INT:
                    p._mass = (int(d["rec_mass_hi"][idx]) << 64) | int(d["rec_mass_lo"][idx])
                else:
                    p._mass = float(d["rec_mass"][idx])
                p._radius = float(d["rec_radius"][idx])
                p._volume = float(d["rec_volume"][idx])
                p._planets_eaten = int(d["rec_eaten"][idx])
                p._alive = False
                p._velocity = np.array([0, 0])
                p.acceleration = np.array([0, 0])
                p._position = None
                p._energy = 0
        ev = eng.events(self._events_seen)
        self._events_seen += ev.shape[0]
        self._merge_events.extend((int(f), int(a), int(b)) for f, a, b in ev)

    # ------------------------------------------------------------------ bulk (array) interface
    @classmethod
    def from_arrays(cls, mass, radius, position, velocity=None, immobile=None, names=None, **universe_kw):
        """Build a universe from SoA arrays without paying Planet.__init__ per planet (large N).

        mass[n] (float or Python-int object array), radius[n], position[n,2], velocity[n,2],
        immobile[n] bool.  Planets get slim proxies whose position/velocity are views into the
        given arrays; no random numbers are drawn (colour is white, names are 'p<index>')."""
        import math as _m
        u = cls(**universe_kw)
        n = len(radius)
        position = np.ascontiguousarray(position, dtype=float)
        velocity = np.zeros((n, 2)) if velocity is None else np.ascontiguousarray(velocity, dtype=float)
        immobile = np.zeros(n, dtype=bool) if immobile is None else np.asarray(immobile, dtype=bool)
        radius = np.asarray(radius, dtype=float)
        planets = []
        for k in range(n):
            p = Planet.__new__(Planet)
            m = mass[k]
            p.__dict__.update(universe=u, name=names[k] if names is not None else "p%d" % k,
                              _mass=m if isinstance(m, int) else float(m), _radius=float(radius[k]),
                              _volume=((4 / 3) * _m.pi) * (float(radius[k]) ** 3), density=0.0,   # as Planet.__init__
                              _position=position[k], _velocity=velocity[k],
                              _energy=0, color=(255, 255, 255), _immobile=bool(immobile[k]), history=[],
                              bound=False, clicked=False, tracked=False, _alive=True, _planets_eaten=0, index=k)
            p.density = p._mass / p._volume
            planets.append(p)
        u.planets = planets
        u._host_dirty = True
        return u

    def arrays(self):
        """Current ACTIVE planets as SoA numpy arrays straight from the device (no Planet objects are
        touched): dict(id, pos, vel, mass, radius, volume, eaten, flags[, energy])."""
        if self._engine is None or self._host_dirty:
            self._push()
        d = self._engine.download()
        if d["status"]["error"]:
            self._engine.check_errors()
        return {k: d[k] for k in ("id", "pos", "vel", "mass", "radius", "volume", "eaten", "flags", "energy")}

    def view(self):
        """(n, id, pos, radius[, energy ...]) of the active planets straight from the device, for a
        front-end: neither the Planet objects nor the death records are touched."""
        if self._engine is None or self._host_dirty:
            self._push()
        return self._engine.view(energy=self.track_energy)

    def peek(self, slot):
        """One active planet by its slot in `view()` (position, velocity, mass, radius, eaten, energy)."""
        return self._engine.peek(slot)

    # ------------------------------------------------------------------ stepping
    def step(self, n=1, dt=None, collisions=None):
        """Headless stepping: n frames of update_planet_lists -> interact -> destroy_escaped
        (universe_old.py:345-350,361) on the GPU, asynchronously.  Nothing is copied back
        until a planet attribute, a planet list or ``merge_events`` is read."""
        if dt is not None:
            self.dt = dt
        if collisions is not None:
            self.collisions = collisions
        if self.paused or n <= 0:
            self.iterations += max(0, n)
            return self
        if self._host_dirty or self._engine is None:
            self._push()
        self._engine.step(n, self.dt, self.G, self.r_max, collisions=self.collisions, bounded=self.bounded,
                          fp32=(self.precision == "fp32"), energy=self.track_energy)
        self._frames += n
        self.iterations += n
        self._device_ahead = True
        return self

    def interact(self):
        """One physics update on the current active planets (universe_old.py:166-212)."""
        if self.paused:
            return
        if self._host_dirty or self._engine is None:
            self._push()
        # interact() alone never destroys escaped planets: that is simulate()'s job (:349-350)
        self._engine.step(1, self.dt, self.G, self.r_max, collisions=self.collisions, bounded=False,
                          fp32=(self.precision == "fp32"), energy=self.track_energy)
        self._frames += 1
        self._device_ahead = True

    def synchronize(self):
        """Block until the enqueued frames are done and surface device-side errors."""
        if self._engine is not None:
            self._engine.check_errors()
        return self

    def simulate(self, width=1600, height=1000, speed=1e6, fps=33, trail_length=1000, collisions=True,
                 track_all=False, run_while=(lambda x: True), scale=1e-7, start_paused=False,
                 headless=True, steps=None):
        """Frame loop with the reference's signature (universe_old.py:306-309).

        headless=True (default here): no window, no event pump; runs `steps` frames in one
        burst when given, else one frame per ``run_while(self)`` evaluation.  headless=False
        needs pygame and drives the optional viewer in cosmosim_b200.viewer."""
        self.collisions = collisions
        self.trail_length = trail_length
        self.default_scale = scale
        self.width, self.height = width, height
        self.fps, self.speed = fps, speed
        self.dt_default = speed / fps
        self.dt = self.dt_default
        self.paused = start_paused
        if track_all:
            for p in self.planets:
                p.tracked = True
        self.context = {"scale": scale, "offset": np.array([0.0, 0.0]), "origin": np.array([width / 2, height / 2])}
        if not headless:
            from .. import viewer
            return viewer.run(self, run_while)
        self.running = True
        if steps is not None:
            self.step(int(steps))
        else:
            while self.running and run_while(self):
                self.step(1)
        self.running = False
        return self
