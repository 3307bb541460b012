"""The reference's NEW (3-D, batch) API -- ``Object`` / ``State`` / ``Universe`` of
cosmosim/core/universe.py -- with ``State.interact`` executed by the CUDA library.

Same names, arguments and results as the reference:
  Object(mass, density, position, velocity=[0,0,0], name=None, color=None)   universe.py:21-70
  State(objects, dt=1, G=6.674e-11, iteration=0) .masses() .radii() .positions() .velocities()
      .colors() .interact(collisions=True) .save(f)                          universe.py:73-133
  Universe(objects, dt, iterations, outpath=None, filesize=1000).run()       universe.py:136-167
``State.save`` writes one pickle per call exactly like the reference; the pickles name the classes
``cosmosim.core.universe.State`` / ``.Object``, so the reference's replay tools
(cosmosim/core/animation.py:32-45) read files written here and ``load_states`` reads files written
by either implementation.

While a State is only stepped, its objects live on the device; ``state.objects`` (or any of the
array getters) pulls them back, and a later ``interact`` re-uploads because the caller may have
edited them.  ``State.step(n)`` runs n steps without returning to the host.
"""
from __future__ import annotations

import copy
import io
import math
import os
import pickle
import random

import numpy as np

from ..util import functions as F
from .planet import generate_name

AU = 1.496e11
ME = 5.972e24
RE = 6.371e6
MS = 1.989e30
RS = 6.9634e8
DAYTIME = 86400
_G = 6.674e-11

_REF_MODULE = "cosmosim.core.universe"


class Object:
    """universe.py:21-70."""

    def __init__(self, mass, density, position, velocity=[0, 0, 0], name=None, color=None):
        self.exists = True
        self.mass = mass
        self.density = density
        self.position = np.array(position).astype(float)
        self.velocity = np.array(velocity).astype(float)
        self.name = name or generate_name()
        self.color = color or (int(255 * random.random()), int(255 * random.random()), int(255 * random.random()))

    def __getstate__(self):
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_cosmo")}

    def get_volume(self):
        return self.mass / self.density

    def get_radius(self):
        return ((3 * self.get_volume()) / (4 * math.pi)) ** (1 / 3)

    def absorb(self, obj):
        """Inelastic merge on the host (universe.py:39-46); the device resolver does the same arithmetic."""
        m_total = self.mass + obj.mass
        self.position = ((self.mass * self.position) + (obj.mass * obj.position)) / m_total
        self.velocity = ((self.mass * self.velocity) + (obj.mass * obj.velocity)) / m_total
        self.density = ((self.mass * self.density) + (obj.mass * obj.density)) / m_total
        self.mass += obj.mass
        obj.destroy()

    def destroy(self):
        self.exists = False
        self.mass = 0.0
        self.volume = 0.0
        self.radius = 0.0
        self.velocity = np.array([0.0, 0.0, 0.0])
        self.position = np.array([0.0, 0.0, 0.0])

    def create_satellite(self, distance=None, mass=None, density=None, theta=None, name=None, color=None, G=_G):
        """Circular-orbit satellite in the z = 0 plane (universe.py:56-70); consumes the `random`
        stream in the reference's order (distance, mass, theta, then the colour)."""
        distance = distance or random.randint(int(self.radius * 5), int(self.radius * 100))
        mass = mass or random.random() * self.mass
        density = density or self.density
        theta = theta or 2 * math.pi * random.random()
        v_mag = math.sqrt((self.mass * G) / distance)
        pos = F.to_cartesian(distance, theta)
        pos_norm = pos / np.linalg.norm(pos)
        v = np.append(F.rotation(v_mag * pos_norm, math.pi / 2), [0])
        pos = np.append(pos, [0])
        return Object(mass, density, pos, v, name, color)


def _python_mass(flags, hi, lo, f):
    return (int(hi) << 64) | int(lo) if flags & 1 else float(f)


class State:
    """universe.py:73-133.  `precision`: "fp64" reproduces the reference's arithmetic; "fp32" is the
    fast all-pairs flavour (state and integrator stay FP64)."""

    def __init__(self, objects, dt=1, G=_G, iteration=0, precision="fp64", device=None, group=None):
        self._objects = objects
        self.group = group           # torch.distributed process group: rows sharded over its ranks
        self.dt = dt
        self.G = G
        self._iteration = iteration
        self.precision = precision
        self.device = device
        self._engine = None
        self._uploaded = None        # the Object list as uploaded (ids index into it)
        self._device_ahead = False   # the device has stepped since the host objects were last valid
        self._host_dirty = True      # the host objects may have been edited since the last upload
        self._events_seen = 0
        self._next_index = 0
        # Object.absorb calls so far: (iteration, absorber, absorbed, absorbed still existed), objects
        # numbered in the order this State first saw them (the initial list order)
        self.absorb_calls = []
        self._bulk = None            # from_arrays(): the host truth is a dict of arrays, not Object instances

    @classmethod
    def from_arrays(cls, mass, density, position, velocity=None, dt=1, G=_G, iteration=0, precision="fp64",
                    device=None, group=None):
        """A State of len(mass) bodies from column data, without a Python Object per body (large N).
        mass / density: numpy arrays or sequences of Python numbers (ints stay exact, as with Object);
        position / velocity: [n, 3].  `state.objects` materialises Object instances on first use;
        masses() / positions() / velocities() / radii() / arrays() read the arrays directly."""
        from .. import engine3d
        st = cls(None, dt=dt, G=G, iteration=iteration, precision=precision, device=device, group=group)
        st._bulk = engine3d.arrays_from_columns(mass, density, position, velocity)
        return st

    def _materialise(self):
        """Bulk mode -> Object instances (numbered by their original row)."""
        b = self._bulk
        objs = []
        for k in range(b["mass_f64"].shape[0]):
            fl = int(b["flags"][k])
            o = Object(_python_mass(fl, b["mass_hi"][k], b["mass_lo"][k], b["mass_f64"][k]),
                       int(b["density"][k]) if fl & 2 else float(b["density"][k]), b["pos"][k], b["vel"][k],
                       name="body%d" % int(b["id"][k]), color=(255, 255, 255))
            o._cosmo_index = int(b["id"][k])
            objs.append(o)
        self._next_index = max([self._next_index] + [o._cosmo_index + 1 for o in objs])
        self._objects, self._bulk = objs, None
        self._host_dirty = True      # device ids are re-issued from the new list at the next step

    # -------------------------------------------------------------- host <-> device
    @property
    def objects(self):
        self._pull()
        if self._objects is None:
            self._materialise()
        self._host_dirty = True      # the caller may edit what it gets
        return self._objects

    def _host_objects(self):
        """The up-to-date Object list for internal readers that do not edit it."""
        self._pull()
        if self._objects is None:
            self._materialise()
        return self._objects

    def _column(self, key):
        """Bulk mode: one column of the up-to-date arrays (None when the State holds Objects)."""
        if self._objects is not None:
            return None
        self._pull()
        return self._bulk[key]

    @objects.setter
    def objects(self, value):
        self._pull()
        self._objects = value
        self._host_dirty = True

    @property
    def iteration(self):
        return self._iteration

    @iteration.setter
    def iteration(self, value):
        self._pull()                 # otherwise the next pull would overwrite it with the device's counter
        self._iteration = value
        self._host_dirty = True

    def _push(self):
        from .. import engine3d
        if self._objects is None:                # bulk mode: the arrays go up as they are
            n = self._bulk["mass_f64"].shape[0]
            if self._engine is None or self._engine.cap < n:
                self._engine = engine3d.Engine3D(n, device=self.device, group=self.group)
            self._engine.upload(self._bulk, iteration=self._iteration)
            self._uploaded = None
            self._events_seen = self._engine.status()["n_events"]
            self._host_dirty = False
            self._device_ahead = False
            return
        objs = self._objects
        for o in objs:
            if getattr(o, "_cosmo_index", None) is None:
                o._cosmo_index = self._next_index
                self._next_index += 1
        if self._engine is None or self._engine.cap < len(objs):
            self._engine = engine3d.Engine3D(len(objs), device=self.device, group=self.group)
        self._engine.upload(engine3d.arrays_from_objects(objs), iteration=self._iteration)
        self._uploaded = list(objs)
        self._events_seen = self._engine.status()["n_events"]
        self._host_dirty = False
        self._device_ahead = False

    def _pull(self):
        if not self._device_ahead:
            return
        eng = self._engine
        eng.check_errors()
        d = eng.download()
        ev = eng.events(self._events_seen)
        self._events_seen += len(ev)
        if self._objects is None:                # bulk mode: ids are the original rows
            self.absorb_calls.extend((int(a), int(b), int(c), int(e)) for a, b, c, e in ev)
            self._bulk = {"pos": d["pos"], "vel": d["vel"], "mass_f64": d["mass"], "mass_hi": d["mass_hi"],
                          "mass_lo": d["mass_lo"], "density": d["density"], "radius": d["radius"],
                          "flags": d["flags"], "id": d["id"]}
            self._iteration = d["status"]["iteration"]
            self._device_ahead = False
            return
        src = self._uploaded
        self.absorb_calls.extend((int(a), src[int(b)]._cosmo_index, src[int(c)]._cosmo_index, int(e))
                                 for a, b, c, e in ev)
        alive = set()
        out = []
        for k in range(d["mass"].shape[0]):
            o = src[int(d["id"][k])]
            alive.add(int(d["id"][k]))
            fl = int(d["flags"][k])
            o.mass = _python_mass(fl, d["mass_hi"][k], d["mass_lo"][k], d["mass"][k])
            o.density = int(d["density"][k]) if fl & 2 else float(d["density"][k])
            o.position = d["pos"][k].copy()
            o.velocity = d["vel"][k].copy()
            out.append(o)
        for k, o in enumerate(src):
            if k not in alive and o.exists:
                o.destroy()
        # in place: the reference removes absorbed objects from the very list the caller handed over
        # (universe.py:126-128), so a second State / Universe built on that list starts from the survivors
        self._objects[:] = out
        self._iteration = d["status"]["iteration"]
        self._device_ahead = False      # device ids keep indexing `src`, so no re-upload is needed

    # -------------------------------------------------------------- the reference's getters
    def masses(self):
        col = self._column("mass_f64")
        return col.copy() if col is not None else np.array([o.mass for o in self._host_objects()])

    def radii(self):
        col = self._column("radius")
        return col.copy() if col is not None else np.array([o.get_radius() for o in self._host_objects()])

    def positions(self):
        col = self._column("pos")
        return col.copy() if col is not None else np.array([o.position for o in self._host_objects()])

    def velocities(self):
        col = self._column("vel")
        return col.copy() if col is not None else np.array([o.velocity for o in self._host_objects()])

    def colors(self):
        return [o.color for o in self._host_objects()]

    # -------------------------------------------------------------- stepping
    def step(self, n=1, collisions=True):
        """n x interact() on the GPU without returning to the host."""
        if self._host_dirty or self._engine is None:
            self._pull()
            self._push()
        self._engine.step(self.dt, self.G, collisions=bool(collisions), fp32=(self.precision == "fp32"), steps=n)
        self._iteration += int(n)
        self._device_ahead = True

    def interact(self, collisions=True):
        """universe.py:94-129."""
        self.step(1, collisions)

    def arrays(self):
        """Device state as numpy arrays without touching the Object list (pos/vel [n,3], mass, density,
        radius, id, flags)."""
        if self._engine is None or (self._host_dirty and not self._device_ahead):
            self._push()
        return self._engine.download()

    # -------------------------------------------------------------- persistence
    def __getstate__(self):
        return {"objects": self._host_objects(), "dt": self.dt, "G": self.G, "iteration": self._iteration}

    def __setstate__(self, d):
        self.__init__(d["objects"], dt=d.get("dt", 1), G=d.get("G", _G), iteration=d.get("iteration", 0))

    def __deepcopy__(self, memo):
        return State(copy.deepcopy(self._host_objects(), memo), dt=self.dt, G=self.G, iteration=self._iteration,
                     precision=self.precision, device=self.device, group=self.group)

    def save(self, f):
        """universe.py:131-132: one pickle of the State appended to the open file."""
        f.write(dumps_reference(self))


def dumps_reference(state) -> bytes:
    """Pickle a State so that it names the reference's classes (cosmosim.core.universe.State / Object).
    Protocol 3 keeps class references as plain `c<module>\\n<name>\\n` opcodes, which are renamed in place."""
    raw = pickle.dumps(state, protocol=3)
    mine = State.__module__.encode()
    for cls in (b"State", b"Object"):
        raw = raw.replace(b"c" + mine + b"\n" + cls + b"\n", b"c" + _REF_MODULE.encode() + b"\n" + cls + b"\n")
    return raw


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module in (_REF_MODULE, State.__module__) and name in ("State", "Object"):
            return {"State": State, "Object": Object}[name]
        return super().find_class(module, name)


def load_states(path):
    """Read every State of a data directory (Universe.run(outpath=...)) or of one .dat file, written by
    this package or by the reference; mirrors the loader of cosmosim/core/animation.py:32-45.
    The format is the reference's (a stream of pickles): only load files you trust."""
    files = [path] if os.path.isfile(path) else [os.path.join(path, f) for f in os.listdir(path)]
    states = []
    for fn in files:
        with open(fn, "rb") as f:
            buf = io.BytesIO(f.read())
        while True:
            try:
                states.append(_Unpickler(buf).load())
            except EOFError:
                break
    return states


class Universe:
    """universe.py:136-167."""

    def __init__(self, objects, dt, iterations, outpath=None, filesize=1000, precision="fp64"):
        self.objects = objects
        self.dt = dt
        self.iterations = iterations
        self.outpath = outpath
        self.filesize = filesize
        self.precision = precision

    def run(self, save_every=1):
        """universe.py:141-167: `iterations` steps; every state goes to `outpath` as a stream of pickles
        (`filesize` states per file) or, without an outpath, into the returned list.

        The steps between two saves run as ONE burst on the GPU (State.step) and the objects are only
        pulled back to the host when a state is written or kept.  save_every=1 is the reference's
        behaviour (every step is saved); a larger value keeps every save_every-th state and lets the
        device run ahead in between."""
        try:
            from tqdm import tqdm
        except ImportError:  # pragma: no cover
            def tqdm(it, **kw):
                return it
        save_every = max(1, int(save_every))
        state = State(self.objects, dt=self.dt, precision=self.precision)
        total = int(self.iterations)
        # the save points: every save_every-th step, and the last one
        points = list(range(save_every, total + 1, save_every))
        if total > 0 and (not points or points[-1] != total):
            points.append(total)
        if not self.outpath:
            states, done = [], 0
            for upto in tqdm(points, desc="Running simulation"):
                state.step(upto - done)
                done = upto
                states.append(copy.deepcopy(state))
            return states
        if not os.path.isdir(self.outpath):
            os.mkdir(self.outpath)
        # the reference empties the directory (universe.py:147-149); only its own data files here
        for f in os.listdir(self.outpath):
            if f.endswith(".dat") and f[:-4].isdigit():
                os.remove(os.path.join(self.outpath, f))
        nfiles = math.ceil(len(points) / self.filesize)
        print(f"A total of {nfiles} data files will be created.")
        # the reference calls state.interact(self.dt) here (universe.py:155): dt lands in the
        # `collisions` argument, so a zero dt would switch collisions off
        collisions = bool(self.dt)
        done = 0
        for n in range(nfiles):
            with open(os.path.join(self.outpath, f"{n}.dat"), "ab+") as f:
                for upto in tqdm(points[n * self.filesize:(n + 1) * self.filesize], desc=f"Writing file {n}"):
                    state.step(upto - done, collisions)
                    done = upto
                    state.save(f)
        return None
