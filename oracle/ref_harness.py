"""Drive the UNMODIFIED reference headlessly -- TEST INFRASTRUCTURE ONLY.

Runs only in the build container, where /root/reference exists; nothing under
tests/ (gpu marker), smoke() or bench.py imports this module.  It is used by
``oracle/gen_golden.py`` to produce the committed fixtures in tests/golden/ and
by the CPU tests that pin ``oracle/cosmo_oracle.c`` against the real reference
when the reference tree is present.

How the reference is made to run:
  * ``oracle/pygame_stub`` precedes the reference on sys.path (pygame absent);
  * ``random.randint`` is wrapped to int()-cast its bounds -- Python 3.12
    rejects the float bounds the reference passes at universe_old.py:53,55,
    planet.py:47 and examples/star_with_disk.py:38-39;
  * the frame loop is bounded through the reference's own ``run_while`` hook
    (universe_old.py:308,338);
  * per-step tensors are captured by wrapping the module-level names
    ``acc_blas`` / ``pairwise_distances`` of universe_old (universe_old.py:6,8)
    and the merge stream by wrapping ``Planet.absorb`` (planet.py:74).
No reference source is copied; the reference is imported from where it lies.
"""
from __future__ import annotations

import os
import random
import sys
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("COSMOSIM_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))

AU = 1.496e11
ME = 5.972e24
MS = 1.989e30
RS = 6.9634e8


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cosmosim", "core", "universe_old.py"))


_loaded = None


def load_reference():
    """Import the reference's old-API modules; returns (universe_old, planet, functions, blas)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    stub = os.path.join(_HERE, "pygame_stub")
    for p in (REFERENCE_ROOT, stub):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, stub)
    if not getattr(random.randint, "_cosmo_patched", False):
        _orig = random.randint

        def _randint(a, b):
            return _orig(int(a), int(b))

        _randint._cosmo_patched = True
        random.randint = _randint
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import cosmosim.core.universe_old as uo  # noqa: E402
    import cosmosim.core.planet as pl  # noqa: E402
    import cosmosim.util.functions as fn  # noqa: E402
    import cosmosim.util.blas as bl  # noqa: E402
    assert uo.__file__.startswith(REFERENCE_ROOT), uo.__file__
    _loaded = (uo, pl, fn, bl)
    return _loaded


# ----------------------------------------------------------------------------------------
# scene builders: the two BASELINE example scripts restated as functions of a seed
# (examples/star_with_disk.py:12-43, examples/cloud.py:6-16) using the REFERENCE classes.
# ----------------------------------------------------------------------------------------
def build_star_with_disk(seed: int, n_planets: int = 1000):
    uo, pl, fn, _ = load_reference()
    random.seed(seed)
    u = uo.Universe()
    star = u.create_planet(mass=MS, radius=RS, position=[0, 0], immobile=True,
                           name="Sol", color=(255, 255, 0))
    for _ in range(n_planets):
        dist = random.randint(0.05 * AU, 1 * AU)
        mass = random.randint(0.1 * ME, 100 * ME)
        star.create_satellite(distance=dist, mass=mass, radius=fn.get_radius(mass, 3000))
    return u


def build_cloud(seed: int, n_planets: int = 1000, omega: float = 1e-9):
    import math
    uo, pl, fn, _ = load_reference()
    random.seed(seed)
    u = uo.Universe()
    for _ in range(n_planets):
        p = u.random_planet()
        p.velocity = fn.rotation(p.position * omega, math.pi / 2)
    return u


def snapshot(u):
    """SoA copy of the reference universe's full planet list (insertion order)."""
    n = len(u.planets)
    out = {
        "mass_f64": np.array([float(p.mass) for p in u.planets], dtype=np.float64),
        "mass_is_int": np.array([isinstance(p.mass, int) for p in u.planets], dtype=np.bool_),
        "mass_int_hi": np.zeros(n, dtype=np.uint64),
        "mass_int_lo": np.zeros(n, dtype=np.uint64),
        "radius": np.array([float(p.radius) for p in u.planets], dtype=np.float64),
        "volume": np.array([float(p.volume) for p in u.planets], dtype=np.float64),
        "pos": np.zeros((n, 2), dtype=np.float64),
        "vel": np.zeros((n, 2), dtype=np.float64),
        "alive": np.array([bool(p.alive) for p in u.planets], dtype=np.bool_),
        "immobile": np.array([bool(p.immobile) for p in u.planets], dtype=np.bool_),
        "eaten": np.array([int(p.planets_eaten) for p in u.planets], dtype=np.int64),
    }
    for k, p in enumerate(u.planets):
        if isinstance(p.mass, int):
            out["mass_int_hi"][k] = (p.mass >> 64) & 0xFFFFFFFFFFFFFFFF
            out["mass_int_lo"][k] = p.mass & 0xFFFFFFFFFFFFFFFF
        if p.alive:
            out["pos"][k] = p.position
            out["vel"][k] = p.velocity
    return out


class Recorder:
    """Capture per-step tensors and the merge stream of a reference run."""

    def __init__(self, u, keep_steps=()):
        self.u = u
        self.keep = set(int(s) for s in keep_steps)
        self.events = []          # (step, absorber_global_idx, absorbed_global_idx)
        self.steps = {}           # step -> dict of arrays
        self._cur = None

    def run(self, steps, collisions=True, speed=1e6, fps=33):
        uo, pl, fn, bl = load_reference()
        u = self.u
        index_of = {id(p): k for k, p in enumerate(u.planets)}
        rec = self

        orig_acc, orig_pd, orig_absorb = uo.acc_blas, uo.pairwise_distances, pl.Planet.absorb
        orig_upl = uo.Universe.update_planet_lists

        def upl(self_u):
            orig_upl(self_u)
            if self_u.iterations in rec.keep:
                rec._cur = {"pre": snapshot(self_u),
                            "active_idx": np.array([index_of[id(p)] for p in self_u.active_planets],
                                                   dtype=np.int64)}
            else:
                rec._cur = None

        def acc(pos, mas, G=1):
            a = orig_acc(pos, mas, G)
            if rec._cur is not None:
                rec._cur["acc"] = np.array(a[:, :2], dtype=np.float64)
            return a

        def pd(p, **kw):
            d = orig_pd(p, **kw)
            if rec._cur is not None:
                rec._cur["p_new"] = np.array(p, dtype=np.float64)
                radii = np.array([q.radius for q in u.active_planets])
                flags = np.clip(d, 1, None) <= np.add.outer(radii, radii)
                np.fill_diagonal(flags, False)
                ii, jj = np.nonzero(flags)
                rec._cur["flag_pairs"] = np.stack([ii, jj], axis=1).astype(np.int64)
            return d

        def absorb(self_p, other):
            if self_p.alive and other.alive:
                rec.events.append((u.iterations, index_of[id(self_p)], index_of[id(other)]))
            return orig_absorb(self_p, other)

        uo.acc_blas, uo.pairwise_distances, pl.Planet.absorb = acc, pd, absorb
        uo.Universe.update_planet_lists = upl
        try:
            def cond(U):
                # runs at the top of every frame; harvest the previous frame here
                prev = U.iterations - 1
                if rec._cur is not None and prev in rec.keep:
                    rec._cur["post"] = snapshot(U)
                    rec.steps[prev] = rec._cur
                    rec._cur = None
                return U.iterations < steps
            u.simulate(run_while=cond, collisions=collisions, speed=speed, fps=fps)
        finally:
            uo.acc_blas, uo.pairwise_distances, pl.Planet.absorb = orig_acc, orig_pd, orig_absorb
            uo.Universe.update_planet_lists = orig_upl
        return self
