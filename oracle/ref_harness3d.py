"""Drive the UNMODIFIED reference's new API (cosmosim/core/universe.py: Object / State /
Universe.run) -- TEST INFRASTRUCTURE ONLY, build container only (needs /root/reference).

The one thing that keeps State.interact from running with the scikit-learn in this image is the
``force_all_finite=`` keyword it passes to pairwise_distances (universe.py:107; removed in
sklearn 1.8).  The harness rebinds the module-level name ``pairwise_distances`` of the
reference module to a wrapper that drops that keyword and calls the real sklearn function with
the remaining arguments -- no reference source is modified or copied.
"""
from __future__ import annotations

import random

import numpy as np

from . import ref_harness as rh

AU, ME, MS = rh.AU, rh.ME, rh.MS
DS = 1408

_mod = None


def load_reference3d():
    global _mod
    if _mod is not None:
        return _mod
    rh.load_reference()                      # sys.path, pygame stub, randint patch
    import cosmosim.core.universe as un      # noqa: E402
    from sklearn.metrics import pairwise_distances as skpd

    def pairwise_distances(p, n_jobs=None, force_all_finite=True, **kw):
        return skpd(p, n_jobs=n_jobs, **kw)

    un.pairwise_distances = pairwise_distances
    assert un.__file__.startswith(rh.REFERENCE_ROOT), un.__file__
    _mod = un
    return un


def build_disk(seed: int, n_planets: int = 1000, d_min=0.1 * AU, d_max=2 * AU, density=3000,
               z_sigma: float = 0.0):
    """examples/testing/in-memory.py:12-42 (defaults) / save-data.py:11-40 as a function of a seed.
    z_sigma > 0 additionally lifts the planets out of the plane (the examples are planar; State
    itself is fully 3-D) by drawing z offsets and z velocities AFTER the example's own RNG draws."""
    un = load_reference3d()
    random.seed(seed)
    star = un.Object(mass=MS, density=DS, position=[0, 0, 0], name="Sol", color=(255, 255, 0))
    planets = []
    for _ in range(n_planets):
        dist = random.randint(d_min, d_max)
        mass = random.randint(0.1 * ME, 100 * ME)
        planets.append(star.create_satellite(distance=dist, mass=mass, density=density))
    if z_sigma:
        rng = np.random.default_rng(seed)
        for p in planets:
            p.position[2] = rng.normal() * z_sigma
            p.velocity[2] = rng.normal() * 50.0
    return [star, *planets]


def snapshot(objects):
    """SoA copy of a State.objects list.  Objects carry `_id` (insertion index)."""
    n = len(objects)
    out = {
        "pos": np.array([o.position for o in objects], dtype=np.float64).reshape(n, 3),
        "vel": np.array([o.velocity for o in objects], dtype=np.float64).reshape(n, 3),
        "mass_f64": np.array([float(o.mass) for o in objects], dtype=np.float64),
        "mass_is_int": np.array([isinstance(o.mass, int) for o in objects], dtype=np.bool_),
        "mass_int_hi": np.zeros(n, dtype=np.uint64),
        "mass_int_lo": np.zeros(n, dtype=np.uint64),
        "dens_f64": np.array([float(o.density) for o in objects], dtype=np.float64),
        "dens_is_int": np.array([isinstance(o.density, int) for o in objects], dtype=np.bool_),
        "id": np.array([o._id for o in objects], dtype=np.int64),
    }
    for k, o in enumerate(objects):
        if isinstance(o.mass, int):
            out["mass_int_hi"][k] = (o.mass >> 64) & 0xFFFFFFFFFFFFFFFF
            out["mass_int_lo"][k] = o.mass & 0xFFFFFFFFFFFFFFFF
    return out


def snapshot_with_ids(objects):
    for k, o in enumerate(objects):
        o._id = k
    return snapshot(objects)


class Recorder3D:
    """Step a reference State and capture per-step tensors + the Object.absorb call stream."""

    def __init__(self, objects, dt, G=6.674e-11, keep_steps=()):
        un = load_reference3d()
        for k, o in enumerate(objects):
            o._id = k
        self.state = un.State(objects, dt=dt, G=G)
        self.keep = set(int(s) for s in keep_steps)
        self.events = []      # (iteration, absorber id, absorbed id, absorbed existed)
        self.steps = {}
        self.n_objects = []

    def run(self, steps, collisions=True):
        un = load_reference3d()
        rec, st = self, self.state
        orig_acc, orig_pd, orig_absorb = un.acc_blas, un.pairwise_distances, un.Object.absorb
        cur = {}

        def acc(pos, mas, G=1):
            a = orig_acc(pos, mas, G)
            if cur.get("on"):
                cur["acc"] = np.array(a, dtype=np.float64)
            return a

        def pd(p, **kw):
            d = orig_pd(p, **kw)
            if cur.get("on"):
                cur["p_new"] = np.array(p, dtype=np.float64)
                r = st.radii()
                cur["radius"] = np.array(r, dtype=np.float64)
                fl = d <= np.add.outer(r, r)
                np.fill_diagonal(fl, False)
                ii, jj = np.nonzero(fl)
                cur["flag_pairs"] = np.stack([ii, jj], axis=1).astype(np.int64)
            return d

        def absorb(self_o, other):
            rec.events.append((st.iteration, self_o._id, other._id, int(bool(other.exists))))
            return orig_absorb(self_o, other)

        un.acc_blas, un.pairwise_distances, un.Object.absorb = acc, pd, absorb
        try:
            for _ in range(steps):
                it = st.iteration
                cur.clear()
                if it in rec.keep:
                    cur["on"] = True
                    cur["pre"] = snapshot(st.objects)
                st.interact(collisions)
                rec.n_objects.append(len(st.objects))
                if it in rec.keep:
                    cur["post"] = snapshot(st.objects)
                    cur.pop("on")
                    rec.steps[it] = dict(cur)
        finally:
            un.acc_blas, un.pairwise_distances, un.Object.absorb = orig_acc, orig_pd, orig_absorb
        return self
