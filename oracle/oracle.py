"""ctypes front-end of oracle/cosmo_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module (see the header of cosmo_oracle.c).  The product package
cosmosim_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcosmo_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_up = C.POINTER(C.c_uint64)
_bp = C.POINTER(C.c_uint8)


class _State(C.Structure):
    _fields_ = [("n", C.c_int64), ("pos", _dp), ("vel", _dp), ("mass_f64", _dp),
                ("mass_hi", _up), ("mass_lo", _up), ("mass_is_int", _bp),
                ("radius", _dp), ("volume", _dp), ("alive", _bp), ("immobile", _bp),
                ("eaten", _ip)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("cosmo_oracle.c", "state3d_oracle.c")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_accel_rows.argtypes = [C.c_int64, _dp, _dp, C.c_double, C.c_int64, C.c_int64, _dp]
        L.oracle_accel_rows_plain.argtypes = [C.c_int64, _dp, _dp, C.c_double, C.c_int64, C.c_int64, _dp]
        L.oracle_d2_rows.argtypes = [C.c_int64, _dp, C.c_int64, C.c_int64, _dp]
        L.oracle_integrate.argtypes = [C.c_int64, _dp, _dp, _dp, C.c_double, _dp, _dp]
        L.oracle_flags_rows.argtypes = [C.c_int64, _dp, _dp, C.c_int64, C.c_int64, _ip, C.c_int64]
        L.oracle_flags_rows.restype = C.c_int64
        L.oracle_dist_rows.argtypes = [C.c_int64, _dp, C.c_int64, C.c_int64, _dp]
        L.oracle_energies.argtypes = [C.c_int64, _dp, _dp, _dp, C.c_double, _dp, _dp]
        L.oracle_frame.argtypes = [C.POINTER(_State), C.c_double, C.c_double, C.c_int, C.c_int,
                                   C.c_double, _ip, _dp, _dp, _dp, _ip, _ip, _ip, _ip, C.c_int64]
        L.oracle_frame.restype = C.c_int64
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _c64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def accel(pos, mass, G, i_begin=0, i_end=None, compensated=True):
    """A1 (blas.py:4-41): accelerations [rows,2] of targets i_begin..i_end from all n sources.
    compensated=False: plain ascending accumulation (the CPU-baseline timing leg of bench.py)."""
    pos, mass = _c64(pos), _c64(mass)
    n = mass.shape[0]
    i_end = n if i_end is None else i_end
    out = np.empty((i_end - i_begin, 2), dtype=np.float64)
    fn = lib().oracle_accel_rows if compensated else lib().oracle_accel_rows_plain
    fn(n, _d(pos), _d(mass), float(G), i_begin, i_end, _d(out))
    return out


def d2_matrix(pos, i_begin=0, i_end=None):
    pos = _c64(pos)
    n = pos.shape[0]
    i_end = n if i_end is None else i_end
    out = np.empty((i_end - i_begin, n), dtype=np.float64)
    lib().oracle_d2_rows(n, _d(pos), i_begin, i_end, _d(out))
    return out


def integrate(p0, v0, a, dt):
    """A2 (universe_old.py:177-178)."""
    p0, v0, a = _c64(p0), _c64(v0), _c64(a)
    p, v = np.empty_like(p0), np.empty_like(v0)
    lib().oracle_integrate(p0.shape[0], _d(p0), _d(v0), _d(a), float(dt), _d(p), _d(v))
    return p, v


def flags(p, radius, i_begin=0, i_end=None):
    """A3 (universe_old.py:180,192-193): flagged ordered pairs (i,j), row-major."""
    p, radius = _c64(p), _c64(radius)
    n = radius.shape[0]
    i_end = n if i_end is None else i_end
    cap = 1 << 16
    while True:
        pairs = np.empty((cap, 2), dtype=np.int64)
        cnt = lib().oracle_flags_rows(n, _d(p), _d(radius), i_begin, i_end,
                                      pairs.ctypes.data_as(_ip), cap)
        if cnt <= cap:
            return pairs[:cnt].copy()
        cap = int(cnt)


def dist_matrix(p, i_begin=0, i_end=None):
    p = _c64(p)
    n = p.shape[0]
    i_end = n if i_end is None else i_end
    out = np.empty((i_end - i_begin, n), dtype=np.float64)
    lib().oracle_dist_rows(n, _d(p), i_begin, i_end, _d(out))
    return out


def energies(p, v, m, G):
    """A7 (universe_old.py:180-189,210): per-planet K+U, (total energy, angular momentum)."""
    p, v, m = _c64(p), _c64(v), _c64(m)
    n = m.shape[0]
    e, tot = np.empty(n), np.empty(2)
    lib().oracle_energies(n, _d(p), _d(v), _d(m), float(G), _d(e), _d(tot))
    return e, float(tot[0]), float(tot[1])


class OracleUniverse:
    """SoA state + frame stepping following universe_old.py:345-350 (see cosmo_oracle.c)."""

    FIELDS = ("pos", "vel", "mass_f64", "mass_int_hi", "mass_int_lo", "mass_is_int", "radius",
              "volume", "alive", "immobile", "eaten")

    def __init__(self, snap, G=6.674e-11, bounded=True, r_max=100 * 1.496e11):
        self.G, self.bounded, self.r_max = float(G), bool(bounded), float(r_max)
        self.n = int(snap["radius"].shape[0])
        self.pos = np.ascontiguousarray(snap["pos"], dtype=np.float64).copy()
        self.vel = np.ascontiguousarray(snap["vel"], dtype=np.float64).copy()
        self.mass_f64 = np.ascontiguousarray(snap["mass_f64"], dtype=np.float64).copy()
        self.mass_int_hi = np.ascontiguousarray(snap["mass_int_hi"], dtype=np.uint64).copy()
        self.mass_int_lo = np.ascontiguousarray(snap["mass_int_lo"], dtype=np.uint64).copy()
        self.mass_is_int = np.ascontiguousarray(snap["mass_is_int"]).astype(np.uint8)
        self.radius = np.ascontiguousarray(snap["radius"], dtype=np.float64).copy()
        self.volume = np.ascontiguousarray(snap["volume"], dtype=np.float64).copy()
        self.alive = np.ascontiguousarray(snap["alive"]).astype(np.uint8)
        self.immobile = np.ascontiguousarray(snap["immobile"]).astype(np.uint8)
        self.eaten = np.ascontiguousarray(snap["eaten"], dtype=np.int64).copy()
        self.iterations = 0

    def snapshot(self):
        out = {k: getattr(self, k).copy() for k in self.FIELDS}
        for k in ("mass_is_int", "alive", "immobile"):
            out[k] = out[k].astype(np.bool_)
        return out

    def _cstate(self):
        return _State(self.n, _d(self.pos), _d(self.vel), _d(self.mass_f64),
                      self.mass_int_hi.ctypes.data_as(_up), self.mass_int_lo.ctypes.data_as(_up),
                      self.mass_is_int.ctypes.data_as(_bp), _d(self.radius), _d(self.volume),
                      self.alive.ctypes.data_as(_bp), self.immobile.ctypes.data_as(_bp),
                      self.eaten.ctypes.data_as(_ip))

    def frame(self, dt, collisions=True, record=False, cap=1 << 14):
        """One frame; returns dict(events=[(absorber, absorbed)], ...) with insertion indices."""
        n = self.n
        st = self._cstate()
        mass_pre = self.mass_f64.copy()
        nfl, nev = C.c_int64(0), C.c_int64(0)
        events = np.empty((cap, 2), dtype=np.int64)
        if record:
            act = np.empty(n, dtype=np.int64)
            acc = np.empty((n, 2)); pn = np.empty((n, 2)); vn = np.empty((n, 2))
            fp = np.empty((cap, 2), dtype=np.int64)
            na = lib().oracle_frame(C.byref(st), self.G, float(dt), int(collisions), int(self.bounded),
                                    self.r_max, act.ctypes.data_as(_ip), _d(acc), _d(pn), _d(vn),
                                    fp.ctypes.data_as(_ip), C.byref(nfl), events.ctypes.data_as(_ip),
                                    C.byref(nev), cap)
        else:
            na = lib().oracle_frame(C.byref(st), self.G, float(dt), int(collisions), int(self.bounded),
                                    self.r_max, None, None, None, None, None, C.byref(nfl),
                                    events.ctypes.data_as(_ip), C.byref(nev), cap)
        if nev.value > cap or nfl.value > cap:
            raise RuntimeError("oracle_frame: event/flag capacity exceeded")
        self.iterations += 1
        out = {"n_active": int(na), "events": events[:nev.value].copy(), "n_flags": int(nfl.value)}
        if record:
            out.update(active_idx=act[:na].copy(), acc=acc[:na].copy(), p_new=pn[:na].copy(),
                       v_new=vn[:na].copy(), flag_pairs=fp[:nfl.value].copy())
            if na:
                e, tot, ang = energies(pn[:na], vn[:na], mass_pre[act[:na]], self.G)
                out.update(energy=e, total_energy=tot, angular_momentum=ang)
        return out
