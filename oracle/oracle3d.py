"""ctypes front-end of oracle/state3d_oracle.c (the reference's new-API step, State.interact,
cosmosim/core/universe.py:94-129) -- TEST INFRASTRUCTURE ONLY, same rules as oracle/oracle.py."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as _o

_dp, _ip, _up, _bp = _o._dp, _o._ip, _o._up, _o._bp
_bound = False


class _State3(C.Structure):
    _fields_ = [("n", C.c_int64), ("pos", _dp), ("vel", _dp), ("mass_f64", _dp), ("mass_hi", _up),
                ("mass_lo", _up), ("mass_is_int", _bp), ("dens_f64", _dp), ("dens_is_int", _bp),
                ("id", _ip)]


def lib():
    global _bound
    L = _o.lib()
    if not _bound:
        L.oracle3_truediv.argtypes = [C.c_uint64] * 4
        L.oracle3_truediv.restype = C.c_double
        for name in ("oracle3_accel_rows", "oracle3_accel_rows_plain"):
            getattr(L, name).argtypes = [C.c_int64, _dp, _dp, C.c_double, C.c_int64, C.c_int64, _dp]
        L.oracle3_d2_rows.argtypes = [C.c_int64, _dp, C.c_int64, C.c_int64, _dp]
        L.oracle3_dist_rows.argtypes = [C.c_int64, _dp, C.c_int64, C.c_int64, _dp]
        L.oracle3_integrate.argtypes = [C.c_int64, _dp, _dp, _dp, C.c_double, _dp, _dp]
        L.oracle3_flags_rows.argtypes = [C.c_int64, _dp, _dp, C.c_int64, C.c_int64, _ip, C.c_int64]
        L.oracle3_flags_rows.restype = C.c_int64
        L.oracle3_radii.argtypes = [C.POINTER(_State3), _dp]
        L.oracle3_frame.argtypes = [C.POINTER(_State3), C.c_double, C.c_double, C.c_int, _dp, _dp, _dp, _dp,
                                    _ip, C.c_int64, _ip, _ip, C.c_int64]
        L.oracle3_frame.restype = C.c_int64
        _bound = True
    return L


_d, _c64 = _o._d, _o._c64


def truediv(a: int, b: int) -> float:
    m = (1 << 64) - 1
    return float(lib().oracle3_truediv(a >> 64, a & m, b >> 64, b & m))


def accel(pos, mass, G, i_begin=0, i_end=None, compensated=True):
    """acc_blas with (n,3) positions (blas.py:4-41, call site universe.py:101)."""
    pos, mass = _c64(pos), _c64(mass)
    n = mass.shape[0]
    i_end = n if i_end is None else i_end
    out = np.empty((i_end - i_begin, 3), dtype=np.float64)
    fn = lib().oracle3_accel_rows if compensated else lib().oracle3_accel_rows_plain
    fn(n, _d(pos), _d(mass), float(G), i_begin, i_end, _d(out))
    return out


def d2_matrix(pos, i_begin=0, i_end=None):
    pos = _c64(pos)
    n = pos.shape[0]
    i_end = n if i_end is None else i_end
    out = np.empty((i_end - i_begin, n))
    lib().oracle3_d2_rows(n, _d(pos), i_begin, i_end, _d(out))
    return out


def dist_matrix(p, i_begin=0, i_end=None):
    p = _c64(p)
    n = p.shape[0]
    i_end = n if i_end is None else i_end
    out = np.empty((i_end - i_begin, n))
    lib().oracle3_dist_rows(n, _d(p), i_begin, i_end, _d(out))
    return out


def integrate(p0, v0, a, dt):
    p0, v0, a = _c64(p0), _c64(v0), _c64(a)
    p, v = np.empty_like(p0), np.empty_like(v0)
    lib().oracle3_integrate(p0.shape[0], _d(p0), _d(v0), _d(a), float(dt), _d(p), _d(v))
    return p, v


def flags(p, radius, i_begin=0, i_end=None):
    p, radius = _c64(p), _c64(radius)
    n = radius.shape[0]
    i_end = n if i_end is None else i_end
    cap = 1 << 16
    while True:
        pairs = np.empty((cap, 2), dtype=np.int64)
        cnt = lib().oracle3_flags_rows(n, _d(p), _d(radius), i_begin, i_end, pairs.ctypes.data_as(_ip), cap)
        if cnt <= cap:
            return pairs[:cnt].copy()
        cap = int(cnt)


class OracleState:
    """SoA copy of a State (universe.py:73-79) stepped by oracle3_frame.

    snap: dict with pos (n,3), vel (n,3), mass_f64, mass_int_hi, mass_int_lo, mass_is_int,
    dens_f64, dens_is_int, id."""

    FIELDS = ("pos", "vel", "mass_f64", "mass_int_hi", "mass_int_lo", "mass_is_int", "dens_f64",
              "dens_is_int", "id")

    def __init__(self, snap, dt=1.0, G=6.674e-11):
        self.dt, self.G = float(dt), float(G)
        self.n = int(snap["mass_f64"].shape[0])
        self.pos = np.ascontiguousarray(snap["pos"], dtype=np.float64).copy()
        self.vel = np.ascontiguousarray(snap["vel"], dtype=np.float64).copy()
        self.mass_f64 = np.ascontiguousarray(snap["mass_f64"], dtype=np.float64).copy()
        self.mass_int_hi = np.ascontiguousarray(snap["mass_int_hi"], dtype=np.uint64).copy()
        self.mass_int_lo = np.ascontiguousarray(snap["mass_int_lo"], dtype=np.uint64).copy()
        self.mass_is_int = np.ascontiguousarray(snap["mass_is_int"]).astype(np.uint8)
        self.dens_f64 = np.ascontiguousarray(snap["dens_f64"], dtype=np.float64).copy()
        self.dens_is_int = np.ascontiguousarray(snap["dens_is_int"]).astype(np.uint8)
        self.id = np.ascontiguousarray(snap["id"], dtype=np.int64).copy()
        self.iteration = 0

    def snapshot(self):
        out = {k: getattr(self, k)[:self.n].copy() for k in self.FIELDS}
        for k in ("mass_is_int", "dens_is_int"):
            out[k] = out[k].astype(np.bool_)
        return out

    def _cstate(self):
        return _State3(self.n, _d(self.pos), _d(self.vel), _d(self.mass_f64),
                       self.mass_int_hi.ctypes.data_as(_up), self.mass_int_lo.ctypes.data_as(_up),
                       self.mass_is_int.ctypes.data_as(_bp), _d(self.dens_f64),
                       self.dens_is_int.ctypes.data_as(_bp), self.id.ctypes.data_as(_ip))

    def radii(self):
        r = np.empty(self.n)
        st = self._cstate()
        lib().oracle3_radii(C.byref(st), _d(r))
        return r

    def interact(self, collisions=True, record=False, cap=1 << 14):
        """One State.interact; returns dict(events=(k,3) [absorber id, absorbed id, existed], ...)."""
        n = self.n
        st = self._cstate()
        events = np.empty((cap, 4), dtype=np.int64)
        npairs = C.c_int64(0)
        if record:
            acc, pn, vn, rad = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3)), np.empty(n)
            pairs = np.empty((cap, 2), dtype=np.int64)
            nev = lib().oracle3_frame(C.byref(st), self.dt, self.G, int(bool(collisions)), _d(acc), _d(pn),
                                      _d(vn), _d(rad), pairs.ctypes.data_as(_ip), cap, C.byref(npairs),
                                      events.ctypes.data_as(_ip), cap)
        else:
            nev = lib().oracle3_frame(C.byref(st), self.dt, self.G, int(bool(collisions)), None, None, None,
                                      None, None, 0, C.byref(npairs), events.ctypes.data_as(_ip), cap)
        if nev < 0:
            raise OverflowError("oracle3_frame: an integer mass/density product exceeded 128 bits")
        if nev > cap or (record and npairs.value > cap):
            raise RuntimeError("oracle3_frame: capacity exceeded")
        self.n = int(st.n)
        self.iteration += 1
        out = {"events": events[:nev, :3].copy(), "n_pairs": int(npairs.value)}
        if record:
            out.update(acc=acc, p_new=pn, v_new=vn, radius=rad, flag_pairs=pairs[:npairs.value].copy())
        return out
