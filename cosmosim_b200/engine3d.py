"""Device-resident SoA object state and launch sequence of the 3-D step (the reference's new API,
``State.interact``, cosmosim/core/universe.py:94-129) -- binds include/cosmosim_b200_state3d.h.

As in engine.py, PyTorch only owns the device buffers; all arithmetic is in libcosmosim_b200.so.
The reference rebuilds numpy arrays from its Object list on every step (universe.py:95-98); here
the SoA buffers are the source of truth while stepping.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch

from . import _abi
from ._abi import (ERR3_INT_OVERFLOW, F3_DENS_INT, F3_MASS_INT, OPT3_COLLISIONS, OPT3_DEFER_INTEGRATE, OPT3_F32_SYM, OPT3_F64_SYM, OPT3_FP32, OPT3_STORE_ACC, PAD,
                   ST_ERROR, ST_N_ACTIVE, ST_N_DEAD, ST_N_EVENTS, ST_N_PAIRS, ST_STEP, STATUS_LEN)
from ._bound import BoundTracker
from .engine import _round_up, split_mass

MAX_INT_DENSITY = 1 << 53


def split_density(density):
    """Python density -> (float(density), is_int).  Integer densities are held as exact doubles."""
    if isinstance(density, (int, np.integer)) and not isinstance(density, (bool, np.bool_)):
        d = int(density)
        return float(d), 0 <= d < MAX_INT_DENSITY
    return float(density), False


class Engine3D(BoundTracker):
    JSPLIT_MAX = 128

    def __init__(self, n_total, device=None, pair_cap=None, event_cap=None, group=None):
        if not torch.cuda.is_available():
            raise _abi.CosmoError("cosmosim_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _abi.lib()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
        with torch.cuda.device(self.device):
            _abi.check(self.lib.cosmo_device_info(C.byref(sm), C.byref(maj), C.byref(mnr)), "cosmo_device_info")
            # load the post-scheduled kernels now (csrc/sym_tuned.cu): never inside a CUDA-graph capture
            self.tuned_kernels = int(self.lib.cosmo_sym_tuned_available())
        self.n_total = int(n_total)
        self.cap = max(PAD, _round_up(self.n_total, PAD))
        self.group = group
        self.rank, self.world = 0, 1
        if group is not None:
            import torch.distributed as dist
            from . import sharded
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
            self.cap = sharded.required_capacity(self.n_total, self.world)
        self.pair_cap = int(pair_cap or max(1 << 16, 4 * self.cap))
        self.event_cap = int(event_cap or max(1 << 16, 4 * self.n_total))
        self._alloc()
        self._init_bound()           # status snapshots / launch bound / grid control (_bound.py)
        self.use_graphs = True       # small N: replay a captured CUDA graph of the step (launch-bound regime)
        self.graph_threshold = 16384
        self._graphs = {}
        self.pos_exp = 0
        self.mass_exp = 0
        self.steps = 0
        self._jsplit_cache = {}
        self.grid_threshold = 16384  # objects from which the grid broad phase beats the all-pairs test
        self.sym64_threshold = 16384 # FP64: from here on the symmetric kernel (each pair once) wins
        self.sym32_threshold = 32768 # FP32: same for the symmetric FP32 kernel
        self.sym_chunk = 0           # column blocks per CTA of the symmetric kernel (0 = automatic)
        self.near_field = True       # FP32 flavours: FP64 near-field corrections on
        self._grid_ready = False     # a stage B in grid mode has left a tuned grid on the device
        self.resolve_mode = None     # stage C: None = automatic, 0 = sequential loop, 1 = parallel components
        self.resolve_parallel_from = 512   # flagged pairs per step from which the parallel stage C pays off
        self._colbuf = None

    def _alloc(self):
        dev, cap = self.device, self.cap
        f64, i64, i32, u8, f32 = torch.float64, torch.int64, torch.int32, torch.uint8, torch.float32
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        self.t = t = {}
        t["pos"] = z(3, cap, dtype=f64); t["vel"] = z(3, cap, dtype=f64)
        t["mass"] = z(cap, dtype=f64); t["mass_hi"] = z(cap, dtype=i64); t["mass_lo"] = z(cap, dtype=i64)
        t["density"] = z(cap, dtype=f64); t["radius"] = z(cap, dtype=f64)
        t["id"] = z(cap, dtype=i32); t["flags"] = z(cap, dtype=u8)
        t["status"] = z(STATUS_LEN, dtype=i64)
        t["events"] = z(self.event_cap, 4, dtype=i64)
        t["h"] = z(cap, dtype=f64); t["pk"] = z(4, cap, dtype=f32); t["acc"] = z(3, cap, dtype=f64)
        t["pos_new"] = z(3, cap, dtype=f64); t["vel_new"] = z(3, cap, dtype=f64); t["xx_new"] = z(cap, dtype=f64)
        t["rinf"] = z(cap, dtype=f64)
        t["partial"] = z(self.JSPLIT_MAX, 3, cap, dtype=f64)
        t["tile_ctr"] = z(cap // 64 + 4, dtype=i32)
        t["pairs"] = z(self.pair_cap, dtype=i64); t["pairs_sorted"] = z(self.pair_cap, dtype=i64)
        t["exists"] = z(cap, dtype=u8)
        t["cur_pos"] = z(3, cap, dtype=f64); t["cur_vel"] = z(3, cap, dtype=f64)
        t["touched"] = torch.full((cap,), -1, dtype=i32, device=dev)
        t["scan_in"] = z(cap + 1, dtype=i32); t["scan_out"] = z(cap + 1, dtype=i32)
        t["scan_blk"] = z(_abi.SCAN_BLOCKS + 1, dtype=i32)
        t["t_pos"] = z(3, cap, dtype=f64); t["t_vel"] = z(3, cap, dtype=f64)
        for k, dt_ in (("mass", f64), ("mass_hi", i64), ("mass_lo", i64), ("density", f64), ("radius", f64),
                       ("id", i32), ("flags", u8)):
            t["t_" + k] = z(cap, dtype=dt_)
        self.grid_cells = int(min(_abi.SCAN_BLOCKS * 4096 - 1, max(1 << 16, 8 * cap)))
        self.giant_cap = 1 << 16
        t["cell_of"] = z(cap, dtype=i32); t["cell_items"] = z(cap, dtype=i32)
        t["cell_count"] = z(self.grid_cells + 1, dtype=i32); t["cell_start"] = z(self.grid_cells + 1, dtype=i32)
        t["giants"] = z(self.giant_cap, dtype=i32)
        t["chunk_bounds"] = z(cap // 1024 + 2, dtype=i32); t["heavy_rows"] = z(cap, dtype=i32)
        t["grid_dev"] = z(256, dtype=u8); t["tune_hist"] = z(_abi.TUNE_HIST_LEN, dtype=i32)
        # parallel stage C (crowded steps)
        t["row_count"] = z(cap + 1, dtype=i32); t["row_start"] = z(cap + 1, dtype=i32); t["label"] = z(cap, dtype=i32)
        t["comp_root"] = z(self.pair_cap, dtype=i32); t["comp_items"] = z(self.pair_cap, dtype=i32)
        t["pair_flag"] = z(self.pair_cap, dtype=i32); t["scan_a"] = z(self.pair_cap + 1, dtype=i32)
        t["corr"] = z(3, cap, dtype=f64)                       # FP64 near-field corrections of the FP32 flavours
        p = lambda k: C.c_void_p(t[k].data_ptr())  # noqa: E731
        st = _abi.State3()
        st.cap, st.n_total, st.event_cap = cap, self.n_total, self.event_cap
        for k in ("pos", "vel", "mass", "mass_hi", "mass_lo", "density", "radius", "id", "flags", "status", "events"):
            setattr(st, k, p(k))
        sc = _abi.Scratch3()
        for k in ("h", "pk", "acc", "pos_new", "vel_new", "xx_new", "rinf", "partial", "tile_ctr", "pairs",
                  "pairs_sorted", "exists", "cur_pos", "cur_vel", "touched", "scan_in", "scan_out", "scan_blk",
                  "t_pos", "t_vel", "t_mass", "t_mass_hi", "t_mass_lo", "t_density", "t_radius", "t_id", "t_flags",
                  "cell_of", "cell_count", "cell_start", "cell_items", "giants", "chunk_bounds", "heavy_rows",
                  "grid_dev", "tune_hist", "row_count", "row_start", "label", "comp_root", "comp_items", "pair_flag",
                  "scan_a", "corr"):
            setattr(sc, k, p(k))
        sc.jsplit_max, sc.pair_cap = self.JSPLIT_MAX, self.pair_cap
        sc.grid_cells, sc.giant_cap = self.grid_cells, self.giant_cap
        self.st, self.sc = st, sc

    # ------------------------------------------------------------------ host <-> device
    def upload(self, arrays, iteration=0):
        """arrays: pos[n,3], vel[n,3], mass_f64[n], mass_hi[n], mass_lo[n] (uint64), density[n], radius[n]
        (Object.get_radius() evaluated by the caller), flags[n] (F3_* bits), id[n]."""
        n = int(arrays["mass_f64"].shape[0])
        if n > self.cap:
            raise ValueError("upload of %d objects exceeds capacity %d" % (n, self.cap))
        t = self.t

        def put(key, src, dtype, planes=False):
            a = np.ascontiguousarray(src, dtype=dtype)
            if planes:
                a = np.ascontiguousarray(a.reshape(n, 3).T)
            if a.dtype == np.uint64:
                a = a.view(np.int64)
            h = torch.from_numpy(a)
            if planes:
                t[key][:, :n].copy_(h)
            else:
                t[key][:n].copy_(h)

        put("pos", arrays["pos"], np.float64, True); put("vel", arrays["vel"], np.float64, True)
        put("mass", arrays["mass_f64"], np.float64)
        put("mass_hi", arrays["mass_hi"], np.uint64); put("mass_lo", arrays["mass_lo"], np.uint64)
        put("density", arrays["density"], np.float64); put("radius", arrays["radius"], np.float64)
        put("id", arrays["id"], np.int32); put("flags", arrays["flags"], np.uint8)
        status = torch.zeros(STATUS_LEN, dtype=torch.int64)
        status[ST_N_ACTIVE] = n
        status[ST_STEP] = int(iteration)
        status[ST_N_EVENTS] = self.t["status"].cpu()[ST_N_EVENTS]
        t["status"].copy_(status)
        t["touched"].fill_(-1)
        self.n_upper = n
        self.steps = int(iteration)
        torch.cuda.current_stream(self.device).synchronize()
        self._status_host.copy_(status)
        self._seed_snapshot(status)
        pos = np.asarray(arrays["pos"], dtype=np.float64)
        mass = np.asarray(arrays["mass_f64"], dtype=np.float64)
        pmax = float(np.max(np.abs(pos))) if pos.size else 1.0
        mmax = float(np.max(np.abs(mass))) if mass.size else 1.0
        self.pos_exp = (math.frexp(pmax)[1] - 6) if pmax > 0 and math.isfinite(pmax) else 0
        self.mass_exp = (math.frexp(mmax)[1] - 20) if mmax > 0 and math.isfinite(mmax) else 0

    def _tick(self):
        return self.steps

    def status(self):
        s = self.t["status"].cpu().numpy()
        self.n_upper = int(s[ST_N_ACTIVE])
        return {"n_active": int(s[ST_N_ACTIVE]), "iteration": int(s[ST_STEP]), "n_events": int(s[ST_N_EVENTS]),
                "n_pairs": int(s[ST_N_PAIRS]), "n_dead": int(s[ST_N_DEAD]), "error": int(s[ST_ERROR])}

    def check_errors(self):
        st = self.status()
        err = st["error"]
        if err:
            what = [name for bit, name in ((1, "flagged-pair buffer overflow"), (2, "event log overflow"),
                                           (4, "non-finite acceleration (d^2 < 0 or NaN input)"),
                                           (8, "broad-phase giant list overflow"),
                                           (ERR3_INT_OVERFLOW, "integer mass x density beyond 128 bits")) if err & bit]
            raise _abi.CosmoError("device reported: " + ", ".join(what or ["error bits 0x%x" % err]))
        return st

    def download(self):
        """Objects of State.objects in list order (numpy, positions/velocities as [n,3])."""
        st = self.status()
        n = st["n_active"]
        t = self.t
        out = {k: t[k][:n].cpu().numpy() for k in ("mass", "density", "radius", "id", "flags")}
        out["pos"] = np.ascontiguousarray(t["pos"][:, :n].cpu().numpy().T)
        out["vel"] = np.ascontiguousarray(t["vel"][:, :n].cpu().numpy().T)
        out["mass_hi"] = t["mass_hi"][:n].cpu().numpy().view(np.uint64)
        out["mass_lo"] = t["mass_lo"][:n].cpu().numpy().view(np.uint64)
        out["status"] = st
        return out

    def events(self, start=0):
        """Object.absorb call stream rows (iteration, absorber id, absorbed id, absorbed existed)."""
        n = min(self.status()["n_events"], self.event_cap)
        return self.t["events"][start:n].cpu().numpy()

    def last_step_tensors(self, n):
        """acc / pos_new / vel_new ([n,3] numpy) and the flagged pairs of the last step, in the slot
        order the step STARTED with."""
        t = self.t
        out = {k: np.ascontiguousarray(t[k][:, :n].cpu().numpy().T) for k in ("acc", "pos_new", "vel_new")}
        m = min(int(self.t["status"].cpu()[ST_N_PAIRS]), self.pair_cap)
        keys = np.sort(t["pairs"][:m].cpu().numpy().view(np.uint64))
        out["flag_pairs"] = np.stack([(keys >> np.uint64(32)).astype(np.int64),
                                      (keys & np.uint64(0xffffffff)).astype(np.int64)], axis=1).reshape(-1, 2)
        return out

    # ------------------------------------------------------------------ stepping
    def _ensure_colbuf(self, n, fp32=False):
        """Column-partial buffer of the symmetric kernels.  FP64: three planes of `cap` doubles per
        512-object row block (24 B x N^2 / 512: 201 MB at N = 65536); FP32: three planes of floats per
        1024-object row block (12 B x N^2 / 1024: 0.8 GB at N = 262144).  Rows are counted in units of
        three planes of doubles (a float row takes half of one)."""
        rows_local = -(-(-(-n // (1024 if fp32 else 512))) // self.world)   # row blocks are dealt round-robin to the ranks
        rows = -(-rows_local // 2) if fp32 else rows_local
        if self._colbuf is None or self._colbuf.shape[0] < rows:
            self._colbuf = torch.empty(rows, 3, self.cap, dtype=torch.float64, device=self.device)
            self.sc.colbuf = C.c_void_p(self._colbuf.data_ptr())
            self.sc.colbuf_rows = rows

    def params(self, dt, G, collisions=True, fp32=False, store_acc=False, i_begin=0, i_end=None, sym=None, near=None):
        n = self.n_upper
        p = _abi.StepParams3()
        p.G, p.dt = float(G), float(dt)
        p.opts = (OPT3_COLLISIONS if collisions else 0) | (OPT3_FP32 if fp32 else 0) | (OPT3_STORE_ACC if store_acc else 0)
        if self.world > 1:
            from . import sharded
            i_begin, i_end, _ = sharded.shard_rows(n, self.world, self.rank)
            p.shard_index, p.shard_count = self.rank, self.world
        else:
            p.shard_index, p.shard_count = 0, 1
        whole = self.world > 1 or (int(i_begin) == 0 and (i_end is None or int(i_end) >= n))
        if sym is None:
            sym = n >= (self.sym32_threshold if fp32 else self.sym64_threshold)
        if sym and whole and n > (1024 if fp32 else 512):
            self._ensure_colbuf(n, fp32)
            p.opts |= (OPT3_F32_SYM if fp32 else OPT3_F64_SYM) | (OPT3_DEFER_INTEGRATE if self.world > 1 else 0)
            nblk = -(-n // (1024 if fp32 else 512))
            want = (nblk * nblk) // (2 * 64 * 2 * 148 * self.world)     # >= ~64 block pairs per resident CTA slot
            p.sym_chunk = int(self.sym_chunk or min(32, max(1, -(-nblk // self.JSPLIT_MAX), want)))
            if os.environ.get("COSMO3_SYM_CHUNK"):
                p.sym_chunk = max(int(os.environ["COSMO3_SYM_CHUNK"]), -(-nblk // self.JSPLIT_MAX))
        p.n_upper, p.i_begin, p.i_end = n, int(i_begin), int(n if i_end is None else i_end)
        p.pos_exp, p.mass_exp = self.pos_exp, self.mass_exp
        rows = p.i_end - p.i_begin
        key = (rows, n, bool(fp32))
        js = self._jsplit_cache.get(key)
        if js is None:
            if len(self._jsplit_cache) > 4096:
                self._jsplit_cache.clear()
            js = (self.lib.cosmo3_suggest_jsplit(rows, n, 1 if fp32 else 0, self.JSPLIT_MAX),
                  self.lib.cosmo3_suggest_jsplit(rows, n, 2, 256))
            self._jsplit_cache[key] = js
        p.jsplit, p.jsplit_detect = js
        if collisions and not self._force_all_pairs and n >= self.grid_threshold:
            # grid tuned on the device from the step's own positions and radii (detect_mode 2) unless
            # the caller pinned one (detect_mode 1)
            if self._host_grid is not None:
                p.detect_mode = 1
                p.cell_size, p.grid_origin_x, p.grid_origin_y, p.cells_per_side, p.giant_radius = self._host_grid
            else:
                p.detect_mode = 2
        # stage C: the sequential loop is the cheaper one for a handful of flagged pairs; crowded steps
        # (judged by the pair count of the status snapshot BOUND_LAG steps back) take the parallel one
        # FP32 flavours: pairs closer than 2^14 FP32 ulps of their coordinates are re-evaluated in FP64
        # (csrc/nearfield.cu), binned with the grid the previous step's stage B left on the device
        if fp32 and whole and (self.near_field if near is None else near):
            p.near_mode = 1 if self._grid_ready else 2
        mode = self.resolve_mode
        p.resolve_mode = int(mode if mode is not None else self.pairs_lag > self.resolve_parallel_from)
        return p

    def _uses_grid(self, collisions=True):
        return bool(collisions and not self._force_all_pairs and self.n_upper >= self.grid_threshold)

    def _refresh_graph_bound(self, margin):
        """CUDA-graph replay: non-blocking refresh of the launch bound from the newest completed
        replay's snapshot (the graph copies the status block into _status_host)."""
        seen_n, seen_s = int(self._status_host[ST_N_ACTIVE]), int(self._status_host[ST_STEP])
        if self._status_epoch <= seen_s <= self.steps and 0 < seen_n < margin * self.n_upper:
            self.n_upper = seen_n

    def _step_graphed(self, steps, dt, G, collisions, fp32, sym):
        """Launch-bound regime (a step is ~13 kernels of a few microseconds): capture cosmo3_step + the
        status snapshot once into a CUDA graph and replay it.  The kernels read the object count from
        the device, so a graph stays valid while objects merge; it is re-captured when the launch
        bound has shrunk by more than 10 %."""
        done = 0
        while done < steps:
            self._refresh_graph_bound(0.9)
            p = self.params(dt, G, collisions, fp32, False, sym=sym)
            key = bytes(memoryview(p))
            g = self._graphs.get(key)
            if g is None:
                if len(self._graphs) > 16:
                    self._graphs.clear()
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    sp = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
                    _abi.check(self.lib.cosmo3_step(C.byref(self.st), C.byref(self.sc), C.byref(p), sp),
                               "cosmo3_step (graph capture)")
                    self._status_host.copy_(self.t["status"], non_blocking=True)
                self._graphs[key] = g
            burst = min(steps - done, 64)
            for _ in range(burst):
                g.replay()
            done += burst
            self.steps += burst

    def step(self, dt, G, collisions=True, fp32=False, store_acc=False, steps=1, stage_events=None, sym=None, near=None):
        """`steps` x State.interact on this device.  The host never reads object data here; the only
        wait is on the status snapshot of BOUND_LAG steps ago (_bound.py).  stage_events: a list
        that receives, per step, four CUDA events bracketing stages A, B and C (bench.py)."""
        stream = torch.cuda.current_stream(self.device)
        sp = C.c_void_p(stream.cuda_stream)
        L, st, sc = self.lib, self.st, self.sc
        with torch.cuda.device(self.device):
            if (self.use_graphs and self.world == 1 and stage_events is None and not store_acc and int(steps) >= 8
                    and self.n_upper <= self.graph_threshold and not self._uses_grid(collisions)):
                return self._step_graphed(int(steps), dt, G, collisions, fp32, sym)
            for _ in range(int(steps)):
                self._refresh_bound()
                p = self.params(dt, G, collisions, fp32, store_acc, sym=sym, near=near)
                if self.world > 1:
                    from . import sharded
                    sharded.frame3d(self, p, sp, stage_events)
                elif stage_events is None:
                    _abi.check(L.cosmo3_step(C.byref(st), C.byref(sc), C.byref(p), sp), "cosmo3_step")
                else:
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                    ev[0].record(stream)
                    _abi.check(L.cosmo3_accel_integrate(C.byref(st), C.byref(sc), C.byref(p), sp), "cosmo3_accel_integrate")
                    ev[1].record(stream)
                    _abi.check(L.cosmo3_detect_pairs(C.byref(st), C.byref(sc), C.byref(p), sp), "cosmo3_detect_pairs")
                    ev[2].record(stream)
                    _abi.check(L.cosmo3_resolve_commit(C.byref(st), C.byref(sc), C.byref(p), sp), "cosmo3_resolve_commit")
                    ev[3].record(stream)
                    stage_events.append(ev)
                self.steps += 1
                self._push_snapshot()
                if p.detect_mode != 0:
                    self._grid_ready = True


    # ------------------------------------------------------------------ host-buffer path (e2e)
    E2E_IN = ("pos", "vel", "mass", "mass_hi", "mass_lo", "density", "radius", "id", "flags")
    E2E_OUT = ("pos", "vel", "mass", "mass_hi", "mass_lo", "density", "radius", "id", "flags")

    def make_host_staging(self, arrays):
        """Pinned host mirrors of the state (what a host-resident caller hands over per step)."""
        n = int(arrays["mass_f64"].shape[0])
        src = {"pos": np.ascontiguousarray(np.asarray(arrays["pos"], dtype=np.float64).T),
               "vel": np.ascontiguousarray(np.asarray(arrays["vel"], dtype=np.float64).T),
               "mass": arrays["mass_f64"], "mass_hi": np.asarray(arrays["mass_hi"]).view(np.int64),
               "mass_lo": np.asarray(arrays["mass_lo"]).view(np.int64), "density": arrays["density"],
               "radius": arrays["radius"], "id": arrays["id"], "flags": arrays["flags"]}
        stg = {"n": n, "in": {}, "out": {}}
        for k in self.E2E_IN:
            stg["in"][k] = torch.from_numpy(np.ascontiguousarray(src[k])).pin_memory()
        for k in self.E2E_OUT:
            shape = (3, n) if k in ("pos", "vel") else (n,)
            stg["out"][k] = torch.empty(shape, dtype=self.t[k].dtype).pin_memory()
        stg["status_in"] = torch.zeros(STATUS_LEN, dtype=torch.int64).pin_memory()
        stg["status_out"] = torch.zeros(STATUS_LEN, dtype=torch.int64).pin_memory()
        stg["h2d_bytes"] = sum(v.numel() * v.element_size() for v in stg["in"].values()) + STATUS_LEN * 8
        stg["d2h_bytes"] = sum(v.numel() * v.element_size() for v in stg["out"].values()) + STATUS_LEN * 8
        return stg

    def step_from_host(self, stg, dt, G, **kw):
        """One State.interact the way a host-resident caller sees it: state H2D from pinned memory, the
        step, the new state D2H, stream synchronised."""
        n, t = stg["n"], self.t
        for k in self.E2E_IN:
            dst = t[k][:, :n] if k in ("pos", "vel") else t[k][:n]
            dst.copy_(stg["in"][k], non_blocking=True)
        stg["status_in"].zero_()
        stg["status_in"][ST_N_ACTIVE] = n
        stg["status_in"][ST_STEP] = self.steps
        t["status"].copy_(stg["status_in"], non_blocking=True)
        self.n_upper = n
        self._seed_snapshot(stg["status_in"])     # snapshots of earlier steps are void
        self.step(dt, G, **kw)
        for k in self.E2E_OUT:
            src = t[k][:, :n] if k in ("pos", "vel") else t[k][:n]
            stg["out"][k].copy_(src, non_blocking=True)
        stg["status_out"].copy_(t["status"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return stg["status_out"]


def arrays_from_objects(objects):
    """Host Object list (reference attribute names: mass, density, position, velocity) -> upload arrays.
    The radius is Object.get_radius() evaluated by Python itself, exactly as the reference does."""
    n = len(objects)
    out = {"pos": np.zeros((n, 3)), "vel": np.zeros((n, 3)), "mass_f64": np.zeros(n),
           "mass_hi": np.zeros(n, dtype=np.uint64), "mass_lo": np.zeros(n, dtype=np.uint64),
           "density": np.zeros(n), "radius": np.zeros(n), "flags": np.zeros(n, dtype=np.uint8),
           "id": np.arange(n, dtype=np.int32)}
    for k, o in enumerate(objects):
        out["pos"][k] = o.position
        out["vel"][k] = o.velocity
        f, is_int, hi, lo = split_mass(o.mass)
        d, d_int = split_density(o.density)
        out["mass_f64"][k], out["mass_hi"][k], out["mass_lo"][k] = f, hi, lo
        out["density"][k] = d
        out["flags"][k] = (F3_MASS_INT if is_int else 0) | (F3_DENS_INT if d_int else 0)
        out["radius"][k] = ((3 * (o.mass / o.density)) / (4 * math.pi)) ** (1 / 3)
    return out


def arrays_from_columns(mass, density, position, velocity=None):
    """Column data -> upload arrays, without a Python Object per body.  `mass` / `density` may be numpy
    float arrays or sequences of Python numbers (ints stay exact ints, as with Object)."""
    pos = np.ascontiguousarray(np.asarray(position, dtype=np.float64).reshape(-1, 3))
    n = pos.shape[0]
    vel = np.zeros((n, 3)) if velocity is None else np.ascontiguousarray(np.asarray(velocity, dtype=np.float64).reshape(n, 3))
    masses = mass.tolist() if isinstance(mass, np.ndarray) else list(mass)
    dens = density.tolist() if isinstance(density, np.ndarray) else (list(density) if np.ndim(density) else [density] * n)
    if len(masses) != n or len(dens) != n:
        raise ValueError("mass / density / position lengths differ")
    out = {"pos": pos, "vel": vel, "mass_f64": np.zeros(n), "mass_hi": np.zeros(n, dtype=np.uint64),
           "mass_lo": np.zeros(n, dtype=np.uint64), "density": np.zeros(n), "radius": np.zeros(n),
           "flags": np.zeros(n, dtype=np.uint8), "id": np.arange(n, dtype=np.int32)}
    four_pi = 4 * math.pi
    for k in range(n):
        m, d = masses[k], dens[k]
        f, is_int, hi, lo = split_mass(m)
        df, d_int = split_density(d)
        out["mass_f64"][k], out["mass_hi"][k], out["mass_lo"][k], out["density"][k] = f, hi, lo, df
        out["flags"][k] = (F3_MASS_INT if is_int else 0) | (F3_DENS_INT if d_int else 0)
        out["radius"][k] = ((3 * (m / d)) / four_pi) ** (1 / 3)      # Object.get_radius() in Python arithmetic
    return out
