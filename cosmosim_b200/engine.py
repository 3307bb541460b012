"""Device-resident SoA planet state and the per-frame launch sequence.

PyTorch is used for exactly two things here: owning the device buffers (torch tensors whose
``data_ptr()`` goes through the C ABI) and, in the sharded mode, ``torch.distributed``
collectives.  All arithmetic happens in libcosmosim_b200.so.

Replaces the body of ``Universe.interact`` (reference cosmosim/core/universe_old.py:166-212):
the reference rebuilds numpy arrays from its Planet list every frame (:170-172,192); here the
SoA buffers are the source of truth while stepping and the Planet objects are refreshed
lazily.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch

from . import _abi
from ._bound import BoundTracker
from ._abi import (OPT_BOUNDED, OPT_COLLISIONS, OPT_DEFER_INTEGRATE, OPT_ENERGY, OPT_F64_SYM,
                   OPT_F32_SYM, OPT_FP32, OPT_STORE_ACC, PAD, ST_ERROR, ST_N_ACTIVE, ST_N_DEAD, ST_N_ESCAPED, ST_N_EVENTS,
                   ST_N_PAIRS, ST_STEP, STATUS_LEN)

_MASK64 = (1 << 64) - 1
MAX_INT_MASS = 1 << 127


def split_mass(mass):
    """Python mass -> (float(mass), is_int, hi, lo) following Python numeric semantics."""
    if isinstance(mass, (int, np.integer)) and not isinstance(mass, (bool, np.bool_)):
        m = int(mass)
        if 0 <= m < MAX_INT_MASS:
            return float(m), True, (m >> 64) & _MASK64, m & _MASK64
        return float(m), False, 0, 0
    return float(mass), False, 0, 0


def _round_up(n, q):
    return (int(n) + q - 1) // q * q


class Engine(BoundTracker):
    """One GPU's replica of the world (optionally responsible for a shard of target rows)."""

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
        self.sm_count = sm.value
        self.n_total = int(n_total)
        self.cap = max(PAD, _round_up(self.n_total, PAD))
        self.pair_cap = int(pair_cap or max(1 << 16, 4 * self.cap))
        self.event_cap = int(event_cap or max(1 << 16, 2 * self.n_total))
        self.group = group
        self.rank, self.world = 0, 1
        if group is not None:
            import torch.distributed as dist
            from . import sharded
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
            self.cap = sharded.required_capacity(self.n_total, self.world)
        self._alloc()
        self._init_bound()          # status snapshots / launch bound / grid control (_bound.py)
        self.pos_exp = 0
        self.mass_exp = 0
        self.frames = 0
        self.grid_threshold = 8192  # planets from which the grid broad phase + parallel resolver pay off
        self.small_threshold = 4096  # below: small-N kernel shape (more, shorter CTAs)
        self.sym_threshold = 32768   # FP32: from here on the symmetric kernel (each pair once) wins
        self.sym64_threshold = 16384  # FP64: same for the symmetric FP64 kernel
        self.sym_chunk = 0           # column blocks per CTA of the symmetric kernel (0 = automatic)
        self.sym_rows = 0            # row blocks per CTA of the symmetric FP32 kernel (0 = automatic)
        self.use_graphs = True      # small N: replay a captured CUDA graph of the frame (launch-bound regime)
        self.graph_threshold = 16384
        self._graphs = {}
        self._params_cache = {}
        self.near_field = True      # FP32 path: FP64 near-field corrections on
        self._grid_ready = False    # a stage B in grid mode has left a tuned grid on the device

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        dev, cap, nt = self.device, self.cap, max(1, self.n_total)
        f64, i64, i32, u8, f32 = torch.float64, torch.int64, torch.int32, torch.uint8, torch.float32
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        self.t = t = {}
        # persistent state
        t["pos"] = z(cap, 2, dtype=f64); t["vel"] = z(cap, 2, dtype=f64)
        t["mass"] = z(cap, dtype=f64); t["mass_hi"] = z(cap, dtype=i64); t["mass_lo"] = z(cap, dtype=i64)
        t["radius"] = z(cap, dtype=f64); t["volume"] = z(cap, dtype=f64)
        t["eaten"] = z(cap, dtype=i64); t["id"] = z(cap, dtype=i32); t["flags"] = z(cap, dtype=u8)
        t["status"] = z(STATUS_LEN, dtype=i64)
        t["energy"] = z(cap, dtype=f64); t["totals"] = z(4, dtype=f64); t["t_energy"] = z(cap, dtype=f64)
        t["rec_mass"] = z(nt, dtype=f64); t["rec_mass_hi"] = z(nt, dtype=i64); t["rec_mass_lo"] = z(nt, dtype=i64)
        t["rec_radius"] = z(nt, dtype=f64); t["rec_volume"] = z(nt, dtype=f64); t["rec_eaten"] = z(nt, dtype=i64)
        t["rec_death"] = torch.full((nt,), -1, dtype=i32, device=dev); t["rec_flags"] = z(nt, dtype=u8)
        t["events"] = z(self.event_cap, 3, dtype=i64)
        # scratch
        t["h"] = z(cap, dtype=f64)
        t["pk_x"] = z(cap, dtype=f32); t["pk_y"] = z(cap, dtype=f32); t["pk_m"] = z(cap, dtype=f32)
        t["acc"] = z(cap, 2, dtype=f64)
        # pos_new | vel_new | xx_new live in one allocation so the sharded mode gathers once
        t["new"] = z(cap, 5, dtype=f64)
        t["rinf"] = z(cap, dtype=f64)
        t["partial"] = z(self.JSPLIT_MAX, cap, 2, dtype=f64)
        t["tile_ctr"] = z(cap // 64 + 4, dtype=i32)
        t["alive"] = z(cap, dtype=u8); t["escaped"] = z(cap, dtype=u8)
        t["pairs"] = z(self.pair_cap, dtype=i64); t["pairs_sorted"] = z(self.pair_cap, dtype=i64)
        t["cur_pos"] = z(cap, 2, dtype=f64); t["cur_vel"] = z(cap, 2, dtype=f64)
        t["mod_row"] = z(cap, dtype=i32); t["mod_step"] = torch.full((cap,), -1, dtype=i32, device=dev)
        t["chunk_alive"] = z(cap // 1024 + 2, dtype=i32)
        # CSR of the flagged pairs + parallel resolver
        t["row_count"] = z(cap + 1, dtype=i32); t["row_start"] = z(cap + 1, dtype=i32)
        t["pmin"] = torch.full((cap,), 0x7fffffff, dtype=i32, device=dev)
        t["pmax"] = torch.full((cap,), -1, dtype=i32, device=dev)
        t["scan_blk"] = z(_abi.SCAN_BLOCKS + 1, dtype=i32)
        t["pair_flag"] = z(self.pair_cap, dtype=i32); t["cplx_list"] = z(self.pair_cap, dtype=i32)
        t["scan_a"] = z(self.pair_cap + 1, dtype=i32)
        t["ev_tmp"] = z(self.pair_cap, 2, dtype=i64)
        t["label"] = z(cap, dtype=i32); t["comp_items"] = z(self.pair_cap, dtype=i32)
        t["comp_root"] = z(self.pair_cap, dtype=i32)
        # uniform-grid broad phase
        self.grid_cells = int(min(_abi.SCAN_BLOCKS * 4096 - 1, max(1 << 16, 8 * cap)))
        self.giant_cap = 1 << 16
        t["cell_of"] = z(cap, dtype=i32); t["cell_items"] = z(cap, dtype=i32)
        t["cell_count"] = z(self.grid_cells + 1, dtype=i32); t["cell_start"] = z(self.grid_cells + 1, dtype=i32)
        t["giants"] = z(self.giant_cap, dtype=i32)
        t["grid_dev"] = z(256, dtype=u8)                       # cosmo_grid[2]: stage B, near field
        t["corr"] = z(cap, 2, dtype=f64)
        t["tune_hist"] = z(_abi.TUNE_HIST_LEN, dtype=i32)
        for k, dt_ in (("pos", f64), ("vel", f64)):
            t["t_" + k] = z(cap, 2, dtype=dt_)
        for k, dt_ in (("mass", f64), ("mass_hi", i64), ("mass_lo", i64), ("radius", f64), ("volume", f64),
                       ("eaten", i64), ("id", i32), ("flags", u8)):
            t["t_" + k] = z(cap, dtype=dt_)
        # the three per-frame outputs of stage A are SoA planes of t["new"]
        new = t["new"].view(-1)
        self._pos_new = new[: 2 * cap]
        self._vel_new = new[2 * cap: 4 * cap]
        self._xx_new = new[4 * cap: 5 * cap]

        p = lambda k: C.c_void_p(t[k].data_ptr())  # noqa: E731
        st = _abi.State()
        st.cap, st.n_total = cap, self.n_total
        for k in ("pos", "vel", "mass", "mass_hi", "mass_lo", "radius", "volume", "eaten", "id", "flags",
                  "status", "energy", "totals", "rec_mass", "rec_mass_hi", "rec_mass_lo", "rec_radius", "rec_volume",
                  "rec_eaten", "rec_death", "rec_flags", "events"):
            setattr(st, k, p(k))
        st.event_cap = self.event_cap
        sc = _abi.Scratch()
        for k in ("h", "pk_x", "pk_y", "pk_m", "acc", "rinf", "partial", "tile_ctr", "alive", "escaped",
                  "pairs", "pairs_sorted", "cur_pos", "cur_vel", "mod_row", "mod_step", "chunk_alive",
                  "t_pos", "t_vel", "t_mass", "t_mass_hi", "t_mass_lo", "t_radius", "t_volume",
                  "t_eaten", "t_id", "t_flags", "t_energy", "row_count", "row_start", "pmin", "pmax", "scan_blk",
                  "pair_flag", "cplx_list", "scan_a", "ev_tmp", "label", "comp_items", "comp_root", "cell_of", "cell_items", "cell_count", "cell_start",
                  "giants", "grid_dev", "tune_hist", "corr"):
            setattr(sc, k, p(k))
        sc.pos_new = C.c_void_p(self._pos_new.data_ptr())
        sc.vel_new = C.c_void_p(self._vel_new.data_ptr())
        sc.xx_new = C.c_void_p(self._xx_new.data_ptr())
        sc.jsplit_max = self.JSPLIT_MAX
        sc.pair_cap = self.pair_cap
        sc.grid_cells = self.grid_cells
        sc.giant_cap = self.giant_cap
        self.st, self.sc = st, sc

    # ------------------------------------------------------------------ host <-> device
    def upload(self, arrays, frame=0):
        """Load the ACTIVE planets (insertion order) from host numpy arrays.

        arrays: pos[n,2], vel[n,2], mass_f64[n], mass_hi[n], mass_lo[n] (uint64), flags[n] (uint8),
        radius[n], volume[n], eaten[n], id[n].  Pinned staging keeps the copy asynchronous.
        """
        n = int(arrays["radius"].shape[0])
        if n > self.cap:
            raise ValueError("upload of %d planets exceeds capacity %d" % (n, self.cap))
        t = self.t

        def put(key, src, dtype):
            a = np.ascontiguousarray(src, dtype=dtype)
            h = torch.from_numpy(a)
            if a.dtype == np.uint64:
                h = torch.from_numpy(a.view(np.int64))
            dst = t[key]
            dst[:n].copy_(h.reshape(dst[:n].shape), non_blocking=False)

        put("pos", arrays["pos"], np.float64); put("vel", arrays["vel"], np.float64)
        put("mass", arrays["mass_f64"], np.float64)
        put("mass_hi", arrays["mass_hi"], np.uint64); put("mass_lo", arrays["mass_lo"], np.uint64)
        put("radius", arrays["radius"], np.float64); put("volume", arrays["volume"], np.float64)
        put("eaten", arrays["eaten"], np.int64); put("id", arrays["id"], np.int32)
        put("flags", arrays["flags"], np.uint8)
        status = torch.zeros(STATUS_LEN, dtype=torch.int64)
        status[ST_N_ACTIVE] = n
        status[ST_STEP] = int(frame)
        keep = t["status"].cpu()
        status[ST_N_EVENTS] = keep[ST_N_EVENTS]
        status[ST_N_ESCAPED] = keep[ST_N_ESCAPED]
        t["status"].copy_(status)
        t["mod_step"].fill_(-1)
        self.n_upper = n
        self.frames = int(frame)
        torch.cuda.current_stream(self.device).synchronize()
        self._status_host.copy_(status)
        self._seed_snapshot(status)
        self._choose_scales(np.asarray(arrays["pos"], dtype=np.float64), np.asarray(arrays["mass_f64"], dtype=np.float64))

    def _tick(self):
        return self.frames

    STATE_KEYS = ("pos", "vel", "mass", "mass_hi", "mass_lo", "radius", "volume", "eaten", "id", "flags")

    def save_state(self):
        """Device-side copy of the active planets + status (for `restore_state`); synchronises once."""
        n = self.status()["n_active"]
        return {"n": n, "frames": self.frames, "status": self.t["status"].clone(),
                "status_host": self.t["status"].cpu(), "fields": {k: self.t[k][:n].clone() for k in self.STATE_KEYS},
                "rec_death": self.t["rec_death"].clone()}

    def restore_state(self, saved):
        """Put a `save_state` copy back (device-to-device, asynchronous, no host data involved).  The
        merge log / death records written since are kept; the event counter is rewound with the status."""
        n = saved["n"]
        for k in self.STATE_KEYS:
            self.t[k][:n].copy_(saved["fields"][k], non_blocking=True)
        self.t["status"].copy_(saved["status"], non_blocking=True)
        self.t["rec_death"].copy_(saved["rec_death"], non_blocking=True)
        self.t["mod_step"].fill_(-1)
        self.frames = saved["frames"]
        self.n_upper = n
        self._seed_snapshot(saved["status_host"])

    def _grid_changed(self):
        self._params_cache.clear()

    def _choose_scales(self, pos, mass):
        # exact power-of-two rescaling for the FP32 path: max |coord| -> [32, 64), max mass -> [2^19, 2^20)
        pmax = float(np.max(np.abs(pos))) if pos.size else 1.0
        mmax = float(np.max(np.abs(mass))) if mass.size else 1.0
        self.pos_exp = (math.frexp(pmax)[1] - 6) if pmax > 0 and math.isfinite(pmax) else 0
        self.mass_exp = (math.frexp(mmax)[1] - 20) if mmax > 0 and math.isfinite(mmax) else 0

    def status(self):
        """Synchronising read of the device status block."""
        s = self.t["status"].cpu().numpy()
        self.n_upper = int(s[ST_N_ACTIVE])
        return {"n_active": int(s[ST_N_ACTIVE]), "frame": int(s[ST_STEP]), "n_events": int(s[ST_N_EVENTS]),
                "n_pairs": int(s[ST_N_PAIRS]), "n_dead": int(s[ST_N_DEAD]), "error": int(s[ST_ERROR]),
                "n_escaped": int(s[ST_N_ESCAPED]), "n_giants": int(s[_abi.ST_N_GIANTS]),
                "n_candidates": int(s[_abi.ST_N_CAND])}

    def check_errors(self):
        st = self.status()
        err = st["error"]
        if err:
            what = [name for bit, name in ((1, "flagged-pair buffer overflow"), (2, "event log overflow"),
                                           (4, "non-finite acceleration (d^2 < 0 or NaN input)"),
                                           (8, "broad-phase giant list overflow")) if err & bit]
            what = what or ["error bits 0x%x" % err]
            raise _abi.CosmoError("device reported: " + ", ".join(what))
        return st

    def download(self):
        """Active planets (slot order == insertion order of survivors) + records of the dead."""
        st = self.status()
        n = st["n_active"]
        t = self.t
        out = {k: t[k][:n].cpu().numpy() for k in ("pos", "vel", "mass", "radius", "volume", "eaten", "id", "flags",
                                                   "energy")}
        tot = t["totals"].cpu().numpy()
        out["total_energy"], out["angular_momentum"] = float(tot[0]), float(tot[1])
        out["mass_hi"] = t["mass_hi"][:n].cpu().numpy().view(np.uint64)
        out["mass_lo"] = t["mass_lo"][:n].cpu().numpy().view(np.uint64)
        for k in ("rec_mass", "rec_radius", "rec_volume", "rec_eaten", "rec_death", "rec_flags"):
            out[k] = t[k].cpu().numpy()
        out["rec_mass_hi"] = t["rec_mass_hi"].cpu().numpy().view(np.uint64)
        out["rec_mass_lo"] = t["rec_mass_lo"].cpu().numpy().view(np.uint64)
        out["status"] = st
        return out

    def view(self, energy=False):
        """(n, id, pos, radius[, energy, totals]) of the active planets: the one small device->host copy
        a front-end needs per displayed frame (no death records, no masses, no Planet objects)."""
        n = self.status()["n_active"]
        t = self.t
        out = {"n": n, "id": t["id"][:n].cpu().numpy(), "pos": t["pos"][:n].cpu().numpy(),
               "radius": t["radius"][:n].cpu().numpy()}
        if energy:
            out["energy"] = t["energy"][:n].cpu().numpy()
            tot = t["totals"].cpu().numpy()
            out["total_energy"], out["angular_momentum"] = float(tot[0]), float(tot[1])
        return out

    def peek(self, slot):
        """One active planet (by slot) as Python numbers: a few words device->host."""
        t, k = self.t, int(slot)
        fl = int(t["flags"][k])
        mass = ((int(t["mass_hi"][k]) & _MASK64) << 64 | (int(t["mass_lo"][k]) & _MASK64)) if fl & 2 else float(t["mass"][k])
        return {"id": int(t["id"][k]), "pos": t["pos"][k].cpu().numpy(), "vel": t["vel"][k].cpu().numpy(), "mass": mass,
                "radius": float(t["radius"][k]), "eaten": int(t["eaten"][k]), "energy": float(t["energy"][k])}

    def events(self, start=0):
        """Merge stream rows (frame, absorber id, absorbed id) from `start` on."""
        n = min(self.status()["n_events"], self.event_cap)
        return self.t["events"][start:n].cpu().numpy()

    def last_frame_tensors(self):
        """acc / pos_new / vel_new of the last frame in the slot order the frame STARTED with."""
        cap = self.cap
        return {"acc": self.t["acc"], "pos_new": self._pos_new.view(cap, 2), "vel_new": self._vel_new.view(cap, 2)}

    # ------------------------------------------------------------------ stepping
    def _sym_tiling(self, n):
        """(row blocks per CTA, column blocks per CTA) of the symmetric FP32 kernel: about 16 CTAs per
        resident slot and rank keep the tail short; within that, as many row blocks per CTA as possible
        (the column-partial traffic and buffer shrink by that factor), and no more column chunks than the
        row-partial buffer has rows (JSPLIT_MAX)."""
        nblk = -(-n // 1024)
        target = (nblk * nblk) // (2 * 16 * 2 * self.sm_count * self.world) if nblk > 1 else 1
        # (a CTA works on two row blocks at once, so 2 is the smallest tiling that keeps all its warps busy;
        # sym_rows = 1 remains available for tests)
        rows = 4 if target >= 16 else 2
        chunk = int(min(32, max(1, -(-nblk // self.JSPLIT_MAX), target // rows)))
        return rows, chunk

    def _ensure_colbuf(self, fp64=False):
        """Column-partial buffer of the symmetric kernels.  FP32: one double2 row of `cap` (= two float2
        rows) per SUPER row (sym_rows row blocks of 1024 planets): 16 B x N^2 / (1024 sym_rows ranks),
        4.3 GB at N = 2^20 on one GPU with sym_rows = 4.  HBM capacity is what makes the deterministic,
        atomic-free reduction affordable."""
        def rows_for(n):
            local = -(-(-(-n // PAD)) // self.world) + 1        # row blocks of this rank (+1: partial round)
            r = self.sym_rows or self._sym_tiling(n)[0]
            return 2 * -(-local // r)
        rows = rows_for(max(1, self.n_upper))
        if not fp64 and ("colbuf" not in self.t or self.t["colbuf"].shape[0] < rows):
            # allocate once for the whole run: as planets merge the active count shrinks and the tiling
            # drops to fewer row blocks per CTA (more, smaller colbuf rows) -- never reallocate mid-run
            rows = max(rows_for(n) for n in range(PAD, max(1, self.n_upper) + PAD, PAD))
        if fp64:                                  # 512-planet row blocks, double2 = two float2 rows each
            rows = 2 * -(-(self.cap // 512) // self.world)
        if "colbuf" not in self.t or self.t["colbuf"].shape[0] < rows:
            self.t["colbuf"] = torch.empty(rows, self.cap, 2, dtype=torch.float32, device=self.device)
            self.sc.colbuf = C.c_void_p(self.t["colbuf"].data_ptr())
            self.sc.colbuf_rows = rows

    def params(self, dt, G, r_max, **kw):
        """Step parameters for the current launch bound (a private copy the caller may edit)."""
        return _abi.StepParams.from_buffer_copy(self._params(dt, G, r_max, **kw))

    def _params(self, dt, G, r_max, **kw):
        """Cached step parameters (shared object: do not modify): they only depend on the arguments,
        the launch bound and a few tunables."""
        key = (self.n_upper, dt, G, r_max, tuple(sorted(kw.items())), self.near_field, self._grid_ready,
               self._force_all_pairs, self._host_grid, self.grid_threshold, self.small_threshold,
               self.sym_threshold, self.sym64_threshold, self.sym_chunk, self.sym_rows, self.pos_exp, self.mass_exp)
        p = self._params_cache.get(key)
        if p is None:
            if len(self._params_cache) > 64:
                self._params_cache.clear()
            p = self._params_cache[key] = self._make_params(dt, G, r_max, **kw)
        return p

    def _make_params(self, dt, G, r_max, collisions=True, bounded=True, fp32=False, store_acc=False,
                     jsplit=None, f32_sym=None, energy=False, f64_sym=None, near=None):
        p = _abi.StepParams()
        p.G, p.dt, p.r_max = float(G), float(dt), float(r_max)
        n = self.n_upper
        sym = bool(fp32 and (f32_sym if f32_sym is not None else n >= self.sym_threshold))
        sym64 = bool((not fp32) and (f64_sym if f64_sym is not None else n >= self.sym64_threshold))
        if sym:
            self._ensure_colbuf()
        if sym64:
            self._ensure_colbuf(fp64=True)
        p.opts = ((OPT_COLLISIONS if collisions else 0) | (OPT_BOUNDED if bounded else 0) |
                  (OPT_FP32 if fp32 else 0) | (OPT_STORE_ACC if store_acc else 0) |
                  (OPT_F32_SYM if sym else 0) |
                  (OPT_F64_SYM if sym64 else 0) |
                  (OPT_DEFER_INTEGRATE if ((sym or sym64) and self.world > 1) else 0) | (OPT_ENERGY if energy else 0))
        if (sym or sym64) and self.world > 1:
            from . import sharded
            if sharded.setup_peer_exchange(self):
                p.opts |= _abi.OPT_PEER_EXCHANGE      # sharded.frame points scratch.peer_acc at the frame's parity
        if sym:
            p.sym_rows, p.sym_chunk = self._sym_tiling(n)
            if self.sym_chunk:
                p.sym_chunk = int(self.sym_chunk)
            if self.sym_rows:
                p.sym_rows = int(self.sym_rows)
        elif self.sym_chunk or not sym64:
            p.sym_chunk = int(self.sym_chunk)
        else:
            # column blocks per CTA: enough CTAs (~16 per resident slot and rank) to keep the tail short,
            # few enough that the row partials fit `partial` (JSPLIT_MAX rows)
            nblk = -(-n // 512)
            want = (nblk * nblk) // (2 * 16 * 2 * self.sm_count * self.world) if nblk > 1 else 1
            p.sym_chunk = int(min(32, max(1, -(-nblk // self.JSPLIT_MAX), want)))
        p.n_upper = n
        if self.world > 1:
            from . import sharded
            p.i_begin, p.i_end, _ = sharded.shard_rows(n, self.world, self.rank)
        else:
            p.i_begin, p.i_end = 0, n
        p.pos_exp, p.mass_exp = self.pos_exp, self.mass_exp
        rows = max(1, p.i_end - p.i_begin)
        p.small_tiles = 1 if n <= self.small_threshold else 0
        use_grid = bool(collisions and not self._force_all_pairs and n >= self.grid_threshold)
        if use_grid:
            # grid tuned ON THE DEVICE from the frame's own positions and radii (detect_mode 2) unless
            # the caller pinned one (detect_mode 1): the host never reads planet data while stepping
            p.resolve_mode = 1
            if self._host_grid is not None:
                p.detect_mode = 1
                p.cell_size, p.grid_origin_x, p.grid_origin_y, p.cells_per_side, p.giant_radius = self._host_grid
            else:
                p.detect_mode = 2
        else:
            p.detect_mode, p.resolve_mode = 0, 0
        p.jsplit = p.jsplit_detect = int(jsplit) if jsplit else 1
        if not jsplit:
            with torch.cuda.device(self.device):
                if not (sym or sym64):        # the one-sided kernels split the source range
                    kern = (4 if p.small_tiles else 0) if not fp32 else 1
                    p.jsplit = self.lib.cosmo_suggest_jsplit(rows, max(1, n), kern, self.JSPLIT_MAX)
                if collisions and not use_grid:
                    p.jsplit_detect = self.lib.cosmo_suggest_jsplit(rows, max(1, n), 5 if p.small_tiles else 3,
                                                                    self.JSPLIT_MAX)
        p.shard_index, p.shard_count = self.rank, self.world
        # FP32 path: pairs closer than 2^14 FP32 ulps of their coordinates are re-evaluated in FP64
        # (csrc/nearfield.cu), binned with the grid the previous frame's stage B left on the device
        if fp32 and (self.near_field if near is None else near):
            p.near_mode = 1 if self._grid_ready else 2
        return p

    def step(self, n_frames, dt, G, r_max, stage_events=None, **kw):
        """Enqueue n_frames frames on the current stream.  The host never reads planet data here; the
        only wait is on the status snapshot of BOUND_LAG frames ago (see __init__), which bounds the
        queue depth and does not drain the GPU.

        stage_events: optional list; for every frame a tuple of four torch.cuda.Event
        (start, after stage A, after stage B, after stage C) is appended -- bench.py uses them
        to time the all-pairs kernel on the launching stream."""
        if self.n_upper <= 0 or n_frames <= 0:
            return
        p = self._params(dt, G, r_max, **kw)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        st, sc, pp = C.byref(self.st), C.byref(self.sc), C.byref(p)
        lib = self.lib
        if (self.use_graphs and self.world == 1 and stage_events is None and n_frames >= 8
                and self.n_upper <= self.graph_threshold):
            return self._step_graphed(int(n_frames), dt, G, r_max, kw)
        for _ in range(int(n_frames)):
            if self._refresh_bound():
                p = self._params(dt, G, r_max, **kw)
                pp = C.byref(p)
            if self.world > 1:
                self._step_sharded(p, stream, stage_events)
            elif stage_events is None:
                _abi.check(lib.cosmo_step(st, sc, pp, stream), "cosmo_step")
            else:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
                ev[0].record()
                _abi.check(lib.cosmo_begin_frame(st, sc, pp, stream), "cosmo_begin_frame")
                ev[1].record()
                _abi.check(lib.cosmo_accel_integrate(st, sc, pp, stream), "cosmo_accel_integrate")
                if p.opts & OPT_ENERGY:
                    _abi.check(lib.cosmo_energies(st, sc, pp, stream), "cosmo_energies")
                ev[2].record()
                _abi.check(lib.cosmo_detect_pairs(st, sc, pp, stream), "cosmo_detect_pairs")
                ev[3].record()
                _abi.check(lib.cosmo_resolve_commit(st, sc, pp, stream), "cosmo_resolve_commit")
                ev[4].record()
                stage_events.append(ev)
            self.frames += 1
            self._push_snapshot()
            if p.detect_mode != 0 and not self._grid_ready:
                self._grid_ready = True
                p = self._params(dt, G, r_max, **kw)
                pp = C.byref(p)

    def _step_graphed(self, n_frames, dt, G, r_max, kw):
        """Launch-bound regime (N of a few thousand: a frame is ~10 kernels of a few microseconds):
        capture cosmo_step + the status snapshot once into a CUDA graph and replay it.  Kernels read
        the active count from the device, so a graph stays valid while planets merge; it is
        re-captured when the launch bound has shrunk by more than 10 %."""
        done = 0
        while done < n_frames:
            seen_n, seen_f = int(self._status_host[ST_N_ACTIVE]), int(self._status_host[ST_STEP])
            if self._status_epoch <= seen_f <= self.frames and 0 < seen_n < 0.9 * self.n_upper:
                self.n_upper = seen_n
            p = self.params(dt, G, r_max, **kw)
            key = bytes(memoryview(p))
            g = self._graphs.get(key)
            if g is None:
                if len(self._graphs) > 16:
                    self._graphs.clear()
                st, sc, pp = C.byref(self.st), C.byref(self.sc), C.byref(p)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
                    _abi.check(self.lib.cosmo_step(st, sc, pp, stream), "cosmo_step (graph capture)")
                    self._status_host.copy_(self.t["status"], non_blocking=True)
                self._graphs[key] = g
            burst = min(n_frames - done, 64)
            for _ in range(burst):
                g.replay()
            done += burst
            self.frames += burst

    def _step_sharded(self, p, stream, stage_events=None):
        from . import sharded
        sharded.frame(self, p, stream, stage_events)

    # ------------------------------------------------------------------ host-buffer (e2e) path
    E2E_IN = ("pos", "vel", "mass", "mass_hi", "mass_lo", "radius", "volume", "eaten", "id", "flags")
    E2E_OUT = ("pos", "vel", "mass", "radius", "id")

    def make_host_staging(self, arrays):
        """Pinned host mirrors of the state (what a host-resident caller hands over per frame) and the
        cosmo_host_frame that points at them."""
        n = int(arrays["radius"].shape[0])
        src = {"pos": arrays["pos"], "vel": arrays["vel"], "mass": arrays["mass_f64"],
               "mass_hi": np.asarray(arrays["mass_hi"]).view(np.int64), "mass_lo": np.asarray(arrays["mass_lo"]).view(np.int64),
               "radius": arrays["radius"], "volume": arrays["volume"], "eaten": arrays["eaten"],
               "id": arrays["id"], "flags": arrays["flags"]}
        stg = {"n": n, "in": {}, "out": {}}
        for k in self.E2E_IN:
            stg["in"][k] = torch.from_numpy(np.ascontiguousarray(src[k])).pin_memory()
        for k in self.E2E_OUT:
            stg["out"][k] = torch.empty_like(self.t[k][:n], device="cpu").pin_memory()
        stg["status_in"] = torch.zeros(STATUS_LEN, dtype=torch.int64).pin_memory()
        stg["status_out"] = torch.zeros(STATUS_LEN, dtype=torch.int64).pin_memory()
        stg["events_out"] = torch.zeros(max(1, n), 3, dtype=torch.int64).pin_memory()
        f = _abi.HostFrame()
        f.n = n
        for k in self.E2E_IN:
            setattr(f, k, C.c_void_p(stg["in"][k].data_ptr()))
        for k in self.E2E_OUT:
            setattr(f, k + "_out", C.c_void_p(stg["out"][k].data_ptr()))
        f.status_out = C.c_void_p(stg["status_out"].data_ptr())
        f.events_out = C.c_void_p(stg["events_out"].data_ptr())
        f.events_cap = max(1, n)
        stg["frame"] = f
        self._e2e_row_bytes_out = sum(self.t[k][:1].numel() * self.t[k].element_size() for k in self.E2E_OUT)
        stg["h2d_bytes"] = sum(v.numel() * v.element_size() for v in stg["in"].values()) + STATUS_LEN * 8
        stg["d2h_bytes"] = STATUS_LEN * 8          # updated by every frame_from_host: survivors + events + status
        stg["path"] = ("cosmo_step_host (C ABI): pinned host SoA -> H2D -> frame -> status D2H -> survivors + merge "
                       "events D2H, stream synchronised")
        return stg

    def frame_from_host(self, stg, dt, G, r_max, **kw):
        """One frame the way a host-resident caller sees it.  One GPU: a single call of the C ABI's
        cosmo_step_host (state H2D from pinned memory, the frame, survivors and merge events D2H,
        stream synchronised).  Sharded: the same steps sequenced here, because the exchanges between
        the stages are torch.distributed collectives.  Returns the device status block."""
        n = stg["n"]
        t = self.t
        if self.world == 1:
            self.n_upper = n
            p = self._params(dt, G, r_max, **kw)
            f = stg["frame"]
            f.frame = self.frames
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            with torch.cuda.device(self.device):
                _abi.check(self.lib.cosmo_step_host(C.byref(self.st), C.byref(self.sc), C.byref(p), C.byref(f), stream),
                           "cosmo_step_host")
            self.frames += 1
            if p.detect_mode != 0:
                self._grid_ready = True
            so = stg["status_out"]
            self._seed_snapshot(so)
            stg["d2h_bytes"] = (int(so[ST_N_ACTIVE]) * self._e2e_row_bytes_out + min(int(so[ST_N_EVENTS]), n) * 24
                                + STATUS_LEN * 8)
            return so
        for k in self.E2E_IN:
            t[k][:n].copy_(stg["in"][k].view(t[k][:n].shape), non_blocking=True)
        stg["status_in"].zero_()
        stg["status_in"][ST_N_ACTIVE] = n
        stg["status_in"][ST_STEP] = self.frames
        t["status"].copy_(stg["status_in"], non_blocking=True)
        self.n_upper = n
        self._seed_snapshot(stg["status_in"])     # snapshots of earlier frames are void
        self.step(1, dt, G, r_max, **kw)
        for k in self.E2E_OUT:
            stg["out"][k].copy_(t[k][:n], non_blocking=True)
        stg["status_out"].copy_(t["status"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        stg["d2h_bytes"] = sum(v.numel() * v.element_size() for v in stg["out"].values()) + STATUS_LEN * 8
        stg["path"] = "Engine.frame_from_host (sharded): pinned host SoA -> H2D -> frame with NCCL exchanges -> D2H, stream sync per step"
        return stg["status_out"]
