"""Host-side launch bound and grid control shared by Engine (2-D frame) and Engine3D (3-D step).

While stepping, the host never reads planet data.  What it does need is an UPPER BOUND of the
active count to size kernel grids (the kernels read the exact count from the device).  After every
step the device status block is copied asynchronously into a ring of pinned host buffers, each
followed by an event.  The bound of step f is the count in the snapshot taken BOUND_LAG steps
earlier (counts only shrink, so a stale value is a safe bound): waiting for THAT event keeps at
most BOUND_LAG steps in flight, never drains the GPU, and -- replicas being bit-identical -- gives
every rank of a sharded run the same bound without any communication.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from ._abi import ST_N_ACTIVE, ST_STEP, STATUS_LEN


class BoundTracker:
    SNAP_RING = 4
    BOUND_LAG = 2

    def _init_bound(self):
        self._snap = [torch.zeros(STATUS_LEN, dtype=torch.int64).pin_memory() for _ in range(self.SNAP_RING)]
        self._snap_ev = [torch.cuda.Event() for _ in range(self.SNAP_RING)]
        self._snap_stamp = [-1] * self.SNAP_RING      # step counter the snapshot belongs to
        self._snap_async = [False] * self.SNAP_RING   # True: wait for the event before reading
        self._status_host = torch.zeros(STATUS_LEN, dtype=torch.int64).pin_memory()   # CUDA-graph replay only
        self._status_epoch = 0       # step counter at the last upload (guards against stale reads)
        self._host_grid = None       # caller-supplied grid (cell edge, origin x, origin y, cells per side, giant radius)
        self._force_all_pairs = False
        self.n_upper = 0
        self.pairs_lag = 0           # flagged pairs of the step BOUND_LAG back (sizes stage C)

    # the engines count their steps under different names (Universe.iterations / State.iteration)
    def _tick(self) -> int:
        raise NotImplementedError

    def _seed_snapshot(self, status):
        """The host just wrote the device status block itself: snapshots older than this are void."""
        now = self._tick()
        self._status_epoch = now
        k = now % self.SNAP_RING
        self._snap[k].copy_(status)
        self._snap_stamp[k], self._snap_async[k] = now, False

    def _push_snapshot(self):
        """Asynchronous copy of the status block after a step (the step counter already counts it)."""
        now = self._tick()
        k = now % self.SNAP_RING
        self._snap[k].copy_(self.t["status"], non_blocking=True)
        self._snap_ev[k].record(torch.cuda.current_stream(self.device))
        self._snap_stamp[k], self._snap_async[k] = now, True

    def _refresh_bound(self):
        """Launch bound of the next step from the snapshot BOUND_LAG steps back; True if it changed."""
        want = self._tick() - self.BOUND_LAG
        if want < self._status_epoch:
            return False
        k = want % self.SNAP_RING
        if self._snap_stamp[k] != want:
            return False
        if self._snap_async[k]:
            self._snap_ev[k].synchronize()
            self._snap_async[k] = False
        seen_n, seen_f = int(self._snap[k][ST_N_ACTIVE]), int(self._snap[k][ST_STEP])
        if seen_f == want:
            self.pairs_lag = int(self._snap[k][_abi.ST_N_PAIRS])
        if seen_f == want and 0 < seen_n < self.n_upper:
            self.n_upper = seen_n
            return True
        return False

    # ------------------------------------------------------------------ broad-phase grid
    @property
    def grid(self):
        """Broad-phase grid of the last step as (cell edge, centre x, centre y, cells per side, giant
        radius) -- a synchronising read of the device-resident cosmo_grid, for tests and diagnostics --
        or None when the all-pairs predicate is in use.  Assigning a tuple (cell edge, origin x,
        origin y, cells per side, giant radius) pins the grid (detect_mode 1), assigning None forces
        the all-pairs predicate; `del eng.grid` returns to the device-side tuning (detect_mode 2)."""
        if self._force_all_pairs or self.n_upper < self.grid_threshold:
            return None
        g = self.grid_info()
        return (g["h"], g["cx"], g["cy"], g["c"], g["rg"]) if g["source"] else ()

    def grid_info(self):
        """All fields of the device-resident cosmo_grid (synchronising read; diagnostics)."""
        raw = self.t["grid_dev"].cpu().numpy().tobytes()[:C.sizeof(_abi.Grid)]
        g = _abi.Grid.from_buffer_copy(raw)
        return {name: getattr(g, name) for name, _ in _abi.Grid._fields_}

    @grid.setter
    def grid(self, value):
        self._force_all_pairs = value is None
        self._host_grid = None if value is None else tuple(value)
        self._grid_changed()

    @grid.deleter
    def grid(self):
        self._force_all_pairs, self._host_grid = False, None
        self._grid_changed()

    def pin_grid(self):
        """Pin the grid the device tuned for the last step (it otherwise adapts every step, using the
        candidate density the previous step measured)."""
        h, cx, cy, c, rg = self.grid
        self.grid = (h, cx - 0.5 * c * h, cy - 0.5 * c * h, c, rg)

    def _grid_changed(self):
        pass
