"""Multi-GPU frame: target rows sharded over the ranks of one node, state replicated.

One process per GPU (torch.distributed, NCCL over NVLink 5 / NVSwitch).  Per frame:
  stage A on this rank's rows            -> pos_new / vel_new / xx_new of the shard
  all-gather of the three planes         (the collision test needs everyone's NEW position,
                                          reference universe_old.py:178-180)
  stage B on this rank's rows            -> flagged ordered pairs (i, j) of the shard
  all-gather of the per-rank pair lists  -> identical pair set on every rank
  stage C replicated                     -> identical state on every rank, no further traffic
The partition helpers are pure functions of (n, world) so the gloo/CPU tests can drive the
exchange protocol without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _abi
from ._abi import PAD


def shard_rows(n: int, world: int, rank: int):
    """Contiguous, tile-aligned row shard [begin, end) of rank; every rank's `per` is equal."""
    per = (-(-n // world) + PAD - 1) // PAD * PAD if n > 0 else PAD
    return min(n, rank * per), min(n, (rank + 1) * per), per


def required_capacity(n_total: int, world: int) -> int:
    """Slots needed so that world equal shards of whole tiles fit."""
    return shard_rows(max(1, n_total), world, 0)[2] * world


def gather_planes(planes, per: int, rank: int, world: int, group=None):
    """In-place all-gather of [cap*w] planes whose rows [rank*per, (rank+1)*per) are local."""
    for plane, width in planes:
        full = plane[: world * per * width]
        mine = plane[rank * per * width: (rank + 1) * per * width]
        dist.all_gather_into_tensor(full, mine, group=group)


def exchange_pairs(local_pairs: torch.Tensor, local_count: torch.Tensor, xchg: torch.Tensor,
                   rank: int, world: int, group=None):
    """All-gather (count | pairs[:xcap]) rows.  xchg: int64 [world, 1 + xcap], updated in place.
    (Reference form of the exchange protocol for the CPU tests; the GPU frame lets stage B write its
    pairs straight into the row and seals the count with one tiny kernel, see _detect_and_exchange.)"""
    xcap = xchg.shape[1] - 1
    xchg[rank, 0] = local_count.reshape(())
    xchg[rank, 1:] = local_pairs[:xcap]
    dist.all_gather_into_tensor(xchg.view(-1), xchg[rank], group=group)
    return xchg


def _detect_and_exchange(eng, detect, seal, append, st, sc, pp, stream, ev_after_detect=None):
    """Stage B on this rank's share + the exchange of the flagged pairs, without a single copy or
    eager torch kernel: the detection kernels append straight into this rank's row of the exchange
    buffer (scratch.pairs is pointed there for the call), one thread seals the count into row[0]
    and resets the counter, NCCL all-gathers the rows, and one launch re-emits every rank's pairs
    into scratch.pairs."""
    if not hasattr(eng, "_xchg"):
        # a rank's share of the flagged pairs is balanced by the round-robin chunking of stage B, so
        # a fraction of pair_cap is plenty (overflow raises COSMO_ERR_PAIR_OVERFLOW, never silent)
        xcap = max(1024, min(eng.pair_cap // eng.world, max(1 << 16, eng.cap // eng.world)))
        eng._xchg = torch.zeros(eng.world, 1 + xcap, dtype=torch.int64, device=eng.device)
        eng._xchg_flat, eng._xchg_row = eng._xchg.view(-1), eng._xchg[eng.rank]
    xcap = eng._xchg.shape[1] - 1
    row_ptr = eng._xchg_row.data_ptr()
    keep = (eng.sc.pairs, eng.sc.pair_cap)
    eng.sc.pairs, eng.sc.pair_cap = row_ptr + 8, xcap
    try:
        _abi.check(detect(st, sc, pp, stream), "detect_pairs")
    finally:
        eng.sc.pairs, eng.sc.pair_cap = keep
    if ev_after_detect is not None:
        ev_after_detect.record()          # stage B kernels done, pair exchange not yet
    _abi.check(seal(st, C.c_void_p(row_ptr), stream), "seal_pair_row")
    dist.all_gather_into_tensor(eng._xchg_flat, eng._xchg_row, group=eng.group)
    _abi.check(append(st, sc, C.c_void_p(eng._xchg.data_ptr()), eng.world, xcap, stream), "append_pair_rows")


_PEER_CACHE = {}
PEER_BARRIER_TIMEOUT_MS = 20000      # a lost rank turns into an error, not a hang


def setup_peer_exchange(eng):
    """Map every rank's exchange buffer into this process (NVLink peer memory) through
    torch.distributed's symmetric-memory plumbing; the data path itself is in the CUDA kernels
    (k_sym_reduce stores into the peers, k_sym_integrate adds the slots in rank order).
    Two buffers (frame parity) so one barrier per frame is enough.  Returns False -- and the frame
    falls back to an NCCL all-reduce -- if symmetric memory cannot be set up here."""
    if getattr(eng, "_peer", None) is not None:
        return bool(eng._peer)
    try:
        # opt-in (experimental): COSMO_PEER_EXCHANGE=1.  The default exchange is NCCL.
        if os.environ.get("COSMO_PEER_EXCHANGE", "0") != "1":
            raise RuntimeError("not enabled (set COSMO_PEER_EXCHANGE=1)")
        import torch.distributed._symmetric_memory as symm_mem
        if eng.world > 16:
            raise RuntimeError("more than COSMO_MAX_PEERS ranks")
        key = (eng.world, eng.cap, str(eng.device))
        if key not in _PEER_CACHE:      # one symmetric buffer per process and shape, never freed:
            numel = 2 * eng.world * eng.cap * 2                   # [parity][rank slot][cap][2] doubles
            buf = symm_mem.empty(numel, dtype=torch.float64, device=eng.device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, eng.group if eng.group is not None else dist.group.WORLD)
            _PEER_CACHE[key] = {"buf": buf, "hdl": hdl, "ptrs": [int(x) for x in hdl.buffer_ptrs],
                                "half_bytes": eng.world * eng.cap * 2 * 8}
        eng._peer = _PEER_CACHE[key]
        eng.sc.n_peers = eng.world
        eng._peer["hdl"].barrier(0, PEER_BARRIER_TIMEOUT_MS)
    except Exception as e:                                        # no fabric / IPC support in this container
        eng._peer = False
        eng._peer_error = "%s: %s" % (type(e).__name__, e)
    return bool(eng._peer)


def frame(eng, p, stream, stage_events=None):
    """One frame on engine `eng` (params p already hold this rank's row shard).

    stage_events: optional list receiving (start, prep, stage A + exchange, stage B + exchange,
    stage C) torch.cuda.Event tuples for bench.py."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)] if stage_events is not None else None
    if ev:
        ev[0].record()
    lib = eng.lib
    st, sc, pp = C.byref(eng.st), C.byref(eng.sc), C.byref(p)
    n = p.n_upper
    _, _, per = shard_rows(n, eng.world, eng.rank)
    if p.opts & _abi.OPT_PEER_EXCHANGE:
        # exchange buffer of this frame's parity, as mapped from every rank (two buffers: a rank may
        # already store frame t+1 while a slower one still reads frame t)
        off = (eng.frames & 1) * eng._peer["half_bytes"]
        for r, base in enumerate(eng._peer["ptrs"]):
            eng.sc.peer_acc[r] = base + off
    _abi.check(lib.cosmo_begin_frame(st, sc, pp, stream), "cosmo_begin_frame")
    if ev:
        ev[1].record()
    _abi.check(lib.cosmo_accel_integrate(st, sc, pp, stream), "cosmo_accel_integrate")
    if ev:
        ev[5].record()          # stage A kernels done, exchange not yet
    if p.opts & _abi.OPT_PEER_EXCHANGE:
        # symmetric kernels, peer exchange: the reduce kernel above has already stored this rank's
        # partial sums into every peer over NVLink; one device-side barrier, then every rank adds
        # the ranks' slots in rank order (deterministic, identical everywhere) and integrates
        eng._peer["hdl"].barrier(0, PEER_BARRIER_TIMEOUT_MS)
        _abi.check(lib.cosmo_integrate(st, sc, pp, stream), "cosmo_integrate")
    elif p.opts & _abi.OPT_DEFER_INTEGRATE:
        # symmetric kernels, NCCL: every rank holds partial sums for ALL planets (its row blocks +
        # the column sides of its pairs) -> all-reduce, then every rank integrates all rows
        dist.all_reduce(eng.t["acc"][:n], op=dist.ReduceOp.SUM, group=eng.group)
        _abi.check(lib.cosmo_integrate(st, sc, pp, stream), "cosmo_integrate")
    else:
        gather_planes(((eng._pos_new, 2), (eng._vel_new, 2), (eng._xx_new, 1)), per, eng.rank, eng.world, eng.group)
    if ev:
        ev[2].record()
    if p.opts & _abi.OPT_ENERGY:      # diagnostics: replicated over all rows (new positions are complete here)
        _abi.check(lib.cosmo_energies(st, sc, pp, stream), "cosmo_energies")
    if p.opts & _abi.OPT_COLLISIONS:
        _detect_and_exchange(eng, lib.cosmo_detect_pairs, lib.cosmo_seal_pair_row, lib.cosmo_append_pair_rows,
                             st, sc, pp, stream, ev[6] if ev else None)
    elif ev:
        ev[6].record()
    if ev:
        ev[3].record()
    _abi.check(lib.cosmo_resolve_commit(st, sc, pp, stream), "cosmo_resolve_commit")
    if ev:
        ev[4].record()
        stage_events.append(ev)


def frame3d(eng, p, stream, stage_events=None):
    """One State.interact on Engine3D `eng`, rows sharded over the ranks (params p hold this rank's
    shard).  Same protocol as `frame`: stage A on the shard -> all-gather of the seven new planes (or,
    with the symmetric kernel, all-reduce of the partial sums + replicated integrate) -> stage B on
    the shard -> all-gather of the pair lists -> replicated stage C."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if stage_events is not None else None
    if ev:
        ev[0].record()
    lib = eng.lib
    st, sc, pp = C.byref(eng.st), C.byref(eng.sc), C.byref(p)
    n = p.n_upper
    _, _, per = shard_rows(n, eng.world, eng.rank)
    _abi.check(lib.cosmo3_accel_integrate(st, sc, pp, stream), "cosmo3_accel_integrate")
    if p.opts & _abi.OPT3_DEFER_INTEGRATE:
        dist.all_reduce(eng.t["acc"], op=dist.ReduceOp.SUM, group=eng.group)
        _abi.check(lib.cosmo3_integrate(st, sc, pp, stream), "cosmo3_integrate")
    else:
        planes = [(eng.t[k][d], 1) for k in ("pos_new", "vel_new") for d in range(3)] + [(eng.t["xx_new"], 1)]
        gather_planes(planes, per, eng.rank, eng.world, eng.group)
    if ev:
        ev[1].record()
    if p.opts & _abi.OPT3_COLLISIONS:
        _detect_and_exchange(eng, lib.cosmo3_detect_pairs, lib.cosmo3_seal_pair_row, lib.cosmo3_append_pair_rows,
                             st, sc, pp, stream)
    if ev:
        ev[2].record()
    _abi.check(lib.cosmo3_resolve_commit(st, sc, pp, stream), "cosmo3_resolve_commit")
    if ev:
        ev[3].record()
        stage_events.append(ev)
