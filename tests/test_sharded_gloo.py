"""Host-side logic of the multi-GPU frame on CPU: world_size-2 gloo process group.

The row partition, the in-place plane all-gather and the pair-list exchange of
cosmosim_b200/sharded.py are device-agnostic torch code; here each rank produces its shard
with the CPU oracle (test infrastructure) and the gathered result must equal the unsharded
oracle.  The CUDA kernels themselves are covered by the -m gpu tests and tools/sharded_check.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partition():
    from cosmosim_b200.sharded import required_capacity, shard_rows
    for n in (1, 511, 512, 513, 1001, 65536, 1048576, 531572):
        for world in (1, 2, 4, 8):
            spans = [shard_rows(n, world, r) for r in range(world)]
            per = spans[0][2]
            assert per % 1024 == 0 and all(s[2] == per for s in spans)
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0] or (a[1] == n and b[0] == n)
            assert required_capacity(n, world) >= max(n, world * per)


def _worker(rank, world, port, n, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cosmosim_b200 import scenes, sharded
        from oracle import oracle as orc
        arr = scenes.star_disk_arrays(n, seed=4)
        arr["radius"] = arr["radius"] * 12.0          # make collisions plentiful
        G, dt = 6.674e-11, 1e6 / 33
        b, e, per = sharded.shard_rows(n, world, rank)
        cap = sharded.required_capacity(n, world)
        # stage A on this rank's rows
        acc = orc.accel(arr["pos"], arr["mass_f64"], G, b, e)
        p_loc, v_loc = orc.integrate(arr["pos"][b:e], arr["vel"][b:e], acc, dt)
        pos_new = torch.zeros(cap * 2, dtype=torch.float64)
        vel_new = torch.zeros(cap * 2, dtype=torch.float64)
        xx_new = torch.zeros(cap, dtype=torch.float64)
        pos_new[2 * b:2 * e] = torch.from_numpy(p_loc.reshape(-1))
        vel_new[2 * b:2 * e] = torch.from_numpy(v_loc.reshape(-1))
        xx_new[b:e] = torch.from_numpy(p_loc[:, 0] * p_loc[:, 0] + p_loc[:, 1] * p_loc[:, 1])
        sharded.gather_planes(((pos_new, 2), (vel_new, 2), (xx_new, 1)), per, rank, world)
        p_all = pos_new[: 2 * n].view(n, 2).numpy()
        # stage B on this rank's rows, then the pair-list exchange
        pairs = orc.flags(p_all, arr["radius"], b, e)
        keys = torch.from_numpy((pairs[:, 0] << 32 | pairs[:, 1]).astype(np.int64))
        xcap = 1 << 14
        local = torch.zeros(xcap, dtype=torch.int64)
        local[: len(keys)] = keys
        xchg = torch.zeros(world, 1 + xcap, dtype=torch.int64)
        sharded.exchange_pairs(local, torch.tensor(len(keys)), xchg, rank, world)
        merged = torch.cat([xchg[r, 1:1 + int(xchg[r, 0])] for r in range(world)]).numpy()
        if rank == 0:
            out_q.put((p_all.copy(), np.sort(merged)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_exchange_equals_unsharded_oracle():
    from cosmosim_b200 import scenes
    from oracle import oracle as orc
    n, world = 1500, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    p_all, merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    arr = scenes.star_disk_arrays(n, seed=4)
    arr["radius"] = arr["radius"] * 12.0
    acc = orc.accel(arr["pos"], arr["mass_f64"], 6.674e-11)
    p_ref, _ = orc.integrate(arr["pos"], arr["vel"], acc, 1e6 / 33)
    assert np.array_equal(p_all, p_ref)                      # sharded rows reassemble bit-exactly
    pairs = orc.flags(p_ref, arr["radius"])
    ref = np.sort((pairs[:, 0] << 32 | pairs[:, 1]).astype(np.int64))
    assert len(ref) > 10 and np.array_equal(merged, ref)     # same flagged set on every rank


# ---------------------------------------------------------------------------------------------
# 3-D step (cosmosim_b200.sharded.frame3d): seven planes of width 1 instead of three interleaved ones
# ---------------------------------------------------------------------------------------------
def _worker3d(rank, world, port, n, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cosmosim_b200 import scenes, sharded
        from oracle import oracle3d as o3
        arr = scenes.disk3d_arrays(n, seed=4, r_in=0.1, r_out=0.2, density=10.0)
        G, dt = 6.674e-11, 600.0
        b, e, per = sharded.shard_rows(n, world, rank)
        cap = sharded.required_capacity(n, world)
        acc = o3.accel(arr["pos"], arr["mass_f64"], G, b, e)
        p_loc, v_loc = o3.integrate(arr["pos"][b:e], arr["vel"][b:e], acc, dt)
        pos_new, vel_new = torch.zeros(3, cap, dtype=torch.float64), torch.zeros(3, cap, dtype=torch.float64)
        xx_new = torch.zeros(cap, dtype=torch.float64)
        pos_new[:, b:e] = torch.from_numpy(np.ascontiguousarray(p_loc.T))
        vel_new[:, b:e] = torch.from_numpy(np.ascontiguousarray(v_loc.T))
        xx_new[b:e] = torch.from_numpy((p_loc[:, 0] ** 2 + p_loc[:, 2] ** 2) + p_loc[:, 1] ** 2)
        planes = [(t[d], 1) for t in (pos_new, vel_new) for d in range(3)] + [(xx_new, 1)]
        sharded.gather_planes(planes, per, rank, world)
        p_all = np.ascontiguousarray(pos_new[:, :n].numpy().T)
        pairs = o3.flags(p_all, arr["radius"], b, e)
        keys = torch.from_numpy((pairs[:, 0] << 32 | pairs[:, 1]).astype(np.int64))
        xcap = 1 << 14
        local = torch.zeros(xcap, dtype=torch.int64)
        local[: len(keys)] = keys
        xchg = torch.zeros(world, 1 + xcap, dtype=torch.int64)
        sharded.exchange_pairs(local, torch.tensor(len(keys)), xchg, rank, world)
        merged = torch.cat([xchg[r, 1:1 + int(xchg[r, 0])] for r in range(world)]).numpy()
        if rank == 0:
            out_q.put((p_all.copy(), np.sort(merged)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_exchange_equals_unsharded_oracle_3d():
    from cosmosim_b200 import scenes
    from oracle import oracle3d as o3
    n, world = 1500, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker3d, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    p_all, merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    arr = scenes.disk3d_arrays(n, seed=4, r_in=0.1, r_out=0.2, density=10.0)
    acc = o3.accel(arr["pos"], arr["mass_f64"], 6.674e-11)
    p_ref, _ = o3.integrate(arr["pos"], arr["vel"], acc, 600.0)
    assert np.array_equal(p_all, p_ref)
    pairs = o3.flags(p_ref, arr["radius"])
    ref = np.sort((pairs[:, 0] << 32 | pairs[:, 1]).astype(np.int64))
    assert len(ref) > 10 and np.array_equal(merged, ref)
