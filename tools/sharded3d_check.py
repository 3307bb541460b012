"""Multi-GPU parity of the 3-D step: the row-sharded State.interact (one process per GPU, NCCL) must
leave the same absorb stream and discrete state as the single-GPU step, positions to rounding, and
identical replicas on every rank.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tools/sharded3d_check.py [--steps 3]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmosim_b200 import scenes  # noqa: E402
from cosmosim_b200.engine3d import Engine3D  # noqa: E402


def crowded(n, seed):
    arr = scenes.disk3d_arrays(n, seed=seed, r_in=0.1, r_out=0.3, thickness=0.01, density=30.0)
    return arr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    DT, G = 600.0, 6.674e-11
    ok = True
    # (n, fp32, symmetric kernel): one-sided gather path, all-pairs and grid detection, symmetric all-reduce path
    # n = 1000: the last rank's CONTIGUOUS row range is empty while it still owns symmetric row blocks
    for n, fp32, sym in ((1000, False, True), (1000, False, False), (3000, False, False), (20000, False, False), (20000, False, True), (40000, True, False),
                         (40000, True, True)):
        arr = crowded(n, 3)
        single = Engine3D(n, device="cuda:%d" % local)
        shard = Engine3D(n, device="cuda:%d" % local, group=dist.group.WORLD)
        single.upload(arr)
        shard.upload(arr)
        same, worst, replicas = True, 0.0, True
        for s in range(a.steps):
            single.step(DT, G, fp32=fp32, sym=sym)
            shard.step(DT, G, fp32=fp32, sym=sym)
            d1, d2 = single.download(), shard.download()
            e1, e2 = single.events(), shard.events()
            fok = all(np.array_equal(d1[k], d2[k]) for k in ("id", "mass", "mass_hi", "mass_lo", "density", "flags"))
            fok = fok and np.array_equal(e1, e2) and d1["status"]["error"] == 0 and d2["status"]["error"] == 0
            tol = 1e-12 if not fp32 else 1e-6      # FP32 sums: another split of the source range, FP32 rounding
            for k in ("pos", "vel"):
                if d1[k].shape != d2[k].shape:
                    fok = False
                    continue
                rel = np.abs(d1[k] - d2[k]).max() / max(1e-300, np.abs(d1[k]).max())
                worst = max(worst, rel)
                fok = fok and rel <= tol
            if not fok and rank == 0:
                print("  step %d mismatch: ids %s mass %s events %s (%d vs %d) err %d/%d pairs %d/%d worst %.2e" % (
                    s, np.array_equal(d1["id"], d2["id"]), np.array_equal(d1["mass"], d2["mass"]),
                    np.array_equal(e1, e2), len(e1), len(e2), d1["status"]["error"], d2["status"]["error"],
                    d1["status"]["n_pairs"], d2["status"]["n_pairs"], worst), flush=True)
            same = same and fok
            digest = torch.tensor([float(np.sum(d2["pos"])), float(np.sum(d2["vel"])), float(len(e2)),
                                   float(d2["status"]["n_active"])], dtype=torch.float64, device="cuda")
            gathered = [torch.zeros_like(digest) for _ in range(world)]
            dist.all_gather(gathered, digest)
            replicas = replicas and all(torch.equal(g, gathered[0]) for g in gathered)
            arr2 = {"pos": d1["pos"], "vel": d1["vel"], "mass_f64": d1["mass"], "mass_hi": d1["mass_hi"],
                    "mass_lo": d1["mass_lo"], "density": d1["density"], "radius": d1["radius"], "flags": d1["flags"],
                    "id": d1["id"]}
            shard.upload(arr2, iteration=s + 1)       # re-sync: n-body is chaotic
            single.upload(arr2, iteration=s + 1)
        same = same and replicas
        if rank == 0:
            print("3-D n=%d fp32=%s sym=%s steps=%d: absorb calls=%d objects=%d detect=%s sharded==single: %s "
                  "(worst rel %.1e) replicas_identical: %s" % (n, fp32, sym, a.steps, len(e1), d1["status"]["n_active"],
                                                               "grid" if shard.grid is not None else "all-pairs", same,
                                                               worst, replicas), flush=True)
        ok = ok and same
    if rank == 0:
        print("ALL OK" if ok else "FAILED", flush=True)
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
