"""Multi-GPU parity: the row-sharded frame (one process per GPU, NCCL) must leave bit-identical
state and merge stream as the single-GPU frame.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/sharded_check.py [--bodies 20000] [--frames 4] [--fp32]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmosim_b200 import scenes  # noqa: E402
from cosmosim_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bodies", dest="n", type=int, default=20000)
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--fp32", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    DT, G, RMAX = 1e6 / 33, 6.674e-11, 100 * 1.496e11
    ok = True
    for kind, n in (("disk", a.n), ("cloud", a.n), ("disk", 1001)):
        arr = scenes.star_disk_arrays(n, seed=2) if kind == "disk" else scenes.cloud_arrays(n, seed=2)
        if kind == "cloud":
            arr["radius"] = arr["radius"] * 40.0
        single = Engine(n, device="cuda:%d" % local)
        single.upload(arr)
        shard = Engine(n, device="cuda:%d" % local, group=dist.group.WORLD)
        shard.upload(arr)
        same, worst, replicas = True, 0.0, True
        for f in range(a.frames):
            single.step(1, DT, G, RMAX, fp32=a.fp32)
            shard.step(1, DT, G, RMAX, fp32=a.fp32)
            d1, d2 = single.download(), shard.download()
            e1, e2 = single.events(), shard.events()
            # discrete state and the merge stream: bit-exact.  pos/vel: the FP64 all-pairs sum is
            # split over a different number of source ranges when the rows are sharded (another,
            # equally valid summation order), so they agree to rounding per frame, not bitwise.
            fok = all(np.array_equal(d1[k], d2[k]) for k in ("id", "mass", "radius", "volume", "eaten"))
            fok = fok and np.array_equal(e1, e2) and d1["status"]["error"] == 0 and d2["status"]["error"] == 0
            for k in ("pos", "vel"):
                if d1[k].shape != d2[k].shape:
                    fok = False
                    continue
                rel = np.abs(d1[k] - d2[k]).max() / max(1e-300, np.abs(d1[k]).max())
                worst = max(worst, rel)
                fok = fok and rel <= 1e-12
            if not fok and rank == 0:
                print("  frame %d mismatch: ids %s mass %s events %s (%d vs %d) err %d/%d pairs %d/%d" % (
                    f, np.array_equal(d1["id"], d2["id"]), np.array_equal(d1["mass"], d2["mass"]),
                    np.array_equal(e1, e2), len(e1), len(e2), d1["status"]["error"], d2["status"]["error"],
                    d1["status"]["n_pairs"], d2["status"]["n_pairs"]), flush=True)
            same = same and fok
            # every rank must hold the same replica of the sharded result, bit for bit
            digest = torch.tensor([float(np.sum(d2["pos"])), float(np.sum(d2["vel"])), float(len(e2)),
                                   float(d2["status"]["n_active"])], dtype=torch.float64, device="cuda")
            gathered = [torch.zeros_like(digest) for _ in range(world)]
            dist.all_gather(gathered, digest)
            replicas = replicas and all(torch.equal(g, gathered[0]) for g in gathered)
            # re-sync: the next frame starts from identical bits on both engines (n-body is chaotic)
            arr2 = {"pos": d1["pos"], "vel": d1["vel"], "mass_f64": d1["mass"], "mass_hi": d1["mass_hi"],
                    "mass_lo": d1["mass_lo"], "flags": d1["flags"], "radius": d1["radius"], "volume": d1["volume"],
                    "eaten": d1["eaten"], "id": d1["id"]}
            shard.upload(arr2, frame=f + 1)
            single.upload(arr2, frame=f + 1)
        same = same and replicas
        if rank == 0:
            print("%s n=%d fp32=%s frames=%d: merges=%d active=%d sharded==single: %s (worst rel %.1e) replicas_identical: %s exchange: %s" %
                  (kind, n, a.fp32, a.frames, len(e1), d1["status"]["n_active"], same, worst, replicas,
                   "peer stores" if getattr(shard, "_peer", None) else "nccl"), flush=True)
        ok = ok and same
    if rank == 0:
        print("exchange:", "peer stores" if getattr(shard, "_peer", None) else "nccl (%s)" % getattr(shard, "_peer_error", "-"), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
