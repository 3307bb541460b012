"""Where does a sharded FP32 frame differ from the single-GPU frame, bit for bit?  (development)
    python -m torch.distributed.run --nproc-per-node 2 ... tools/probe_shard_bits.py [n] [frames]"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmosim_b200 import scenes
from cosmosim_b200.engine import Engine
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
DT, G, RMAX = 1e6 / 33, 6.674e-11, 100 * 1.496e11
arr = scenes.star_disk_arrays(n, seed=0)
for near in (False, True):
    single = Engine(n, device="cuda:%d" % local); single.upload(arr)
    shard = Engine(n, device="cuda:%d" % local, group=dist.group.WORLD); shard.upload(arr)
    for f in range(frames):
        m = single.status()["n_active"]
        single.step(1, DT, G, RMAX, fp32=True, near=near)
        shard.step(1, DT, G, RMAX, fp32=True, near=near)
        p1 = single._pos_new.view(-1, 2)[:m].cpu().numpy(); p2 = shard._pos_new.view(-1, 2)[:m].cpu().numpy()
        diff = np.nonzero((p1 != p2).any(axis=1))[0]
        d1, d2 = single.download(), shard.download()
        e1, e2 = single.events(), shard.events()
        if rank == 0:
            rel = (np.abs(p1 - p2).max() / np.abs(p1).max()) if len(diff) else 0.0
            print("near=%d frame %d n=%d: rows with different pos_new bits %d (max rel %.1e); events equal %s; ids equal %s" % (
                near, f, m, len(diff), rel, np.array_equal(e1, e2), np.array_equal(d1["id"], d2["id"])), flush=True)
            if len(diff):
                k = int(diff[0])
                print("    first row %d: single %r sharded %r" % (k, p1[k].tolist(), p2[k].tolist()), flush=True)
        a2 = {"pos": d1["pos"], "vel": d1["vel"], "mass_f64": d1["mass"], "mass_hi": d1["mass_hi"], "mass_lo": d1["mass_lo"],
              "flags": d1["flags"], "radius": d1["radius"], "volume": d1["volume"], "eaten": d1["eaten"], "id": d1["id"]}
        shard.upload(a2, frame=f + 1); single.upload(a2, frame=f + 1)
    del single, shard
    torch.cuda.empty_cache()
dist.destroy_process_group()
