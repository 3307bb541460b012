"""FP32 fast path: per-row relative acceleration error vs the FP64 oracle, near field on / off.
Not a test (pytest does not collect it): a measurement script, kept under tests/ because it uses the
oracle, which only tests / smoke() / the CPU leg of bench.py may do.  python tests/probe_fp32_err.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cosmosim_b200 import scenes
from cosmosim_b200.engine import Engine
from oracle import oracle as orc
AU = 1.496e11; DT = 1e6 / 33; G = 6.674e-11


def relerr(a, ref):
    return np.linalg.norm(a - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-300)


for name, arr, nrows in (("disk20k", scenes.star_disk_arrays(20000, seed=1), 20000), ("cloud20k", scenes.cloud_arrays(20000, seed=2), 20000),
                         ("disk1m", scenes.star_disk_arrays(1 << 20, seed=0), 4096)):
    n = arr["radius"].shape[0]
    rows = np.arange(n) if nrows >= n else np.sort(np.random.default_rng(0).choice(n, nrows, replace=False))
    # oracle on contiguous row blocks containing the sampled rows is wasteful; evaluate row by row in blocks of 64
    ref = np.zeros((len(rows), 2))
    if nrows >= n:
        ref = orc.accel(arr["pos"], arr["mass_f64"], G)
    else:
        blocks = np.unique(rows // 64)
        cache = {}
        for b in blocks:
            cache[b] = orc.accel(arr["pos"], arr["mass_f64"], G, int(b * 64), int(min(n, b * 64 + 64)))
        ref = np.stack([cache[r // 64][r % 64] for r in rows])
    for sym in (False, True):
        if n > 100000 and not sym:
            continue
        for near in (False, True):
            eng = Engine(n); eng.upload(arr)
            evs = []
            eng.step(1, DT, G, 100 * AU, collisions=False, store_acc=True, fp32=True, f32_sym=sym, near=near, stage_events=evs)
            torch.cuda.synchronize()
            print("   stage A %.3f ms" % evs[0][1].elapsed_time(evs[0][2]), end=" ")
            acc = eng.last_frame_tensors()["acc"][:n].cpu().numpy()[rows]
            eng.check_errors()
            rel = relerr(acc, ref)
            print("%-9s sym=%d near=%d  median %.2e  p99 %.2e  p99.9 %.2e  max %.2e  (rows %d)" % (
                name, sym, near, np.median(rel), np.quantile(rel, 0.99), np.quantile(rel, 0.999), rel.max(), len(rows)), flush=True)
            if near:
                w = np.argsort(rel)[-3:]
                pos = arr["pos"]
                for k in w:
                    i = rows[k]
                    d = np.hypot(*(pos - pos[i]).T); d[i] = np.inf
                    j = int(np.argmin(d))
                    print("     worst row %d rel %.2e |p|=%.3g AU nearest %.3g m (r_i+r_j=%.3g) Rc=%.3g" % (
                        i, rel[k], np.hypot(*pos[i]) / AU, d[j], arr["radius"][i] + arr["radius"][j], np.abs(pos[[i, j]]).max() / 512))
            del eng
