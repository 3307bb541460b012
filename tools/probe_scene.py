"""Where are the planets of the default bench scene after k frames?  (development)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cosmosim_b200 import scenes
from cosmosim_b200.engine import Engine
AU = 1.496e11; DT = 1e6 / 33; G = 6.674e-11
n = 1 << 20
arr = scenes.star_disk_arrays(n, seed=0)
eng = Engine(n); eng.upload(arr)
for frames in (1, 7, 16):
    eng.step(frames, DT, G, 100 * AU, fp32=True)
    d = eng.download()
    pos, rad = d["pos"], d["radius"]
    r = np.hypot(pos[:, 0], pos[:, 1])
    qs = [0.001, 0.01, 0.05, 0.1, 0.25, 0.5, 0.75, 0.9, 0.95, 0.99, 0.999]
    print("frame", d["status"]["frame"], "n", len(r))
    print(" |p|/AU quantiles", " ".join("%g:%.3g" % (q, v / AU) for q, v in zip(qs, np.quantile(r, qs))))
    print(" radius quantiles", " ".join("%g:%.3g" % (q, v) for q, v in zip(qs, np.quantile(rad, qs))), "max", rad.max())
    g = eng.grid_info()
    h = g["h"]
    ix = np.floor(pos / h).astype(np.int64)
    key = ix[:, 0] * (1 << 32) + ix[:, 1]
    u, cnt = np.unique(key, return_counts=True)
    top = np.sort(cnt)[::-1]
    print(" h=%.3g cells used %d; top occupancies %s; planets in cells with >14 occupants: %d; sum c^2/n = %.1f" % (
        h, len(u), top[:12], cnt[cnt > 14].sum(), (cnt.astype(float) ** 2).sum() / len(r)))
    big = u[np.argsort(cnt)[::-1][:5]]
    for k in big:
        print("   cell at (%.4g, %.4g) AU" % ((k >> 32) * h / AU, ((k & 0xffffffff) - (1 << 32 if (k & 0x80000000) else 0)) * h / AU))
    vel = d["vel"]; sp = np.hypot(vel[:, 0], vel[:, 1])
    print(" speed quantiles", " ".join("%g:%.3g" % (q, v) for q, v in zip(qs, np.quantile(sp, qs))))
