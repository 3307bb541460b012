"""Per-frame view of the device-tuned broad-phase grid on the default bench scene (development)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cosmosim_b200 import scenes, _abi
from cosmosim_b200.engine import Engine
AU = 1.496e11; DT = 1e6 / 33; G = 6.674e-11
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 25
arr = scenes.star_disk_arrays(n, seed=0)
eng = Engine(n); eng.upload(arr)
for f in range(frames):
    evs = []
    eng.step(1, DT, G, 100 * AU, stage_events=evs, fp32=True)
    torch.cuda.synchronize()
    ev = evs[0]
    s = eng.t["status"].cpu().numpy()
    g = eng.grid
    gi = eng.grid_info()
    rad = eng.t["radius"][: int(s[0])]
    print("f=%2d n=%7d A=%7.2f B=%6.3f C=%6.3f  h=%.3e c=%4d rg=%.3e giants=%d cand/n=%.1f heavy=%d pairs=%d rmax=%.3e r99.9=%.3e box=%.3e rho=%.3e" % (
        f, s[0], ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[3].elapsed_time(ev[4]), g[0], g[3], g[4],
        s[_abi.ST_N_GIANTS], s[_abi.ST_N_CAND] / max(1, s[0]), s[_abi.ST_N_HEAVY], s[_abi.ST_N_PAIRS],
        float(rad.max()), float(torch.quantile(rad[:: max(1, rad.numel() // 100000)], 0.999)), gi["box"], gi["rho"]), flush=True)
