"""Stage-A time of frame 0 of the default bench scene, fresh upload each repetition (development)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cosmosim_b200 import scenes
from cosmosim_b200.engine import Engine
AU = 1.496e11; DT = 1e6 / 33; G = 6.674e-11
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
arr = scenes.star_disk_arrays(n, seed=0)
eng = Engine(n)
ts = []
for rep in range(4):
    eng.upload(arr)
    evs = []
    eng.step(1, DT, G, 100 * AU, stage_events=evs, fp32=True, collisions=False, near=False)
    torch.cuda.synchronize()
    ts.append(evs[0][1].elapsed_time(evs[0][2]))
print("stage A ms:", " ".join("%.3f" % t for t in ts), " best %.3f" % min(ts))
