import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmosim_b200 import _abi, scenes
from cosmosim_b200.engine import Engine
DT, G, AU = 1e6/33, 6.674e-11, 1.496e11
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
arr = scenes.star_disk_arrays(n, seed=0)
eng = Engine(n); eng.upload(arr)
print("grid (h, ox, oy, c, rg):", eng.grid)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
st, sc = C.byref(eng.st), C.byref(eng.sc)
for frame in range(6):
    p = eng.params(DT, G, 100*AU, fp32=True)
    pp = C.byref(p)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    _abi.check(eng.lib.cosmo_begin_frame(st, sc, pp, stream)); _abi.check(eng.lib.cosmo_accel_integrate(st, sc, pp, stream))
    ev[0].record()
    _abi.check(eng.lib.cosmo_detect_pairs(st, sc, pp, stream))
    ev[1].record()
    torch.cuda.synchronize()
    s = eng.t["status"].cpu().numpy()
    tot = int(eng.t["cell_start"][p.cells_per_side**2].item())
    cs = eng.t["cell_start"][:p.cells_per_side**2+1].cpu().numpy()
    pop = np.diff(cs)
    b = eng.t["chunk_alive"][:tot//1024+2].cpu().numpy()
    _abi.check(eng.lib.cosmo_resolve_commit(st, sc, pp, stream)); torch.cuda.synchronize()
    eng.retune(); torch.cuda.synchronize()
    print("frame", frame, "n", s[0], "detect ms %.3f"%ev[0].elapsed_time(ev[1]), "pairs", s[3], "cand", s[7], "giants", s[8], "max cell pop", pop.max(), "cells", p.cells_per_side, "h %.3e"%p.cell_size, "chunk sizes max", np.diff(b).max(), "min", np.diff(b).min())
