import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cosmosim_b200 import _abi, scenes, sharded
from cosmosim_b200.engine import Engine
from test_gpu_parity import _detect_pairs
DT, G, AU = 1e6/33, 6.674e-11, 1.496e11
n = 20000
arr = scenes.cloud_arrays(n, seed=2); arr["radius"] = arr["radius"] * 40.0
eng = Engine(n); eng.upload(arr)
print("grid", eng.grid)
p = eng.params(DT, G, 100*AU)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
st, sc = C.byref(eng.st), C.byref(eng.sc)
_abi.check(eng.lib.cosmo_begin_frame(st, sc, C.byref(p), stream)); _abi.check(eng.lib.cosmo_accel_integrate(st, sc, C.byref(p), stream))
full = _detect_pairs(eng, p)
ng = int(eng.t["status"][_abi.ST_N_GIANTS].item()); giants = set(eng.t["giants"][:ng].cpu().numpy().tolist())
print("full", len(full), "giants", ng)
parts = []
for r in range(2):
    q = eng.params(DT, G, 100*AU); q.i_begin, q.i_end, _ = sharded.shard_rows(n, 2, r); q.shard_index, q.shard_count = r, 2
    parts.append(_detect_pairs(eng, q)); print("rank", r, "rows", q.i_begin, q.i_end, "pairs", len(parts[-1]), "giants now", int(eng.t["status"][_abi.ST_N_GIANTS].item()))
u, cnt = np.unique(np.concatenate(parts), return_counts=True)
dup = u[cnt > 1]
print("dups", len(dup))
for k in dup[:10]:
    i, j = int(k >> 32), int(k & 0xffffffff)
    print(i, j, "i giant" if i in giants else "", "j giant" if j in giants else "", "in0", k in parts[0], "in1", k in parts[1], "dup within part0", (parts[0]==k).sum(), (parts[1]==k).sum())
missing = np.setdiff1d(full, u); extra = np.setdiff1d(u, full)
print("missing", len(missing), "extra", len(extra))
