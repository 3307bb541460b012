"""Stage-A time of the symmetric 3-D FP64 kernel for several column-chunk sizes."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from cosmosim_b200 import engine3d  # noqa: E402
from probe_state3d import cloud, timed  # noqa: E402

for n in (32768, 65536, 131072):
    eng = engine3d.Engine3D(n)
    eng.upload(cloud(n))
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for chunk in (0, 1, 2, 3, 4, 6, 8, 16):
        if chunk:
            os.environ["COSMO3_SYM_CHUNK"] = str(chunk)
        else:
            os.environ.pop("COSMO3_SYM_CHUNK", None)
        p = eng.params(60.0, 6.674e-11, True, False)
        t = timed(lambda: eng.lib.cosmo3_accel_integrate(C.byref(eng.st), C.byref(eng.sc), C.byref(p), sp), 10)
        print("n=%d chunk=%d (used %d) A %.3f ms  %.3e int/s  frac of 1.845e13 DP instr/s: %.3f"
              % (n, chunk, p.sym_chunk, t, n * n / t * 1e3, n * n * 14.5 / t * 1e3 / 1.845e13), flush=True)
