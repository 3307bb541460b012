"""Quick timing of the 3-D step's stages (CUDA events on the launching stream)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from cosmosim_b200 import engine3d  # noqa: E402


def cloud(n, seed=0):
    rng = np.random.default_rng(seed)
    AU, ME, MS = 1.496e11, 5.972e24, 1.989e30
    pos = rng.normal(size=(n, 3)) * AU * np.array([1.0, 0.7, 0.2])
    vel = rng.normal(size=(n, 3)) * 1e4
    mass = rng.uniform(0.1, 100.0, n) * ME
    mass[0] = MS
    dens = np.full(n, 3000.0)
    radius = ((3 * (mass / dens)) / (4 * np.pi)) ** (1 / 3)
    return {"pos": pos, "vel": vel, "mass_f64": mass, "mass_hi": np.zeros(n, np.uint64), "mass_lo": np.zeros(n, np.uint64),
            "density": dens, "radius": radius, "flags": np.zeros(n, np.uint8), "id": np.arange(n, dtype=np.int32)}


def timed(fn, reps):
    s = torch.cuda.current_stream()
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(reps):
        fn()
    b.record(s)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if __name__ == "__main__":
  for n, fp32, reps in ((1001, False, 200), (10001, False, 50), (65536, False, 10), (65536, True, 20), (262144, True, 5)):
      eng = engine3d.Engine3D(n)
      eng.upload(cloud(n))
      L, st, sc = eng.lib, eng.st, eng.sc
      sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
      p = eng.params(60.0, 6.674e-11, True, fp32)
      ta = timed(lambda: L.cosmo3_accel_integrate(C.byref(st), C.byref(sc), C.byref(p), sp), reps)
      tb = timed(lambda: L.cosmo3_detect_pairs(C.byref(st), C.byref(sc), C.byref(p), sp), reps)
      eng.t["status"][3] = 0
      tc = timed(lambda: L.cosmo3_resolve_commit(C.byref(st), C.byref(sc), C.byref(p), sp), reps)
      eng.upload(cloud(n))
      tall = timed(lambda: eng.step(60.0, 6.674e-11, True, fp32), reps)
      if n <= 16384:
          eng.step(60.0, 6.674e-11, True, fp32, steps=64)      # captures the graph
          tg = timed(lambda: eng.step(60.0, 6.674e-11, True, fp32, steps=64), 4) / 64
          print("n=%d graph replay: %.4f ms per step" % (n, tg))
      ops = (8.0 if n >= 32768 else 12.0) if fp32 else (14.5 if n >= 16384 else 25.0)
      print("n=%d fp32=%s jsplit=%d/%d  A %.3f ms (%.3e int/s, %.2e pipe-instr/s)  B %.3f ms  C %.3f ms  step %.3f ms  status %s"
            % (n, fp32, p.jsplit, p.jsplit_detect, ta, n * n / ta * 1e3, n * n * ops / ta * 1e3, tb, tc, tall, eng.check_errors()))
