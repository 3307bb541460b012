"""GPU probe: pipe micro-benchmarks + per-stage timings of the frame at several N.
Writes JSON lines to stdout; used while developing (bench.py is the contract benchmark)."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmosim_b200 import _abi, scenes  # noqa: E402
from cosmosim_b200.engine import Engine  # noqa: E402

KINDS = {0: "FFMA", 1: "FFMA2", 2: "DFMA", 3: "MUFU.RSQ", 4: "DADD", 5: "DMUL", 6: "FMUL", 7: "FADD", 8: "MUFU.RSQ64H"}


def microbench(L, sm_count):
    out = torch.zeros(sm_count * 8 * 1024, dtype=torch.float64, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    for kind, name in KINDS.items():
        best = 0.0
        for threads, bps in ((256, 4), (512, 2), (1024, 1), (1024, 2)):
            blocks = sm_count * bps
            ops = C.c_int64()
            iters = 2000 if kind in (2, 4, 5, 3, 8) else 8000
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _abi.check(L.cosmo_microbench(kind, blocks, threads, iters, C.c_void_p(out.data_ptr()), C.byref(ops), stream))
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                rate = ops.value * blocks * threads / (ms * 1e-3)   # thread-instructions / s
                best = max(best, rate)
        res[name] = best
    return res


def time_stages(eng, dt, G, reps, **kw):
    L = eng.lib
    p = eng.params(dt, G, 100 * 1.496e11, **kw)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    st, sc, pp = C.byref(eng.st), C.byref(eng.sc), C.byref(p)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4 * reps + 1)]
    tA = tB = tC = 0.0
    for r in range(reps + 2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        _abi.check(L.cosmo_begin_frame(st, sc, pp, stream)); _abi.check(L.cosmo_accel_integrate(st, sc, pp, stream))
        e[1].record()
        _abi.check(L.cosmo_detect_pairs(st, sc, pp, stream))
        e[2].record()
        _abi.check(L.cosmo_resolve_commit(st, sc, pp, stream))
        e[3].record()
        torch.cuda.synchronize()
        if r >= 2:
            tA += e[0].elapsed_time(e[1]); tB += e[1].elapsed_time(e[2]); tC += e[2].elapsed_time(e[3])
    return tA / reps, tB / reps, tC / reps, p.jsplit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1001,16384,65536")
    ap.add_argument("--f32-sizes", default="65536,262144")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-micro", action="store_true")
    a = ap.parse_args()
    L = _abi.lib()
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    if not a.no_micro:
        mb = microbench(L, sm)
        print(json.dumps({"microbench_thread_instr_per_s": mb, "sm_count": sm}), flush=True)
    DT, G = 1e6 / 33, 6.674e-11
    for n in [int(x) for x in a.sizes.split(",") if x]:
        eng = Engine(n); eng.upload(scenes.star_disk_arrays(n, seed=1))
        for js in (None,):
            tA, tB, tC, jsplit = time_stages(eng, DT, G, a.reps, jsplit=js)
            print(json.dumps({"n": n, "prec": "fp64", "jsplit": jsplit, "ms_accel": tA, "ms_detect": tB, "ms_resolve": tC,
                              "accel_inter_per_s": n * n / (tA * 1e-3), "status": eng.status()}), flush=True)
    for n in [int(x) for x in a.f32_sizes.split(",") if x]:
        for scalar in (False, True):
            eng = Engine(n); eng.upload(scenes.star_disk_arrays(n, seed=1))
            tA, tB, tC, jsplit = time_stages(eng, DT, G, a.reps, fp32=True, f32_scalar=scalar, collisions=False)
            print(json.dumps({"n": n, "prec": "fp32", "scalar": scalar, "jsplit": jsplit, "ms_accel": tA, "ms_detect": tB,
                              "ms_resolve": tC, "accel_inter_per_s": n * n / (tA * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
