"""Packed-FP32 pipe probes (cosmo_microbench kinds 0, 1, 9-12): what the FMA pipe sustains for operand
patterns like those of k_accel_f32_sym, at the kernel's own occupancy (16 warps per SM) and at 32."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cosmosim_b200 import _abi

L = _abi.lib()
dev = torch.device("cuda", 0)
sm = torch.cuda.get_device_properties(dev).multi_processor_count
out = torch.zeros(sm * 8 * 1024, dtype=torch.float64, device=dev)
stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
NAMES = {0: "FFMA scalar, constant operands", 1: "FFMA2, constant operands", 9: "FFMA2, three changing operands",
         10: "FFMA2, two changing operands + constant", 11: "pair chain of the symmetric kernel (registers only)",
         12: "pair chain without the MUFUs", 2: "DFMA, constant operands", 13: "DFMA, three changing operands",
         14: "DFMA, two changing operands + constant"}
for kind, iters in ((0, 4000), (1, 4000), (9, 4000), (10, 4000), (11, 4000), (12, 4000), (2, 2000), (13, 2000), (14, 2000)):
    for bps in (2, 4):
        best = 0.0
        ops = C.c_int64()
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _abi.check(L.cosmo_microbench(kind, sm * bps, 256, iters, C.c_void_p(out.data_ptr()), C.byref(ops), stream),
                       "cosmo_microbench")
            e1.record()
            torch.cuda.synchronize()
            best = max(best, ops.value * sm * bps * 256 / (e0.elapsed_time(e1) * 1e-3))
        lanes = 1 if kind in (0, 2, 13, 14) else 2
        print("kind %2d  %d CTAs of 256 per SM: %.3e instr/s = %.2f TFLOP/s  (%s)" % (kind, bps, best, best * lanes * 2 / 1e12, NAMES[kind]))
