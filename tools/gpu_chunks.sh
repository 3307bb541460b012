#!/bin/bash
set -u
for c in 0 2 4 8 16; do
  COSMO_SYM_CHUNK=$c timeout 300 python bench.py --workload star_disk_64k_fp64 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/ck_$c.json 2>gpurun_out/ck_$c.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ck_$c.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("f64 chunk $c accel_ms %.3f"%r["kernel_ms"])
except Exception as e: print("chunk $c failed", e)
PY
done
for c in 0 8 16 32; do
  COSMO_SYM_CHUNK=$c timeout 300 python bench.py --workload cloud_256k_fp32 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/ck32_$c.json 2>gpurun_out/ck32_$c.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ck32_$c.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("f32 256k chunk $c accel_ms %.3f"%r["kernel_ms"])
except Exception as e: print("chunk $c failed", e)
PY
done
