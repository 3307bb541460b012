#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "fp32 or sharded_detection or energies" 2>&1 | tail -3
for v in 0 1; do
  for w in star_disk_1m_fp32 cloud_256k_fp32; do
    COSMO_SYM_VARIANT=$v timeout 300 python bench.py --workload $w --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/var_${v}_$w.json 2> gpurun_out/var_${v}_$w.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/var_${v}_$w.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('variant $v $w accel_ms %.3f rate %.3e frac %.4f launches %d'%(r['kernel_ms'], r['interactions_per_s_kernel'], r['frac'], d['gpu_launches']))
except Exception as e: print('variant $v $w failed', e)
PY
  done
done
