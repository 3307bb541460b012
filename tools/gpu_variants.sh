#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 0 1 2 3 4; do
  for w in star_disk_1m_fp32 cloud_256k_fp32; do
    COSMO_SYM_VARIANT=$v timeout 300 python bench.py --workload $w --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/var_${v}_$w.json 2> gpurun_out/var_${v}_$w.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/var_${v}_$w.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('variant $v $w accel_ms %.3f rate %.3e'%(r['kernel_ms'], r['interactions_per_s_kernel']))
except Exception as e: print('variant $v $w failed', e)
PY
  done
done
