#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload star_disk_64k_fp64 --no-cpu-baseline --no-e2e > gpurun_out/bench_64k_b.json 2> gpurun_out/bench_64k_b.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_64k_b.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.3e'%d['value'], 'ms/step', d['ms_per_step'], 'frac', r['frac'], r['stage_ms'], 'jsplit', d['jsplit'])
PY
