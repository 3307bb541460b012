#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "fp32_fast or dense_clusters or 65536" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -8 gpurun_out/pytest_gpu2.log
timeout 600 python bench.py --workload star_disk_1m_fp32 > gpurun_out/bench_star_disk_1m_fp32.json 2> gpurun_out/bench_star_disk_1m_fp32.err; echo "1m rc=$?"
tail -n 3 gpurun_out/bench_star_disk_1m_fp32.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_star_disk_1m_fp32.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.3e'%d['value'], 'ms/step', d['ms_per_step'], 'frac', r['frac'], r['stage_ms'], 'e2e', d['e2e'], 'cpu', d['cpu_baseline'], d['final_state'], d['n_active_per_step'], d['pairs_last_frame'])
PY
