#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for w in star_disk_1m_fp32 cloud_256k_fp32; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/b3_$w.json 2> gpurun_out/b3_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b3_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.4f'%r['frac'], r['stage_ms'], d['final_state'], 'pairs', d['pairs_last_frame'])
    except Exception as e: print(f, e)
PY
tail -n 5 gpurun_out/b3_*.err
