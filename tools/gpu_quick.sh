#!/bin/bash
# tests + the contract bench on the main workloads (no profiler)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for w in star_disk_1m_fp32 star_disk_64k_fp64 cloud_256k_fp32; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
timeout 300 python bench.py --workload star_disk_1001_fp64 --steps 200 --no-cpu-baseline > gpurun_out/bench_star_disk_1001_fp64.json 2> gpurun_out/bench_1001.err; echo "1001 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    r=d.get('roofline') or {}
    print(f.split('/')[-1], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac', r.get('frac'), 'stage_ms', r.get('stage_ms'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('final_state'), d.get('detect_mode'), 'pairs', d.get('pairs_last_frame'), 'n/step', d.get('n_active_per_step'))
PY
tail -n 3 gpurun_out/bench_*.err
