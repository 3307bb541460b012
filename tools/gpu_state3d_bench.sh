#!/bin/bash
# bench lines of the 3-D step (no profiler)
set -u
mkdir -p gpurun_out
for w in state3d_64k_fp64 state3d_10k_fp64 state3d_256k_fp32; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_state3d_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    r=d.get('roofline') or {}
    print(f.split('/')[-1], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac', r.get('frac'), 'stage_ms', r.get('stage_ms'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('final_state'), 'pairs', d.get('pairs_last_step'), 'launches', d.get('gpu_launches'), d.get('clocks'))
PY
tail -n 3 gpurun_out/bench_state3d_*.err
