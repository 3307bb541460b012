#!/bin/bash
# end-of-round verification on one GPU: full GPU suite, the contract bench (default workload), the other lines
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"
for w in star_disk_64k_fp64 state3d_64k_fp64 state3d_10k_fp64 state3d_256k_fp32; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json,glob
for f in ['gpurun_out/bench_default.json','gpurun_out/bench_reference.json']+sorted(glob.glob('gpurun_out/bench_st*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    r=d.get('roofline') or {}
    print(f.split('/')[-1], d.get('impl','b200'), 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel', r.get('kernel'), 'frac', r.get('frac'), 'stage_ms', r.get('stage_ms'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'launches', d.get('gpu_launches'), (d.get('clocks') or {}).get('reasons'))
PY
tail -n 2 gpurun_out/bench_*.err | grep -v "^$" | tail -20
