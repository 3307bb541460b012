#!/bin/bash
set -u
mkdir -p gpurun_out
for w in star_disk_1m_fp32 star_disk_64k_fp64; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/b2_$w.json 2> gpurun_out/b2_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b2_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f.split('/')[-1], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac %.4f'%r['frac'], r['stage_ms'], d['final_state'], 'pairs', d['pairs_last_frame'], 'jsplit', d['jsplit'])
PY
tail -n 3 gpurun_out/b2_*.err
