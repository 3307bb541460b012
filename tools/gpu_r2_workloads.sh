#!/bin/bash
# every bench workload once (short), to make sure all lines of bench.py still print (development)
set -u
mkdir -p gpurun_out
for w in star_disk_64k_fp64 cloud_256k_fp32 star_disk_1001_fp64 state3d_64k_fp64 state3d_10k_fp64 state3d_256k_fp32; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 2 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_$w.json') if l.startswith('{')][-1])
    r=d['roofline']
    print('$w rc=$rc value %.3e ms/step %.3f frac %s kernel %s e2e_ms %s stage %s' % (d['value'], d['ms_per_step'], r and round(r['frac'],3), r and r['kernel'], d['e2e'] and round(d['e2e']['ms_per_step'],3), r and r['stage_ms']))
except Exception as e:
    print('$w rc=$rc parse failed', e); print(open('gpurun_out/r2_bench_$w.err').read()[-600:])
PY
done
