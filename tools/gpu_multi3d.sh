#!/bin/bash
# 3-D step on NG GPUs: sharded == single parity, then bench lines
set -u
NG=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $RUN tools/sharded3d_check.py --steps 3 > gpurun_out/sharded3d_check_$NG.log 2>&1; echo "sharded3d check rc=$?"
grep -E "sharded==single|mismatch|ALL OK|FAILED|Error|error" gpurun_out/sharded3d_check_$NG.log | tail -20
for w in state3d_64k_fp64 state3d_256k_fp32; do
  timeout 600 $RUN bench.py --gpus $NG --workload $w --no-cpu-baseline > gpurun_out/bench_${w}_${NG}gpu.json 2> gpurun_out/bench_${w}_${NG}gpu.err; echo "$w rc=$?"
  tail -n 1 gpurun_out/bench_${w}_${NG}gpu.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
    print(d['config']['workload'], 'gpus', d['n_gpus'], 'ms/step %.3f'%d['ms_per_step'], 'value %.3e'%d['value'], 'kernel', r['kernel'], 'frac', r['frac'], r['stage_ms'], 'e2e', (d.get('e2e') or {}).get('value'))
except Exception as e: print('unreadable', e)
"
  tail -n 5 gpurun_out/bench_${w}_${NG}gpu.err
done
