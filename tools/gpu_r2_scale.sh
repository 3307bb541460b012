#!/bin/bash
# round 2: bench at N GPUs only (run under gpurun --gpus N).  usage: gpu_r2_scale.sh <N> [steps] [warmup] [tag]
set -u
N=${1:-8}; K=${2:-20}; W=${3:-5}; TAG=${4:-a}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/sharded_check.py --bodies 50000 --frames 3 --fp32 > gpurun_out/r2_sharded_check_${N}gpu.log 2>&1; echo "check fp32 rc=$?"
grep "sharded==single" gpurun_out/r2_sharded_check_${N}gpu.log | tail -4
timeout 900 $TR bench.py --gpus $N --steps $K --warmup $W > gpurun_out/r2_bench_${N}gpu_$TAG.json 2> gpurun_out/r2_bench_${N}gpu_$TAG.err; echo "bench $N rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_${N}gpu_$TAG.json') if l.startswith('{')][-1])
    r=d['roofline']
    print('N=%d value %.4g ms/step %.3f frac %.3f outside %.3f digest %s identical %s' % (d['n_gpus'], d['value'], d['ms_per_step'], r['frac'], r['outside_stage_timers_ms'], d['state_digest'], d['replicas_identical']))
    print('   stage', r['stage_ms'], 'xchg', r['exchange_ms'])
    print('   full_n', d['full_n'] and (d['full_n']['ms_per_step'], d['full_n']['stage_ms']))
    print('   e2e', d['e2e'] and d['e2e']['ms_per_step'], d['final_state'])
except Exception as e:
    print('parse failed', e)
PY
