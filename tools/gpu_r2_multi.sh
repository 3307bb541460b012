#!/bin/bash
# round 2: N-GPU pass.  usage: gpu_r2_multi.sh <N> [steps] [warmup]   (run under gpurun --gpus N)
set -u
N=${1:-2}; K=${2:-20}; W=${3:-5}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/sharded_check.py --bodies 20000 --frames 4 > gpurun_out/r2_sharded_check_${N}gpu.log 2>&1; echo "check fp64 rc=$?"
timeout 600 $TR tools/sharded_check.py --bodies 50000 --frames 4 --fp32 >> gpurun_out/r2_sharded_check_${N}gpu.log 2>&1; echo "check fp32 rc=$?"
timeout 600 $TR tools/sharded3d_check.py >> gpurun_out/r2_sharded3d_check_${N}gpu.log 2>&1; echo "check 3d rc=$?"
grep -v "^W\|^\[W\|warn" gpurun_out/r2_sharded_check_${N}gpu.log | tail -8
tail -4 gpurun_out/r2_sharded3d_check_${N}gpu.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then CMD="python bench.py --gpus 1 --steps $K --warmup $W --no-cpu-baseline"; else CMD="$TR bench.py --gpus $n --steps $K --warmup $W"; fi
  timeout 900 $CMD > gpurun_out/r2_bench_${n}of${N}gpu.json 2> gpurun_out/r2_bench_${n}of${N}gpu.err; echo "bench $n rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_${n}of${N}gpu.json') if l.startswith('{')][-1])
    r=d['roofline']
    print('N=%d value %.4g ms/step %.3f frac %.3f outside %.3f digest %s identical %s' % (d['n_gpus'], d['value'], d['ms_per_step'], r['frac'], r['outside_stage_timers_ms'], d['state_digest'], d['replicas_identical']))
    print('   stage', r['stage_ms'], 'xchg', r['exchange_ms'])
    print('   full_n', d['full_n'] and (d['full_n']['ms_per_step'], d['full_n']['stage_ms']))
    print('   e2e', d['e2e'] and d['e2e']['ms_per_step'])
except Exception as e:
    print('parse failed', e)
PY
done
