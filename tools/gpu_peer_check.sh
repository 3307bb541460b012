#!/bin/bash
# experimental peer-store exchange: short, strictly time-limited check + bench on NG GPUs
set -u
NG=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29577"
COSMO_PEER_EXCHANGE=1 timeout 150 $RUN tools/sharded_check.py --bodies 70000 --frames 3 --fp32 > gpurun_out/peer_check_$NG.log 2>&1; echo "peer fp32 rc=$?"
COSMO_PEER_EXCHANGE=1 timeout 150 $RUN tools/sharded_check.py --bodies 40000 --frames 3 >> gpurun_out/peer_check_$NG.log 2>&1; echo "peer fp64 rc=$?"
grep -E "sharded==single|mismatch|Error" gpurun_out/peer_check_$NG.log | tail -12
for mode in 1 0; do
COSMO_PEER_EXCHANGE=$mode timeout 200 $RUN bench.py --gpus $NG --steps 5 --warmup 3 --no-e2e > gpurun_out/peer_bench_${mode}_g$NG.json 2> gpurun_out/peer_bench_${mode}_g$NG.err; echo "bench mode $mode rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/peer_bench_${mode}_g$NG.json").read().strip().splitlines()[-1])
    print('mode $mode', 'ms/step %.3f'%d['ms_per_step'], d['roofline']['stage_ms'], d['exchange'], d['final_state'])
except Exception as e: print('mode $mode failed', e)
PY
done
