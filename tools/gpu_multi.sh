#!/bin/bash
# multi-GPU validation: sharded == single parity, then the contract bench under torchrun
set -u
NG=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $RUN tools/sharded_check.py --bodies 20000 --frames 4 > gpurun_out/sharded_check_$NG.log 2>&1; echo "sharded fp64 rc=$?"
timeout 600 $RUN tools/sharded_check.py --bodies 70000 --frames 3 --fp32 >> gpurun_out/sharded_check_$NG.log 2>&1; echo "sharded fp32 rc=$?"
grep -E "sharded==single|Error|error" gpurun_out/sharded_check_$NG.log | tail -12
timeout 900 $RUN bench.py --gpus $NG --workload star_disk_1m_fp32 > gpurun_out/bench_1m_fp32_g$NG.json 2> gpurun_out/bench_1m_fp32_g$NG.err; echo "bench rc=$?"
#timeout 900 $RUN bench.py --gpus $NG --workload cloud_256k_fp32 > gpurun_out/bench_256k_fp32_g$NG.json 2> gpurun_out/bench_256k_fp32_g$NG.err; echo "bench rc=$?"
tail -n 4 gpurun_out/bench_1m_fp32_g$NG.err
python - <<PY
import json
for f in ("gpurun_out/bench_1m_fp32_g$NG.json","gpurun_out/bench_256k_fp32_g$NG.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e'%d['value'], 'ms/step', d['ms_per_step'], 'n_gpus', d['n_gpus'], 'e2e', (d.get('e2e') or {}).get('value'), d['final_state'], d['clocks'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_1m_fp32_g$NG.json").read().strip().splitlines()[-1])
print(d['roofline']['stage_ms'], 'frac', d['roofline']['frac'])
PY
