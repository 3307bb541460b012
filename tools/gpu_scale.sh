#!/bin/bash
# contract bench under torchrun at NG GPUs (default workload)
set -u
NG=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544"
timeout 600 $RUN bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/scale_1m_fp32_g$NG.json 2> gpurun_out/scale_1m_fp32_g$NG.err; echo "bench rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/scale_1m_fp32_g$NG.err | tail -n 5
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_1m_fp32_g$NG.json").read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value %.3e'%d['value'], 'ms/step', d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), d['final_state'], d['clocks'])
print(d['roofline']['stage_ms'], 'exchange', d['roofline'].get('exchange_ms'), 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
PY
