#!/bin/bash
# round 2 experiment: 8 row planets per lane in k_accel_f32_sym (COSMO_SYM_RL=8) against the shipped 4
set -u
mkdir -p gpurun_out
COSMO_SYM_RL=8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp32 or symmetric or sym" > gpurun_out/r2_rl8_pytest.log 2>&1; echo "pytest rl8 rc=$?"; tail -5 gpurun_out/r2_rl8_pytest.log
for rl in 4 8 4 8; do
  echo "RL=$rl"; COSMO_SYM_RL=$rl timeout 300 python tools/probe_stage_a.py 2>&1 | tail -1
done | tee gpurun_out/r2_rl8_probe.log
