#!/bin/bash
# round 2: the post-scheduled k_accel_f32_sym against the ptxas schedule -- same bits, stage-A time of both
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tuned_sass or super_tiles" > gpurun_out/r2_tuned_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tuned_pytest.log
for t in 0 1 0 1; do
  echo "COSMO_SYM_TUNED=$t"; COSMO_SYM_TUNED=$t timeout 300 python tools/probe_stage_a.py 2>&1 | tail -1
done | tee gpurun_out/r2_tuned_probe.log
