#!/bin/bash
# round 2: the 3-D near field (k_nearfield3) -- the 3-D GPU tests, then the FP32 error probe of the 2-D path is untouched
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_state3d.py -m gpu -x -q > gpurun_out/r2_near3_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_near3_pytest.log
