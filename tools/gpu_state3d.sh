#!/bin/bash
# 3-D step: parity tests + a quick timing probe (no profiler)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_state3d.py -m gpu -q > gpurun_out/pytest_state3d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_state3d.log
tail -40 gpurun_out/pytest_state3d.log
[ -n "${SKIP_PROBE:-}" ] || timeout 300 python tools/probe_state3d.py > gpurun_out/probe_state3d.log 2>&1; echo "probe rc=$?"
[ -n "${SKIP_PROBE:-}" ] || tail -20 gpurun_out/probe_state3d.log
