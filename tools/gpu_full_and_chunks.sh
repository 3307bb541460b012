#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/probe_state3d_chunks.py > gpurun_out/probe_chunks3d.log 2>&1; echo "chunks rc=$?"
cat gpurun_out/probe_chunks3d.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
