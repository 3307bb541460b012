#!/bin/bash
# 3-D step: bench lines, then the ncu launch list of the 64k FP64 line (same command, after it exited 0)
set -u
mkdir -p gpurun_out
bash tools/gpu_state3d_bench.sh
CMD="python bench.py --workload state3d_64k_fp64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain_s3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_state3d_64k_fp64.csv $CMD > gpurun_out/ncu_launch_s3.log 2>&1
echo "ncu launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_state3d_64k_fp64.csv
