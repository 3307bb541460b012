#!/bin/bash
# ncu launch list (per-launch device time) of a short default-bench run.  usage: gpu_r2_launches.sh <tag> [warmup] [extra bench args]
set -u
TAG=${1:-a}; W=${2:-12}; shift; shift
CMD="python bench.py --steps 2 --warmup $W --no-e2e --no-cpu-baseline --no-extras $*"
$CMD > gpurun_out/r2_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_$TAG.csv $CMD > gpurun_out/r2_ncu_launch_$TAG.log 2>&1
echo "rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_$TAG.csv
