#!/bin/bash
# ncu launch list (per-launch device time) of a short bench run.  usage: gpu_ncu_launches.sh <workload> <tag>
set -u
W=${1:-star_disk_1m_fp32}; TAG=${2:-v}
CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${W}_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "rc=$?"
python tools/summarize_launches.py gpurun_out/launches_${W}_$TAG.csv
