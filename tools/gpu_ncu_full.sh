#!/bin/bash
# ncu --set full capture of one kernel.  usage: gpu_ncu_full.sh <workload> <kernel-regex> <tag> [skip]
set -u
W=${1:-cloud_256k_fp32}; K=${2:-k_accel_f32_sym}; TAG=${3:-sym}; SKIP=${4:-1}
CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log
