#!/bin/bash
# round 2 profiles: launch list of the default bench + ncu --set full of the two symmetric kernels.
# usage: gpu_r2_ncu.sh <tag>
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-extras"
# 1. launch list (per-launch device time) of the default workload, frames 6..7 (n ~ 600k)
$B --steps 2 --warmup 5 > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_star_disk_1m_fp32.csv \
    $B --steps 2 --warmup 5 > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list rc=$?"
# 2. full capture of the FP32 symmetric kernel: launch of frame 1 (n ~ 947k)
ncu --set full --clock-control none --import-source on -k regex:k_accel_f32_sym -s 1 -c 1 -o gpurun_out/${TAG}_prof_sym1m \
    $B --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu_full_sym1m.log 2>&1
echo "full f32 rc=$?"
# 3. full capture of the FP64 symmetric kernel at N = 65536
ncu --set full --clock-control none --import-source on -k regex:k_accel_f64_sym -s 1 -c 1 -o gpurun_out/${TAG}_prof_sym64 \
    $B --workload star_disk_64k_fp64 --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu_full_sym64.log 2>&1
echo "full f64 rc=$?"
ls -la gpurun_out/${TAG}_*
