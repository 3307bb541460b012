#!/bin/bash
# One GPU session: tests, contract bench, ncu launch list and full capture of the dominant kernels.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1m_fp32.json 2> gpurun_out/bench_1m_fp32.err; echo "bench1m rc=$?"
timeout 300 python bench.py --workload star_disk_64k_fp64 > gpurun_out/bench_64k_fp64.json 2> gpurun_out/bench_64k_fp64.err; echo "bench64k rc=$?"
timeout 300 python bench.py --workload cloud_256k_fp32 --no-cpu-baseline > gpurun_out/bench_256k_fp32.json 2> gpurun_out/bench_256k_fp32.err; echo "bench256k rc=$?"
timeout 300 python bench.py --workload star_disk_1001_fp64 --steps 200 --no-cpu-baseline > gpurun_out/bench_1001_fp64.json 2> gpurun_out/bench_1001.err; echo "bench1001 rc=$?"
CMD="python bench.py --workload cloud_256k_fp32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_256k_fp32.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_accel_f32 -s 1 -c 2 -o gpurun_out/prof_accel_f32 $CMD > gpurun_out/ncu_full_f32.log 2>&1
echo "ncu f32 rc=$?"
CMD2="python bench.py --workload star_disk_64k_fp64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/ncu_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_accel_f64\|k_detect -s 2 -c 2 -o gpurun_out/prof_accel_f64 $CMD2 > gpurun_out/ncu_full_f64.log 2>&1
echo "ncu f64 rc=$?"
cat gpurun_out/bench_1m_fp32.json gpurun_out/bench_64k_fp64.json gpurun_out/bench_256k_fp32.json gpurun_out/bench_1001_fp64.json
