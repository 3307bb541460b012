#!/bin/bash
# round 2 experiment: row planets per lane (COSMO_SYM_RL) x unroll (COSMO_SYM_UN) of k_accel_f32_sym
set -u
mkdir -p gpurun_out
for v in "4 2" "4 4" "8 1" "8 2" "4 1"; do
  set -- $v
  echo "RL=$1 UN=$2"; COSMO_SYM_RL=$1 COSMO_SYM_UN=$2 timeout 300 python tools/probe_stage_a.py 2>&1 | tail -1
done | tee gpurun_out/r2_rl8_probe.log
