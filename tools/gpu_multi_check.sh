#!/bin/bash
set -u
NG=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $RUN tools/sharded_check.py --bodies 20000 --frames 4 > gpurun_out/sharded_check_$NG.log 2>&1; echo "sharded fp64 rc=$?"
timeout 600 $RUN tools/sharded_check.py --bodies 70000 --frames 3 --fp32 >> gpurun_out/sharded_check_$NG.log 2>&1; echo "sharded fp32 rc=$?"
grep -E "sharded==single|mismatch|exchange:" gpurun_out/sharded_check_$NG.log | tail -20
