#!/bin/bash
set -u
NG=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $RUN tools/sharded3d_check.py --steps 3 > gpurun_out/sharded3d_check_$NG.log 2>&1; echo "sharded3d check rc=$?"
grep -E "sharded==single|mismatch|ALL OK|FAILED" gpurun_out/sharded3d_check_$NG.log | tail -20
