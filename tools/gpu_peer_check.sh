#!/bin/bash
# experimental peer-store exchange: short, strictly time-limited check on NG GPUs
set -u
NG=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29577"
COSMO_PEER_EXCHANGE=1 timeout 150 $RUN tools/sharded_check.py --bodies 70000 --frames 3 --fp32 > gpurun_out/peer_check_$NG.log 2>&1; echo "peer fp32 rc=$?"
grep -E "sharded==single|mismatch|exchange:|Error" gpurun_out/peer_check_$NG.log | tail -12
