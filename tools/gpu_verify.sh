#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
python - <<'PY'
# graph capture with the symmetric kernel forced at a small size (attribute opt-in must already be in place)
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from cosmosim_b200 import scenes
from cosmosim_b200.engine3d import Engine3D
arr = scenes.disk3d_arrays(4000, seed=1)
for fp32 in (False, True):
    a, b = Engine3D(4000), Engine3D(4000)
    a.upload(arr); b.upload(arr)
    a.step(60.0, 6.674e-11, fp32=fp32, sym=True, steps=16)          # graph replay
    for _ in range(16):
        b.step(60.0, 6.674e-11, fp32=fp32, sym=True)                # eager
    da, db = a.download(), b.download()
    print("fp32", fp32, "graphs", len(a._graphs), "graph == eager:", np.array_equal(da["pos"], db["pos"]), a.check_errors()["iteration"])
PY
