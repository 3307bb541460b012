#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python - <<'PY'
import time, sys
sys.path.insert(0,'.')
import torch
from cosmosim_b200 import scenes
for name in ("star_with_disk","cloud"):
    u = getattr(scenes,name)(0)
    u.simulate(steps=10); u.synchronize()
    t=time.perf_counter(); u.simulate(steps=1000); u.synchronize(); dt=time.perf_counter()-t
    print(name, "1000 frames through Universe.simulate(steps=1000): %.3f s  -> %.1f us/frame, %d merges, %d active"%(dt, dt*1e3, len(u.merge_events), len(u.get_active_planets())))
from cosmosim_b200.core.state import State
objs = scenes.star_with_disk_objects(0, 1000)
st = State(objs, dt=60.0)
st.step(10); torch.cuda.synchronize()
t=time.perf_counter(); st.step(1000); torch.cuda.synchronize(); dt=time.perf_counter()-t
print("3-D in-memory scene: 1000 steps through State.step(1000): %.3f s -> %.1f us/step, %d objects" % (dt, dt*1e3, len(st.objects)))
PY
