#!/bin/bash
set -u
timeout 600 python -m pytest tests -m gpu -q -x -k "graph_replay or free_run or lockstep" 2>&1 | tail -4
timeout 300 python bench.py --workload star_disk_1001_fp64 --steps 1000 --no-cpu-baseline --no-e2e > gpurun_out/bench_1001_graph.json 2> gpurun_out/bench_1001_graph.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1001_graph.json').read().strip().splitlines()[-1])
print('value %.3e'%d['value'], 'ms/step', d['ms_per_step'], 'steps/s', d['steps_per_s'], d['final_state'])
PY
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
PY
