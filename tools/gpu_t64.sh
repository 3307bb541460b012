#!/bin/bash
set -u
timeout 800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --workload star_disk_64k_fp64 --no-cpu-baseline --no-e2e > gpurun_out/b4_64k.json 2>gpurun_out/b4_64k.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/b4_64k.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("value %.3e"%d["value"], d["ms_per_step"], 'frac', r["frac"], r["stage_ms"], d["final_state"])
except Exception as e: print("bench failed", e)
PY
tail -n 3 gpurun_out/b4_64k.err
