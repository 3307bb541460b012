#!/bin/bash
# round 2, first GPU pass: GPU suite, default bench line, launch list.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('value %.4g ms/step %.3f frac %.3f stage %s outside %.3f' % (d['value'], d['ms_per_step'], r['frac'], r['stage_ms'], r['outside_stage_timers_ms']))
    print('full_n', d['full_n'])
    print('e2e', d['e2e'])
    s=d['secondary']; print('secondary', s and (s['value'], s['ms_per_step'], s['roofline']['frac'], s['roofline']['stage_ms']))
    print('digest', d['state_digest'], d['final_state'])
except Exception as e:
    print('parse failed', e)
PY
