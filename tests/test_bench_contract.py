"""bench.py on CPU: the `--impl reference` arm (the only arm that runs without a GPU) prints one
well-formed JSON line with the contract's keys for the 2-D and the 3-D workloads; the grid tuning
used by both engines respects its invariants.  No GPU."""
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("workload", ["star_disk_1m_fp32", "state3d_64k_fp64"])
def test_reference_arm_prints_the_contract_line(workload):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--bodies", "3000", "--steps", "2", "--warmup", "1", "--cpu-seconds", "0.5"],
                         capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pairwise_interactions_per_s" and d["unit"] == "interactions/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 1e6
    assert d["config"]["workload"] == workload and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


@pytest.mark.timeout(300)
def test_reference_arm_uses_all_cores_under_torchrun_env():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the reference arm must not shrink to one core
    there (round-1 SCALE ratios at N >= 2 were inflated ~10x by exactly that)."""
    cores = []
    for extra in ({}, {"OMP_NUM_THREADS": "1", "WORLD_SIZE": "2", "RANK": "0", "LOCAL_RANK": "0"}):
        env = dict(os.environ, **extra)
        if not extra:
            env.pop("OMP_NUM_THREADS", None)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--bodies", "2000",
                              "--steps", "1", "--warmup", "0", "--cpu-seconds", "0.2"],
                             capture_output=True, text=True, timeout=280, cwd=ROOT, env=env)
        assert out.returncode == 0, out.stderr[-2000:]
        cores.append(json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]["cores"])
    assert cores[0] == cores[1] == len(os.sched_getaffinity(0))
