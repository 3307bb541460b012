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


def test_grid_tuning_invariants():
    from cosmosim_b200.engine import tune_grid
    rng = np.random.default_rng(0)
    for n, spread, big in ((20000, 1.5e11, 7e8), (300000, 7e12, 0.0), (50000, 1.0, 0.0)):
        pos = rng.normal(size=(n, 2)) * spread
        pos[:5] *= 1e4                                    # a few planets flung far out must not stretch the box
        radius = rng.uniform(1e6, 3e7, n) * (spread / 1.5e11)
        if big:
            radius[0] = big
        grid_cells, w = 1 << 22, 32
        h, ox, oy, c, rg = tune_grid(pos, radius, grid_cells, w)
        assert c * c <= grid_cells and c - 2 * w >= 1
        assert h >= 2.0 * rg > 0 and math.isfinite(ox) and math.isfinite(oy)
        lin = (c - 2 * w) * h
        assert lin < 50 * spread                          # the box follows the bulk, not the outliers
        giants = int((radius * (1 + 2.0 ** -30) > rg).sum())
        assert giants <= 1024                             # the cost model never brute-forces more than that
