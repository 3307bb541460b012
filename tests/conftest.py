import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


SNAP_KEYS = ("pos", "vel", "mass_f64", "mass_int_hi", "mass_int_lo", "mass_is_int", "radius", "volume",
             "alive", "immobile", "eaten")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def snap_from(npz, prefix):
    return {k: npz[prefix + k] for k in SNAP_KEYS}


def engine_arrays_from_snapshot(snap):
    """Reference-style snapshot (all planets, insertion order) -> Engine.upload arrays (alive only)."""
    al = np.asarray(snap["alive"], dtype=bool)
    idx = np.nonzero(al)[0]
    flags = (np.asarray(snap["immobile"], dtype=bool)[idx].astype(np.uint8) * 1 |
             np.asarray(snap["mass_is_int"], dtype=bool)[idx].astype(np.uint8) * 2)
    return {"pos": snap["pos"][idx], "vel": snap["vel"][idx], "mass_f64": snap["mass_f64"][idx],
            "mass_hi": snap["mass_int_hi"][idx], "mass_lo": snap["mass_int_lo"][idx], "flags": flags,
            "radius": snap["radius"][idx], "volume": snap["volume"][idx], "eaten": snap["eaten"][idx],
            "id": idx.astype(np.int32)}


def relerr_rows(a, b):
    """Per-row relative error of 2-vectors."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.linalg.norm(b, axis=1)
    den[den == 0] = 1.0
    return np.linalg.norm(a - b, axis=1) / den


# ---------------------------------------------------------------------------------------------
# 3-D (new API) fixtures: tests/golden/state3d_*.npz, oracle/gen_golden3d.py
# ---------------------------------------------------------------------------------------------
S3_KEYS = ("pos", "vel", "mass_f64", "mass_int_hi", "mass_int_lo", "mass_is_int", "dens_f64", "dens_is_int", "id")


def snap3_from(npz, prefix):
    return {k: npz[prefix + k] for k in S3_KEYS}


def python_numbers3(snap):
    """(masses, densities) of a 3-D snapshot as the Python numbers the reference held."""
    masses, dens = [], []
    for k in range(len(snap["mass_f64"])):
        if snap["mass_is_int"][k]:
            masses.append((int(snap["mass_int_hi"][k]) << 64) | int(snap["mass_int_lo"][k]))
        else:
            masses.append(float(snap["mass_f64"][k]))
        dens.append(int(snap["dens_f64"][k]) if snap["dens_is_int"][k] else float(snap["dens_f64"][k]))
    return masses, dens


def engine3d_arrays_from_snapshot(snap):
    """3-D snapshot -> Engine3D.upload arrays; radii by Object.get_radius() in Python (universe.py:33-37)."""
    import math
    masses, dens = python_numbers3(snap)
    n = len(masses)
    radius = np.array([((3 * (m / d)) / (4 * math.pi)) ** (1 / 3) for m, d in zip(masses, dens)])
    flags = (np.asarray(snap["mass_is_int"], dtype=bool).astype(np.uint8) * 1 |
             np.asarray(snap["dens_is_int"], dtype=bool).astype(np.uint8) * 2)
    return {"pos": snap["pos"], "vel": snap["vel"], "mass_f64": snap["mass_f64"], "mass_hi": snap["mass_int_hi"],
            "mass_lo": snap["mass_int_lo"], "density": snap["dens_f64"], "radius": radius, "flags": flags,
            "id": np.asarray(snap["id"], dtype=np.int32) if "id" in snap else np.arange(n, dtype=np.int32)}
