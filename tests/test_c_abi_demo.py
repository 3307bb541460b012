"""The C ABI from plain C: both headers compile as strict C99, a C program links against
libcosmosim_b200.so without Python or torch, and (on the GPU box) drives cosmo_accel_host."""
import os
import shutil
import subprocess

import pytest

from cosmosim_b200 import _abi, _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
SRC = os.path.join(ROOT, "examples", "c_abi_demo.c")

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None or not os.path.isdir(os.path.join(CUDA, "include")),
                                reason="needs gcc and the CUDA headers")


def build_demo(tmp_path):
    _abi.lib()                                   # makes sure the library is built
    exe = str(tmp_path / "c_abi_demo")
    lib_dir = os.path.dirname(_build.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", _build.INCLUDE, "-isystem",
           os.path.join(CUDA, "include"), SRC, "-L", lib_dir, "-lcosmosim_b200", "-L", os.path.join(CUDA, "lib64"),
           "-lcudart", "-lm", "-o", exe]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    env = dict(os.environ, LD_LIBRARY_PATH=lib_dir + ":" + os.path.join(CUDA, "lib64") + ":" +
               os.environ.get("LD_LIBRARY_PATH", ""))
    return exe, env


def test_headers_are_c99_and_the_library_links_from_c(tmp_path):
    exe, env = build_demo(tmp_path)
    out = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, env=env, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK (no device used)" in out.stdout and "abi version %d" % _abi.ABI_VERSION in out.stdout


@pytest.mark.gpu
def test_c_program_drives_the_gpu_path(tmp_path):
    exe, env = build_demo(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK") and "cosmo_accel_host: 4097 planets" in out.stdout


@pytest.mark.gpu
def test_c_program_steps_whole_frames_through_cosmo_step_host(tmp_path):
    """cosmo_workspace_init + cosmo_params_default + cosmo_step_host from plain C: three frames of a
    crowded 2049-planet disk (host buffers in, survivors and merge events out), replayed by the CPU
    oracle from the same initial state: merge stream, ids, masses, eaten counts exact, radii to one
    ulp, positions and velocities <= 1e-12."""
    import numpy as np
    from oracle import oracle as orc
    exe, env = build_demo(tmp_path)
    path = str(tmp_path / "frames.bin")
    out = subprocess.run([exe, "--step", path], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK")
    raw = open(path, "rb").read()
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a

    n0, frames = (int(x) for x in take(np.int64, 4)[:2])
    pos0, vel0 = take(np.float64, 2 * n0).reshape(n0, 2), take(np.float64, 2 * n0).reshape(n0, 2)
    mass0, radius0, volume0, flags0 = take(np.float64, n0), take(np.float64, n0), take(np.float64, n0), take(np.uint8, n0)
    events = []
    for f in range(frames):
        ne = int(take(np.int64, 1)[0])
        events += [tuple(int(x) for x in r) for r in take(np.int64, 3 * ne).reshape(ne, 3)]
    n1, total = (int(x) for x in take(np.int64, 2))
    ids, pos1, vel1 = take(np.int32, n1), take(np.float64, 2 * n1).reshape(n1, 2), take(np.float64, 2 * n1).reshape(n1, 2)
    mass1, radius1, eaten1 = take(np.float64, n1), take(np.float64, n1), take(np.int64, n1)
    assert off == len(raw) and total == len(events) >= 10
    snap = {"pos": pos0, "vel": vel0, "mass_f64": mass0, "mass_int_hi": np.zeros(n0, np.uint64),
            "mass_int_lo": np.zeros(n0, np.uint64), "mass_is_int": np.zeros(n0, bool), "radius": radius0,
            "volume": volume0, "alive": np.ones(n0, bool), "immobile": (flags0 & 1) != 0, "eaten": np.zeros(n0, np.int64)}
    ou = orc.OracleUniverse(snap, G=6.674e-11)
    want = []
    for f in range(frames):
        want += [(f, int(a), int(b)) for a, b in ou.frame(1e6 / 33)["events"]]
    assert events == want
    post = ou.snapshot()
    assert np.array_equal(ids, np.nonzero(post["alive"])[0])
    assert np.array_equal(mass1, post["mass_f64"][ids]) and np.array_equal(eaten1, post["eaten"][ids])
    assert np.abs(radius1 - post["radius"][ids]).max() <= np.spacing(post["radius"][ids]).max()
    for got, ref in ((pos1, post["pos"][ids]), (vel1, post["vel"][ids])):
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
