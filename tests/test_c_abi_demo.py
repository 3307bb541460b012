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
