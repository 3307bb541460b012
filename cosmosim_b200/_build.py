"""In-tree build of libcosmosim_b200.so (hand-written sm_100a CUDA behind a C ABI).

nvcc cross-compiles without a GPU; the resulting .so sits next to this file, is
git-ignored, and travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB_PATH = os.path.join(_HERE, "libcosmosim_b200.so")
SOURCES = ("accel.cu", "accel_sym.cu", "accel_sym64.cu", "detect.cu", "resolve.cu", "energy.cu", "scan.cu", "microbench.cu", "abi.cu", "state3d.cu", "state3d_sym.cu")
HEADERS = ("cosmo_device.cuh", "cosmo_internal.h", "pred.cuh", "integrate.cuh", "state3d.cuh")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libcosmosim_b200.so")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, h) for h in ("cosmosim_b200.h", "cosmosim_b200_state3d.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a into one shared library; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-o", LIB_PATH] + \
          [os.path.join(CSRC, f) for f in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libcosmosim_b200.so")
    with open(os.path.join(_HERE, "build_ptxas.log"), "w") as f:
        f.write(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
