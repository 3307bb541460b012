"""In-tree build of libcosmosim_b200.so (hand-written sm_100a CUDA behind a C ABI).

nvcc cross-compiles without a GPU; the resulting .so sits next to this file, is
git-ignored, and travels to the GPU box with the repo snapshot.  Every .cu is compiled to its
own object (in parallel, only when it or a header changed) and the objects are linked into
one shared library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB_PATH = os.path.join(_HERE, "libcosmosim_b200.so")
OBJ_DIR = os.path.join(_HERE, "build")
SOURCES = ("accel.cu", "accel_sym.cu", "accel_sym64.cu", "nearfield.cu", "detect.cu", "resolve.cu", "energy.cu",
           "scan.cu", "microbench.cu", "abi.cu", "host_step.cu", "state3d.cu", "state3d_sym.cu")
HEADERS = ("cosmo_device.cuh", "cosmo_internal.h", "pred.cuh", "integrate.cuh", "grid.cuh", "state3d.cuh")
PUBLIC_HEADERS = ("cosmosim_b200.h", "cosmosim_b200_state3d.h")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libcosmosim_b200.so")


def _sources():
    return list(SOURCES)


def _header_mtime() -> float:
    deps = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(INCLUDE, h) for h in PUBLIC_HEADERS]
    return max(os.path.getmtime(d) for d in deps)


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return _header_mtime() > t or any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a into one shared library; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = _header_mtime()
    logs = {}

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        log = obj + ".log"
        if (not force and os.path.exists(obj) and os.path.exists(log)
                and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_t)):
            logs[src] = open(log).read()
            return 0, obj
        proc = subprocess.run([nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-c", path, "-o", obj],
                              capture_output=True, text=True)
        logs[src] = proc.stdout + proc.stderr
        if proc.returncode == 0:
            with open(log, "w") as f:
                f.write(logs[src])
        return proc.returncode, obj

    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    failed = [s for s, (rc, _) in zip(srcs, results) if rc != 0]
    if verbose or failed:
        for s in (failed or srcs):
            print("==== %s\n%s" % (s, logs[s]))
    if failed:
        raise RuntimeError("nvcc failed on %s" % ", ".join(failed))
    proc = subprocess.run([nvcc, "--shared", "-o", LIB_PATH] + [o for _, o in results], capture_output=True, text=True)
    if proc.returncode != 0:
        print(proc.stdout, proc.stderr)
        raise RuntimeError("nvcc failed linking libcosmosim_b200.so")
    with open(os.path.join(_HERE, "build_ptxas.log"), "w") as f:
        for s in srcs:
            f.write("==== %s\n%s" % (s, logs[s]))
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
