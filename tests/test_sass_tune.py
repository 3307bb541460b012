"""The SASS post-scheduler of the symmetric kernels (cosmosim_b200/sass_tune.py) on the CPU: it runs at
build time on cubins nvcc cross-compiles here; cuobjdump disassembles them without a GPU.  The bitwise
equality of the tuned kernels with ptxas's schedule is a GPU test (test_tuned_sass_equals_ptxas_schedule*)."""
import os
import shutil
import subprocess

import pytest

from cosmosim_b200 import _build, sass_tune

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("nvcc") is None,
                                reason="needs nvcc and cuobjdump")


@pytest.fixture(scope="module")
def cubins(tmp_path_factory):
    d = tmp_path_factory.mktemp("cubins")
    out = {}
    for src, kernels in _build.TUNED:
        cubin = str(d / (src[:-3] + ".cubin"))
        flags = [f for f in _build.NVCC_FLAGS if f not in ("-Xcompiler", "-fPIC", "-Xptxas", "-v")]
        subprocess.run([_build._nvcc()] + flags + ["-I", _build.INCLUDE, "-I", _build.CSRC, "-cubin", "-o", cubin,
                                                    os.path.join(_build.CSRC, src)], check=True, capture_output=True)
        out[src] = (cubin, kernels, str(d / (src[:-3] + "_tuned.cubin")))
    return out


def test_every_registered_kernel_is_tuned_and_verified(cubins):
    """tune_cubin re-disassembles its output and compares it instruction by instruction with the plan (same
    opcodes, registers and immediates, only accumulations moved / their product operands swapped, nothing
    outside the loops touched); here: it accepts every kernel the build registers and finds reuse."""
    for src, (cubin, kernels, tuned) in cubins.items():
        stats = sass_tune.tune_cubin(cubin, tuned, kernels)
        assert [s["kernel"] for s in stats] == list(kernels)
        for st in stats:
            assert st["moved"] >= st["accumulations"] // 2, st["kernel"]
            assert st["reuse_flags_on_moved"] >= st["moved"] // 2, st["kernel"]
        assert os.path.getsize(tuned) == os.path.getsize(cubin)
        # deterministic: the same input gives the same bytes
        again = tuned + ".again"
        sass_tune.tune_cubin(cubin, again, kernels)
        assert open(again, "rb").read() == open(tuned, "rb").read()


def test_new_order_respects_every_dependency_and_latency(cubins):
    """Independent re-check on the patched SASS of the FP32 kernel: inside the tuned loop every register is
    written and read in an order that is consistent with the original def-use chains, and every dependent
    pair of two-cycle instructions is at least the pipeline latency apart by the stall counts."""
    cubin, kernels, tuned = cubins["accel_sym.cu"]
    sass_tune.tune_cubin(cubin, tuned, kernels)
    k = kernels[-1]
    before = sass_tune._disassemble(cubin, k)
    after = sass_tune._disassemble(tuned, k)
    (first, last), = sass_tune._find_loops(before)
    assert sorted(t.replace(".reuse", "") for _, t, _, _ in before[first:last + 1] if not t.startswith("FFMA2")) == \
        sorted(t.replace(".reuse", "") for _, t, _, _ in after[first:last + 1] if not t.startswith("FFMA2"))
    body = [sass_tune._parse(a, t, lo, hi, i) for i, (a, t, lo, hi) in enumerate(after[first:last + 1])]
    lat = sass_tune._latency_table(before)
    fma_lat = min(v for (p, c), v in lat.items() if p in ("FFMA2", "FMUL2", "FADD2") and c in ("FFMA2", "FMUL2", "FADD2"))
    t, issue, writer = 0, {}, {}
    for i, x in enumerate(body):
        for r in x.src:
            w = writer.get(r)
            if w is not None and body[w].base in ("FFMA2", "FMUL2", "FADD2") and x.base in ("FFMA2", "FMUL2", "FADD2"):
                assert t - issue[w] >= fma_lat, (x.text, body[w].text)
        issue[i] = t
        for r in x.dst:
            writer[r] = i
        t += max(1, x.stall())
    # reuse flags only where the next instruction really reads that register in that slot
    for x, y in zip(body, body[1:]):
        if ".reuse" in x.text and x.base in sass_tune.PACKED:
            flagged = [s for s, a in enumerate(x.text.split(",")[1:]) if ".reuse" in a]
            xs = [a.strip().replace(".reuse", "") for a in x.text.split(",")[1:]]
            ys = [a.strip().replace(".reuse", "") for a in y.text.split(",")[1:]]
            for s in flagged:
                if x.base == y.base:
                    assert s < len(ys) and xs[s].lstrip("-") == ys[s].lstrip("-"), (x.text, y.text)


def test_dataflow_of_every_tuned_loop_is_unchanged(cubins):
    """Independent of the tool's own dependency graph: in every tuned loop of every registered kernel each
    source register of each instruction is produced by the SAME instruction as before (or is live-in as
    before), and the last writer of every register at the end of the body is the same instruction -- i.e.
    the patched body computes the same dataflow graph, whatever order it issues it in."""
    for src, (cubin, kernels, tuned) in cubins.items():
        stats = sass_tune.tune_cubin(cubin, tuned, kernels)
        for st in stats:
            before = sass_tune._disassemble(cubin, st["kernel"])
            assert st["loops"], st["kernel"]
            for loop in st["loops"]:
                first, order = loop["first"], loop["order"]
                body = [sass_tune._parse(a, t, lo, hi, i) for i, (a, t, lo, hi) in enumerate(before[first:first + len(order)])]

                def flow(seq):
                    writer, prod = {}, {}
                    for i in seq:
                        x = body[i]
                        prod[i] = tuple(sorted((r, writer.get(r, -1)) for r in set(x.src)))
                        for r in x.dst:
                            writer[r] = i
                    return prod, writer

                p0, w0 = flow(range(len(body)))
                p1, w1 = flow(order)
                assert p0 == p1, (st["kernel"], hex(loop["loop"][0]))
                assert w0 == w1, (st["kernel"], hex(loop["loop"][0]))


def test_the_build_shipped_tuned_kernels():
    """The in-tree build ran the tuner on every registered kernel (none declined)."""
    import cosmosim_b200
    cosmosim_b200.build()
    log = open(os.path.join(_build.OBJ_DIR, "sass_tune.log")).read()
    assert "declined" not in log, log
    for _, kernels in _build.TUNED:
        for k in kernels:
            assert k in log
    import re
    sizes = [int(m) for m in re.findall(r"\{kTunedCubin\d+, (\d+)\}", open(_build.TUNED_INC).read())]
    assert len(sizes) == len(_build.TUNED) and all(s > 100000 for s in sizes), sizes


def test_operand_bandwidth_model_reproduces_the_probes(tmp_path):
    """tools/sass_model.py (one issue slot per instruction, two for packed FP32 / FP64 ones, two registers
    per cycle from the register file, `.reuse` hits free) against the fractions of the pipe peak measured
    on the B200 by tools/probe_pipe.py (profiles/r02_probe_pipe.log): within 0.03 for every probe."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import sass_model
    cubin = str(tmp_path / "microbench.cubin")
    flags = [f for f in _build.NVCC_FLAGS if f not in ("-Xcompiler", "-fPIC", "-Xptxas", "-v")]
    subprocess.run([_build._nvcc()] + flags + ["-I", _build.INCLUDE, "-I", _build.CSRC, "-cubin", "-o", cubin,
                                                os.path.join(_build.CSRC, "microbench.cu")], check=True, capture_output=True)
    measured = {9: 49.3 / 74.1, 10: 74.0 / 74.1, 11: 60.6 / 74.1, 12: 64.7 / 74.1, 13: 24.7 / 36.7, 14: 36.7 / 36.7}
    kernels = sass_model.disassemble(cubin)
    seen = set()
    for name, ins in kernels.items():
        m = __import__("re").search(r"k_microbench2ILi(\d+)E", name)
        if not m:
            continue
        kind = int(m.group(1))
        big = max((sass_model.model(ins, lo, hi) for lo, hi in sass_model.loops(ins)), key=lambda r: r["two_cycle_pipe_cycles"])
        assert abs(big["pipe_fraction"] - measured[kind]) < 0.03, (kind, big["pipe_fraction"], measured[kind])
        seen.add(kind)
    assert seen == set(measured)
