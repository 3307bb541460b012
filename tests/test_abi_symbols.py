"""The C-ABI library loads and exports every symbol the headers under include/ declare; the
host-callable exact-arithmetic helpers agree with Python's own numeric semantics.  No GPU."""
import ctypes as C
import math
import os
import random
import re

import pytest

from cosmosim_b200 import _abi, _build

HEADERS = [os.path.join(_build.INCLUDE, h) for h in sorted(os.listdir(_build.INCLUDE)) if h.endswith(".h")]


def header_functions():
    names = set()
    for header in HEADERS:
        src = open(header).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names.update(re.findall(r"\b(cosmo3?_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_builds_and_loads():
    L = _abi.lib()
    assert L.cosmo_abi_version() == _abi.ABI_VERSION


def test_every_declared_symbol_is_exported_and_bound():
    L = _abi.lib()
    declared = header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), "header declares %s but the library does not export it" % name
        assert name in _abi.SYMBOLS, "no ctypes prototype for %s" % name
    assert sorted(_abi.SYMBOLS) == declared


def test_struct_layouts_match_header_sizes():
    import ctypes as C
    # 2 int32 + 20 pointers + int64  /  pointers and ints as declared in the header
    assert C.sizeof(_abi.State) == 8 + 22 * 8 + 8
    assert C.sizeof(_abi.StepParams) == 3 * 8 + 16 * 4 + 3 * 8 + 2 * 4 + 8


def test_bad_arguments_are_rejected_without_a_gpu():
    L = _abi.lib()
    assert L.cosmo_accel(None, None, 0, 1.0, 0, 0, 0, None, None, None) == -1
    assert b"bad argument" in L.cosmo_last_error()
    assert L.cosmo_accel_workspace_bytes(1000) > 1000 * 100


def test_u128_to_double_is_python_float():
    L = _abi.lib()
    rnd = random.Random(5)
    vals = [0, 1, 2 ** 53, 2 ** 53 + 1, 2 ** 64 - 1, 2 ** 64, 2 ** 64 + 2 ** 11, 2 ** 64 + 2 ** 11 + 1,
            2 ** 127 - 1, 5972000000000000327155712 * 1000, 260918406660053865733901484]
    vals += [rnd.getrandbits(rnd.randint(1, 127)) for _ in range(20000)]
    # ties: exactly halfway between two doubles
    for e in (54, 70, 100, 126):
        base = rnd.getrandbits(52) | (1 << 52)
        vals += [(base << (e - 52)) + (1 << (e - 53)), ((base | 1) << (e - 52)) + (1 << (e - 53))]
    for v in vals:
        assert L.cosmo_host_u128_to_double(v >> 64, v & (2 ** 64 - 1)) == float(v), v


def test_mass_ge_is_python_comparison():
    L = _abi.lib()
    rnd = random.Random(7)

    def enc(m):
        if isinstance(m, int):
            return 1, m >> 64, m & (2 ** 64 - 1), float(m)
        return 0, 0, 0, m

    cases = [(5, 5), (5, 5.0), (5.0, 5), (2 ** 80 + 1, float(2 ** 80)), (float(2 ** 80), 2 ** 80 + 1),
             (2 ** 80, float(2 ** 80)), (3, 2.5), (2, 2.5), (2.5, 2), (2.5, 3), (1.989e30, 5 * 10 ** 26),
             (10 ** 30, 1.989e30), (0, 0.0), (1, 1e-300), (10 ** 38, 1e38), (1e38, 10 ** 38)]
    for _ in range(20000):
        a = rnd.getrandbits(rnd.randint(1, 120))
        b = float(a) if rnd.random() < 0.3 else rnd.random() * 10 ** rnd.randint(0, 36)
        cases += [(a, b), (b, a), (a, a + rnd.randint(-2, 2))]
    for a, b in cases:
        if isinstance(b, int) and b < 0:
            continue
        assert bool(L.cosmo_host_mass_ge(*enc(a), *enc(b))) == (a >= b), (a, b)


def test_pow_third_matches_libm_pow():
    """planet.py:82 uses C pow(x, 1/3).  pow_third is correctly rounded; glibc's pow misrounds
    ~0.1% of inputs by one ulp (true result within 0.001 ulp of a tie), so: never more than one
    ulp apart and equal in > 99.8% of draws."""
    L = _abi.lib()
    rnd = random.Random(11)
    n, diff = 100000, 0
    for _ in range(n):
        x = 10.0 ** rnd.uniform(-10, 70) * (1 + rnd.random())
        want, got = x ** (1 / 3), L.cosmo_host_pow_third(x)
        if want != got:
            diff += 1
            assert abs(want - got) <= math.ulp(want)
    assert diff < 0.002 * n


def test_ctypes_structs_match_the_c_compilers_layout(tmp_path):
    """sizeof / offsetof of every struct of both headers as gcc sees them == the ctypes mirrors."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    pairs = [("cosmo_state", _abi.State), ("cosmo_scratch", _abi.Scratch), ("cosmo_step_params", _abi.StepParams),
             ("cosmo_grid", _abi.Grid), ("cosmo_host_frame", _abi.HostFrame), ("cosmo3_state", _abi.State3),
             ("cosmo3_scratch", _abi.Scratch3), ("cosmo3_step_params", _abi.StepParams3)]
    lines = []
    for cname, ct in pairs:
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "cosmosim_b200.h"\n#include "cosmosim_b200_state3d.h"\n'
                   'int main(void) {\n' + "\n".join(lines) + '\nreturn 0;\n}\n')
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    proc = subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe)], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, ct in pairs:
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)
