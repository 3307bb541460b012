"""Parity of the CUDA 3-D step (include/cosmosim_b200_state3d.h; the reference's State.interact,
cosmosim/core/universe.py:94-129) against the reference fixtures and the CPU oracle.  B200 only.

Bars: absorb-call stream, flagged pairs, ids, masses and densities bit-exact; accelerations,
positions and velocities within 1e-12 (summation order is the only freedom); merged radii within
one ulp (glibc pow is not correctly rounded, the device's pow(x, 1/3) is); FP32 flavour within the
bounds stated in the test."""
import math
import os
import random

import numpy as np
import pytest

from conftest import engine3d_arrays_from_snapshot, load_golden, python_numbers3, relerr_rows, snap3_from

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    from cosmosim_b200 import _abi, engine3d
    from oracle import oracle3d
    assert torch.cuda.is_available()
    return engine3d, oracle3d, _abi


def _abi_bits(mods):
    return mods[2]


def synthetic_cloud(n, seed, thick=0.2):
    rng = np.random.default_rng(seed)
    AU, ME, MS = 1.496e11, 5.972e24, 1.989e30
    pos = rng.normal(size=(n, 3)) * AU * np.array([1.0, 0.7, thick])
    vel = rng.normal(size=(n, 3)) * 1e4
    mass = rng.uniform(0.1, 100.0, n) * ME
    mass[0] = MS
    dens = np.full(n, 3000.0)
    snap = {"pos": pos, "vel": vel, "mass_f64": mass, "mass_int_hi": np.zeros(n, np.uint64),
            "mass_int_lo": np.zeros(n, np.uint64), "mass_is_int": np.zeros(n, bool), "dens_f64": dens,
            "dens_is_int": np.zeros(n, bool), "id": np.arange(n)}
    return snap


def step_once(engine3d, snap, dt, G, collisions=True, fp32=False, sym=None, resolve_mode=None, stage_events=None):
    arr = engine3d_arrays_from_snapshot(snap)
    n = len(arr["mass_f64"])
    eng = engine3d.Engine3D(n)
    eng.resolve_mode = resolve_mode
    eng.upload(arr)
    eng.step(dt, G, collisions=collisions, fp32=fp32, store_acc=True, sym=sym, stage_events=stage_events)
    eng.check_errors()
    return eng, eng.last_step_tensors(n)


def test_accel_fixture_against_reference(mods):
    engine3d, o3, _ = mods
    g = load_golden("state3d_accel_n2001.npz")
    n = len(g["mass"])
    snap = synthetic_cloud(n, 0)
    snap["pos"], snap["mass_f64"] = g["pos"], g["mass"]
    eng, out = step_once(engine3d, snap, 1.0, float(g["G"]), collisions=False)
    assert relerr_rows(out["acc"], g["acc"]).max() < 1e-12


@pytest.mark.parametrize("n,rows,sym", [(1000, None, False), (9001, None, False), (9001, None, True), (1537, None, True),
                                        (65536, 3000, False), (65536, 3000, True)])
def test_accel_against_oracle(mods, n, rows, sym):
    """sym: the symmetric kernel (each unordered pair once; default from 16384 objects on)."""
    engine3d, o3, _ = mods
    snap = synthetic_cloud(n, n)
    eng, out = step_once(engine3d, snap, 60.0, 6.674e-11, collisions=False, sym=sym)
    assert bool(eng.params(60.0, 6.674e-11, False, sym=sym).opts & _abi_bits(mods).OPT3_F64_SYM) == sym
    i0, i1 = (0, n) if rows is None else (n // 2 - rows // 2, n // 2 + rows // 2)
    want = o3.accel(snap["pos"], snap["mass_f64"], 6.674e-11, i0, i1)
    assert relerr_rows(out["acc"][i0:i1], want).max() < 1e-12
    p, v = o3.integrate(snap["pos"][i0:i1], snap["vel"][i0:i1], out["acc"][i0:i1], 60.0)
    assert np.array_equal(out["pos_new"][i0:i1], p) and np.array_equal(out["vel_new"][i0:i1], v)
    # Newton's third law as a whole-matrix property: sum m a = 0 up to rounding
    ma = out["acc"] * snap["mass_f64"][:, None]
    assert np.abs(ma.sum(0)).max() < 1e-9 * np.abs(ma).sum(0).max()


@pytest.mark.parametrize("name,grid", [("disk_planar", False), ("disk_thick", False), ("inmemory", False),
                                       ("disk_planar", True), ("disk_thick", True)])
def test_kept_steps_match_reference(mods, name, grid):
    """grid=True forces the uniform-grid broad phase (normally used from 16384 objects on)."""
    engine3d, o3, _ = mods
    g = load_golden("state3d_%s.npz" % name)
    for s in g["kept"]:
        pre = snap3_from(g, "s%d_pre_" % s)
        arr = engine3d_arrays_from_snapshot(pre)
        assert np.array_equal(arr["radius"], g["s%d_radius" % s])
        n = len(arr["mass_f64"])
        eng = engine3d.Engine3D(n)
        if grid:
            eng.grid_threshold = 0
        eng.upload(arr, iteration=int(s))
        assert (eng.grid is not None) == grid
        eng.step(float(g["dt"]), float(g["G"]), store_acc=True)
        eng.check_errors()
        out = eng.last_step_tensors(n)
        assert relerr_rows(out["acc"], g["s%d_acc" % s]).max() < 1e-12
        assert np.abs(out["pos_new"] - g["s%d_p_new" % s]).max() <= 1e-12 * np.abs(g["s%d_p_new" % s]).max()
        assert np.array_equal(out["flag_pairs"], g["s%d_flag_pairs" % s])
        ev = eng.events()
        want_ev = g["s%d_events" % s]
        assert np.array_equal(ev[:, 1:].reshape(-1, 3), want_ev.reshape(-1, 3))
        assert np.all(ev[:, 0] == s)
        d, want = eng.download(), snap3_from(g, "s%d_post_" % s)
        assert np.array_equal(d["id"], want["id"])
        assert np.array_equal(d["mass"], want["mass_f64"])
        assert np.array_equal(d["mass_hi"], want["mass_int_hi"]) and np.array_equal(d["mass_lo"], want["mass_int_lo"])
        assert np.array_equal(d["density"], want["dens_f64"])
        assert np.array_equal(d["flags"] & 1, want["mass_is_int"].astype(np.uint8))
        assert np.array_equal((d["flags"] >> 1) & 1, want["dens_is_int"].astype(np.uint8))
        for k, w in (("pos", want["pos"]), ("vel", want["vel"])):
            assert np.abs(d[k] - w).max() <= 1e-12 * np.abs(w).max(), (s, k)
        masses, dens = python_numbers3(want)
        r_want = np.array([((3 * (m / q)) / (4 * math.pi)) ** (1 / 3) for m, q in zip(masses, dens)])
        assert np.all(np.abs(d["radius"] - r_want) <= np.spacing(r_want))


@pytest.mark.parametrize("name,graphed", [("disk_planar", False), ("disk_thick", False), ("disk_planar", True),
                                          ("disk_thick", True)])
def test_free_run_absorb_stream(mods, name, graphed):
    """graphed: all steps in one call, i.e. through the CUDA-graph replay used for small N."""
    engine3d, o3, _ = mods
    g = load_golden("state3d_%s.npz" % name)
    init = snap3_from(g, "init_")
    arr = engine3d_arrays_from_snapshot(init)
    eng = engine3d.Engine3D(len(arr["mass_f64"]))
    eng.upload(arr)
    if graphed:
        eng.step(float(g["dt"]), float(g["G"]), steps=int(g["steps"]))
        assert len(eng._graphs) >= 1
        counts = [eng.status()["n_active"]]
        assert counts[0] == g["n_objects"][-1]
    else:
        counts = []
        for s in range(int(g["steps"])):
            eng.step(float(g["dt"]), float(g["G"]))
            counts.append(eng.status()["n_active"])
        assert np.array_equal(np.array(counts), g["n_objects"])
    eng.check_errors()
    assert eng.status()["iteration"] == int(g["steps"])
    assert np.array_equal(eng.events(), g["events"])
    d, want = eng.download(), snap3_from(g, "final_")
    assert np.array_equal(d["id"], want["id"])
    assert np.array_equal(d["flags"] & 1, want["mass_is_int"].astype(np.uint8))
    assert np.allclose(d["mass"], want["mass_f64"], rtol=1e-15, atol=0)
    assert np.abs(d["pos"] - want["pos"]).max() <= 1e-9 * np.abs(want["pos"]).max()


def test_micro_scenes(mods):
    engine3d, o3, _ = mods
    g = load_golden("state3d_micro.npz")
    for name in g["names"]:
        snap = snap3_from(g, "%s_pre_" % name)
        ev = []
        for s in (0, 1):
            arr = engine3d_arrays_from_snapshot(snap)
            assert np.array_equal(arr["radius"], g["%s_radius%d" % (name, s)]), name
            n = len(arr["mass_f64"])
            eng = engine3d.Engine3D(n)
            eng.upload(arr, iteration=s)
            eng.step(1.0, float(g["%s_G" % name]), store_acc=True)
            eng.check_errors()
            acc = eng.last_step_tensors(n)["acc"]
            want_acc = g["%s_acc%d" % (name, s)]
            assert np.abs(acc - want_acc).max() <= 1e-13 * np.abs(want_acc).max(), (name, s)
            ev += [(int(a), int(b), int(c), int(e)) for a, b, c, e in eng.events()]
            d, want = eng.download(), snap3_from(g, "%s_post%d_" % (name, s))
            assert np.array_equal(d["id"], want["id"]), (name, s)
            assert np.array_equal(d["mass"], want["mass_f64"]) and np.array_equal(d["density"], want["dens_f64"]), (name, s)
            assert np.array_equal(d["mass_hi"], want["mass_int_hi"]) and np.array_equal(d["mass_lo"], want["mass_int_lo"])
            assert np.array_equal(d["flags"] & 1, want["mass_is_int"].astype(np.uint8)), (name, s)
            assert np.array_equal((d["flags"] >> 1) & 1, want["dens_is_int"].astype(np.uint8)), (name, s)
            # merged velocities can cancel to (almost) zero: the scale is the step's own velocity change
            scale = {"pos": np.abs(want["pos"]).max(), "vel": max(np.abs(want["vel"]).max(), np.abs(want_acc).max())}
            for k in ("pos", "vel"):
                assert np.abs(d[k] - want[k]).max() <= 1e-13 * scale[k], (name, s, k)
            snap = want
        assert np.array_equal(np.array(ev, dtype=np.int64).reshape(-1, 4), g["%s_events" % name]), name


@pytest.mark.parametrize("n,squeeze,resolve_mode,min_calls", [(6000, 0.02, 0, 100), (6000, 0.02, 1, 100),
                                                              (60000, 0.06, 1, 10000)])
def test_crowded_step_against_oracle(mods, n, squeeze, resolve_mode, min_calls):
    """A crowded 3-D cloud (hundreds to tens of thousands of flagged pairs, chains, destroyed-object
    re-absorbs) for one step: pairs and the absorb stream must equal the oracle's exactly, with the
    sequential stage C and with the parallel one (row-major counting sort + connected components)."""
    engine3d, o3, _ = mods
    snap = synthetic_cloud(n, 77, thick=0.5)
    snap["pos"] *= squeeze                   # 0.02: radii ~ 1e7 m in a ~3e9 m cloud
    rnd = random.Random(5)
    ints = [rnd.randint(int(0.1 * 5.972e24), int(100 * 5.972e24)) for _ in range(n)]
    snap["mass_is_int"][:] = True
    snap["mass_int_hi"] = np.array([m >> 64 for m in ints], dtype=np.uint64)
    snap["mass_int_lo"] = np.array([m & ((1 << 64) - 1) for m in ints], dtype=np.uint64)
    snap["mass_f64"] = np.array([float(m) for m in ints])
    snap["dens_f64"][:] = 30.0
    snap["dens_is_int"][:] = True
    orc = o3.OracleState(snap, dt=60.0)
    want = orc.interact(True, record=True, cap=1 << 22)
    assert len(want["events"]) > min_calls and (want["events"][:, 2] == 0).sum() > 0
    evs = []
    eng, out = step_once(engine3d, snap, 60.0, 6.674e-11, resolve_mode=resolve_mode, stage_events=evs)
    assert eng.params(60.0, 6.674e-11).resolve_mode == resolve_mode
    print("3-D stage C, n=%d, %d flagged pairs, %d absorb calls, resolve_mode %d: %.3f ms" % (
        n, len(want["flag_pairs"]), len(want["events"]), resolve_mode, evs[0][2].elapsed_time(evs[0][3])))
    assert np.array_equal(out["flag_pairs"], want["flag_pairs"])
    assert np.array_equal(eng.events()[:, 1:], want["events"])
    d, post = eng.download(), orc.snapshot()
    assert np.array_equal(d["id"], post["id"])
    assert np.array_equal(d["mass"], post["mass_f64"]) and np.array_equal(d["density"], post["dens_f64"])
    assert np.array_equal(d["mass_hi"], post["mass_int_hi"]) and np.array_equal(d["mass_lo"], post["mass_int_lo"])
    assert np.abs(d["pos"] - post["pos"]).max() <= 1e-12 * np.abs(post["pos"]).max()
    assert np.abs(d["vel"] - post["vel"]).max() <= 1e-10 * np.abs(post["vel"]).max()


@pytest.mark.parametrize("thick", [0.02, 0.6])
def test_grid_broad_phase_equals_all_pairs(mods, thick):
    """N = 40000 (grid broad phase on the (x, y) projection, thin disk and thick cloud): flagged pairs
    and absorb stream must equal the all-pairs kernel's and the oracle's flagged set."""
    engine3d, o3, _ = mods
    n = 40000
    snap = synthetic_cloud(n, 9, thick=thick)
    snap["pos"] *= 0.05
    snap["dens_f64"][:] = 100.0
    arr = engine3d_arrays_from_snapshot(snap)
    res = []
    for force_all_pairs in (False, True):
        eng = engine3d.Engine3D(n)
        if force_all_pairs:
            eng.grid_threshold = 1 << 30
        eng.upload(arr)
        assert (eng.grid is None) == force_all_pairs
        eng.step(60.0, 6.674e-11)
        st = eng.check_errors()
        res.append((eng.last_step_tensors(n), eng.events(), eng.download(), st))
    (ta, ea, da, sa), (tb, eb, db, sb) = res
    assert len(ta["flag_pairs"]) > 50
    assert np.array_equal(ta["flag_pairs"], tb["flag_pairs"]) and np.array_equal(ea, eb)
    assert np.array_equal(da["id"], db["id"]) and np.array_equal(da["pos"], db["pos"])
    want = o3.flags(ta["pos_new"], arr["radius"])
    assert np.array_equal(ta["flag_pairs"], want)


@pytest.mark.parametrize("sym", [False, True])
def test_fp32_flavour_accuracy(mods, sym):
    """FP32 all-pairs sum against the FP64 oracle: direct differences after an exact power-of-two
    rescaling, so the error is the FP32 rounding of the pair terms: median < 1e-5, 99 % < 1e-4,
    max < 1e-2 of |a| (rows with a near neighbour carry the largest relative error)."""
    engine3d, o3, _ = mods
    n = 8192
    snap = synthetic_cloud(n, 3)
    eng, out = step_once(engine3d, snap, 60.0, 6.674e-11, collisions=False, fp32=True, sym=sym)
    assert bool(eng.params(60.0, 6.674e-11, False, True, sym=sym).opts & _abi_bits(mods).OPT3_F32_SYM) == sym
    want = o3.accel(snap["pos"], snap["mass_f64"], 6.674e-11)
    err = relerr_rows(out["acc"], want)
    assert np.median(err) < 1e-5 and np.quantile(err, 0.99) < 1e-4 and err.max() < 1e-2
    a = out["acc"]
    v = snap["vel"] + a * 60.0
    assert np.array_equal(out["vel_new"], v) and np.array_equal(out["pos_new"], snap["pos"] + v * 60.0)


def crowded_thin_disk(n):
    """Thin disk with 200 planted pairs 1e6..1e8 m apart (an FP32 coordinate at 1 AU is known to ~1e4 m)."""
    snap = synthetic_cloud(n, 11, thick=0.001)
    rng = np.random.default_rng(5)
    twins = rng.choice(np.arange(1, n), size=400, replace=False)      # object k gets a close companion
    for q, k in enumerate(twins[:200]):
        d = rng.normal(size=3)
        snap["pos"][twins[200 + q]] = snap["pos"][k] + d / np.linalg.norm(d) * 10.0 ** rng.uniform(6.0, 8.0)
    return snap


@pytest.mark.parametrize("sym", [False, True])
def test_fp32_near_field_3d_emulated_ranks(mods, sym):
    """The sharded protocol of the 3-D FP32 step with the near field on, emulated on one GPU: every "rank"
    runs stage A on its share (symmetric kernel: row blocks and cell-sorted near-field chunks dealt round
    robin, partial sums added like the all-reduce does; one-sided kernel: its own row range with its own
    corrections) -- the result must be the single-GPU accelerations to 1e-12 (every near pair corrected
    exactly once, whatever the rank count)."""
    import ctypes as C
    import torch
    from cosmosim_b200 import sharded
    engine3d, o3, abi = mods
    n, G = 40960, 6.674e-11
    arr = engine3d_arrays_from_snapshot(crowded_thin_disk(n))
    eng = engine3d.Engine3D(n)
    eng.upload(arr)
    eng.step(60.0, G, collisions=False, fp32=True, store_acc=True, sym=sym)
    eng.check_errors()
    assert eng.params(60.0, G, False, True, sym=sym).near_mode == 2
    base = eng.last_step_tensors(n)["acc"]
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for world in (2, 3):
        e2 = engine3d.Engine3D(n)
        e2.upload(arr)
        total = np.zeros((n, 3))
        for r in range(world):
            p = e2.params(60.0, G, False, True, True, sym=sym)
            assert p.near_mode == 2
            p.shard_index, p.shard_count = r, world
            i0, i1, _ = sharded.shard_rows(n, world, r)
            if sym:
                p.opts |= abi.OPT3_DEFER_INTEGRATE
            else:
                p.i_begin, p.i_end = i0, i1
            abi.check(e2.lib.cosmo3_accel_integrate(C.byref(e2.st), C.byref(e2.sc), C.byref(p), stream), "cosmo3_accel_integrate")
            acc = e2.t["acc"][:, :n].cpu().numpy().T
            if sym:
                total += acc
            else:
                total[i0:i1] = acc[i0:i1]
        e2.check_errors()
        got = total * G if sym else total
        assert np.abs(got - base).max() <= 1e-12 * np.abs(base).max(), (world, sym)


@pytest.mark.parametrize("sym", [False, True])
def test_fp32_near_field_3d(mods, sym):
    """FP64 near field of the 3-D FP32 flavours (csrc/nearfield.cu, k_nearfield3): a thin crowded disk
    with planted close pairs (separations of 1e6..1e8 m at 1 AU, where an FP32 coordinate is known to
    ~1e4 m).  With the pass on the stated bound of the 2-D path holds for every row (median <= 1e-5,
    max <= 1e-3 of |a|); switched off, the planted rows are far outside it.  The second step takes the
    grid the first step's stage B left on the device (near_mode 1), the first tunes its own (2)."""
    engine3d, o3, _ = mods
    n = 40960
    snap = crowded_thin_disk(n)
    G = 6.674e-11
    want = o3.accel(snap["pos"], snap["mass_f64"], G)
    arr = engine3d_arrays_from_snapshot(snap)
    errs = {}
    for near in (True, False):
        eng = engine3d.Engine3D(n)
        eng.upload(arr)
        eng.step(60.0, G, collisions=False, fp32=True, store_acc=True, sym=sym, near=near)
        eng.check_errors()
        p = eng.params(60.0, G, False, True, sym=sym, near=near)
        assert bool(p.opts & _abi_bits(mods).OPT3_F32_SYM) == sym and p.near_mode == (2 if near else 0)
        errs[near] = relerr_rows(eng.last_step_tensors(n)["acc"], want)
    on, off = errs[True], errs[False]
    assert np.median(on) < 1e-5 and on.max() < 1e-3, (np.median(on), on.max())
    assert off.max() > 10 * on.max(), (off.max(), on.max())
    # collisions on (grid broad phase): the next step bins the near field with stage B's grid
    eng = engine3d.Engine3D(n)
    eng.grid_threshold = 1024
    eng.upload(arr)
    eng.step(60.0, G, collisions=True, fp32=True, store_acc=True, sym=sym)
    assert eng.params(60.0, G, True, True, sym=sym).near_mode == 1
    st = eng.check_errors()
    d = eng.download()
    m = st["n_active"]
    want2 = o3.accel(d["pos"], d["mass"], G)
    eng.step(60.0, G, collisions=False, fp32=True, store_acc=True, sym=sym)
    eng.check_errors()
    err2 = relerr_rows(eng.last_step_tensors(m)["acc"], want2)
    assert np.median(err2) < 1e-5 and err2.max() < 1e-3, (np.median(err2), err2.max())


@pytest.mark.parametrize("fp32", [False, True])
def test_tuned_sass_equals_ptxas_schedule_3d(mods, fp32):
    """The symmetric 3-D kernels run from post-scheduled copies of their SASS (cosmosim_b200/sass_tune.py);
    same instructions, same arithmetic: accelerations and new state must be the SAME BITS as those of the
    kernels as ptxas scheduled them (COSMO_SYM_TUNED=0)."""
    import os
    engine3d, o3, _ = mods
    n = 20000 if not fp32 else 40000
    arr = engine3d_arrays_from_snapshot(synthetic_cloud(n, 21, thick=0.05))
    old = os.environ.get("COSMO_SYM_TUNED")
    got = {}
    try:
        for tuned in ("0", "1"):
            os.environ["COSMO_SYM_TUNED"] = tuned
            eng = engine3d.Engine3D(n)
            eng.upload(arr)
            eng.step(60.0, 6.674e-11, collisions=False, fp32=fp32, store_acc=True, sym=True, near=False)
            eng.check_errors()
            t = eng.last_step_tensors(n)
            got[tuned] = np.concatenate([t[k] for k in ("acc", "pos_new", "vel_new")], axis=1)
    finally:
        if old is None:
            os.environ.pop("COSMO_SYM_TUNED", None)
        else:
            os.environ["COSMO_SYM_TUNED"] = old
    assert np.isfinite(got["1"]).all()
    assert np.array_equal(got["0"].view(np.uint64), got["1"].view(np.uint64))


def test_host_api_run_and_pickles(mods, tmp_path):
    """Object / State / Universe.run through the drop-in module: in-memory run vs the reference
    fixture, and the on-disk pickle stream (universe.py:151-158) read back."""
    from cosmosim_b200.core import state as S
    g = load_golden("state3d_inmemory.npz")
    init = snap3_from(g, "init_")
    masses, dens = python_numbers3(init)

    def build():
        return [S.Object(m, q, p, v, name="o%d" % k, color=(1, 2, 3))
                for k, (m, q, p, v) in enumerate(zip(masses, dens, init["pos"], init["vel"]))]

    st = S.State(build(), dt=float(g["dt"]))
    st.step(int(g["steps"]))
    want = snap3_from(g, "final_")
    assert st.iteration == int(g["steps"]) and len(st.objects) == len(want["id"])
    assert np.abs(st.positions() - want["pos"]).max() <= 1e-9 * np.abs(want["pos"]).max()
    assert st.absorb_calls == [tuple(int(x) for x in e) for e in g["events"]]
    assert all(isinstance(o.mass, int) for o in st.objects[1:]) and isinstance(st.objects[0].mass, float)

    out = str(tmp_path / "run") + os.sep
    S.Universe(build(), float(g["dt"]), 5, outpath=out, filesize=3).run()
    assert sorted(os.listdir(out)) == ["0.dat", "1.dat"]
    states = S.load_states(out)
    assert sorted(s.iteration for s in states) == [1, 2, 3, 4, 5]
    mem = S.Universe(build(), float(g["dt"]), 5).run()
    last = [s for s in states if s.iteration == 5][0]
    assert np.array_equal(last.positions(), mem[-1].positions())
    assert b"ccosmosim.core.universe\nState\n" in open(os.path.join(out, "0.dat"), "rb").read()
    # bursts between save points: the device runs ahead, the saved states are the same ones
    some = S.Universe(build(), float(g["dt"]), 5).run(save_every=2)
    assert [s.iteration for s in some] == [2, 4, 5]
    for s in some:
        assert np.array_equal(s.positions(), mem[s.iteration - 1].positions())
    # the caller's list is pruned in place, like the reference's State.objects (universe.py:126-128), so a
    # second run on the same list starts from the survivors (ints stay ints, nothing sits on the star)
    objs = build()
    st2 = S.State(objs, dt=float(g["dt"]))
    st2.step(int(g["steps"]))
    assert st2.objects is objs and len(objs) == len(want["id"]) and all(o.exists for o in objs)
    again = S.Universe(objs, float(g["dt"]), 1).run()
    assert len(again[0].objects) <= len(want["id"]) and all(isinstance(o.mass, int) for o in again[0].objects[1:])
    # a host-side edit of the step counter survives the next step
    st2.iteration = 100
    st2.step(1)
    assert st2.iteration == 101


@pytest.mark.parametrize("n", [1, 2, 3, 63, 64, 65, 1023, 1024, 1025, 2049])
def test_ragged_sizes_against_oracle(mods, n):
    """Object counts around the tile / padding boundaries (incl. a lone object): one crowded step must
    equal the oracle's (pairs, absorb stream, survivors)."""
    engine3d, o3, _ = mods
    snap = synthetic_cloud(max(n, 2), 100 + n, thick=0.3)
    snap = {k: v[:n].copy() for k, v in snap.items()}
    snap["pos"] *= 1e-3 * max(1.0, n / 64.0) ** (1.0 / 3.0)   # constant density: radii ~1e7 m, plenty of contacts
    snap["vel"] *= 0.0
    snap["mass_f64"][:] = np.random.default_rng(n).uniform(0.1, 100.0, n) * 5.972e24
    orc = o3.OracleState(snap, dt=10.0)
    want = orc.interact(True, record=True, cap=1 << 18)
    eng, out = step_once(engine3d, snap, 10.0, 6.674e-11)
    assert np.array_equal(out["flag_pairs"], want["flag_pairs"])
    assert np.array_equal(eng.events()[:, 1:].reshape(-1, 3), want["events"].reshape(-1, 3))
    d, post = eng.download(), orc.snapshot()
    assert d["status"]["n_active"] == orc.n and d["status"]["iteration"] == 1
    assert np.array_equal(d["id"], post["id"]) and np.array_equal(d["mass"], post["mass_f64"])
    if n > 1:
        assert relerr_rows(out["acc"], want["acc"]).max() < 1e-11
        assert np.abs(d["pos"] - post["pos"]).max() <= 1e-12 * np.abs(post["pos"]).max()
    else:
        assert np.array_equal(out["acc"], np.zeros((1, 3))) and np.array_equal(d["pos"], post["pos"])


def test_empty_state_and_collisions_off(mods):
    engine3d, o3, _ = mods
    eng = engine3d.Engine3D(0)
    snap = synthetic_cloud(2, 1)
    empty = {k: v[:0] for k, v in snap.items()}
    eng.upload(engine3d_arrays_from_snapshot(empty))
    eng.step(60.0, 6.674e-11, steps=3)
    st = eng.check_errors()
    assert st["n_active"] == 0 and st["iteration"] == 3 and st["n_events"] == 0
    # collisions=False: overlapping objects pass through each other (universe.py:106,115)
    snap = synthetic_cloud(500, 2)
    snap["pos"] *= 1e-4
    eng, out = step_once(engine3d, snap, 1.0, 6.674e-11, collisions=False)
    st = eng.check_errors()
    assert st["n_active"] == 500 and st["n_events"] == 0 and st["n_pairs"] == 0


def test_overflow_is_reported_not_silent(mods):
    """More flagged pairs than the pair buffer holds: the step must raise, never drop pairs silently."""
    engine3d, o3, abi = mods
    n = 600
    snap = synthetic_cloud(n, 5)
    snap["pos"] *= 1e-6                       # everything overlaps everything: n (n - 1) flagged pairs
    snap["vel"] *= 0.0
    dt = 1e-6                                 # (the pull of the 1 M_sun object at 1e5 m would scatter them in 1 s)
    arr = engine3d_arrays_from_snapshot(snap)
    eng = engine3d.Engine3D(n, pair_cap=1024)
    eng.upload(arr)
    eng.step(dt, 6.674e-11)
    with pytest.raises(abi.CosmoError, match="pair buffer overflow"):
        eng.check_errors()
    # with room for all pairs the same scene collapses into the heaviest row objects, like the oracle
    eng = engine3d.Engine3D(n, pair_cap=n * n)
    eng.upload(arr)
    eng.step(dt, 6.674e-11)
    orc = o3.OracleState(snap, dt=dt)
    want = orc.interact(True, cap=1 << 20)
    assert want["n_pairs"] == n * (n - 1)
    assert eng.check_errors()["n_active"] == orc.n
    assert np.array_equal(eng.events()[:, 1:], want["events"])


def test_bulk_state_from_arrays_equals_object_state(mods):
    """State.from_arrays (no Python Object per body) steps exactly like a State built from Objects,
    reports the same absorb calls, and can be turned into Objects mid-run."""
    from cosmosim_b200.core import state as S
    g = load_golden("state3d_disk_planar.npz")
    init = snap3_from(g, "init_")
    masses, dens = python_numbers3(init)
    objs = [S.Object(m, q, p, v, name="o%d" % k, color=(1, 2, 3))
            for k, (m, q, p, v) in enumerate(zip(masses, dens, init["pos"], init["vel"]))]
    a = S.State(objs, dt=float(g["dt"]))
    b = S.State.from_arrays(masses, dens, init["pos"], init["vel"], dt=float(g["dt"]))
    assert np.array_equal(a.radii(), b.radii()) and np.array_equal(b.positions(), init["pos"])
    a.step(12)
    b.step(12)
    assert np.array_equal(a.positions(), b.positions()) and np.array_equal(a.velocities(), b.velocities())
    want = [tuple(int(x) for x in e) for e in g["events"] if e[0] < 12]
    assert a.absorb_calls == want and b.absorb_calls == want and b.iteration == 12
    assert b._objects is None                                   # still arrays only
    bo = b.objects                                              # materialise mid-run
    assert len(bo) == len(a.objects) and [type(o.mass) for o in bo] == [type(o.mass) for o in a.objects]
    assert [o._cosmo_index for o in bo] == [o._cosmo_index for o in a.objects]
    a.step(8)
    b.step(8)
    assert np.array_equal(a.positions(), b.positions())
    want = [tuple(int(x) for x in e) for e in g["events"] if e[0] < 20]
    assert a.absorb_calls == want and b.absorb_calls == want


@pytest.mark.parametrize("sym", [False, True])
def test_savedata_scene_free_run(mods, sym):
    """examples/testing/save-data.py at its real size (10001 objects, dt = 600) against the reference's own
    run: 4 steps, 66 absorb calls; one-sided kernel (the default at this size) and the symmetric one."""
    engine3d, o3, _ = mods
    g = load_golden("state3d_savedata.npz")
    arr = engine3d_arrays_from_snapshot(snap3_from(g, "init_"))
    n = len(arr["mass_f64"])
    assert np.array_equal(arr["radius"], g["s0_radius"])
    eng = engine3d.Engine3D(n)
    eng.upload(arr)
    counts = []
    for s in range(int(g["steps"])):
        eng.step(float(g["dt"]), float(g["G"]), store_acc=(s == 0), sym=sym)
        if s == 0:
            out = eng.last_step_tensors(n)
            assert relerr_rows(out["acc"], g["s0_acc"]).max() < 1e-12
            assert np.array_equal(out["flag_pairs"], g["s0_flag_pairs"])
        counts.append(eng.status()["n_active"])
    eng.check_errors()
    assert np.array_equal(np.array(counts), g["n_objects"])
    assert np.array_equal(eng.events(), g["events"])
    d, want = eng.download(), snap3_from(g, "final_")
    assert np.array_equal(d["id"], want["id"]) and np.array_equal(d["mass"], want["mass_f64"])
    assert np.array_equal(d["density"], want["dens_f64"])
    assert np.abs(d["pos"] - want["pos"]).max() <= 1e-11 * np.abs(want["pos"]).max()
