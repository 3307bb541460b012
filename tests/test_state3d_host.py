"""Host side of the 3-D API (cosmosim_b200/core/state.py, mirror of the reference's
cosmosim/core/universe.py) and the C ABI of include/cosmosim_b200_state3d.h -- no GPU."""
import ctypes as C
import io
import math
import pickle
import random

import numpy as np
import pytest

from cosmosim_b200 import _abi
from cosmosim_b200.core import state as S
from oracle import ref_harness as rh


def test_struct_layouts_and_argument_checks():
    assert C.sizeof(_abi.State3) == 8 + 11 * 8 + 8
    assert C.sizeof(_abi.StepParams3) == 2 * 8 + 8 * 4 + 4 * 4 + 4 * 8 + 4 * 4
    assert C.sizeof(_abi.Scratch3) == 29 * 8 + 7 * 8 + 8 + 2 * 8 + 2 * 8 + 7 * 8 + 8    # + grid_dev, tune_hist, parallel stage C, corr
    L = _abi.lib()
    assert L.cosmo3_step(None, None, None, None) == -1
    assert b"bad argument" in L.cosmo_last_error()
    assert L.cosmo3_suggest_jsplit(65536, 65536, 0, 64) >= 1


def test_exact_helpers_follow_python():
    L = _abi.lib()
    rnd = random.Random(2)
    m64 = (1 << 64) - 1
    for _ in range(5000):
        a = rnd.getrandbits(rnd.randint(1, 127)) + 1
        b = rnd.getrandbits(rnd.randint(1, 100)) + 1
        assert L.cosmo3_host_truediv(a >> 64, a & m64, b >> 64, b & m64) == a / b
    off = 0
    for _ in range(5000):
        m = rnd.randint(int(0.1 * 5.972e24), int(100 * 5.972e24))
        d = rnd.choice([3000, 1408, 30, 2999.5])
        want = ((3 * (m / d)) / (4 * math.pi)) ** (1 / 3)
        got = L.cosmo3_host_radius(1, m >> 64, m & m64, float(m), int(isinstance(d, int)), float(d))
        assert abs(got - want) <= math.ulp(want)      # glibc pow is within an ulp of the correctly rounded value
        off += got != want
    assert off < 25


def test_object_semantics():
    a = S.Object(10, 3, [0, 0, 0], [1, 0, 0], name="a", color=(1, 1, 1))
    b = S.Object(5, 2, [3, 0, 0], name="b", color=(1, 1, 1))
    assert b.velocity.dtype == np.float64 and list(b.velocity) == [0, 0, 0]
    a.absorb(b)
    assert a.mass == 15 and isinstance(a.mass, int) and a.density == 40 / 15 and not b.exists and b.mass == 0.0
    assert np.array_equal(a.position, [1.0, 0, 0]) and np.array_equal(b.position, [0, 0, 0])
    a.absorb(b)                                   # the reference re-absorbs destroyed objects
    assert isinstance(a.mass, float) and a.mass == 15.0
    random.seed(4)
    star = S.Object(S.MS, 1408, [0, 0, 0], name="s", color=(1, 1, 1))
    sat = star.create_satellite(distance=int(S.AU), mass=int(S.ME), density=3000, name="p", color=(1, 1, 1))
    assert sat.position[2] == 0 and abs(np.linalg.norm(sat.position) - S.AU) < 1e-3
    assert abs(np.linalg.norm(sat.velocity) - math.sqrt(S.MS * 6.674e-11 / int(S.AU))) < 1e-6
    assert abs(np.dot(sat.position, sat.velocity)) < 1e-3 * S.AU


def test_pickle_stream_names_the_reference_classes():
    objs = [S.Object(7 + k, 3000, [k, 2 * k, -k], [0.5, 0, k], name="n%d" % k, color=(k, k, k)) for k in range(4)]
    st = S.State(objs, dt=60, iteration=3)
    buf = io.BytesIO()
    st.save(buf)
    st.save(buf)
    raw = buf.getvalue()
    assert raw.count(b"ccosmosim.core.universe\nState\n") == 2 and b"cosmosim_b200" not in raw
    buf.seek(0)
    back = [S._Unpickler(buf).load(), S._Unpickler(buf).load()]
    assert back[0].iteration == 3 and back[0].dt == 60 and len(back[0].objects) == 4
    assert np.array_equal(back[1].positions(), st.positions()) and back[1].objects[2].mass == 9


@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")
def test_pickles_load_with_the_reference_classes_and_back():
    from oracle import ref_harness3d as r3
    un = r3.load_reference3d()
    objs = [S.Object(7 + k, 3000, [k, 2 * k, -k], [0.5, 0, k], name="n%d" % k, color=(k, k, k)) for k in range(4)]
    buf = io.BytesIO()
    S.State(objs, dt=60, iteration=3).save(buf)
    buf.seek(0)
    ref_state = pickle.load(buf)                  # the way cosmosim/core/animation.py:41 reads it
    assert type(ref_state) is un.State and type(ref_state.objects[0]) is un.Object
    assert ref_state.dt == 60 and ref_state.iteration == 3
    assert np.array_equal(ref_state.positions(), np.array([o.position for o in objs]))
    assert ref_state.objects[1].get_radius() == objs[1].get_radius()
    buf2 = io.BytesIO()
    ref_state.save(buf2)                          # written by the reference, read here
    buf2.seek(0)
    mine = S._Unpickler(buf2).load()
    assert isinstance(mine, S.State) and mine.objects[3].mass == 10 and mine.iteration == 3


def test_scene_builder_reproduces_the_reference_initial_state():
    """scenes.star_with_disk_objects(seed) = examples/testing/in-memory.py:12-42 on this package's Object:
    positions, velocities, masses (Python ints) and densities equal the reference's for the same seed
    (fixture written by the unmodified reference, oracle/gen_golden3d.py)."""
    from conftest import load_golden, python_numbers3, snap3_from
    from cosmosim_b200 import scenes
    g = load_golden("state3d_inmemory.npz")
    init = snap3_from(g, "init_")
    objs = scenes.star_with_disk_objects(0, 1000)
    assert len(objs) == 1001
    assert np.array_equal(np.array([o.position for o in objs]), init["pos"])
    assert np.array_equal(np.array([o.velocity for o in objs]), init["vel"])
    masses, dens = python_numbers3(init)
    assert [o.mass for o in objs] == masses and [type(o.mass) for o in objs] == [type(m) for m in masses]
    assert [o.density for o in objs] == dens
    from cosmosim_b200.engine3d import arrays_from_objects
    arr = arrays_from_objects(objs)
    assert np.array_equal(arr["radius"], g["s0_radius"]) and arr["pos"].shape == (1001, 3)
    assert np.array_equal(arr["flags"] & 1, init["mass_is_int"].astype(np.uint8))
