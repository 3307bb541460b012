"""Host-side mirror of the reference API: scene builders replay the reference's RNG draw
order (fixtures made by the reference), lazy-sync bookkeeping, drop-in import aliasing.  No GPU."""
import random

import numpy as np
import pytest

import cosmosim_b200
from cosmosim_b200 import Planet, Universe, scenes
from cosmosim_b200.engine import split_mass
from cosmosim_b200.util import functions as F
from conftest import load_golden


def universe_snapshot(u):
    ps = u.planets
    return {"pos": np.array([p._position for p in ps]), "vel": np.array([p._velocity for p in ps]),
            "mass": [p._mass for p in ps], "radius": np.array([p._radius for p in ps]),
            "volume": np.array([p._volume for p in ps]), "immobile": np.array([bool(p._immobile) for p in ps])}


@pytest.mark.parametrize("name,builder", [("star_with_disk", scenes.star_with_disk), ("cloud", scenes.cloud)])
def test_seeded_scene_equals_reference_initial_state(name, builder):
    g = load_golden(f"scene_{name}_init.npz")
    u = builder(int(g["seed"]))
    s = universe_snapshot(u)
    assert len(u.planets) == g["radius"].shape[0]
    assert np.array_equal(s["pos"], g["pos"])
    assert np.array_equal(s["vel"], g["vel"])
    assert np.array_equal(s["radius"], g["radius"])
    assert np.array_equal(s["volume"], g["volume"])
    assert np.array_equal(s["immobile"], g["immobile"])
    for k, m in enumerate(s["mass"]):
        f, is_int, hi, lo = split_mass(m)
        assert f == g["mass_f64"][k] and is_int == bool(g["mass_is_int"][k])
        assert hi == int(g["mass_int_hi"][k]) and lo == int(g["mass_int_lo"][k])
    assert u.G == float(g["G"]) and u.r_max == float(g["r_max"]) and u.bounded == bool(g["bounded"])


def test_gather_layout():
    u = scenes.star_with_disk(3, n_planets=20)
    u.planets[5]._alive = False
    a = u._gather()
    assert a["id"].tolist() == [k for k in range(21) if k != 5]
    assert a["flags"][0] == 1 and set(a["flags"][1:].tolist()) == {2}
    assert a["pos"].shape == (20, 2) and a["pos"].dtype == np.float64


def test_functions_helpers():
    assert F.get_radius(3000 * (4 / 3) * np.pi, 3000) == pytest.approx(1.0)
    y, x = F.to_cartesian(2.0, 0.0)
    assert (y, x) == (0.0, 2.0)
    assert np.allclose(F.rotation(np.array([1.0, 0.0]), np.pi / 2), [0.0, 1.0])
    assert np.allclose(F.normalize(np.array([3.0, 4.0])), [0.6, 0.8])
    assert np.array_equal(F.normalize(np.zeros(2)), np.zeros(2))
    # 3-D variants (reference functions.py:16-20,27-36): [x, y, z] with x the cosine term; R_y(theta) R_x(phi)
    assert np.allclose(F.to_cartesian_3d(2.0, 0.0, np.pi / 2), [2.0, 0.0, 0.0])
    assert np.allclose(F.to_cartesian_3d(2.0, np.pi / 2, np.pi / 2, origin=(1, 10, 100)), [10.0, 3.0, 100.0])
    assert np.allclose(F.to_cartesian_3d(2.0, 0.3, 0.0), [0.0, 0.0, 2.0])
    assert np.allclose(F.rotation_3d(np.array([0.0, 1.0, 0.0]), 0.0, np.pi / 2), [0.0, 0.0, 1.0])
    assert np.allclose(F.rotation_3d(np.array([0.0, 0.0, 1.0]), np.pi / 2, 0.0), [1.0, 0.0, 0.0])


def test_planet_defaults_and_rng_usage():
    random.seed(1)
    p = Planet(mass=10, radius=2.0, position=[1, 2])
    after = random.random()
    random.seed(1)
    [random.random() for _ in range(3)]          # colour = three draws, name = none
    assert after == random.random()
    assert p.volume == ((4 / 3) * np.pi) * (2.0 ** 3) and p.alive and p.planets_eaten == 0
    assert p.position.dtype == float and p.velocity.tolist() == [0.0, 0.0]
    q = Planet(mass=3, radius=1.0, position=[0, 0], color=(1, 2, 3), name="x")
    p.absorb(q)
    assert p.mass == 13 and not q.alive and q.position is None and p.planets_eaten == 1


def test_host_mutation_marks_dirty():
    u = Universe()
    p = u.create_planet(mass=1.0, radius=1.0, position=[0, 0], color=(1, 1, 1), name="a")
    u._host_dirty = False
    p.velocity = [1, 2]
    assert u._host_dirty and p.velocity.tolist() == [1.0, 2.0]
    assert p.universe is u and p.index == 0
    u.update_planet_lists()
    assert u.active_planets == [p] and u.absorbed_planets == [] and u.escaped_planets == []


def test_install_as_cosmosim_aliases():
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "cosmosim" or k.startswith("cosmosim.")}
    try:
        cosmosim_b200.install_as_cosmosim()
        from cosmosim.core.universe import Universe as U2
        from cosmosim.util.functions import get_radius
        import cosmosim.core.planet as pl
        assert U2 is Universe and pl.Planet is Planet and get_radius is F.get_radius
        cosmosim_b200.install_as_cosmosim(api="new")      # the 3-D batch API under the same import path
        from cosmosim.core.universe import Object, State
        from cosmosim_b200.core import state as S
        assert Object is S.Object and State is S.State
    finally:
        for k in [k for k in sys.modules if k == "cosmosim" or k.startswith("cosmosim.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_stepping_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    u = scenes.cloud(0, n_planets=4)
    with pytest.raises(Exception) as e:
        u.step(1)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_from_arrays_builds_the_same_gather_as_planet_objects():
    u1 = scenes.star_with_disk(4, n_planets=50)
    g1 = u1._gather()
    masses = np.array([p._mass for p in u1.planets], dtype=object)
    u2 = Universe.from_arrays(masses, g1["radius"], g1["pos"], g1["vel"], immobile=(g1["flags"] & 1) != 0)
    g2 = u2._gather()
    for k in g1:
        assert np.array_equal(g1[k], g2[k]), k
    assert u2.planets[3].universe is u2 and u2.planets[3].index == 3 and u2.planets[0].immobile
    u2.planets[7].velocity = [1.0, 2.0]          # setters still work on the slim proxies
    assert u2._host_dirty and u2.planets[7].velocity.tolist() == [1.0, 2.0]
