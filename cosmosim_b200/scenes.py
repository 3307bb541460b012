"""Seeded scene builders: the BASELINE example scripts as functions, and bulk synthetic states.

star_with_disk / cloud restate reference examples/star_with_disk.py:12-43 and
examples/cloud.py:6-16 on this package's Universe (same RNG draw order, so a seed gives the
reference's initial state bit-for-bit -- tests/test_host_api.py pins that against fixtures
produced by the reference).  The *_arrays builders make the large synthetic configurations
of BASELINE.json directly as SoA numpy arrays (no Python object per planet).
"""
from __future__ import annotations

import math
import random

import numpy as np

from .core.planet import AU, ME, MS, RS
from .core.universe import Universe
from .util import functions as F


def star_with_disk(seed=0, n_planets=1000, **universe_kw):
    random.seed(seed)
    u = Universe(**universe_kw)
    star = u.create_planet(mass=MS, radius=RS, position=[0, 0], immobile=True, name="Sol", color=(255, 255, 0))
    for _ in range(n_planets):
        dist = random.randint(int(0.05 * AU), int(1 * AU))
        mass = random.randint(int(0.1 * ME), int(100 * ME))
        star.create_satellite(distance=dist, mass=mass, radius=F.get_radius(mass, 3000))
    return u


def cloud(seed=0, n_planets=1000, omega=1e-9, **universe_kw):
    random.seed(seed)
    u = Universe(**universe_kw)
    for _ in range(n_planets):
        p = u.random_planet()
        p.velocity = F.rotation(p.position * omega, math.pi / 2)
    return u


def _finish(pos, vel, mass, radius, immobile):
    n = mass.shape[0]
    flags = np.zeros(n, dtype=np.uint8)
    flags[immobile] = 1
    return {"pos": np.ascontiguousarray(pos), "vel": np.ascontiguousarray(vel), "mass_f64": mass,
            "mass_hi": np.zeros(n, dtype=np.uint64), "mass_lo": np.zeros(n, dtype=np.uint64),
            "flags": flags, "radius": radius, "volume": (4.0 / 3.0) * math.pi * radius ** 3,
            "eaten": np.zeros(n, dtype=np.int64), "id": np.arange(n, dtype=np.int32)}


def star_disk_arrays(n, seed=0, G=6.674e-11, density=3000.0, r_in=0.05, r_out=1.0):
    """Immobile 1 M_sun star at slot 0 + n-1 float-mass planets on circular orbits, radius
    uniform in [r_in, r_out] AU like examples/star_with_disk.py (BASELINE configs 3 and 5)."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(r_in, r_out, n - 1) * AU
    th = rng.uniform(0.0, 2.0 * np.pi, n - 1)
    pos = np.zeros((n, 2))
    pos[1:, 0], pos[1:, 1] = r * np.cos(th), r * np.sin(th)
    speed = np.sqrt(G * MS / r)
    vel = np.zeros((n, 2))
    vel[1:, 0], vel[1:, 1] = -speed * np.sin(th), speed * np.cos(th)
    mass = np.empty(n)
    mass[0] = MS
    mass[1:] = rng.uniform(0.1, 100.0, n - 1) * ME
    radius = np.cbrt(3.0 * (mass / density) / (4.0 * np.pi))
    radius[0] = RS
    return _finish(pos, vel, mass, radius, np.array([0]))


def cloud_arrays(n, seed=0, omega=1e-9, density=4000.0, r_out=50.0):
    """Rotating cloud like examples/cloud.py scaled up (BASELINE config 4)."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(0.0, r_out, n) * AU
    th = rng.uniform(0.0, 2.0 * np.pi, n)
    pos = np.stack([r * np.sin(th), r * np.cos(th)], axis=1)
    vel = np.stack([-pos[:, 1] * omega, pos[:, 0] * omega], axis=1)
    mass = rng.uniform(0.1, 1000.0, n) * ME
    radius = np.cbrt(3.0 * (mass / density) / (4.0 * np.pi))
    return _finish(pos, vel, mass, radius, np.array([], dtype=np.int64))


# ---------------------------------------------------------------------------------------------
# 3-D (new API) scenes
# ---------------------------------------------------------------------------------------------
def star_with_disk_objects(seed=0, n_planets=1000, d_min=0.1 * AU, d_max=2 * AU, density=3000, star_density=1408):
    """reference examples/testing/in-memory.py:12-42 (defaults; save-data.py uses 10000 planets between
    0.1 and 0.5 AU) on this package's Object: same RNG draw order, so a seed gives the reference's
    initial object list."""
    from .core.state import Object
    random.seed(seed)
    star = Object(mass=MS, density=star_density, position=[0, 0, 0], name="Sol", color=(255, 255, 0))
    planets = []
    for _ in range(n_planets):
        dist = random.randint(int(d_min), int(d_max))
        mass = random.randint(int(0.1 * ME), int(100 * ME))
        planets.append(star.create_satellite(distance=dist, mass=mass, density=density))
    return [star, *planets]


def disk3d_arrays(n, seed=0, G=6.674e-11, density=3000.0, r_in=0.1, r_out=2.0, thickness=0.02):
    """Bulk synthetic version of the same scene for Engine3D.upload: 1 M_sun star at slot 0 + n-1
    float-mass planets on circular orbits between r_in and r_out AU, lifted out of the plane by a
    Gaussian z offset of `thickness` x radius (the example scripts are planar; State is 3-D)."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(r_in, r_out, n - 1) * AU
    th = rng.uniform(0.0, 2.0 * np.pi, n - 1)
    pos = np.zeros((n, 3))
    pos[1:, 0], pos[1:, 1] = r * np.cos(th), r * np.sin(th)
    pos[1:, 2] = rng.normal(size=n - 1) * thickness * r
    speed = np.sqrt(G * MS / r)
    vel = np.zeros((n, 3))
    vel[1:, 0], vel[1:, 1] = -speed * np.sin(th), speed * np.cos(th)
    mass = np.empty(n)
    mass[0] = MS
    mass[1:] = rng.uniform(0.1, 100.0, n - 1) * ME
    dens = np.full(n, float(density))
    dens[0] = 1408.0
    radius = np.array([((3 * (m / d)) / (4 * math.pi)) ** (1 / 3) for m, d in zip(mass.tolist(), dens.tolist())])
    return {"pos": pos, "vel": vel, "mass_f64": mass, "mass_hi": np.zeros(n, dtype=np.uint64),
            "mass_lo": np.zeros(n, dtype=np.uint64), "density": dens, "radius": radius,
            "flags": np.zeros(n, dtype=np.uint8), "id": np.arange(n, dtype=np.int32)}
