"""Host-side ``Planet``: the reference's public object (cosmosim/core/planet.py:18-91) kept as
a thin proxy over the device-resident SoA state.

Attributes that the physics update mutates (position, velocity, mass, radius, volume, alive,
planets_eaten, energy) are properties: reading one pulls the device state back if the GPU has
stepped since the last sync; writing one marks the host copy as the newer one so the next
step re-uploads.  Everything else (name, color, tracked, ...) is plain data.
"""
from __future__ import annotations

import math
import random
import secrets

import numpy as np

from ..util import functions as F

AU = 1.496e11
ME = 5.972e24
RE = 6.371e6
MS = 1.989e30
RS = 6.9634e8

_SYNCED = ("position", "velocity", "mass", "radius", "volume", "alive", "planets_eaten", "energy")

_CONS = "bcdfghjklmnprstvwz"
_VOWS = "aeiou"


def generate_name():
    """Pronounceable random name.  Uses `secrets`, never `random`: the reference's name
    generator (cosmosim/util/pronounceable/main.py:99-104) does not consume the seeded RNG
    stream either, so scenes built under random.seed() stay identical."""
    n = 2 + secrets.randbelow(3)
    word = "".join(secrets.choice(_CONS) + secrets.choice(_VOWS) for _ in range(n))
    return word.capitalize()


def _synced_property(name):
    slot = "_" + name

    def getter(self):
        u = self.universe
        if u is not None and u._device_ahead:
            u._pull()
        return getattr(self, slot)

    def setter(self, value):
        u = self.universe
        if u is not None:
            if u._device_ahead:
                u._pull()
            u._host_dirty = True
        if name in ("position", "velocity") and value is not None:
            value = np.array(value).astype(float)
        setattr(self, slot, value)

    return property(getter, setter)


class Planet:
    """Planet(mass, radius, position, velocity=[0,0], color=None, immobile=False, name=None)."""

    def __init__(self, mass, radius, position, velocity=(0.0, 0.0), color=None, immobile=False,
                 name=None, universe=None):
        self.universe = None  # set by Universe.add_planet (planet.py:22, universe_old.py:43)
        self.name = name
        self._mass = mass
        self._radius = radius
        self._volume = ((4 / 3) * math.pi) * (radius ** 3)
        self.density = mass / self._volume
        self._position = np.array(position).astype(float)
        self._velocity = np.array(velocity).astype(float)
        self._energy = 0
        self.color = color
        self._immobile = immobile
        self.history = []
        self.bound = False
        self.clicked = False
        self.tracked = False
        self._alive = True
        self._planets_eaten = 0
        self.index = -1  # insertion index in Universe.planets
        if not self.name:
            self.name = generate_name()
        if not self.color:
            # three draws from the seeded stream, as the reference does (planet.py:41-42)
            self.color = (int(255 * random.random()), int(255 * random.random()), int(255 * random.random()))

    @property
    def immobile(self):
        return self._immobile

    @immobile.setter
    def immobile(self, value):
        u = self.universe
        if u is not None:
            if u._device_ahead:
                u._pull()
            u._host_dirty = True
        self._immobile = value

    def create_satellite(self, distance=None, mass=None, radius=None, theta=None, name=None, color=None):
        """Planet on a circular counter-clockwise orbit of this body's mass (planet.py:44-63).

        Faithful to the reference's conventions: the orbit is centred on the ORIGIN and the
        speed sqrt(M G / distance) is not offset by this planet's own velocity; random draws
        happen in the order distance, mass, theta."""
        G = self.universe.G
        if not distance:
            distance = random.randint(int(self.radius * 5), int(self.radius * 100))
        if not mass:
            mass = random.random() * self.mass
        if not radius:
            radius = F.get_radius(mass, 1)
        if theta is None:
            theta = 2 * math.pi * random.random()
        speed = math.sqrt((self.mass * G) / distance)
        pos = F.to_cartesian(distance, theta)[::-1]
        direction = pos / np.linalg.norm(pos)
        vel = F.rotation(speed * direction, math.pi / 2)
        return self.universe.create_planet(mass=mass, radius=radius, position=pos, velocity=vel,
                                           name=name, color=color)

    def update_history(self, trail_length):
        if self.alive and self.tracked:
            self.history.append(np.array([self.position[0], self.position[1]]))
        else:
            self.history = []
        if len(self.history) > trail_length:
            del self.history[0]

    def absorb(self, planet):
        """Host-side inelastic merge (planet.py:74-85) for user code that merges by hand; the
        stepping path performs the same arithmetic on the device."""
        if self.alive and planet.alive:
            if not self.immobile:
                total = self.mass + planet.mass
                self.position = ((self.mass * self.position) + (planet.mass * planet.position)) / total
                self.velocity = ((self.mass * self.velocity) + (planet.mass * planet.velocity)) / total
            self.mass += planet.mass
            self.volume += planet.volume
            self.radius = ((3 * self.volume) / (4 * math.pi)) ** (1 / 3)
            planet.destroy()
            self.planets_eaten += 1 + planet.planets_eaten

    def destroy(self):
        """planet.py:86-91: dead planets keep mass and radius, lose position."""
        self.alive = False
        self.velocity = np.array([0, 0])
        self.acceleration = np.array([0, 0])
        self.position = None
        self.energy = 0

    def __repr__(self):
        return "Planet(%r, mass=%r, alive=%r)" % (self.name, self._mass, self._alive)


for _name in _SYNCED:
    setattr(Planet, _name, _synced_property(_name))
