"""Geometry helpers used by the scene generators (host only).

Same names, argument meaning and results as the reference's cosmosim/util/functions.py
(normalize :5, to_cartesian :11, rotation :22, get_tangent :37, get_radius :43); the
screen-coordinate helpers are UI and out of scope.
"""
from __future__ import annotations

import math

import numpy as np


def normalize(v):
    """v / |v| (v itself when |v| == 0)."""
    length = np.linalg.norm(v)
    return v if length == 0 else v / length


def to_cartesian(r, theta, origin=(0, 0)):
    """Polar -> [r sin(theta) + origin[0], r cos(theta) + origin[1]].

    Note the component order: the reference returns (sine, cosine) and its callers rely on
    it (random_planet keeps it, create_satellite reverses it)."""
    first = r * np.sin(theta) + origin[0]
    second = r * np.cos(theta) + origin[1]
    return [first, second]


def rotation(v, theta):
    """Rotate the 2-vector v counter-clockwise by theta."""
    c, s = np.cos(theta), np.sin(theta)
    return np.matmul(np.array([[c, -s], [s, c]]), v)


def get_tangent(position, reference=(0, 0)):
    """Unit vector perpendicular to (reference - position), rotated clockwise."""
    towards = normalize(np.asarray(reference) - position)
    return np.array([[0, 1], [-1, 0]]) @ towards


def get_radius(mass, density):
    """Radius of a homogeneous sphere of the given mass and density."""
    volume = mass / density
    return ((3 * volume) / (4 * math.pi)) ** (1 / 3)


def to_cartesian_3d(r, theta, phi, origin=(0, 0, 0)):
    """Spherical -> Cartesian with the reference's conventions (functions.py:16-20): theta is the
    azimuth measured as in `to_cartesian`, phi the polar angle from +z; the sine term takes origin[0]
    and the cosine term origin[1] (as in the 2-D helper), and the result is ordered [x, y, z] with
    x = the cosine term."""
    sin_term = r * np.sin(theta) * np.sin(phi) + origin[0]
    cos_term = r * np.cos(theta) * np.sin(phi) + origin[1]
    height = r * np.cos(phi) + origin[2]
    return [cos_term, sin_term, height]


def rotation_3d(v, theta, phi):
    """Rotate the 3-vector v by phi about the x axis, then by theta about the y axis
    (functions.py:27-36: R = R_y(theta) @ R_x(phi))."""
    ct, st, cp, sp = np.cos(theta), np.sin(theta), np.cos(phi), np.sin(phi)
    about_y = np.array([[ct, 0, st], [0, 1, 0], [-st, 0, ct]])
    about_x = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    return np.matmul(np.matmul(about_y, about_x), v)
