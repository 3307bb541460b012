"""The optional front-end (cosmosim_b200/viewer.py; reference universe_old.py:89-164,214-304 and
planet.py:65-72,93-118).  Pure geometry on the CPU; the interactive flow on the GPU, driven headlessly
through the no-op pygame stand-in of oracle/pygame_stub with scripted events."""
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AU = 1.496e11


def stub_pygame():
    path = os.path.join(ROOT, "oracle", "pygame_stub")
    sys.path.insert(0, path)
    try:
        sys.modules.pop("pygame", None)
        import pygame
    finally:
        sys.path.remove(path)
        sys.modules.pop("pygame", None)
    return pygame


def ev(pg, kind, **kw):
    return types.SimpleNamespace(type=getattr(pg, kind), **kw)


def test_screen_geometry_and_hit_test():
    from cosmosim_b200 import viewer
    ctx = {"scale": 1e-9, "offset": np.array([0.0, 0.0]), "origin": np.array([800.0, 500.0])}
    pos = np.array([[0.0, 0.0], [1.0 * AU, 0.0], [0.0, 2.0 * AU]])
    q = viewer.screen_coordinates(pos, **ctx)
    assert np.allclose(q, [[800, 500], [800 + 149.6, 500], [800, 500 - 299.2]])      # y points up in the world
    rad = np.array([7e8, 6e6, 6e6])
    assert viewer.hit_test(pos, rad, ctx, (800, 500)) == 0
    assert viewer.hit_test(pos, rad, ctx, (955, 505)) == 1          # within the 10-pixel minimum
    assert viewer.hit_test(pos, rad, ctx, (900, 300)) == -1
    assert viewer.hit_test(np.zeros((0, 2)), np.zeros(0), ctx, (0, 0)) == -1
    b = viewer.panel_buttons()
    assert b["track"][0] + b["track"][2] == b["destroy"][0] and b["close"][1] == viewer.INFO_Y


@pytest.mark.gpu
def test_interactive_flow_with_scripted_events():
    from cosmosim_b200 import Universe, viewer
    pg = stub_pygame()
    u = Universe(track_energy=True)
    u.create_planet(mass=1.989e30, radius=6.9634e8, position=[0, 0], immobile=True, name="sun", color=(255, 255, 0))
    planets = [u.create_planet(mass=5.972e24 * (k + 1), radius=6.4e6, position=[(0.3 + 0.2 * k) * AU, 0.0],
                               velocity=[0.0, 3.0e4], name="p%d" % k, color=(0, 100 + 20 * k, 255)) for k in range(4)]
    script = []
    pg.event.get = lambda: script.pop(0) if script else []
    # configure like simulate() does, then drive the viewer frame by frame
    u.collisions, u.trail_length, u.default_scale, u.width, u.height = True, 5, 1e-9, 1600, 1000
    u.context = {"scale": 1e-9, "offset": np.array([0.0, 0.0]), "origin": np.array([800.0, 500.0])}
    v = viewer.Viewer(u, pg)
    v.frame()
    assert u.iterations == 1 and v.view["n"] == 5 and v.hud[0] == "Total Planets: 5"
    assert v.hud[1].startswith("Bound Planets: ") and v.hud[3] == "Destroyed Planets: 0" and "Total Energy" in v.hud[5]
    # keys: pause / resume, speed, pan; mouse: zoom, reset
    dt0 = u.dt
    script.append([ev(pg, "KEYDOWN", key=pg.K_SPACE)])
    v.frame()
    frames_on_device = u._frames
    assert u.paused and u.iterations == 2
    script.append([ev(pg, "KEYDOWN", key=pg.K_SPACE), ev(pg, "KEYDOWN", key=pg.K_KP_PLUS), ev(pg, "KEYDOWN", key=pg.K_w),
                   ev(pg, "MOUSEBUTTONDOWN", button=4, pos=(0, 0))])
    v.frame()
    assert not u.paused and u._frames == frames_on_device + 1 and u.dt == dt0 * 1.05
    assert np.allclose(u.context["offset"], [0, 50 / 1e-9]) and u.context["scale"] == 1e-9 * 1.05
    script.append([ev(pg, "MOUSEBUTTONDOWN", button=3, pos=(0, 0))])
    v.frame()
    assert u.dt == dt0 and u.context["scale"] == 1e-9 and np.allclose(u.context["offset"], 0)
    # click on the second planet -> panel; TRACK -> trail grows (capped by trail_length); DESTROY -> gone
    k = int(np.nonzero(v.view["id"] == planets[1].index)[0][0])
    q = viewer.screen_coordinates(v.view["pos"][k], **u.context)
    script.append([ev(pg, "MOUSEBUTTONDOWN", button=1, pos=tuple(q)), ev(pg, "MOUSEBUTTONUP", button=1, pos=tuple(q + 3))])
    v.frame()
    assert u.clicked_planet is planets[1] and planets[1].clicked and v.clicked_planet_text()[0] == "P1"
    b = viewer.panel_buttons()
    centre = lambda r: (r[0] + r[2] / 2, r[1] + r[3] / 2)  # noqa: E731
    script.append([ev(pg, "MOUSEBUTTONUP", button=1, pos=centre(b["track"]))])
    for _ in range(8):
        v.frame()
    assert planets[1].tracked and len(planets[1].history) == 5
    assert np.allclose(planets[1].history[-1], planets[1].position)
    n_before = v.view["n"]
    script.append([ev(pg, "MOUSEBUTTONUP", button=1, pos=centre(b["destroy"]))])
    v.frame()
    assert u.clicked_planet is None and not planets[1].alive and v.view["n"] == n_before - 1
    assert v.hud[0] == "Total Planets: %d" % (n_before - 1) and v.hud[3] == "Destroyed Planets: 1"
    assert planets[1].index not in set(int(i) for i in v.view["id"])
    # close button on another planet, window close ends the loop
    k = int(np.nonzero(v.view["id"] == planets[2].index)[0][0])
    q = viewer.screen_coordinates(v.view["pos"][k], **u.context)
    script.append([ev(pg, "MOUSEBUTTONUP", button=1, pos=tuple(q))])
    v.frame()
    assert u.clicked_planet is planets[2]
    script.append([ev(pg, "MOUSEBUTTONUP", button=1, pos=centre(b["close"]))])
    v.frame()
    assert u.clicked_planet is None and not planets[2].clicked
    script.append([ev(pg, "QUIT")])
    it = u.iterations
    v.run(run_while=lambda U: U.iterations < it + 50)
    assert not u.running and u.iterations == it + 1


@pytest.mark.gpu
def test_simulate_with_a_window_runs_the_viewer():
    """simulate(headless=False) goes through the viewer with the reference's run_while hook."""
    from cosmosim_b200 import scenes, viewer
    pg = stub_pygame()
    sys.modules["pygame"] = pg
    try:
        u = scenes.star_with_disk(0, n_planets=50)
        u.simulate(scale=2e-9, run_while=lambda U: U.iterations < 12, headless=False)
        assert u.iterations == 12 and not u.running and len(u.get_active_planets()) >= 40
    finally:
        sys.modules.pop("pygame", None)
    assert viewer.run is not None
