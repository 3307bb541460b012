"""Optional pygame front-end on top of the device-resident state.

Not part of the timed path.  It mirrors the interactive half of the reference's Universe
(cosmosim/core/universe_old.py:89-164 input handling, :214-304 HUD and planet panel, :338-362 frame
loop) and Planet.draw / update_history (cosmosim/core/planet.py:65-72,93-118), re-designed around
the fact that the planets live on the GPU:

* per displayed frame ONE small device->host copy of (id, position, radius[, energy]) of the active
  planets (`Engine.view`) feeds drawing, hit-testing and the HUD -- the Planet objects are not
  refreshed, so a million circles do not cost a million Python attribute syncs;
* the HUD counts come from the device status block (active, merged so far);
* only a CLICKED or TRACKED planet is materialised as a Planet object (panel text, trail history);
* "DESTROY" is a host-initiated kill: Planet.destroy() marks the host copy newer, and the next frame
  re-uploads the survivors (exactly what any host-side mutation does).

pygame is imported lazily, so the package works without it; tests drive this module with the no-op
pygame stand-in of oracle/pygame_stub and scripted events.
"""
from __future__ import annotations

import numpy as np

WHITE, BLACK = (255, 255, 255), (0, 0, 0)
ME = 5.972e24
# planet panel geometry (universe_old.py:231-239)
INFO_X, INFO_Y, INFO_WIDTH, INFO_HEIGHT, BTN_SIZE, TXT_PAD = 10, 150, 370, 150, 30, 10


def screen_coordinates(p, scale, offset, origin):
    """World -> screen, vectorised over planets (cosmosim/util/functions.py:48-49)."""
    return origin + np.multiply(p, np.array([1, -1])) * scale + offset * scale


def hit_test(pos, radius, context, click, min_pixels=10.0):
    """Index of the first active planet under `click` or -1 (universe_old.py:147: a planet is hit
    within max(radius * scale, 10 px) of its screen position; first in list order wins)."""
    if len(radius) == 0:
        return -1
    q = screen_coordinates(np.asarray(pos, dtype=float), **context)
    d = np.linalg.norm(q - np.asarray(click, dtype=float), axis=1)
    hit = np.nonzero(d <= np.maximum(np.asarray(radius) * context["scale"], min_pixels))[0]
    return int(hit[0]) if hit.size else -1


def panel_buttons():
    """Rectangles (x, y, w, h) of the planet panel's close / track / destroy buttons."""
    x2, y2 = INFO_X + INFO_WIDTH, INFO_Y + INFO_HEIGHT
    return {"close": (x2 - BTN_SIZE, INFO_Y, BTN_SIZE, BTN_SIZE),
            "track": (INFO_X, y2, INFO_WIDTH / 2, BTN_SIZE),
            "destroy": (INFO_X + INFO_WIDTH / 2, y2, INFO_WIDTH / 2, BTN_SIZE)}


def _inside(rect, pos):
    x, y, w, h = rect
    return x <= pos[0] < x + w and y <= pos[1] < y + h


class Viewer:
    def __init__(self, universe, pygame_module=None):
        if pygame_module is None:
            try:
                import pygame as pygame_module
            except ImportError as e:  # pragma: no cover - pygame is not part of the GPU image
                raise RuntimeError("simulate(headless=False) needs pygame; use headless=True or step()") from e
        self.pg = pygame_module
        self.u = universe
        self.mouse_xy = (0, 0)
        self.view = {"n": 0, "id": np.zeros(0, np.int32), "pos": np.zeros((0, 2)), "radius": np.zeros(0)}
        self.hud = []
        pg = self.pg
        pg.init()
        self.clock = pg.time.Clock()
        self.font = pg.font.SysFont(None, 24)
        self.screen = pg.display.set_mode([universe.width, universe.height])
        pg.display.set_caption("cosmosim_b200")

    # ------------------------------------------------------------------ input (universe_old.py:89-164)
    def handle_user_input(self, event):
        pg, u, ctx = self.pg, self.u, self.u.context
        if event.type == pg.QUIT:
            u.running = False
        elif event.type == pg.KEYDOWN:
            if event.key == pg.K_SPACE:
                u.paused = not u.paused
            if event.key == pg.K_w:
                ctx["offset"] = ctx["offset"] + np.array([0, 50]) / ctx["scale"]
            if event.key == pg.K_a:
                ctx["offset"] = ctx["offset"] + np.array([50, 0]) / ctx["scale"]
            if event.key == pg.K_s:
                ctx["offset"] = ctx["offset"] + np.array([0, -50]) / ctx["scale"]
            if event.key == pg.K_d:
                ctx["offset"] = ctx["offset"] + np.array([-50, 0]) / ctx["scale"]
            if event.key == pg.K_KP_PLUS:
                u.dt *= 1.05
            if event.key == pg.K_KP_MINUS:
                u.dt *= 0.95
        elif event.type == pg.MOUSEBUTTONDOWN:
            if event.button == 1:        # left: drag the view
                u.dragging = True
                self.mouse_xy = tuple(event.pos)
            elif event.button == 3:      # right: reset view and speed
                ctx["offset"] = np.array([0.0, 0.0])
                ctx["scale"] = u.default_scale
                u.dt = u.dt_default
            elif event.button == 4:      # wheel: zoom 5 %
                ctx["scale"] *= 1.05
            elif event.button == 5:
                ctx["scale"] *= 0.95
        elif event.type == pg.MOUSEBUTTONUP:
            if event.button == 1:
                u.dragging = False
                self._click(tuple(getattr(event, "pos", None) or self.pg.mouse.get_pos()))
        elif event.type == pg.MOUSEMOTION:
            if u.dragging:
                x0, y0 = event.pos
                dx = int((x0 - self.mouse_xy[0]) / (5 * ctx["scale"]))
                dy = int((y0 - self.mouse_xy[1]) / (5 * ctx["scale"]))
                ctx["offset"] = ctx["offset"] + np.array([dx, dy])

    def _click(self, pos):
        u = self.u
        cp = u.clicked_planet
        buttons = panel_buttons() if cp is not None else {}
        if cp is not None and _inside(buttons["close"], pos):
            cp.clicked = False
            u.clicked_planet = None
        elif cp is not None and _inside(buttons["track"], pos):
            cp.tracked = not cp.tracked
        elif cp is not None and _inside(buttons["destroy"], pos):
            cp.clicked = False
            cp.tracked = False
            cp.destroy()                 # host-initiated kill: the next frame re-uploads the survivors
            u.clicked_planet = None
        else:
            k = hit_test(self.view["pos"], self.view["radius"], u.context, pos)
            if k >= 0:
                if cp is not None:
                    cp.clicked = False
                u.clicked_planet = u.planets[int(self.view["id"][k])]
                u.clicked_planet.clicked = True

    # ------------------------------------------------------------------ one displayed frame
    def frame(self):
        """screen clear -> input -> one physics frame -> draw -> HUD -> flip (universe_old.py:338-361)."""
        pg, u = self.pg, self.u
        self.screen.fill(BLACK)
        for event in pg.event.get():
            self.handle_user_input(event)
        u.step(1)                        # a paused universe only counts the frame (universe_old.py:361)
        self.view = u.view()
        self._update_histories()
        self._draw_planets()
        self.hud = self.universe_info_text()
        for i, text in enumerate(self.hud):
            self.screen.blit(self.font.render(text, True, WHITE), (20, 20 * (i + 1)))
        self._draw_clicked_planet_panel()
        self._draw_simulation_text()
        pg.display.flip()
        self.clock.tick(u.fps)

    def _update_histories(self):
        """Trails (planet.py:65-72) for the tracked planets only; positions come from the frame's view."""
        u, v = self.u, self.view
        tracked = [p for p in u.planets if p.tracked or p.history]
        if not tracked:
            return
        slot = {int(i): k for k, i in enumerate(v["id"])}
        for p in tracked:
            k = slot.get(p.index)
            if p.tracked and k is not None:
                p.history.append(v["pos"][k].copy())
                if len(p.history) > u.trail_length:
                    del p.history[0]
            else:
                p.history = []

    def _draw_planets(self):
        pg, u, v, ctx = self.pg, self.u, self.view, self.u.context
        if v["n"] == 0:
            return
        q = screen_coordinates(v["pos"], **ctx).astype(int)
        rad = np.maximum(1, (v["radius"] * ctx["scale"]).astype(int))
        # only what is on the screen is handed to pygame
        on = np.nonzero((q[:, 0] + rad >= 0) & (q[:, 0] - rad < u.width) & (q[:, 1] + rad >= 0) & (q[:, 1] - rad < u.height))[0]
        for k in on:
            p = u.planets[int(v["id"][k])]
            if p.clicked:                # highlight ring (planet.py:100-103)
                pg.draw.circle(self.screen, WHITE, tuple(q[k]), max(int(rad[k] * 2), 4))
                pg.draw.circle(self.screen, BLACK, tuple(q[k]), max(int(rad[k] * 1.85), 2))
            if p.tracked and len(p.history) > 1:      # fading trail (planet.py:104-116)
                h = screen_coordinates(np.array(p.history), **ctx).astype(int)
                for i in range(1, len(h)):
                    alpha = i / len(h)
                    pg.draw.line(self.screen, tuple(alpha * c for c in p.color), tuple(h[i - 1]), tuple(h[i]), 1)
            pg.draw.circle(self.screen, p.color, tuple(q[k]), int(rad[k]))

    # ------------------------------------------------------------------ HUD (universe_old.py:214-304)
    def universe_info_text(self):
        u, v = self.u, self.view
        n_active = int(v["n"])
        lines = ["Total Planets: %d" % n_active]
        if "energy" in v:
            bound = int((v["energy"] < 0).sum())
            lines += ["Bound Planets: %d" % bound, "Escaped Planets: %d" % (n_active - bound)]
        else:
            lines += ["Bound Planets: n/a (track_energy=False)", "Escaped Planets: n/a"]
        lines.append("Destroyed Planets: %d" % (len(u.planets) - n_active))
        if "energy" in v:
            lines += ["Net Angular Momentum: {:.2e}".format(v["angular_momentum"]),
                      "Total Energy: {:.2e}".format(v["total_energy"])]
        return lines

    def _slot_of(self, planet):
        k = np.nonzero(self.view["id"] == planet.index)[0]
        return int(k[0]) if k.size else -1

    def clicked_planet_text(self):
        """Panel lines of the clicked planet from ONE small device read of its slot (the Planet objects
        stay unsynchronised)."""
        p = self.u.clicked_planet
        k = self._slot_of(p) if p is not None else -1
        if k < 0:
            return []
        s = self.u.peek(k)
        vel = s["vel"]
        return [str(p.name).upper(), "Mass: {:.2e} Earth masses".format(s["mass"] / ME),
                "Radius: {:.2e} km".format(s["radius"] / 1000),
                "Velocity: [{:.2e}, {:.2e}] {:.2e} m/s".format(vel[0], vel[1], float(np.linalg.norm(vel))),
                "Energy: {:.2e}".format(round(s["energy"], 2)), "Planets eaten: %i" % s["eaten"]]

    def _draw_clicked_planet_panel(self):
        pg, u = self.pg, self.u
        p = u.clicked_planet
        if p is None:
            return
        if self._slot_of(p) < 0:         # merged or escaped since it was clicked
            p.clicked = False
            u.clicked_planet = None
            return
        x2, y2 = INFO_X + INFO_WIDTH, INFO_Y + INFO_HEIGHT
        b = panel_buttons()
        pg.draw.rect(self.screen, WHITE, ((INFO_X, INFO_Y), (INFO_WIDTH, INFO_HEIGHT)), 2)
        pg.draw.line(self.screen, WHITE, (INFO_X, INFO_Y + BTN_SIZE), (x2, INFO_Y + BTN_SIZE), 1)
        pg.draw.rect(self.screen, WHITE, (b["close"][:2], b["close"][2:]), 2)
        self.screen.blit(self.font.render("X", True, WHITE), (x2 - BTN_SIZE + TXT_PAD, INFO_Y + TXT_PAD))
        pg.draw.rect(self.screen, WHITE, (b["track"][:2], b["track"][2:]), 0 if p.tracked else 2)
        self.screen.blit(self.font.render("TRACK", True, BLACK if p.tracked else WHITE), (INFO_X + TXT_PAD, y2 + TXT_PAD - 1))
        pg.draw.rect(self.screen, WHITE, (b["destroy"][:2], b["destroy"][2:]), 2)
        self.screen.blit(self.font.render("DESTROY", True, WHITE), (INFO_X + INFO_WIDTH / 2 + TXT_PAD, y2 + TXT_PAD - 1))
        for i, text in enumerate(self.clicked_planet_text()):
            y = INFO_Y + TXT_PAD if i == 0 else INFO_Y + 2 * TXT_PAD + 20 * i
            self.screen.blit(self.font.render(text, True, WHITE), (20, y))

    def _draw_simulation_text(self):
        u = self.u
        fps = 1.0 / ((self.clock.get_time() + 1) / 1000)
        for i, text in enumerate(("FPS: %.0f" % fps, "Scale: {:.2e}".format(u.context["scale"]),
                                  "Speed: %.2f" % round(u.speed, 1))):
            self.screen.blit(self.font.render(text, True, WHITE), (u.width * 0.90, 20 * (i + 1)))

    # ------------------------------------------------------------------ loop
    def run(self, run_while=lambda u: True):
        u = self.u
        u.running = True
        while u.running and run_while(u):
            self.frame()
        self.pg.quit()
        u.running = False
        return u


def run(universe, run_while=lambda u: True, pygame_module=None):
    return Viewer(universe, pygame_module).run(run_while)
