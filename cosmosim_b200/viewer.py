"""Optional pygame front-end on top of the device-resident state.

Not part of the timed path.  Mirrors the reference's frame loop (universe_old.py:338-362): one
physics frame per displayed frame, planets drawn as circles, SPACE pauses, keypad +/- scale dt,
window close stops.  Needs pygame; imported lazily so the package works without it."""
from __future__ import annotations

import numpy as np


def run(universe, run_while=lambda u: True):
    try:
        import pygame
    except ImportError as e:  # pragma: no cover - pygame is not part of the GPU image
        raise RuntimeError("simulate(headless=False) needs pygame; use headless=True or step()") from e
    pygame.init()
    screen = pygame.display.set_mode([universe.width, universe.height])
    pygame.display.set_caption("cosmosim_b200")
    clock = pygame.time.Clock()
    universe.running = True
    ctx = universe.context
    while universe.running and run_while(universe):
        for event in pygame.event.get():
            if event.type == pygame.QUIT:
                universe.running = False
            elif event.type == pygame.KEYDOWN:
                if event.key == pygame.K_SPACE:
                    universe.paused = not universe.paused
                elif event.key == pygame.K_KP_PLUS:
                    universe.dt *= 1.05
                elif event.key == pygame.K_KP_MINUS:
                    universe.dt *= 0.95
        universe.step(1)
        screen.fill((0, 0, 0))
        for p in universe.get_active_planets():       # pulls the device state once per frame
            q = ctx["origin"] + np.multiply(p.position, np.array([1, -1])) * ctx["scale"] + ctx["offset"] * ctx["scale"]
            pygame.draw.circle(screen, p.color, q.astype(int), max(1, int(p.radius * ctx["scale"])))
        pygame.display.flip()
        clock.tick(universe.fps)
    pygame.quit()
    universe.running = False
    return universe
