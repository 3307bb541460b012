"""Generate tests/golden/*.npz by running the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py [--steps 1000] [--out tests/golden]

Fixtures written (all produced by reference code, none by this repo's oracle or kernels):
  scene_<name>_init.npz    initial SoA state of the seeded example scene (pins the RNG draw
                           order of the scene builders)
  scene_<name>_run.npz     1000-step free run: full merge stream (step, absorber, absorbed),
                           final state, per-step active counts
  scene_<name>_steps.npz   a handful of single steps with full pre-state, acc_blas output,
                           new positions, flagged pairs and post-state (per-step parity)
  micro_merges.npz         the six 2-3 body merge micro-scenes of SURVEY.md Appendix B
  accel_n4001.npz          acc_blas on a seeded synthetic disk, N = 4001
  scene_disk6001.npz       star + 6000 planets (radii x4 so that merges are plentiful), 3 frames of the
                           reference: initial state, merge stream, final state, acc / flagged pairs of frame 0
  energies.npz             total_energy / angular_momentum / Planet.energy over six frames
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402

SNAP_KEYS = ("pos", "vel", "mass_f64", "mass_int_hi", "mass_int_lo", "mass_is_int", "radius",
             "volume", "alive", "immobile", "eaten")


def flat(prefix, snap):
    return {f"{prefix}{k}": snap[k] for k in SNAP_KEYS}


def run_scene(name, builder, seed, steps, out_dir, n_keep_merge=4, n_keep_plain=3):
    t0 = time.time()
    u = builder(seed)
    init = rh.snapshot(u)
    np.savez_compressed(os.path.join(out_dir, f"scene_{name}_init.npz"), seed=seed, G=u.G,
                        r_max=u.r_max, bounded=u.bounded, **flat("", init))
    # pass 1: events only
    rec = rh.Recorder(u).run(steps)
    events = np.array(rec.events, dtype=np.int64).reshape(-1, 3)
    final = rh.snapshot(u)
    np.savez_compressed(os.path.join(out_dir, f"scene_{name}_run.npz"), seed=seed, steps=steps,
                        dt=1e6 / 33, events=events, **flat("final_", final))
    print(f"[{name}] pass 1: {steps} steps, {len(events)} merges, {time.time() - t0:.0f}s", flush=True)

    # pass 2: re-run the same seed and keep selected steps in full
    merge_steps = sorted(set(int(s) for s in events[:, 0]))
    # prefer steps that show both absorber orders (higher index absorbs / lower index absorbs)
    hi = [int(s) for s, a, b in events if a > b]
    lo = [int(s) for s, a, b in events if a < b]
    keep = []
    for pool in (hi, lo, merge_steps):
        for s in pool:
            if s not in keep and len(keep) < n_keep_merge:
                keep.append(s)
                break
    for s in merge_steps:
        if len(keep) >= n_keep_merge:
            break
        if s not in keep:
            keep.append(s)
    plain = [s for s in (0, 1, steps // 2, steps - 1) if s not in merge_steps][:n_keep_plain]
    keep = sorted(set(keep + plain))
    u2 = builder(seed)
    last = max(keep) + 1
    rec2 = rh.Recorder(u2, keep_steps=keep).run(last + 1)
    ev2 = np.array(rec2.events, dtype=np.int64).reshape(-1, 3)
    assert np.array_equal(ev2, events[events[:, 0] <= last]), "reference run is not reproducible"
    blob = {"steps": np.array(keep, dtype=np.int64), "dt": 1e6 / 33, "G": u2.G}
    for s in keep:
        st = rec2.steps[s]
        blob.update(flat(f"s{s}_pre_", st["pre"]))
        blob.update(flat(f"s{s}_post_", st["post"]))
        blob[f"s{s}_active_idx"] = st["active_idx"]
        blob[f"s{s}_acc"] = st["acc"]
        blob[f"s{s}_p_new"] = st["p_new"]
        blob[f"s{s}_flag_pairs"] = st["flag_pairs"]
        blob[f"s{s}_events"] = events[events[:, 0] == s][:, 1:]
    np.savez_compressed(os.path.join(out_dir, f"scene_{name}_steps.npz"), **blob)
    print(f"[{name}] pass 2: kept steps {keep}, {time.time() - t0:.0f}s", flush=True)


def micro_merges(out_dir):
    """Six micro-scenes (G=1, dt=1, bodies on the x axis) run through the reference."""
    uo, pl, fn, bl = rh.load_reference()
    scenes = {
        # name: list of (mass, radius, x, vx, immobile)
        "heavy_low_index": [(10, 1.0, 0.0, 0.0, False), (3, 1.0, 1.5, 0.0, False)],
        "heavy_high_index": [(3, 1.0, 0.0, 0.0, False), (10, 1.0, 1.5, 0.0, False)],
        "tie_lower_wins": [(5, 1.0, 0.0, 0.0, False), (5, 1.0, 1.5, 0.0, False)],
        "chain_live_masses": [(4, 1.0, 0.0, 0.0, False), (3, 1.0, 1.5, 0.0, False), (6, 1.0, 3.0, 0.0, False)],
        "immobile_absorber": [(10.0, 1.0, 0.0, 0.0, True), (3, 1.0, 1.5, 0.25, False)],
        "immobile_absorbed": [(3.0, 1.0, 0.0, 0.0, True), (10, 1.0, 1.5, 0.25, False)],
        "float_masses": [(2.5, 1.0, 0.0, 0.125, False), (2.5, 1.0, 1.5, -0.5, False), (0.75, 0.5, 40.0, 0.0, False)],
    }
    blob = {"names": np.array(list(scenes))}
    for name, bodies in scenes.items():
        u = uo.Universe(G=1.0)
        for (m, r, x, vx, imm) in bodies:
            u.create_planet(mass=m, radius=r, position=[x, 0.25 * x], velocity=[vx, 0.0],
                            immobile=imm, color=(1, 1, 1), name="b")
        pre = rh.snapshot(u)
        rec = rh.Recorder(u, keep_steps=[0, 1]).run(2, speed=1.0, fps=1)
        blob.update(flat(f"{name}_pre_", pre))
        blob.update(flat(f"{name}_post1_", rec.steps[0]["post"]))
        blob.update(flat(f"{name}_post2_", rec.steps[1]["post"]))
        blob[f"{name}_events"] = np.array(rec.events, dtype=np.int64).reshape(-1, 3)
        blob[f"{name}_acc0"] = rec.steps[0]["acc"]
    np.savez_compressed(os.path.join(out_dir, "micro_merges.npz"), **blob)
    print("[micro] done", flush=True)


def energy_fixture(out_dir, frames=6):
    """HUD numbers of the reference (universe_old.py:181-189,210): total_energy, angular_momentum
    and Planet.energy after each of a few frames of two small seeded scenes."""
    uo, pl, fn, bl = rh.load_reference()
    blob = {}
    for name, builder, seed, n in (("disk", rh.build_star_with_disk, 11, 300), ("cloud", rh.build_cloud, 12, 300)):
        u = builder(seed, n)
        if name == "cloud":                      # make a few merges happen in six frames
            for p in u.planets:
                p.radius *= 300.0
                p.volume = ((4 / 3) * np.pi) * (p.radius ** 3)
        blob.update(flat(f"{name}_init_", rh.snapshot(u)))
        tot, ang, en, alive = [], [], [], []

        def cond(U):
            if U.iterations > 0:
                tot.append(float(U.total_energy)); ang.append(float(U.angular_momentum))
                en.append(np.array([float(p.energy) for p in U.planets]))
                alive.append(np.array([bool(p.alive) for p in U.planets]))
            return U.iterations < frames
        u.simulate(run_while=cond)
        blob[f"{name}_total_energy"] = np.array(tot); blob[f"{name}_angular_momentum"] = np.array(ang)
        blob[f"{name}_energy"] = np.array(en); blob[f"{name}_alive"] = np.array(alive)
        blob[f"{name}_G"] = u.G
    np.savez_compressed(os.path.join(out_dir, "energies.npz"), dt=1e6 / 33, frames=frames, **blob)
    print("[energy] done", flush=True)


def synthetic_disk(n, seed):
    """Config-3 style synthetic disk (float masses): shared with the tests via the fixture."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(0.05, 1.0, n - 1) * rh.AU
    th = rng.uniform(0, 2 * np.pi, n - 1)
    pos = np.zeros((n, 2))
    pos[1:, 0] = r * np.cos(th)
    pos[1:, 1] = r * np.sin(th)
    mass = np.empty(n)
    mass[0] = rh.MS
    mass[1:] = rng.uniform(0.1, 100.0, n - 1) * rh.ME
    return pos, mass


def disk6001(out_dir, n_planets=6000, seed=3, frames=3, inflate=4.0):
    """A scene beyond the small-N kernel shapes (N > 4096), run by the reference itself."""
    t0 = time.time()
    u = rh.build_star_with_disk(seed, n_planets)
    for p in u.planets[1:]:
        p.radius *= inflate
        p.volume = ((4 / 3) * np.pi) * (p.radius ** 3)
    init = rh.snapshot(u)
    rec = rh.Recorder(u, keep_steps=[0]).run(frames)
    events = np.array(rec.events, dtype=np.int64).reshape(-1, 3)
    blob = {"dt": 1e6 / 33, "G": u.G, "frames": frames, "events": events, "s0_acc": rec.steps[0]["acc"],
            "s0_flag_pairs": rec.steps[0]["flag_pairs"], "s0_active_idx": rec.steps[0]["active_idx"]}
    blob.update(flat("init_", init))
    blob.update(flat("final_", rh.snapshot(u)))
    np.savez_compressed(os.path.join(out_dir, "scene_disk6001.npz"), **blob)
    print(f"[disk6001] {frames} frames, {len(events)} merges, {len(rec.steps[0]['flag_pairs'])} flagged pairs in "
          f"frame 0, {time.time() - t0:.0f}s", flush=True)


def accel_fixture(out_dir, n=4001, seed=7):
    uo, pl, fn, bl = rh.load_reference()
    pos, mass = synthetic_disk(n, seed)
    acc = bl.acc_blas(pos, mass, 6.674e-11)[:, :2]
    np.savez_compressed(os.path.join(out_dir, f"accel_n{n}.npz"), pos=pos, mass=mass,
                        G=6.674e-11, acc=np.ascontiguousarray(acc), seed=seed)
    print("[accel] done", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                   "tests", "golden"))
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    todo = a.only.split(",") if a.only else ["micro", "accel", "energy", "disk", "cloud", "disk6001"]
    if "disk6001" in todo:
        disk6001(a.out)
    if "energy" in todo:
        energy_fixture(a.out)
    if "micro" in todo:
        micro_merges(a.out)
    if "accel" in todo:
        accel_fixture(a.out)
    if "disk" in todo:
        run_scene("star_with_disk", rh.build_star_with_disk, 0, a.steps, a.out)
    if "cloud" in todo:
        run_scene("cloud", rh.build_cloud, 0, a.steps, a.out)
