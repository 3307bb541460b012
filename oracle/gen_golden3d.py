"""Generate tests/golden/state3d_*.npz by running the UNMODIFIED reference's new API
(cosmosim/core/universe.py) -- TEST INFRASTRUCTURE.  Build container only.

    python oracle/gen_golden3d.py [--out tests/golden]

Fixtures (all produced by reference code through oracle/ref_harness3d.py):
  state3d_disk_planar.npz   dense planar disk (300 planets, density 30, dt 600, 40 steps): init,
                            Object.absorb call stream (iteration, absorber, absorbed, existed),
                            objects per step, final state, selected steps in full
  state3d_disk_thick.npz    same with out-of-plane offsets (400 planets, density 20): true 3-D
  state3d_inmemory.npz      examples/testing/in-memory.py scene (seed 0), 200 of its 1000 steps
  state3d_savedata.npz      examples/testing/save-data.py scene (10000 planets, 0.1-0.5 AU, dt 600, seed 0),
                            4 of its 5000 steps: init, absorb stream, final state, acc / pairs of step 0
  state3d_micro.npz         2-3 body scenes for every branch of the absorb loop
  state3d_accel_n2001.npz   acc_blas on a seeded 3-D cloud
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness3d as r3  # noqa: E402

SNAP_KEYS = ("pos", "vel", "mass_f64", "mass_int_hi", "mass_int_lo", "mass_is_int", "dens_f64",
             "dens_is_int", "id")


def flat(prefix, snap):
    return {f"{prefix}{k}": snap[k] for k in SNAP_KEYS}


def run_scene(name, out_dir, steps, dt, n_keep=6, **build_kw):
    t0 = time.time()
    objs = r3.build_disk(**build_kw)
    init = r3.snapshot_with_ids(objs)
    rec = r3.Recorder3D(objs, dt).run(steps)
    events = np.array(rec.events, dtype=np.int64).reshape(-1, 4)
    final = r3.snapshot(rec.state.objects)
    # second pass: keep steps with dead re-absorbs, steps with merges, a few plain steps
    dead = sorted(set(int(e[0]) for e in events if not e[3]))
    merge = sorted(set(int(e[0]) for e in events))
    keep = []
    for pool in (dead, merge, [0, 1, steps - 1]):
        for s in pool:
            if s not in keep and len(keep) < n_keep:
                keep.append(s)
    keep = sorted(keep)
    objs2 = r3.build_disk(**build_kw)
    rec2 = r3.Recorder3D(objs2, dt, keep_steps=keep).run(max(keep) + 1)
    ev2 = np.array(rec2.events, dtype=np.int64).reshape(-1, 4)
    assert np.array_equal(ev2, events[events[:, 0] <= max(keep)]), "reference run is not reproducible"
    blob = {"dt": float(dt), "G": 6.674e-11, "steps": steps, "events": events,
            "n_objects": np.array(rec.n_objects, dtype=np.int64), "kept": np.array(keep, dtype=np.int64)}
    blob.update(flat("init_", init))
    blob.update(flat("final_", final))
    for s in keep:
        st = rec2.steps[s]
        blob.update(flat(f"s{s}_pre_", st["pre"]))
        blob.update(flat(f"s{s}_post_", st["post"]))
        for k in ("acc", "p_new", "radius", "flag_pairs"):
            blob[f"s{s}_{k}"] = st[k]
        blob[f"s{s}_events"] = events[events[:, 0] == s][:, 1:]
    np.savez_compressed(os.path.join(out_dir, f"state3d_{name}.npz"), **blob)
    print(f"[{name}] {steps} steps, {len(events)} absorb calls ({int((events[:, 3] == 0).sum())} on destroyed "
          f"objects), kept {keep}, {time.time() - t0:.0f}s", flush=True)


def savedata(out_dir, steps=4):
    """The reference needs ~40 s per step at this size (its absorb loop walks all N^2 flags in Python)."""
    t0 = time.time()
    kw = dict(seed=0, n_planets=10000, d_min=0.1 * r3.AU, d_max=0.5 * r3.AU, density=3000)
    objs = r3.build_disk(**kw)
    init = r3.snapshot_with_ids(objs)
    rec = r3.Recorder3D(objs, 600, keep_steps=[0]).run(steps)
    events = np.array(rec.events, dtype=np.int64).reshape(-1, 4)
    final = r3.snapshot(rec.state.objects)
    st0 = rec.steps[0]
    blob = {"dt": 600.0, "G": 6.674e-11, "steps": steps, "events": events,
            "n_objects": np.array(rec.n_objects, dtype=np.int64), "s0_acc": st0["acc"],
            "s0_flag_pairs": st0["flag_pairs"], "s0_radius": st0["radius"]}
    blob.update(flat("init_", init))
    blob.update(flat("final_", final))
    np.savez_compressed(os.path.join(out_dir, "state3d_savedata.npz"), **blob)
    print(f"[savedata] {steps} steps, {len(events)} absorb calls, {len(st0['flag_pairs'])} flagged pairs in step 0, "
          f"{time.time() - t0:.0f}s", flush=True)


def micro(out_dir):
    un = r3.load_reference3d()
    # (mass, density, x, vx): bodies on a line, radii ~ (3 m / (4 pi rho))^(1/3)
    scenes = {
        "row_heavier_absorbs": [(10, 1, 0.0, 0.0), (3, 1, 1.5, 0.0)],
        "row_lighter_waits": [(3, 1, 0.0, 0.0), (10, 1, 1.5, 0.0)],
        "tie_row_wins": [(5, 1, 0.0, 0.0), (5, 1, 1.5, 0.0)],
        "chain_grown_mass": [(4, 1, 0.0, 0.0), (3, 1, 1.2, 0.0), (6, 1, 2.4, 0.0)],
        "dead_reabsorbed": [(50, 1, 0.0, 0.0), (1, 1, 2.4, 0.0), (40, 1, 4.7, 0.0)],
        "float_masses": [(2.5, 1.5, 0.0, 0.125), (2.5, 0.5, 1.2, -0.5), (0.75, 1.0, 40.0, 0.0)],
        "int_mass_float_density": [(7, 0.75, 0.0, 0.0), (5, 2, 1.3, 0.25)],
        "big_int_masses": [(597200000000000032715571200, 3000, 0.0, 0.0),
                           (59720000000000003271557121, 3000, 3.0e6, 10.0),
                           (5972000000000000327155713, 2000, -2.5e6, -10.0)],
    }
    blob = {"names": np.array(list(scenes))}
    for name, bodies in scenes.items():
        objs = [un.Object(mass=m, density=d, position=[x, 0.25 * x, -0.5 * x], velocity=[vx, 0.0, 0.1 * vx])
                for (m, d, x, vx) in bodies]
        pre = r3.snapshot_with_ids(objs)
        rec = r3.Recorder3D(objs, 1.0, G=0.01 if name != "big_int_masses" else 6.674e-11, keep_steps=[0, 1]).run(2)
        blob.update(flat(f"{name}_pre_", pre))
        blob[f"{name}_G"] = rec.state.G
        for s in (0, 1):
            blob.update(flat(f"{name}_post{s}_", rec.steps[s]["post"]))
            blob[f"{name}_acc{s}"] = rec.steps[s]["acc"]
            blob[f"{name}_radius{s}"] = rec.steps[s]["radius"]
        blob[f"{name}_events"] = np.array(rec.events, dtype=np.int64).reshape(-1, 4)
    np.savez_compressed(os.path.join(out_dir, "state3d_micro.npz"), **blob)
    print("[micro]", {k: len(blob[f"{k}_events"]) for k in scenes}, flush=True)


def accel(out_dir, n=2001, seed=5):
    un = r3.load_reference3d()
    rng = np.random.default_rng(seed)
    pos = rng.normal(size=(n, 3)) * r3.AU * np.array([1.0, 0.7, 0.2])
    mass = rng.uniform(0.1, 100.0, n) * r3.ME
    mass[0] = r3.MS
    acc = un.acc_blas(pos, mass, 6.674e-11)
    np.savez_compressed(os.path.join(out_dir, f"state3d_accel_n{n}.npz"), pos=pos, mass=mass, G=6.674e-11,
                        acc=np.ascontiguousarray(acc))
    print("[accel] done", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                   "tests", "golden"))
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    todo = a.only.split(",") if a.only else ["micro", "accel", "planar", "thick", "inmemory", "savedata"]
    if "micro" in todo:
        micro(a.out)
    if "accel" in todo:
        accel(a.out)
    if "planar" in todo:
        run_scene("disk_planar", a.out, 40, 600, seed=1, n_planets=300, d_min=0.1 * r3.AU, d_max=0.12 * r3.AU,
                  density=30)
    if "thick" in todo:
        run_scene("disk_thick", a.out, 40, 600, seed=2, n_planets=400, d_min=0.1 * r3.AU, d_max=0.13 * r3.AU,
                  density=20, z_sigma=2e8)
    if "inmemory" in todo:
        run_scene("inmemory", a.out, 200, 60, n_keep=3, seed=0, n_planets=1000)
    if "savedata" in todo:
        savedata(a.out)
