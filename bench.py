#!/usr/bin/env python
"""Contract benchmark of the per-timestep physics update (BASELINE.json north_star).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A "step" is one frame of the hot path (all-pairs acceleration + fused integrator + collision
detection + ordered merges + compaction) over one synthetic scene.  Default workload is the
configuration the metric is quoted on: star+disk, N = 1,048,576, FP32 fast path, merges on
(BASELINE.json configs[4]); with N > 1 GPUs the target rows are sharded (strong scaling,
same scene).  Prints ONE JSON line (rank 0).

Metric: ordered pairwise interactions per second = N_active^2 * K / time (the reference's
acc_blas also evaluates all N^2 ordered pairs).  `value` is timed with inputs resident in
HBM; `e2e` repeats the measurement through the host-buffer path (state H2D from pinned
memory + frame + results D2H every step).  `roofline` is for the dominant kernel (stage A),
timed with CUDA events on the launching stream inside the timed region, against the FMA-pipe
peak measured by the in-process micro-benchmark (MEASURED_PEAKS.json has no CUDA-core peak;
its HBM figure is used for the hbm sub-object).  `cpu_baseline` times the CPU oracle (a C port
of the reference's algorithm, all host threads) on a bounded row sample of the same scene.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

AU = 1.496e11
DT = 1e6 / 33
G = 6.674e-11
R_MAX = 100 * AU

WORKLOADS = {
    # name: (builder, n, precision, collisions)
    "star_disk_1m_fp32": ("star_disk", 1 << 20, "fp32", True),
    "star_disk_64k_fp64": ("star_disk", 1 << 16, "fp64", True),
    "cloud_256k_fp32": ("cloud", 1 << 18, "fp32", True),
    "star_disk_1001_fp64": ("star_disk", 1001, "fp64", True),
    # the reference's NEW (3-D, batch) API, State.interact: 1-GPU lines
    "state3d_64k_fp64": ("disk3d", 1 << 16, "fp64", True),
    "state3d_10k_fp64": ("disk3d", 10001, "fp64", True),     # size of examples/testing/save-data.py
    "state3d_256k_fp32": ("disk3d", 1 << 18, "fp32", True),
}
DT3 = 60.0          # dt of examples/testing/in-memory.py
# pipe instructions per ordered interaction of the 3-D kernels
INSTR3_PER_INTERACTION = {"fp32": 12.0, "fp64": 25.0}
# FMA-pipe instructions per ordered interaction of the dominant kernel
INSTR_PER_INTERACTION = {"fp32": 9, "fp64": 20}


def build_arrays(kind, n, seed=0):
    from cosmosim_b200 import scenes
    return scenes.star_disk_arrays(n, seed=seed) if kind == "star_disk" else scenes.cloud_arrays(n, seed=seed)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ CPU legs
def host_threads():
    """Threads the CPU arm uses: every core this process may run on, whatever OMP_NUM_THREADS says
    (torch.distributed.run exports OMP_NUM_THREADS=1, which must not shrink the reference arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_leg(arrays, prec_n, seconds_target=12.0, steps=1, warmup=0):
    """Time the CPU oracle (port of the reference's acc_blas + integrator + collision predicate)
    on a bounded sample of target rows of the same scene, all host threads."""
    from oracle import oracle as orc
    n = arrays["radius"].shape[0]
    pos, vel, mass, rad = arrays["pos"], arrays["vel"], arrays["mass_f64"], arrays["radius"]
    orc.set_threads(host_threads())
    cores = orc.num_threads()

    def sample(rows):
        t0 = time.perf_counter()
        a = orc.accel(pos, mass, G, 0, rows, compensated=False)
        p, v = orc.integrate(pos[:rows], vel[:rows], a, DT)
        pn = pos.copy()
        pn[:rows] = p
        orc.flags(pn, rad, 0, rows)
        return time.perf_counter() - t0

    rows = min(n, max(cores, 8))
    sample(rows)                                       # warms the library and the thread pool
    for _ in range(4):                                 # grow the sample towards the time target
        t = sample(rows)
        want = int(min(n, max(rows, seconds_target * rows / max(t, 1e-9))))
        if want <= rows * 1.25 or t >= 0.5 * seconds_target:
            break
        rows = min(want, rows * 16)
    times = []
    for _ in range(warmup):
        sample(rows)
    for _ in range(max(1, steps)):
        times.append(sample(rows))
    t = float(np.mean(times))
    return {"value": rows * n / t, "unit": "interactions/s", "cores": cores, "kind": "port",
            "sample": "oracle/cosmo_oracle.c accel+integrate+collision predicate on target rows [0,%d) of the "
                      "N=%d scene (%d x %d ordered pairs, %.1f s per step, OpenMP %d threads)"
                      % (rows, n, rows, n, t, cores),
            "ms_per_step_sample": t * 1e3, "rows": rows}


# ------------------------------------------------------------------------------------ B200 arm helpers
def state_digest(eng):
    """(discrete, full) sha1 digests of the device state.  discrete = ids, mass / radius / volume bits,
    eaten counts and the whole merge stream: equal across GPU counts when the sharded run reproduced the
    single-GPU one (the driver's 1/2/4/8 sweep can compare them).  full adds position / velocity bits,
    which agree only to rounding across GPU counts (the cross-rank sum is another summation order)
    but must be identical on all replicas of one run."""
    import hashlib
    d = eng.download()
    ev = eng.events()
    h = hashlib.sha1()
    for k in ("id", "mass", "radius", "volume", "eaten"):
        h.update(np.ascontiguousarray(d[k]).tobytes())
    h.update(np.ascontiguousarray(ev).tobytes())      # the whole merge stream (frame, absorber, absorbed)
    h.update(np.array([d["status"]["n_events"], d["status"]["frame"], d["status"]["n_escaped"]], dtype=np.int64).tobytes())
    discrete = h.hexdigest()[:16]
    for k in ("pos", "vel"):
        h.update(np.ascontiguousarray(d[k]).tobytes())
    return discrete, h.hexdigest()[:16]


def timed_frames(eng, K, kw, flush, barrier, world, restore=None):
    """K frames bracketed by barrier + synchronize, device-timed (max over ranks), with per-stage
    events.  restore: a save_state() copy put back before EVERY frame (outside the per-frame events
    but inside nothing else: its frames all start from the same full-size state)."""
    import torch
    import torch.distributed as dist
    stage_events = []
    status_log = torch.zeros(K + 1, 16, dtype=torch.int64).pin_memory()   # async snapshots, no sync
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.lib.cosmo_launch_count()
    barrier()
    e0.record()
    for k in range(K):
        if restore is not None:
            eng.restore_state(restore)
        status_log[k].copy_(eng.t["status"], non_blocking=True)
        eng.step(1, DT, G, R_MAX, stage_events=stage_events, **kw)
        flush.zero_()
    e1.record()
    launches = int(eng.lib.cosmo_launch_count() - launches0)     # counted inside the library
    status_log[K].copy_(eng.t["status"], non_blocking=True)
    barrier()
    if restore is None:
        ms = e0.elapsed_time(e1)
    else:      # frames only (first to last stage event of each frame): the state restore is not part of a step
        ms = float(sum(ev[0].elapsed_time(ev[4]) for ev in stage_events))
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=eng.device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    n_per_step = status_log[:K, 0].numpy().astype(np.float64)
    return ms, n_per_step, stage_events, launches


def stage_table(stage_events, world):
    f = lambda a, b: float(np.mean([ev[a].elapsed_time(ev[b]) for ev in stage_events]))  # noqa: E731
    tab = {"prep": f(0, 1), "accel_integrate" + ("+exchange" if world > 1 else ""): f(1, 2),
           "detect" + ("+exchange" if world > 1 else ""): f(2, 3), "resolve_commit": f(3, 4)}
    tab["sum"] = float(sum(tab.values()))
    return tab


def roofline_block(eng, prec, kw, peaks, stage_events, n_per_step, world, workload, ms_per_step):
    from cosmosim_b200 import _abi
    accel_each = np.array([ev[1].elapsed_time(ev[2]) for ev in stage_events])
    accel_ms = float(np.mean(accel_each))
    n_eff = math.sqrt(float(np.mean(n_per_step ** 2)))
    # interactions/s of stage A on THIS GPU (rank 0 evaluates 1/world of the pairs; with world > 1
    # the stage time includes the all-reduce / all-gather that follows the kernel)
    kernel_rate = float(np.sum(n_per_step ** 2) / world / np.sum(accel_each * 1e-3))
    opts_now = eng.params(DT, G, R_MAX, **kw).opts
    sym = bool(opts_now & (_abi.OPT_F32_SYM | _abi.OPT_F64_SYM))
    # the symmetric kernels evaluate each unordered pair once: 12 FMA-pipe (FP32) / 23 FP64-pipe
    # instructions per pair = 6 / 11.5 per ordered interaction; the
    # one-sided kernels need 9 (FP32) / 20 (FP64)
    ipi = ({"fp32": 6.0, "fp64": 11.5}[prec]) if sym else INSTR_PER_INTERACTION[prec]
    achieved = kernel_rate * ipi * 2 / 1e12                            # FMA-equivalent TFLOP/s
    peak = peaks.get("ffma_tflops" if prec == "fp32" else "dfma_tflops")
    # algorithmic HBM bytes of stage A per launch: sources read once (L2-resident afterwards)
    # + targets' state read + new state written
    bytes_per_body = (12 + 8 + 16 + 16 + 40) if prec == "fp32" else (32 + 16 + 16 + 40)
    hbm_gbs = n_eff * bytes_per_body / (accel_ms * 1e-3) / 1e9
    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = mp.get("hbm_gbs", 6650.0)
    kname = "k_accel_%s%s" % ("f32" if prec == "fp32" else "f64", "_sym" if sym else "")
    traffic, traffic_note = None, None
    try:   # DRAM bytes per launch from the committed ncu capture of this kernel on this workload
        ent = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("%s|%s" % (kname, workload), {})
        if ent.get("bytes"):
            # the capture was taken at n_capture planets; the kernel's DRAM traffic is dominated by the
            # column-partial buffer, which scales like n^2 -> scale to this run's RMS active count
            nc = float(ent.get("n_capture") or n_eff)
            traffic = float(ent["bytes"]) * (n_eff / nc) ** 2 / max(1, world)
            traffic_note = "ncu capture at n=%d (%s) scaled by (n_eff/n_capture)^2%s" % (
                nc, ent.get("source", "profiles/"), " / ranks" if world > 1 else "")
    except Exception:
        pass
    tab = stage_table(stage_events, world)
    return {"bound": "fma_pipe_" + prec, "kernel": kname,
            "frac_vs_one_sided_convention": (kernel_rate * INSTR_PER_INTERACTION[prec] * 2 / 1e12 / peak) if peak else None,
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": (achieved / peak) if peak else None, "traffic": traffic, "traffic_note": traffic_note,
            "algorithmic_bytes": n_eff * bytes_per_body,
            "peak_source": "in-process %s micro-benchmark (cosmo_microbench) on this GPU; "
                           "MEASURED_PEAKS.json holds no CUDA-core peak" % ("FFMA / FFMA2 (the better of the two)" if prec == "fp32" else "DFMA"),
            "instruction_mix_ceiling": ({"tflops": peaks.get("pair_chain_tflops"),
                                         "frac_of_peak": peaks["pair_chain_tflops"] / peak,
                                         "kernel_frac_of_ceiling": achieved / peaks["pair_chain_tflops"],
                                         "what": "the kernel's pair chain (2 FADD2 + 6 FFMA2 + 4 FMUL2 + 2 MUFU per two pairs) "
                                                 "on registers only, no shared memory / shuffles / barriers "
                                                 "(cosmo_microbench kind 11): operand bandwidth of the register file"}
                                        if prec == "fp32" and peak and peaks.get("pair_chain_tflops") else None),
            "instr_per_interaction": ipi, "kernel_ms": accel_ms, "n_eff": n_eff,
            "interactions_per_s_kernel": kernel_rate, "stage_ms": tab,
            "outside_stage_timers_ms": ms_per_step - tab["sum"],
            "exchange_ms": ({"after_stage_a": float(np.mean([ev[5].elapsed_time(ev[2]) for ev in stage_events])),
                             "after_stage_b": float(np.mean([ev[6].elapsed_time(ev[3]) for ev in stage_events]))}
                            if world > 1 else None),
            "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak, "frac": hbm_gbs / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in mp else "fallback"},
            "mufu_rsq_per_s_peak": peaks.get("mufu_rsq_per_s")}


def secondary_block(workload, device, peaks, flush, barrier, group=None, world=1, steps=5, warmup=3):
    """Another BASELINE configuration measured in the same run, so that it has a driver-run record of
    its own (value, ms per step, roofline, stage table): config 3 (star+disk, N = 65,536, FP64 parity
    path) on one GPU, config 4 (rotating cloud, N = 262,144, FP32, rows sharded) on N > 1 GPUs."""
    import torch
    import torch.distributed as dist
    from cosmosim_b200.engine import Engine
    kind, n, prec, collisions = WORKLOADS[workload]
    arrays = build_arrays(kind, n)
    eng = Engine(n, device=device, group=group)
    eng.upload(arrays)
    kw = dict(collisions=collisions, bounded=True, fp32=(prec == "fp32"))
    for _ in range(warmup):
        eng.step(1, DT, G, R_MAX, **kw)
        flush.zero_()
    barrier()
    ms, n_per_step, stage_events, launches = timed_frames(eng, steps, kw, flush, barrier, world)
    st = eng.check_errors()
    roof = roofline_block(eng, prec, kw, peaks, stage_events, n_per_step, world, workload, ms / steps)
    return {"workload": workload, "n_bodies": n, "precision": prec, "n_gpus": world, "steps": steps, "warmup": warmup,
            "value": float(np.sum(n_per_step ** 2)) / (ms * 1e-3), "unit": "interactions/s",
            "ms_per_step": ms / steps, "steps_per_s": steps / (ms * 1e-3), "gpu_launches": launches,
            "n_active_per_step": [int(x) for x in n_per_step], "merges": st["n_events"], "roofline": roof,
            "parity": ("FP64 path: accelerations / positions <= 1e-12 vs the reference per step, merges bit-exact"
                       if prec == "fp64" else "FP32 path + FP64 near field: median <= 1e-5, max <= 1e-3 per step")
                      + " (tests/test_gpu_parity.py)"}


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="star_disk_1m_fp32", choices=sorted(WORKLOADS))
    ap.add_argument("--bodies", dest="n", type=int, default=0, help="override the body count (development)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the full-N and FP64 secondary blocks")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    a = ap.parse_args()

    kind, n, prec, collisions = WORKLOADS[a.workload]
    if a.n:
        n = a.n
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(3, a.warmup) if a.impl == "b200" else max(0, a.warmup)
    K = max(1, a.steps)
    config = {"workload": a.workload, "scene": kind, "n_bodies": n, "precision": prec, "collisions": collisions,
              "dt": DT, "parallelism": "rows%d" % world if world > 1 else "single",
              "l2": "160 MiB memset (> the 126 MB L2) between steps, inside the timed region"}

    if kind == "disk3d":
        return main_state3d(a, n, prec, collisions, rank, world, local_rank, W, K, config)

    # ---------------------------------------------------------------- reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return 0
        arrays = build_arrays(kind, n)
        leg = cpu_oracle_leg(arrays, n, seconds_target=min(a.cpu_seconds, 100.0 / (K + W)), steps=K, warmup=W)
        line = {"impl": "reference", "metric": "pairwise_interactions_per_s", "value": leg["value"],
                "unit": "interactions/s", "n_gpus": 0, "steps": K, "warmup": W,
                "ms_per_step": leg["ms_per_step_sample"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": leg["value"], "unit": "interactions/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "steps_per_s_extrapolated": leg["value"] / (float(n) * n)}
        print(json.dumps(line), flush=True)
        return 0

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    from cosmosim_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD
    arrays = build_arrays(kind, n)
    eng = Engine(n, device="cuda:%d" % local_rank, group=group)
    eng.upload(arrays)
    kw = dict(collisions=collisions, bounded=True, fp32=(prec == "fp32"))
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=eng.device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()              # started early: nvidia-smi needs ~0.1 s before its first sample
    # pipe peaks for the roofline denominator (tiny, before the timed region)
    peaks = measure_pipe_peaks(eng, prec) if rank == 0 else {}
    full_state = eng.save_state() if not a.no_extras else None     # the scene at its full size, on the device
    for _ in range(W):
        eng.step(1, DT, G, R_MAX, **kw)
        flush.zero_()
    barrier()
    warm_samples = len(sampler.lines)
    ms, n_per_step, stage_events, gpu_launches = timed_frames(eng, K, kw, flush, barrier, world)
    clocks = None
    if rank == 0:
        timed_samples = len(sampler.lines) - warm_samples
        if timed_samples >= 3:       # enough samples inside the timed region: use only those
            sampler.lines = sampler.lines[warm_samples:]
        clocks = sampler.stop()
        clocks["window"] = "timed region" if timed_samples >= 3 else "pipe micro-benchmark + warm-up + timed region (timed region shorter than 3 samples of 20 ms)"
    st = eng.check_errors()
    # interactions actually evaluated: N_active^2 per frame, and N_active shrinks as planets merge
    interactions = float(np.sum(n_per_step ** 2))
    n_eff = math.sqrt(interactions / K)
    value = interactions / (ms * 1e-3)
    config["n_eff"] = n_eff          # RMS active count over the timed frames (the scene merges ~7 % of its planets per frame)
    roofline = roofline_block(eng, prec, kw, peaks, stage_events, n_per_step, world, a.workload, ms / K)
    digest, digest_full = state_digest(eng)
    replicas_identical = None
    if world > 1:
        got = [None] * world
        dist.all_gather_object(got, digest_full)
        replicas_identical = len(set(got)) == 1

    # ---------------------------------------------------------------- full-N frames (device-timed)
    full_n = None
    if full_state is not None:
        Kf = min(K, 5)
        eng.restore_state(full_state)
        eng.step(1, DT, G, R_MAX, **kw)          # warm this size's launch shapes
        fms, fn, fev, _ = timed_frames(eng, Kf, kw, flush, barrier, world, restore=full_state)
        eng.check_errors()
        full_n = {"n_bodies": int(fn[0]), "steps": Kf, "ms_per_step": fms / Kf, "steps_per_s": Kf / (fms * 1e-3),
                  "value": float(np.sum(fn ** 2)) / (fms * 1e-3), "unit": "interactions/s",
                  "stage_ms": stage_table(fev, world),
                  "what": "every frame starts from the scene's initial state (N = %d, restored device-to-device "
                          "outside the frame's events): ms per TRUE full-size frame, device-timed" % int(fn[0])}

    # ---------------------------------------------------------------- sharded == single, asserted in this run
    sharded_vs_single = None
    if world > 1 and full_state is not None:
        # one full-size frame from the scene's initial state on the sharded engine and on a single-GPU
        # engine of this rank: merge stream, ids, masses, radii must be identical, positions agree to
        # rounding (the cross-rank sum is another summation order for a handful of rows per frame --
        # which is also why the discrete digests of LONG runs may part ways: the scene is chaotic)
        single = Engine(n, device="cuda:%d" % local_rank)
        single.upload(arrays)
        ev_single0 = single.status()["n_events"]
        single.step(1, DT, G, R_MAX, **kw)
        eng.restore_state(full_state)
        ev_shard0 = eng.status()["n_events"]
        eng.step(1, DT, G, R_MAX, **kw)
        d1, d2 = single.download(), eng.download()
        e1, e2 = single.events(ev_single0), eng.events(ev_shard0)
        same = (d1["id"].shape == d2["id"].shape and all(np.array_equal(d1[k], d2[k]) for k in ("id", "mass", "radius", "eaten"))
                and np.array_equal(e1[:, 1:], e2[:, 1:]))
        rel = float(np.abs(d1["pos"] - d2["pos"]).max() / np.abs(d1["pos"]).max()) if same else None
        # 1e-9: a row whose FP32 sum carries a huge near-pair term that the FP64 near-field pass cancels
        # again turns a 1-ulp difference of that sum into ~1e-11 of its displacement
        ok = torch.tensor([1 if (same and rel is not None and rel <= 1e-9) else 0], device=eng.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        sharded_vs_single = {"frames": 1, "n_bodies": n, "merges": int(len(e1)), "discrete_state_and_merge_stream_identical": bool(same),
                             "max_rel_pos_diff": rel, "ok_on_all_ranks": bool(ok.item())}
        del single
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- e2e (host buffers)
    e2e = None
    if not a.no_e2e:
        eng.upload(arrays)
        stg = eng.make_host_staging(arrays)
        for _ in range(2):
            eng.frame_from_host(stg, DT, G, R_MAX, **kw)
        barrier()
        t0 = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(K):
            eng.frame_from_host(stg, DT, G, R_MAX, **kw)
        f1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e_ms = max(f0.elapsed_time(f1), 0.0)
        if world > 1:
            tms = torch.tensor([e_ms], dtype=torch.float64, device=eng.device)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            e_ms = float(tms.item())
        # every e2e step re-uploads the same n-planet host state, so each evaluates n^2 pairs
        e2e = {"value": float(n) * n * K / (e_ms * 1e-3), "unit": "interactions/s",
               "h2d_bytes_per_step": int(stg["h2d_bytes"]), "d2h_bytes_per_step": int(stg["d2h_bytes"]),
               "ms_per_step": e_ms / K, "wall_ms_per_step": wall_ms / K, "n_bodies_per_step": n,
               "path": stg.get("path", "Engine.frame_from_host: pinned host SoA -> H2D -> frame -> D2H, stream sync per step")}

    secondary = None
    exchange = (("nvlink peer stores + symmetric-memory barrier" if getattr(eng, "_peer", None)
                 else "nccl all-reduce of the partial sums + all-gather of the pair lists") if world > 1 else None)
    if not a.no_extras and a.workload == "star_disk_1m_fp32":
        del eng
        torch.cuda.empty_cache()
        if world == 1:      # BASELINE config 3
            secondary = secondary_block("star_disk_64k_fp64", "cuda:%d" % local_rank, peaks, flush, barrier)
        else:               # BASELINE config 4: N = 262,144 FP32, rows sharded over the ranks
            secondary = secondary_block("cloud_256k_fp32", "cuda:%d" % local_rank, peaks, flush, barrier, group, world)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_oracle_leg(arrays, n, seconds_target=a.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "pairwise_interactions_per_s", "value": value, "unit": "interactions/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms / K, "steps_per_s": K / (ms * 1e-3),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if prec == "fp32" else "f64", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches,
                "roofline": roofline, "cpu_baseline": cpu, "full_n": full_n, "secondary": secondary,
                "final_state": {"n_active": st["n_active"], "merges": st["n_events"], "escaped": st["n_escaped"]},
                "state_digest": digest, "state_digest_full": digest_full, "replicas_identical": replicas_identical,
                "sharded_vs_single": sharded_vs_single,
                "exchange": exchange,
                "n_active_per_step": [int(x) for x in n_per_step],
                "pairs_last_frame": st["n_pairs"], "detect_mode": "grid (tuned on the device)" if st["n_active"] >= 8192 else "all-pairs",
                "fp32_bound": ("per-step relative acceleration error vs the FP64 reference formula: see "
                               "tests/test_gpu_parity.py::test_n1m_fp32_full_size_properties") if prec == "fp32" else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------ 3-D step
def cpu_oracle3_leg(arrays, seconds_target=12.0, steps=1, warmup=0):
    """CPU oracle of State.interact (oracle/state3d_oracle.c: accel + integrate + collision predicate)
    on a bounded sample of target rows of the same scene, all host threads."""
    from oracle import oracle as orc
    from oracle import oracle3d as o3
    n = arrays["mass_f64"].shape[0]
    pos, vel, mass, rad = arrays["pos"], arrays["vel"], arrays["mass_f64"], arrays["radius"]
    orc.set_threads(host_threads())
    cores = orc.num_threads()

    def sample(rows):
        t0 = time.perf_counter()
        acc = o3.accel(pos, mass, G, 0, rows, compensated=False)
        p, v = o3.integrate(pos[:rows], vel[:rows], acc, DT3)
        pn = pos.copy()
        pn[:rows] = p
        o3.flags(pn, rad, 0, rows)
        return time.perf_counter() - t0

    rows = min(n, max(cores, 8))
    sample(rows)
    for _ in range(4):
        t = sample(rows)
        want = int(min(n, max(rows, seconds_target * rows / max(t, 1e-9))))
        if want <= rows * 1.25 or t >= 0.5 * seconds_target:
            break
        rows = min(want, rows * 16)
    for _ in range(warmup):
        sample(rows)
    t = float(np.mean([sample(rows) for _ in range(max(1, steps))]))
    return {"value": rows * n / t, "unit": "interactions/s", "cores": cores, "kind": "port",
            "sample": "oracle/state3d_oracle.c accel+integrate+collision predicate on target rows [0,%d) of the "
                      "N=%d 3-D scene (%d x %d ordered pairs, %.1f s per step, OpenMP %d threads)"
                      % (rows, n, rows, n, t, cores), "ms_per_step_sample": t * 1e3, "rows": rows}


def main_state3d(a, n, prec, collisions, rank, world, local_rank, W, K, config):
    """Bench lines of the 3-D step (one GPU): same JSON contract as the main path."""
    from cosmosim_b200 import scenes
    config = dict(config, dt=DT3, scene="disk3d (star + thick disk, 0.1-2 AU)", api="State.interact (universe.py:94-129)")
    arrays = scenes.disk3d_arrays(n, seed=0)
    if a.impl == "reference":
        if rank != 0:
            return 0
        leg = cpu_oracle3_leg(arrays, seconds_target=min(a.cpu_seconds, 100.0 / (K + W)), steps=K, warmup=W)
        print(json.dumps({"impl": "reference", "metric": "pairwise_interactions_per_s", "value": leg["value"],
                          "unit": "interactions/s", "n_gpus": 0, "steps": K, "warmup": W,
                          "ms_per_step": leg["ms_per_step_sample"], "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
                          "e2e": {"value": leg["value"], "unit": "interactions/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}), flush=True)
        return 0
    import torch
    import torch.distributed as dist
    from cosmosim_b200.engine3d import Engine3D
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = Engine3D(n, device="cuda:%d" % local_rank, group=group)
    eng.sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    eng.upload(arrays)
    kw = dict(collisions=collisions, fp32=(prec == "fp32"))
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=eng.device)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    peaks = measure_pipe_peaks(eng, prec) if rank == 0 else {}
    for _ in range(W):
        eng.step(DT3, G, **kw)
        flush.zero_()
    barrier()
    warm_samples = len(sampler.lines)
    stage_events = []
    status_log = torch.zeros(K + 1, 16, dtype=torch.int64).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.lib.cosmo_launch_count()
    e0.record()
    for k in range(K):
        status_log[k].copy_(eng.t["status"], non_blocking=True)
        eng.step(DT3, G, stage_events=stage_events, **kw)
        flush.zero_()
    e1.record()
    gpu_launches = int(eng.lib.cosmo_launch_count() - launches0)
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=eng.device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    clocks = None
    if rank == 0:
        timed_samples = len(sampler.lines) - warm_samples
        if timed_samples >= 3:
            sampler.lines = sampler.lines[warm_samples:]
        clocks = sampler.stop()
        clocks["window"] = "timed region" if timed_samples >= 3 else "pipe micro-benchmark + warm-up + timed region"
    st = eng.check_errors()
    n_per_step = status_log[:K, 0].numpy().astype(np.float64)
    interactions = float(np.sum(n_per_step ** 2))
    value = interactions / (ms * 1e-3)
    accel_each = np.array([ev[0].elapsed_time(ev[1]) for ev in stage_events])
    accel_ms = float(np.mean(accel_each))
    kernel_rate = float(interactions / world / np.sum(accel_each * 1e-3))   # this GPU's share of the pairs
    from cosmosim_b200 import _abi
    sym = bool(eng.params(DT3, G, **kw).opts & (_abi.OPT3_F64_SYM | _abi.OPT3_F32_SYM))
    # the symmetric kernels evaluate each unordered pair once: 29 FP64-pipe / 16 FMA-pipe instructions per pair
    ipi = ({"fp32": 8.0, "fp64": 14.5}[prec]) if sym else INSTR3_PER_INTERACTION[prec]
    achieved = kernel_rate * ipi * 2 / 1e12
    peak = peaks.get("ffma_tflops" if prec == "fp32" else "dfma_tflops")
    kname = "k3_accel_%s%s" % ("f32" if prec == "fp32" else "f64", "_sym" if sym else "")
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
            "%s|%s" % (kname, a.workload), {}).get("bytes")
    except Exception:
        pass
    roofline = {"bound": "fma_pipe_" + prec, "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if peak else None, "traffic": traffic,
                "frac_vs_one_sided_convention": (kernel_rate * INSTR3_PER_INTERACTION[prec] * 2 / 1e12 / peak)
                if peak else None,
                "peak_source": "in-process %s micro-benchmark (cosmo_microbench) on this GPU"
                               % ("FFMA" if prec == "fp32" else "DFMA"),
                "instr_per_interaction": ipi, "kernel_ms": accel_ms, "interactions_per_s_kernel": kernel_rate,
                "stage_ms": {"prep+accel_integrate" + ("+exchange" if world > 1 else ""): accel_ms,
                             "detect": float(np.mean([ev[1].elapsed_time(ev[2]) for ev in stage_events])),
                             "resolve_commit": float(np.mean([ev[2].elapsed_time(ev[3]) for ev in stage_events]))}}
    e2e = None
    if not a.no_e2e:
        eng.upload(arrays)
        stg = eng.make_host_staging(arrays)
        for _ in range(2):
            eng.step_from_host(stg, DT3, G, **kw)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        f0.record()
        for _ in range(K):
            eng.step_from_host(stg, DT3, G, **kw)
        f1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e_ms = f0.elapsed_time(f1)
        if world > 1:
            tms = torch.tensor([e_ms], dtype=torch.float64, device=eng.device)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            e_ms = float(tms.item())
        e2e = {"value": float(n) * n * K / (e_ms * 1e-3), "unit": "interactions/s",
               "h2d_bytes_per_step": int(stg["h2d_bytes"]), "d2h_bytes_per_step": int(stg["d2h_bytes"]),
               "ms_per_step": e_ms / K, "wall_ms_per_step": wall_ms / K,
               "path": "Engine3D.step_from_host: pinned host SoA -> H2D -> cosmo3_step -> D2H, stream sync per step"}
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_oracle3_leg(arrays, seconds_target=a.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return 0
    print(json.dumps({"metric": "pairwise_interactions_per_s", "value": value, "unit": "interactions/s", "n_gpus": world,
                      "steps": K, "warmup": W, "ms_per_step": ms / K, "steps_per_s": K / (ms * 1e-3),
                      "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "f32" if prec == "fp32" else "f64", "data": "synthetic", "config": config,
                      "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline,
                      "cpu_baseline": cpu, "final_state": {"n_active": st["n_active"], "absorb_calls": st["n_events"]},
                      "n_active_per_step": [int(x) for x in n_per_step], "pairs_last_step": st["n_pairs"],
                      "jsplit": eng.params(DT3, G, **kw).jsplit}), flush=True)
    return 0


def measure_pipe_peaks(eng, prec):
    """FFMA / DFMA / MUFU peaks (thread-instructions per second) from the in-library
    micro-benchmark: the measured denominators SURVEY.md section 8d asks for."""
    import torch
    from cosmosim_b200 import _abi
    L = eng.lib
    out = torch.zeros(eng.sm_count * 2 * 1024, dtype=torch.float64, device=eng.device)
    stream = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)

    def rate(kind, iters, shapes=((512, 2), (1024, 1), (1024, 2))):
        best = 0.0
        for threads, bps in shapes:
            ops = C.c_int64()
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _abi.check(L.cosmo_microbench(kind, eng.sm_count * bps, threads, iters, C.c_void_p(out.data_ptr()),
                                              C.byref(ops), stream), "cosmo_microbench")
                e1.record()
                torch.cuda.synchronize()
                best = max(best, ops.value * eng.sm_count * bps * threads / (e0.elapsed_time(e1) * 1e-3))
        return best

    ffma, dfma, mufu = rate(0, 8000), rate(2, 3000), rate(3, 3000)
    # packed FP32: FFMA2 with operands from the reuse cache (the pipe's peak: two FMAs per lane and
    # instruction), and the pair chain of k_accel_f32_sym on registers only (kind 11: the ceiling of the
    # kernel's instruction mix -- an FFMA2 with three live register-pair operands takes 3 cycles, the
    # register file delivers two pairs per two cycles; profiles/r02_kernel_experiments.md)
    ffma2, chain = rate(1, 4000), rate(11, 2000, ((256, 2), (256, 4)))   # kinds 9-12: CTAs of 256
    return {"ffma_per_s": ffma, "dfma_per_s": dfma, "mufu_rsq_per_s": mufu, "ffma2_per_s": ffma2,
            "ffma_tflops": max(2 * ffma, 4 * ffma2) / 1e12, "ffma_scalar_tflops": 2 * ffma / 1e12,
            "ffma2_tflops": 4 * ffma2 / 1e12, "pair_chain_tflops": 4 * chain / 1e12, "dfma_tflops": 2 * dfma / 1e12}


if __name__ == "__main__":
    sys.exit(main())
