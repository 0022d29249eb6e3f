#!/usr/bin/env python
"""bench.py — particle-substeps/s of the MPM substep (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--scene jelly_collision]

One "step" = one MPM substep (bin -> P2G -> grid -> G2P/advance) over the whole synthetic scene.
At N=1 the workload is BASELINE.json configs[1]: the 1.02 M-particle two-block Neo-Hookean jelly
collision (squishy_volumes_b200/scenes.py:jelly_collision, side=80).  For N>1 the blocks grow along x with the
GPU count (weak scaling, 1.02 M particles per GPU) and the domain is slab-decomposed: every substep the
neighbours exchange grid-halo sums after P2G and migrating particles after the advance, written straight
into the neighbour's HBM over NVLink (CUDA-IPC mailboxes; NCCL send/recv is the fallback), DESIGN.md §8.

Printed JSON keys follow the driver contract; in particular
  value     device-timed throughput, state already resident in HBM (CUDA events on the library's
            own stream around the substep loop, `svb_last_advance_ms`), max over ranks;
  e2e       same metric through the public API with HOST (page-locked) buffers: upload of the IoState
            (H2D) + produce_next_state (K substeps + D2H of the IoState) inside the timed region; the
            handle (allocations, communicator) is created before the clock starts;
  roofline  dominant kernel: algorithmic bytes per launch / its mean launch time (CUDA events per
            stage, a separate instrumented pass) against MEASURED_PEAKS.json's HBM copy bandwidth;
  cpu_baseline   the oracle (C++ restatement of the reference's CPU path, OpenMP) on a bounded
            sample of the same scene on this box's host cores.

`--impl reference` times that CPU restatement on the same scene with all host threads (the
reference's Rust crate cannot be built here: no cargo/rustc, un-vendored deps).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from squishy_volumes_b200 import abi as _abi, scenes  # noqa: E402
from squishy_volumes_b200.types import RunParameters  # noqa: E402

METRIC = "particle-substeps/sec"
UNIT = "particle-substeps/s"
# algorithmic (compulsory) bytes per particle-substep
A_P2G = 120.0
A_G2P = 56.0 + 96.0
A_GRID = 8.0
A_NOSORT = 280.0
A_SORT_EXTRA = 272.0
E2E_REPEATS = 5


def measured_traffic(scene_name: str, particles_per_gpu: int, kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from a committed `ncu --set full` capture of exactly
    this workload (profiles/traffic.json, filled by profiles/summarize.py), or None: a number from another scene or size is
    not this run's traffic."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None
    hit = table.get(f"{scene_name}:{int(particles_per_gpu)}:{kernel}")
    return float(hit["bytes"]) if hit else None


def make_scene(name: str, scale: float, length: int = 1):
    """The five BASELINE.json configs at `scale` times their named particle count.  The collider scenes start IN CONTACT
    (scenes.py: `contact=True`) so that a short timed run exercises collide, meld and the return mapping, not free fall."""
    if name == "jelly_collision":
        side = max(4, int(round(80 * scale ** (1.0 / 3.0))))
        return scenes.jelly_collision(side=side, length=length)
    if name == "elastic_cube":
        return scenes.elastic_cube(side=max(4, int(round(46 * scale ** (1.0 / 3.0)))))
    if name == "sand_torus":
        return scenes.sand_torus(side=max(8, int(round(200 * scale ** (1.0 / 3.0)))), contact=True)
    if name == "dam_break":
        s = scale ** (1.0 / 3.0)
        return scenes.dam_break(nx=max(8, int(400 * s)), ny=max(4, int(200 * s)), nz=max(4, int(200 * s)), contact=True)
    if name == "mixed":
        return scenes.mixed(side=max(16, int(round(400 * scale ** (1.0 / 3.0)))), contact=True)
    raise SystemExit(f"unknown scene {name}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.15)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except Exception:
                continue
            for nme, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_leg(scene, steps: int, warmup: int, budget_s: float):
    """Oracle (reference-algorithm CPU restatement) on this box's host cores; bounded sample."""
    import oracle.oracle as orc
    orc.build()
    L = orc.lib()
    # all the host threads the box offers, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for every rank)
    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    L.svo_set_threads(int(usable))
    cores = int(L.svo_max_threads())
    o = orc.OracleState.from_io_state(scene.io_state, scene.frame_input)
    o._sync_keyframes(scene.frame_input)
    dt = scene.time_step
    k = 0
    for _ in range(warmup):
        k += 1
        L.svo_advance(o._h, (k - 0.5) * dt, dt, 0, None)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        k += 1
        L.svo_advance(o._h, (k - 0.5) * dt, dt, 0, None)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    el = time.perf_counter() - t0
    return {"value": scene.n * done / el, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": 1e3 * el / done, "steps_done": done,
            "sample": f"{done} substeps of the same {scene.n}-particle scene after {warmup} warm-up substeps (C++/OpenMP restatement of the reference CPU path; Rust crate not buildable here)"}


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        # torch.distributed (NCCL) carries the control plane: unique-id hand-over, barriers, max over ranks.  The data
        # path — halo sums and migrating particles — runs on the library's own NCCL communicator (svb_comm_init).
        dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def all_max(dist, local, value: float) -> float:
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(dist, local, value: float) -> float:
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(dist, local):
    import torch
    torch.cuda.synchronize(local)
    if dist is not None:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize(local)


def slab_parity_check(dist, rank, world, local, fresh_unique_id):
    """Outside the timed region: a small fast-approach jelly collision (8 192 particles, 24 substeps, particles cross every
    cut) on the SAME `world` slab ranks against the single-GPU run of the same scene on rank 0.  Integer fields (flags,
    collider bits, cell keys) must agree exactly, floats within the per-substep-compounded tolerance of tests/parity.py.
    Returns the dict printed as "slab_parity" (driver-side evidence that the multi-GPU path computes the same thing)."""
    from squishy_volumes_b200 import slabs
    from squishy_volumes_b200.state import B200State
    sc = scenes.jelly_collision(side=16)
    sc.io_state.particles.velocities[:, 0] *= 6.0
    sc.frame_input.consts.frames_per_second = 1
    params = RunParameters(target_time=23.5 * sc.time_step, max_time_step=sc.time_step)
    st = slabs.SlabState.from_io_state(sc.io_state, sc.frame_input, rank, world, local, fresh_unique_id())
    err = st.advance(None, sc.frame_input, params)
    idx, rows = st.resident()
    gathered = [None] * world
    dist.all_gather_object(gathered, (idx, rows, None if err is None else err.status))
    st.close()
    out = None
    if rank == 0:
        single = B200State.from_io_state(sc.io_state, sc.frame_input, device=local)
        want, err1 = single.produce_next_state(None, sc.frame_input, params)
        single.close()
        got = slabs.assemble(sc.n, [(i, r) for i, r, _ in gathered], sc.io_state.particles)
        h = sc.frame_input.consts.scaled_grid_node_size()
        cell = lambda x: np.floor(x / np.float32(h) - np.float32(0.5)).astype(np.int32)
        rel = {}
        for f in ("positions", "velocities", "position_gradients", "velocity_gradients"):
            a, b = getattr(got, f).astype(np.float64), getattr(want.particles, f).astype(np.float64)
            rel[f] = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
        start_owner = slabs.slab_of(sc.io_state.particles.positions, h, st.plan)
        end_owner = np.empty(sc.n, np.int32)
        for r, (i, _, _) in enumerate(gathered):
            end_owner[np.asarray(i, np.int64)] = r
        out = {"scene": f"jelly_collision side=16 ({sc.n} particles), 24 substeps, approach speed x6", "ranks": world,
               "integer_fields_exact": bool(np.array_equal(got.flags, want.particles.flags) and np.array_equal(got.collider_bits, want.particles.collider_bits)
                                            and np.array_equal(cell(got.positions), cell(want.particles.positions))),
               "max_rel_err_vs_single_gpu": rel, "tolerance": 2e-3, "within_tolerance": bool(max(rel.values()) <= 2e-3),
               "particles_that_changed_rank": int(np.count_nonzero(start_owner != end_owner)),
               "errors": [e for _, _, e in gathered if e is not None] + ([] if err1 is None else [err1.status])}
    return out


def main():
    # the contract is ONE JSON line on stdout: libraries that chat on fd 1 (NCCL prints its version there) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)    # ~5 output frames at 24 fps, dt = 1e-3: a timed region of tens of milliseconds
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="jelly_collision")
    ap.add_argument("--scale", type=float, default=1.0, help="particle-count multiplier of the named scene")
    ap.add_argument("--strong", action="store_true", help="N > 1: split ONE scene of the named size over the GPUs (default for every scene but jelly_collision)")
    ap.add_argument("--adaptive", action="store_true", help="adaptive time steps (the reference's default mode); max_time_step = 4 x the scene's fixed dt")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (development A/B runs)")
    ap.add_argument("--length", type=int, default=0, help="development: jelly_collision stretched along x like the weak-scaling scene of that many GPUs, whatever N is")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    steps, warmup = max(1, args.steps), max(0, args.warmup)

    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    # N > 1.  jelly_collision (the default workload): weak scaling — the blocks grow along x with the GPU count, 1.02 M particles
    # per GPU.  Every other scene, or --strong: ONE scene of the named size is cut into N slabs (BASELINE configs 4 and 5).
    weak = world_env > 1 and args.scene == "jelly_collision" and not args.strong
    scene = make_scene(args.scene, args.scale, length=args.length or (world_env if weak else 1))
    scene.frame_input.consts.frames_per_second = 1  # one long frame: the bench never crosses a keyframe boundary
    dt = scene.time_step
    max_dt = 4.0 * dt if args.adaptive else dt
    state_mb = scene.n / max(world_env, 1) * 144 / 1e6   # 9 quads of 16 bytes per particle (svb_device.cuh)
    config = {"workload": f"{args.scene}: {scene.description}", "particles_total": scene.n, "particles_per_gpu": scene.n // max(world_env, 1),
              "time_step": dt, "adaptive_time_steps": bool(args.adaptive), "max_time_step": max_dt,
              "rebin": "every substep (counting sort on (tile, cell); the physical permutation rides on the G2P write)",
              "l2": f"per-GPU particle state {state_mb:.0f} MB (144 B/particle: 34 words in 9 quads) + grid vs the 126 MB L2: " + ("larger than L2, no explicit flush" if state_mb > 126 else "NOT larger than L2 - the HBM fractions of this run are partly L2 numbers"),
              "decomposition": "single GPU" if world_env == 1 else f"{world_env} slabs along x ({'weak: the scene grows with N' if weak else 'strong: one scene split N ways'}); halo-column sums + particle migration every substep over peer memory (NVLink, CUDA IPC), NCCL fallback"}
    scaling = "weak" if (weak or world_env == 1) else "strong"

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return 0
        leg = cpu_leg(scene, steps, warmup, budget_s=150.0)
        line = {"impl": "reference", "metric": METRIC, "value": leg["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": leg["steps_done"], "warmup": warmup,
                "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), file=json_out, flush=True)
        return 0

    import torch
    from squishy_volumes_b200.state import B200State
    rank, world, local, dist = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    fi = scene.frame_input
    t0 = scene.io_state.time

    def run_params(state, k):
        # fixed dt: exactly k substeps.  adaptive: the same simulated time span, however many substeps the limits make of it
        return RunParameters(target_time=state.time + ((k - 0.5) * dt if not args.adaptive else k * dt), max_time_step=max_dt, adaptive_time_steps=args.adaptive)

    def fresh_unique_id():
        """An NCCL unique id serves one communicator: every slab state gets its own, handed over by rank 0."""
        from squishy_volumes_b200 import abi
        import ctypes as C
        box = [None]
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            assert abi.load().svb_comm_unique_id(buf) == 0
            box[0] = bytes(buf)
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def make_state(io_state):
        """-> (object with .advance/.time, the B200State that owns the device handle)"""
        if world == 1:
            st = B200State.from_io_state(io_state, fi, device=local)
            return st, st
        from squishy_volumes_b200 import slabs
        st = slabs.SlabState.from_io_state(io_state, fi, rank, world, local, fresh_unique_id())
        return st, st.inner

    slab_parity = slab_parity_check(dist, rank, world, local, fresh_unique_id) if world > 1 else None

    # ---------------- device-resident throughput
    state, inner = make_state(scene.io_state)
    if warmup:
        state.advance(None, fi, run_params(state, warmup))   # contact has begun, buffers have settled
    rebalanced = None
    if world > 1:
        # the cuts were planned on the initial particle positions; material has moved since ("rebalanced by particle
        # count every K substeps" — here once, between the warm-up and the timed region, like a frame loop would between frames)
        try:
            rebalanced = bool(state.rebalance(dist))
        except Exception as exc:   # a refused plan (e.g. a cut that would leave its neighbour slabs) is not an error of the run
            rebalanced = f"not done: {exc}"
    if world == 1:
        inner.snapshot()                                      # device-side copy: the extra passes below repeat exactly the timed one
    launches0 = inner.kernel_launches
    sub0 = state.substeps
    if world > 1:
        inner.exchange_waits(reset=True)
    barrier(dist, local)
    waits = None
    with ClockSampler(local) as clocks:
        state.advance(None, fi, run_params(state, steps))
        ms = inner.last_advance_ms
        done = state.substeps - sub0
        if world > 1:
            waits = {k: (round(v / max(done, 1), 5) if not k.endswith("_tiles") else int(v)) for k, v in inner.exchange_waits().items()}
            waits["resident_rows_at_end"] = int(_abi.load().svb_particle_count(inner._h))
        barrier(dist, local)
        gpu_launches = inner.kernel_launches - launches0
        # the timed pass lasts tens of milliseconds, one nvidia-smi query a good part of a second: keep the same load on the GPU
        # (single GPU: the timed pass itself, restored from the snapshot) until the sampler has seen it a few times
        extra, t_stop = 0, time.perf_counter() + 6.0
        # (slab ranks cannot rewind: their extra passes must leave room for the stage pass inside the one loaded frame, 1 s at fps = 1)
        span = (steps + 1) * dt
        while len(clocks.samples) < 3 and (time.perf_counter() < t_stop if world == 1 else (extra < 10 and state.time + 2.2 * span < 0.95)):
            if world == 1:
                inner.restore()
            state.advance(None, fi, run_params(state, steps))
            extra += 1
    if world == 1:
        inner.restore()                                       # the stage pass below measures the timed pass's state as well
    ms_max = all_max(dist, local, ms)
    total_particles = float(scene.n)
    value = total_particles * done / (ms_max * 1e-3)

    # ---------------- per-stage pass (instrumented: one event pair + sync per stage) -> roofline
    inner.enable_stage_timing(True)
    sub1 = state.substeps
    state.advance(None, fi, run_params(state, steps))
    stage_steps = max(state.substeps - sub1, 1)
    stages = {k: v / stage_steps for k, v in inner.stage_times().items()}
    inner.enable_stage_timing(False)
    stages_by_rank = None
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: round(v, 5) for k, v in stages.items() if v})
        stages_by_rank = gathered
    peak, peak_kind = measured_peak()
    n_local = scene.n / world
    dom = max(("p2g", "g2p"), key=lambda s: stages.get(s, 0.0))
    alg_bytes = n_local * ((A_P2G if dom == "p2g" else A_G2P) + A_GRID / 2)
    dom_ms = max(stages.get(dom, 0.0), 1e-9)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    sub_s = ms_max / done * 1e-3
    whole_nosort = scene.n * A_NOSORT / sub_s / 1e9
    whole_rebin = scene.n * (A_NOSORT + A_SORT_EXTRA) / sub_s / 1e9
    other = "g2p" if dom == "p2g" else "p2g"
    other_bytes = n_local * ((A_P2G if other == "p2g" else A_G2P) + A_GRID / 2)
    other_ms = max(stages.get(other, 0.0), 1e-9)
    roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peak, "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)", "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic(args.scene, scene.n // world, "k_" + dom),
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms,
                "second_kernel": {"kernel": "k_" + other, "achieved": other_bytes / (other_ms * 1e-3) / 1e9, "frac": other_bytes / (other_ms * 1e-3) / 1e9 / peak,
                                  "algorithmic_bytes_per_launch": other_bytes, "launch_ms": other_ms, "traffic": measured_traffic(args.scene, scene.n // world, "k_" + other)},
                "whole_substep": {"algorithmic_bytes_compulsory": scene.n * A_NOSORT, "achieved": whole_nosort, "frac": whole_nosort / (peak * world),
                                  "note": "A_nosort = 280 B per particle-substep (SURVEY.md 8d); the design performs no separate re-sort pass",
                                  "with_rebin_credit": {"algorithmic_bytes": scene.n * (A_NOSORT + A_SORT_EXTRA), "achieved": whole_rebin, "frac": whole_rebin / (peak * world),
                                                        "note": "credits the 272 B a separate physical re-sort every substep would move (SURVEY.md 8d A_sort); shown for comparison only"}},
                "stage_ms_per_substep": stages, "stage_note": "rank 0, instrumented pass (one event pair and a sync per stage)"}
    if stages_by_rank is not None:
        roofline["stage_ms_per_substep_by_rank"] = stages_by_rank   # exchange stages include the wait for the neighbour
        gathered = [None] * world
        dist.all_gather_object(gathered, waits)
        # the timed region itself: time the exchange kernels spent waiting, per substep, by device-side clocks.  The two "send_for_*_boundary"
        # waits run on the second stream beside the interior tiles; the "recv" waits are the ones that stall a rank
        roofline["exchange_wait_ms_per_substep_by_rank"] = gathered
    state.close()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": int(done), "warmup": warmup, "ms_per_step": ms_max / done, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "roofline": roofline, "gpu_launches": int(gpu_launches),
            "clocks": clocks.summary()}
    if slab_parity is not None or world > 1:
        line["slab_parity"] = slab_parity
        line["config"]["rebalanced_after_warmup"] = rebalanced

    # ---------------- end to end through the public API with host buffers (page-locked, as the contract asks)
    if not args.no_e2e:
        line["e2e"] = e2e_leg(args, scene, fi, dist, rank, world, local, make_state, run_params, steps, dt, max_dt, t0, total_particles)
    if rank == 0 and not args.no_cpu:
        leg = cpu_leg(scene, 1000, 2, budget_s=args.cpu_budget)
        line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if dist is not None:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), file=json_out, flush=True)
    return 0


def e2e_leg(args, scene, fi, dist, rank, world, local, make_state, run_params, steps, dt, max_dt, t0, total_particles):
    """Same metric through the public API with HOST buffers: every pass uploads the IoState (H2D from page-locked memory), runs the
    substeps and downloads the IoState (D2H into page-locked memory).  The handle (device allocations, and for N > 1 the NCCL
    communicator + CUDA-IPC mailboxes) is session setup, like the CUDA context: it is created before the clock starts.
    Reported: the MEDIAN of E2E_REPEATS passes (wall clock, max over ranks)."""
    import dataclasses
    import torch
    from squishy_volumes_b200.state import B200State
    from squishy_volumes_b200.types import IoState, Particles
    keep = []

    def pinned(a):
        t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype)), pin_memory=True)
        keep.append(t)
        v = t.numpy()
        v[...] = a
        return v
    fields = ("flags", "mass", "initial_volume", "mu_or_bulk_modulus", "lambda_or_exponent", "sand_alpha", "viscosity_dynamic", "viscosity_bulk", "positions",
              "position_gradients", "velocities", "velocity_gradients", "elastic_energies", "collider_bits")
    target = t0 + ((steps - 0.5) * dt if not args.adaptive else steps * dt)
    params = RunParameters(target_time=target, max_time_step=max_dt, adaptive_time_steps=args.adaptive)
    times = []
    if world == 1:
        host_state = IoState(scene.io_state.time, Particles(**{f.name: pinned(getattr(scene.io_state.particles, f.name)) for f in dataclasses.fields(Particles)}))
        out_buffers = Particles(**{f.name: pinned(getattr(scene.io_state.particles, f.name)) for f in dataclasses.fields(Particles)})
        h2d = sum(getattr(host_state.particles, f).nbytes for f in fields)
        st2 = B200State.from_io_state(host_state, fi, device=local)
        st2.produce_next_state(None, fi, run_params(st2, 2), out=out_buffers)   # first-use costs (kernel loading, keyframe upload, table allocation) are session setup too
        for _ in range(E2E_REPEATS):
            barrier(dist, local)
            te = time.perf_counter()
            st2.upload(host_state)
            out, err = st2.produce_next_state(None, fi, params, out=out_buffers)
            barrier(dist, local)
            times.append(time.perf_counter() - te)
            e2e_done = st2.substeps
        st2.close()
        what = (f"B200State.upload (H2D of the whole state, page-locked) + produce_next_state ({e2e_done} substeps + D2H of the IoState into page-locked arrays), wall clock, "
                f"median of {E2E_REPEATS} passes; handle created beforehand")
    else:
        from squishy_volumes_b200 import slabs
        st2, inner2 = make_state(scene.io_state)       # session: communicator, mailboxes, allocations
        st2.advance(None, fi, run_params(st2, 2))
        st2.resident()                                  # first use of the download kernels is session setup too
        hgrid = fi.consts.scaled_grid_node_size()
        local_rows, idx = slabs.split_state(scene.io_state, hgrid, st2.plan, rank)
        host_state = IoState(scene.io_state.time, Particles(**{f.name: pinned(getattr(local_rows.particles, f.name)) for f in dataclasses.fields(Particles)}))
        room = int(host_state.particles.n * 3 // 2 + 65536)
        big = Particles.empty(room)
        out_buffers = Particles(**{f.name: pinned(getattr(big, f.name)) for f in dataclasses.fields(Particles)})
        del big
        h2d = all_sum(dist, local, float(sum(getattr(host_state.particles, f).nbytes for f in fields)))   # summed over ranks
        for _ in range(E2E_REPEATS):
            barrier(dist, local)
            te = time.perf_counter()
            st2.upload(host_state, idx)                     # H2D of this rank's slab
            st2.advance(None, fi, params)
            idx_out, rows = st2.resident(out=out_buffers)   # D2H of the rows this rank holds now
            barrier(dist, local)
            times.append(all_max(dist, local, time.perf_counter() - te))
            e2e_done = st2.substeps
        st2.close()
        what = (f"per rank: SlabState.upload (H2D of its slab, page-locked) + {e2e_done} substeps with halo exchange and migration + D2H of the resident rows into "
                f"page-locked arrays; wall clock, max over ranks, median of {E2E_REPEATS} passes; communicator / mailboxes / handle created beforehand, host-side split and re-assembly outside")
    e2e_s = float(np.median(times))
    return {"value": total_particles * e2e_done / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / e2e_done, "d2h_bytes_per_step": h2d / e2e_done, "what": what,
            "seconds_per_pass": [round(t, 5) for t in times]}


if __name__ == "__main__":
    sys.exit(main())
