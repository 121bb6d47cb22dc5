#!/usr/bin/env python
"""bench.py -- particle-updates/s of the tiSPHi hot path on B200 (BASELINE.json metric), one JSON line on stdout.

Workload (config.workload): BASELINE config C4, the 3D WCSPH dambreak (Wendland C2, dummy walls, "LF" integrator)
with 10 240 000 fluid + 2 719 788 dummy = 12 959 788 particles on 269 x 136 x 56 cells; one "step" = one
SPHBase.step() = grid build + kernel correction + 2 x one_step + integrator + advect_pos + post-step.
`--workload c3` (soil: the test2 column collapse under Drucker-Prager + CSPM + RK4, refined) and `--workload c5`
(uniform box: grid build + neighbour count + density sum) time the other BASELINE configurations the same way.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA engine (MIXED precision: fp32 sweeps,
                                                                  fp64 positions/densities)
  python bench.py --impl reference ...                           the CPU restatement of the reference (oracle/, float64,
                                                                  OpenMP on all host cores) on the SAME scene, a bounded
                                                                  number of steps (Taichi itself is not installable here)
value        device-timed (CUDA events), state resident in HBM.
e2e          the same metric through the C ABI with HOST buffers: every step uploads the particle state from pinned
             host memory (sph_add_particles), runs sph_step(1) and reads the state back (sph_read_state).
roofline     dominant kernel class, algorithmic bytes / CUDA-event time measured inside the timed region.
cpu_baseline the oracle timed on this box's host cores (rank 0, N = 1 only).
N > 1        the same scene slab-partitioned over N GPUs (strong scaling): one process per GPU, the device-driven slab
             step of csrc/slab.cu (messages stored into the neighbours' inboxes over NVLink peer mappings); the run
             first asserts that the slab run equals the single-GPU run bit for bit (config.slab_parity).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "particle-updates/s", "particle-updates/s"
# The reference's 3D WCSPH scheme diverges about 35 steps after rest at ANY resolution (its 3D Wendland normaliser is
# 8x the textbook value, SURVEY H8: density 1.3 rho0 at step 30, 1e11 rho0 at step 36 in the float64 oracle).
# A run longer than the stable horizon is therefore cut into legs of at most STABLE_STEPS steps, each started from the
# restored initial state (a device-to-device re-upload inside the timed region, < 1 % of a leg).
STABLE_STEPS = 20
STRAIGHT_STEPS = 30       # warm-up + timed steps up to here run straight from rest (rho_max < 1.35 rho0, no crowded cell yet)
CPU_MAX_STEPS = 3         # timed steps of the CPU arm on the full C4 scene (about 10 s each on 32 threads)


def run_in_legs(run_steps, restore, total):
    done = 0
    while done < total:
        restore()
        leg = min(STABLE_STEPS, total - done)
        run_steps(leg)
        done += leg


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ncu_evidence(kernel):
    """DRAM traffic per launch and the unit that bounds the kernel, from the COMMITTED ncu --set full captures
    (profiles/*ncu_evidence.json, written by tools/ncu_evidence.py from the .ncu-rep of this bench command): static
    evidence from an earlier run of the same code, not measured in this run -- the JSON line says so."""
    for name in ("r2_ncu_evidence.json", "r1_ncu_evidence.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                ev = json.load(f).get(kernel)
            if ev:
                return dict(ev, source=f"profiles/{name} (static: captured under ncu by an earlier run of this command)")
        except Exception:
            continue
    return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi style clock / throttle-reason samples during the timed region (pynvml)."""

    def __init__(self, index=0, period=0.1):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.period = period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            log("clock sampling unavailable:", e)
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def c4_workload(n, n_fluid, n_wall, cells, dt):
    return (f"C4 3D WCSPH dambreak (Wendland C2, dummy walls, LF): N={n} ({n_fluid} fluid + {n_wall} wall), "
            f"cells={cells}, dt={dt!r}")


# ---------------------------------------------------------------------------------------------- CPU baseline
def host_threads():
    """Threads the CPU arm uses: every host core, set explicitly (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference(scale, steps, warmup=0):
    """Oracle (float64 restatement of the reference, OpenMP) on the C4 scene (scale 1 = the benchmarked scene itself).
    Like the GPU arm it never takes more than STABLE_STEPS consecutive steps from rest."""
    from oracle import oracle as orc
    from tisphi_b200 import scenes
    L = orc.lib()
    L.orc_set_threads(host_threads())
    threads = int(L.orc_max_threads())
    scene = scenes.dambreak3d(scale=scale, precision="f64")
    t_build = time.perf_counter()
    o = orc.Oracle.from_scene(scene, serial=0)
    t_build = time.perf_counter() - t_build
    n = o.n
    for _ in range(min(warmup, 2)):
        o.step()
    steps = max(1, min(steps, STABLE_STEPS - 2))
    t0 = time.perf_counter()
    for _ in range(steps):
        assert o.step() == 0
    dt = time.perf_counter() - t0
    del o
    same = "the benchmarked C4 scene itself" if scale == 1.0 else f"C4 scene coarsened x{1 / scale:g}"
    return {"value": n * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{same} (N={n}), {steps} steps after {min(warmup, 2)} warm-up, float64, OpenMP omp_get_max_threads()={threads}, "
                      f"{dt:.1f} s (+ {t_build:.1f} s scene build); restated CPU baseline (Taichi not installable in this image)"}, n, dt, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return          # under torchrun only rank 0 times the CPU reference arm
    steps = max(1, min(args.steps, CPU_MAX_STEPS))
    base, n, dt, steps = cpu_reference(args.cpu_scale, steps, warmup=min(args.warmup, 1))
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": C4_WORKLOAD_FULL if args.cpu_scale == 1.0 else "C4 3D WCSPH dambreak, bounded sample: " + base["sample"],
                       "note": f"CPU arm: {steps} timed steps of the same scene (each step costs seconds on the host)"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


C4_WORKLOAD_FULL = c4_workload(12959788, 10240000, 2719788, 2048704, 2.4999999999999998e-05)


# ---------------------------------------------------------------------------------------------- algorithmic bytes
def algorithmic_bytes(kernel, n, n_fluid, n_wall, cells):
    """Algorithmic bytes per launch of a kernel class (DESIGN.md section 'Kernels and their roofline')."""
    table = {
        # read {x,y,z,V | v~x,v~y,v~z,rho~ | type | p} = 36 B of every particle, write {d_rho, d_vel} = 16 B per fluid
        "tile_fluid": 36 * n + 16 * n_fluid,
        # wall pass: read 36 B + write 16 B per wall particle (the gathered kernel only visits wall cells in reach of flow)
        "tile_wall": 52 * n_wall,
        # masks: read {x,y,z,flow} = 16 B per particle; the words themselves are internal traffic
        "tile_mask": 16 * n,
        # reorder (+ init_real2tmp, + SoA / AoS tile payloads): read perm, key, x, m_V, v, rho, p, type, id0 = 84 B,
        # write x, xs, ps4, SoA, v, v~, rho, rho~, p, type, id0, key = 136 B
        "reorder": 220 * n,
        "cell_id": 24 * n + 8 * n, "rank": 12 * n, "scatter_index": 12 * n, "scan": 12 * cells,
        # EOS + tile payloads + dry walls (2 launches per LF step, the second carries advect_LF_half): average per launch
        "wc_eos": 119 * n_fluid + 40 * n_wall,
        # advect_LF + advect_pos + advect_something in one kernel: read 88 B, write 64 B per real particle
        "advect": 152 * n_fluid,
        "init_real2tmp": 52 * n_fluid, "advect_pos": 64 * n_fluid, "post": 40 * n_fluid,
    }
    # (cspm_f, wc_wall, wc_fluid on the cell-tile path are flagged-cells-only launches that return at once when no cell is
    # flagged: dividing the full algorithmic bytes by their duration would be meaningless, so they have no entry)
    return table.get(kernel)


# ---------------------------------------------------------------------------------------------- our arm
def scene_options(args):
    """Engine options (not part of the reference's scene schema) the bench can toggle for A/B runs."""
    opt = {}
    if args.lists is not None:
        opt["neighbourLists"] = bool(args.lists)
    return opt


def run_ours(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        return run_ours_multi(args, rank, local, world)
    torch.cuda.set_device(local)
    if args.workload == "c5":
        from bench_workloads import c5 as bench_c5
        return bench_c5.run(args, local, log, ClockSampler, measured_peaks)
    if args.workload == "c3":
        from bench_workloads import soil as bench_soil
        return bench_soil.run(args, local, log, ClockSampler, measured_peaks, host_threads)
    from tisphi_b200 import scenes
    from tisphi_b200.eng.simulation import Simulation, SimConfiger

    scene = scenes.dambreak3d(scale=args.scale, precision=args.precision, **scene_options(args))
    t0 = time.time()
    sim = Simulation(SimConfiger(config=scene), device=f"cuda:{local}")
    ps, solver, eng = sim.ps, sim.solver, sim.ps.engine
    n = ps.particle_num[None]
    typ = ps.pt.mat_type
    n_fluid = int((typ == 1).sum())
    n_wall = n - n_fluid
    log(f"scene built: N={n} (fluid {n_fluid}, wall {n_wall}), cells={ps.grid_num_total}, dt={solver.dt[None]!r}, {time.time() - t0:.1f}s")

    # host copy of the initial state for the end-to-end leg (pinned)
    pin = lambda t: t.detach().cpu().contiguous().pin_memory()
    h_x, h_rho, h_typ = pin(ps.pt.x), pin(ps.pt.density), pin(ps.pt.mat_type)
    h_v = pin(ps.pt.v.double())
    # device-resident copy of the initial state (legs longer than the stable horizon, the stirred leg)
    d_x, d_rho, d_typ = ps.pt.x.clone().contiguous(), ps.pt.density.clone().contiguous(), ps.pt.mat_type.clone().contiguous()
    d_v = ps.pt.v.double().contiguous()

    def restore(x=None):
        eng.call("sph_clear_particles")
        eng.call("sph_add_particles", n, (d_x if x is None else x).data_ptr(), d_v.data_ptr(), d_rho.data_ptr(), d_typ.data_ptr())

    replay = args.warmup + args.steps > STRAIGHT_STEPS                # the default 3 + 20 runs straight through
    solver.run_steps(min(args.warmup, STABLE_STEPS) if replay else args.warmup)
    torch.cuda.synchronize()
    launches0 = eng.L.sph_launch_count(eng.h)
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(eng.stream)
        if replay:
            run_in_legs(solver.run_steps, restore, args.steps)
        else:
            solver.run_steps(args.steps)
        e1.record(eng.stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof = eng.profile_read()
    eng.profile(False)
    launches = eng.L.sph_launch_count(eng.h) - launches0
    bad = eng.L.sph_read_bad_cells(eng.h)
    nflag = int(eng.L.sph_read_flagged_cells(eng.h))
    value = n * args.steps / (ms * 1e-3)
    assert bool(torch.isfinite(ps.pt.v).all()), "non-finite velocities after the timed region"

    # roofline of the dominant kernel class
    peak, peak_kind = measured_peaks()
    dom = max(prof.items(), key=lambda kv: kv[1][0])
    dom_name, (dom_ms, dom_cnt) = dom
    ab = algorithmic_bytes(dom_name, n, n_fluid, n_wall, ps.grid_num_total)
    achieved = ab / (dom_ms / dom_cnt * 1e-3) / 1e9 if ab else None
    kernel_share = {k: round(v[0] / ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    kernel_ms = {k: round(v[0] / v[1], 4) for k, v in prof.items()}
    kernel_gbs = {}
    for k, v in prof.items():
        b = algorithmic_bytes(k, n, n_fluid, n_wall, ps.grid_num_total)
        if b:
            kernel_gbs[k] = round(b / (v[0] / v[1] * 1e-3) / 1e9, 1)
    step_bytes = 208 * n + 104 * n_wall + 12 * ps.grid_num_total          # SURVEY 8(d) formula, per step
    ev = ncu_evidence(dom_name) or {}
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": ev.get("dram_bytes_per_launch"),
                "traffic_source": ev.get("source"), "peak_kind": peak_kind,
                "note": "the neighbour sweeps are bound by shared-memory gather bandwidth and instruction issue, not by HBM "
                        "(SURVEY 8d): frac is the HBM view of the dominant kernel, `limiter` the ncu view of what binds it; "
                        "HBM-bound kernels (reorder, integrators) are in kernel_gbs",
                "limiter": ev.get("limiter"),
                "kernel_share_of_step": kernel_share, "kernel_ms": kernel_ms, "kernel_gbs": kernel_gbs,
                "whole_step_algorithmic_gbs": step_bytes * args.steps / (ms * 1e-3) / 1e9}

    # secondary: a STIRRED state.  The timed region above starts from the rest lattice (<= 27 particles per cell, no cell
    # flagged); here every fluid particle is displaced by U(-0.2 d, 0.2 d) per axis first, so that cells are unevenly
    # filled and the > 32-per-cell fallback to the generic kernels shows up in the number if it is hit.
    stirred = None
    if args.stirred_steps > 0:
        d = 2 * scene["Configuration"]["particleRadius"]
        g = torch.Generator(device=d_x.device).manual_seed(1234)
        jit = (torch.rand(d_x.shape, generator=g, device=d_x.device, dtype=torch.float64) - 0.5) * (0.4 * d)
        x_st = torch.where((d_typ == 1)[:, None], d_x + jit, d_x).contiguous()
        restore(x_st)
        solver.run_steps(3)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(eng.stream)
        solver.run_steps(args.stirred_steps)
        s1.record(eng.stream)
        torch.cuda.synchronize()
        sms = s0.elapsed_time(s1)
        stirred = {"ms_per_step": sms / args.stirred_steps, "value": n * args.stirred_steps / (sms * 1e-3), "steps": args.stirred_steps,
                   "warmup": 3, "flagged_cells": int(eng.L.sph_read_flagged_cells(eng.h)),
                   "finite": bool(torch.isfinite(ps.pt.v).all()),
                   "state": "fluid particles displaced by U(-0.2 d, 0.2 d) per axis from the rest lattice (seed 1234)"}

    # end-to-end through the C ABI with HOST buffers: every step uploads the particle state from pinned host memory
    # (sph_add_particles), runs sph_step(1) and reads the state back (sph_read_state*).  Steps are independent jobs, so
    # they are pipelined over `depth` contexts on their own streams (upload of job k+1 | step of job k | download of
    # job k-1 overlap; PCIe is full duplex); depth 1 is the strictly serial variant, reported alongside.
    from tisphi_b200 import _lib as L_

    def make_job(engine):
        return {"eng": engine,
                "x": torch.empty((n, 3), dtype=torch.float64).pin_memory(), "v": torch.empty((n, 4), dtype=engine.real).pin_memory(),
                "rho": torch.empty(n, dtype=torch.float64).pin_memory(), "p": torch.empty(n, dtype=engine.real).pin_memory(),
                "id": torch.empty(n, dtype=torch.int32).pin_memory()}

    def submit(job):
        e = job["eng"]
        e.call("sph_clear_particles")
        e.call("sph_add_particles", n, h_x.data_ptr(), h_v.data_ptr(), h_rho.data_ptr(), h_typ.data_ptr())
        e.call("sph_step", 1)
        e.call("sph_read_state_async", job["x"].data_ptr(), job["v"].data_ptr(), job["rho"].data_ptr(), job["p"].data_ptr(),
               job["id"].data_ptr())

    def run_e2e(jobs, steps):
        for j in jobs:                                   # untimed warm-up of every context
            submit(j)
        for j in jobs:
            j["eng"].call("sph_synchronize")
        t0 = time.perf_counter()
        for k in range(steps):
            j = jobs[k % len(jobs)]
            if k >= len(jobs):
                j["eng"].call("sph_synchronize")         # the job that used this context before has been read back
            submit(j)
        for j in jobs:
            j["eng"].call("sph_synchronize")
        return time.perf_counter() - t0

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    jobs = [make_job(eng)]
    serial_s = run_e2e(jobs, e2e_steps)
    depth = max(1, args.e2e_depth)
    for _ in range(depth - 1):
        extra = L_.Engine(eng.params, eng.n_max, device=f"cuda:{local}", stream=torch.cuda.Stream(device=local))
        jobs.append(make_job(extra))
    pipe_steps = max(e2e_steps, 3 * depth)
    e2e_s = run_e2e(jobs, pipe_steps) if depth > 1 else serial_s
    if depth == 1:
        pipe_steps = e2e_steps
    assert bool(torch.isfinite(jobs[-1]["v"]).all()) and int(jobs[-1]["id"].max()) == n - 1, "end-to-end result is not a particle state"
    h2d = n * (24 + 24 + 8 + 4)
    d2h = n * (24 + 4 * jobs[0]["v"].element_size() + 8 + jobs[0]["p"].element_size() + 4)
    e2e = {"value": n * pipe_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": pipe_steps, "ms_per_step": e2e_s / pipe_steps * 1e3, "pipeline_depth": depth,
           "serial_ms_per_step": serial_s / e2e_steps * 1e3, "serial_value": n * e2e_steps / serial_s}
    for j in jobs[1:]:
        j["eng"].close()

    cpu = None
    if not args.no_cpu:
        cpu, _, _, _ = cpu_reference(args.cpu_scale, args.cpu_steps, warmup=1)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.precision != "f64" else "f64", "data": "synthetic",
            "config": {"workload": c4_workload(n, n_fluid, n_wall, ps.grid_num_total, solver.dt[None]),
                       "scale": args.scale,
                       "precision": "mixed: fp32 sweeps, fp64 positions+densities" if args.precision != "f64" else "f64",
                       "l2": "state (>= 2 GB) larger than L2, no flush needed", "bad_cells": int(bad), "flagged_cells": nflag,
                       "legs": (f"{-(-args.steps // STABLE_STEPS)} legs of <= {STABLE_STEPS} steps from the restored initial state "
                                "(the reference's 3D scheme diverges ~35 steps after rest)") if replay else "one run from rest"},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "real_particle_updates_per_s": n_fluid * args.steps / (ms * 1e-3), "stirred": stirred}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- our arm, N > 1
def slab_parity_check(rank, local, world, log):
    """The driver's single-GPU test box cannot run the multi-GPU tests, so the bench itself asserts, on THIS box, that
    the slab-partitioned run equals the single-GPU run bit for bit: BASELINE config C1 (2D) and C4 coarsened x5 (3D),
    3 steps each, every rank one slab, rank 0 additionally the whole scene."""
    import copy
    import numpy as np
    import torch
    import torch.distributed as dist
    from tisphi_b200 import scenes
    from tisphi_b200.eng.simulation import Simulation, SimConfiger
    from tisphi_b200.parallel import SlabSimulation
    out = {}
    with open(os.path.join(ROOT, "data", "scenes", "test1_db_water.json")) as f:
        c1 = json.load(f)
    c1["Configuration"]["precision"] = "f32"
    cases = {"c1_2d": c1, "c4_coarse_3d": scenes.dambreak3d(scale=0.2, precision="f32")}
    fields = ["x", "v", "density", "pressure", "d_vel", "id0", "grid_ids"]
    out["c4_full_hash"] = slab_parity_hash(rank, local, world, log, scenes.dambreak3d(scale=1.0, precision="f32"), fields)
    for name, scene in cases.items():
        ok, mine, slab, ref = True, None, None, None
        try:                                         # whatever fails locally, every rank reaches every collective
            slab = SlabSimulation(SimConfiger(config=copy.deepcopy(scene)), f"cuda:{local}", rank, world, transport="p2p")
        except Exception as e:
            ok = False
            log(f"[rank {rank}] slab parity {name}: {type(e).__name__}: {e}")
        flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not int(flag):
            out[name] = False
            continue
        try:
            ref = Simulation(SimConfiger(config=copy.deepcopy(scene)), device=f"cuda:{local}") if rank == 0 else None
            slab.run_steps(3)
            mine = {f: slab.owned(f).detach().cpu().numpy() for f in fields}
        except Exception as e:
            ok = False
            log(f"[rank {rank}] slab parity {name}: {type(e).__name__}: {e}")
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0 and ok and all(g is not None for g in gathered):
            ref.solver.run_steps(3)
            torch.cuda.synchronize()
            for f in fields:
                got = np.concatenate([g[f] for g in gathered])
                want = getattr(ref.ps.pt, f).detach().cpu().numpy()
                if got.shape != want.shape or not np.array_equal(got, want):
                    ok = False
                    log(f"slab parity: {name}: {f} differs from the single-GPU run")
        elif any(g is None for g in gathered):
            ok = False
        torch.cuda.synchronize()
        dist.barrier()
        slab.close()
        flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out[name] = bool(int(flag))
    return out


def state_hash(torch, cols, id0):
    """Order-independent 64-bit hash of per-particle state: every word of every field folded per particle (LCG steps,
    wrap-around int64), weighted by the particle's creation index, summed.  Equal sums <=> bit-identical state (w.h.p.)."""
    n = id0.numel()
    acc = torch.zeros(n, dtype=torch.int64, device=id0.device)
    for t in cols:
        t = t.contiguous()
        w = (t.view(torch.int64) if t.element_size() == 8 else t.view(torch.int32).to(torch.int64)).reshape(n, -1)
        for j in range(w.shape[1]):
            acc = acc * 6364136223846793005 + w[:, j] + 1442695040888963407
    acc = (acc ^ (acc >> 31)) * (2 * id0.to(torch.int64) + 1)
    return acc.sum()


def slab_parity_hash(rank, local, world, log, scene, fields, steps=3):
    """The benchmarked scene itself at FULL size: `steps` steps on the slabs and (rank 0) on one GPU, compared through
    state_hash of the owned particles (the arrays are too large to gather through the host)."""
    import copy
    import torch
    import torch.distributed as dist
    from tisphi_b200.eng.simulation import Simulation, SimConfiger
    from tisphi_b200.parallel import SlabSimulation
    ok, slab, h = True, None, torch.zeros(2, dtype=torch.int64, device=f"cuda:{local}")
    try:
        slab = SlabSimulation(SimConfiger(config=copy.deepcopy(scene)), f"cuda:{local}", rank, world, transport="p2p")
    except Exception as e:
        ok = False
        log(f"[rank {rank}] slab parity (full): {type(e).__name__}: {e}")
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not int(flag):
        return False
    try:
        slab.run_steps(steps)
        slab.sync()
        cols = [slab.owned(f) for f in fields if f != "id0"]
        h[0] = state_hash(torch, cols, slab.owned("id0"))
        h[1] = slab.owned("id0").numel()
    except Exception as e:
        ok = False
        log(f"[rank {rank}] slab parity (full): {type(e).__name__}: {e}")
    dist.all_reduce(h)                                   # int64 sums wrap around
    torch.cuda.synchronize()
    dist.barrier()
    slab.close()
    del slab
    torch.cuda.empty_cache()
    if rank == 0 and ok:
        try:
            ref = Simulation(SimConfiger(config=copy.deepcopy(scene)), device=f"cuda:{local}")
            ref.solver.run_steps(steps)
            torch.cuda.synchronize()
            want = state_hash(torch, [getattr(ref.ps.pt, f) for f in fields if f != "id0"], ref.ps.pt.id0)
            ok = int(h[1]) == ref.ps.pt.id0.numel() and int(want) == int(h[0])
            if not ok:
                log(f"slab parity (full): hash {int(h[0])} over {int(h[1])} particles != single-GPU {int(want)} over {ref.ps.pt.id0.numel()}")
            del ref
            torch.cuda.empty_cache()
        except Exception as e:
            ok = False
            log(f"slab parity (full) reference: {type(e).__name__}: {e}")
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(int(flag))


def run_ours_multi(args, rank, local, world):
    """C4 slab-partitioned over `world` GPUs (strong scaling): one process per GPU."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from tisphi_b200 import scenes
    from tisphi_b200.eng.simulation import SimConfiger
    from tisphi_b200.parallel import SlabSimulation

    if args.workload == "c5":
        from bench_workloads import c5 as bench_c5
        bench_c5.run_multi(args, rank, local, world, log, ClockSampler, measured_peaks)
        dist.destroy_process_group()
        return
    parity = slab_parity_check(rank, local, world, log) if not args.no_parity else None
    scene = scenes.dambreak3d(scale=args.scale, precision=args.precision, **scene_options(args))
    t0 = time.time()
    transport = args.transport
    try:
        sim = SlabSimulation(SimConfiger(config=scene), f"cuda:{local}", rank, world, transport=transport, wall_weight=args.wall_weight)
    except RuntimeError as e:                           # raised on every rank together (parallel.py::_connect_p2p)
        if transport != "p2p":
            raise
        log(f"[rank {rank}] {e}; falling back to torch.distributed send / recv")
        transport = "dist"
        sim = SlabSimulation(SimConfiger(config=scene), f"cuda:{local}", rank, world, transport=transport, wall_weight=args.wall_weight)
    eng, drv, ps = sim.ps.engine, sim.driver, sim.ps
    n_global = ps.global_particle_num
    n_own0 = eng.n
    log(f"[rank {rank}] slab built: columns {sim.columns}, {n_own0} of {n_global} particles, transport {transport}, {time.time() - t0:.1f}s")

    # host copy of this rank's initial state for the end-to-end leg (pinned); device copy for the legs
    pin = lambda t: t.detach().cpu().contiguous().pin_memory()
    h_x, h_rho, h_typ, h_id = pin(ps.pt.x), pin(ps.pt.density), pin(ps.pt.mat_type), pin(ps.pt.id0)
    h_v = pin(ps.pt.v.double())
    d_x, d_rho, d_typ = ps.pt.x.clone().contiguous(), ps.pt.density.clone().contiguous(), ps.pt.mat_type.clone().contiguous()
    d_v, d_id = ps.pt.v.double().contiguous(), ps.pt.id0.clone().contiguous()
    n_init = len(d_rho)
    nf_init = int((d_typ == 1).sum())

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def restore():
        eng.call("sph_clear_particles")
        eng.call("sph_add_particles", n_init, d_x.data_ptr(), d_v.data_ptr(), d_rho.data_ptr(), d_typ.data_ptr())
        eng.field("ID0", count=n_init).copy_(d_id)
        drv.reset()

    replay = args.warmup + args.steps > STRAIGHT_STEPS
    sim.run_steps(min(args.warmup, STABLE_STEPS) if replay else args.warmup)
    barrier()
    launches0 = eng.L.sph_launch_count(eng.h)
    ex0 = drv.exchanges
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(eng.stream)
        if replay:
            run_in_legs(sim.run_steps, restore, args.steps)
        else:
            sim.run_steps(args.steps)
        e1.record(eng.stream)
        barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms = float(ms_t)
    launches = eng.L.sph_launch_count(eng.h) - launches0
    exchanges = drv.exchanges - ex0
    sim.sync()                                          # raises if the slab step set an error bit
    own_count = drv.own_count
    finite = bool(torch.isfinite(ps.pt.v).all())
    value = n_global * args.steps / (ms * 1e-3)

    # per-rank timeline (a separate, profiled leg: CUDA events around every kernel class): compute, message kernels, waits
    tl_steps = 4 if (replay or args.warmup + args.steps + 4 <= STRAIGHT_STEPS) else 0
    timeline = None
    if tl_steps:
        if replay:
            restore()
            sim.run_steps(2)
        barrier()
        eng.profile(True)
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record(eng.stream)
        sim.run_steps(tl_steps)
        t1e.record(eng.stream)
        barrier()
        prof = eng.profile_read()
        eng.profile(False)
        tot = t0e.elapsed_time(t1e) / tl_steps
        wait = prof.get("halo_wait", (0.0, 0))[0] / tl_steps
        halo = prof.get("halo", (0.0, 0))[0] / tl_steps
        comp = sum(v[0] for k, v in prof.items() if k not in ("halo", "halo_wait")) / tl_steps
        timeline = {"step_ms": round(tot, 3), "compute_ms": round(comp, 3), "message_kernels_ms": round(halo, 3),
                    "wait_ms": round(wait, 3), "kernel_ms": {k: round(v[0] / tl_steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
    else:
        prof = {}
    stats = {"launches": int(launches), "own": int(own_count), "exchanges": int(exchanges), "ms": float(e0.elapsed_time(e1)),
             "finite": finite, "timeline": timeline, "n_fluid": nf_init}
    allv = [None] * world
    dist.all_gather_object(allv, stats)
    assert all(v["finite"] for v in allv), "non-finite velocities after the timed region"
    assert sum(v["own"] for v in allv) == n_global, "particles lost or duplicated by the migration"

    # end to end: every step uploads this rank's particles from pinned host memory and reads its owned state back
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    n0 = len(h_rho)
    cap = eng.n_max
    out_x = torch.empty((cap, 3), dtype=torch.float64).pin_memory()
    out_v = torch.empty((cap, 4), dtype=eng.real).pin_memory()
    out_rho = torch.empty(cap, dtype=torch.float64).pin_memory()
    out_p = torch.empty(cap, dtype=eng.real).pin_memory()
    out_id = torch.empty(cap, dtype=torch.int32).pin_memory()

    def e2e_step():
        eng.call("sph_clear_particles")
        eng.call("sph_add_particles", n0, h_x.data_ptr(), h_v.data_ptr(), h_rho.data_ptr(), h_typ.data_ptr())
        eng.field("ID0", count=n0).copy_(h_id, non_blocking=True)
        drv.reset()
        sim.run_steps(1)
        sim.sync()                                      # the host learns the particle count (owned + ghosts)
        eng.call("sph_read_state", out_x.data_ptr(), out_v.data_ptr(), out_rho.data_ptr(), out_p.data_ptr(), out_id.data_ptr())

    e2e_step()
    # where an end-to-end step spends its time (one extra, untimed step with a synchronize between the parts)
    barrier()
    tb = [time.perf_counter()]
    eng.call("sph_clear_particles")
    eng.call("sph_add_particles", n0, h_x.data_ptr(), h_v.data_ptr(), h_rho.data_ptr(), h_typ.data_ptr())
    eng.field("ID0", count=n0).copy_(h_id, non_blocking=True)
    torch.cuda.synchronize(); tb.append(time.perf_counter())
    drv.reset()
    sim.run_steps(1)
    sim.sync(); tb.append(time.perf_counter())
    eng.call("sph_read_state", out_x.data_ptr(), out_v.data_ptr(), out_rho.data_ptr(), out_p.data_ptr(), out_id.data_ptr())
    tb.append(time.perf_counter())
    e2e_parts = {"upload_ms": round((tb[1] - tb[0]) * 1e3, 3), "step_ms": round((tb[2] - tb[1]) * 1e3, 3), "readback_ms": round((tb[3] - tb[2]) * 1e3, 3)}
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t)
    io = torch.tensor([n0 * (24 + 24 + 8 + 4 + 4), eng.n * (24 + 4 * out_v.element_size() + 8 + out_p.element_size() + 4)],
                      dtype=torch.int64, device=f"cuda:{local}")
    dist.all_reduce(io)
    if rank == 0:
        peak, peak_kind = measured_peaks()
        roofline = None
        tl0 = allv[0]["timeline"]
        if tl0:
            km = {k: v for k, v in tl0["kernel_ms"].items() if k not in ("halo", "halo_wait")}
            dom_name = max(km, key=km.get)
            n_loc = allv[0]["own"]
            nf_loc = allv[0]["n_fluid"]
            ab = algorithmic_bytes(dom_name, n_loc, nf_loc, n_loc - nf_loc, ps.grid_num_total)
            per_launch_ms = km[dom_name] / (2 if dom_name in ("tile_fluid", "tile_wall", "wc_eos") else 1)
            achieved = ab / (per_launch_ms * 1e-3) / 1e9 if ab else None
            roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak if achieved else None, "traffic": None, "peak_kind": peak_kind,
                        "rank": 0, "note": "rank 0's slab, from the profiled timeline leg; neighbour sweeps are fp32-issue / shared-memory bound, not HBM bound"}
        tls = [v["timeline"] for v in allv]
        limiter = None
        if all(tls):
            comp = [t["compute_ms"] for t in tls]
            step = max(t["step_ms"] for t in tls)
            limiter = {"max_compute_ms": max(comp), "mean_compute_ms": round(sum(comp) / world, 3),
                       "imbalance": round(max(comp) / (sum(comp) / world), 3),
                       "max_message_kernels_ms": max(t["message_kernels_ms"] for t in tls),
                       "min_wait_ms": min(t["wait_ms"] for t in tls), "step_ms": step,
                       "reading": "step = slowest rank's compute + its message kernels + the wait nobody can avoid (min_wait); "
                                  "the other ranks' extra wait is load imbalance"}
        n_fluid_g, cells = 10240000, ps.grid_num_total
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if args.precision != "f64" else "f64", "data": "synthetic",
                "config": {"workload": (C4_WORKLOAD_FULL if args.scale == 1.0 else f"C4 3D WCSPH dambreak scale={args.scale}: N={n_global}, cells={cells}"),
                           "partition": f"slab-partitioned over {world} GPUs along x (strong scaling: the same scene on every N)",
                           "scale": args.scale, "transport": ("device-driven slab step, messages stored into peer-mapped inboxes over NVLink (csrc/slab.cu)"
                                                              if transport == "p2p" else "torch.distributed send / recv driven from Python"),
                           "precision": "mixed: fp32 sweeps, fp64 positions+densities" if args.precision != "f64" else "f64",
                           "slab_parity": (all(parity.values()) if parity else None), "slab_parity_cases": parity,
                           "columns": [list(c) for c in ps.slab_columns],
                           "owned_particles": [v["own"] for v in allv],
                           "exchanges_per_step": allv[0]["exchanges"] / args.steps,
                           "rank_ms_per_step": [round(v["ms"] / args.steps, 3) for v in allv],
                           "rank_timeline_ms_per_step": tls, "limiter": limiter,
                           "l2": "state larger than L2, no flush needed",
                           "legs": "legs of <= 20 steps from the restored initial state" if replay else "one run from rest"},
                "clocks": clocks.summary(),
                "e2e": {"value": n_global * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(io[0]),
                        "d2h_bytes_per_step": int(io[1]), "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                        "rank0_parts_ms": e2e_parts},
                "gpu_launches": int(sum(v["launches"] for v in allv)), "roofline": roofline, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if transport == "p2p":
        barrier()
        sim.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c3", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="lattice refinement of the C4 scene (1 = 12.96 M particles)")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-depth", type=int, default=4, help="contexts the end-to-end jobs are pipelined over (1 = serial)")
    ap.add_argument("--cpu-scale", type=float, default=1.0, help="coarsening of the CPU arm's scene (1 = the benchmarked scene itself)")
    ap.add_argument("--cpu-steps", type=int, default=2, help="timed steps of the cpu_baseline leg (about 10 s each on 32 threads)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the slab == single-GPU bit-exactness check")
    ap.add_argument("--stirred-steps", type=int, default=10, help="timed steps of the secondary, jittered-state measurement (0 = off)")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "dist"])
    ap.add_argument("--wall-weight", type=float, default=0.15, help="work of a wall particle relative to a fluid particle (column partition)")
    ap.add_argument("--size", type=float, default=None, help="c5: particles (default 1e7); c3: refinement of the test2 geometry")
    ap.add_argument("--soil", default="dp", choices=["dp", "mui"], help="c3 workload: Drucker-Prager + CSPM + RK4 (C3) or mu(I) + LF (C2)")
    ap.add_argument("--lists", type=int, default=None, help="1 / 0: neighbour round lists on / off (default: engine default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
