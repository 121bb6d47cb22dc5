"""bench.py --workload c3: the soil solvers on BASELINE config C3's geometry, refined.

Scene: data/scenes/test2_cc_sand_dp_rk4_cspm.json (2D granular column collapse: Drucker-Prager + CSPM + RK4 + XSPH,
dummy walls) with the particle radius divided by `--size` (default 18: 1800 x 900 = 1 620 000 soil + 21 654 dummy
particles).  `--soil mui` runs BASELINE config C2 instead (mu(I) + "LF" + XSPH on the same geometry).
One "step" = one SPHBase.step(): grid build, CSPM_f (+ CSPM_L), four (two) one_steps of three sweeps each, the RK4
("LF") integrator kernels, advect_pos with XSPH, the post-step (return mapping / regularisation sweep).
"""
import copy
import json
import os
import time

METRIC, UNIT = "particle-updates/s", "particle-updates/s"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def soil_scene(refine, soil="dp", precision="f32"):
    name = "test2_cc_sand_dp_rk4_cspm.json" if soil == "dp" else "test2_cc_sand_muI.json"
    with open(os.path.join(ROOT, "data", "scenes", name)) as f:
        sc = json.load(f)
    sc["Configuration"]["particleRadius"] = sc["Configuration"]["particleRadius"] / refine
    sc["Configuration"]["precision"] = precision
    return sc


def _cpu_baseline(refine, soil, steps, threads):
    from oracle import oracle as orc
    L = orc.lib()
    L.orc_set_threads(threads)
    o = orc.Oracle.from_scene(soil_scene(refine, soil, "f64"), serial=0)
    o.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        assert o.step() == 0
    dt = time.perf_counter() - t0
    return {"value": o.n * steps / dt, "unit": UNIT, "cores": int(L.orc_max_threads()), "kind": "port",
            "sample": f"the same geometry refined x{refine:g} (N={o.n}), {steps} steps after 1 warm-up, float64, "
                      f"OpenMP omp_get_max_threads()={int(L.orc_max_threads())}, {dt:.1f} s; restated CPU baseline"}


def run(args, local, log, ClockSampler, measured_peaks, host_threads):
    import torch
    from tisphi_b200.eng.simulation import Simulation, SimConfiger
    from tisphi_b200 import _lib as L_
    refine = float(args.size or 18)
    soil = getattr(args, "soil", "dp")
    scene = soil_scene(refine, soil, args.precision)
    t0 = time.time()
    sim = Simulation(SimConfiger(config=copy.deepcopy(scene)), device=f"cuda:{local}")
    ps, solver, eng = sim.ps, sim.solver, sim.ps.engine
    n = ps.particle_num[None]
    n_soil = int((ps.pt.mat_type == 2).sum())
    log(f"scene built: N={n} ({n_soil} soil), cells={ps.grid_num_total}, dt={solver.dt[None]!r}, {time.time() - t0:.1f}s")
    pin = lambda t: t.detach().cpu().contiguous().pin_memory()
    h_x, h_rho, h_typ = pin(ps.pt.x), pin(ps.pt.density), pin(ps.pt.mat_type)
    h_v = pin(ps.pt.v.double())
    solver.run_steps(args.warmup)
    torch.cuda.synchronize()
    launches0 = eng.L.sph_launch_count(eng.h)
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(eng.stream)
        solver.run_steps(args.steps)
        e1.record(eng.stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof = eng.profile_read()
    eng.profile(False)
    launches = eng.L.sph_launch_count(eng.h) - launches0
    value = n * args.steps / (ms * 1e-3)
    assert bool(torch.isfinite(ps.pt.v).all()), "non-finite velocities after the timed region"
    peak, peak_kind = measured_peaks()
    kernel_ms = {k: round(v[0] / v[1], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    share = {k: round(v[0] / ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    dom = next(iter(kernel_ms))
    # soil sweep: read the 68-byte state + stage values of every particle, write derivatives (40 B) per soil particle
    ab = {"dp_soil": 68 * n + 40 * n_soil, "mui_soil1": 68 * n + 40 * n_soil, "mui_soil3": 68 * n + 16 * n_soil,
          "soil_wall": 104 * (n - n_soil), "reorder": 300 * n, "advect": 152 * n_soil, "tile_mask": 16 * n,
          "tile_soil": 68 * n + 40 * n_soil, "tile_soil_wall": 104 * (n - n_soil)}
    achieved = ab[dom] / (kernel_ms[dom] * 1e-3) / 1e9 if dom in ab else None
    step_bytes = 650 * n if soil == "dp" else 400 * n
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": None, "peak_kind": peak_kind,
                "kernel_ms": kernel_ms, "kernel_share_of_step": share,
                "whole_step_algorithmic_gbs": step_bytes * args.steps / (ms * 1e-3) / 1e9,
                "note": "neighbour sweeps are fp32-issue / gather bound, not HBM bound (SURVEY 8d)"}

    # end to end: upload the particle state from pinned host memory, initial stress, one step, state back to the host
    out = {"x": torch.empty((n, 3), dtype=torch.float64).pin_memory(), "v": torch.empty((n, 4), dtype=eng.real).pin_memory(),
           "rho": torch.empty(n, dtype=torch.float64).pin_memory(), "id": torch.empty(n, dtype=torch.int32).pin_memory()}

    def e2e_step():
        eng.call("sph_clear_particles")
        eng.call("sph_add_particles", n, h_x.data_ptr(), h_v.data_ptr(), h_rho.data_ptr(), h_typ.data_ptr())
        if soil == "dp":
            eng.call("sph_init_stress")
        eng.call("sph_step", 1)
        eng.call("sph_read_state", out["x"].data_ptr(), out["v"].data_ptr(), out["rho"].data_ptr(), None, out["id"].data_ptr())

    e2e_step()
    k = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(k):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    cpu = None if args.no_cpu else _cpu_baseline(min(refine, 6.0), soil, 2, host_threads())
    cfg = scene["Configuration"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.precision != "f64" else "f64", "data": "synthetic",
            "config": {"workload": (f"C3 2D granular column collapse, Drucker-Prager + CSPM + RK4 + XSPH" if soil == "dp" else
                                    "C2 2D granular column collapse, mu(I) + LF + XSPH") +
                                   f" (test2_cc_sand geometry refined x{refine:g}): N={n} ({n_soil} soil), cells={ps.grid_num_total}, dt={solver.dt[None]!r}",
                       "precision": "mixed: fp32 sweeps, fp64 positions+densities" if args.precision != "f64" else "f64",
                       "sweeps": "cell-tile" if getattr(eng.params, "fast", 0) and cfg.get("soilTiles", True) else "generic",
                       "l2": "state larger than L2, no flush needed"},
            "clocks": clocks.summary(),
            "e2e": {"value": n * k / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * 60, "d2h_bytes_per_step": n * (24 + 4 * out["v"].element_size() + 8 + 4),
                    "steps": k, "ms_per_step": e2e_s / k * 1e3},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
