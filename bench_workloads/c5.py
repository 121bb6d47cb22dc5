"""bench.py --workload c5: BASELINE config C5, the synthetic 3D uniform box.

One "update" = one particle through grid build (cell id, histogram, scan, stable counting sort, reorder) + neighbour
count + density sum -- the bare for_all_neighbors iteration of the reference (eng/particle_system.py:216-269) with the
density task of eng/solver_sph_wc.py:30-31.  N > 1: the same box slab-partitioned along x (strong scaling), every sweep
preceded by the migration / halo exchange of the native slab step (csrc/slab.cu).
"""
import json
import os
import time

import numpy as np

METRIC, UNIT = "particle-updates/s", "particle-updates/s"


def _cpu_baseline(n_target, threads, log):
    """oracle (float64, OpenMP): grid build + neighbour count + density sum on a bounded box."""
    from oracle import oracle as orc
    from tisphi_b200.c5 import box_positions, box_params
    L = orc.lib()
    L.orc_set_threads(threads)
    x, n_side = box_positions(n_target)
    Pe = box_params(n_side)
    P = orc.OrcParams()
    P.dim, P.kernel, P.kcorr, P.ti, P.xsph, P.solver, P.serial, P.wc_fresh = 3, 1, 0, 1, 0, 1, 0, 0
    for a in range(3):
        P.gn[a], P.vstart[a], P.g[a] = Pe.gn[a], Pe.vstart[a], 0.0
    P.h, P.support, P.grid_size, P.m_V0, P.eps = Pe.h, Pe.support, Pe.grid_size, Pe.m_V0, 1e-8
    P.dt, P.rho0, P.visc, P.stiff, P.gamma_, P.vsound = Pe.dt, Pe.rho0, Pe.visc, Pe.stiff, Pe.gamma_, Pe.vsound
    o = orc.Oracle(P, x, np.zeros_like(x), np.ones(len(x)), np.ones(len(x), dtype=np.int32))
    o.grid_build(); o.neighbor_count(); o.density_sum()          # warm-up
    reps, t0 = 2, time.perf_counter()
    for _ in range(reps):
        o.grid_build(); o.neighbor_count(); o.density_sum()
    dt = (time.perf_counter() - t0) / reps
    return {"value": len(x) / dt, "unit": UNIT, "cores": int(L.orc_max_threads()), "kind": "port",
            "sample": f"uniform box N={len(x)}, {reps} sweeps (sort + count + density as separate oracle calls), float64, "
                      f"OpenMP omp_get_max_threads()={int(L.orc_max_threads())}, {dt:.2f} s per sweep"}


def run(args, local, log, ClockSampler, measured_peaks, host_threads=None):
    import torch
    from tisphi_b200.c5 import UniformBox
    n_target = int(args.size or 1e7)
    t0 = time.time()
    box = UniformBox(n_target, device=f"cuda:{local}")
    eng, n = box.engine, box.n
    cells = box.params.gn[0] * box.params.gn[1] * box.params.gn[2]
    log(f"box built: N={n} ({box.n_side}^3), cells={cells}, {time.time() - t0:.1f}s")
    for _ in range(max(3, args.warmup)):
        box.sweep()
    torch.cuda.synchronize()
    launches0 = eng.L.sph_launch_count(eng.h)
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(eng.stream)
        for _ in range(args.steps):
            box.sweep()
        e1.record(eng.stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof = eng.profile_read()
    eng.profile(False)
    launches = eng.L.sph_launch_count(eng.h) - launches0
    value = n * args.steps / (ms * 1e-3)
    cnt = box.count
    peak, peak_kind = measured_peaks()
    kernel_ms = {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    # per-kernel algorithmic bytes: the grid-build kernels are the HBM-bound ones
    ab = {"cell_id": 32 * n, "scan": 12 * cells, "scatter_index": 12 * n, "rank": 12 * n, "reorder": 220 * n,
          "tile_mask": 16 * n, "c5_sweep": 20 * n + 8 * n}
    kernel_gbs = {k: round(ab[k] / (v * 1e-3) / 1e9, 1) for k, v in kernel_ms.items() if k in ab and v > 0}
    grid_ms = sum(kernel_ms.get(k, 0.0) for k in ("cell_id", "scan", "scatter_index", "rank", "reorder"))
    grid_bytes = 92 * n + 12 * cells
    dom = next(iter(kernel_ms))
    achieved = ab[dom] / (kernel_ms[dom] * 1e-3) / 1e9 if dom in ab else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": None, "peak_kind": peak_kind,
                "kernel_ms_per_sweep": kernel_ms, "kernel_gbs": kernel_gbs,
                "grid_build": {"ms": round(grid_ms, 4), "algorithmic_bytes": grid_bytes, "gbs": round(grid_bytes / (grid_ms * 1e-3) / 1e9, 1),
                               "frac_of_hbm": round(grid_bytes / (grid_ms * 1e-3) / 1e9 / peak, 4)},
                "whole_sweep_algorithmic_gbs": (72 * n + 12 * cells) * args.steps / (ms * 1e-3) / 1e9,
                "note": "the mask and sweep kernels are fp32-pipe / shared-memory bound (SURVEY 8d); the grid-build kernels are the HBM-bound part"}

    # end to end: positions from pinned host memory -> sweep -> counts and densities back to the host
    h_x = torch.from_numpy(box.x).pin_memory()
    h_v = torch.zeros((n, 3), dtype=torch.float64).pin_memory()
    h_rho = torch.ones(n, dtype=torch.float64).pin_memory()
    h_typ = torch.ones(n, dtype=torch.int32).pin_memory()
    o_cnt = torch.empty(n, dtype=torch.int32).pin_memory()
    o_rho = torch.empty(n, dtype=torch.float32).pin_memory()

    def e2e_step():
        eng.call("sph_clear_particles")
        eng.call("sph_add_particles", n, h_x.data_ptr(), h_v.data_ptr(), h_rho.data_ptr(), h_typ.data_ptr())
        box.sweep()
        with torch.cuda.stream(eng.stream):
            o_cnt.copy_(box.count, non_blocking=True)
            o_rho.copy_(box.rho, non_blocking=True)
        eng.call("sph_synchronize")

    e2e_step()
    k = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(k):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    assert int(o_cnt.max()) == int(cnt.max())
    cpu = None if args.no_cpu else _cpu_baseline(min(n_target, 2_000_000), host_threads() if host_threads else os.cpu_count(), log)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"C5 synthetic 3D uniform box: N={n} ({box.n_side}^3 jittered lattice, seed 1234), cells={cells}; "
                                   "one step = grid build + neighbour count + density sum",
                       "mean_neighbours": float(cnt.float().mean()), "max_neighbours": int(cnt.max()),
                       "flagged_cells": int(eng.L.sph_read_flagged_cells(eng.h)), "l2": "state larger than L2, no flush needed"},
            "clocks": clocks.summary(),
            "e2e": {"value": n * k / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * 60, "d2h_bytes_per_step": n * 8, "steps": k,
                    "ms_per_step": e2e_s / k * 1e3},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def run_multi(args, rank, local, world, log, ClockSampler, measured_peaks):
    """The same box slab-partitioned along x over `world` GPUs (strong scaling): every sweep = sph_grid_build on the slab
    (selection of the face columns, push into the neighbours' inboxes, ONE sort of the virtual concatenation) +
    sph_density_sweep on the owned columns.  One process per GPU, csrc/slab.cu."""
    import torch
    import torch.distributed as dist
    from tisphi_b200 import _lib
    from tisphi_b200.c5 import box_positions, box_params
    from tisphi_b200.parallel import partition_columns, connect_p2p
    n_target = int(args.size or 1e7)
    x, n_side = box_positions(n_target)
    n = len(x)
    P = box_params(n_side)
    gs, ncol = P.grid_size, int(P.gn[0])
    cx = np.clip(((x[:, 0] - P.vstart[0]) / gs).astype(np.int64), 0, ncol - 1)
    counts = np.bincount(cx, minlength=ncol)
    cols = partition_columns(counts.astype(np.float64), world)
    a, b = cols[rank]
    mine = np.nonzero((cx >= a) & (cx < b))[0]
    cap = int(1.3 * (len(mine) + counts[max(a - 1, 0):a].sum() + counts[b:b + 1].sum())) + 4 * int(counts.max()) + 1024
    eng = _lib.Engine(P, cap, device=f"cuda:{local}")
    eng.add_particles(x[mine], np.zeros((len(mine), 3)), np.ones(len(mine)), np.ones(len(mine), dtype=np.int32))
    eng.field("ID0").copy_(torch.from_numpy(mine.astype(np.int32)).to(eng.device))
    del x
    face_cap = 2 * int(counts.max()) + 4096
    drv, inbox, ipc = connect_p2p(eng, rank, world, (a, b), face_cap)
    count = torch.empty(cap, dtype=torch.int32, device=eng.device)
    rho = torch.empty(cap, dtype=eng.real, device=eng.device)

    def sweep():
        eng.call("sph_grid_build")
        eng.call("sph_density_sweep", count.data_ptr(), rho.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        sweep()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(eng.stream)
        for _ in range(args.steps):
            sweep()
        e1.record(eng.stream)
        barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=eng.device)
    dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms = float(ms_t)
    drv.sync()
    own = slice(drv.own_first, drv.own_first + drv.own_count)
    tot = torch.tensor([drv.own_count, int(count[own].long().sum()), int(count[own].max())], dtype=torch.int64, device=eng.device)
    allv = [torch.zeros_like(tot) for _ in range(world)]
    dist.all_gather(allv, tot)
    assert sum(int(v[0]) for v in allv) == n, "particles lost or duplicated by the slab exchange"
    pairs = sum(int(v[1]) for v in allv)
    assert pairs % 2 == 0, "the neighbour relation across the slab faces is not symmetric"
    if rank == 0:
        cells = int(P.gn[0]) * int(P.gn[1]) * int(P.gn[2])
        line = {"metric": METRIC, "value": n * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"C5 synthetic 3D uniform box: N={n} ({n_side}^3 jittered lattice, seed 1234), cells={cells}; "
                                       "one step = grid build + neighbour count + density sum",
                           "partition": f"slab-partitioned over {world} GPUs along x; every step starts with the face exchange of csrc/slab.cu",
                           "columns": [list(c) for c in cols], "owned_particles": [int(v[0]) for v in allv],
                           "mean_neighbours": pairs / n, "max_neighbours": max(int(v[2]) for v in allv),
                           "l2": "state larger than L2, no flush needed"},
                "clocks": clocks.summary(), "e2e": None, "gpu_launches": int(eng.L.sph_launch_count(eng.h)), "roofline": None,
                "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    barrier()
    for p in ipc:
        eng.L.sph_ipc_close(p)
    eng.L.sph_ipc_free(inbox)
