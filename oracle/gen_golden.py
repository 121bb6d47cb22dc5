#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE'S OWN SOURCES.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference); nothing on the GPU box or in
the product path imports it.  The reference (Rabmelon/tiSPHi, eng/*.py) is imported unmodified; its only missing
dependency, Taichi 1.2.2, is replaced by the serial emulator in oracle/ti_shim (see that module's header for the
semantics assumed).  The run therefore has the reference's single-threaded semantics
(``ti.init(arch=ti.cpu, cpu_max_num_threads=1)``, run_simulation.py:22): top-level loops execute in index order,
in-place reads see earlier iterations' writes.

Each fixture is one ``.npz`` holding, for every requested step s:
  ``s{step}/grid/...``  state right after ``initialize_particle_system()`` + ``calc_kernel_corr()`` of that step
                        (captured by wrapping ``solver.init_real2tmp``; ``SPHBase.step`` itself is untouched):
                        grid_ids, id0 (sorted order), grid_particle_num (inclusive offsets), CSPM_f, CSPM_L,
                        neighbour counts obtained through the reference's own ``for_all_neighbors``;
  ``s{step}/end/...``   every dynamic member of the particle struct after ``step()`` returned;
plus ``meta`` (json: scene, dt, constants, shim OOB read count).

Usage:  python oracle/gen_golden.py --list | --case NAME [--case NAME ...] | --all [--jobs 8]
"""
import argparse
import copy
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("TISPHI_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")


def _scene(base, cfg=None, block=None, materials=None):
    with open(os.path.join(REF, "data", "scenes", base)) as f:
        sc = json.load(f)
    sc["Configuration"].update(cfg or {})
    if block:
        sc["Blocks"][0].update(block)
    if materials is not None:
        sc["Materials"] = materials
    return sc


def _indenter(water=False):
    with open(os.path.join(REF, "data", "scenes", "test4_in_ver.json")) as f:
        sc = json.load(f)
    sc["Configuration"].update(particleRadius=0.0005, domainEnd=[0.05, 0.03, 0.05])
    sc["Blocks"][0].update(size=[0.05, 0.02, 0.05])
    sc["Blocks"][1].update(translation=[0.021, 0.02, 0.0], size=[0.008, 0.004, 0.05])
    if water:
        with open(os.path.join(REF, "data", "scenes", "test1_db_water.json")) as f:
            w = json.load(f)
        fluid = dict(w["Materials"][0], matId=0)
        rigid = [m for m in sc["Materials"] if m["matType"] == 11][0]
        sc["Materials"] = [fluid, dict(rigid, matId=1)]
        sc["Configuration"].update(simulationMethod=1, xsph=False, timeStepSizeMin=w["Configuration"]["timeStepSizeMin"])
        sc["Blocks"][1].update(velocity=[0.1, -0.25, 0.0])
    return sc


def _dynrigid(method=3, wall=False):
    """The shrunken test4 scene with the rigid block DYNAMIC: it falls onto / is thrown into the soil bed,
    feels the reaction of the momentum sums (dp:164-165, muI:45-46), is shape-matched back to a rigid transform every
    step (base:472-499) and, with ``wall``, starts against the left domain face (collision clamp, base:525-601)."""
    sc = _indenter()
    sc["Configuration"].update(simulationMethod=method)
    sc["Blocks"][1].update(isDynamic=1, velocity=[0.3, -0.25, 0.0])
    if wall:
        sc["Blocks"][1].update(translation=[0.0, 0.0205, 0.0], velocity=[-0.5, -0.1, 0.0])
    return sc


def _plate():
    with open(os.path.join(REF, "data", "scenes", "test5_in_hor.json")) as f:
        sc = json.load(f)
    sc["Configuration"].update(particleRadius=0.0005, domainEnd=[0.04, 0.025, 0.05])
    b = sc["Blocks"]
    b[0].update(translation=[0, 0, 0], size=[0.04, 0.012, 0.05])
    b[1].update(translation=[0, 0.016, 0], size=[0.04, 0.004, 0.05])
    b[2].update(translation=[0, 0.012, 0], size=[0.01, 0.004, 0.05])
    b[3].update(translation=[0.012, 0.012, 0], size=[0.028, 0.004, 0.05])
    b[4].update(translation=[0.01, 0.012, 0], size=[0.002, 0.004, 0.05])
    return sc


WATER3D = dict(is2D=False, particleRadius=0.01, domainStart=[0.0, 0.0, 0.0], domainEnd=[0.24, 0.2, 0.18],
               timeStepSizeMin=1e-6)

# name -> (scene dict factory, steps to snapshot)
CASES = {
    # full shipped scene (BASELINE config C1): N = 6422
    "c1_test1_wc_lf": (lambda: _scene("test1_db_water.json"), [1, 2, 3]),
    # shrunken dambreak, same parameters, long horizon
    "wc2d_small_lf": (lambda: _scene("test1_db_water.json", dict(domainEnd=[1.0, 0.6, 0.5]),
                                     dict(size=[0.4, 0.3, 0.1])), [1, 2, 10, 50, 100]),
    "wc2d_small_se_cubic": (lambda: _scene("test1_db_water.json",
                                           dict(domainEnd=[1.0, 0.6, 0.5], timeIntegration=1, kernel=0),
                                           dict(size=[0.4, 0.3, 0.1])), [1, 2, 10]),
    "wc2d_small_rk4_cspm": (lambda: _scene("test1_db_water.json",
                                           dict(domainEnd=[1.0, 0.6, 0.5], timeIntegration=4, kernelCorrection=1),
                                           dict(size=[0.4, 0.3, 0.1])), [1, 2, 10]),
    # BASELINE config C2: mu(I) on the test2 geometry (shrunken: 40 x 25 soil particles)
    "mui2d_small_lf": (lambda: _scene("test2_cc_sand.json",
                                      dict(domainEnd=[0.2, 0.08, 0.05], simulationMethod=2),
                                      dict(size=[0.08, 0.05, 0.05])), [1, 2, 10, 50]),
    # BASELINE config C3: DP + CSPM + RK4 (shrunken)
    "dp2d_small_rk4_cspm": (lambda: _scene("test2_cc_sand.json",
                                           dict(domainEnd=[0.2, 0.08, 0.05], simulationMethod=3, kernelCorrection=1,
                                                timeIntegration=4),
                                           dict(size=[0.08, 0.05, 0.05])), [1, 2, 10, 30]),
    # shipped test2 parameters (DP, LF, xsph, no CSPM), shrunken
    "dp2d_small_lf": (lambda: _scene("test2_cc_sand.json", dict(domainEnd=[0.2, 0.08, 0.05]),
                                     dict(size=[0.08, 0.05, 0.05])), [1, 2, 10]),
    # full-size C2 / C3, first steps only
    "c2_test2_mui_lf": (lambda: _scene("test2_cc_sand.json", dict(simulationMethod=2)), [1, 2]),
    "c3_test2_dp_rk4_cspm": (lambda: _scene("test2_cc_sand.json",
                                            dict(simulationMethod=3, kernelCorrection=1, timeIntegration=4)), [1]),
    # static rigid indenter with a prescribed velocity pressed into a DP soil bed (the shipped test4
    # scene shrunken: bed 50 x 20 particles, indenter 8 x 4, v = (0, -0.25, 0), isDynamic 0), "LF", XSPH
    "dp2d_indenter_lf": (lambda: _indenter(), [1, 2, 10, 30]),
    # the same indenter pushed sideways through a water bed under WCSPH (type 11 in the wall loop of wc:90-106, 125-126)
    "wc2d_indenter_lf": (lambda: _indenter(water=True), [1, 2, 10]),
    # the shipped test5 scene shrunken: four soil blocks (an L-shaped bed with a gap) and a static rigid plate pushed
    # sideways at 0.64 m/s (oracle-only fixture: several objects of one material + a horizontal indenter)
    "dp2d_plate_lf": (lambda: _plate(), [1, 2, 10]),
    # tiny 3D dambreak with the C4 parameter set
    "wc3d_tiny_lf": (lambda: _scene("test1_db_water.json", dict(WATER3D), dict(size=[0.16, 0.12, 0.12])), [1, 2, 3]),
    # ---- round 2, SURVEY 8 f4: the cubic-spline kernel through the soil sweeps (kernel = 0)
    "dp2d_small_lf_cubic": (lambda: _scene("test2_cc_sand.json", dict(domainEnd=[0.2, 0.08, 0.05], kernel=0),
                                           dict(size=[0.08, 0.05, 0.05])), [1, 2, 10]),
    "mui2d_small_lf_cubic": (lambda: _scene("test2_cc_sand.json",
                                            dict(domainEnd=[0.2, 0.08, 0.05], simulationMethod=2, kernel=0),
                                            dict(size=[0.08, 0.05, 0.05])), [1, 2, 10]),
    # ---- round 2, SURVEY 8 f2: DYNAMIC rigid body (shape matching + collision clamp) under the two soil solvers
    "dp2d_dynrigid_lf": (lambda: _dynrigid(3), [1, 2, 10, 30]),
    "mui2d_dynrigid_lf": (lambda: _dynrigid(2), [1, 2, 10]),
    "dp2d_dynrigid_wall_lf": (lambda: _dynrigid(3, wall=True), [1, 2, 10]),
    # ---- round 2, SURVEY 8 f3: the other boundary modes (1 enforced collision, 3 repulsive particles, 4 dummy + repulsive)
    "wc2d_rep_lf": (lambda: _scene("test1_db_water.json", dict(domainEnd=[1.0, 0.6, 0.5], boundary=3),
                                   dict(size=[0.4, 0.3, 0.1])), [1, 2, 10]),
    "wc2d_dummyrep_lf": (lambda: _scene("test1_db_water.json", dict(domainEnd=[1.0, 0.6, 0.5], boundary=4),
                                        dict(size=[0.4, 0.3, 0.1])), [1, 2, 10]),
    "wc2d_collision_lf": (lambda: _scene("test1_db_water.json", dict(domainEnd=[1.0, 0.6, 0.5], boundary=1),
                                         dict(size=[0.4, 0.3, 0.1])), [1, 2, 10, 40]),
    "mui2d_dummyrep_lf": (lambda: _scene("test2_cc_sand.json",
                                         dict(domainEnd=[0.2, 0.08, 0.05], simulationMethod=2, boundary=4),
                                         dict(size=[0.08, 0.05, 0.05])), [1, 2, 10]),
    # ---- round 2: the BASELINE configs over their full horizons (100 steps for C1-C3) ----
    "c1_test1_wc_lf_h100": (lambda: _scene("test1_db_water.json"), [1, 10, 50, 100]),
    "c2_test2_mui_lf_h100": (lambda: _scene("test2_cc_sand.json", dict(simulationMethod=2)), [1, 10, 50, 100]),
    "c3_test2_dp_rk4_cspm_h30": (lambda: _scene("test2_cc_sand.json",
                                                dict(simulationMethod=3, kernelCorrection=1, timeIntegration=4)),
                                 [1, 10, 30]),
    # 3D dambreak with the C4 parameter set, N = 20 772 (fluid 24 x 16 x 20)
    "wc3d_20k_lf": (lambda: _scene("test1_db_water.json",
                                   dict(WATER3D, particleRadius=0.005, domainEnd=[0.4, 0.24, 0.2]),
                                   dict(size=[0.24, 0.16, 0.2])), [1, 10, 20]),
}

SCALARS = ["mat_type", "id0", "obj_id", "grid_ids", "m_V", "density", "mass", "pressure", "CSPM_f", "d_density",
           "flag_retmap", "strain_equ", "d_strain_equ", "strain_equ_p", "d_strain_equ_p", "density_tmp"]
VECTORS = ["x", "v", "d_vel", "v_tmp"]
MATRICES = ["stress", "CSPM_L", "d_stress", "v_grad", "stress_tmp"]


def _col(pt, name, n):
    col = pt.cols[name][:n]
    if name in SCALARS:
        return np.array(col)
    return np.array([m.d for m in col], dtype=np.float64)


def run_case(name):
    sys.path.insert(0, os.path.join(HERE, "ti_shim"))
    sys.path.insert(0, REF)
    import taichi as ti                                   # the shim
    from eng.simulation import Simulation, SimConfiger    # the reference, unmodified

    factory, steps = CASES[name]
    scene = factory()
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(scene, f)
    cfg = SimConfiger(f.name)
    case = Simulation(cfg)
    os.unlink(f.name)
    ps, solver = case.ps, case.solver
    n = ps.particle_num[None]
    out = {}
    state = {"step": 0}

    def count_task(i, j, ret):
        ret.add(1)

    orig_init_real2tmp = solver.init_real2tmp

    def hooked_init_real2tmp():
        s = state["step"] + 1
        if s in steps:
            pre = f"s{s}/grid/"
            for nm in ("grid_ids", "id0", "CSPM_f", "CSPM_L", "m_V", "x"):
                out[pre + nm] = _col(ps.pt, nm, n)
            out[pre + "grid_particle_num"] = np.array(ps.grid_particle_num.d, dtype=np.int64)
            cnt = np.zeros(n, dtype=np.int64)
            for i in range(n):
                box = ti._Box(0)
                ps.for_all_neighbors(i, count_task, box)
                cnt[i] = box.v
            out[pre + "neighbor_count"] = cnt
        orig_init_real2tmp()

    solver.init_real2tmp = hooked_init_real2tmp

    t0 = time.time()
    for s in range(1, max(steps) + 1):
        solver.step()
        state["step"] = s
        if s in steps:
            pre = f"s{s}/end/"
            for nm in SCALARS + VECTORS + MATRICES:
                out[pre + nm] = _col(ps.pt, nm, n)
            pos, data = ps.dump()                    # the reference's own export API (ps:459-545)
            out[pre + "dump_keys"] = np.array(sorted(list(pos) + list(data)))
            assert np.array_equal(data["id0"], out[pre + "id0"])
        print(f"[{name}] step {s}/{max(steps)}  {time.time() - t0:.1f}s", flush=True)

    meta = dict(case=name, scene=scene, steps=steps, n=int(n), dt=float(solver.dt[None]),
                dim=int(ps.dim), grid_num=[int(v) for v in ps.grid_num], grid_size=float(ps.grid_size),
                vdomain_start=[float(v) for v in ps.vdomain_start], support_radius=float(ps.support_radius),
                smoothing_len=float(ps.smoothing_len), m_V0=float(ps.m_V0),
                shim_oob_reads=int(ti.OOB_READS[0]), semantics="serial (ti_shim)",
                solver=type(solver).__name__)
    for k in ("alpha_fric", "k_c", "G", "K", "vsound", "mu"):
        if hasattr(solver, k):
            meta[k] = float(getattr(solver, k))
    out["meta"] = np.array(json.dumps(meta))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"[{name}] wrote {name}.npz  ({time.time() - t0:.1f}s)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--case", action="append", default=[])
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--jobs", type=int, default=8)
    a = ap.parse_args()
    if a.list:
        print("\n".join(CASES))
        return
    names = list(CASES) if a.all else a.case
    if len(names) == 1:
        run_case(names[0])
        return
    import subprocess
    procs = []
    for nm in names:
        while len([p for p in procs if p.poll() is None]) >= a.jobs:
            time.sleep(1.0)
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--case", nm]))
    rc = [p.wait() for p in procs]
    sys.exit(max(rc) if rc else 0)


if __name__ == "__main__":
    main()
