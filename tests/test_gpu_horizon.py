"""Long-horizon parity on the BASELINE configurations themselves (100 steps for C1-C3).

Fixtures `*_h100` / `*_h30` / `wc3d_20k_lf` were produced by executing the reference's own sources under the serial
Taichi emulator for the whole horizon (oracle/gen_golden.py; about an hour of CPU each).  At every snapshot the integer
work (cell ids, sorted order, cell offsets, float64 neighbour counts) must be bit-exact; floats are held to the stated
tolerances below, BOTH in the max-norm (relmax) and element-wise with an absolute floor of 1e-3 of the field's scale
(relelem).  Every tolerance and exclusion of this file is restated in DESIGN.md section 2.

Exclusions (and why):
  * strain_equ_p / d_strain_equ_p: the reference integrates lambda * g_p / d_stress[a][b] over the components with
    |d_stress| > 1e-8 (dp:111-118, 203-206) -- a division by near-zero stress-rate components; the reference's OWN
    float64 result is not reproducible to better than 5e-2 by a float64 restatement (tests/test_oracle_golden.py).
  * with XSPH (C2, C3) the serial reference moves particles IN PLACE while neighbours still read them (base:231-238);
    the engine evaluates XSPH on a snapshot.  The difference is part of the tolerance (it is what bounds C3's 1e-6).
  * mu(I) `stress` is the Shepard-regularised OUTPUT field (muI:151-156, in place = Gauss-Seidel in the serial
    reference, snapshot here); it never feeds the dynamics and differs at the 1e-1 level between the two
    evaluation orders of the reference itself.  The dynamic state (x, v, density) and stress_tmp are compared.
  * mixed precision: wall particles whose only flow neighbours sit exactly on the support sphere (CSPM_f > 1e3, see
    test_gpu_parity._well_conditioned) are masked out of pressure / velocity comparisons.
"""
import numpy as np
import pytest

from helpers import Golden, relmax, relelem, make_sim, engine_fields

pytestmark = pytest.mark.gpu

# (max-norm tolerance, element-wise tolerance) per field; about 10x what the B200 run measured (profiles/r2_horizon_parity.log)
F64_WC = {"x": (1e-12, 1e-10), "v": (1e-11, 1e-8), "density": (1e-13, 1e-13), "pressure": (1e-10, 1e-7), "d_vel": (1e-10, 1e-7),
          "d_density": (1e-10, 1e-7)}
MIXED_WC3D_20 = {"x": (1e-8, 1e-6), "density": (5e-6, 5e-6), "v": (2e-4, 5e-2), "pressure": (2e-4, 5e-2)}
MIXED_WC_100 = {"x": (1e-9, 1e-6), "density": (1e-8, 1e-8), "v": (1e-5, 1e-2), "pressure": (1e-5, 1e-2)}
# DP + CSPM + RK4 with XSPH: the float64 figures are the in-place vs snapshot XSPH of the reference (module docstring)
F64_DP_30 = {"x": (1e-9, 1e-6), "v": (1e-7, 1e-5), "density": (1e-12, 1e-12), "stress": (1e-7, 1e-6), "d_vel": (1e-7, 1e-5),
             "d_stress": (1e-5, 1e-3), "strain_equ": (1e-6, 1e-6)}
MIXED_DP_30 = {"x": (1e-9, 1e-6), "density": (1e-8, 1e-8), "v": (1e-4, 2e-2), "stress": (5e-4, 1e-2)}
F64_MUI_100 = {"x": (1e-6, 1e-4), "v": (1e-4, 1e-2), "density": (1e-8, 1e-8), "stress_tmp": (1e-3, 1e-1)}
MIXED_MUI_100 = {"x": (1e-5, 1e-4), "density": (1e-6, 1e-6), "v": (2e-2, 1.0), "stress_tmp": (5e-2, None)}


def _grid_snapshot(sim, g, s, f64):
    ps = sim.ps
    ps.initialize_particle_system()
    sim.solver.calc_kernel_corr()
    assert np.array_equal(ps.pt.grid_ids.cpu().numpy(), g.grid(s, "grid_ids")), f"cell ids differ at step {s}"
    assert np.array_equal(ps.pt.id0.cpu().numpy(), g.grid(s, "id0")), f"sorted order differs at step {s}"
    assert np.array_equal(ps.grid_particle_num.cpu().numpy(), g.grid(s, "grid_particle_num")), f"cell offsets differ at step {s}"
    if f64:
        assert np.array_equal(ps.neighbor_count().cpu().numpy(), g.grid(s, "neighbor_count")), f"neighbour counts differ at step {s}"


def _run_horizon(name, prec, tol, ok_mask=False, flags=None, report=None):
    g = Golden(name)
    sim = make_sim(g.scene, precision=prec)
    assert sim.ps.particle_num[None] == g.meta["n"] and sim.solver.dt[None] == g.meta["dt"]
    done = 0
    worst = {}
    for s in g.steps:
        sim.solver.run_steps(s - 1 - done)
        done = s - 1
        # integer work of step s: exact while the float64 state still determines it exactly (every snapshot in float64;
        # the first in mixed precision -- afterwards a particle within 1e-7 of a cell face may sit on the other side)
        if prec == "f64" or s == 1:
            _grid_snapshot(sim, g, s, prec == "f64")
        sim.solver.run_steps(1)
        done = s
        got = engine_fields(sim)
        same_order = np.array_equal(got["id0"], g.end(s, "id0"))
        assert same_order or prec != "f64", f"{name}: sorted order differs after step {s}"
        pa, pb = (slice(None), slice(None)) if same_order else (np.argsort(got["id0"]), np.argsort(g.end(s, "id0")))
        keep = np.ones(len(got["id0"]), dtype=bool)
        if ok_mask:
            keep = (np.abs(got["CSPM_f"][pa]) < 1e3) & (np.abs(g.end(s, "CSPM_f")[pb]) < 1e3)
            assert keep.mean() > 0.9
        if flags is not None:
            mism = int((got["flag_retmap"][pa] != g.end(s, "flag_retmap")[pb]).sum())
            assert mism <= flags * len(keep), f"{name}[{prec}] step {s}: {mism} return-mapping flags differ"
        for f, (tmax, telem) in tol.items():
            a, b = got[f][pa][keep], g.end(s, f)[pb][keep]
            emax, eelem = relmax(a, b), relelem(a, b)
            worst[f] = (max(worst.get(f, (0, 0))[0], emax), max(worst.get(f, (0, 0))[1], eelem))
            assert emax < tmax, f"{name}[{prec}] step {s} {f}: max-norm rel err {emax:.3e} (stated {tmax:g})"
            # (telem None: a field with a max(., 0) clamp -- mu(I) p = max(c^2 (rho - rho0), 0) -- whose entries switch between
            # 0 and a small value across precisions: only the max-norm is meaningful)
            assert telem is None or eelem < telem, f"{name}[{prec}] step {s} {f}: element-wise rel err {eelem:.3e} (stated {telem:g})"
    print(f"\nHORIZON {name}[{prec}] steps {g.steps}: " + " ".join(f"{f}={v[0]:.1e}/{v[1]:.1e}" for f, v in worst.items()))
    assert sim.ps.engine.L.sph_read_bad_cells(sim.ps.engine.h) == 0
    return sim, g


def test_c1_100_steps_f64():
    """BASELINE config C1 (test1_db_water.json as shipped), snapshots 1, 10, 50, 100: the 2D dambreak density field."""
    _run_horizon("c1_test1_wc_lf_h100", "f64", F64_WC)


def test_c1_100_steps_mixed():
    _run_horizon("c1_test1_wc_lf_h100", "f32", MIXED_WC_100, ok_mask=True)


def test_c3_30_steps_f64():
    """BASELINE config C3 (DP + CSPM + RK4 + XSPH on the test2 geometry), snapshots 1, 10, 30; flags bit-exact."""
    _run_horizon("c3_test2_dp_rk4_cspm_h30", "f64", F64_DP_30, flags=0.0)


def test_c3_30_steps_mixed():
    # threshold branches (f > 1e-4, |d_stress| > 1e-8, max(rho0, rho)) flip for single particles between precisions
    # (SURVEY H26): at most 0.5 % of the return-mapping flags may differ
    _run_horizon("c3_test2_dp_rk4_cspm_h30", "f32", MIXED_DP_30, flags=5e-3)


def test_c2_100_steps_f64():
    """BASELINE config C2 (mu(I) + "LF" + XSPH on the test2 geometry), snapshots 1, 10, 50, 100.  The float64 figures are
    the in-place vs snapshot XSPH of the reference (module docstring); `stress` (regularised output) is excluded."""
    _run_horizon("c2_test2_mui_lf_h100", "f64", F64_MUI_100)


def test_c2_100_steps_mixed():
    _run_horizon("c2_test2_mui_lf_h100", "f32", MIXED_MUI_100)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_wc3d_20k_20_steps(prec):
    """3D WCSPH dambreak with the C4 parameter set at 20 772 particles against the reference run, snapshots 1, 10, 20
    (the reference's 3D scheme diverges after ~35 steps)."""
    _run_horizon("wc3d_20k_lf", prec, F64_WC if prec == "f64" else MIXED_WC3D_20, ok_mask=prec != "f64")


def test_dump_matches_reference_keys():
    """ParticleSystem.dump (ps:459-545): the same keys as the reference's own dump() call recorded in the fixture."""
    g = Golden("c1_test1_wc_lf_h100")
    sim = make_sim(g.scene, precision="f64")
    sim.solver.run_steps(1)
    pos, data = sim.ps.dump()
    assert sorted(list(pos) + list(data)) == [str(k) for k in g.end(1, "dump_keys")]
    assert np.array_equal(np.asarray(data["id0"]), g.end(1, "id0"))
    assert relmax(np.asarray(data["density"]), g.end(1, "density")) < 1e-12


# ------------------------------------------------------------------------------------------ the benchmarked scene, coarsened
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_c4_coarse_vs_oracle(prec):
    """BASELINE config C4 (the benchmarked 3D dambreak) coarsened x8, against the float64 oracle at steps 1, 10 and 20
    (the reference's 3D scheme diverges after ~35 steps).  float64: bit-exact integers, floats to
    1e-10; mixed precision: the stated one-step 1e-5, 5e-5 at 10 steps, 2e-4 at 20 steps."""
    from oracle import oracle as orc
    from tisphi_b200 import scenes
    scene = scenes.dambreak3d(scale=0.125, precision=prec)
    sim = make_sim(scene, precision=prec)
    o = orc.Oracle.from_scene(scene, serial=0)
    assert sim.ps.particle_num[None] == o.n
    tol = {s: {"density": 1e-11, "pressure": 1e-10, "d_vel": 1e-10, "v": 1e-10, "x": 1e-12} for s in (1, 10, 20)} if prec == "f64" else \
          {1: {"density": 1e-8, "pressure": 1e-5, "d_vel": 1e-5, "v": 1e-5, "x": 1e-10},
           10: {"density": 1e-7, "pressure": 5e-5, "d_vel": 5e-5, "v": 5e-5, "x": 1e-9},
           20: {"density": 5e-6, "pressure": 2e-4, "d_vel": 2e-4, "v": 2e-4, "x": 1e-8}}
    done = 0
    for s in (1, 10, 20):
        sim.solver.run_steps(s - done)
        for _ in range(s - done):
            assert o.step() == 0
        done = s
        got = engine_fields(sim)
        if prec == "f64":
            assert np.array_equal(got["id0"], o.id0) and np.array_equal(got["grid_ids"], o.grid_ids)
            pa = pb = slice(None)
        else:
            pa, pb = np.argsort(got["id0"]), np.argsort(o.id0)
        keep = (np.abs(got["CSPM_f"][pa]) < 1e3) & (np.abs(o.CSPM_f[pb]) < 1e3)
        msg = []
        for f, t in tol[s].items():
            e = relmax(got[f][pa][keep], getattr(o, f)[pb][keep])
            msg.append(f"{f}={e:.1e}")
            assert e < t, f"C4/8[{prec}] step {s} {f}: rel err {e:.3e} (stated {t:g})"
        print(f"\nHORIZON c4_coarse[{prec}] step {s}: " + " ".join(msg))
