"""GPU parity tests: the CUDA engine (through the reference-shaped Python API and the C ABI) against
  (a) the golden fixtures produced by executing the reference's own sources (oracle/gen_golden.py), and
  (b) the CPU oracle (oracle/sph_oracle.c) on the same inputs.
Integer work (cell ids, sorted order, cell offsets, neighbour counts) must be bit-exact; floating-point fields in
float64 mode agree to summation-order noise; mixed float32 mode within 1e-5 relative after one step.
"""
import numpy as np
import pytest

from helpers import Golden, relmax, sym6_to_9, make_sim, engine_fields

pytestmark = pytest.mark.gpu

# the *_indenter_* fixtures hold a static rigid block (type 11) with a prescribed velocity
# the *_rep_* / *_dummyrep_* / *_collision_* fixtures run the other boundary modes (3, 4, 1: SURVEY 8 f3)
WC_CASES = ["wc2d_small_lf", "wc2d_small_se_cubic", "wc2d_small_rk4_cspm", "wc3d_tiny_lf", "c1_test1_wc_lf", "wc2d_indenter_lf",
            "wc2d_rep_lf", "wc2d_dummyrep_lf", "wc2d_collision_lf"]
SOIL_CASES = ["mui2d_small_lf", "dp2d_small_rk4_cspm", "dp2d_small_lf", "c2_test2_mui_lf", "c3_test2_dp_rk4_cspm",
              "dp2d_indenter_lf", "mui2d_dummyrep_lf", "dp2d_small_lf_cubic", "mui2d_small_lf_cubic",
              # DYNAMIC rigid body: reaction terms, shape matching, collision clamp
              "mui2d_dynrigid_lf", "dp2d_dynrigid_wall_lf", "dp2d_dynrigid_lf"]
ALL_CASES = WC_CASES + SOIL_CASES

F64_TOL = 1e-9
# strain_equ_p integrates lambda*g_p/d_stress[a][b] (dp:111-118, 203-206): ill-conditioned by the reference's formula
FIELD_TOL = {"strain_equ_p": 5e-2, "d_strain_equ_p": 5e-2}
# fields the serial reference computes with in-place reads that a parallel engine evaluates on a snapshot:
# XSPH positions (base:231-238) and the mu(I) regularised stress (muI:151-156, output-only field, H25)
RACY = {"mui": {"stress"}, "xsph": {"x"}}
STATE = ["x", "v", "density", "m_V", "pressure", "d_density", "d_vel", "density_tmp", "v_tmp", "CSPM_f"]
SOIL_STATE = ["stress", "d_stress", "v_grad", "strain_equ", "strain_equ_p", "stress_tmp"]


def _grid_checks(sim, g, s, exact_counts=True):
    ps = sim.ps
    ps.initialize_particle_system()
    sim.solver.calc_kernel_corr()
    assert np.array_equal(ps.pt.grid_ids.cpu().numpy(), g.grid(s, "grid_ids")), f"cell ids differ at step {s}"
    assert np.array_equal(ps.pt.id0.cpu().numpy(), g.grid(s, "id0")), f"sorted order differs at step {s}"
    assert np.array_equal(ps.grid_particle_num.cpu().numpy(), g.grid(s, "grid_particle_num")), f"cell offsets differ at step {s}"
    if exact_counts:
        assert np.array_equal(ps.neighbor_count().cpu().numpy(), g.grid(s, "neighbor_count")), f"neighbour counts differ at step {s}"
    return ps


@pytest.mark.parametrize("name", ALL_CASES)
def test_f64_matches_reference_fixtures(name):
    """float64 engine vs the reference run: bit-exact integers at every snapshot, floats to 1e-9."""
    g = Golden(name)
    sim = make_sim(g.scene, precision="f64")
    assert sim.ps.particle_num[None] == g.meta["n"]
    assert sim.solver.dt[None] == g.meta["dt"]
    cfg = g.scene["Configuration"]
    racy = set()
    if cfg["simulationMethod"] == 2:
        racy |= RACY["mui"]
    if cfg["xsph"]:
        racy |= RACY["xsph"]
    last = max(g.steps) if (("small_lf" in name or "indenter" in name or "rep" in name or "collision" in name) and not cfg["xsph"]) or "tiny" in name or "dynrigid" in name \
        else min(max(g.steps), 10)
    # with XSPH the serial reference moves particles in place: positions drift from the snapshot evaluation by
    # O(dt * |v| * 1e-4) per step, so only the first snapshots are compared tightly
    if cfg["xsph"]:
        last = 1
    for s in range(1, last + 1):
        if s in g.steps:
            # (a shape-matched rigid body is its rest lattice rotated by an SVD: its lattice pairs sit exactly on the
            # support sphere, so their COUNT -- never their contribution -- depends on the SVD's last bit after step 1)
            ps = _grid_checks(sim, g, s, exact_counts="dynrigid" not in name or s == 1)
            assert relmax(ps.pt.CSPM_f.cpu().numpy(), g.grid(s, "CSPM_f")) < F64_TOL
            if cfg["kernelCorrection"] == 1:
                assert relmax(ps.pt.CSPM_L.cpu().numpy().reshape(-1, 9), g.grid(s, "CSPM_L")) < F64_TOL
        sim.solver.step()
        if s in g.steps:
            got = engine_fields(sim)
            assert np.array_equal(got["id0"], g.end(s, "id0"))
            if cfg["simulationMethod"] == 3:
                assert np.array_equal(got["flag_retmap"], g.end(s, "flag_retmap"))
            fields = STATE + (SOIL_STATE if cfg["simulationMethod"] != 1 else [])
            for f in fields:
                if f in racy:
                    continue
                err = relmax(got[f], g.end(s, f))
                assert err < FIELD_TOL.get(f, F64_TOL), f"{name}: step {s} field {f}: rel err {err:.3e}"
    assert sim.ps.engine.L.sph_read_bad_cells(sim.ps.engine.h) == 0


@pytest.mark.parametrize("name", ["mui2d_small_lf", "dp2d_small_lf", "dp2d_small_rk4_cspm", "wc2d_small_lf",
                                  "mui2d_dynrigid_lf", "dp2d_dynrigid_wall_lf", "dp2d_dynrigid_lf", "mui2d_dummyrep_lf"])
def test_f64_matches_oracle_jacobi_long(name):
    """float64 engine vs the oracle in snapshot ('jacobi') mode over the whole horizon of the fixture, all fields."""
    from oracle import oracle as orc
    g = Golden(name)
    sim = make_sim(g.scene, precision="f64")
    o = orc.Oracle.from_scene(g.scene, serial=0)
    nsteps = min(max(g.steps), 30)
    for s in range(1, nsteps + 1):
        sim.solver.step()
        assert o.step() == 0
        if s in (1, 2, 5, 10, nsteps):
            got = engine_fields(sim)
            assert np.array_equal(got["id0"], o.id0)
            assert np.array_equal(got["grid_ids"], o.grid_ids)
            fields = STATE + (SOIL_STATE if g.scene["Configuration"]["simulationMethod"] != 1 else [])
            for f in fields:
                err = relmax(got[f], getattr(o, f))
                # errors grow with step count through threshold branches (DP) - stated tolerance 1e-7 at 30 steps
                tol = FIELD_TOL.get(f, 1e-9 if s <= 2 else 1e-7)
                assert err < tol, f"{name}: step {s} field {f}: rel err {err:.3e}"


@pytest.mark.parametrize("name", ["wc2d_small_lf", "wc3d_tiny_lf", "wc2d_small_rk4_cspm", "dp2d_small_rk4_cspm", "mui2d_small_lf"])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_native_step_loop_equals_python_orchestration(name, prec):
    """sph_step (whole step enqueued natively: fused pointwise stages -- init_real2tmp in the reorder, the RK4 stage kernels
    in one pass -- and the per-step neighbour lists) must be bit-identical to SPHBase.step() driven from Python."""
    g = Golden(name)
    a = make_sim(g.scene, precision=prec)
    b = make_sim(g.scene, precision=prec)
    for _ in range(3):
        a.solver.step()
    b.solver.run_steps(3)
    fa, fb = engine_fields(a), engine_fields(b)
    keys = ["id0", "x", "v", "density", "pressure", "d_vel"] + (["stress", "d_stress", "strain_equ"] if "stress" in fa else [])
    for k in keys:
        assert np.array_equal(fa[k], fb[k]), k


@pytest.mark.parametrize("name", ["wc2d_small_lf", "wc3d_tiny_lf"])
def test_native_step_loop_equals_python_orchestration_mixed(name):
    """Cell-tile path: sph_step forms the Shepard sums inside the first wall / fluid pass, the Python-orchestrated step
    calls the stand-alone kernel -- the two must still agree bit for bit (the multi-GPU parity test relies on it)."""
    g = Golden(name)
    a = make_sim(g.scene, precision="f32")
    b = make_sim(g.scene, precision="f32")
    for _ in range(3):
        a.solver.step()
    b.solver.run_steps(3)
    fa, fb = engine_fields(a), engine_fields(b)
    for k in ("id0", "x", "v", "density", "pressure", "d_vel", "CSPM_f"):
        assert np.array_equal(fa[k], fb[k]), k


# -------------------------------------------------------------------------------------------- mixed precision
MIXED_TOL_1 = 1e-5     # north_star: density, pressure, acceleration, stress within 1e-5 relative after one step


def _well_conditioned(f_mine, f_ref):
    """Wall particles whose ONLY flow neighbours sit exactly on the support sphere of the initial lattice have a Shepard
    sum of ~1e-64 (float64) or ~1e-27 (float32) or exactly 0, depending on how |x_ij| < support rounds.
    The reference then extrapolates p = (sum V p w) / (sum V w) as a ratio of two such numbers.  Whether the pair is
    inside is rounding noise in either precision, so those particles (CSPM_f > 1e3, normal values are 1..10) are
    excluded from float32-vs-float64 comparisons; their kernel gradients towards the fluid are ~0, so they do not
    influence any flow particle."""
    return (np.abs(f_mine) < 1e3) & (np.abs(f_ref) < 1e3)


@pytest.mark.parametrize("name", ALL_CASES)
def test_mixed_one_step_within_1e5(name):
    g = Golden(name)
    sim = make_sim(g.scene, precision="f32")
    cfg = g.scene["Configuration"]
    _grid_checks(sim, g, 1, exact_counts=False)      # cell ids / order / offsets are float64 work in both modes
    sim.solver.step()
    got = engine_fields(sim)
    assert np.array_equal(got["id0"], g.end(1, "id0"))
    fields = ["density", "pressure", "d_vel", "v", "d_density", "x"]
    if cfg["simulationMethod"] == 3:
        fields += ["stress", "d_stress"]
    if cfg["simulationMethod"] == 2:
        fields += ["stress_tmp"]
    ok = _well_conditioned(got["CSPM_f"], g.end(1, "CSPM_f"))
    assert ok.mean() > 0.9
    for f in fields:
        a, b = got[f][ok], g.end(1, f)[ok]
        if f == "d_stress":
            # a stress RATE: the plastic-multiplier branch (f >= -eps_f and sqrt(J2) > eps, dp:194) flips for a few
            # particles between precisions, so the bound is on all but 0.2 % of the particles (at least 3: the small scenes hold ~1400)
            scale = np.max(np.abs(b))
            bad = np.max(np.abs(a - b), axis=1) > MIXED_TOL_1 * scale
            assert bad.sum() <= max(3, 2e-3 * len(bad)), f"{name}: d_stress differs for {bad.sum()} particles"
            continue
        err = relmax(a, b)
        assert err < MIXED_TOL_1, f"{name}: field {f}: rel err {err:.3e}"


def test_mixed_neighbor_counts_bit_exact_vs_oracle():
    """The float32 predicate (cell-local coordinates) restated in the oracle gives identical counts."""
    from oracle import oracle as orc
    for name in ("c1_test1_wc_lf", "wc3d_tiny_lf"):
        g = Golden(name)
        sim = make_sim(g.scene, precision="f32")
        o = orc.Oracle.from_scene(g.scene, serial=0)
        for s in range(3):
            sim.ps.initialize_particle_system()
            o.grid_build()
            assert np.array_equal(sim.ps.neighbor_count().cpu().numpy(), o.neighbor_count(f32=True)), (name, s)
            sim.solver.step()
            o.x[:] = sim.ps.pt.x.cpu().numpy()      # same positions on both sides


def test_mixed_100_step_horizon_wc():
    """Stated tolerance over a 100-step horizon (2D dambreak, 'LF'): rho 1e-6, v 1e-3, p 2e-2 (max-norm relative)
    against the reference fixture.  Pressure is the Tait EOS of a density known to ~1e-7, amplified by gamma*k/p."""
    g = Golden("wc2d_small_lf")
    sim = make_sim(g.scene, precision="f32")
    sim.solver.run_steps(100)
    got = engine_fields(sim)
    assert np.array_equal(got["id0"], g.end(100, "id0"))
    assert relmax(got["density"], g.end(100, "density")) < 1e-6
    assert relmax(got["x"], g.end(100, "x")) < 1e-6
    assert relmax(got["v"], g.end(100, "v")) < 1e-3
    assert relmax(got["pressure"], g.end(100, "pressure")) < 2e-2


def test_f64_100_step_horizon_wc():
    g = Golden("wc2d_small_lf")
    sim = make_sim(g.scene, precision="f64")
    sim.solver.run_steps(100)
    got = engine_fields(sim)
    assert np.array_equal(got["id0"], g.end(100, "id0"))
    for f in ("density", "x", "v", "pressure", "d_vel"):
        assert relmax(got[f], g.end(100, f)) < 1e-8, f


# -------------------------------------------------------------------------------------------- properties at size
def _box_scene(n_side, is2d=False):
    d = 0.01
    L = n_side * d
    return {
        "Configuration": dict(is2D=is2d, domainStart=[0.0, 0.0, 0.0], domainEnd=[2 * L, 1.5 * L, L if not is2d else 0.5],
                              gravitation=[0.0, -9.81, 0.0], particleRadius=d / 2, kappa=2.0, kh=1.5, simulationMethod=1,
                              timeStepSizeMin=1e-6, boundary=2, kernel=1, kernelCorrection=0, timeIntegration=2,
                              xsph=False, colorTitle=0, colorGroup=0, showBdyPts=True),
        "Materials": [dict(matId=0, matType=1, density0=1000.0, viscosity=0.01, stiffness=50000.0, exponent=7.0,
                           color=[50, 100, 200])],
        "Blocks": [dict(objectId=0, materialId=0, translation=[0.0, 0.0, 0.0], size=[L, L, L if not is2d else 0.5],
                        velocity=[0.0, 0.0, 0.0], rotationAxis=[0, 0, 1], rotationAngle=0)],
    }


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_large_3d_sort_properties(prec):
    """~1M fluid particles: sortedness, permutation, histogram/offset consistency, idempotence of the grid build,
    symmetry of the neighbour relation (sum of counts is even) and agreement of the two precisions' step."""
    import torch
    sim = make_sim(_box_scene(100), precision=prec)
    ps = sim.ps
    n = ps.particle_num[None]
    assert n > 1_000_000
    sim.solver.run_steps(2)
    ps.initialize_particle_system()
    gid = ps.pt.grid_ids
    assert bool((gid[1:] >= gid[:-1]).all()), "cell ids not sorted"
    id0 = ps.pt.id0.long()
    assert bool((torch.sort(id0).values == torch.arange(n, device=id0.device)).all()), "id0 is not a permutation"
    cell_end = ps.grid_particle_num.long()
    hist = torch.bincount(gid.long(), minlength=ps.grid_num_total)
    assert bool((torch.cumsum(hist, 0) == cell_end).all())
    assert bool((ps.grid_particle_num_temp.long() == hist).all())
    # stability: inside a cell the previous relative order is kept -> a second build is the identity
    before = id0.clone()
    ps.initialize_particle_system()
    assert bool((ps.pt.id0.long() == before).all())
    cnt = ps.neighbor_count().long()
    if prec == "f64":       # global float64 coordinates: the relation is exactly symmetric
        assert int(cnt.sum()) % 2 == 0
    assert int(cnt.max()) < 400
    assert sim.ps.engine.L.sph_read_bad_cells(sim.ps.engine.h) == 0
    assert bool(torch.isfinite(ps.pt.v).all()) and bool(torch.isfinite(ps.pt.density).all())


def test_density_sum_matches_oracle_small_3d():
    from oracle import oracle as orc
    g = Golden("wc3d_tiny_lf")
    for prec, tol in (("f64", 1e-12), ("f32", 2e-6)):
        sim = make_sim(g.scene, precision=prec)
        o = orc.Oracle.from_scene(g.scene, serial=0)
        sim.ps.initialize_particle_system()
        o.grid_build()
        got = sim.ps.density_sum().cpu().numpy()
        assert relmax(got, o.density_sum()) < tol, prec


# -------------------------------------------------------------------------------------------- cell-tile fast path
@pytest.mark.parametrize("name", ["wc2d_small_lf", "wc3d_tiny_lf", "c1_test1_wc_lf", "wc2d_small_se_cubic"])
def test_tile_path_equals_generic_path(name):
    """MIXED engine: cell-tile kernels (TMA-staged tiles + neighbour bit masks) vs the generic per-particle sweeps.
    Same pairs, same order, same arithmetic apart from r*r vs r2 in one denominator: agreement to float32 rounding."""
    g = Golden(name)
    a = make_sim(g.scene, precision="f32", fastSweeps=True)
    b = make_sim(g.scene, precision="f32", fastSweeps=False)
    assert a.ps.engine.params.fast >= 1 and b.ps.engine.params.fast == 0
    a.ps.initialize_particle_system()
    b.ps.initialize_particle_system()
    a.solver.calc_kernel_corr()
    b.solver.calc_kernel_corr()
    fa, fb = a.ps.pt.CSPM_f.cpu().numpy(), b.ps.pt.CSPM_f.cpu().numpy()
    ok = _well_conditioned(fa, fb)
    assert relmax(fa[ok], fb[ok]) < 2e-6
    assert np.array_equal(a.ps.neighbor_count().cpu().numpy(), b.ps.neighbor_count().cpu().numpy())
    for s in range(5):
        a.solver.step()
        b.solver.step()
    fa, fb = engine_fields(a), engine_fields(b)
    assert np.array_equal(fa["id0"], fb["id0"])
    ok = _well_conditioned(fa["CSPM_f"], fb["CSPM_f"])
    for k in ("x", "v", "density", "pressure", "d_vel", "d_density", "v_tmp", "CSPM_f"):
        assert relmax(fa[k][ok], fb[k][ok]) < 1e-5, k


@pytest.mark.parametrize("name", ["wc2d_small_lf", "wc3d_tiny_lf", "c1_test1_wc_lf", "wc2d_small_rk4"])
def test_neighbour_round_lists_replay_bit_exact(name):
    """fast = 2 records the neighbours the first fluid pass of a step finds and replays them in the later one_steps
    (LF: 1, RK4: 3): same pairs, same order, same arithmetic => bit-identical to walking the masks again (fast = 1)."""
    import copy
    if name == "wc2d_small_rk4":
        scene = copy.deepcopy(Golden("wc2d_small_lf").scene)
        scene["Configuration"]["timeIntegration"] = 4
    else:
        scene = Golden(name).scene
    a = make_sim(scene, precision="f32", neighbourLists=True)
    b = make_sim(scene, precision="f32", neighbourLists=False)
    assert a.ps.engine.params.fast == 2 and b.ps.engine.params.fast == 1
    for s in range(6):
        a.solver.step()
        b.solver.step()
    fa, fb = engine_fields(a), engine_fields(b)
    for k in ("id0", "x", "v", "density", "pressure", "d_vel", "d_density", "v_tmp", "CSPM_f"):
        assert np.array_equal(fa[k], fb[k]), k


@pytest.mark.parametrize("name", ["wc2d_small_lf", "wc3d_tiny_lf", "c1_test1_wc_lf"])
def test_tile_masks_reproduce_the_neighbour_predicate(name):
    """The per-step neighbour bit masks (one evaluation per cell pair + warp transpose) must select exactly the pairs
    of the float32 predicate: popcounts of every flow particle == generic neighbour count == the oracle's restatement."""
    import torch
    from oracle import oracle as orc
    g = Golden(name)
    sim = make_sim(g.scene, precision="f32")
    o = orc.Oracle.from_scene(g.scene, serial=0)
    eng = sim.ps.engine
    for s in range(4):
        sim.ps.initialize_particle_system()
        sim.solver.calc_kernel_corr()
        out = torch.empty(eng.n, dtype=torch.int32, device=eng.device)
        eng.call("sph_neighbor_count_masks", out.data_ptr())
        got = out.cpu().numpy()
        want = sim.ps.neighbor_count().cpu().numpy()
        flow = sim.ps.pt.mat_type.cpu().numpy() > 0
        assert (got[flow] >= 0).all(), "a flow particle was left to the generic path on a lattice scene"
        assert np.array_equal(got[flow], want[flow]), (name, s)
        o.x[:] = sim.ps.pt.x.cpu().numpy()
        o.grid_build()
        assert np.array_equal(want, o.neighbor_count(f32=True)), (name, s)
        sim.solver.step()
        o.x[:] = sim.ps.pt.x.cpu().numpy()


def test_tile_path_crowded_cells_fall_back():
    """kh = 2 puts 4^3 = 64 particles in a cell (> 32): every cell is flagged and the generic kernels must take over."""
    import copy
    from oracle import oracle as orc
    g = Golden("wc3d_tiny_lf")
    scene = copy.deepcopy(g.scene)
    scene["Configuration"]["kh"] = 2.0
    sim = make_sim(scene, precision="f32")
    o = orc.Oracle.from_scene(scene, serial=0)
    sim.solver.step()
    assert o.step() == 0
    got = engine_fields(sim)
    assert np.array_equal(got["id0"], o.id0)
    ok = _well_conditioned(got["CSPM_f"], o.CSPM_f)
    for k in ("density", "pressure", "d_vel", "v"):
        assert relmax(got[k][ok], getattr(o, k)[ok]) < 1e-5, k


def test_tile_path_partially_flagged_cells():
    """Two fluid blocks that overlap (second lattice shifted by half a spacing) put up to 54 particles into the cells
    of the overlap: those cells and their stencil neighbours fall back to the generic kernels while the rest of the
    scene stays on the tile path.  The mixed step must agree with the all-generic engine to float32 rounding.
    (One and two steps only: the overlap is a pressure bomb, not a flow.)"""
    import copy
    import torch
    from tisphi_b200 import scenes
    scene = scenes.dambreak3d(scale=0.5, precision="f32")
    scene["Configuration"]["domainEnd"] = [1.0, 0.6, 0.4]
    scene["Blocks"][0].update(size=[0.4, 0.3, 0.4])
    blk = copy.deepcopy(scene["Blocks"][0])
    blk.update(objectId=1, translation=[0.305, 0.005, 0.005], size=[0.2, 0.2, 0.2])
    scene["Blocks"].append(blk)
    a = make_sim(scene, precision="f32", fastSweeps=True)
    b = make_sim(scene, precision="f32", fastSweeps=False)
    eng = a.ps.engine
    a.ps.initialize_particle_system()
    a.solver.calc_kernel_corr()
    out = torch.empty(eng.n, dtype=torch.int32, device=eng.device)
    eng.call("sph_neighbor_count_masks", out.data_ptr())
    flow = a.ps.pt.mat_type > 0
    left_out, n_flow = int(((out < 0) & flow).sum()), int(flow.sum())
    assert 0 < left_out < n_flow, (left_out, n_flow)            # some cells flagged, most not
    on_tile = (out >= 0) & flow
    assert torch.equal(out[on_tile], a.ps.neighbor_count()[on_tile])
    assert int(torch.bincount(a.ps.pt.grid_ids.long()).max()) > 32
    for s, tol in enumerate((1e-5, 1e-4)):
        a.solver.step()
        b.solver.step()
        fa, fb = engine_fields(a), engine_fields(b)
        # particles are compared by identity: under this violent start a rounding difference can move a particle
        # across a cell border, which changes the sorted order but not the physics
        ia, ib = np.argsort(fa["id0"]), np.argsort(fb["id0"])
        ok = _well_conditioned(fa["CSPM_f"][ia], fb["CSPM_f"][ib])
        # the derivatives of the second step sit on top of the explosion of the first: only the state is compared there
        keys = ("x", "v", "density", "pressure", "d_vel", "d_density", "CSPM_f") if s == 0 else ("x", "v", "density", "pressure")
        errs = {k: relmax(fa[k][ia][ok], fb[k][ib][ok]) for k in keys}
        assert all(e < tol for e in errs.values()), (s, errs)


@pytest.mark.parametrize("name", ["c1_test1_wc_lf_h100", "wc3d_20k_lf", "c2_test2_mui_lf_h100"])
def test_mixed_neighbor_counts_bracketed_by_reference_positions(name):
    """Pins the float32 (cell-local) neighbour predicate to the REFERENCE run, not to the oracle's restatement of it:
    the fixture's float64 positions of every snapshot are injected, the MIXED engine sorts and counts, and every count
    must lie between #{r < support (1 - tol)} and #{r < support (1 + tol)} evaluated in float64 on those positions (KD
    tree), tol = 1e-6 -- i.e. the MIXED neighbour set differs from the reference's only by pairs within 1e-6 (relative)
    of the support sphere, of which a rest lattice has many.  The reference's own counts satisfy the same
    bracket, and wherever the bracket is tight (no borderline pair) the MIXED count equals the reference's exactly."""
    import torch
    from scipy.spatial import cKDTree
    g = Golden(name)
    tol = 1e-6
    for s in g.steps:
        sim = make_sim(g.scene, precision="f32")
        ps = sim.ps
        sup = float(ps.support_radius)
        xs, id0 = g.grid(s, "x"), g.grid(s, "id0")
        x_creation = np.empty_like(xs)
        x_creation[id0] = xs
        ps.pt.x.copy_(torch.from_numpy(x_creation).to(ps.pt.x.device))          # before the first sort: creation order
        ps.initialize_particle_system()
        mine_sorted = ps.neighbor_count().cpu().numpy()
        mine = np.empty_like(mine_sorted)
        mine[ps.pt.id0.cpu().numpy()] = mine_sorted                              # by creation index
        ref = np.empty_like(mine)
        ref[id0] = g.grid(s, "neighbor_count")
        tree = cKDTree(x_creation)
        lo = tree.query_ball_point(x_creation, sup * (1 - tol), return_length=True) - 1
        hi = tree.query_ball_point(x_creation, sup * (1 + tol), return_length=True) - 1
        assert np.all((lo <= ref) & (ref <= hi)), (name, s, "the reference's own counts leave the bracket")
        assert np.all((lo <= mine) & (mine <= hi)), (name, s, int(((mine < lo) | (mine > hi)).sum()))
        tight = lo == hi
        assert np.array_equal(mine[tight], ref[tight]), (name, s)
        if sim.solver_type == 1:                                                 # the cell-tile path's bit masks (flow particles)
            out = torch.empty(ps.engine.n, dtype=torch.int32, device=ps.engine.device)
            sim.solver.calc_kernel_corr()                                        # builds the masks
            ps.engine.call("sph_neighbor_count_masks", out.data_ptr())
            m_sorted = out.cpu().numpy()
            assert np.array_equal(m_sorted[m_sorted >= 0], mine_sorted[m_sorted >= 0]), (name, s, "mask counts differ from the walk")
            assert (m_sorted >= 0).sum() > 0
