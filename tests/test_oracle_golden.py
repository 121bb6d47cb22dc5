"""The CPU oracle (oracle/sph_oracle.c) against fixtures produced by running the reference's own sources
(serial Taichi emulation, oracle/gen_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from helpers import Golden, relmax
from oracle import oracle as orc

CASES = ["wc2d_small_lf", "wc2d_small_se_cubic", "wc2d_small_rk4_cspm", "mui2d_small_lf", "dp2d_small_rk4_cspm",
         "dp2d_small_lf", "wc3d_tiny_lf", "c1_test1_wc_lf", "c2_test2_mui_lf", "c3_test2_dp_rk4_cspm",
         # static rigid indenter with a prescribed velocity (type 11 in the wall loops, d_vel = 0)
         "dp2d_indenter_lf", "wc2d_indenter_lf",
         # shipped test5 shrunken: four soil blocks + a static rigid plate pushed sideways
         "dp2d_plate_lf",
         # the cubic-spline kernel through the soil sweeps
         "dp2d_small_lf_cubic", "mui2d_small_lf_cubic",
         # DYNAMIC rigid body: reaction of the momentum sums, shape matching, collision clamp
         "mui2d_dynrigid_lf", "dp2d_dynrigid_wall_lf", "dp2d_dynrigid_lf",
         # boundary modes 3 (repulsive particles), 4 (dummy + repulsive), 1 (enforced collision)
         "wc2d_rep_lf", "wc2d_dummyrep_lf", "wc2d_collision_lf", "mui2d_dummyrep_lf",
         # round 2: the BASELINE configs over their full horizons (C1, C2: 100 steps, C3: 30 steps; ~1 h of emulator each)
         "c1_test1_wc_lf_h100", "c2_test2_mui_lf_h100", "c3_test2_dp_rk4_cspm_h30",
         # 3D dambreak with the C4 parameter set, 20 772 particles, steps 1 / 10 / 20 (2.6 h of emulator)
         "wc3d_20k_lf"]

# float64 restatement of the same serial algorithm: only summation-order / libm noise is allowed
TOL = 1e-9
# strain_equ_p integrates lambda*g_p/d_stress[a][b] over components with |d_stress| > 1e-8 (dp:111-118, 203-206):
# a division by near-zero stress-rate components, i.e. ill-conditioned by the reference's own formula.
FIELD_TOL = {"strain_equ_p": 5e-2}
FIELDS = ["x", "v", "density", "m_V", "pressure", "stress", "d_density", "d_vel", "d_stress", "v_grad", "strain_equ",
          "strain_equ_p", "density_tmp", "v_tmp", "stress_tmp", "CSPM_f", "CSPM_L"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_run(name):
    g = Golden(name)
    o = orc.Oracle.from_scene(g.scene, serial=1)
    assert o.n == g.meta["n"]
    assert o.P.dt == g.meta["dt"]
    assert [int(v) for v in o.D["grid_num"]] == g.meta["grid_num"]
    last = max(g.steps) if name.endswith("small_lf") or name.endswith("_cubic") or "tiny" in name or "indenter" in name or "plate" in name or "_h" in name or "rep" in name or "collision" in name or "dynrigid" in name or "20k" in name else min(max(g.steps), 10)
    for s in range(1, last + 1):
        if s in g.steps:
            # state right after the grid build + kernel correction of step s
            o.grid_build()
            o.calc_kernel_corr()
            assert np.array_equal(o.grid_ids, g.grid(s, "grid_ids")), f"cell ids differ at step {s}"
            assert np.array_equal(o.id0, g.grid(s, "id0")), f"sorted order differs at step {s}"
            assert np.array_equal(o.cell_end, g.grid(s, "grid_particle_num")), f"cell offsets differ at step {s}"
            # a shape-matched rigid body is its rest lattice rotated by a third-party SVD (ti.polar_decompose): its
            # lattice pairs sit exactly ON the support sphere, so from the second step on their count (never
            # their contribution, W = 0 there) depends on the last bit of that SVD -- checked for the first step only
            if "dynrigid" not in name or s == 1:
                assert np.array_equal(o.neighbor_count(), g.grid(s, "neighbor_count")), f"neighbour counts differ at step {s}"
            assert relmax(o.CSPM_f, g.grid(s, "CSPM_f")) < TOL
            assert relmax(o.CSPM_L, g.grid(s, "CSPM_L")) < TOL
        bad = o.step()          # (re-runs the idempotent grid build)
        assert bad == 0
        if s in g.steps:
            assert np.array_equal(o.id0, g.end(s, "id0"))
            assert np.array_equal(o.flag_retmap, g.end(s, "flag_retmap"))
            for f in FIELDS:
                err = relmax(getattr(o, f), g.end(s, f))
                assert err < FIELD_TOL.get(f, TOL), f"{name}: step {s} field {f}: rel err {err:.3e}"
