"""The headless driver on the GPU: exports of a real run, checkpoint / resume, the run_simulation.py entry point."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import Golden, make_sim, engine_fields

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_headless_run_exports_and_stops(tmp_path):
    from tisphi_b200.eng import ui_sim as U
    g = Golden("wc2d_small_lf")
    sim = make_sim(g.scene, precision="f64", stepsPerRenderUpdate=5, exitAtStep=20, exportEveryRender=2, exportCSV=True,
                   exportVTK=True, exportFrame=True)
    res = U.ui_sim(sim, out_dir=str(tmp_path), log=lambda *a: None)
    assert res["steps"] == 20 and res["reason"] == "exit" and res["bad_cells"] == 0
    names = sorted(os.path.basename(f) for f in res["files"])
    assert names == sorted([f"{p}{s}{q}" for s in ("000000", "000010", "000020")
                            for p, q in (("sim.csv.", ".csv"), ("sim.vtk.", ".vtu"), ("", ".npz"))])
    assert os.path.exists(os.path.join(res["simpath"], "_info.txt"))
    assert "Solver: Weakly Compressible" in open(os.path.join(res["simpath"], "_info.txt")).read()
    # the last CSV holds the state the engine holds
    rows = np.loadtxt(os.path.join(res["simpath"], "sim.csv.000020.csv"), delimiter=",")
    f = engine_fields(sim)
    assert np.array_equal(rows[:, 0].astype(np.int64), f["id0"])
    assert np.allclose(rows[:, 3:6], f["x"], rtol=0, atol=1e-15) and np.allclose(rows[:, 9], f["density"], rtol=1e-15)
    vtu = U.read_vtu(os.path.join(res["simpath"], "sim.vtk.000020.vtu"))
    assert np.array_equal(vtu["points"], f["x"]) and np.array_equal(vtu["density"], f["density"])


@pytest.mark.parametrize("name,prec", [("wc2d_small_lf", "f64"), ("wc2d_small_lf", "f32"), ("dp2d_small_lf", "f64"),
                                       ("mui2d_small_lf", "f64"), ("mui2d_small_lf", "f32")])    # mu(I): wall v_tmp is carried state (H27)
def test_checkpoint_resume_is_bit_exact(tmp_path, name, prec):
    from tisphi_b200.eng import ui_sim as U
    g = Golden(name)
    a = make_sim(g.scene, precision=prec, stepsPerRenderUpdate=4)
    U.ui_sim(a, max_steps=8, log=lambda *x: None)
    ck = U.save_checkpoint(str(tmp_path / "ck.npz"), a, 8)
    U.ui_sim(a, max_steps=8, log=lambda *x: None)                 # 8 more steps (the counter restarts, the state does not)
    b = make_sim(g.scene, precision=prec, stepsPerRenderUpdate=4)
    assert U.load_checkpoint(ck, b) == 8
    b.solver.run_steps(8)
    fa, fb = engine_fields(a), engine_fields(b)
    for k in ("id0", "x", "v", "density"):
        assert np.array_equal(fa[k], fb[k]), k


def test_run_simulation_entry_point(tmp_path):
    scene = os.path.join(ROOT, "data", "scenes", "test1_db_water.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "run_simulation.py"), "--scene_file", scene, "--max_steps", "20",
                        "--out_dir", str(tmp_path)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "========== SIMULATION ==========" in r.stdout and "========== END ==========" in r.stdout
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["steps"] == 20 and res["scene"] == "test1_db_water" and res["bad_cells"] == 0


def test_viewer_members_and_host_neighbour_helper():
    """assign_value_color / v_maxmin / set_color / copy2vis (ps:380-407, base:721-789) and the host-side
    for_all_neighbors helper against the engine's own neighbour counts."""
    import torch
    g = Golden("wc2d_small_lf")
    sim = make_sim(g.scene, precision="f64", colorTitle=7, colorGroup=0, showBdyPts=False)
    sim.solver.run_steps(5)
    ps, pt = sim.ps, sim.ps.pt
    ps.initialize_particle_system()
    sim.solver.assign_value_color()
    ps.v_maxmin(-1, -1, 0, 0)
    ps.set_color()
    ps.copy2vis(0.5)
    flow = pt.mat_type == 1
    p = pt.pressure.double()
    assert torch.equal(pt.val[flow].double(), p[flow])                       # colorTitle 7 = pressure
    assert ps.vmax[None] == float(p[flow].max()) and ps.vmin[None] == float(p[flow].min())
    col = pt.color
    assert col.shape == (ps.particle_num[None], 3) and float(col.min()) >= 0.0 and float(col.max()) <= 1.0
    hi = int(torch.argmax(torch.where(flow, p, torch.full_like(p, -1e300))))
    assert float(col[hi][0]) > float(col[hi][2])                              # highest pressure is drawn at the red end
    assert torch.allclose(pt.pos2vis[flow].double(), (pt.x[flow] * 0.5).float().double())
    # viewer members follow their particle through a sort
    before = {int(i): float(v) for i, v in zip(pt.id0[:50].tolist(), pt.val[:50].tolist())}
    sim.solver.run_steps(3)
    after = {int(i): float(v) for i, v in zip(pt.id0.tolist(), pt.val.tolist())}
    assert all(after[i] == v for i, v in before.items())
    # host for_all_neighbors == device neighbour count, and the helper's cell of x_i is the engine's grid id
    ps.initialize_particle_system()
    counts = ps.neighbor_count().cpu().numpy()
    gid = pt.grid_ids.cpu().numpy()
    for i in (0, 17, ps.particle_num[None] // 2, ps.particle_num[None] - 1):
        seen = []
        ps.for_all_neighbors(i, lambda a, b, ret: ret.append(b), seen)
        assert len(seen) == counts[i] and seen == sorted(seen) and i not in seen, i
        idx = ps.pos_to_index(pt.x[i].cpu().numpy())
        assert int(idx[0] * ps.grid_num[1] + idx[1]) == gid[i]
    # hydrostatic init_pressure (base:263-271) on fluid particles only
    sim.solver.init_pressure(1000.0)
    y = pt.x[:, 1]
    want = 1000.0 * 9.81 * (y[flow].max() - y[flow])
    assert torch.allclose(pt.pressure[flow].double(), want.double(), rtol=1e-12, atol=1e-9)


def test_mesh_body_points_become_particles(monkeypatch):
    """Bodies (ps:83-91, 176-199): the voxelised points of a mesh enter the particle set after the blocks and before the
    walls, with the body's material.  The voxeliser itself is trimesh (absent in this image), so load_body is replaced by
    a fake that returns a small lattice -- everything downstream of it is the real path."""
    import copy
    import torch
    from tisphi_b200.eng import particle_system as PS
    from tisphi_b200.eng.simulation import Simulation, SimConfiger
    g = Golden("wc2d_small_lf")
    scene = copy.deepcopy(g.scene)
    d = 2 * scene["Configuration"]["particleRadius"]
    pts = np.array([[0.6 + (i + 0.5) * d, 0.1 + (j + 0.5) * d, 0.0] for i in range(4) for j in range(5)])
    scene["Bodies"] = [dict(objectId=7, materialId=0, geometryFile="unused.obj", translation=[0, 0, 0], scale=[1, 1, 1],
                            rotationAxis=[0, 0, 1], rotationAngle=0.0, velocity=[0.0, -1.0, 0.0], isDynamic=1)]
    monkeypatch.setattr(PS, "load_body", lambda body, vox_len: pts.copy())
    sim = Simulation(SimConfiger(config=scene))
    n0 = g.meta["n"]
    assert sim.ps.particle_num[None] == n0 + len(pts)
    obj = sim.ps.pt.obj_id.cpu().numpy()
    idx = np.nonzero(obj == 7)[0]
    assert len(idx) == len(pts)
    assert np.allclose(sim.ps.pt.x.cpu().numpy()[idx], pts) and np.allclose(sim.ps.pt.v.cpu().numpy()[idx, 1], -1.0)
    sim.solver.run_steps(5)
    assert bool(torch.isfinite(sim.ps.pt.v).all())
