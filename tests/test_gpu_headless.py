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


@pytest.mark.parametrize("name,prec", [("wc2d_small_lf", "f64"), ("wc2d_small_lf", "f32"), ("dp2d_small_lf", "f64")])
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
