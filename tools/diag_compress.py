"""Diagnostic: block thrown against a wall; compares the tile path (+ fallback) with the all-generic engine step by step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import make_sim, engine_fields, relmax
from tisphi_b200 import scenes
speed = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
scene = scenes.dambreak3d(scale=0.5, precision="f32")
scene["Configuration"]["domainEnd"] = [1.0, 0.6, 0.4]
scene["Blocks"][0].update(size=[0.4, 0.3, 0.4], velocity=[-speed, 0.0, 0.0])
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    a = make_sim(scene, precision="f32", fastSweeps=True)
    b = make_sim(scene, precision="f32", fastSweeps=False)
eng = a.ps.engine
for s in range(8):
    a.solver.run_steps(10)
    for name in ("X", "V", "MASS", "M_V", "DENSITY", "PRESSURE", "MAT_TYPE", "ID0"):
        b.ps.engine.field(name).copy_(a.ps.engine.field(name))
    a.solver.step(); b.solver.step()
    fa, fb = engine_fields(a), engine_fields(b)
    a.ps.initialize_particle_system(); a.solver.calc_kernel_corr()
    out = torch.empty(eng.n, dtype=torch.int32, device=eng.device)
    eng.call("sph_neighbor_count_masks", out.data_ptr())
    flow = a.ps.pt.mat_type > 0
    nb = a.ps.neighbor_count()
    cnt = torch.bincount(a.ps.pt.grid_ids.long())
    line = {k: (relmax(fa[k], fb[k])) for k in ("x", "v", "density", "pressure", "d_vel", "d_density", "CSPM_f")}
    print(s, "flagged-out flow", int(((out < 0) & flow).sum()), "of", int(flow.sum()), "max cell pop", int(cnt.max()),
          "max nb", int(nb.max()), "rho max", float(fa["density"].max()), "|v| max", float(np.abs(fa["v"]).max()),
          "nan a/b", int(np.isnan(fa["v"]).sum()), int(np.isnan(fb["v"]).sum()), "bad", eng.L.sph_read_bad_cells(eng.h),
          {k: f"{v:.2e}" for k, v in line.items()}, flush=True)
