"""Does the thrown-block scene blow up in every engine (physics of the scheme) or only on the tile path?"""
import sys, os, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import make_sim
from tisphi_b200 import scenes
speed = float(sys.argv[1]) if len(sys.argv) > 1 else 15.0
scene = scenes.dambreak3d(scale=0.5, precision="f32")
scene["Configuration"]["domainEnd"] = [1.0, 0.6, 0.4]
scene["Blocks"][0].update(size=[0.4, 0.3, 0.4], velocity=[-speed, 0.0, 0.0])
for label, kw in (("tile f32", dict(precision="f32", fastSweeps=True)), ("generic f32", dict(precision="f32", fastSweeps=False)),
                  ("generic f64", dict(precision="f64"))):
    with contextlib.redirect_stdout(io.StringIO()):
        sim = make_sim(scene, **kw)
    print(label, "dt", sim.solver.dt[None])
    for s in range(12):
        sim.solver.step()
        pt = sim.ps.pt
        fl = pt.mat_type > 0
        v = pt.v[fl]; rho = pt.density[fl]
        print("  step", s + 1, "rho max %.1f min %.1f" % (float(rho.max()), float(rho.min())), "|v|max %.2f" % float(v.norm(dim=1).max()),
              "p max %.3e" % float(pt.pressure.max()), "xmin %.4f" % float(pt.x[fl][:, 0].min()),
              "nan", int(torch.isnan(v).sum()), "bad", sim.ps.engine.L.sph_read_bad_cells(sim.ps.engine.h), flush=True)
        if int(torch.isnan(v).sum()):
            break
