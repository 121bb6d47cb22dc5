import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from helpers import Golden, relmax, make_sim, engine_fields
g = Golden("wc3d_tiny_lf")
for fresh in (False, True):
    a = make_sim(g.scene, precision="f32", fastSweeps=True, wcFresh=fresh)
    b = make_sim(g.scene, precision="f32", fastSweeps=False, wcFresh=fresh)
    for s in range(1, 4):
        for e in (a, b): e.solver.step()
        fa, fb = engine_fields(a), engine_fields(b)
        fl = fa["mat_type"] > 0
        wl = ~fl
        out = [fresh, s]
        for k in ("v", "d_vel", "pressure", "density", "v_tmp", "CSPM_f", "d_density"):
            out.append((k, "flow %.2e" % relmax(fa[k][fl], fb[k][fl]), "wall %.2e" % relmax(fa[k][wl], fb[k][wl])))
        print(out, flush=True)
        if s == 2:
            d = np.abs(fa["pressure"] - fb["pressure"])
            i = np.argsort(-d)[:8]
            print("worst pressure:", [(int(j), int(fa["mat_type"][j]), float(fa["pressure"][j]), float(fb["pressure"][j]), float(fa["CSPM_f"][j]), float(fb["CSPM_f"][j])) for j in i], flush=True)
