"""Shared helpers for the parity tests (fixtures are produced by oracle/gen_golden.py from the reference sources)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))
        self.scene = self.meta["scene"]
        self.steps = self.meta["steps"]

    def end(self, step, field):
        return self.z[f"s{step}/end/{field}"]

    def grid(self, step, field):
        return self.z[f"s{step}/grid/{field}"]


def relmax(a, b):
    """max-norm relative difference: max|a-b| / max|b| (0 if both are identically zero)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    num = np.max(np.abs(a - b)) if a.size else 0.0
    return 0.0 if num == 0.0 else (num / den if den > 0 else np.inf)


def relelem(a, b, floor_frac=1e-3):
    """ELEMENT-WISE relative difference max_k |a_k - b_k| / max(|b_k|, floor) with the absolute floor
    floor = floor_frac * max|b|: entries much smaller than the field's scale are held to the floor, every other entry
    to its own magnitude (the max-norm of relmax leaves small entries unconstrained)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = np.max(np.abs(b))
    if scale == 0.0:
        return 0.0 if np.max(np.abs(a)) == 0.0 else np.inf
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor_frac * scale)))


def sym6_to_9(s6):
    """(n,6) xx,yy,zz,xy,yz,zx -> (n,9) row-major full tensor."""
    xx, yy, zz, xy, yz, zx = [s6[:, k] for k in range(6)]
    return np.stack([xx, xy, zx, xy, yy, yz, zx, yz, zz], axis=1)


def make_sim(scene, precision="f64", device="cuda:0", **extra):
    """Simulation built through the reference-shaped API of tisphi_b200.eng from an in-memory scene dict."""
    import copy
    from tisphi_b200.eng.simulation import Simulation, SimConfiger
    sc = copy.deepcopy(scene)
    sc["Configuration"]["precision"] = precision
    sc["Configuration"].update(extra)
    return Simulation(SimConfiger(config=sc), device=device)


def engine_fields(sim):
    """Every comparable particle member as float64/int64 numpy arrays in current order (tensors as (n, 9))."""
    pt, ps = sim.ps.pt, sim.ps
    out = {}
    for k in ("x", "v", "density", "m_V", "mass", "pressure", "d_density", "d_vel", "density_tmp", "v_tmp", "CSPM_f"):
        out[k] = getattr(pt, k).detach().cpu().double().numpy()
    for k in ("id0", "grid_ids", "mat_type"):
        out[k] = getattr(pt, k).detach().cpu().long().numpy()
    if sim.solver_type != 1:
        for k in ("stress", "d_stress", "stress_tmp"):
            out[k] = sym6_to_9(ps.sym6(k).detach().cpu().double().numpy())
        out["v_grad"] = pt.v_grad.detach().cpu().double().numpy().reshape(-1, 9)
        for k in ("strain_equ", "strain_equ_p", "d_strain_equ", "d_strain_equ_p"):
            out[k] = getattr(pt, k).detach().cpu().double().numpy()
        out["flag_retmap"] = pt.flag_retmap.detach().cpu().long().numpy()
    if sim.ps.params.kcorr == 1:
        out["CSPM_L"] = pt.CSPM_L.detach().cpu().double().numpy().reshape(-1, 9)
    return out
