"""Python front-end of the CPU oracle (ctypes over oracle/sph_oracle.c).   TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this
module; the product (tisphi_b200/) never does.  It restates, independently of the product's host code:
  * scene -> particle set            eng/particle_func.py:176-259, 302-343; eng/particle_system.py:14-131
  * discretisation constants, dt     eng/particle_system.py:32-59; eng/solver_sph_base.py:209-212
  * solver constants                 eng/solver_sph_wc.py:12-21; eng/solver_sph_muI.py:12-25; eng/solver_sph_dp.py:12-35
and drives the C restatement of the step (sph_oracle.c).
"""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libsph_oracle.so")

FIELDS = ["x", "v", "m_V", "density", "mass", "pressure", "stress", "CSPM_f", "CSPM_L", "d_density", "d_vel",
          "d_stress", "v_grad", "strain_equ", "d_strain_equ", "strain_equ_p", "d_strain_equ_p", "density_tmp",
          "v_tmp", "stress_tmp", "d_density_RK", "d_vel_RK", "d_stress_RK", "x0"]
NCOMP = [3, 3, 1, 1, 1, 1, 9, 1, 9, 1, 3, 9, 9, 1, 1, 1, 1, 1, 3, 9, 1, 3, 9, 3]
IFIELDS = ["mat_type", "id0", "grid_ids", "flag_retmap", "obj_id", "is_dynamic"]


def build(force=False):
    """Compile the C restatement (gcc + OpenMP).  Building the checker is not using it."""
    src = os.path.join(HERE, "sph_oracle.c")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", LIB, "-lm"])
    return LIB


class OrcParams(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("dim", "kernel", "kcorr", "ti", "xsph", "solver", "serial", "wc_fresh")] + \
               [("gn", C.c_int * 3)] + \
               [("h", C.c_double), ("support", C.c_double), ("grid_size", C.c_double), ("vstart", C.c_double * 3),
                ("m_V0", C.c_double), ("g", C.c_double * 3), ("dt", C.c_double), ("eps", C.c_double)] + \
               [(k, C.c_double) for k in ("rho0", "visc", "stiff", "gamma_", "coh", "fric", "E", "poi", "dila",
                                          "vsound", "mu", "alpha", "kc", "G", "K", "eps_f")] + \
               [("boundary", C.c_int), ("pad_", C.c_int), ("radius", C.c_double), ("dstart", C.c_double * 3),
                ("dend", C.c_double * 3)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcParams), C.c_int64]
        L.orc_field.restype = C.POINTER(C.c_double)
        L.orc_field.argtypes = [C.c_void_p, C.c_int]
        L.orc_ifield.restype = C.POINTER(C.c_int32)
        L.orc_ifield.argtypes = [C.c_void_p, C.c_int]
        L.orc_cell_end.restype = C.POINTER(C.c_int64)
        L.orc_cell_end.argtypes = [C.c_void_p]
        for fn in ("orc_destroy", "orc_calc_kernel_corr", "orc_init_real2tmp", "orc_one_step", "orc_init_stress",
                   "orc_advect_pos", "orc_post_step", "orc_enforce_boundary", "orc_init_rigid_body", "orc_solve_rigid_body"):
            getattr(L, fn).argtypes = [C.c_void_p]
            getattr(L, fn).restype = None
        L.orc_set_params.argtypes = [C.c_void_p, C.POINTER(OrcParams)]
        L.orc_advect.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_one_step_phase.argtypes = [C.c_void_p, C.c_int]
        L.orc_one_step_phase.restype = None
        L.orc_grid_build.argtypes = [C.c_void_p]
        L.orc_grid_build.restype = C.c_int64
        L.orc_step.argtypes = [C.c_void_p]
        L.orc_step.restype = C.c_int64
        for fn in ("orc_neighbor_count", "orc_neighbor_count_f32", "orc_neighbor_count_f32local", "orc_density_sum",
                   "orc_density_sum_f32pos"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p]
            getattr(L, fn).restype = None
        L.orc_set_threads.argtypes, L.orc_set_threads.restype = [C.c_int], None
        L.orc_max_threads.argtypes, L.orc_max_threads.restype = [], C.c_int
        L.orc_r2_threshold_f32.argtypes = [C.c_float]
        L.orc_r2_threshold_f32.restype = C.c_float
        L.orc_r2_threshold_f64.argtypes = [C.c_double]
        L.orc_r2_threshold_f64.restype = C.c_double
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------- scene -> particles
def _axis(t, size, d):
    off = d if size >= 0 else -d                      # pf:257
    return np.arange(t + off / 2.0, t + size + 1e-5, off)   # pf:258


def cube_positions(lower, size, dim, d):
    """pf:176-196: per-axis aranges, meshgrid(indexing='ij'), x slowest / last axis fastest, z = 0 in 2D."""
    axes = [_axis(lower[a], size[a], d) for a in range(dim)]
    if dim == 2:
        axes.append(np.array([0.0]))
    g = np.array(np.meshgrid(*axes, sparse=False, indexing="ij"), dtype=np.float64)
    return g.reshape(3, -1).transpose().copy()


def discretisation(cfg):
    """ps:14-59."""
    dim = 2 if cfg["is2D"] else 3
    ds, de = np.array(cfg["domainStart"], dtype=np.float64), np.array(cfg["domainEnd"], dtype=np.float64)
    d = 2 * cfg["particleRadius"]
    h = cfg["kh"] * d
    out = dict(dim=dim, d=d, h=h, support=cfg["kappa"] * h, m_V0=d ** dim)
    gs = float(math.ceil(cfg["kappa"] * cfg["kh"])) * d
    vs, ve = ds - gs, de + gs
    if dim == 2:
        de[2] = ds[2] + d
        vs[2] = ds[2]
        ve[2] = vs[2] + d
    gn = np.ceil((ve - vs) / gs).astype(int)
    out.update(grid_size=gs, domain_start=ds, domain_end=de, vstart=vs, vend=ve, grid_num=gn,
               C=int(np.prod(gn[:dim])))
    return out


def dummy_boxes(dim, ds, de, vs, ve):
    """pf:316-343 (no lid)."""
    if dim == 3:
        return [(np.array([vs[0], ds[1], vs[2]]), np.array([ds[0], de[1], de[2]])),
                (np.array([vs[0], ds[1], de[2]]), np.array([de[0], de[1], ve[2]])),
                (np.array([de[0], ds[1], ds[2]]), np.array([ve[0], de[1], ve[2]])),
                (np.array([ds[0], ds[1], vs[2]]), np.array([ve[0], de[1], ds[2]])),
                (vs.copy(), np.array([ve[0], ds[1], ve[2]]))]
    return [(np.array([vs[0], ds[1], ds[2]]), np.array([ds[0], de[1], de[2]])),
            (np.array([vs[0], vs[1], ds[2]]), np.array([ve[0], ds[1], de[2]])),
            (np.array([de[0], ds[1], ds[2]]), np.array([ve[0], de[1], de[2]]))]


def rep_boxes(dim, ds, de, radius):
    """pf:346-374: one layer of repulsive particles ON the domain faces (no lid), half a radius thick."""
    t = radius / 2
    if dim == 3:
        return [(np.array([ds[0] - t, ds[1] + t, ds[2] - t]), np.array([ds[0] + t, de[1] - t, de[2] - t])),
                (np.array([ds[0] - t, ds[1] + t, de[2] - t]), np.array([de[0] - t, de[1] - t, de[2] + t])),
                (np.array([de[0] - t, ds[1] + t, ds[2] + t]), np.array([de[0] + t, de[1] - t, de[2] + t])),
                (np.array([ds[0] + t, ds[1] + t, ds[2] - t]), np.array([de[0] + t, de[1] - t, ds[2] + t])),
                (ds - t, np.array([de[0], ds[1], de[2]]) + t)]
    return [(np.array([ds[0] - t, ds[1] + t, ds[2]]), np.array([ds[0] + t, de[1] - t, de[2]])),
            (np.array([ds[0] - t, ds[1] - t, ds[2]]), np.array([de[0] + t, ds[1] + t, de[2]])),
            (np.array([de[0] - t, ds[1] + t, ds[2]]), np.array([de[0] + t, de[1] - t, de[2]]))]


def build_particles(scene):
    """Particle set in creation order (blocks in JSON order, then dummy boxes): ps:135-174, pf:308-313."""
    cfg = scene["Configuration"]
    D = discretisation(cfg)
    dim, d = D["dim"], D["d"]
    mats = {m["matId"]: m for m in scene.get("Materials", [])}
    xs, vs_, rho, typ = [], [], [], []
    obj, dyn = [], []                                 # object id and is_dynamic per particle (ps:150-174)
    for b in scene.get("Blocks", []):
        for a in range(dim):                          # pf:246-251
            assert b["translation"][a] - D["domain_start"][a] >= 0.0 and \
                b["translation"][a] + b["size"][a] - D["domain_end"][a] <= 0.0, "Block is not in domain!"
        m = mats[b["materialId"]]
        p = cube_positions(np.array(b["translation"], dtype=np.float64), np.array(b["size"], dtype=np.float64), dim, d)
        xs.append(p)
        vs_.append(np.tile(np.array(b["velocity"], dtype=np.float64), (len(p), 1)))
        rho.append(np.full(len(p), float(m["density0"])))
        typ.append(np.full(len(p), int(m["matType"]), dtype=np.int32))
        obj.append(np.full(len(p), int(b["objectId"]), dtype=np.int32))
        dyn.append(np.full(len(p), int(b["isDynamic"]) if m["matType"] > 10 else 1, dtype=np.int32))
    if cfg["boundary"] in (2, 4):
        for lo, hi in dummy_boxes(dim, D["domain_start"], D["domain_end"], D["vstart"], D["vend"]):
            p = cube_positions(lo, hi - lo, dim, d)
            xs.append(p)
            vs_.append(np.zeros_like(p))
            rho.append(np.zeros(len(p)))              # walls: density 0 -> mass 0 (pf:206, ps:282)
            typ.append(np.full(len(p), -1, dtype=np.int32))
            obj.append(np.full(len(p), -1, dtype=np.int32))       # add_boundary: object id = the type (pf:308-313)
            dyn.append(np.ones(len(p), dtype=np.int32))
    if cfg["boundary"] in (3, 4):                     # ps:147-148: spacing = particle radius, type -2
        for lo, hi in rep_boxes(dim, D["domain_start"], D["domain_end"], cfg["particleRadius"]):
            p = cube_positions(lo, hi - lo, dim, cfg["particleRadius"])
            xs.append(p)
            vs_.append(np.zeros_like(p))
            rho.append(np.zeros(len(p)))
            typ.append(np.full(len(p), -2, dtype=np.int32))
            obj.append(np.full(len(p), -2, dtype=np.int32))
            dyn.append(np.ones(len(p), dtype=np.int32))
    build_particles.last_obj_dyn = (np.concatenate(obj), np.concatenate(dyn))
    return D, np.concatenate(xs), np.concatenate(vs_), np.concatenate(rho), np.concatenate(typ)


def calc_dt_cfl(h, vsound, dt_min, cfl=0.2):
    """base:209-212 with Taichi's float modulo a - floor(a/b)*b."""
    dt = cfl * h / vsound
    return max(dt_min, dt - (dt - math.floor(dt / dt_min) * dt_min))


def make_params(scene, serial=1, wc_fresh=0):
    cfg = scene["Configuration"]
    D = discretisation(cfg)
    P = OrcParams()
    P.dim, P.kernel, P.kcorr, P.ti = D["dim"], cfg["kernel"], cfg["kernelCorrection"], cfg["timeIntegration"]
    P.xsph, P.solver, P.serial, P.wc_fresh = int(bool(cfg["xsph"])), cfg["simulationMethod"], serial, wc_fresh
    for a in range(3):
        P.gn[a] = int(D["grid_num"][a])
        P.vstart[a] = float(D["vstart"][a])
        P.g[a] = float(cfg["gravitation"][a])
    P.h, P.support, P.grid_size, P.m_V0, P.eps = D["h"], D["support"], D["grid_size"], D["m_V0"], 1e-8
    P.boundary, P.radius = int(cfg["boundary"]), float(cfg["particleRadius"])
    for a in range(3):
        P.dstart[a], P.dend[a] = float(D["domain_start"][a]), float(D["domain_end"][a])
    fluids = [m for m in scene.get("Materials", []) if m["matType"] == 1]
    soils = [m for m in scene.get("Materials", []) if m["matType"] == 2]
    dt_min = cfg["timeStepSizeMin"]
    if P.solver == 1:
        m = fluids[0]
        P.rho0, P.visc, P.stiff, P.gamma_ = m["density0"], m["viscosity"], m["stiffness"], m["exponent"]
        P.vsound = 60.0                                                   # wc:17
    else:
        m = soils[0]
        P.rho0, P.coh = m["density0"], m["cohesion"]
        P.fric = m["friction"] / 180 * math.pi
        P.E, P.poi = m["EYoungMod"], m["poison"]
        P.dila = m["dilatancy"] / 180 * math.pi
        P.mu = math.tan(P.fric)
        if P.solver == 2:
            P.vsound = 24.0                                               # muI:23
        else:
            P.vsound = math.sqrt(P.E / P.rho0)                            # dp:20-21
            t = math.tan(P.fric)
            P.alpha = t / math.sqrt(9 + 12 * t ** 2)                      # dp:26-29
            P.kc = 3 * P.coh / math.sqrt(9 + 12 * t ** 2)
            P.G = P.E / (2 * (1 + P.poi))
            P.K = P.E / (3 * (1 - 2 * P.poi))
            P.eps_f = 1e-4
    P.dt = calc_dt_cfl(P.h, P.vsound, dt_min)
    return P, D


class Oracle:
    def __init__(self, params, x, v, density, mat_type, obj_id=None, is_dynamic=None):
        self.L = lib()
        self.P = params
        self.n = len(x)
        self.h = self.L.orc_create(C.byref(params), self.n)
        self._views()
        self.x[:] = x
        self.v[:] = v
        self.density[:] = density
        self.m_V[:] = params.m_V0                                         # ps:281-282
        self.mass[:] = params.m_V0 * np.asarray(density)
        self.mat_type[:] = mat_type
        self.id0[:] = np.arange(self.n)                                   # ps:208-211
        self.x0[:] = x
        self.obj_id[:] = 0 if obj_id is None else obj_id
        self.is_dynamic[:] = 1 if is_dynamic is None else is_dynamic
        self.L.orc_init_rigid_body(self.h)                                # base:28
        if params.solver == 3:
            self.L.orc_init_stress(self.h)                                # dp:35

    @classmethod
    def from_scene(cls, scene, serial=1, wc_fresh=0):
        P, D = make_params(scene, serial, wc_fresh)
        D2, x, v, rho, typ = build_particles(scene)
        obj, dyn = build_particles.last_obj_dyn
        o = cls(P, x, v, rho, typ, obj, dyn)
        o.D = D
        return o

    def _views(self):
        n = self.n
        for k, (nm, nc) in enumerate(zip(FIELDS, NCOMP)):
            a = np.ctypeslib.as_array(self.L.orc_field(self.h, k), shape=(n * nc,))
            setattr(self, nm, a if nc == 1 else a.reshape(n, nc))
        for k, nm in enumerate(IFIELDS):
            setattr(self, nm, np.ctypeslib.as_array(self.L.orc_ifield(self.h, k), shape=(n,)))
        Cn = int(self.P.gn[0]) * int(self.P.gn[1]) * (int(self.P.gn[2]) if self.P.dim == 3 else 1)
        self.cell_end = np.ctypeslib.as_array(self.L.orc_cell_end(self.h), shape=(Cn,))

    def set_params(self):
        self.L.orc_set_params(self.h, C.byref(self.P))

    def grid_build(self):
        return self.L.orc_grid_build(self.h)

    def calc_kernel_corr(self):
        self.L.orc_calc_kernel_corr(self.h)

    def step(self):
        return self.L.orc_step(self.h)

    def neighbor_count(self, f32=False, f32global=False):
        """f32: the MIXED engine's predicate (cell-local float32 coordinates); f32global: globally rounded float32."""
        out = np.zeros(self.n, dtype=np.int32)
        fn = self.L.orc_neighbor_count_f32 if f32global else (self.L.orc_neighbor_count_f32local if f32 else self.L.orc_neighbor_count)
        fn(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def density_sum(self, f32pos=False):
        out = np.zeros(self.n, dtype=np.float64)
        (self.L.orc_density_sum_f32pos if f32pos else self.L.orc_density_sum)(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass


def load_scene(path):
    with open(path) as f:
        return json.load(f)
