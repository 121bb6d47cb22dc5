"""ParticleSystem: particle storage (SoA inside one torch arena), uniform grid, counting sort, dump().

Mirror of the reference's eng/particle_system.py (class, attribute and method names kept):
  __init__ ps:12-133, initialize_particles ps:135-148, init_block ps:150-174, set_id0 ps:208-211,
  update_grid_id / counting_sort / initialize_particle_system ps:229-257, dump ps:459-545.
Device work goes through the C ABI of libtisphi_b200 (tisphi_b200/_lib.py); there is no device code in Python.
"""
import math

import numpy as np

from .. import _lib
from .configer_builder import SimConfiger
from .particle_func import (add_boundary, add_cube, calc_cube_particle_num, calc_dummy_boundary, calc_rep_boundary, chk_block_in_domain,
                            count_boundary, get_material, load_body, set_material)


class _Scalar:
    """0-d field shim: ``ps.particle_num[None]`` / ``solver.dt[None]`` keep working."""

    def __init__(self, value=0, on_set=None):
        self._v = value
        self._on_set = on_set

    def __getitem__(self, key):
        return self._v

    def __setitem__(self, key, value):
        self._v = value
        if self._on_set is not None:
            self._on_set(value)


_SCALARS = {"m_V": "M_V", "density": "DENSITY", "mass": "MASS", "pressure": "PRESSURE", "mat_type": "MAT_TYPE",
            "id0": "ID0", "grid_ids": "GRID_IDS", "id_new": "ID_NEW", "CSPM_f": "CSPM_F", "d_density": "D_DENSITY",
            "flag_retmap": "FLAG_RETMAP", "strain_equ": "STRAIN_EQU", "d_strain_equ": "D_STRAIN_EQU",
            "strain_equ_p": "STRAIN_EQU_P", "d_strain_equ_p": "D_STRAIN_EQU_P", "density_tmp": "DENSITY_TMP",
            "d_density_RK": "D_DENSITY_RK"}
_VECTORS = {"x": "X", "v": "V", "d_vel": "D_VEL", "v_tmp": "V_TMP", "d_vel_RK": "D_VEL_RK"}
_SYM = {"stress": "STRESS", "d_stress": "D_STRESS", "stress_tmp": "STRESS_TMP", "d_stress_RK": "D_STRESS_RK"}
_FULL = {"CSPM_L": "CSPM_L", "v_grad": "V_GRAD"}
_CONST = ("obj_id", "mat_id", "is_dynamic", "color", "x0")     # per-particle constants, keyed by id0
_SYM_IDX = [0, 3, 5, 3, 1, 4, 5, 4, 2]                         # xx,yy,zz,xy,yz,zx -> row-major 3x3
_DEAD = {"MLS_beta": (4,), "dist_B": (), "g_p": (), "drag_force": (3,)}   # pf:13-69, dead in the reference


# ---- host-side restatements of the reference's device helpers (pure functions of arrays: also used by the tests)
def pos_to_index(pos, vdomain_start, grid_size):
    """ps:216-218: trunc((pos - vdomain_start) / grid_size) per axis (C-style cast)."""
    q = (np.asarray(pos, dtype=np.float64) - np.asarray(vdomain_start, dtype=np.float64)) / float(grid_size)
    return np.trunc(q).astype(np.int64)


def flatten_grid_index(grid_index, grid_num):
    """ps:220-222: x-major, z fastest."""
    g = np.asarray(grid_index, dtype=np.int64)
    return g[..., 0] * int(grid_num[1]) * int(grid_num[2]) + g[..., 1] * int(grid_num[2]) + g[..., 2]


def value_range(val, flow, givenmax, givenmin, fixmax, fixmin):
    """ps:387-397 v_maxmin: extrema of `val` over flow particles, optionally overridden / capped by the given values."""
    import torch
    sel = val[flow]
    vmax = float(sel.max()) if sel.numel() else -float("inf")
    vmin = float(sel.min()) if sel.numel() else float("inf")
    out_max = vmax if (givenmax == -1 or (vmax < givenmax and not fixmax)) else givenmax
    out_min = vmin if (givenmin == -1 or (vmin > givenmin and not fixmin)) else givenmin
    return out_max, out_min


class _ParticleFields:
    """``ps.pt.<member>``: torch views (no copy) of the members in current, i.e. sorted, order.

    Scalars and vectors are live views; symmetric tensors are stored as 6 components and returned as (n, 3, 3)
    copies (use ``ps.sym6(name)`` for the live 6-component view)."""

    def __init__(self, ps):
        object.__setattr__(self, "_ps", ps)

    def __getattr__(self, name):
        ps = self._ps
        eng = ps.engine
        if name in _SCALARS:
            return eng.field(_SCALARS[name])
        if name in _VECTORS:
            return eng.field(_VECTORS[name])
        if name in _SYM:
            try:
                s6 = eng.field(_SYM[name])
            except _lib.SphError:
                return ps._torch.zeros((eng.n, 3, 3), dtype=eng.real, device=eng.device)
            return s6[:, _SYM_IDX].reshape(-1, 3, 3)
        if name in _FULL:
            return eng.field(_FULL[name]).reshape(-1, 3, 3)
        if name in ("val", "pos2vis") or (name == "color" and ps._vis.get("color") is not None):
            return ps._vis_field(name)
        if name in _CONST:
            idx = eng.field("ID0").long()
            return ps._const_dev(name)[idx]
        if name in _DEAD:                       # members the reference declares but never computes
            shape = (eng.n,) + _DEAD[name]
            return ps._torch.zeros(shape, dtype=eng.real, device=eng.device)
        raise AttributeError(f"Particle has no member {name!r} in this engine")


class ParticleSystem:
    def __init__(self, config: SimConfiger, device="cuda:0", precision=None, slab=None) -> None:
        import torch
        self._torch = torch
        self.cfg = config
        self.dim = 3 if not self.cfg.get_cfg("is2D") else 2
        self.dim3 = 3
        self.domain_start = np.array(self.cfg.get_cfg("domainStart"), dtype=np.float64)
        self.domain_end = np.array(self.cfg.get_cfg("domainEnd"), dtype=np.float64)
        self.domain_size = self.domain_end - self.domain_start
        self.color_title = self.cfg.get_cfg("colorTitle")
        self.color_group = self.cfg.get_cfg("colorGroup")
        self.flag_boundary = self.cfg.get_cfg("boundary")
        self.show_bdy = self.cfg.get_cfg("showBdyPts")
        self.mat_dummy_type, self.mat_rep_type = -1, -2
        self.mat_fluid_type, self.mat_soil_type, self.mat_rigid_type = 1, 2, 11
        self.bdy_none, self.bdy_collision, self.bdy_dummy, self.bdy_rep, self.bdy_dummy_rep = 0, 1, 2, 3, 4
        self.i_dump = _Scalar(0)
        self.rigid_rest_cm = np.zeros(3)                         # rigid bodies are out of scope
        self._vis = {"val": None, "pos2vis": None, "color": None}   # viewer members, keyed by id0 (they survive the sort)

        # discretisation (ps:32-39)
        self.particle_radius = self.cfg.get_cfg("particleRadius")
        self.kappa = self.cfg.get_cfg("kappa")
        self.kh = self.cfg.get_cfg("kh")
        self.particle_diameter = 2 * self.particle_radius
        self.smoothing_len = self.kh * self.particle_diameter
        self.support_radius = self.kappa * self.smoothing_len
        self.m_V0 = self.particle_diameter ** self.dim
        self.particle_num = _Scalar(0)
        self.vmax = _Scalar(0.0)
        self.vmin = _Scalar(0.0)

        # grid (ps:46-59)
        self.grid_size = float(math.ceil(self.kappa * self.kh)) * self.particle_diameter
        self.vdomain_start = self.domain_start - self.grid_size
        self.vdomain_end = self.domain_end + self.grid_size
        self.vdomain_size = self.vdomain_end - self.vdomain_start
        if self.dim == 2:
            self.domain_end[2] = self.domain_start[2] + self.particle_diameter
            self.domain_size[2] = self.particle_diameter
            self.vdomain_start[2] = self.domain_start[2]
            self.vdomain_end[2] = self.vdomain_start[2] + self.particle_diameter
            self.vdomain_size[2] = self.particle_diameter
        self.grid_num = np.ceil(self.vdomain_size / self.grid_size).astype(int)
        self.grid_num_total = int(np.prod(self.grid_num[0:self.dim]))
        print("Grid num:", [int(self.grid_num[i]) for i in range(self.dim)], "total:", self.grid_num_total)

        self.mat_index, self.mat_fluid, self.mat_soil, self.mat_rigid = set_material(self)
        self.object_collection = dict()
        self.object_id_rigid = set()

        # particle counting (ps:72-108)
        self.blocks = self.cfg.get_blocks()
        block_particle_num = 0
        for block in self.blocks:
            chk_block_in_domain(self.domain_start, self.domain_end, block["translation"], block["size"], self.dim)
            particle_num, _ = calc_cube_particle_num(block["translation"], block["size"], self.dim,
                                                     offset=self.particle_diameter)
            block["particleNum"] = particle_num
            self.object_collection[block["objectId"]] = block
            block_particle_num += particle_num
            print("Block %d particle number: %d" % (block["objectId"], particle_num))
        self.bodies = self.cfg.get_bodies()                                          # ps:83-91
        body_particle_num = 0
        for body in self.bodies:
            points = load_body(body, self.particle_diameter)
            body["particleNum"], body["voxelizedPoints"] = points.shape[0], points
            self.object_collection[body["objectId"]] = body
            body_particle_num += points.shape[0]
            print("Body %d particle number: %d" % (body["objectId"], points.shape[0]))
        dummy_particle_num = 0
        if self.flag_boundary in (self.bdy_dummy, self.bdy_dummy_rep):
            self.dummy_boundary = calc_dummy_boundary(self.dim, self.domain_start, self.domain_end, self.vdomain_start,
                                                      self.vdomain_end)
            dummy_particle_num = count_boundary(self.dummy_boundary, self.dim, self.particle_diameter)
            print("Dummy particle number: %d" % dummy_particle_num)
        rep_particle_num = 0
        if self.flag_boundary in (self.bdy_rep, self.bdy_dummy_rep):                 # ps:100-105
            self.rep_boundary = calc_rep_boundary(self.dim, self.domain_start, self.domain_end, self.particle_radius)
            rep_particle_num = count_boundary(self.rep_boundary, self.dim, self.particle_radius)
            print("Repulsive particle number: %d" % rep_particle_num)
        self.particle_max_num = block_particle_num + body_particle_num + dummy_particle_num + rep_particle_num
        print(f"Particle total num: {self.particle_max_num}")

        # engine (replaces the Taichi struct fields pt / pt_buf, the grid counters and the prefix-sum executor)
        self.precision = precision if precision is not None else self.cfg.get_opt("precision", "f64")
        P = _lib.SphParams()
        P.dim = self.dim
        P.kernel = self.cfg.get_cfg("kernel")
        P.kcorr = self.cfg.get_cfg("kernelCorrection")
        P.ti = self.cfg.get_cfg("timeIntegration")
        P.xsph = int(bool(self.cfg.get_cfg("xsph")))
        P.solver = self.cfg.get_cfg("simulationMethod")
        P.precision = {"f64": _lib.PREC_F64, "fp64": _lib.PREC_F64, "f32": _lib.PREC_MIXED, "fp32": _lib.PREC_MIXED,
                       "mixed": _lib.PREC_MIXED}[self.precision]
        P.wc_fresh = int(bool(self.cfg.get_opt("wcFresh", False)))
        # 0: generic sweeps only; 1: cell-tile sweeps; 2: cell-tile sweeps + neighbour round lists replayed by the
        # later one_steps of a step.  Measured on C4 (profiles/): a replayed pass is 8 % faster than walking the masks
        # again but recording costs the first pass as much, so lists pay off with RK4 (three replays) and not with "LF"
        # (one); they also take 448 B per particle.  Default: on for timeIntegration 4 only.
        lists = self.cfg.get_opt("neighbourLists", P.ti == 4)
        P.fast = 0 if not self.cfg.get_opt("fastSweeps", True) else (2 if lists else 1)
        grav = self.cfg.get_cfg("gravitation")
        for a in range(3):
            P.gn[a] = int(self.grid_num[a])
            P.vstart[a] = float(self.vdomain_start[a])
            P.g[a] = float(grav[a])
        P.h, P.support, P.grid_size, P.m_V0, P.eps = self.smoothing_len, self.support_radius, self.grid_size, self.m_V0, 1e-8
        P.boundary, P.radius = int(self.flag_boundary), float(self.particle_radius)
        for a in range(3):
            P.dstart[a], P.dend[a] = float(self.domain_start[a]), float(self.domain_end[a])
        self.params = P
        self.pt = _ParticleFields(self)
        self.pt_buf = self.pt            # the ping-pong buffers are internal to the engine
        self._const = {k: [] for k in _CONST}
        self._const_cache = {}
        self._slab = slab
        self._device = device
        if slab is None:
            self.engine = _lib.Engine(P, max(self.particle_max_num, 1), device=device)
            self.initialize_particles()
        else:
            self._pending = []           # host arrays of the whole scene; only this rank's columns are uploaded
            self.initialize_particles()
            self._upload_slab()
        self.set_id0()
        self._upload_rigid_bodies()
        print("Particle system construction complete!")

    # ------------------------------------------------------------------------------------------ construction
    def initialize_particles(self):
        for block in self.blocks:
            self.init_block(block)
        for body in self.bodies:
            self.init_body(body)
        if self.flag_boundary in (self.bdy_dummy, self.bdy_dummy_rep):
            add_boundary(self, self.dummy_boundary, self.mat_dummy_type, color=[153, 153, 255])
        if self.flag_boundary in (self.bdy_rep, self.bdy_dummy_rep):                 # ps:147-148
            add_boundary(self, self.rep_boundary, self.mat_rep_type, offset=self.particle_radius, color=[170, 17, 255])

    def init_block(self, block):
        mat = get_material(self, block["materialId"])
        mat_type = mat["matType"]
        is_dynamic = True
        if mat_type > 10:                                   # ps:160-162
            self.object_id_rigid.add(block["objectId"])
            is_dynamic = bool(block["isDynamic"])
            if is_dynamic and self.params.solver == _lib.SOLVER_WC:
                raise NotImplementedError("a dynamic rigid body under WCSPH does not run in the reference either: "
                                          "WCSPHSolver.advect_something_func calls init_rigid_body, i.e. a kernel, from "
                                          "kernel scope (eng/solver_sph_wc.py:129-132); use the mu(I) or DP solver")
            if is_dynamic and self._slab is not None:
                raise NotImplementedError("dynamic rigid bodies are not supported on multi-GPU slabs (a body would span ranks)")
        add_cube(self, object_id=block["objectId"], lower_corner=np.array(block["translation"]),
                 cube_size=np.array(block["size"]), velocity=block["velocity"], density=mat["density0"],
                 is_dynamic=is_dynamic, color=np.array([ic / 255 for ic in mat["color"]], dtype=np.float32),
                 mat_id=block["materialId"], mat_type=mat_type)

    def init_body(self, body):
        """ps:176-199: the voxelised points of a mesh body as particles of its material."""
        mat = get_material(self, body["materialId"])
        mat_type, n = mat["matType"], body["particleNum"]
        is_dynamic = True
        if mat_type > 10:
            self.object_id_rigid.add(body["objectId"])
            is_dynamic = bool(body["isDynamic"])
            if is_dynamic and self.params.solver == _lib.SOLVER_WC:
                raise NotImplementedError("a dynamic rigid body under WCSPH does not run in the reference either (wc:129-132)")
        self._add_particles(body["objectId"], n, np.array(body["voxelizedPoints"], dtype=np.float64),
                            np.tile(np.array(body["velocity"], dtype=np.float64), (n, 1)),
                            np.full(n, mat["density0"], dtype=np.float64), np.zeros(n, dtype=np.float64),
                            np.full(n, body["materialId"], dtype=np.int32), np.full(n, mat_type, dtype=np.int32),
                            np.full(n, int(is_dynamic), dtype=np.int32),
                            np.tile(np.array([ic / 255 for ic in mat["color"]], dtype=np.float32), (n, 1)))

    def _add_particles(self, object_id, new_particles_num, new_particles_positions, new_particles_velocity,
                       new_particle_density, new_particle_pressure, new_particles_material_id,
                       new_particles_material_type, new_particles_is_dynamic, new_particles_color):
        """ps:289-314: append host arrays to the device arrays (pressure is always 0 at creation, ps:283)."""
        n = int(new_particles_num)
        if self._slab is not None:
            self._pending.append((np.asarray(new_particles_positions, dtype=np.float64),
                                  np.asarray(new_particles_velocity, dtype=np.float64),
                                  np.asarray(new_particle_density, dtype=np.float64),
                                  np.asarray(new_particles_material_type, dtype=np.int32)))
        else:
            self.engine.add_particles(new_particles_positions, new_particles_velocity, new_particle_density,
                                      new_particles_material_type)
        self._const["obj_id"].append(np.full(n, object_id, dtype=np.int32))
        self._const["mat_id"].append(np.asarray(new_particles_material_id, dtype=np.int32))
        self._const["is_dynamic"].append(np.asarray(new_particles_is_dynamic, dtype=np.int32))
        self._const["color"].append(np.asarray(new_particles_color, dtype=np.float32))
        self._const["x0"].append(np.asarray(new_particles_positions, dtype=np.float64))
        self._const_cache.clear()
        if self._slab is None:
            self.particle_num[None] = self.engine.n

    def _upload_slab(self):
        """Multi-GPU: choose the column partition from the whole scene, upload this rank's columns, keep GLOBAL id0."""
        from ..parallel import column_weights, partition_columns
        slab = self._slab
        x = np.concatenate([p[0] for p in self._pending])
        v = np.concatenate([p[1] for p in self._pending])
        rho = np.concatenate([p[2] for p in self._pending])
        typ = np.concatenate([p[3] for p in self._pending])
        self._pending = None
        n_cols = int(self.grid_num[0])
        w, cx = column_weights(x[:, 0], typ, float(self.vdomain_start[0]), self.grid_size, n_cols,
                               slab.get("wall_weight", 0.15))
        self.slab_columns = slab.get("columns") or partition_columns(w, slab["world"])
        a, b = self.slab_columns[slab["rank"]]
        mine = np.nonzero((cx >= a) & (cx < b))[0]
        counts = np.bincount(cx, minlength=n_cols)
        ghosts = int(counts[max(a - 1, 0):a].sum() + counts[b:b + 1].sum())
        # capacity: this rank's particles + both ghost columns, with head-room for inflow (a dam break front can
        # multiply the population of an initially dry slab); override with the optional scene key "slabCapacity"
        cap = self.cfg.get_opt("slabCapacity", None)
        if cap is None:
            cap = int(1.5 * (len(mine) + ghosts)) + 4 * int(counts.max()) + 1024
        self.global_particle_num = len(x)
        soil = typ == self.mat_soil_type
        self.slab_soil_ymax = float(x[soil, 1].max()) if soil.any() else None      # base:251-255 over the WHOLE scene
        # the most particles one slab message may carry: a boundary column + migrants (multi-GPU)
        self.slab_face_cap = int(self.cfg.get_opt("slabFaceCapacity", 2 * int(counts.max()) + 4096))
        self.engine = _lib.Engine(self.params, max(int(cap), 1), device=self._device)
        if len(mine):
            self.engine.add_particles(x[mine], v[mine], rho[mine], typ[mine])
            self.engine.field("ID0").copy_(self._torch.from_numpy(mine.astype(np.int32)).to(self.engine.device))
        self.particle_num[None] = self.engine.n
        print(f"slab rank {slab['rank']}/{slab['world']}: columns [{a}, {b}) of {n_cols}, {len(mine)} particles, capacity {cap}")

    def _upload_rigid_bodies(self):
        """Dynamic rigid bodies: per creation index the body a particle belongs to and its rest position."""
        self.rigid_dynamic_ids = sorted(o for o in self.object_id_rigid if bool(self.object_collection[o].get("isDynamic", 0)))
        self._rigid_tables = None
        if not self.rigid_dynamic_ids:
            return
        torch = self._torch
        obj = np.concatenate(self._const["obj_id"])
        dyn = np.concatenate(self._const["is_dynamic"])
        body = np.full(len(obj), -1, dtype=np.int32)
        for k, oid in enumerate(self.rigid_dynamic_ids):
            body[(obj == oid) & (dyn != 0)] = k
        x0 = np.ascontiguousarray(np.concatenate(self._const["x0"]), dtype=np.float64)
        dev = self.engine.device
        self._rigid_tables = (torch.from_numpy(body).to(dev), torch.from_numpy(x0).to(dev))       # kept alive: the engine reads them
        self.engine.call("sph_set_rigid_bodies", len(body), self._rigid_tables[0].data_ptr(), self._rigid_tables[1].data_ptr(),
                         len(self.rigid_dynamic_ids))

    def _const_dev(self, name):
        if name not in self._const_cache:
            arr = np.concatenate(self._const[name]) if self._const[name] else np.zeros(0)
            self._const_cache[name] = self._torch.from_numpy(arr).to(self.engine.device)
        return self._const_cache[name]

    def clear_particles(self):
        self.engine.call("sph_clear_particles")
        for k in self._const:
            self._const[k] = []
        self._const_cache.clear()
        self._vis = {"val": None, "pos2vis": None, "color": None}
        self.particle_num[None] = 0

    def set_id0(self):
        """ps:208-211.  id0 is assigned as the running creation index when particles are added."""
        return None

    # ------------------------------------------------------------------------------------------ grid + sort
    def initialize_particle_system(self):
        """ps:254-257: update_grid_id -> prefix sum -> counting_sort, as one native call."""
        self.engine.call("sph_grid_build")

    def update_grid_id(self):
        self.initialize_particle_system()

    def counting_sort(self):
        return None          # part of initialize_particle_system (the three reference kernels are one native call)

    @property
    def grid_particle_num(self):
        """Inclusive scan of the cell histogram (what the reference's field holds after prefix_sum_executor.run)."""
        return self.engine.field("CELL_END", count=self.grid_num_total)

    @property
    def grid_particle_num_temp(self):
        return self.engine.field("CELL_COUNT", count=self.grid_num_total)

    def sym6(self, name):
        return self.engine.field(_SYM[name])

    def neighbor_count(self):
        """Number of j with |x_i - x_j| < support_radius per particle (the for_all_neighbors predicate, ps:259-269)."""
        out = self._torch.empty(self.engine.n, dtype=self._torch.int32, device=self.engine.device)
        self.engine.call("sph_neighbor_count", out.data_ptr())
        return out

    def density_sum(self):
        out = self._torch.empty(self.engine.n, dtype=self.engine.real, device=self.engine.device)
        self.engine.call("sph_density_sum", out.data_ptr())
        return out

    # type predicates (ps:320-374), host-side helpers on type codes
    def is_fluid(self, t):
        return t == self.mat_fluid_type

    def is_soil(self, t):
        return t == self.mat_soil_type

    def is_flow(self, t):
        return (t == self.mat_fluid_type) | (t == self.mat_soil_type)

    def is_real(self, t):
        return t > 0

    def is_dummy(self, t):
        return t == self.mat_dummy_type

    def is_rep(self, t):
        return t == self.mat_rep_type

    def is_bdy(self, t):
        return (t == self.mat_dummy_type) | (t == self.mat_rep_type)

    def is_rigid(self, t):
        return t == self.mat_rigid_type

    # ------------------------------------------------------------------------------------------ device helpers on the host
    def pos_to_index(self, pos):
        return pos_to_index(pos, self.vdomain_start, self.grid_size)

    def flatten_grid_index(self, grid_index):
        return flatten_grid_index(grid_index, self.grid_num)

    def get_flatten_grid_index(self, pos):
        return self.flatten_grid_index(self.pos_to_index(pos))

    def for_all_neighbors(self, i, task, ret):
        """ps:259-269 for ONE particle on the host (a debugging helper; the sweeps do this on the device): calls
        ``task(i, j, ret)`` for every j of the 3^dim cells around i with |x_i - x_j| < support_radius, in the
        reference's order (cells x-major / z fastest, j ascending).  Needs a current grid."""
        x = self.pt.x
        xi = x[int(i)].cpu().numpy()
        cell_end = self.grid_particle_num.cpu().numpy()
        centre = self.pos_to_index(xi)
        zs = (0,) if self.dim == 2 else (-1, 0, 1)
        for ox in (-1, 0, 1):
            for oy in (-1, 0, 1):
                for oz in zs:
                    c = centre + np.array([ox, oy, oz])
                    if np.any(c < 0) or np.any(c[:self.dim] >= self.grid_num[:self.dim]):
                        continue                                  # the reference reads out of range here
                    g = int(self.flatten_grid_index(c if self.dim == 3 else np.array([c[0], c[1], 0])))   # grid_num[2] == 1 in 2D
                    j0 = int(cell_end[g - 1]) if g > 0 else 0
                    j1 = int(cell_end[g])
                    if j1 <= j0:
                        continue
                    xj = x[j0:j1].cpu().numpy()
                    d = xi[None, :] - xj
                    r = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
                    for k in np.nonzero(r < self.support_radius)[0]:
                        j = j0 + int(k)
                        if j != int(i):
                            task(int(i), j, ret)

    def add_particle(self, p, obj_id, x, v, density, pressure, material_id, material_type, is_dynamic, color):
        """ps:274-287: one particle (appended; ``p`` must be the next free index, as in the reference's callers)."""
        assert int(p) == self.particle_num[None], "particles are appended: p must equal particle_num"
        self._add_particles(obj_id, 1, np.asarray(x, dtype=np.float64).reshape(1, 3), np.asarray(v, dtype=np.float64).reshape(1, 3),
                            np.array([density], dtype=np.float64), np.array([pressure], dtype=np.float64),
                            np.array([material_id], dtype=np.int32), np.array([material_type], dtype=np.int32),
                            np.array([int(is_dynamic)], dtype=np.int32), np.asarray(color, dtype=np.float32).reshape(1, 3))

    # ------------------------------------------------------------------------------------------ viewer members (ps:380-407)
    def _vis_store(self, name, values):
        """Viewer members are kept keyed by id0, so that they follow their particle through the sorts."""
        n = max(self.engine.n, int(getattr(self, "global_particle_num", 0)))     # a slab rank holds GLOBAL creation indices
        shape = (n,) + tuple(values.shape[1:])
        buf = self._vis.get(name)
        if buf is None or tuple(buf.shape) != shape or buf.dtype != values.dtype:
            buf = self._torch.zeros(shape, dtype=values.dtype, device=self.engine.device)
            if name == "color":
                base = self._const_dev("color").to(values.dtype)
                buf[:len(base)].copy_(base[:len(buf)])
        buf[self.pt.id0.long()] = values
        self._vis[name] = buf

    def _vis_field(self, name):
        buf = self._vis.get(name)
        n = self.engine.n
        if buf is None:
            shape = (n,) if name == "val" else (n, 3)
            return self._torch.zeros(shape, dtype=self._torch.float32 if name != "val" else self.engine.real, device=self.engine.device)
        return buf[self.pt.id0.long()]

    def _shown(self):
        t = self.pt.mat_type
        real = t > 0
        return real | (t == self.mat_dummy_type) | (t == self.mat_rep_type) if self.show_bdy else real

    def copy2vis(self, w2s_ratio):
        """ps:380-385: positions scaled to the viewer's unit box, float32, for real (and, if shown, boundary) particles."""
        x = self.pt.x
        vis = self._vis_field("pos2vis").clone()
        shown = self._shown()
        vis[shown] = (x[shown] * float(w2s_ratio)).to(self._torch.float32)
        self._vis_store("pos2vis", vis)

    def v_maxmin(self, givenmax=-1, givenmin=-1, fixmax=0, fixmin=0):
        """ps:387-397: range of pt.val over FLOW particles (pt.val is set by solver.assign_value_color)."""
        t = self.pt.mat_type
        flow = (t == self.mat_fluid_type) | (t == self.mat_soil_type)
        self.vmax[None], self.vmin[None] = value_range(self.pt.val, flow, givenmax, givenmin, fixmax, fixmin)

    def set_color(self):
        """ps:399-407: pt.color = color_map((val - vmin) / (vmax - vmin)) for the particles of the colour group."""
        from .colormap import color_map
        t = self.pt.mat_type
        flow = (t == self.mat_fluid_type) | (t == self.mat_soil_type)
        real = t > 0
        if self.color_group == 0:
            sel = flow
        elif self.color_group == 1:
            sel = real
        else:
            sel = real | (t == self.mat_dummy_type) if self.show_bdy else real
        rng = self.vmax[None] - self.vmin[None]
        col = self.pt.color.to(self._torch.float32).clone()
        if rng != 0:
            col[sel] = color_map((self.pt.val[sel].double() - self.vmin[None]) / rng)
        self._vis_store("color", col)

    # ps:412-455: typed copies into caller-provided numpy arrays (``src_arr`` is a ps.pt member)
    def copy_to_numpy(self, np_arr, src_arr):
        np_arr[...] = src_arr.detach().cpu().numpy()

    def copy_to_numpy_vec(self, np_arr, src_arr):
        np_arr[...] = src_arr.detach().cpu().numpy()

    def copy_to_numpy_mat(self, np_arr, src_arr):
        np_arr[...] = src_arr.detach().cpu().numpy()

    def copy_to_numpy_vecxyz(self, np_arr_x, np_arr_y, np_arr_z, src_arr):
        a = src_arr.detach().cpu().numpy()
        np_arr_x[...], np_arr_y[...], np_arr_z[...] = a[:, 0], a[:, 1], a[:, 2]

    def copy_to_numpy_matxyz(self, np_arr_xx, np_arr_yy, np_arr_zz, np_arr_xy, np_arr_yz, np_arr_zx, src_arr):
        a = src_arr.detach().cpu().numpy()
        np_arr_xx[...], np_arr_yy[...], np_arr_zz[...] = a[:, 0, 0], a[:, 1, 1], a[:, 2, 2]
        np_arr_xy[...], np_arr_yz[...], np_arr_zx[...] = a[:, 0, 1], a[:, 1, 2], a[:, 2, 0]

    def copy_to_numpy_vecnorm(self, np_arr, src_arr):
        np_arr[...] = src_arr.detach().double().norm(dim=1).cpu().numpy()

    def copy_to_numpy_mathydro(self, np_arr, src_arr):
        a = src_arr.detach().cpu().numpy()
        np_arr[...] = (a[:, 0, 0] + a[:, 1, 1] + a[:, 2, 2]) / 3.0

    # ------------------------------------------------------------------------------------------ export
    def dump(self):
        """ps:459-545: (positions, data) dicts of float64 / int64 numpy arrays in current (sorted) order."""
        pt = self.pt
        f64 = lambda t: t.detach().to("cpu").double().numpy().copy()
        i64 = lambda t: t.detach().to("cpu").long().numpy().copy()
        x, v = f64(pt.x), f64(pt.v)
        n = len(x)
        try:
            s6 = f64(self.sym6("stress"))
            strain, strain_p, retmap = f64(pt.strain_equ), f64(pt.strain_equ_p), f64(pt.flag_retmap)
        except _lib.SphError:
            s6 = np.zeros((n, 6))
            strain, strain_p, retmap = np.zeros(n), np.zeros(n), np.zeros(n)
        return {"pos.x": x[:, 0].copy(), "pos.y": x[:, 1].copy(), "pos.z": x[:, 2].copy()}, {
            "id0": i64(pt.id0), "objId": i64(pt.obj_id), "material": i64(pt.mat_type), "density": f64(pt.density),
            "vel.x": v[:, 0].copy(), "vel.y": v[:, 1].copy(), "vel.z": v[:, 2].copy(),
            "vel.norm": np.sqrt((v * v).sum(axis=1)),
            "stress.xx": s6[:, 0].copy(), "stress.yy": s6[:, 1].copy(), "stress.zz": s6[:, 2].copy(),
            "stress.xy": s6[:, 3].copy(), "stress.yz": s6[:, 4].copy(), "stress.zx": s6[:, 5].copy(),
            "stress.hydro": (s6[:, 0] + s6[:, 1] + s6[:, 2]) / 3.0,
            "strain_equ": strain, "strain_equ_p": strain_p, "pressure": f64(pt.pressure), "plas_behav": retmap}
