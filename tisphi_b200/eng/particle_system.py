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
from .particle_func import (add_boundary, add_cube, calc_cube_particle_num, calc_dummy_boundary, chk_block_in_domain,
                            count_boundary, get_material, set_material)


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
        if name in _CONST:
            idx = eng.field("ID0").long()
            return ps._const_dev(name)[idx]
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
        self.bodies = self.cfg.get_bodies()
        if self.bodies:
            raise NotImplementedError("mesh Bodies are out of scope of this engine (SURVEY 2.1)")
        dummy_particle_num = 0
        if self.flag_boundary in (self.bdy_dummy, self.bdy_dummy_rep):
            self.dummy_boundary = calc_dummy_boundary(self.dim, self.domain_start, self.domain_end, self.vdomain_start,
                                                      self.vdomain_end)
            dummy_particle_num = count_boundary(self.dummy_boundary, self.dim, self.particle_diameter)
            print("Dummy particle number: %d" % dummy_particle_num)
        if self.flag_boundary in (self.bdy_rep, self.bdy_dummy_rep, self.bdy_collision):
            raise NotImplementedError("boundary modes 1, 3 and 4 are out of scope of this engine (SURVEY 2.1)")
        self.particle_max_num = block_particle_num + dummy_particle_num
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
        print("Particle system construction complete!")

    # ------------------------------------------------------------------------------------------ construction
    def initialize_particles(self):
        for block in self.blocks:
            self.init_block(block)
        if self.flag_boundary in (self.bdy_dummy, self.bdy_dummy_rep):
            add_boundary(self, self.dummy_boundary, self.mat_dummy_type, color=[153, 153, 255])

    def init_block(self, block):
        mat = get_material(self, block["materialId"])
        mat_type = mat["matType"]
        if mat_type > 10:
            raise NotImplementedError("rigid blocks are out of scope of this engine (SURVEY 8f-2)")
        add_cube(self, object_id=block["objectId"], lower_corner=np.array(block["translation"]),
                 cube_size=np.array(block["size"]), velocity=block["velocity"], density=mat["density0"],
                 is_dynamic=True, color=np.array([ic / 255 for ic in mat["color"]], dtype=np.float32),
                 mat_id=block["materialId"], mat_type=mat_type)

    def init_body(self, body):
        raise NotImplementedError("mesh Bodies are out of scope of this engine (SURVEY 2.1)")

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
        self.engine = _lib.Engine(self.params, max(int(cap), 1), device=self._device)
        if len(mine):
            self.engine.add_particles(x[mine], v[mine], rho[mine], typ[mine])
            self.engine.field("ID0").copy_(self._torch.from_numpy(mine.astype(np.int32)).to(self.engine.device))
        self.particle_num[None] = self.engine.n
        print(f"slab rank {slab['rank']}/{slab['world']}: columns [{a}, {b}) of {n_cols}, {len(mine)} particles, capacity {cap}")

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

    # ------------------------------------------------------------------------------------------ export
    def v_maxmin(self):
        val = self.pt.v.norm(dim=1) if self.engine.n else self._torch.zeros(1)
        self.vmax[None] = float(val.max())
        self.vmin[None] = float(val.min())

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
