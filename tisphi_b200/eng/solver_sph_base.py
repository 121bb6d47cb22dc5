"""SPHBase: step skeleton and time integrators (host orchestration only; every kernel is native).

Mirror of the reference's eng/solver_sph_base.py: step base:41-51, substep base:53-61, init_real2tmp base:67-74,
advect_SE/LF_half/LF base:79-114, substep_LF base:116-120, RK4 base:134-180, calc_dt_CFL base:209-212,
advect_pos base:228-238, advect_something base:244-247, init_stress base:249-260, calc_kernel_corr base:363-368.
"""
import math

import numpy as np

from .particle_system import ParticleSystem, _Scalar


class SPHBase:
    def __init__(self, particle_system: ParticleSystem):
        self.ps = particle_system
        cfg = self.ps.cfg
        self.g = np.array(cfg.get_cfg("gravitation"), dtype=np.float64)
        self.flagKernel = cfg.get_cfg("kernel")
        self.flagKernelCorr = cfg.get_cfg("kernelCorrection")
        self.flagTI = cfg.get_cfg("timeIntegration")
        self.flagXSPH = cfg.get_cfg("xsph")
        self.dt_min = cfg.get_cfg("timeStepSizeMin")
        self.dt = _Scalar(self.dt_min, on_set=self._push_dt)
        self.I = np.eye(self.ps.dim)
        self.I3 = np.eye(3)
        self.epsilon = 1e-8
        self.alert_ratio = 0.01
        self._eng = self.ps.engine
        self.init_rigid_body()                                                       # base:28

    # -------------------------------------------------------------------------------------- parameters
    def _push_dt(self, value):
        self.ps.params.dt = float(value)
        self._eng.set_params()

    def _push_params(self, **kw):
        for k, v in kw.items():
            setattr(self.ps.params, k, v)
        self._eng.set_params()

    # -------------------------------------------------------------------------------------- time integration
    def step(self):
        """base:41-51.  solve_rigid_body / enforce_boundary are no-ops for the supported scenes (no dynamic rigid)."""
        self.ps.initialize_particle_system()
        self.calc_kernel_corr()
        self.init_real2tmp()
        self.substep()
        self.advect_pos()
        self.advect_something()
        self.solve_rigid_body()
        self.enforce_boundary()

    def run_steps(self, n):
        """n x step() enqueued by one native call (no Python between kernels)."""
        if self.flagTI == 3:
            self.substep_VV()
        self._eng.call("sph_step", int(n))

    def substep(self):
        if self.flagTI == 1:
            self.substep_SE()
        elif self.flagTI == 2:
            self.substep_LF()
        elif self.flagTI == 3:
            self.substep_VV()
        elif self.flagTI == 4:
            self.substep_RK()

    def one_step(self):
        self._eng.call("sph_one_step")

    def init_real2tmp(self):
        self._eng.call("sph_init_real2tmp")

    def advect_SE(self):
        self._eng.call("sph_advect", 0, 0)

    def substep_SE(self):
        self.one_step()
        self.advect_SE()

    def advect_LF_half(self):
        self._eng.call("sph_advect", 1, 0)

    def advect_LF(self):
        self._eng.call("sph_advect", 0, 0)

    def substep_LF(self):
        self.one_step()
        self.advect_LF_half()
        self.one_step()
        self.advect_LF()

    def substep_VV(self):
        # the reference calls an undefined advect_VV_half here (base:126-130)
        raise AttributeError("'%s' object has no attribute 'advect_VV_half'" % type(self).__name__)

    def advect_RK_4(self):
        self._eng.call("sph_advect", 2, 0)

    def init_RK(self):
        self._eng.call("sph_advect", 3, 0)

    def update_RK(self, m):
        self._eng.call("sph_advect", 4, int(m))

    def advect_RK(self):
        self._eng.call("sph_advect", 5, 0)

    def substep_RK(self):
        self.init_RK()
        for stage, m in enumerate((1, 2, 2, 1)):
            self.one_step()
            self.update_RK(m)
            if stage < 3:
                self.advect_RK_4()
        self.advect_RK()

    # -------------------------------------------------------------------------------------- assist
    def calc_dt_CFL(self, CFL_component, vsound, dt_min):
        """base:209-212 with Taichi's float modulo ``a - floor(a / b) * b``."""
        dt = CFL_component * self.ps.smoothing_len / vsound
        return max(dt_min, dt - (dt - math.floor(dt / dt_min) * dt_min))

    def advect_pos(self):
        self._eng.call("sph_advect_pos")

    def advect_something(self):
        self._eng.call("sph_post_step")

    def advect_something_func(self, i):
        raise NotImplementedError("per-particle hooks run inside the native post-step kernel")

    def init_stress(self, density0=None, fric=None):
        ymax = getattr(self.ps, "slab_soil_ymax", None)     # one slab of a multi-GPU scene: the GLOBAL soil top
        if ymax is not None:
            self._eng.call("sph_init_stress_ymax", float(ymax))
        else:
            self._eng.call("sph_init_stress")

    def calc_kernel_corr(self):
        self._eng.call("sph_calc_kernel_corr")

    def calc_CSPM_f(self):
        self._eng.call("sph_calc_kernel_corr")

    def calc_CSPM_L(self):
        self._eng.call("sph_calc_kernel_corr")

    def init_rigid_body(self):
        """base:467-470: the rest centre of mass of every dynamic rigid body (-> ps.rigid_rest_cm[object id])."""
        ids = getattr(self.ps, "rigid_dynamic_ids", [])
        if not ids:
            return
        self._eng.call("sph_init_rigid_body")
        out = np.zeros((len(ids), 3), dtype=np.float64)
        self._eng.call("sph_rigid_rest_cm", out.ctypes.data)
        self.ps.rigid_rest_cm = {oid: out[k].copy() for k, oid in enumerate(ids)}

    def solve_rigid_body(self):
        """base:472-499: shape matching of every dynamic rigid body (a no-op without one)."""
        self._eng.call("sph_solve_rigid_body")

    def enforce_boundary(self):
        """base:525-601: with ``boundary == 1`` flow particles are clamped into the domain box and reflected."""
        self._eng.call("sph_enforce_boundary")

    def assign_value_color(self):
        """base:721-789: pt.val := the scalar selected by ``colorTitle`` for real (and, if shown, dummy) particles."""
        ps, pt = self.ps, self.ps.pt
        f = _COLOR_VALUES.get(int(ps.color_title))
        if f is None:
            return None
        t = pt.mat_type
        shown = (t > 0) | (t == ps.mat_dummy_type) if ps.show_bdy else (t > 0)
        val = ps._vis_field("val").clone()
        new = f(pt, ps).to(val.dtype)
        val[shown] = new[shown]
        ps._vis_store("val", val)
        return None

    def init_pressure(self, density0):
        """base:263-271: hydrostatic pressure below the highest fluid particle (not called by the reference's solvers)."""
        pt = self.ps.pt
        fluid = pt.mat_type == self.ps.mat_fluid_type
        if not bool(fluid.any()):
            return None
        y = pt.x[:, 1]
        ymax = y[fluid].max()
        p = pt.pressure
        p[fluid] = (-float(density0) * float(self.g[1]) * (ymax - y[fluid])).to(p.dtype)
        return None


def _norm(v):
    return v.double().norm(dim=1)


# colorTitle -> value (base:725-787); signs as in the reference (compression positive for 32, 52, 53, 57)
_COLOR_VALUES = {
    1: lambda pt, ps: pt.id0.double(), 2: lambda pt, ps: pt.density, 21: lambda pt, ps: pt.d_density,
    3: lambda pt, ps: _norm(pt.v), 31: lambda pt, ps: pt.v[:, 0], 32: lambda pt, ps: -pt.v[:, 1], 33: lambda pt, ps: pt.v[:, 2],
    34: lambda pt, ps: _norm(pt.d_vel), 35: lambda pt, ps: pt.d_vel[:, 0], 36: lambda pt, ps: pt.d_vel[:, 1], 37: lambda pt, ps: pt.d_vel[:, 2],
    4: lambda pt, ps: _norm(pt.x), 41: lambda pt, ps: pt.x[:, 0], 42: lambda pt, ps: pt.x[:, 1], 43: lambda pt, ps: pt.x[:, 2],
    44: lambda pt, ps: _norm(pt.x - pt.x0),
    51: lambda pt, ps: pt.stress[:, 0, 0], 52: lambda pt, ps: -pt.stress[:, 1, 1], 53: lambda pt, ps: -pt.stress[:, 2, 2],
    54: lambda pt, ps: pt.stress[:, 0, 1], 55: lambda pt, ps: pt.stress[:, 1, 2], 56: lambda pt, ps: pt.stress[:, 2, 0],
    57: lambda pt, ps: -(pt.stress[:, 0, 0] + pt.stress[:, 1, 1] + pt.stress[:, 2, 2]) / 3.0,
    61: lambda pt, ps: pt.strain_equ, 62: lambda pt, ps: pt.strain_equ_p, 7: lambda pt, ps: pt.pressure,
    8: lambda pt, ps: pt.flag_retmap.double(), 100: lambda pt, ps: pt.grid_ids.double(),
    101: lambda pt, ps: ps._torch.rand(ps.engine.n, device=ps.engine.device, dtype=ps._torch.float64),
    102: lambda pt, ps: pt.CSPM_L[:, 0, 0], 103: lambda pt, ps: _norm(pt.v_tmp),
}
