"""MUISPHSolver (mirror of eng/solver_sph_muI.py:6-25; one_step muI:62-132 runs natively via sph_one_step)."""
import math

from .solver_sph_base import SPHBase


class MUISPHSolver(SPHBase):
    def __init__(self, particle_system):
        super().__init__(particle_system)
        print("μ(I) SPH starts to serve!")
        mat = self.ps.mat_soil[0]                  # only the first soil material is used (muI:12-17)
        self.density0 = mat["density0"]
        self.coh = mat["cohesion"]
        self.fric = mat["friction"] / 180 * math.pi
        self.E = mat["EYoungMod"]
        self.poi = mat["poison"]
        self.dila = mat["dilatancy"] / 180 * math.pi
        self.eta_0 = 0.0
        self.mu = math.tan(self.fric)
        self.vsound = 24.0                         # hard-coded in the reference (muI:23)
        self.vsound2 = self.vsound ** 2
        self._push_params(rho0=float(self.density0), coh=float(self.coh), fric=self.fric, E=float(self.E),
                          poi=float(self.poi), dila=self.dila, mu=self.mu, vsound=self.vsound)
        self.dt[None] = self.calc_dt_CFL(CFL_component=0.2, vsound=self.vsound, dt_min=self.dt_min)
