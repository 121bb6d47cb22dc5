"""DPSPHSolver (mirror of eng/solver_sph_dp.py:6-35; one_step dp:210-274 runs natively via sph_one_step)."""
import math

from .solver_sph_base import SPHBase


class DPSPHSolver(SPHBase):
    def __init__(self, particle_system):
        super().__init__(particle_system)
        print("Drucker-Prager SPH starts to serve!")
        mat = self.ps.mat_soil[0]                  # only the first soil material is used (dp:12-17)
        self.density0 = mat["density0"]
        self.coh = mat["cohesion"]
        self.fric = mat["friction"] / 180 * math.pi
        self.E = mat["EYoungMod"]
        self.poi = mat["poison"]
        self.dila = mat["dilatancy"] / 180 * math.pi
        self.vsound2 = self.E / self.density0
        self.vsound = math.sqrt(self.vsound2)
        self.eps_f = 1e-4
        t = math.tan(self.fric)
        self.alpha_fric = t / math.sqrt(9 + 12 * t ** 2)
        self.k_c = 3 * self.coh / math.sqrt(9 + 12 * t ** 2)
        self.G = self.E / (2 * (1 + self.poi))
        self.K = self.E / (3 * (1 - 2 * self.poi))
        self._push_params(rho0=float(self.density0), coh=float(self.coh), fric=self.fric, E=float(self.E),
                          poi=float(self.poi), dila=self.dila, mu=t, vsound=self.vsound, alpha=self.alpha_fric,
                          kc=self.k_c, G=self.G, K=self.K, eps_f=self.eps_f)
        self.dt[None] = self.calc_dt_CFL(CFL_component=0.2, vsound=self.vsound, dt_min=self.dt_min)
        self.init_stress(self.density0, self.fric)
