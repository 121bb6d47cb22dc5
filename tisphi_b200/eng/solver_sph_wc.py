"""WCSPHSolver (mirror of eng/solver_sph_wc.py:6-21; one_step wc:82-126 runs natively via sph_one_step)."""
from .solver_sph_base import SPHBase


class WCSPHSolver(SPHBase):
    def __init__(self, particle_system):
        super().__init__(particle_system)
        print("WCSPH starts to serve!")
        mat = self.ps.mat_fluid[0]                 # only the first fluid material is used (wc:12-15)
        self.density0 = mat["density0"]
        self.viscosity = mat["viscosity"]
        self.stiffness = mat["stiffness"]
        self.exponent = mat["exponent"]
        self.vsound = 60                           # hard-coded in the reference (wc:17)
        self._push_params(rho0=float(self.density0), visc=float(self.viscosity), stiff=float(self.stiffness),
                          gamma_=float(self.exponent), vsound=float(self.vsound))
        self.dt[None] = self.calc_dt_CFL(CFL_component=0.2, vsound=self.vsound, dt_min=self.dt_min)
