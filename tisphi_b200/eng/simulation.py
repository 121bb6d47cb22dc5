"""Assembly (mirror of eng/simulation.py:7-22)."""
from .configer_builder import SimConfiger
from .particle_system import ParticleSystem
from .solver_sph_wc import WCSPHSolver
from .solver_sph_muI import MUISPHSolver
from .solver_sph_dp import DPSPHSolver

_SOLVERS = {1: WCSPHSolver, 2: MUISPHSolver, 3: DPSPHSolver}


class Simulation:
    def __init__(self, config: SimConfiger, device="cuda:0", slab=None) -> None:
        """``slab``: dict(rank, world, wall_weight, columns) for one rank of a multi-GPU run (tisphi_b200/parallel.py)."""
        self.cfg = config
        self.solver_type = self.cfg.get_cfg("simulationMethod")
        if self.solver_type not in _SOLVERS:      # the reference fails after building the particle system
            raise NotImplementedError(f"Solver type {self.solver_type} has not been implemented.")
        self.ps = ParticleSystem(self.cfg, device=device, slab=slab)
        self.solver = self.build_solver()

    def build_solver(self):
        try:
            cls = _SOLVERS[self.solver_type]
        except KeyError:
            raise NotImplementedError(f"Solver type {self.solver_type} has not been implemented.")
        return cls(self.ps)
