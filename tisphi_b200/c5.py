"""BASELINE.json config C5: synthetic 3D uniform box, neighbour search + density sweep.

A cubic box of n_side = round(N^(1/3)) lattice points per side, spacing d = 1, positions = lattice + U(-0.25 d, 0.25 d)
jitter (numpy default_rng(1234)), kappa * kh = 3 => support = cell edge = 3 d, the grid padded by one cell, uniform
mass.  One "update" of this workload = grid build (cell ids, counting sort, reorder) + neighbour count (int32) +
density sum  sum_j m_j W_ij  per particle -- the bare ``for_all_neighbors`` iteration of the reference
(eng/particle_system.py:216-269) with the density task of eng/solver_sph_wc.py:30-31, through the C ABI
(sph_grid_build, sph_density_sweep: count and density in one walk of the neighbours).
"""
import math

import numpy as np

from . import _lib


def box_positions(n_target, seed=1234, d=1.0):
    n_side = max(1, int(round(n_target ** (1.0 / 3.0))))
    ax = (np.arange(n_side, dtype=np.float64) + 0.5) * d
    x = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1).reshape(-1, 3)
    x += np.random.default_rng(seed).uniform(-0.25 * d, 0.25 * d, size=x.shape)
    return x, n_side


def box_params(n_side, precision="f32", d=1.0, fast=1):
    """SphParams of the box: h = 1.5 d, support = grid_size = 3 d, one cell of padding on every side."""
    P = _lib.SphParams()
    P.dim, P.kernel, P.kcorr, P.ti, P.xsph, P.solver = 3, 1, 0, 1, 0, _lib.SOLVER_WC
    P.precision = _lib.PREC_MIXED if precision in ("f32", "mixed") else _lib.PREC_F64
    P.wc_fresh, P.fast = 0, int(fast)                           # fast: the sweep walks the per-step neighbour bit masks
    gs = 3.0 * d
    cells = int(math.ceil(n_side * d / gs)) + 2
    for a in range(3):
        P.gn[a] = cells
        P.vstart[a] = -gs
        P.g[a] = 0.0
    P.h, P.support, P.grid_size, P.m_V0, P.eps = 1.5 * d, 3.0 * d, gs, d ** 3, 1e-8
    P.dt, P.rho0, P.visc, P.stiff, P.gamma_, P.vsound = 1e-4, 1.0, 0.0, 1.0, 7.0, 60.0
    return P


class UniformBox:
    """Engine + particles of one C5 instance on one device."""

    def __init__(self, n_target, device="cuda:0", precision="f32", seed=1234, fast=1):
        import torch
        self.torch = torch
        self.x, self.n_side = box_positions(n_target, seed)
        self.n = len(self.x)
        self.params = box_params(self.n_side, precision, fast=fast)
        self.engine = _lib.Engine(self.params, self.n, device=device)
        self.engine.add_particles(self.x, np.zeros_like(self.x), np.ones(self.n), np.ones(self.n, dtype=np.int32))
        self.count = torch.empty(self.n, dtype=torch.int32, device=self.engine.device)
        self.rho = torch.empty(self.n, dtype=self.engine.real, device=self.engine.device)

    def sweep(self):
        """grid build + neighbour count + density sum, enqueued on the engine's stream."""
        e = self.engine
        e.call("sph_grid_build")
        e.call("sph_density_sweep", self.count.data_ptr(), self.rho.data_ptr())

    def algorithmic_bytes(self):
        cells = self.params.gn[0] * self.params.gn[1] * self.params.gn[2]
        return 72 * self.n + 12 * cells                         # sort + density sweep
