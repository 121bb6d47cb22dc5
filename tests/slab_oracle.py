"""Slab-engine adapter over the CPU oracle (TEST INFRASTRUCTURE): lets tisphi_b200.parallel.SlabDriver run under
gloo on CPU, so that the migration / halo protocol is checked bit-for-bit against a single-process oracle run."""
import ctypes as C

import numpy as np
import torch

from oracle import oracle as orc

STATE = list(orc.FIELDS) + list(orc.IFIELDS)


class OracleSlabEngine:
    def __init__(self, params, x, v, density, mat_type, id0, stress=None, obj_id=None, is_dynamic=None):
        self.P = params
        self.ti, self.xsph, self.solver = params.ti, params.xsph, params.solver
        self.o = self._alloc(len(x))
        o = self.o
        o.x[:], o.v[:], o.density[:], o.mat_type[:], o.id0[:] = x, v, density, mat_type, id0
        o.m_V[:] = params.m_V0
        o.mass[:] = params.m_V0 * np.asarray(density)
        o.obj_id[:] = 0 if obj_id is None else obj_id            # Oracle.__init__ defaults; static rigid blocks carry is_dynamic = 0
        o.is_dynamic[:] = 1 if is_dynamic is None else is_dynamic
        o.x0[:] = x
        if stress is not None:
            o.stress[:] = stress
        self.state_fields = STATE
        if self.solver == 1:
            self._phase = {0: ["v_tmp", "density_tmp", "pressure"], 1: []}
            self.deriv_fields = ["d_density", "d_vel"]
        elif self.solver == 2:
            self._phase = {0: ["stress_tmp", "pressure"], 1: ["v_tmp", "density_tmp", "stress_tmp"], 2: []}
            self.deriv_fields = ["d_density", "d_vel"]
        else:
            self._phase = {0: ["stress_tmp"], 1: ["v_tmp", "density_tmp", "stress_tmp"], 2: []}
            self.deriv_fields = ["d_density", "d_vel", "d_stress"]
        self.needs_final_ghosts = bool(self.xsph) or self.solver == 2
        self.post_fields = ["x"] if (self.solver == 2 and self.xsph) else []
        self.owned = None

    def _alloc(self, n):
        o = orc.Oracle.__new__(orc.Oracle)
        o.L, o.P, o.n = orc.lib(), self.P, int(n)
        o.h = o.L.orc_create(C.byref(self.P), max(o.n, 1))
        o._views()
        return o

    @property
    def n(self):
        return self.o.n

    def _arr(self, name):
        return getattr(self.o, name)

    def phase_fields(self, p):
        return list(self._phase[p])

    def new_buffer(self, nbytes):
        return torch.zeros(int(nbytes), dtype=torch.uint8)

    def new_counts(self):
        return torch.zeros(2, dtype=torch.int64)

    def _row_bytes(self, name):
        a = self._arr(name)
        return a.dtype.itemsize * (1 if a.ndim == 1 else a.shape[1])

    def message_bytes(self, fields, count):
        return sum(self._row_bytes(f) * int(count) for f in fields)

    def pack(self, fields, first, count, buf):
        b, off = buf.numpy(), 0
        for f in fields:
            raw = np.ascontiguousarray(self._arr(f)[first:first + count]).view(np.uint8).reshape(-1)
            b[off:off + raw.size] = raw
            off += raw.size

    def _unpack_into(self, o, fields, first, count, buf):
        b, off = buf.numpy(), 0
        for f in fields:
            a = getattr(o, f)
            nb = self._row_bytes(f) * count
            a[first:first + count] = b[off:off + nb].view(a.dtype).reshape((count,) + a.shape[1:])
            off += nb

    def unpack(self, fields, first, count, buf):
        self._unpack_into(self.o, fields, first, count, buf)

    def select_columns(self, which, first, count, cx_lo, cx_hi):
        x = self.o.x[first:first + count, 0]
        cx = ((x - self.P.vstart[0]) / self.P.grid_size).astype(np.int64)           # ps:216-218
        if not hasattr(self, "_sel"):
            self._sel = {}
        self._sel[which] = first + np.nonzero((cx >= cx_lo) & (cx <= cx_hi))[0]       # stable: previous order kept

    def select_counts(self):
        sel = getattr(self, "_sel", {})
        out = (len(sel.get(0, ())), len(sel.get(1, ())))
        return out

    def pack_selected(self, which, fields, count, buf):
        idx = self._sel[which]
        assert len(idx) == count
        b, off = buf.numpy(), 0
        for f in fields:
            raw = np.ascontiguousarray(self._arr(f)[idx]).view(np.uint8).reshape(-1)
            b[off:off + raw.size] = raw
            off += raw.size

    def replace(self, keep_first, keep_count, left, n_left, right, n_right):
        self._sel = {}
        new = self._alloc(n_left + keep_count + n_right)
        for f in STATE:
            getattr(new, f)[n_left:n_left + keep_count] = self._arr(f)[keep_first:keep_first + keep_count]
        if n_left:
            self._unpack_into(new, STATE, 0, n_left, left)
        if n_right:
            self._unpack_into(new, STATE, n_left + keep_count, n_right, right)
        self.o = new

    def column_starts(self, cols):
        gn = self.P.gn
        nyz = int(gn[1]) * (int(gn[2]) if self.P.dim == 3 else 1)
        out = []
        for cx in cols:
            if cx <= 0:
                out.append(0)
            elif cx >= gn[0]:
                out.append(self.o.n)
            else:
                out.append(int(self.o.cell_end[cx * nyz - 1]))
        return out

    def set_owned_columns(self, a, b):
        self.owned = (a, b)          # the oracle computes ghosts too; their results are overwritten by the refreshes

    def grid_build(self):
        assert self.o.grid_build() == 0

    def calc_kernel_corr(self):
        self.o.calc_kernel_corr()

    def init_real2tmp(self):
        self.o.L.orc_init_real2tmp(self.o.h)

    def num_phases(self):
        return 2 if self.solver == 1 else 3

    def one_step_phase(self, p):
        self.o.L.orc_one_step_phase(self.o.h, int(p))

    def advect(self, kind, m):
        self.o.L.orc_advect(self.o.h, int(kind), int(m))

    def advect_pos(self):
        self.o.L.orc_advect_pos(self.o.h)

    def post_step(self):
        self.o.L.orc_post_step(self.o.h)

    def enforce_boundary(self):
        self.o.L.orc_enforce_boundary(self.o.h)
