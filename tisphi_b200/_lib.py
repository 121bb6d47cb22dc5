"""ctypes binding of libtisphi_b200.so (C ABI declared in include/tisphi_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this module raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TISPHI_B200_LIB") or os.path.join(HERE, "libtisphi_b200.so")   # override: A/B runs of two builds

PREC_F64, PREC_MIXED = 0, 1
SOLVER_WC, SOLVER_MUI, SOLVER_DP = 1, 2, 3

FIELDS = ["X", "V", "MASS", "M_V", "DENSITY", "DENSITY_TMP", "V_TMP", "PRESSURE", "MAT_TYPE", "ID0", "GRID_IDS",
          "STRESS", "STRESS_TMP", "STRAIN_EQU", "STRAIN_EQU_P", "FLAG_RETMAP", "CSPM_F", "CSPM_L", "D_DENSITY", "D_VEL",
          "D_STRESS", "V_GRAD", "D_STRAIN_EQU", "D_STRAIN_EQU_P", "D_DENSITY_RK", "D_VEL_RK", "D_STRESS_RK", "XS",
          "CELL_END", "CELL_COUNT", "ID_NEW", "PK4"]
FIELD_ID = {name: k for k, name in enumerate(FIELDS)}

# every symbol include/tisphi_b200.h declares (tests check that the library exports all of them)
SYMBOLS = ["sph_arena_bytes", "sph_create", "sph_destroy", "sph_last_error", "sph_set_params", "sph_field_info",
           "sph_add_particles", "sph_num_particles", "sph_clear_particles", "sph_real_bytes", "sph_read_state", "sph_read_state_async",
           "sph_synchronize", "sph_grid_build",
           "sph_calc_kernel_corr", "sph_calc_kernel_corr_deferred", "sph_init_real2tmp", "sph_one_step", "sph_advect", "sph_advect_pos", "sph_post_step",
           "sph_init_stress", "sph_init_stress_ymax", "sph_enforce_boundary", "sph_set_rigid_bodies", "sph_init_rigid_body", "sph_solve_rigid_body", "sph_rigid_rest_cm", "sph_step", "sph_neighbor_count", "sph_neighbor_count_masks", "sph_density_sum", "sph_density_sweep", "sph_read_bad_cells", "sph_read_flagged_cells",
           "sph_launch_count", "sph_num_phases", "sph_one_step_phase", "sph_set_owned_columns",
           "sph_column_starts", "sph_state_fields", "sph_message_bytes", "sph_pack_fields", "sph_unpack_fields",
           "sph_replace_particles", "sph_select_columns", "sph_select_counts", "sph_pack_selected", "sph_profile_enable", "sph_profile_num_kernels",
           "sph_profile_name", "sph_profile_read", "sph_params_size",
           "sph_slab_inbox_bytes", "sph_slab_init", "sph_slab_connect", "sph_slab_sync", "sph_slab_epoch",
           "sph_ipc_alloc", "sph_ipc_free", "sph_ipc_get_handle", "sph_ipc_open", "sph_ipc_close"]


class SphParams(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("dim", "kernel", "kcorr", "ti", "xsph", "solver", "precision", "wc_fresh")] + \
               [("gn", C.c_int32 * 3), ("fast", C.c_int32)] + \
               [("h", C.c_double), ("support", C.c_double), ("grid_size", C.c_double), ("vstart", C.c_double * 3),
                ("m_V0", C.c_double), ("g", C.c_double * 3), ("dt", C.c_double), ("eps", C.c_double)] + \
               [(k, C.c_double) for k in ("rho0", "visc", "stiff", "gamma_", "coh", "fric", "E", "poi", "dila",
                                          "vsound", "mu", "alpha", "kc", "G", "K", "eps_f")] + \
               [("boundary", C.c_int32), ("pad_", C.c_int32), ("radius", C.c_double), ("dstart", C.c_double * 3),
                ("dend", C.c_double * 3)]


class SphError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (build it first with ``python -m tisphi_b200._build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SphError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.sph_arena_bytes.restype, L.sph_arena_bytes.argtypes = i64, [C.POINTER(SphParams), i64]
    L.sph_create.restype, L.sph_create.argtypes = vp, [C.POINTER(SphParams), i64, vp, i64, vp]
    L.sph_destroy.restype, L.sph_destroy.argtypes = None, [vp]
    L.sph_last_error.restype, L.sph_last_error.argtypes = C.c_char_p, [vp]
    L.sph_set_params.restype, L.sph_set_params.argtypes = C.c_int, [vp, C.POINTER(SphParams)]
    L.sph_field_info.restype = C.c_int
    L.sph_field_info.argtypes = [vp, C.c_int, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.sph_add_particles.restype, L.sph_add_particles.argtypes = C.c_int, [vp, i64, vp, vp, vp, vp]
    L.sph_num_particles.restype, L.sph_num_particles.argtypes = i64, [vp]
    L.sph_clear_particles.restype, L.sph_clear_particles.argtypes = C.c_int, [vp]
    L.sph_real_bytes.restype, L.sph_real_bytes.argtypes = i32, [vp]
    L.sph_read_state.restype, L.sph_read_state.argtypes = C.c_int, [vp, vp, vp, vp, vp, vp]
    L.sph_read_state_async.restype, L.sph_read_state_async.argtypes = C.c_int, [vp, vp, vp, vp, vp, vp]
    L.sph_synchronize.restype, L.sph_synchronize.argtypes = C.c_int, [vp]
    for fn in ("sph_grid_build", "sph_calc_kernel_corr", "sph_calc_kernel_corr_deferred", "sph_init_real2tmp", "sph_one_step", "sph_advect_pos",
               "sph_post_step", "sph_init_stress", "sph_enforce_boundary", "sph_init_rigid_body", "sph_solve_rigid_body"):
        getattr(L, fn).restype, getattr(L, fn).argtypes = C.c_int, [vp]
    L.sph_advect.restype, L.sph_advect.argtypes = C.c_int, [vp, C.c_int, C.c_int]
    L.sph_set_rigid_bodies.restype, L.sph_set_rigid_bodies.argtypes = C.c_int, [vp, i64, vp, vp, i32]
    L.sph_rigid_rest_cm.restype, L.sph_rigid_rest_cm.argtypes = C.c_int, [vp, vp]
    L.sph_init_stress_ymax.restype, L.sph_init_stress_ymax.argtypes = C.c_int, [vp, C.c_double]
    L.sph_step.restype, L.sph_step.argtypes = C.c_int, [vp, C.c_int]
    L.sph_neighbor_count.restype, L.sph_neighbor_count.argtypes = C.c_int, [vp, vp]
    L.sph_density_sum.restype, L.sph_density_sum.argtypes = C.c_int, [vp, vp]
    L.sph_density_sweep.restype, L.sph_density_sweep.argtypes = C.c_int, [vp, vp, vp]
    L.sph_neighbor_count_masks.restype, L.sph_neighbor_count_masks.argtypes = C.c_int, [vp, vp]
    L.sph_read_bad_cells.restype, L.sph_read_bad_cells.argtypes = i64, [vp]
    L.sph_launch_count.restype, L.sph_launch_count.argtypes = i64, [vp]
    L.sph_read_flagged_cells.restype, L.sph_read_flagged_cells.argtypes = i64, [vp]
    L.sph_num_phases.restype, L.sph_num_phases.argtypes = C.c_int, [vp]
    L.sph_one_step_phase.restype, L.sph_one_step_phase.argtypes = C.c_int, [vp, C.c_int]
    L.sph_set_owned_columns.restype, L.sph_set_owned_columns.argtypes = C.c_int, [vp, i32, i32]
    L.sph_column_starts.restype, L.sph_column_starts.argtypes = C.c_int, [vp, i32, C.POINTER(i32), C.POINTER(i64)]
    L.sph_state_fields.restype, L.sph_state_fields.argtypes = C.c_int, [vp, C.POINTER(i32), i32]
    L.sph_message_bytes.restype, L.sph_message_bytes.argtypes = i64, [vp, i32, C.POINTER(i32), i64]
    L.sph_pack_fields.restype, L.sph_pack_fields.argtypes = C.c_int, [vp, i32, C.POINTER(i32), i64, i64, vp]
    L.sph_unpack_fields.restype, L.sph_unpack_fields.argtypes = C.c_int, [vp, i32, C.POINTER(i32), i64, i64, vp]
    L.sph_select_columns.restype, L.sph_select_columns.argtypes = C.c_int, [vp, i32, i64, i64, i32, i32]
    L.sph_select_counts.restype, L.sph_select_counts.argtypes = C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]
    L.sph_pack_selected.restype, L.sph_pack_selected.argtypes = C.c_int, [vp, i32, i32, C.POINTER(i32), i64, vp]
    L.sph_replace_particles.restype, L.sph_replace_particles.argtypes = C.c_int, [vp, i64, i64, vp, i64, vp, i64]
    L.sph_profile_enable.restype, L.sph_profile_enable.argtypes = C.c_int, [vp, C.c_int]
    L.sph_profile_num_kernels.restype, L.sph_profile_num_kernels.argtypes = C.c_int, []
    L.sph_profile_name.restype, L.sph_profile_name.argtypes = C.c_char_p, [C.c_int]
    L.sph_profile_read.restype, L.sph_profile_read.argtypes = C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    L.sph_params_size.restype, L.sph_params_size.argtypes = i64, []
    L.sph_slab_inbox_bytes.restype, L.sph_slab_inbox_bytes.argtypes = i64, [vp, i64]
    L.sph_slab_init.restype, L.sph_slab_init.argtypes = C.c_int, [vp, i32, i32, i32, i32, i64, vp, i64]
    L.sph_slab_connect.restype, L.sph_slab_connect.argtypes = C.c_int, [vp, vp, vp]
    L.sph_slab_sync.restype = C.c_int
    L.sph_slab_sync.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.sph_slab_epoch.restype, L.sph_slab_epoch.argtypes = i64, [vp]
    L.sph_ipc_alloc.restype, L.sph_ipc_alloc.argtypes = vp, [i64]
    L.sph_ipc_free.restype, L.sph_ipc_free.argtypes = None, [vp]
    L.sph_ipc_get_handle.restype, L.sph_ipc_get_handle.argtypes = C.c_int, [vp, vp]
    L.sph_ipc_open.restype, L.sph_ipc_open.argtypes = vp, [vp]
    L.sph_ipc_close.restype, L.sph_ipc_close.argtypes = None, [vp]
    if L.sph_params_size() != C.sizeof(SphParams):
        raise SphError('SphParams layout mismatch between tisphi_b200/_lib.py and include/tisphi_b200.h')
    _lib = L
    return L


class Engine:
    """One SphCtx bound to a caller-owned torch arena on one CUDA device/stream."""

    def __init__(self, params, n_max, device="cuda:0", stream=None):
        import torch
        if not torch.cuda.is_available():
            raise SphError("tisphi_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.L = load()
        self.torch = torch
        self.device = torch.device(device)
        self.params = params
        self.n_max = int(n_max)
        nbytes = self.L.sph_arena_bytes(C.byref(params), self.n_max)
        with torch.cuda.device(self.device):
            self.arena = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = self.arena.data_ptr()
            self._pad = (-base) % 256
            self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
            self.h = self.L.sph_create(C.byref(params), self.n_max, base + self._pad, nbytes, self.stream.cuda_stream)
        if not self.h:
            raise SphError("sph_create failed (bad parameters or arena)")
        self.real = torch.float64 if params.precision == PREC_F64 else torch.float32

    def close(self):
        if getattr(self, "h", None):
            self.L.sph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise SphError(f"libtisphi_b200 error {rc}: {self.L.sph_last_error(self.h).decode()}")

    def call(self, name, *args):
        with self.torch.cuda.device(self.device):
            self.check(getattr(self.L, name)(self.h, *args))

    @property
    def n(self):
        return int(self.L.sph_num_particles(self.h))

    def set_params(self):
        self.check(self.L.sph_set_params(self.h, C.byref(self.params)))

    def field(self, name, count=None):
        """torch view (no copy) of a particle member in its current buffer."""
        torch = self.torch
        off, nc, stride, kind = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
        rc = self.L.sph_field_info(self.h, FIELD_ID[name], C.byref(off), C.byref(nc), C.byref(stride), C.byref(kind))
        if rc != 0:
            raise SphError(f"field {name} is not allocated for this solver configuration")
        dt = {0: torch.float64, 1: self.real, 2: torch.int32}[kind.value]
        es = torch.empty((), dtype=dt).element_size()
        if count is None:
            count = self.n
        start = self._pad + off.value
        if count == 0:
            return torch.empty((0,) if nc.value == 1 else (0, nc.value), dtype=dt, device=self.device)
        nbytes = ((count - 1) * stride.value + nc.value) * es
        flat = self.arena[start:start + nbytes].view(dt)
        if nc.value == 1:
            return flat.as_strided((count,), (stride.value,))
        return flat.as_strided((count, nc.value), (stride.value, 1))

    def profile(self, on=True):
        self.check(self.L.sph_profile_enable(self.h, int(on)))

    def profile_read(self):
        """{kernel class: (total ms, launches)} since the last read (CUDA events on the engine's stream)."""
        k = self.L.sph_profile_num_kernels()
        ms, cnt = (C.c_double * k)(), (C.c_int64 * k)()
        self.check(self.L.sph_profile_read(self.h, ms, cnt))
        return {self.L.sph_profile_name(i).decode(): (ms[i], cnt[i]) for i in range(k) if cnt[i] > 0}

    def add_particles(self, x, v, density, mat_type):
        import numpy as np
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        density = np.ascontiguousarray(density, dtype=np.float64)
        mat_type = np.ascontiguousarray(mat_type, dtype=np.int32)
        n = len(density)
        assert x.shape == (n, 3) and v.shape == (n, 3) and mat_type.shape == (n,)
        self.call("sph_add_particles", n, x.ctypes.data, v.ctypes.data, density.ctypes.data, mat_type.ctypes.data)
        self.torch.cuda.synchronize(self.device)          # host buffers may be pageable: keep them alive until copied
