"""Spatial slab decomposition over the GPUs of one box: column partition, migration + halo exchange, phase-wise step.

The reference is single-device; this module is what lets one process per GPU own the x-columns
[a, b) of the GLOBAL uniform grid.  Properties the design relies on:

  * the flattened cell id is x-major (eng/particle_system.py:221-222): after the counting sort every column is ONE
    contiguous index range of every member array, so halo and migration messages are plain ranges;
  * interactions reach one cell (support radius == cell edge): one ghost column per side is enough, provided the
    ghosts are refreshed after every top-level loop of one_step (the phases of sph_one_step_phase);
  * the counting sort is stable: if arrivals from the lower-x neighbour are placed in front of the local particles
    and arrivals from the higher-x neighbour behind them, every rank's sorted order is the restriction of the
    single-GPU global order -- cell ids, order and every floating-point sum are bit-identical to a 1-GPU run.

Per step and rank (neighbours L = rank - 1, R = rank + 1):
  1. stable selection (no sort) of the owned particles whose NEW column is <= a, resp. >= b - 1
  2. send them to L / R (migrants + my boundary column in one message); what arrives holds the neighbour's
     boundary column (my ghosts) and its migrants into my slab                                       (1 exchange)
  3. ONE sort of [from L][mine, last step's ghosts dropped][from R]
  4. kernel correction, init_real2tmp, then for every phase of every one_step: run it on the owned columns,
     refresh the ghost columns with the members that phase wrote; integrator kernels run on ghosts too.
There is no collective in the step: only neighbour send/recv (NCCL over NVLink on GPUs, gloo in the CPU tests).

The driver below is engine-agnostic (duck-typed "slab engine"); CudaSlabEngine binds it to libtisphi_b200 through
the C ABI.  tests/ bind it to the CPU oracle to check the protocol against a single-process run under gloo.
"""
import ctypes as C

import numpy as np


# ---------------------------------------------------------------------------------------------- partition
def column_weights(x, mat_type, vstart_x, grid_size, n_cols, wall_weight=0.15):
    """Work estimate per x-column: flow particles weigh 1, wall particles ``wall_weight`` (two short sweeps)."""
    cx = ((np.asarray(x, dtype=np.float64) - vstart_x) / grid_size).astype(np.int64)      # ps:216-218, C cast
    cx = np.clip(cx, 0, n_cols - 1)
    w = np.where(np.asarray(mat_type) > 0, 1.0, wall_weight)
    return np.bincount(cx, weights=w, minlength=n_cols), cx


def partition_columns(weights, world):
    """Contiguous column ranges [(a_0, b_0), ...] with near-equal weight, every rank at least one column."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if world > n:
        raise ValueError(f"{world} ranks but only {n} grid columns")
    cum = np.concatenate([[0.0], np.cumsum(w)])
    bounds = [0]
    for k in range(1, world):
        target = cum[-1] * k / world
        c = int(np.searchsorted(cum, target, side="left"))
        if c > 0 and abs(cum[c - 1] - target) <= abs(cum[min(c, n)] - target):
            c -= 1
        c = max(c, bounds[-1] + 1)              # at least one column per rank
        c = min(c, n - (world - k))             # leave one for each rank still to come
        bounds.append(c)
    bounds.append(n)
    return [(bounds[k], bounds[k + 1]) for k in range(world)]


# ---------------------------------------------------------------------------------------------- the protocol
class SlabDriver:
    """Runs SPHBase.step (eng/solver_sph_base.py:41-51) on one slab.  ``eng`` is a slab engine (see module doc):

    n, ti, xsph, solver; grid_build(), column_starts(list) -> list, state_fields, phase_fields(phase), deriv_fields,
    select_columns(which, first, count, cx_lo, cx_hi), select_counts() -> (n0, n1), pack_selected(which, fields, count, buf),
    message_bytes(fields, count), new_buffer(nbytes), pack(fields, first, count, buf), unpack(fields, first, count, buf),
    replace(keep_first, keep_count, left_buf, n_left, right_buf, n_right), set_owned_columns(a, b),
    calc_kernel_corr(), init_real2tmp(), num_phases(), one_step_phase(p), advect(kind, m), advect_pos(), post_step(),
    enforce_boundary(),
    new_counts() -> int64 tensor[2] on the transport device.
    """

    def __init__(self, eng, columns, rank, world, n_grid_cols, group=None, check=False):
        import torch.distributed as dist
        self.dist = dist
        self.eng = eng
        self.a, self.b = int(columns[0]), int(columns[1])
        self.rank, self.world = rank, world
        self.left = rank - 1 if rank > 0 else None
        self.right = rank + 1 if rank < world - 1 else None
        self.n_cols = n_grid_cols
        self.group = group
        self.check = check
        self.own_first, self.own_count = 0, eng.n          # before the first step a rank holds its own particles only
        self.col_start = None                               # {column: first index} after the last sort
        self.ghost_l = (0, 0)                               # (first, count) of the ghost column a - 1
        self.ghost_r = (0, 0)
        self.send_l = (0, 0)                                # my column a      -> ghost column of L
        self.send_r = (0, 0)                                # my column b - 1  -> ghost column of R
        self.exchanges = 0
        self.bytes_sent = 0
        self.t_compute = 0.0
        self.t_comm = 0.0                                   # host seconds spent waiting in send/recv (incl. device drain)
        eng.set_owned_columns(self.a, self.b)

    def reset(self):
        """The engine's particle set was replaced from outside (fresh upload): it holds owned particles only, unsorted."""
        self.own_first, self.own_count = 0, self.eng.n
        self.col_start = None
        self.ghost_l = self.ghost_r = self.send_l = self.send_r = (0, 0)

    # ---- transport ------------------------------------------------------------------------------------
    def _sendrecv(self, send_l, send_r, recv_l, recv_r):
        """Neighbour exchange of already packed buffers (None = nothing in that direction)."""
        dist = self.dist
        ops = []
        if self.left is not None:
            if recv_l is not None and recv_l.numel():
                ops.append(dist.P2POp(dist.irecv, recv_l, self.left, group=self.group))
            if send_l is not None and send_l.numel():
                ops.append(dist.P2POp(dist.isend, send_l, self.left, group=self.group))
                self.bytes_sent += send_l.numel() * send_l.element_size()
        if self.right is not None:
            if send_r is not None and send_r.numel():
                ops.append(dist.P2POp(dist.isend, send_r, self.right, group=self.group))
                self.bytes_sent += send_r.numel() * send_r.element_size()
            if recv_r is not None and recv_r.numel():
                ops.append(dist.P2POp(dist.irecv, recv_r, self.right, group=self.group))
        if ops:
            import time
            t0 = time.perf_counter()
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            self.t_comm += time.perf_counter() - t0
            self.exchanges += 1

    def _exchange_counts(self, n_to_l, n_to_r):
        e = self.eng
        out_l, out_r, in_l, in_r = e.new_counts(), e.new_counts(), e.new_counts(), e.new_counts()
        out_l[0], out_r[0] = n_to_l, n_to_r
        in_l.zero_()
        in_r.zero_()
        self._sendrecv(out_l, out_r, in_l, in_r)
        return int(in_l[0]), int(in_r[0])

    # ---- step 1-3: migration + halo -------------------------------------------------------------------
    def redistribute(self):
        """One exchange that carries both the migrants and the neighbour's boundary column, then ONE sort.

        Nothing is sorted before the exchange: the receiver's stable counting sort only needs arrivals in the sender's
        previous order, so the sender picks them with a stable selection on the NEW cell column.  Particles move less
        than a cell per step, so only the two old columns at each face have to be looked at."""
        e = self.eng
        a, b = self.a, self.b
        of, oc = self.own_first, self.own_count
        if self.col_start is None:                      # first step: no column table yet, look at everything owned
            reg_l = reg_r = (of, oc)
        else:
            cs = self.col_start
            reg_l = (cs[a], cs[min(a + 2, b)] - cs[a])
            reg_r = (cs[max(b - 2, a)], cs[b] - cs[max(b - 2, a)])
        if self.left is not None:
            e.select_columns(0, reg_l[0], reg_l[1], -(1 << 30), a)         # migrants to L + my first column
        if self.right is not None:
            e.select_columns(1, reg_r[0], reg_r[1], b - 1, 1 << 30)        # my last column + migrants to R
        n_to_l, n_to_r = e.select_counts() if (self.left is not None or self.right is not None) else (0, 0)
        if self.left is None:
            n_to_l = 0
        if self.right is None:
            n_to_r = 0
        n_from_l, n_from_r = self._exchange_counts(n_to_l, n_to_r)
        fields = e.state_fields
        buf = lambda cnt: e.new_buffer(e.message_bytes(fields, cnt)) if cnt else None
        out_l, out_r, in_l, in_r = buf(n_to_l), buf(n_to_r), buf(n_from_l), buf(n_from_r)
        if n_to_l:
            e.pack_selected(0, fields, n_to_l, out_l)
        if n_to_r:
            e.pack_selected(1, fields, n_to_r, out_r)
        self._sendrecv(out_l, out_r, in_l, in_r)
        e.replace(of, oc, in_l, n_from_l, in_r, n_from_r)                  # [from L][mine, ghosts dropped][from R]
        e.grid_build()
        cols = sorted({a - 1, a, a + 1, min(a + 2, b), max(b - 2, a), b - 1, b, b + 1})
        c = dict(zip(cols, e.column_starts(cols)))
        if c[a - 1] != 0 or c[b + 1] != e.n:
            raise RuntimeError(f"rank {self.rank}: particles moved more than one column in a step "
                               f"({c[a - 1]} below, {e.n - c[b + 1]} above the slab and its ghost columns)")
        if self.left is None and c[a] != 0 or self.right is None and c[b] != e.n:
            raise RuntimeError(f"rank {self.rank}: particles beyond the outermost slab")
        self.col_start = c
        self.ghost_l = (c[a - 1], c[a] - c[a - 1]) if self.left is not None else (0, 0)
        self.ghost_r = (c[b], c[b + 1] - c[b]) if self.right is not None else (0, 0)
        self.own_first, self.own_count = c[a], c[b] - c[a]
        self.send_l = (c[a], c[a + 1] - c[a]) if self.left is not None else (0, 0)
        self.send_r = (c[b - 1], c[b] - c[b - 1]) if self.right is not None else (0, 0)
        if self.check:                          # both sides of a face must agree on the column population
            gl, gr = self._exchange_counts(self.send_l[1], self.send_r[1])
            assert (self.left is None or gl == self.ghost_l[1]) and (self.right is None or gr == self.ghost_r[1]), \
                f"rank {self.rank}: ghost / boundary column sizes disagree"

    # ---- step 4: ghost refresh ------------------------------------------------------------------------
    def refresh_ghosts(self, fields):
        """Owners send the listed members of their boundary columns; ghosts are overwritten (same order, same count)."""
        if not fields:
            return
        e = self.eng
        buf = lambda cnt: e.new_buffer(e.message_bytes(fields, cnt)) if cnt else None
        out_l, out_r = buf(self.send_l[1]), buf(self.send_r[1])
        in_l, in_r = buf(self.ghost_l[1]), buf(self.ghost_r[1])
        if out_l is not None:
            e.pack(fields, self.send_l[0], self.send_l[1], out_l)
        if out_r is not None:
            e.pack(fields, self.send_r[0], self.send_r[1], out_r)
        self._sendrecv(out_l, out_r, in_l, in_r)
        if in_l is not None:
            e.unpack(fields, self.ghost_l[0], self.ghost_l[1], in_l)
        if in_r is not None:
            e.unpack(fields, self.ghost_r[0], self.ghost_r[1], in_r)

    def one_step(self, last=False):
        e = self.eng
        np_ = e.num_phases()
        for p in range(np_):
            e.one_step_phase(p)
            final = p == np_ - 1
            fields = e.phase_fields(p) + (e.deriv_fields if final else [])
            if final and last and not e.needs_final_ghosts:
                continue                        # nothing reads the ghosts' derivatives after the last one_step
            self.refresh_ghosts(fields)

    def step(self):
        """SPHBase.step (base:41-51) with substep (base:53-61)."""
        e = self.eng
        self.redistribute()
        e.calc_kernel_corr()
        e.init_real2tmp()
        if e.ti == 1:
            self.one_step(last=True)
            e.advect(0, 0)
        elif e.ti == 2:
            self.one_step()
            e.advect(1, 0)
            self.one_step(last=True)
            e.advect(0, 0)
        elif e.ti == 4:
            e.advect(3, 0)
            for s, m in enumerate((1, 2, 2, 1)):
                self.one_step(last=s == 3)
                e.advect(4, m)
                if s < 3:
                    e.advect(2, 0)
            e.advect(5, 0)
        else:
            raise AttributeError("timeIntegration 3 is broken in the reference (base:126-130)")
        e.advect_pos()
        if e.post_fields:
            self.refresh_ghosts(e.post_fields)
        e.post_step()
        e.enforce_boundary()                    # base:51 (pointwise; a no-op unless boundary == 1)

    def run_steps(self, n):
        for _ in range(n):
            self.step()


# ---------------------------------------------------------------------------------------------- CUDA binding
class CudaSlabEngine:
    """Slab-engine view of one libtisphi_b200 ctx (tisphi_b200/_lib.Engine)."""

    def __init__(self, engine, ti, xsph, solver):
        import torch
        from . import _lib
        self.torch = torch
        self.e = engine
        self.L = engine.L
        self.ti, self.xsph, self.solver = ti, xsph, solver
        F = _lib.FIELD_ID
        arr = (C.c_int32 * 16)()
        k = self.L.sph_state_fields(engine.h, arr, 16)
        self.state_fields = [int(arr[i]) for i in range(k)]
        fast = True
        try:
            engine.field("PK4", count=1)
        except _lib.SphError:
            fast = False
        wall_out = [F["V_TMP"], F["DENSITY_TMP"], F["PRESSURE"]] + ([F["PK4"]] if fast else [])
        if solver == 1:
            self._phase = {0: wall_out, 1: []}
            self.deriv_fields = [F["D_DENSITY"], F["D_VEL"]]
        elif solver == 2:
            self._phase = {0: [F["STRESS_TMP"], F["PRESSURE"]], 1: [F["V_TMP"], F["DENSITY_TMP"], F["STRESS_TMP"]], 2: []}
            self.deriv_fields = [F["D_DENSITY"], F["D_VEL"]]
        else:
            self._phase = {0: [F["STRESS_TMP"]], 1: [F["V_TMP"], F["DENSITY_TMP"], F["STRESS_TMP"]], 2: []}
            self.deriv_fields = [F["D_DENSITY"], F["D_VEL"], F["D_STRESS"]]
        # XSPH and the mu(I) regularisation sweep read neighbours after the last integrator kernel
        self.needs_final_ghosts = bool(xsph) or solver == 2
        self.post_fields = [F["X"]] if (solver == 2 and xsph) else []   # ghosts' XSPH sums are incomplete
        self._cache = {}

    @property
    def n(self):
        return self.e.n

    def _ids(self, fields):
        key = tuple(fields)
        if key not in self._cache:
            self._cache[key] = (C.c_int32 * len(fields))(*fields)
        return self._cache[key]

    def phase_fields(self, p):
        return list(self._phase[p])

    def new_buffer(self, nbytes):
        return self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.e.device)

    def new_counts(self):
        return self.torch.zeros(2, dtype=self.torch.int64, device=self.e.device)

    def message_bytes(self, fields, count):
        return int(self.L.sph_message_bytes(self.e.h, len(fields), self._ids(fields), int(count)))

    def pack(self, fields, first, count, buf):
        self.e.call("sph_pack_fields", len(fields), self._ids(fields), int(first), int(count), buf.data_ptr())

    def unpack(self, fields, first, count, buf):
        self.e.call("sph_unpack_fields", len(fields), self._ids(fields), int(first), int(count), buf.data_ptr())

    def select_columns(self, which, first, count, cx_lo, cx_hi):
        self.e.call("sph_select_columns", int(which), int(first), int(count), int(cx_lo), int(cx_hi))

    def select_counts(self):
        n0, n1 = C.c_int64(), C.c_int64()
        self.e.call("sph_select_counts", C.byref(n0), C.byref(n1))
        return int(n0.value), int(n1.value)

    def pack_selected(self, which, fields, count, buf):
        self.e.call("sph_pack_selected", int(which), len(fields), self._ids(fields), int(count), buf.data_ptr())

    def replace(self, keep_first, keep_count, left, n_left, right, n_right):
        self.e.call("sph_replace_particles", int(keep_first), int(keep_count), left.data_ptr() if left is not None else None,
                    int(n_left), right.data_ptr() if right is not None else None, int(n_right))

    def column_starts(self, cols):
        k = len(cols)
        cx, out = (C.c_int32 * k)(*[int(c) for c in cols]), (C.c_int64 * k)()
        self.e.call("sph_column_starts", k, cx, out)
        return [int(v) for v in out]

    def set_owned_columns(self, a, b):
        self.e.call("sph_set_owned_columns", int(a), int(b))

    def grid_build(self):
        self.e.call("sph_grid_build")

    def calc_kernel_corr(self):
        self.e.call("sph_calc_kernel_corr_deferred")

    def init_real2tmp(self):
        self.e.call("sph_init_real2tmp")

    def num_phases(self):
        return int(self.L.sph_num_phases(self.e.h))

    def one_step_phase(self, p):
        self.e.call("sph_one_step_phase", int(p))

    def advect(self, kind, m):
        self.e.call("sph_advect", int(kind), int(m))

    def advect_pos(self):
        self.e.call("sph_advect_pos")

    def post_step(self):
        self.e.call("sph_post_step")

    def enforce_boundary(self):
        self.e.call("sph_enforce_boundary")


class NativeSlab:
    """The slab protocol run by the library itself (tisphi_b200/csrc/slab.cu): device-resident counts and column table,
    messages stored by the packing kernels straight into the neighbours' inboxes, no host round trip inside a step.

    ``inbox_ptr`` is device memory of ``inbox_bytes(engine, face_cap)`` bytes owned by the caller (exportable over CUDA
    IPC when the neighbours are other processes)."""

    def __init__(self, engine, rank, world, columns, face_cap, inbox_ptr, inbox_bytes):
        self.e, self.L = engine, engine.L
        self.rank, self.world = rank, world
        self.a, self.b = int(columns[0]), int(columns[1])
        self.inbox_ptr = int(inbox_ptr)
        engine.call("sph_slab_init", rank, world, self.a, self.b, int(face_cap), inbox_ptr, int(inbox_bytes))
        self.own_first, self.own_count = 0, engine.n

    @staticmethod
    def inbox_bytes(engine, face_cap):
        return int(engine.L.sph_slab_inbox_bytes(engine.h, int(face_cap)))

    def connect(self, left_ptr, right_ptr):
        self.e.call("sph_slab_connect", left_ptr, right_ptr)

    def run_steps(self, n):
        self.e.call("sph_step", int(n))

    def sync(self):
        """Reads the device control block back (count, owned range); raises if the step set an error bit."""
        n, of, oc, err = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        self.e.call("sph_slab_sync", C.byref(n), C.byref(of), C.byref(oc), C.byref(err))
        self.own_first, self.own_count = int(of.value), int(oc.value)
        return int(n.value)

    def reset(self):
        """The particle set was replaced from outside (sph_clear_particles + sph_add_particles)."""
        self.own_first, self.own_count = 0, self.e.n

    @property
    def exchanges(self):
        return int(self.L.sph_slab_epoch(self.e.h))


def connect_p2p(eng, rank, world, columns, face_cap, group=None):
    """Allocate this rank's inbox, exchange CUDA IPC handles, map the neighbours' inboxes, arm the native slab step.
    Every rank takes part in every collective whatever happens locally, then all ranks raise together if any of them
    failed.  Returns (NativeSlab, inbox pointer, [mapped neighbour pointers])."""
    import torch.distributed as dist
    L = eng.L
    nbytes = NativeSlab.inbox_bytes(eng, face_cap)
    problem, drv, handle, ipc = None, None, (C.c_ubyte * 64)(), []
    with eng.torch.cuda.device(eng.device):
        inbox = L.sph_ipc_alloc(nbytes)
        if not inbox:
            problem = f"cudaMalloc of the {nbytes}-byte slab inbox failed"
        elif L.sph_ipc_get_handle(inbox, handle) != 0:
            problem = "cudaIpcGetMemHandle failed for the slab inbox"
        else:
            drv = NativeSlab(eng, rank, world, columns, face_cap, inbox, nbytes)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle) if problem is None else None, group=group)
        ptrs = [None, None]
        if problem is None:
            for side, nb in ((0, rank - 1), (1, rank + 1)):
                if 0 <= nb < world and handles[nb] is not None:
                    p = L.sph_ipc_open((C.c_ubyte * 64).from_buffer_copy(handles[nb]))
                    if not p:
                        problem = f"cudaIpcOpenMemHandle failed for the inbox of rank {nb} (no peer access between the GPUs?)"
                        break
                    ipc.append(p)
                    ptrs[side] = p
        problems = [None] * world
        dist.all_gather_object(problems, problem, group=group)
        if any(problems):
            for p in ipc:
                L.sph_ipc_close(p)
            if inbox:
                L.sph_ipc_free(inbox)
            raise RuntimeError("slab p2p transport unavailable: " + "; ".join(f"rank {r}: {p}" for r, p in enumerate(problems) if p))
        drv.connect(ptrs[0], ptrs[1])
    dist.barrier(group=group)
    return drv, inbox, ipc


class SlabSimulation:
    """``Simulation`` for one rank of a slab-partitioned run (same scene JSON; torch.distributed must be initialised).

    Every rank builds the scene on the host, keeps the particles of its columns (id0 stays the GLOBAL creation index,
    ps:208-211) and steps them.  ``transport``: "p2p" = the native device-driven step of slab.cu with the inboxes
    exchanged over CUDA IPC (one process per GPU of one box); "dist" = SlabDriver over torch.distributed send / recv
    (NCCL or gloo).  ``owned`` / ``gather_state`` reassemble arrays for parity checks."""

    def __init__(self, config, device, rank, world, group=None, wall_weight=0.15, columns=None, check=False,
                 transport="p2p"):
        from .eng.simulation import Simulation
        self.rank, self.world = rank, world
        self.sim = Simulation(config, device=device, slab=dict(rank=rank, world=world, wall_weight=wall_weight,
                                                                 columns=columns))
        self.ps, self.solver = self.sim.ps, self.sim.solver
        cfg = config
        self.columns = self.ps.slab_columns[rank]
        self.transport = transport
        self._ipc = []
        if transport == "p2p":
            self.driver = self._connect_p2p(group)
        else:
            self.engine = CudaSlabEngine(self.ps.engine, cfg.get_cfg("timeIntegration"), bool(cfg.get_cfg("xsph")),
                                         cfg.get_cfg("simulationMethod"))
            self.driver = SlabDriver(self.engine, self.columns, rank, world, int(self.ps.grid_num[0]), group=group, check=check)

    def _connect_p2p(self, group):
        drv, self._inbox, self._ipc = connect_p2p(self.ps.engine, self.rank, self.world, self.columns, self.ps.slab_face_cap, group)
        return drv

    def close(self):
        L = self.ps.engine.L
        with self.ps.engine.torch.cuda.device(self.ps.engine.device):
            for p in self._ipc:
                L.sph_ipc_close(p)
            self._ipc = []
            if getattr(self, "_inbox", None):
                self.ps.engine.torch.cuda.synchronize(self.ps.engine.device)
                L.sph_ipc_free(self._inbox)
                self._inbox = None

    def run_steps(self, n):
        self.driver.run_steps(n)

    def sync(self):
        if self.transport == "p2p":
            return self.driver.sync()
        return self.ps.engine.n

    def owned(self, name):
        """torch view of a member restricted to the particles this rank owns (current order)."""
        self.sync()
        d = self.driver
        return getattr(self.ps.pt, name)[d.own_first:d.own_first + d.own_count]


class LocalSlabGroup:
    """``world`` slabs of one scene as independent contexts (own arena, own stream) of ONE process on ONE device, connected
    through plain device pointers: the native slab step on a single-GPU box (tests/test_gpu_slab.py).  The ranks are
    stepped one step at a time in turn -- nothing ever blocks the host, every wait is a one-block kernel -- so each
    stream's waits are answered by kernels the host enqueues right afterwards on the other streams."""

    def __init__(self, config_factory, world, device="cuda:0", columns=None, wall_weight=0.15):
        import torch
        from .eng.simulation import Simulation
        torch.cuda.init()
        mode = C.c_int(0)
        try:                                     # CUmoduleLoadingMode: 1 eager, 2 lazy
            C.CDLL("libcuda.so.1").cuModuleGetLoadingMode(C.byref(mode))
        except OSError:
            pass
        if mode.value == 2:
            raise RuntimeError("LocalSlabGroup needs CUDA_MODULE_LOADING=EAGER set before CUDA initialises: with lazy "
                               "loading the first launch of a kernel synchronises with the spinning wait kernel of "
                               "another slab of this process (one process per GPU is not affected)")
        self.world = world
        self.sims, self.drivers, self._inboxes = [], [], []
        for r in range(world):
            with torch.cuda.stream(torch.cuda.Stream(device=device)):
                sim = Simulation(config_factory(), device=device, slab=dict(rank=r, world=world, wall_weight=wall_weight,
                                                                          columns=columns))
            self.sims.append(sim)
        torch.cuda.synchronize(device)
        for r, sim in enumerate(self.sims):
            eng = sim.ps.engine
            nbytes = NativeSlab.inbox_bytes(eng, sim.ps.slab_face_cap)
            box = torch.zeros(nbytes + 256, dtype=torch.uint8, device=device)
            off = (-box.data_ptr()) % 256
            self._inboxes.append(box)
            self.drivers.append(NativeSlab(eng, r, world, sim.ps.slab_columns[r], sim.ps.slab_face_cap, box.data_ptr() + off, nbytes))
        for r, d in enumerate(self.drivers):
            d.connect(self.drivers[r - 1].inbox_ptr if r > 0 else None, self.drivers[r + 1].inbox_ptr if r < world - 1 else None)
        self.global_particle_num = self.sims[0].ps.global_particle_num

    def run_steps(self, n):
        for _ in range(n):
            for d in self.drivers:
                d.run_steps(1)

    def gather(self, name):
        """Member ``name`` of the owned particles of every rank, concatenated in rank order (= the global sorted order)."""
        import torch
        parts = []
        for sim, d in zip(self.sims, self.drivers):
            d.sync()
            parts.append(getattr(sim.ps.pt, name)[d.own_first:d.own_first + d.own_count].clone())
        return torch.cat(parts)
