/*
 * tisphi_b200.h -- C ABI of the B200-native SPH engine (libtisphi_b200.so).
 *
 * This is the drop-in boundary for the data-parallel hot path of Rabmelon/tiSPHi.  The reference has no FFI: its
 * hot path sits behind Python methods whose bodies are Taichi kernels.  Each entry point below replaces one of
 * those methods (file:line into /root/reference); the Python classes in tisphi_b200/eng keep the reference's
 * names and call these through ctypes.  INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every array is a device pointer inside ONE caller-owned arena (a torch.uint8 tensor);
 *     the library never allocates device memory after sph_create (which allocates nothing on the device either).
 *   - all work is enqueued on the cudaStream_t given to sph_create; nothing synchronises unless documented.
 *   - return value: 0 = ok, < 0 = error (sph_last_error gives the text).  No C++ exceptions cross the boundary.
 *   - one ctx per (device, stream); calls on one ctx are not re-entrant; different ctxs are independent.
 *
 * Precision: SPH_PREC_F64 computes everything in float64 (the reference's default_fp, run_simulation.py:23).
 *            SPH_PREC_MIXED keeps positions and densities in float64 and everything else in float32; neighbour
 *            distances are evaluated on cell-local float32 coordinates.
 */
#ifndef TISPHI_B200_H
#define TISPHI_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SPH_PREC_F64 = 0, SPH_PREC_MIXED = 1 };
enum { SPH_SOLVER_WC = 1, SPH_SOLVER_MUI = 2, SPH_SOLVER_DP = 3 };   /* simulationMethod, eng/simulation.py:14-22 */

/* Scene-derived constants.  Filled by the host exactly as the reference's constructors compute them
 * (eng/particle_system.py:32-59, eng/solver_sph_base.py:12-28, solver_sph_wc.py:12-21, solver_sph_muI.py:12-25,
 * solver_sph_dp.py:12-35). */
typedef struct SphParams {
    int32_t dim;            /* 2 | 3                                              ps:14 */
    int32_t kernel;         /* 0 cubic spline, 1 Wendland C2                      base:15 */
    int32_t kcorr;          /* 0 none, 1 CSPM, 2 (MLS stub: zero gradient)        base:16 */
    int32_t ti;             /* 1 SE, 2 "LF" (explicit midpoint), 4 RK4            base:17, 53-61 */
    int32_t xsph;           /* 0 | 1                                              base:18 */
    int32_t solver;         /* SPH_SOLVER_*                                                */
    int32_t precision;      /* SPH_PREC_*                                                  */
    int32_t wc_fresh;       /* 0: wall pressure reads p_j as the serial reference does (EOS value for j < i, previous
                               value for j > i, wc:86-103); 1: race-free variant (EOS value for every j)            */
    int32_t gn[3];          /* grid_num                                           ps:57 */
    int32_t fast;           /* 0: generic sweeps only; 1: cell-tile fast sweeps where they apply; 2: additionally record
                               the neighbours found by the first fluid pass of a step and replay them in the later ones */
    double h, support, grid_size, vstart[3], m_V0, g[3], dt, eps;
    double rho0, visc, stiff, gamma_;                                   /* wc:12-15 */
    double coh, fric, E, poi, dila, vsound, mu, alpha, kc, G, K, eps_f; /* muI:12-24, dp:12-29 */
    /* boundary treatment (ps:18, 25): 0 none, 1 enforced collision, 2 dummy particles, 3 repulsive particles,
     * 4 dummy + repulsive.  Modes 3 / 4 add calc_repulsive_force (base:675-689) of type -2 neighbours to the momentum
     * sums (wc:119-121, muI:121-123); mode 1 makes sph_enforce_boundary clamp flow particles into the domain box. */
    int32_t boundary, pad_;
    double radius, dstart[3], dend[3];                                  /* particleRadius, domainStart, domainEnd */
} SphParams;

/* Particle members (names of eng/particle_func.py:13-69).  KIND: 0 = float64, 1 = engine real (float64 or float32
 * by precision), 2 = int32.  A member is described by (byte offset in the arena, component count, element stride,
 * kind); offsets of the members that are permuted by the sort change at every sph_grid_build (ping-pong). */
enum SphField {
    SPH_F_X = 0,          /* f64 x3                        pt.x            */
    SPH_F_V,              /* real x3 (stride 4; .w = mass) pt.v            */
    SPH_F_MASS,           /* real (view of V.w)            pt.mass         */
    SPH_F_M_V,            /* real (view of XS.w)           pt.m_V          */
    SPH_F_DENSITY,        /* f64                           pt.density      */
    SPH_F_DENSITY_TMP,    /* f64                           pt.density_tmp  */
    SPH_F_V_TMP,          /* real x3 (stride 4; .w = density_tmp rounded to real)  pt.v_tmp */
    SPH_F_PRESSURE,       /* real                          pt.pressure     */
    SPH_F_MAT_TYPE,       /* i32                           pt.mat_type     */
    SPH_F_ID0,            /* i32                           pt.id0          */
    SPH_F_GRID_IDS,       /* i32                           pt.grid_ids     */
    SPH_F_STRESS,         /* real x6 xx,yy,zz,xy,yz,zx     pt.stress       */
    SPH_F_STRESS_TMP,     /* real x6                       pt.stress_tmp   */
    SPH_F_STRAIN_EQU,     /* real                          pt.strain_equ   */
    SPH_F_STRAIN_EQU_P,   /* real                          pt.strain_equ_p */
    SPH_F_FLAG_RETMAP,    /* i32                           pt.flag_retmap  */
    SPH_F_CSPM_F,         /* real                          pt.CSPM_f       */
    SPH_F_CSPM_L,         /* real x9 row-major             pt.CSPM_L       */
    SPH_F_D_DENSITY,      /* real                          pt.d_density    */
    SPH_F_D_VEL,          /* real x3 (stride 4)            pt.d_vel        */
    SPH_F_D_STRESS,       /* real x6                       pt.d_stress     */
    SPH_F_V_GRAD,         /* real x9 row-major             pt.v_grad       */
    SPH_F_D_STRAIN_EQU,   /* real                          pt.d_strain_equ */
    SPH_F_D_STRAIN_EQU_P, /* real                          pt.d_strain_equ_p */
    SPH_F_D_DENSITY_RK,   /* real                          pt.d_density_RK */
    SPH_F_D_VEL_RK,       /* real x3 (stride 4)            pt.d_vel_RK     */
    SPH_F_D_STRESS_RK,    /* real x6                       pt.d_stress_RK  */
    SPH_F_XS,             /* real x3 (stride 4): sweep coordinates (global in F64, cell-local in MIXED) */
    SPH_F_CELL_END,       /* i32 x C: inclusive scan = grid_particle_num after prefix_sum.run (ps:256) */
    SPH_F_CELL_COUNT,     /* i32 x C: histogram = grid_particle_num_temp (ps:235-236)                 */
    SPH_F_ID_NEW,         /* i32: destination index of the last sort, pt.id_new (ps:245)              */
    SPH_F_PK4,            /* real x4: v_tmp.xyz, pressure / density_tmp^2 -- tile payload of the fluid pass; only
                             allocated when the cell-tile fast path is (MIXED precision WCSPH without CSPM_L)  */
    SPH_F_NUM
};

typedef struct SphCtx SphCtx;

/* bytes of device memory the caller must provide for n_max particles (includes all scratch) */
int64_t sph_arena_bytes(const SphParams *p, int64_t n_max);
/* arena: device pointer (256-byte aligned), stream: cudaStream_t (NULL = default stream) */
SphCtx *sph_create(const SphParams *p, int64_t n_max, void *arena, int64_t arena_bytes, void *stream);
void sph_destroy(SphCtx *ctx);
const char *sph_last_error(SphCtx *ctx);
int sph_set_params(SphCtx *ctx, const SphParams *p);           /* dt, g, material constants (not sizes)       */
int sph_field_info(SphCtx *ctx, int field, int64_t *offset_bytes, int32_t *ncomp, int32_t *stride, int32_t *kind);

/* ParticleSystem.add_particle / _add_particles / set_id0 (ps:274-314, 208-211): host (pinned or pageable)
 * arrays -> device, appended after the particles already present.  x, v: n x 3 float64; density: n float64;
 * mat_type: n int32.  Sets m_V = m_V0, mass = m_V0 * density, pressure = 0, id0 = running index. */
int sph_add_particles(SphCtx *ctx, int64_t n, const double *x, const double *v, const double *density,
                      const int32_t *mat_type);
int64_t sph_num_particles(SphCtx *ctx);
int sph_clear_particles(SphCtx *ctx);
/* device -> host copies of the dump() state (ps:459-545) in current (sorted) order; any pointer may be NULL.
 * Layout of the host buffers, n = sph_num_particles(ctx), R = sph_real_bytes(ctx) (8 in SPH_PREC_F64, 4 in MIXED):
 *   x         n x 3 float64                                   (24 n bytes)
 *   v         n x 4 engine reals: vx, vy, vz, mass            (4 R n bytes -- NOT n x 3 float64 like sph_add_particles)
 *   density   n float64                                       ( 8 n bytes)
 *   pressure  n engine reals                                  ( R n bytes)
 *   id0       n int32                                         ( 4 n bytes)
 * The arrays are copied as they are stored (no conversion pass).  Synchronises the stream. */
int32_t sph_real_bytes(SphCtx *ctx);
int sph_read_state(SphCtx *ctx, double *x, void *v, double *density, void *pressure, int32_t *id0);
/* the same copies enqueued on the ctx's stream without waiting (pinned host buffers); sph_synchronize waits for
 * everything enqueued on the ctx.  Several ctxs on different streams pipeline upload / step / download. */
int sph_read_state_async(SphCtx *ctx, double *x, void *v, double *density, void *pressure, int32_t *id0);
int sph_synchronize(SphCtx *ctx);

/* ParticleSystem.initialize_particle_system (ps:254-257): cell ids, histogram, inclusive scan, stable counting
 * sort, reorder of every carried member. */
int sph_grid_build(SphCtx *ctx);
/* SPHBase.calc_kernel_corr (base:363-368): CSPM_f always, CSPM_L when kcorr == 1 */
int sph_calc_kernel_corr(SphCtx *ctx);
/* the same, except that on the cell-tile path (MIXED precision WCSPH) the CSPM_f of particles whose sums the next
 * one_step forms anyway (walls: loop A, fluid: loop B) is completed by that one_step instead of an extra sweep. */
int sph_calc_kernel_corr_deferred(SphCtx *ctx);
/* SPHBase.init_real2tmp (base:67-74) */
int sph_init_real2tmp(SphCtx *ctx);
/* <Solver>.one_step (wc:82-126, muI:62-132, dp:210-274) */
int sph_one_step(SphCtx *ctx);
/* pointwise integrator kernels (base:79-170).  kind: 0 advect_SE / advect_LF, 1 advect_LF_half, 2 advect_RK_4,
 * 3 init_RK, 4 update_RK(m), 5 advect_RK */
int sph_advect(SphCtx *ctx, int kind, int m);
/* SPHBase.advect_pos (base:228-238), XSPH evaluated on a snapshot */
int sph_advect_pos(SphCtx *ctx);
/* SPHBase.advect_something (base:244-247 -> wc:129-132 | muI:134-156 | dp:276-296) */
int sph_post_step(SphCtx *ctx);
/* SPHBase.enforce_boundary (base:525-601): with boundary == 1 every flow particle outside the domain box (no lid) is put
 * back on its face and loses (1 + 0.3) of its normal velocity; a no-op in the other modes (dynamic rigid bodies: see
 * sph_solve_rigid_body) */
int sph_enforce_boundary(SphCtx *ctx);
/* ---- DYNAMIC rigid bodies (type 11 blocks with isDynamic), soil solvers only -- under WCSPH the reference
 * itself cannot run them (wc:129-132 calls a kernel from kernel scope).  The caller provides two device tables indexed by
 * the CREATION index id0 (they must stay alive): the body index of each particle (0 .. n_bodies-1; -1: not part of a
 * dynamic rigid body) and its rest position x0 (n_ids x 3 float64).  The sweeps then give those particles d_vel = g minus
 * the reaction of every momentum term they take part in (muI:45-46, 111-112; dp:164-165, 233-234), XSPH moves them like
 * any dynamic particle (base:234), sph_enforce_boundary clamps them into the domain box, and
 *   sph_init_rigid_body   stores each body's rest centre of mass (SPHBase.init_rigid_body, base:467-470);
 *   sph_solve_rigid_body  shape matching (solve_constraints, base:478-499): x_i := cm + R (x0_i - rest_cm) with R the
 *                         rotation of the polar decomposition of sum m (x - cm)(x0 - rest_cm)^T. */
int sph_set_rigid_bodies(SphCtx *ctx, int64_t n_ids, const int32_t *body_of_id0_dev, const double *x0_of_id0_dev, int32_t n_bodies);
int sph_init_rigid_body(SphCtx *ctx);
int sph_solve_rigid_body(SphCtx *ctx);
int sph_rigid_rest_cm(SphCtx *ctx, double *out_host);        /* n_bodies x 3 (ps.rigid_rest_cm); synchronises */
/* SPHBase.init_stress (base:249-260) */
int sph_init_stress(SphCtx *ctx);
/* the same with y_max given by the caller: a ctx that holds one slab of a scene must use the highest soil particle of
 * the WHOLE scene, not of its own columns */
int sph_init_stress_ymax(SphCtx *ctx, double y_max);
/* SPHBase.step (base:41-51) repeated nsteps times, enqueued without host round trips */
int sph_step(SphCtx *ctx, int nsteps);

/* stand-alone sweeps on the current grid (BASELINE config C5; parity of the neighbour predicate, ps:259-269) */
int sph_neighbor_count(SphCtx *ctx, int32_t *out_dev);          /* n int32  */
int sph_density_sum(SphCtx *ctx, void *out_dev);                /* n real: sum_j mass_j W_ij (wc:30-31) */
/* both in ONE walk of the neighbours (BASELINE config C5).  With the cell-tile path allocated (fast != 0, MIXED) the
 * walk follows the per-step neighbour bit masks (built here if they are not current): count = popcount, density over
 * the set bits with shared-memory tiles; wall particles and crowded cells go through the generic walk. */
int sph_density_sweep(SphCtx *ctx, int32_t *count_dev, void *rho_dev);
/* cell-tile path only, after sph_calc_kernel_corr: the count read back from the per-step neighbour bit masks for
 * flow particles of representable cells, -1 for every other particle (parity probe of the mask kernel) */
int sph_neighbor_count_masks(SphCtx *ctx, int32_t *out_dev);

/* number of particles whose cell fell outside the grid since the last call.  Synchronises. */
int64_t sph_read_bad_cells(SphCtx *ctx);
/* cells the cell-tile path left to the generic per-particle kernels at the last mask build (a stencil cell holds more
 * than 32 particles, or a tile overflowed); 0 without the fast path.  Synchronises. */
int64_t sph_read_flagged_cells(SphCtx *ctx);
/* how many kernels the library has launched on this ctx since creation */
int64_t sph_launch_count(SphCtx *ctx);
/* sizeof(SphParams) as compiled, so that bindings can verify their mirror of the struct */
int64_t sph_params_size(void);
/* per-kernel-class device timing with CUDA events on the engine's stream (used by bench.py for the roofline line).
 * ms_by_kernel / launches_by_kernel: arrays of sph_profile_num_kernels() entries; reading synchronises and resets. */
int sph_profile_enable(SphCtx *ctx, int on);
int sph_profile_num_kernels(void);
const char *sph_profile_name(int id);
int sph_profile_read(SphCtx *ctx, double *ms_by_kernel, int64_t *launches_by_kernel);

/* ---- multi-GPU slab decomposition (DESIGN.md "Multi-GPU"; the reference is single-device) -----------
 * A rank owns the x-columns [cx_begin, cx_end) of the GLOBAL grid and additionally holds one ghost column on each
 * side.  The flattened cell id is x-major (ps:221-222), so after sph_grid_build every column is ONE contiguous
 * index range of every member array: halo and migration messages are plain ranges, never gathers.
 * A message is a device buffer holding, for each listed member in order, `count` consecutive elements of that
 * member (element = stride x sizeof(kind)), each section padded to 16 bytes.                                      */
/* <Solver>.one_step split at its top-level loops (the points where ghost columns must be refreshed):
 *   WCSPH  0: loop A (EOS + wall extrapolation, wc:86-106)        1: loop B (continuity + momentum, wc:108-126)
 *   mu(I)  0: loop 1 (muI:67-92)   1: loop 2 walls (muI:95-109)   2: loop 3 momentum (muI:115-128)
 *   DP     0: loop 1 (dp:215-217)  1: loop 2 walls (dp:220-231)   2: loop 3 (dp:237-270)                          */
int sph_num_phases(SphCtx *ctx);
int sph_one_step_phase(SphCtx *ctx, int phase);
/* sweeps only update particles whose cell column lies in [cx_begin, cx_end); pointwise kernels still run on all.
 * (0, grid_num[0]) restores the single-GPU behaviour. */
int sph_set_owned_columns(SphCtx *ctx, int32_t cx_begin, int32_t cx_end);
/* index of the first particle of each listed column (cx <= 0 -> 0, cx >= grid_num[0] -> n) after the last
 * sph_grid_build.  Synchronises the stream. */
int sph_column_starts(SphCtx *ctx, int32_t ncols, const int32_t *cx, int64_t *start_out);
/* the members that travel with a particle through the sort (the migration / halo record); returns their number */
int sph_state_fields(SphCtx *ctx, int32_t *fields_out, int32_t capacity);
int64_t sph_message_bytes(SphCtx *ctx, int32_t nfields, const int32_t *fields, int64_t count);
int sph_pack_fields(SphCtx *ctx, int32_t nfields, const int32_t *fields, int64_t first, int64_t count, void *msg_dev);
int sph_unpack_fields(SphCtx *ctx, int32_t nfields, const int32_t *fields, int64_t first, int64_t count,
                      const void *msg_dev);
/* Migration without a sort: a STABLE index list (previous order kept) of the particles of [first, first + count)
 * whose NEW cell column lies in [cx_lo, cx_hi]; two lists can be pending (which = 0 | 1).  sph_select_counts
 * returns both lengths (synchronises); sph_pack_selected gathers the listed particles into a message. */
int sph_select_columns(SphCtx *ctx, int32_t which, int64_t first, int64_t count, int32_t cx_lo, int32_t cx_hi);
int sph_select_counts(SphCtx *ctx, int64_t *n0, int64_t *n1);
int sph_pack_selected(SphCtx *ctx, int32_t which, int32_t nfields, const int32_t *fields, int64_t count, void *msg_dev);
/* particle set := [n_left particles of left_msg][keep_count current particles from keep_first][n_right of right_msg]
 * (messages hold the sph_state_fields members).  Arrivals from the lower-x neighbour go in front and those from
 * the higher-x neighbour behind, so that the next stable sort reproduces the single-GPU global order. */
int sph_replace_particles(SphCtx *ctx, int64_t keep_first, int64_t keep_count, const void *left_msg, int64_t n_left,
                          const void *right_msg, int64_t n_right);


/* ---- the slab step without host round trips (tisphi_b200/csrc/slab.cu) ---------------------------------------------
 * The calls above let a host program run the slab protocol itself (tisphi_b200/parallel.py::SlabDriver does, over
 * torch.distributed).  The calls below move the whole protocol onto the device: after sph_slab_init + sph_slab_connect,
 * sph_step(ctx, nsteps) runs SPHBase.step (base:41-51) on the slab -- migration, the one sort, every ghost refresh --
 * with the particle count and the column table kept in device memory and every message STORED by the packing kernel
 * straight into the neighbour's inbox (a peer mapping over NVLink), completed by a system-scope flag the receiver's
 * stream waits on.  Results are bit-identical to the single-GPU run (the stable sort, see above).
 *   inbox: device memory of sph_slab_inbox_bytes() bytes, 256-byte aligned, owned by the caller; with one process per
 *   GPU it must be its own cudaMalloc allocation so that it can be exported (sph_ipc_*).  face_cap: the most particles
 *   one message may carry (a boundary column + migrants); exceeding it sets an error bit, it never overruns. */
int64_t sph_slab_inbox_bytes(SphCtx *ctx, int64_t face_cap);
int sph_slab_init(SphCtx *ctx, int32_t rank, int32_t world, int32_t cx_begin, int32_t cx_end, int64_t face_cap, void *inbox,
                  int64_t inbox_bytes);
/* the neighbours' inboxes as device pointers valid on THIS device (NULL at the ends of the row of slabs) */
int sph_slab_connect(SphCtx *ctx, void *left_inbox, void *right_inbox);
/* While the slab steps, the host does not know the particle count.  This reads the device control block back:
 * count (owned + ghosts), the index range of the owned particles, sticky error bits (timeout 1, ghost / boundary
 * column sizes disagree 2, capacity 4, a particle moved more than one column 8, face capacity 16; != 0 returns -4).
 * Synchronises the stream.  sph_num_particles / sph_read_state* use the count of the last sync. */
int sph_slab_sync(SphCtx *ctx, int64_t *n, int64_t *own_first, int64_t *own_count, int32_t *err_bits);
int64_t sph_slab_epoch(SphCtx *ctx);                 /* exchanges enqueued so far (every rank counts the same) */
/* CUDA IPC plumbing for one process per GPU: allocate an exportable inbox, export it as a 64-byte handle, map a
 * neighbour's handle (peer access is enabled lazily).  Handles travel through any host channel (torch.distributed). */
void *sph_ipc_alloc(int64_t bytes);
void sph_ipc_free(void *dev_ptr);
int sph_ipc_get_handle(void *dev_ptr, void *handle64);
void *sph_ipc_open(const void *handle64);
void sph_ipc_close(void *dev_ptr);

#ifdef __cplusplus
}
#endif
#endif
