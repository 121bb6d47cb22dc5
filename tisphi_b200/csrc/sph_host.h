// sph_host.h -- host-side engine state shared by the translation units of libtisphi_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "sph_dev.cuh"

namespace sph {
struct SlabState;
// ---- native multi-GPU slab step (slab.cu); the sort (grid.cu) reads these too -------------------------------------
constexpr int SLAB_HDR = 256;                 // bytes in front of the message buffers of an inbox
struct SlabCtl {                 // device-resident; ctl->n is what Dev::ndev points to
    int n;
    int own_first, own_count;
    int ghost_first[2], ghost_count[2];     // [0]: ghost column a - 1, [1]: ghost column b
    int send_first[2], send_count[2];       // [0]: my column a (the left neighbour's ghosts), [1]: my column b - 1
    int reg_first[2], reg_count[2];         // where the next redistribution looks for particles that leave / are on a face
    int sel_count[2];
    int err;
    unsigned done[2];                       // completion counters of the push kernels, per side
    int src_nl, src_first, src_count;       // input of the pending sort: arrivals from L, then the OLD own range, then arrivals from R
};
struct InboxHdr {
    unsigned long long flag[2];             // [side the message came from]: epoch of the newest complete message
    int count[2][2];                        // [side][epoch parity]: particles in that message
};
// What the slab sort reads: the VIRTUAL concatenation [arrivals from L | own range | arrivals from R] -- arrivals stay in
// the inbox, own particles in their current buffers; nothing is copied together before the reorder kernel gathers.
// sec[k]: byte offset of the k-th carried member's section inside a message (order of sph_state_fields).
struct VSrc {
    const SlabCtl *ctl;
    const char *inbox;
    int has0, has1;
    unsigned parity;
    long long msg_cap;
    long long sec[12];
};
// after the scan: the column table of the slab (own / ghost / boundary ranges), written by one thread of k_scatter_index
struct ColTab { SlabCtl *ctl; int a, b, gn0, nyz, has0, has1; };
__host__ __device__ __forceinline__ char *inbox_msg(char *inbox, int side, unsigned parity, long long msg_cap) {
    return inbox + SLAB_HDR + (long long)(side * 2 + (int)parity) * msg_cap;
}
}

struct FieldSlot {
    int64_t off[2];      // byte offsets of the two ping-pong buffers (off[1] == off[0] when not carried)
    int cur;             // which buffer is current
    int ncomp, stride, kind;   // kind: 0 f64, 1 real, 2 i32
    bool present;
    int64_t view_shift;  // extra bytes (views MASS / M_V into the .w lane of V / XS)
};

struct SphCtx {
    SphParams p;
    int64_t n_max, n;
    int C;
    cudaStream_t stream;
    char err[512];
    char *arena;
    int64_t arena_bytes;
    FieldSlot f[SPH_F_NUM];
    // scratch (byte offsets)
    int64_t off_gid_unsorted, off_slot, off_perm, off_tmpidx, off_pnew, off_bad, off_scan_tiles, off_x_alt_unused;
    int64_t off_pw4;
    int64_t off_ps4, off_pk4, off_mask, off_nflow, off_cellflag, off_nflag, off_cellinfo, off_worklist;
    int64_t off_psoa, off_cellflow, soa_stride, wl_stride, off_nlist, off_lrounds;
    bool use_list, list_valid;   // neighbour round lists allocated / filled by a fluid pass since the last mask build
    bool fast;           // cell-tile fast path allocated (MIXED precision, WCSPH, no CSPM_L)
    int mask_words;
    int scan_tiles;
    int64_t launches;
    int real_bytes;      // sizeof engine real
    bool soil, rk, has_L;
    int press_cur;       // which of (PRESSURE buffer, pnew) ... handled through field table
    double r2thr64;
    bool prof_on;
    int prof_open;
    int64_t launches_by_kernel[32];
    void *prof_state;
    float r2thr32;
    bool shep_wall_pending;
    bool shep_pending;   // tile path: CSPM_f of flow particles is still to be formed by the next fluid pass
    bool fuse_init, fuse_half;   // sph_step only: init_real2tmp rides in the reorder kernel / advect_LF_half in k_tile_prep
    int own0, own1;      // owned x-columns [own0, own1) (multi-GPU slabs); the whole grid on one GPU
    int64_t off_slabctl; // device-resident control block of the native slab step (slab.cu)
    sph::SlabState *slab;   // native multi-GPU slab step (sph_slab_init); null on one GPU
    int64_t off_rigid;   // per dynamic rigid body: rest_cm, cm, mass, A, R (RIG_STRIDE doubles each)
    const int *rig_obj; const double *rig_x0; int rig_n;     // caller-owned device tables (sph_set_rigid_bodies)
    int64_t off_sor;     // soil: stress_tmp / density_tmp^2 of every particle, written right before a momentum sweep
    int64_t off_gnl, off_gnl_count;  // per-step neighbour lists of the generic sweeps (0: not allocated)
    int gnl_cap;
    bool gnl_valid;      // built for the current sort and positions
    bool slab_sort;      // grid_build is the sort of a slab redistribution: virtual concatenation, column table, cell sub-range
    bool masks_valid;    // the neighbour masks / work lists belong to the current sort (cleared by every re-sort / upload)
};

#define SPH_CHECK(ctx, call)                                                                          \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return -1;                                                                                \
        }                                                                                             \
    } while (0)

// per-kernel-class device timing (CUDA events on the launch stream), enabled with sph_profile_enable
enum SphKernelId { K_CELL_ID = 0, K_SCAN, K_SCATTER, K_RANK, K_REORDER, K_CSPM_F, K_CSPM_L, K_WC_EOS, K_WC_WALL, K_WC_FLUID,
                   K_MUI_SOIL1, K_SOIL_WALL, K_MUI_SOIL3, K_DP_ADAPT, K_DP_SOIL, K_ADVECT_POS, K_POST, K_POST_SWEEP,
                   K_NEIGHBOR_COUNT, K_DENSITY_SUM, K_OTHER, K_INIT_TMP, K_ADVECT, K_TILE_MASK, K_TILE_FLUID, K_TILE_WALL,
                   K_HALO, K_HALO_WAIT, K_C5, K_NLIST, K_NUM };
void sph_prof_begin(SphCtx *c, int id);
void sph_prof_end(SphCtx *c);

#define SPH_PROF(ctx, id)                                                                             \
    do {                                                                                              \
        if ((ctx)->prof_on) sph_prof_begin(ctx, id);                                                  \
        (ctx)->launches_by_kernel[id]++;                                                              \
    } while (0)

#define SPH_LAUNCH_CHECK(ctx)                                                                         \
    do {                                                                                              \
        (ctx)->launches++;                                                                            \
        if ((ctx)->prof_on) sph_prof_end(ctx);                                                        \
        SPH_CHECK(ctx, cudaGetLastError());                                                           \
    } while (0)

namespace sph {

constexpr int LIST_ROUNDS = 56;     // rounds (of four neighbour slots) a particle's list can hold

template <typename T> Dev<T> make_dev(SphCtx *c, int which = -1);   // which = -1: current buffers, 1: alternates

inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// grid.cu
template <typename T> int grid_build(SphCtx *c);
template <typename T> int select_columns(SphCtx *c, int which, int64_t first, int64_t count, int lo, int hi);
// sweeps.cu
template <typename T> int calc_kernel_corr(SphCtx *c, bool standalone);
template <typename T> int one_step(SphCtx *c, bool last = false);
template <typename T> int one_step_phase(SphCtx *c, int phase);
template <typename T> int advect_pos(SphCtx *c);
template <typename T> int post_step(SphCtx *c);
template <typename T> int enforce_boundary(SphCtx *c);
template <typename T> int init_rigid_body(SphCtx *c);
template <typename T> int solve_rigid_body(SphCtx *c);
template <typename T> int finish_step(SphCtx *c);      // advect_SE/LF + advect_pos + advect_something of WCSPH in one kernel
template <typename T> int neighbor_count(SphCtx *c, int32_t *out);
template <typename T> int density_sum(SphCtx *c, void *out);
template <typename T> int density_sweep(SphCtx *c, int32_t *count_out, void *rho_out);
// sweeps_tile.cu (float only)
int tile_mask(SphCtx *c, bool shepard);
int tile_mask_count(SphCtx *c, int32_t *out);
int tile_density_sweep(SphCtx *c, int32_t *count_out, float *rho_out);
int tile_wc_prep_and_wall(SphCtx *c);
int tile_wc_fluid(SphCtx *c);
// integrate.cu
template <typename T> int init_real2tmp(SphCtx *c);
template <typename T> int advect(SphCtx *c, int kind, int m);
template <typename T> int rk_stage(SphCtx *c, int m, bool first, bool last);   // sph_step only: fused RK4 pointwise stages
template <typename T> int init_stress(SphCtx *c, const double *ymax_ext = nullptr);
template <typename T> int add_particles_finish(SphCtx *c, int64_t first, int64_t count);

void flip(SphCtx *c, int field);
// halo.cu: current (or alternate) buffer of a member that can travel in a message + bytes per particle
bool field_ref(SphCtx *c, int f, bool alt, char **ptr, int *elem_bytes);
// slab.cu (native multi-GPU slab step; no-ops when c->slab is null)
template <typename T> int slab_redistribute(SphCtx *c);                       // migration + halo + ONE sort
int slab_refresh(SphCtx *c, int phase, bool final_phase, bool last_one_step); // ghost columns after a phase of one_step
int slab_refresh_post(SphCtx *c);                                             // ghost columns after advect_pos (mu(I) + XSPH)
int slab_arm(SphCtx *c);
void slab_sort_args(SphCtx *c, VSrc *vs, ColTab *ct, int *cell0, int *cell1);   // grid.cu asks while c->slab_sort is set
void slab_disarm(SphCtx *c);
bool slab_armed(const SphCtx *c);
int64_t slab_exact_n(const SphCtx *c);    // particle count the host knows (as of the last sph_slab_sync while stepping)
const int *slab_ndev(SphCtx *c);          // device-resident count while the native slab step runs, else null
void slab_free(SphCtx *c);

}  // namespace sph
