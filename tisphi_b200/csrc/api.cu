// api.cu -- the C ABI of libtisphi_b200 (include/tisphi_b200.h): arena layout, parameter derivation, dispatch.
#include <math.h>
#include <new>
#include <string.h>
#include "sph_host.h"

using namespace sph;

namespace {

inline int64_t align_up(int64_t v, int64_t a = 256) { return (v + a - 1) / a * a; }

struct FieldSpec { int ncomp, stride, kind; bool carried; int need; };   // need: 0 always, 1 soil, 2 rk, 3 cspm_L, 4 soil+rk, 5 fast
// order == enum FieldSlot
const FieldSpec SPEC[SPH_F_NUM] = {
    /* X            */ {3, 3, 0, true, 0},
    /* V            */ {3, 4, 1, true, 0},
    /* MASS (view)  */ {1, 4, 1, false, 0},
    /* M_V (view)   */ {1, 4, 1, false, 0},
    /* DENSITY      */ {1, 1, 0, true, 0},
    /* DENSITY_TMP  */ {1, 1, 0, false, 0},
    /* V_TMP        */ {3, 4, 1, true, 0},
    /* PRESSURE     */ {1, 1, 1, true, 0},
    /* MAT_TYPE     */ {1, 1, 2, true, 0},
    /* ID0          */ {1, 1, 2, true, 0},
    /* GRID_IDS     */ {1, 1, 2, false, 0},
    /* STRESS       */ {6, 6, 1, true, 1},
    /* STRESS_TMP   */ {6, 6, 1, false, 1},
    /* STRAIN_EQU   */ {1, 1, 1, true, 1},
    /* STRAIN_EQU_P */ {1, 1, 1, true, 1},
    /* FLAG_RETMAP  */ {1, 1, 2, true, 1},
    /* CSPM_F       */ {1, 1, 1, false, 0},
    /* CSPM_L       */ {9, 9, 1, false, 3},
    /* D_DENSITY    */ {1, 1, 1, false, 0},
    /* D_VEL        */ {3, 4, 1, false, 0},
    /* D_STRESS     */ {6, 6, 1, false, 1},
    /* V_GRAD       */ {9, 9, 1, false, 1},
    /* D_STRAIN_EQU */ {1, 1, 1, false, 1},
    /* D_STRAIN_EQU_P*/ {1, 1, 1, false, 1},
    /* D_DENSITY_RK */ {1, 1, 1, false, 2},
    /* D_VEL_RK     */ {3, 4, 1, false, 2},
    /* D_STRESS_RK  */ {6, 6, 1, false, 4},
    /* XS           */ {3, 4, 1, true, 0},
    /* CELL_END     */ {1, 1, 2, false, 0},
    /* CELL_COUNT   */ {1, 1, 2, false, 0},
    /* ID_NEW       */ {1, 1, 2, false, 0},
    /* PK4          */ {4, 4, 1, false, 5},
};

inline int elem_bytes(int kind, int real_bytes) { return kind == 0 ? 8 : (kind == 1 ? real_bytes : 4); }

int64_t n_cells(const SphParams *p) { return (int64_t)p->gn[0] * p->gn[1] * (p->dim == 3 ? p->gn[2] : 1); }

// lays the arena out; when c == nullptr only the size is computed
int64_t layout(const SphParams *p, int64_t n_max, SphCtx *c) {
    const int rb = p->precision == SPH_PREC_F64 ? 8 : 4;
    const bool soil = p->solver != SPH_SOLVER_WC, rk = p->ti == 4, hasL = p->kcorr == 1;
    const int64_t C = n_cells(p);
    // cell-tile fast path scratch: MIXED precision WCSPH without CSPM_L
    // (the payload sign of the tile passes tells flow from non-flow only: repulsive particles, boundary 3 / 4, need the type)
    const bool fast = p->fast && p->precision == SPH_PREC_MIXED && p->solver == SPH_SOLVER_WC && p->kcorr == 0 &&
                      p->boundary != 3 && p->boundary != 4;
    int64_t off = 0;
    for (int f = 0; f < SPH_F_NUM; f++) {
        const FieldSpec &s = SPEC[f];
        bool present = s.need == 0 || (s.need == 1 && soil) || (s.need == 2 && rk) || (s.need == 3 && hasL) ||
                       (s.need == 4 && soil && rk) || (s.need == 5 && fast);
        if (f == SPH_F_MASS || f == SPH_F_M_V) present = true;
        int64_t count = (f == SPH_F_CELL_END || f == SPH_F_CELL_COUNT) ? C + 1 : n_max;
        int64_t bytes = align_up(count * s.stride * elem_bytes(s.kind, rb));
        int64_t o0 = 0, o1 = 0;
        if (present && f != SPH_F_MASS && f != SPH_F_M_V) {
            o0 = off; off += bytes;
            o1 = o0;
            if (s.carried) { o1 = off; off += bytes; }
        }
        if (c) {
            FieldSlot &F = c->f[f];
            F.off[0] = o0; F.off[1] = o1; F.cur = 0;
            F.ncomp = s.ncomp; F.stride = s.stride; F.kind = s.kind; F.present = present; F.view_shift = 0;
        }
    }
    const int64_t ib = align_up(n_max * 4);
    int64_t o_gid = off; off += ib;
    int64_t o_slot = off; off += ib;
    int64_t o_perm = off; off += ib;
    int64_t o_tmp = off; off += ib;
    int64_t o_bad = off; off += 256;
    int64_t o_ctl = off; off += 256;                                   // SlabCtl (slab.cu)
    // Per-step neighbour lists of the generic sweeps (every solver / precision that does not run the cell-tile WCSPH
    // path).  2D: ~28 neighbours among 81 candidates, 64 entries cover compressed states; 3D would need ~160 entries
    // (640 B per particle), so there the sweeps keep walking the cells.  Particles that do not fit walk, too.
    int64_t o_rig = off; off += align_up((int64_t)RIG_MAX * RIG_STRIDE * 8);
    int64_t o_sor = 0;
    if (soil) { o_sor = off; off += align_up(n_max * 6 * (int64_t)rb); }
    const int nl_cap = (p->fast && !fast && p->dim == 2 && n_max < (1ll << 27)) ? 64 : 0;
    int64_t o_nl = 0, o_nc = 0;
    if (nl_cap) { o_nl = off; off += align_up(n_max * 4 * (int64_t)nl_cap); o_nc = off; off += align_up(n_max * 4); }
    const int64_t nt = ((C > n_max ? C : n_max) + 2047) / 2048 + 2;    // scan tiles: cells (grid build) or particles (selection)
    int64_t o_tiles = off; off += align_up(nt * 4);
    const int mask_words = p->dim == 3 ? 27 : 9;
    int64_t o_pw4 = 0;
    int64_t o_ps4 = 0, o_pk4 = 0, o_mask = 0, o_nflow = 0, o_cflag = 0, o_nflag = 0, o_cinfo = 0, o_wl = 0, o_soa = 0, o_cflow = 0;
    const int64_t soa_stride = align_up(n_max * 4 + 64), wl_stride = align_up((C + 1) * 4);
    // neighbour round lists: SE has a single fluid pass per step (nothing to replay); entries are addressed with 32 bits
    const bool lists = fast && p->fast >= 2 && p->ti != 1 && n_max * (int64_t)LIST_ROUNDS < (1ll << 32);
    int64_t o_nlist = 0, o_lrounds = 0;
    if (fast) {
        o_ps4 = off; off += align_up(n_max * 16);
        o_pw4 = off; off += align_up(n_max * 16);
        o_mask = off; off += align_up(n_max * 4 * mask_words);
        o_nflow = off; off += align_up(n_max * 4);
        o_cflag = off; off += align_up(C + 1);
        o_nflag = off; off += 256;
        o_cinfo = off; off += align_up(C + 1);
        o_wl = off; off += 4 * wl_stride;                 // work lists: occupied / flow / wall / mask segments
        o_soa = off; off += 4 * soa_stride;               // psx, psy, psz, psf
        o_cflow = off; off += align_up(C + 1);
        if (lists) {
            o_nlist = off; off += align_up(n_max * 8 * (int64_t)LIST_ROUNDS);
            o_lrounds = off; off += align_up((C + 1) * 4);
        }
    }
    if (c) {
        c->off_pw4 = o_pw4;
        c->off_ps4 = o_ps4; c->off_pk4 = fast ? c->f[SPH_F_PK4].off[0] : o_pk4; c->off_mask = o_mask; c->off_nflow = o_nflow; c->off_cellflag = o_cflag;
        c->off_nflag = o_nflag; c->off_cellinfo = o_cinfo; c->off_worklist = o_wl; c->fast = fast; c->mask_words = mask_words;
        c->off_psoa = o_soa; c->off_cellflow = o_cflow; c->soa_stride = soa_stride; c->wl_stride = wl_stride;
        c->off_nlist = o_nlist; c->off_lrounds = o_lrounds; c->use_list = lists; c->list_valid = false;
        c->off_gid_unsorted = o_gid; c->off_slot = o_slot; c->off_perm = o_perm; c->off_tmpidx = o_tmp;
        c->off_bad = o_bad; c->off_scan_tiles = o_tiles; c->off_slabctl = o_ctl;
        c->off_rigid = o_rig; c->off_sor = o_sor; c->off_gnl = o_nl; c->off_gnl_count = o_nc; c->gnl_cap = nl_cap; c->gnl_valid = false;
        c->real_bytes = rb; c->soil = soil; c->rk = rk; c->has_L = hasL; c->C = (int)C;
    }
    return off;
}

// smallest t with sqrt(t) >= s (so that r2 < t  <=>  sqrt(r2) < s for correctly rounded sqrt)
double r2_threshold64(double s) {
    double t = s * s;
    while (sqrt(t) >= s) t = nextafter(t, 0.0);
    while (sqrt(t) < s) t = nextafter(t, INFINITY);
    return t;
}
float r2_threshold32(float s) {
    float t = s * s;
    while (sqrtf(t) >= s) t = nextafterf(t, 0.0f);
    while (sqrtf(t) < s) t = nextafterf(t, INFINITY);
    return t;
}

double kernel_norm(const SphParams *p) {       // base:300-358 (3D Wendland constant as in the reference, H8)
    const double h1 = 1.0 / p->h;
    double k;
    if (p->kernel == 0) k = p->dim == 2 ? 15.0 / 7.0 / M_PI : 3.0 / 2.0 / M_PI;
    else k = p->dim == 2 ? 7.0 / (4.0 * M_PI) : 21.0 / (2.0 * M_PI);
    double hp = h1;
    for (int a = 1; a < p->dim; a++) hp *= h1;
    return k * hp;
}

template <typename T> int dispatch_step(SphCtx *c, int nsteps);

}  // namespace

// ---------------------------------------------------------------------------------------------- profiling
#include <vector>
namespace {
struct ProfState {
    std::vector<cudaEvent_t> pool;         // event pairs (begin, end)
    std::vector<int> ids;                  // kernel id of each recorded pair
    size_t used = 0;
};
const char *KNAMES[K_NUM] = {"cell_id", "scan", "scatter_index", "rank", "reorder", "cspm_f", "cspm_L", "wc_eos", "wc_wall",
                             "wc_fluid", "mui_soil1", "soil_wall", "mui_soil3", "dp_adapt", "dp_soil", "advect_pos", "post",
                             "post_sweep", "neighbor_count", "density_sum", "other", "init_real2tmp", "advect", "tile_mask",
                             "tile_fluid", "tile_wall", "halo", "halo_wait", "c5_sweep", "nlist_build"};
}
void sph_prof_begin(SphCtx *c, int id) {
    ProfState *ps = (ProfState *)c->prof_state;
    if (ps->used + 2 > ps->pool.size()) {
        for (int k = 0; k < 2; k++) { cudaEvent_t e; cudaEventCreate(&e); ps->pool.push_back(e); }
    }
    ps->ids.push_back(id);
    cudaEventRecord(ps->pool[ps->used], c->stream);
    c->prof_open = 1;
}
void sph_prof_end(SphCtx *c) {
    if (!c->prof_open) return;
    ProfState *ps = (ProfState *)c->prof_state;
    cudaEventRecord(ps->pool[ps->used + 1], c->stream);
    ps->used += 2;
    c->prof_open = 0;
}

namespace sph {

void flip(SphCtx *c, int field) { c->f[field].cur ^= 1; }

template <typename T> Dev<T> make_dev(SphCtx *c, int which) {
    const SphParams &p = c->p;
    Dev<T> d;
    memset(&d, 0, sizeof(d));
    d.n = (int)c->n;
    d.ndev = slab_ndev(c);
    if (c->gnl_valid) {
        d.gnl = (const unsigned *)(c->arena + c->off_gnl); d.gnl_count = (const int *)(c->arena + c->off_gnl_count);
        d.gnl_stride = (int)c->n_max; d.gnl_cap = c->gnl_cap;
    }
    d.dim = p.dim; d.kernel = p.kernel; d.kcorr = p.kcorr; d.solver = p.solver; d.xsph = p.xsph; d.wc_fresh = p.wc_fresh;
    for (int a = 0; a < 3; a++) { d.gn[a] = p.gn[a]; d.vstart[a] = p.vstart[a]; d.g[a] = (T)p.g[a]; }
    if (p.dim == 2) d.gn[2] = 1;
    d.C = c->C;
    d.own0 = c->own0; d.own1 = c->own1;
    d.gs = p.grid_size; d.dt = p.dt; d.m_V0d = p.m_V0;
    d.h = (T)p.h; d.hinv = (T)(1.0 / p.h); d.support = (T)p.support;
    d.r2thr = sizeof(T) == 8 ? (T)c->r2thr64 : (T)c->r2thr32;
    d.eps = (T)p.eps; d.knorm = (T)kernel_norm(&p);
    d.gsT = sizeof(T) == 8 ? (T)0 : (T)p.grid_size;
    d.m_V0 = (T)p.m_V0;
    d.visc_coef = (T)(2 * (p.dim + 2) * p.visc); d.rho0T = (T)p.rho0; d.h2_001 = (T)(0.01 * p.h * p.h);
    d.rho0 = p.rho0; d.stiff = p.stiff; d.gamma_ = p.gamma_; d.vsound = p.vsound;
    d.coh = (T)p.coh; d.mu = (T)p.mu; d.E = (T)p.E; d.alpha = (T)p.alpha; d.kc = (T)p.kc; d.G = (T)p.G; d.K = (T)p.K;
    d.eps_f = (T)p.eps_f; d.sin_dila = (T)sin(p.dila);
    d.damp_c = (T)(-5e-5 * sqrt(p.E) / p.h);
    d.boundary = p.boundary; d.radius_d = p.radius;
    for (int a = 0; a < 3; a++) { d.dstart[a] = p.dstart[a]; d.dend[a] = p.dend[a]; }
    d.rep_k = (T)(0.01 * p.vsound * p.vsound); d.rep_judge = (T)(2.0 * p.radius); d.rep_ginv = (T)(1.0 / (0.75 * p.h));
    auto ptr = [&](int f, bool alt) -> char * {
        const FieldSlot &F = c->f[f];
        if (!F.present) return nullptr;
        int b = F.cur;
        if (alt) b ^= 1;
        return c->arena + F.off[b];
    };
    const bool alt = which == 1;
    d.x = (double *)ptr(SPH_F_X, alt);
    d.rho = (double *)ptr(SPH_F_DENSITY, alt);
    d.rho_t = (double *)ptr(SPH_F_DENSITY_TMP, false);
    d.v4 = (Vec4<T> *)ptr(SPH_F_V, alt);
    d.vt4 = (Vec4<T> *)ptr(SPH_F_V_TMP, alt);
    d.xs4 = (Vec4<T> *)ptr(SPH_F_XS, alt);
    d.press = (T *)ptr(SPH_F_PRESSURE, alt);
    d.pnew = (T *)ptr(SPH_F_PRESSURE, !alt);
    d.type = (int *)ptr(SPH_F_MAT_TYPE, alt);
    d.id0 = (int *)ptr(SPH_F_ID0, alt);
    d.gid = (int *)ptr(SPH_F_GRID_IDS, false);
    d.flag = (int *)ptr(SPH_F_FLAG_RETMAP, alt);
    d.stress = (T *)ptr(SPH_F_STRESS, alt);
    d.stress_t = (T *)ptr(SPH_F_STRESS_TMP, false);
    d.sor = c->off_sor ? (T *)(c->arena + c->off_sor) : nullptr;
    d.rig_obj = c->rig_n > 0 ? c->rig_obj : nullptr; d.rig_x0 = c->rig_x0; d.rig_n = c->rig_n;
    d.rig_buf = (double *)(c->arena + c->off_rigid);
    d.strain = (T *)ptr(SPH_F_STRAIN_EQU, alt);
    d.strain_p = (T *)ptr(SPH_F_STRAIN_EQU_P, alt);
    d.cspm_f = (T *)ptr(SPH_F_CSPM_F, false);
    d.cspm_L = (T *)ptr(SPH_F_CSPM_L, false);
    d.d_rho = (T *)ptr(SPH_F_D_DENSITY, false);
    d.d_vel = (Vec4<T> *)ptr(SPH_F_D_VEL, false);
    d.d_stress = (T *)ptr(SPH_F_D_STRESS, false);
    d.v_grad = (T *)ptr(SPH_F_V_GRAD, false);
    d.d_strain = (T *)ptr(SPH_F_D_STRAIN_EQU, false);
    d.d_strain_p = (T *)ptr(SPH_F_D_STRAIN_EQU_P, false);
    d.d_rho_rk = (T *)ptr(SPH_F_D_DENSITY_RK, false);
    d.d_vel_rk = (Vec4<T> *)ptr(SPH_F_D_VEL_RK, false);
    d.d_stress_rk = (T *)ptr(SPH_F_D_STRESS_RK, false);
    d.cell_end = (int *)ptr(SPH_F_CELL_END, false);
    d.cell_cnt = (int *)ptr(SPH_F_CELL_COUNT, false);
    d.bad = (unsigned long long *)(c->arena + c->off_bad);
    if (c->fast) {
        d.ps4 = (Vec4<T> *)(c->arena + c->off_ps4);
        d.pk4 = (Vec4<T> *)(c->arena + c->off_pk4);
        d.pw4 = (Vec4<T> *)(c->arena + c->off_pw4);
        d.mask = (unsigned *)(c->arena + c->off_mask);
        d.nzw = (unsigned *)(c->arena + c->off_nflow);
        d.cellflag = (unsigned char *)(c->arena + c->off_cellflag);
        d.nflag = (int *)(c->arena + c->off_nflag);
        d.cellinfo = (unsigned char *)(c->arena + c->off_cellinfo);
        for (int k = 0; k < 4; k++) d.worklist[k] = (int *)(c->arena + c->off_worklist + k * c->wl_stride);
        d.wcount = d.nflag + 4;              // list lengths [0..3], cursors [4..] after the flag counter
        d.psx = (T *)(c->arena + c->off_psoa); d.psy = (T *)(c->arena + c->off_psoa + c->soa_stride);
        d.psz = (T *)(c->arena + c->off_psoa + 2 * c->soa_stride); d.psf = (T *)(c->arena + c->off_psoa + 3 * c->soa_stride);
        d.cellflow = (unsigned char *)(c->arena + c->off_cellflow);
        if (c->use_list) { d.nlist = (uint2 *)(c->arena + c->off_nlist); d.lrounds = (int *)(c->arena + c->off_lrounds); }
    }
    return d;
}
template Dev<float> make_dev<float>(SphCtx *, int);
template Dev<double> make_dev<double>(SphCtx *, int);

}  // namespace sph

namespace {

// SPHBase.step (base:41-51) + substep (base:53-61)
template <typename T> int step_once(SphCtx *c) {
    int r;
    // WCSPH without XSPH: pointwise stages ride in neighbouring kernels (identical operations, fewer HBM round trips):
    // init_real2tmp in the reorder kernel, advect_LF_half in the EOS / payload kernel of the cell-tile path, and
    // advect_SE|LF + advect_pos + advect_something in one kernel.
    const bool wc_fused = c->p.solver == SPH_SOLVER_WC && !c->p.xsph && (c->p.ti == 1 || c->p.ti == 2);
    c->fuse_half = false;                                     // (a step that failed half way must not leave the flag behind)
    c->fuse_init = wc_fused;
    if ((r = c->slab ? slab_redistribute<T>(c) : grid_build<T>(c))) return r;
    c->fuse_init = false;
    if ((r = calc_kernel_corr<T>(c, false))) return r;
    if (!wc_fused && (r = init_real2tmp<T>(c))) return r;
    switch (c->p.ti) {
    case 1:
        if ((r = one_step<T>(c, true))) return r;
        if (!wc_fused && (r = advect<T>(c, 0, 0))) return r;
        break;
    case 2:
        if ((r = one_step<T>(c, false))) return r;
        if (wc_fused && c->fast) c->fuse_half = true;          // consumed by the next one_step's first kernel
        else if ((r = advect<T>(c, 1, 0))) return r;
        if ((r = one_step<T>(c, true))) return r;
        if (!wc_fused && (r = advect<T>(c, 0, 0))) return r;
        break;
    case 4: {
        static const int m[4] = {1, 2, 2, 1};
        for (int s = 0; s < 4; s++) {          // init_RK / update_RK / advect_RK_4 / advect_RK fused per stage (integrate.cu)
            if ((r = one_step<T>(c, s == 3))) return r;
            if ((r = rk_stage<T>(c, m[s], s == 0, s == 3))) return r;
        }
    } break;
    default:
        snprintf(c->err, sizeof(c->err), "timeIntegration %d is not runnable (3 is broken in the reference, base:126-130)", c->p.ti);
        return -2;
    }
    if (wc_fused) r = finish_step<T>(c);
    else {
        if ((r = advect_pos<T>(c))) return r;
        if ((r = slab_refresh_post(c))) return r;
        r = post_step<T>(c);
    }
    if (r) return r;
    if ((r = solve_rigid_body<T>(c))) return r;
    return enforce_boundary<T>(c);
}
template <typename T> int dispatch_step(SphCtx *c, int nsteps) {
    for (int s = 0; s < nsteps; s++) {
        int r = step_once<T>(c);
        if (r) return r;
    }
    return 0;
}

}  // namespace

#define DISPATCH(ctx, fn, ...) ((ctx)->p.precision == SPH_PREC_F64 ? fn<double>(__VA_ARGS__) : fn<float>(__VA_ARGS__))

extern "C" {

int64_t sph_arena_bytes(const SphParams *p, int64_t n_max) { return layout(p, n_max, nullptr); }

SphCtx *sph_create(const SphParams *p, int64_t n_max, void *arena, int64_t arena_bytes, void *stream) {
    if (!p || !arena || n_max <= 0 || n_max >= (1ll << 31)) return nullptr;
    if (n_cells(p) <= 0 || n_cells(p) >= (1ll << 31)) return nullptr;
    if (layout(p, n_max, nullptr) > arena_bytes) return nullptr;
    if (((uintptr_t)arena) & 255) return nullptr;
    SphCtx *c = new (std::nothrow) SphCtx();
    if (!c) return nullptr;
    memset(c, 0, sizeof(*c));
    c->p = *p;
    c->n_max = n_max;
    c->n = 0;
    c->arena = (char *)arena;
    c->arena_bytes = arena_bytes;
    c->stream = (cudaStream_t)stream;
    layout(p, n_max, c);
    c->own0 = 0; c->own1 = p->gn[0];
    c->r2thr64 = r2_threshold64(p->support);
    c->r2thr32 = r2_threshold32((float)p->support);
    if (cudaMemsetAsync(arena, 0, (size_t)layout(p, n_max, nullptr), c->stream) != cudaSuccess) { delete c; return nullptr; }
    return c;
}
void sph_destroy(SphCtx *c) {
    if (!c) return;
    if (c->prof_state) {
        ProfState *ps = (ProfState *)c->prof_state;
        for (cudaEvent_t e : ps->pool) cudaEventDestroy(e);
        delete ps;
    }
    slab_free(c);
    delete c;
}
const char *sph_last_error(SphCtx *c) { return c ? c->err : "null ctx"; }

int sph_set_params(SphCtx *c, const SphParams *p) {
    if (p->dim != c->p.dim || p->precision != c->p.precision || p->solver != c->p.solver || p->ti != c->p.ti ||
        p->kcorr != c->p.kcorr || n_cells(p) != c->C || p->fast != c->p.fast || p->boundary != c->p.boundary) {
        snprintf(c->err, sizeof(c->err), "sph_set_params cannot change sizes, solver, precision or buffers");
        return -2;
    }
    c->p = *p;
    c->r2thr64 = r2_threshold64(p->support);
    c->r2thr32 = r2_threshold32((float)p->support);
    c->masks_valid = false; c->gnl_valid = false;
    return 0;
}

int sph_field_info(SphCtx *c, int field, int64_t *offset_bytes, int32_t *ncomp, int32_t *stride, int32_t *kind) {
    if (field < 0 || field >= SPH_F_NUM) return -2;
    int src = field;
    int64_t shift = 0;
    if (field == SPH_F_MASS) { src = SPH_F_V; shift = 3 * c->real_bytes; }
    if (field == SPH_F_M_V) { src = SPH_F_XS; shift = 3 * c->real_bytes; }
    const FieldSlot &F = c->f[src];
    if (!F.present) return -3;
    *offset_bytes = F.off[F.cur] + shift;
    *ncomp = c->f[field].ncomp; *stride = c->f[field].stride; *kind = c->f[field].kind;
    return 0;
}

int sph_add_particles(SphCtx *c, int64_t n, const double *x, const double *v, const double *density, const int32_t *mat_type) {
    if (n <= 0) return 0;
    if (slab_armed(c)) { snprintf(c->err, sizeof(c->err), "sph_add_particles on a stepping slab: sph_clear_particles first"); return -2; }
    c->masks_valid = false; c->gnl_valid = false;
    if (c->n + n > c->n_max) { snprintf(c->err, sizeof(c->err), "particle capacity %lld exceeded", (long long)c->n_max); return -2; }
    const int64_t first = c->n;
    char *X = c->arena + c->f[SPH_F_X].off[c->f[SPH_F_X].cur];
    char *Xalt = c->arena + c->f[SPH_F_X].off[1 - c->f[SPH_F_X].cur];
    char *R = c->arena + c->f[SPH_F_DENSITY].off[c->f[SPH_F_DENSITY].cur];
    char *Ty = c->arena + c->f[SPH_F_MAT_TYPE].off[c->f[SPH_F_MAT_TYPE].cur];
    SPH_CHECK(c, cudaMemcpyAsync(X + first * 24, x, (size_t)n * 24, cudaMemcpyDefault, c->stream));
    SPH_CHECK(c, cudaMemcpyAsync(Xalt + first * 24, v, (size_t)n * 24, cudaMemcpyDefault, c->stream));
    SPH_CHECK(c, cudaMemcpyAsync(R + first * 8, density, (size_t)n * 8, cudaMemcpyDefault, c->stream));
    SPH_CHECK(c, cudaMemcpyAsync(Ty + first * 4, mat_type, (size_t)n * 4, cudaMemcpyDefault, c->stream));
    c->n += n;
    return DISPATCH(c, add_particles_finish, c, first, n);
}
int64_t sph_num_particles(SphCtx *c) { return slab_exact_n(c); }
// Only what a fresh upload does not overwrite has to be zero: the carried members of the live range (both ping-pong
// buffers, so that members sph_add_particles does not set -- v_tmp, pressure, stress ... -- start from 0 like the
// reference's zero-initialised fields) and the small counters; every other array is scratch rewritten before use.
int sph_clear_particles(SphCtx *c) {
    const int64_t live = c->slab ? c->n_max : c->n;
    slab_disarm(c);
    for (int f = 0; f < SPH_F_NUM; f++) {
        FieldSlot &F = c->f[f];
        if (!F.present || f == SPH_F_MASS || f == SPH_F_M_V || f == SPH_F_CELL_END || f == SPH_F_CELL_COUNT) continue;
        const int64_t bytes = live * F.stride * elem_bytes(F.kind, c->real_bytes);
        if (bytes == 0) continue;
        SPH_CHECK(c, cudaMemsetAsync(c->arena + F.off[0], 0, (size_t)bytes, c->stream));
        if (F.off[1] != F.off[0]) SPH_CHECK(c, cudaMemsetAsync(c->arena + F.off[1], 0, (size_t)bytes, c->stream));
    }
    SPH_CHECK(c, cudaMemsetAsync(c->arena + c->off_bad, 0, 512, c->stream));      // bad-cell counter, selection counts, SlabCtl
    c->n = 0;
    c->masks_valid = false; c->gnl_valid = false;
    c->list_valid = false; c->shep_pending = c->shep_wall_pending = false; c->fuse_half = c->fuse_init = false;
    for (int f = 0; f < SPH_F_NUM; f++) c->f[f].cur = 0;
    return 0;
}

int32_t sph_real_bytes(SphCtx *c) { return c->real_bytes; }
int sph_read_state_async(SphCtx *c, double *x, void *v, double *density, void *pressure, int32_t *id0) {
    const int64_t n = slab_exact_n(c);
    if (x) SPH_CHECK(c, cudaMemcpyAsync(x, c->arena + c->f[SPH_F_X].off[c->f[SPH_F_X].cur], (size_t)n * 24, cudaMemcpyDefault, c->stream));
    if (density) SPH_CHECK(c, cudaMemcpyAsync(density, c->arena + c->f[SPH_F_DENSITY].off[c->f[SPH_F_DENSITY].cur], (size_t)n * 8, cudaMemcpyDefault, c->stream));
    if (id0) SPH_CHECK(c, cudaMemcpyAsync(id0, c->arena + c->f[SPH_F_ID0].off[c->f[SPH_F_ID0].cur], (size_t)n * 4, cudaMemcpyDefault, c->stream));
    // v and pressure are engine-real; they are returned in the engine's own precision packed as stored
    // (v: n x 4 reals, pressure: n reals) when the engine is MIXED the caller passes float buffers.
    if (v) SPH_CHECK(c, cudaMemcpyAsync(v, c->arena + c->f[SPH_F_V].off[c->f[SPH_F_V].cur], (size_t)n * 4 * c->real_bytes, cudaMemcpyDefault, c->stream));
    if (pressure) SPH_CHECK(c, cudaMemcpyAsync(pressure, c->arena + c->f[SPH_F_PRESSURE].off[c->f[SPH_F_PRESSURE].cur], (size_t)n * c->real_bytes, cudaMemcpyDefault, c->stream));
    return 0;
}
int sph_synchronize(SphCtx *c) {
    SPH_CHECK(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int sph_read_state(SphCtx *c, double *x, void *v, double *density, void *pressure, int32_t *id0) {
    int r = sph_read_state_async(c, x, v, density, pressure, id0);
    return r ? r : sph_synchronize(c);
}

// on a native slab (sph_slab_init) the build starts with the migration / halo exchange and sorts [from L | own | from R]
int sph_grid_build(SphCtx *c) { return c->slab ? DISPATCH(c, slab_redistribute, c) : DISPATCH(c, grid_build, c); }
int sph_calc_kernel_corr(SphCtx *c) { return DISPATCH(c, calc_kernel_corr, c, true); }
int sph_calc_kernel_corr_deferred(SphCtx *c) { return DISPATCH(c, calc_kernel_corr, c, false); }
int sph_init_real2tmp(SphCtx *c) { return DISPATCH(c, init_real2tmp, c); }
int sph_one_step(SphCtx *c) { return DISPATCH(c, one_step, c, false); }
int sph_advect(SphCtx *c, int kind, int m) { return DISPATCH(c, advect, c, kind, m); }
int sph_advect_pos(SphCtx *c) { return DISPATCH(c, advect_pos, c); }
int sph_post_step(SphCtx *c) { return DISPATCH(c, post_step, c); }
int sph_enforce_boundary(SphCtx *c) { return DISPATCH(c, enforce_boundary, c); }
int sph_set_rigid_bodies(SphCtx *c, int64_t n_ids, const int32_t *body_of_id0_dev, const double *x0_of_id0_dev, int32_t n_bodies) {
    if (n_bodies < 0 || n_bodies > RIG_MAX || (n_bodies > 0 && (!body_of_id0_dev || !x0_of_id0_dev || n_ids <= 0))) {
        snprintf(c->err, sizeof(c->err), "sph_set_rigid_bodies: at most %d dynamic rigid bodies, tables indexed by id0", RIG_MAX); return -2;
    }
    if (n_bodies > 0 && c->p.solver == SPH_SOLVER_WC) {
        snprintf(c->err, sizeof(c->err), "dynamic rigid bodies under WCSPH do not run in the reference either (wc:129-132 calls a kernel from kernel scope)"); return -2;
    }
    if (n_bodies > 0 && c->slab) { snprintf(c->err, sizeof(c->err), "dynamic rigid bodies are not supported on slabs (a body would span ranks)"); return -2; }
    c->rig_obj = body_of_id0_dev; c->rig_x0 = x0_of_id0_dev; c->rig_n = n_bodies;
    return 0;
}
int sph_init_rigid_body(SphCtx *c) { return DISPATCH(c, init_rigid_body, c); }
int sph_solve_rigid_body(SphCtx *c) { return DISPATCH(c, solve_rigid_body, c); }
int sph_rigid_rest_cm(SphCtx *c, double *out_host) {
    for (int b = 0; b < c->rig_n; b++)
        SPH_CHECK(c, cudaMemcpyAsync(out_host + 3 * b, c->arena + c->off_rigid + (size_t)b * RIG_STRIDE * 8, 24, cudaMemcpyDeviceToHost, c->stream));
    SPH_CHECK(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int sph_init_stress(SphCtx *c) { return DISPATCH(c, init_stress, c, nullptr); }
int sph_init_stress_ymax(SphCtx *c, double ymax) { return DISPATCH(c, init_stress, c, &ymax); }
int sph_step(SphCtx *c, int nsteps) { return DISPATCH(c, dispatch_step, c, nsteps); }
int sph_neighbor_count(SphCtx *c, int32_t *out_dev) { return DISPATCH(c, neighbor_count, c, out_dev); }
int sph_density_sum(SphCtx *c, void *out_dev) { return DISPATCH(c, density_sum, c, out_dev); }
int sph_density_sweep(SphCtx *c, int32_t *count_dev, void *rho_dev) { return DISPATCH(c, density_sweep, c, count_dev, rho_dev); }
int sph_neighbor_count_masks(SphCtx *c, int32_t *out_dev) {
    if (!c->fast) { snprintf(c->err, sizeof(c->err), "neighbour masks exist only on the cell-tile path (MIXED precision WCSPH)"); return -2; }
    if (c->n == 0) return 0;
    return tile_mask_count(c, out_dev);
}

int64_t sph_read_bad_cells(SphCtx *c) {
    unsigned long long v = 0;
    if (cudaMemcpyAsync(&v, c->arena + c->off_bad, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -1;
    if (cudaMemsetAsync(c->arena + c->off_bad, 0, 8, c->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    return (int64_t)v;
}
// cells the cell-tile path handed to the generic kernels in the last mask build (a stencil cell with > 32 particles or a
// tile that overflowed); 0 when the fast path is not allocated.  Synchronises.
int64_t sph_read_flagged_cells(SphCtx *c) {
    if (!c->fast) return 0;
    int v = 0;
    if (cudaMemcpyAsync(&v, c->arena + c->off_nflag, 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    return v;
}
int64_t sph_launch_count(SphCtx *c) { return c->launches; }
int64_t sph_params_size(void) { return (int64_t)sizeof(SphParams); }

int sph_profile_enable(SphCtx *c, int on) {
    if (!c->prof_state) c->prof_state = new ProfState();
    c->prof_on = on != 0;
    return 0;
}
int sph_profile_num_kernels(void) { return K_NUM; }
const char *sph_profile_name(int id) { return (id >= 0 && id < K_NUM) ? KNAMES[id] : ""; }
int sph_profile_read(SphCtx *c, double *ms_by_kernel, int64_t *launches_by_kernel) {
    for (int k = 0; k < K_NUM; k++) { ms_by_kernel[k] = 0.0; launches_by_kernel[k] = 0; }
    ProfState *ps = (ProfState *)c->prof_state;
    if (!ps) return 0;
    SPH_CHECK(c, cudaStreamSynchronize(c->stream));
    for (size_t k = 0; k < ps->ids.size() && k < ps->used / 2; k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ps->pool[2 * k], ps->pool[2 * k + 1]) == cudaSuccess) {
            ms_by_kernel[ps->ids[k]] += ms;
            launches_by_kernel[ps->ids[k]]++;
        }
    }
    ps->ids.clear();
    ps->used = 0;
    return 0;
}
int sph_num_phases(SphCtx *c) { return c->p.solver == SPH_SOLVER_WC ? 2 : 3; }
int sph_one_step_phase(SphCtx *c, int phase) { return DISPATCH(c, one_step_phase, c, phase); }
int sph_set_owned_columns(SphCtx *c, int32_t b, int32_t e) {
    if (b < 0 || e > c->p.gn[0] || b >= e) { snprintf(c->err, sizeof(c->err), "owned columns [%d, %d) outside the grid", b, e); return -2; }
    c->own0 = b; c->own1 = e;
    return 0;
}

}  // extern "C"
