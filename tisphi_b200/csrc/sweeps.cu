// sweeps.cu -- generic (any solver, any dimension, stale grids allowed) neighbour sweeps, one thread per particle.
// Each kernel is one top-level loop of the reference's @ti.kernel bodies:
//   calc_CSPM_f / calc_CSPM_L          eng/solver_sph_base.py:386-423
//   WCSPH one_step loops A, B          eng/solver_sph_wc.py:82-126   (tasks wc:33-71, base:186-203, base:647-669)
//   mu(I) one_step loops 1-3           eng/solver_sph_muI.py:62-132
//   DP one_step loops 1-3              eng/solver_sph_dp.py:210-274  (return mapping dp:41-118, Bui 2008 dp:171-208)
//   advect_pos / XSPH                  eng/solver_sph_base.py:223-238
//   advect_something                   wc:129-132, muI:134-156, dp:276-296
// The cell-tile fast paths for the WCSPH hot loop live in sweeps_tile.cu; these kernels are the complete path.
#include "sph_host.h"

namespace sph {

template <typename T> static int rigid_reaction(SphCtx *c);     // dynamic rigid bodies: defined with their kernels below

// --------------------------------------------------------------------------------------------- kernel correction
// flagged-only mode: the cell-tile kernels handled every particle except those of flagged cells
template <typename T> __device__ __forceinline__ bool not_owned(const Dev<T> &c, int i) {
    if (c.own1 - c.own0 >= c.gn[0]) return false;
    const int cx = c.gid[i] / (c.gn[1] * c.gn[2]);
    return cx < c.own0 || cx >= c.own1;
}
template <typename T> __device__ __forceinline__ bool skip_unflagged(const Dev<T> &c, int i) {
    if (not_owned(c, i)) return true;
    if (!c.flagged_only) return false;
    if (*c.nflag == 0) return true;
    return c.cellflag[c.gid[i]] == 0;
}

// Flagged-only launches (the cell-tile kernels did everything else) are a fixed-size grid-stride grid that returns at
// once when no cell is flagged -- the common case -- instead of n / 128 blocks that each find nothing to do.
constexpr int FLAGGED_BLOCKS = 148 * 8;
#define SPH_PARTICLE_KERNEL(NAME, BODY)                                                            \
    template <typename T> __global__ void __launch_bounds__(128) NAME(Dev<T> c) {                  \
        if (c.flagged_only) {                                                                      \
            if (*c.nflag == 0) return;                                                             \
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.N(); i += gridDim.x * blockDim.x) BODY(c, i); \
        } else {                                                                                   \
            const int i = blockIdx.x * blockDim.x + threadIdx.x;                                   \
            if (i < c.N()) BODY(c, i);                                                               \
        }                                                                                          \
    }
template <typename T> inline int sweep_blocks(const Dev<T> &d) { return d.flagged_only ? FLAGGED_BLOCKS : blocks_for(d.n, 128); }

template <typename T> __device__ __forceinline__ void body_cspm_f(const Dev<T> &c, int i) {
    if (skip_unflagged(c, i)) return;
    T s = 0;
    for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) {
        if (is_flow(c.type[j])) s += Vj * kernel_W(c, r);
    });
    c.cspm_f[i] = (s != (T)0) ? (T)1 / s : (T)1;
}
SPH_PARTICLE_KERNEL(k_cspm_f, body_cspm_f)

template <typename T> __device__ __forceinline__ T det3(const T *m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
// calc_CSPM_L (base:400-423): M += V_j (x_j - x_i) (x) gradW_ij over same-type neighbours, then L = M^-1 (2x2 block in 2D)
template <typename T> __device__ __forceinline__ void cspm_L_add(const Dev<T> &c, T M[9], T dx, T dy, T dz, T r, T Vj) {
    const T s = kernel_dW_over_r(c, r);
    const T d[3] = {dx, dy, dz};
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) M[3 * a + b] += Vj * (-d[a]) * (s * d[b]);
}
template <typename T> __device__ __forceinline__ void cspm_L_store(const Dev<T> &c, int i, bool flow, const T M[9]) {
    T L[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (flow) {
        if (c.dim == 2) {
            const T det = M[0] * M[4] - M[1] * M[3];
            if (fabs(det) > c.eps) {
                const T inv = (T)1 / det;
                L[0] = M[4] * inv; L[1] = -M[1] * inv; L[2] = 0;
                L[3] = -M[3] * inv; L[4] = M[0] * inv; L[5] = 0;
                L[6] = 0; L[7] = 0; L[8] = 0;
            }
        } else {
            const T det = det3(M);
            if (fabs(det) > c.eps) {
                const T inv = (T)1 / det;
                L[0] = (M[4] * M[8] - M[5] * M[7]) * inv; L[1] = -(M[1] * M[8] - M[2] * M[7]) * inv; L[2] = (M[1] * M[5] - M[2] * M[4]) * inv;
                L[3] = -(M[3] * M[8] - M[5] * M[6]) * inv; L[4] = (M[0] * M[8] - M[2] * M[6]) * inv; L[5] = -(M[0] * M[5] - M[2] * M[3]) * inv;
                L[6] = (M[3] * M[7] - M[4] * M[6]) * inv; L[7] = -(M[0] * M[7] - M[1] * M[6]) * inv; L[8] = (M[0] * M[4] - M[1] * M[3]) * inv;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 9; a++) c.cspm_L[9 * (size_t)i + a] = L[a];
}
template <typename T> __device__ __forceinline__ void body_cspm_L(const Dev<T> &c, int i) {
    if (not_owned(c, i)) return;
    T M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int ti = c.type[i];
    const bool flow = is_flow(ti);
    if (flow)
        for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) {
            if (c.type[j] == ti) cspm_L_add(c, M, dx, dy, dz, r, Vj);
        });
    cspm_L_store(c, i, flow, M);
}
template <typename T> __global__ void __launch_bounds__(128) k_cspm_L(Dev<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.N()) body_cspm_L(c, i);
}

// The candidate walk of every particle, ONCE per step (positions are frozen from the grid build to advect_pos): it
// records the neighbours for the sweeps that follow (sph_dev.cuh::for_neighbors replays them) and, being the first sweep
// after the grid build anyway, forms the kernel correction on its way -- calc_CSPM_f (base:386-398) and, with CSPM,
// calc_CSPM_L (base:400-423): the same tasks in the same order as k_cspm_f / k_cspm_L.  A particle whose list does not
// fit, or whose position is no longer in the cell it is stored in, is marked -1 and walks the cells in every sweep.
template <typename T> __global__ void __launch_bounds__(128) k_corr_nlist(Dev<T> c, unsigned *__restrict__ nlist, int *__restrict__ ncount,
                                                                         int stride, int cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    int cc[3], sc[3];
    const double xi[3] = {c.x[3 * (size_t)i], c.x[3 * (size_t)i + 1], c.x[3 * (size_t)i + 2]};
    pos_to_cell(c, xi, cc);
    unflatten(c, c.gid[i], sc);
    const bool stored = cc[0] == sc[0] && cc[1] == sc[1] && cc[2] == sc[2];
    const int ti = c.type[i];
    const bool mine = !not_owned(c, i), want_L = c.kcorr == 1 && mine, flowL = want_L && is_flow(ti);
    T S = 0, M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    auto task = [&](int j, T dx, T dy, T dz, T r, T Vj) {
        const int tj = c.type[j];
        if (mine && is_flow(tj)) S += Vj * kernel_W(c, r);
        if (flowL && tj == ti) cspm_L_add(c, M, dx, dy, dz, r, Vj);
    };
    int cnt = -1;
    if (stored) {
        cnt = for_neighbors_walk<T, 2>(c, i, nlist + i, stride, cap, task);
        if (cnt > cap) cnt = -1;
    } else for_neighbors_walk<T, 0>(c, i, nullptr, 0, 0, task);
    ncount[i] = cnt;
    if (mine) c.cspm_f[i] = (S != (T)0) ? (T)1 / S : (T)1;
    if (want_L) cspm_L_store(c, i, flowL, M);
}

// standalone: every CSPM_f is final on return (the API call).  Otherwise (inside sph_step) the tile path leaves the
// Shepard sums of unflagged cells to the wall pass and the first fluid pass, which visit the same neighbours anyway.
template <typename T> int calc_kernel_corr(SphCtx *c, bool standalone) {
    if (c->n == 0) return 0;
    c->gnl_valid = false;
    if (!c->fast && c->off_gnl) {        // generic sweeps with per-step lists: correction and lists in one walk
        Dev<T> d = make_dev<T>(c);
        SPH_PROF(c, K_NLIST);
        k_corr_nlist<T><<<blocks_for(c->n, 128), 128, 0, c->stream>>>(d, (unsigned *)(c->arena + c->off_gnl), (int *)(c->arena + c->off_gnl_count),
                                                                      (int)c->n_max, c->gnl_cap);
        SPH_LAUNCH_CHECK(c);
        c->gnl_valid = true;
        return 0;
    }
    Dev<T> d = make_dev<T>(c);
    if (c->fast) {                       // tile path: masks (+ CSPM_f); the generic kernel completes flagged cells
        int r = tile_mask(c, standalone);
        if (r) return r;
        d.flagged_only = 1;
    }
    SPH_PROF(c, K_CSPM_F);
    k_cspm_f<T><<<sweep_blocks(d), 128, 0, c->stream>>>(d);
    SPH_LAUNCH_CHECK(c);
    if (c->p.kcorr == 1) {
        SPH_PROF(c, K_CSPM_L);
        k_cspm_L<T><<<blocks_for(c->n, 128), 128, 0, c->stream>>>(d);
        SPH_LAUNCH_CHECK(c);
    }
    return 0;
}

// corrected kernel gradient (base:374-383): g = s*d, gc = g | L_i g | 0
template <typename T>
__device__ __forceinline__ void grad_corr(const Dev<T> &c, const T *L, T s, T dx, T dy, T dz, T g[3], T gc[3]) {
    g[0] = s * dx; g[1] = s * dy; g[2] = s * dz;
    if (c.kcorr == 0) { gc[0] = g[0]; gc[1] = g[1]; gc[2] = g[2]; }
    else if (c.kcorr == 1) {
#pragma unroll
        for (int a = 0; a < 3; a++) gc[a] = L[3 * a] * g[0] + L[3 * a + 1] * g[1] + L[3 * a + 2] * g[2];
    } else { gc[0] = gc[1] = gc[2] = 0; }
}
template <typename T> __device__ __forceinline__ void load_L(const Dev<T> &c, int i, T L[9]) {
    if (c.kcorr == 1) {
#pragma unroll
        for (int a = 0; a < 9; a++) L[a] = c.cspm_L[9 * (size_t)i + a];
    } else {
#pragma unroll
        for (int a = 0; a < 9; a++) L[a] = 0;
    }
}

// --------------------------------------------------------------------------------------------------- WCSPH
__device__ __forceinline__ double eos_wc(double rho, double rho0, double stiff, double gamma_) {
    double v = stiff * (pow(rho / rho0, gamma_) - 1.0);      // wc:87-88
    return v > 0.0 ? v : 0.0;
}
// loop A, fluid branch: pointwise EOS into the NEW pressure buffer
template <typename T> __global__ void __launch_bounds__(256) k_wc_eos(Dev<T> c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    const int t = c.type[i];
    if (is_fluid(t)) c.pnew[i] = (T)eos_wc(c.rho_t[i], c.rho0, c.stiff, c.gamma_);
    else if (!is_wall(t)) c.pnew[i] = c.press[i];
}
// loop A, wall branch (wc:90-103).  p_j is read "in place" by the reference; serial semantics are reproduced
// pointwise: fluid j < i already holds EOS(rho~_j) (pnew), fluid j > i still holds the previous pressure (press).
// Static rigid particles (type 11, SURVEY 8 f2: an indenter with a prescribed velocity) are wall particles in loop A
// and keep d_vel = 0 (wc:125-126, muI:131-132, dp:273-274) and the d_density = 0 they were created with; the
// integrators treat them as real particles (base:79-170), so both are written here -- the derivative arrays are
// scratch that does not travel through the sort.
// (a DYNAMIC rigid particle starts every one_step from d_vel = g instead -- wc:105-106, muI:111-112, dp:233-234 -- and
// k_rigid_reaction subtracts the momentum terms it takes part in)
template <typename T> __device__ __forceinline__ void zero_rigid_derivatives(const Dev<T> &c, int i) {
    c.d_rho[i] = 0;
    Vec4<T> z; z.x = z.y = z.z = z.w = 0;
    if (rigid_body_of(c, i) >= 0) { z.x = c.g[0]; z.y = c.g[1]; z.z = c.g[2]; }
    c.d_vel[i] = z;
}
template <typename T> __device__ __forceinline__ void body_wc_wall(const Dev<T> &c, int i) {
    if (!is_wall(c.type[i])) return;
    if (is_rigid(c.type[i]) && !c.flagged_only) zero_rigid_derivatives(c, i);      // the tile path does it in k_tile_prep
    if (skip_unflagged(c, i)) return;
    T Sv0 = 0, Sv1 = 0, Sv2 = 0, Sp = 0;
    for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) {
        const int tj = c.type[j];
        if (is_flow(tj)) {
            const T w = kernel_W(c, r);
            const Vec4<T> vj = c.vt4[j];
            Sv0 += Vj * vj.x * w; Sv1 += Vj * vj.y * w; Sv2 += Vj * vj.z * w;
            const T pj = (is_fluid(tj) && (c.wc_fresh || j < i)) ? c.pnew[j] : c.press[j];
            Sp += Vj * (pj + vj.w * c.g[1] * dy) * w;
        }
    });
    const T f = c.cspm_f[i];
    const Vec4<T> v = c.v4[i];
    Vec4<T> vt;
    vt.x = (T)2 * v.x - Sv0 * f; vt.y = (T)2 * v.y - Sv1 * f; vt.z = (T)2 * v.z - Sv2 * f;
    vt.w = c.rho0T;
    c.vt4[i] = vt;
    c.rho_t[i] = c.rho0;
    const T p = Sp * f;
    const T pc = p > (T)0 ? p : (T)0;
    c.pnew[i] = pc;
    if (c.pk4) { Vec4<T> pk = vt; pk.w = pc / (c.rho0T * c.rho0T); c.pk4[i] = pk; }
}
SPH_PARTICLE_KERNEL(k_wc_wall, body_wc_wall)
// calc_repulsive_force (base:675-689) of a repulsive particle at d = x_i - x_j, |d| = r
template <typename T> __device__ __forceinline__ void rep_force(const Dev<T> &c, T dx, T dy, T dz, T r, T &a0, T &a1, T &a2) {
    const T chi = (r > (T)0 && r < c.rep_judge) ? (T)1 - r / c.rep_judge : (T)0;
    const T gamma = r * c.rep_ginv;
    T f = 0;
    if (gamma > (T)0 && gamma <= (T)(2.0 / 3.0)) f = (T)(2.0 / 3.0);
    else if (gamma > (T)(2.0 / 3.0) && gamma <= (T)1) f = (T)2 * gamma - (T)1.5 * gamma * gamma;
    else if (gamma > (T)1 && gamma < (T)2) f = (T)0.5 * ((T)2 - gamma) * ((T)2 - gamma);
    const T k = c.rep_k * chi * f / (r * r);
    a0 += k * dx; a1 += k * dy; a2 += k * dz;
}
template <typename T> __device__ __forceinline__ bool has_rep(const Dev<T> &c) { return c.boundary == 3 || c.boundary == 4; }
// loop B (wc:108-126): continuity (corrected gradient) + viscosity + pressure (plain gradient, H22)
template <typename T> __device__ __forceinline__ void body_wc_fluid(const Dev<T> &c, int i) {
    if (!is_fluid(c.type[i])) return;
    if (skip_unflagged(c, i)) return;
    const Vec4<T> vi = c.vt4[i];
    const T rhoi = vi.w, pi = c.press[i];
    const T pri = pi / (rhoi * rhoi);
    T L[9];
    load_L(c, i, L);
    T dd = 0, a0 = 0, a1 = 0, a2 = 0;
    for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) {
        const T s = kernel_dW_over_r(c, r);
        T g[3], gc[3];
        grad_corr(c, L, s, dx, dy, dz, g, gc);
        const Vec4<T> vj = c.vt4[j];
        const T ux = vi.x - vj.x, uy = vi.y - vj.y, uz = vi.z - vj.z;
        dd += Vj * ux * gc[0] + Vj * uy * gc[1] + Vj * uz * gc[2];
        const int tj = c.type[j];
        const T vx = ux * dx + uy * dy + uz * dz;
        const T mn = vx < (T)0 ? vx : (T)0;
        T visc = 0;
        if (is_fluid(tj)) visc = c.visc_coef * Vj * mn / (r * r + c.h2_001);
        else if (is_wall(tj)) visc = c.visc_coef * Vj * c.rho0T / rhoi * mn / (r * r + c.h2_001);
        const T rhoj = vj.w;
        const T pres = -c.rho0T * Vj * (pri + c.press[j] / (rhoj * rhoj));
        a0 += visc * g[0] + pres * g[0]; a1 += visc * g[1] + pres * g[1]; a2 += visc * g[2] + pres * g[2];
    });
    if (has_rep(c))                                              // wc:119-121: a second walk, after the first, like the reference
        for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) { if (is_rep(c.type[j])) rep_force(c, dx, dy, dz, r, a0, a1, a2); });
    c.d_rho[i] = dd * rhoi;
    Vec4<T> dv; dv.x = a0 + c.g[0]; dv.y = a1 + c.g[1]; dv.z = a2 + c.g[2]; dv.w = 0;
    c.d_vel[i] = dv;
}
SPH_PARTICLE_KERNEL(k_wc_fluid, body_wc_fluid)

// ------------------------------------------------------------------------------------------ soil: shared pieces
template <typename T> __device__ __forceinline__ T dev_component(const T *t) {     // type_define.py:26-28
    T s = 0;
#pragma unroll
    for (int a = 0; a < 9; a++) s += t[a] * t[a];
    return sqrt(s * (T)2 / (T)3);
}
// Adami wall extrapolation for soil solvers (muI:95-109, dp:220-231; tasks base:647-669)
template <typename T> __device__ __forceinline__ void body_soil_wall(const Dev<T> &c, int i) {
    if (!is_wall(c.type[i])) return;
    if (is_rigid(c.type[i])) zero_rigid_derivatives(c, i);
    if (not_owned(c, i)) return;
    T Sv0 = 0, Sv1 = 0, Sv2 = 0, Sr = 0, Ss[6] = {0, 0, 0, 0, 0, 0};
    for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) {
        if (is_flow(c.type[j])) {
            const T w = kernel_W(c, r);
            const Vec4<T> vj = c.vt4[j];
            const T rj = vj.w;
            Sv0 += Vj * vj.x * w; Sv1 += Vj * vj.y * w; Sv2 += Vj * vj.z * w;
            Sr += Vj * rj * w;
            const T *sj = c.stress_t + 6 * (size_t)j;
            Ss[0] += Vj * (sj[0] + rj * c.g[0] * dx) * w;
            Ss[1] += Vj * (sj[1] + rj * c.g[1] * dy) * w;
            Ss[2] += Vj * (sj[2] + rj * c.g[2] * dz) * w;
            Ss[3] += Vj * sj[3] * w; Ss[4] += Vj * sj[4] * w; Ss[5] += Vj * sj[5] * w;
        }
    });
    const T f = c.cspm_f[i];
    const Vec4<T> v = c.v4[i];
    Vec4<T> vt;
    vt.x = (T)2 * v.x - Sv0 * f; vt.y = (T)2 * v.y - Sv1 * f; vt.z = (T)2 * v.z - Sv2 * f;
    T rt = c.rho0T;
    if (c.solver == SPH_SOLVER_MUI) { const T e = Sr * f; rt = e > c.rho0T ? e : c.rho0T; }   // muI:105
    vt.w = rt;
    c.vt4[i] = vt;
    c.rho_t[i] = (double)rt;
#pragma unroll
    for (int q = 0; q < 6; q++) c.stress_t[6 * (size_t)i + q] = Ss[q] * f;
}
template <typename T> __global__ void __launch_bounds__(128) k_soil_wall(Dev<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.N()) body_soil_wall(c, i);
}

// stress_tmp / density_tmp^2 of every particle (walls and ghosts included), the neighbour payload of the momentum sum
template <typename T> __global__ void __launch_bounds__(256) k_soil_sor(Dev<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    const T rho = c.vt4[i].w, r2 = rho * rho;
    const T *s = c.stress_t + 6 * (size_t)i;
    T *o = c.sor + 6 * (size_t)i;
#pragma unroll
    for (int q = 0; q < 6; q++) o[q] = s[q] / r2;
}
// accumulators of one soil sweep: velocity gradient, continuity sum, momentum sum
template <typename T, bool VG, bool MOM>
__device__ __forceinline__ void soil_sweep(const Dev<T> &c, int i, T vg[9], T *dd, T mom[3]) {
    const Vec4<T> vi = c.vt4[i];
    const T rhoi = vi.w;
    T L[9], si[9];
    load_L(c, i, L);
    T sir[9];                                         // sigma~_i / rho~_i^2, the j-independent quotients of the momentum sum
    if (MOM) {
        sym_load(c.stress_t, (size_t)i, si);
        const T r2i = rhoi * rhoi;
#pragma unroll
        for (int a = 0; a < 9; a++) sir[a] = si[a] / r2i;
    }
    T acc = 0;
#pragma unroll
    for (int a = 0; a < 9; a++) vg[a] = 0;
    mom[0] = mom[1] = mom[2] = 0;
    for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) {
        const T s = kernel_dW_over_r(c, r);
        T g[3], gc[3];
        grad_corr(c, L, s, dx, dy, dz, g, gc);
        const Vec4<T> vj = c.vt4[j];
        if (VG) {
            const T u[3] = {vj.x - vi.x, vj.y - vi.y, vj.z - vi.z};
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) vg[3 * a + b] += Vj * u[a] * gc[b];
            acc += Vj * (-u[0]) * gc[0] + Vj * (-u[1]) * gc[1] + Vj * (-u[2]) * gc[2];
        }
        if (MOM) {      // muI:38-46 / dp:156-165: V_j rho~_j (sigma~_j / rho~_j^2 + sigma~_i / rho~_i^2) . gradW^c
            // The expression divides 18 times per neighbour.  The quotients sigma~_j / rho~_j^2 are a property of particle j
            // alone: k_soil_sor forms them once per particle right before this sweep (the 6 components of the symmetric
            // tensor), the particle's own are formed before the loop.  Same quotients, same products, same order of sums:
            // bit-identical to the expression as written.  (A reciprocal-and-multiply form is cheaper still but multiplies
            // the 30-step velocity error of BASELINE config C3 by 20 -- the sum cancels to ~1e-4 of its terms.)
            const T cf = Vj * vj.w;
            const T *q = c.sor + 6 * (size_t)j;
            const T mxx = cf * (q[0] + sir[0]), myy = cf * (q[1] + sir[4]), mzz = cf * (q[2] + sir[8]);
            const T mxy = cf * (q[3] + sir[1]), myz = cf * (q[4] + sir[5]), mzx = cf * (q[5] + sir[2]);
            T t0 = 0, t1 = 0, t2 = 0;
            t0 += mxx * gc[0]; t0 += mxy * gc[1]; t0 += mzx * gc[2];
            t1 += mxy * gc[0]; t1 += myy * gc[1]; t1 += myz * gc[2];
            t2 += mzx * gc[0]; t2 += myz * gc[1]; t2 += mzz * gc[2];
            mom[0] += t0; mom[1] += t1; mom[2] += t2;
        }
    });
    *dd = acc;
}

// ----------------------------------------------------------------------------------------------------- mu(I)
template <typename T> __device__ __forceinline__ void body_mui_soil1(const Dev<T> &c, int i) {          // muI:67-92
    if (!is_soil(c.type[i])) return;
    if (not_owned(c, i)) return;
    T vg[9], dd, mom[3];
    soil_sweep<T, true, false>(c, i, vg, &dd, mom);
#pragma unroll
    for (int a = 0; a < 9; a++) c.v_grad[9 * (size_t)i + a] = vg[a];
    c.d_rho[i] = dd * c.vt4[i].w;
    double pv = c.vsound * c.vsound * (c.rho[i] - c.rho0);      // rho, not rho~ (muI:80, H17)
    if (pv < 0.0) pv = 0.0;
    const T p = (T)pv;
    c.press[i] = p;
    T sr[9], s2 = 0;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) sr[3 * a + b] = (T)0.5 * (vg[3 * a + b] + vg[3 * b + a]);
#pragma unroll
    for (int a = 0; a < 9; a++) s2 += sr[a] * sr[a];
    const T dbdot = sqrt((T)0.5 * s2) + c.eps;
    const T coef = (T)0 + (c.coh + p * c.mu) / dbdot;             // eta_0 = 0 (muI:19)
    T st[9];
#pragma unroll
    for (int a = 0; a < 9; a++) st[a] = coef * sr[a];
    st[0] -= p; st[4] -= p; st[8] -= p;
    sym_store(c.stress_t, (size_t)i, st);
    const T tr = sr[0] + sr[4] + sr[8];
    sr[0] -= tr / (T)3; sr[4] -= tr / (T)3; sr[8] -= tr / (T)3;
    c.d_strain[i] = dev_component(sr);
}
template <typename T> __global__ void __launch_bounds__(128) k_mui_soil1(Dev<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.N()) body_mui_soil1(c, i);
}
template <typename T> __device__ __forceinline__ void body_mui_soil3(const Dev<T> &c, int i) {          // muI:115-128
    if (!is_soil(c.type[i])) return;
    if (not_owned(c, i)) return;
    T vg[9], dd, mom[3];
    soil_sweep<T, false, true>(c, i, vg, &dd, mom);
    if (has_rep(c))                                              // muI:121-123
        for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) { if (is_rep(c.type[j])) rep_force(c, dx, dy, dz, r, mom[0], mom[1], mom[2]); });
    const Vec4<T> vi = c.vt4[i];
    const T dc = c.damp_c / sqrt(vi.w);                            // -5e-5 sqrt(E / (rho~ h^2)) (base:713-715)
    Vec4<T> dv;
    dv.x = mom[0] + c.g[0] + dc * vi.x; dv.y = mom[1] + c.g[1] + dc * vi.y; dv.z = mom[2] + c.g[2] + dc * vi.z; dv.w = 0;
    c.d_vel[i] = dv;
}
template <typename T> __global__ void __launch_bounds__(128) k_mui_soil3(Dev<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.N()) body_mui_soil3(c, i);
}

// -------------------------------------------------------------------------------------------- Drucker-Prager
template <typename T> __device__ __forceinline__ void from_stress(const Dev<T> &c, const T *s, T dev[9], T *I1, T *sJ2, T *f) {
    const T tr = s[0] + s[4] + s[8];
    T s2 = 0;
#pragma unroll
    for (int a = 0; a < 9; a++) dev[a] = s[a];
    dev[0] -= tr / (T)3; dev[4] -= tr / (T)3; dev[8] -= tr / (T)3;
#pragma unroll
    for (int a = 0; a < 9; a++) s2 += dev[a] * dev[a];
    *I1 = tr; *sJ2 = sqrt((T)0.5 * s2); *f = *sJ2 + c.alpha * tr - c.kc;
}
template <typename T> __device__ __forceinline__ void adapt_stress(const Dev<T> &c, T *s) {   // dp:73-96
    T dev[9], I1, sJ2, f;
    from_stress(c, s, dev, &I1, &sJ2, &f);
    if (f > c.eps_f) {
        if (f > sJ2) {
            const T tmp = (I1 - c.kc / c.alpha) / (T)3;
            s[0] -= tmp; s[4] -= tmp; s[8] -= tmp;
        }
        from_stress(c, s, dev, &I1, &sJ2, &f);
        const T rr = (-I1 * c.alpha + c.kc) / sJ2;
#pragma unroll
        for (int a = 0; a < 9; a++) s[a] = rr * dev[a];
        s[0] += I1 / (T)3; s[4] += I1 / (T)3; s[8] += I1 / (T)3;
    }
}
template <typename T> __device__ __forceinline__ int flag_dp(const Dev<T> &c, const T *s) {   // dp:98-109
    T dev[9], I1, sJ2, f;
    from_stress(c, s, dev, &I1, &sJ2, &f);
    if (f < -c.eps_f) return 0;
    if (f > c.eps_f) return (f >= sJ2) ? 3 : 2;
    return 1;
}
template <typename T> __global__ void __launch_bounds__(256) k_dp_adapt(Dev<T> c) {           // dp:215-217
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    if (!is_soil(c.type[i])) return;
    T s[9];
    sym_load(c.stress_t, (size_t)i, s);
    adapt_stress(c, s);
    sym_store(c.stress_t, (size_t)i, s);
}
template <typename T>
__device__ __forceinline__ void bui2008(const Dev<T> &c, const T *st, const T *vg, T ds[9], T *dse, T *dsep) {  // dp:171-208
    T dev[9], I1, sJ2, f, sr[9], sp[9], J[9], se[9], te[9], tg[9], lam = 0;
    from_stress(c, st, dev, &I1, &sJ2, &f);
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            sr[3 * a + b] = (T)0.5 * (vg[3 * a + b] + vg[3 * b + a]);
            sp[3 * a + b] = (T)0.5 * (vg[3 * a + b] - vg[3 * b + a]);
        }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            J[3 * i + j] = st[3 * i + 0] * sp[3 * j + 0] + st[3 * i + 1] * sp[3 * j + 1] + st[3 * i + 2] * sp[3 * j + 2] +
                           st[0 + j] * sp[3 * i + 0] + st[3 + j] * sp[3 * i + 1] + st[6 + j] * sp[3 * i + 2];
    const T tr = sr[0] + sr[4] + sr[8];
#pragma unroll
    for (int a = 0; a < 9; a++) se[a] = sr[a];
    se[0] -= tr / (T)3; se[4] -= tr / (T)3; se[8] -= tr / (T)3;
#pragma unroll
    for (int a = 0; a < 9; a++) { te[a] = (T)2 * c.G * se[a]; tg[a] = 0; }
    te[0] += c.K * tr; te[4] += c.K * tr; te[8] += c.K * tr;
    const bool plastic = (f >= -c.eps_f && sJ2 > c.eps);
    if (plastic) {
        T ss = 0;
#pragma unroll
        for (int a = 0; a < 9; a++) ss += dev[a] * sr[a];
        lam = ((T)3 * c.alpha * c.K * tr + (c.G / sJ2) * ss) / ((T)27 * c.alpha * c.K * c.sin_dila + c.G);
#pragma unroll
        for (int a = 0; a < 9; a++) tg[a] = lam * (c.G / sJ2 * dev[a]);
        tg[0] = lam * ((T)9 * c.K * c.sin_dila + c.G / sJ2 * dev[0]);
        tg[4] = lam * ((T)9 * c.K * c.sin_dila + c.G / sJ2 * dev[4]);
        tg[8] = lam * ((T)9 * c.K * c.sin_dila + c.G / sJ2 * dev[8]);
    }
#pragma unroll
    for (int a = 0; a < 9; a++) ds[a] = J[a] + te[a] - tg[a];
    *dse = dev_component(se);
    *dsep = 0;
    if (plastic) {
        const T gp = sJ2 + (T)3 * I1 * c.sin_dila;
        T ep[9];
#pragma unroll
        for (int a = 0; a < 9; a++) ep[a] = (fabs(ds[a]) > c.eps ? gp / ds[a] : (T)0) * lam;   // pti.g_p is never written
        const T trp = ep[0] + ep[4] + ep[8];
        ep[0] -= trp / (T)3; ep[4] -= trp / (T)3; ep[8] -= trp / (T)3;
        *dsep = dev_component(ep);
    }
}
template <typename T> __device__ __forceinline__ void body_dp_soil(const Dev<T> &c, int i) {            // dp:237-270
    if (!is_soil(c.type[i])) return;
    if (not_owned(c, i)) return;
    T vg[9], dd, mom[3];
    soil_sweep<T, true, true>(c, i, vg, &dd, mom);
#pragma unroll
    for (int a = 0; a < 9; a++) c.v_grad[9 * (size_t)i + a] = vg[a];
    const Vec4<T> vi = c.vt4[i];
    c.d_rho[i] = dd * vi.w;
    T st[9], ds[9], dse, dsep;
    sym_load(c.stress_t, (size_t)i, st);
    bui2008(c, st, vg, ds, &dse, &dsep);
    sym_store(c.d_stress, (size_t)i, ds);
    c.d_strain[i] = dse;
    c.d_strain_p[i] = dsep;
    const T dc = c.damp_c / sqrt(vi.w);
    Vec4<T> dv;
    dv.x = mom[0] + c.g[0] + dc * vi.x; dv.y = mom[1] + c.g[1] + dc * vi.y; dv.z = mom[2] + c.g[2] + dc * vi.z; dv.w = 0;
    c.d_vel[i] = dv;
}
template <typename T> __global__ void __launch_bounds__(128, sizeof(T) == 4 ? 7 : 4) k_dp_soil(Dev<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.N()) body_dp_soil(c, i);
}

// one top-level loop of <Solver>.one_step (phases documented in include/tisphi_b200.h: sph_one_step_phase)
template <typename T> int one_step_phase(SphCtx *c, int phase) {
    if (c->n == 0) return 0;
    const int n = (int)c->n;
    cudaStream_t st = c->stream;
    const int solver = c->p.solver;
    if (phase < 0 || phase >= (solver == SPH_SOLVER_WC ? 2 : 3)) {
        snprintf(c->err, sizeof(c->err), "solver %d has no phase %d", solver, phase);
        return -2;
    }
    if (solver == SPH_SOLVER_WC && c->fast) {
        if (phase == 0) {
            int r = tile_wc_prep_and_wall(c);
            if (r) return r;
            Dev<T> d = make_dev<T>(c);
            d.flagged_only = 1;
            SPH_PROF(c, K_WC_WALL);
            k_wc_wall<T><<<sweep_blocks(d), 128, 0, st>>>(d);
            SPH_LAUNCH_CHECK(c);
            flip(c, SPH_F_PRESSURE);
        } else {
            int r = tile_wc_fluid(c);
            if (r) return r;
            Dev<T> d = make_dev<T>(c);
            d.flagged_only = 1;
            SPH_PROF(c, K_WC_FLUID);
            k_wc_fluid<T><<<sweep_blocks(d), 128, 0, st>>>(d);
            SPH_LAUNCH_CHECK(c);
        }
    } else if (solver == SPH_SOLVER_WC) {
        Dev<T> d = make_dev<T>(c);
        if (phase == 0) {
            SPH_PROF(c, K_WC_EOS);
            k_wc_eos<T><<<blocks_for(n, 256), 256, 0, st>>>(d);
            SPH_LAUNCH_CHECK(c);
            SPH_PROF(c, K_WC_WALL);
            k_wc_wall<T><<<sweep_blocks(d), 128, 0, st>>>(d);
            SPH_LAUNCH_CHECK(c);
            flip(c, SPH_F_PRESSURE);               // pnew becomes pt.pressure
        } else {
            SPH_PROF(c, K_WC_FLUID);
            k_wc_fluid<T><<<sweep_blocks(d), 128, 0, st>>>(d);
            SPH_LAUNCH_CHECK(c);
        }
    } else if (solver == SPH_SOLVER_MUI) {
        Dev<T> d = make_dev<T>(c);
        if (phase == 0) {
            SPH_PROF(c, K_MUI_SOIL1);
            k_mui_soil1<T><<<blocks_for(n, 128), 128, 0, st>>>(d);
        } else if (phase == 1) {
            SPH_PROF(c, K_SOIL_WALL);
            k_soil_wall<T><<<blocks_for(n, 128), 128, 0, st>>>(d);
        } else {
            SPH_PROF(c, K_OTHER);
            k_soil_sor<T><<<blocks_for(n, 256), 256, 0, st>>>(d);
            SPH_LAUNCH_CHECK(c);
            SPH_PROF(c, K_MUI_SOIL3);
            k_mui_soil3<T><<<blocks_for(n, 128), 128, 0, st>>>(d);
            SPH_LAUNCH_CHECK(c);
            return rigid_reaction<T>(c);
        }
        SPH_LAUNCH_CHECK(c);
    } else if (solver == SPH_SOLVER_DP) {
        Dev<T> d = make_dev<T>(c);
        if (phase == 0) {
            SPH_PROF(c, K_DP_ADAPT);
            k_dp_adapt<T><<<blocks_for(n, 256), 256, 0, st>>>(d);
        } else if (phase == 1) {
            SPH_PROF(c, K_SOIL_WALL);
            k_soil_wall<T><<<blocks_for(n, 128), 128, 0, st>>>(d);
        } else {
            SPH_PROF(c, K_OTHER);
            k_soil_sor<T><<<blocks_for(n, 256), 256, 0, st>>>(d);
            SPH_LAUNCH_CHECK(c);
            SPH_PROF(c, K_DP_SOIL);
            k_dp_soil<T><<<blocks_for(n, 128), 128, 0, st>>>(d);
            SPH_LAUNCH_CHECK(c);
            return rigid_reaction<T>(c);
        }
        SPH_LAUNCH_CHECK(c);
    } else {
        snprintf(c->err, sizeof(c->err), "unknown solver %d", solver);
        return -2;
    }
    return 0;
}
// `last`: the last one_step of a step (multi-GPU slabs skip the ghost refresh of derivatives nobody reads any more)
template <typename T> int one_step(SphCtx *c, bool last) {
    const int np = c->p.solver == SPH_SOLVER_WC ? 2 : 3;
    for (int ph = 0; ph < np; ph++) {
        int r = one_step_phase<T>(c, ph);
        if (r) return r;
        if ((r = slab_refresh(c, ph, ph == np - 1, last))) return r;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------- advect_pos (base:228-238)
template <typename T> __global__ void __launch_bounds__(256) k_advect_pos(Dev<T> c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    if (!is_real(c.type[i])) return;
    const Vec4<T> v = c.v4[i];
    double *x = c.x + 3 * (size_t)i;
    x[0] += c.dt * (double)v.x; x[1] += c.dt * (double)v.y; x[2] += c.dt * (double)v.z;
}
// XSPH on a snapshot: new positions go to the alternate x buffer
template <typename T> __device__ __forceinline__ void body_advect_pos_xsph(const Dev<T> &c, int i) {
    double *xnew = c.xnew;
    const double *x = c.x + 3 * (size_t)i;
    double o0 = x[0], o1 = x[1], o2 = x[2];
    const int ti = c.type[i];
    if (is_real(ti)) {
        const Vec4<T> vi = c.v4[i];
        T s0 = 0, s1 = 0, s2 = 0;
        if (is_dynamic(c, i))                                        // base:234: XSPH only moves dynamic particles
            for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) {
                if (c.type[j] == ti) {
                    const T w = kernel_W(c, r);
                    const Vec4<T> vj = c.v4[j];
                    s0 += Vj * (vj.x - vi.x) * w; s1 += Vj * (vj.y - vi.y) * w; s2 += Vj * (vj.z - vi.z) * w;
                }
            });
        o0 += c.dt * (double)(vi.x + (T)0.5 * s0); o1 += c.dt * (double)(vi.y + (T)0.5 * s1); o2 += c.dt * (double)(vi.z + (T)0.5 * s2);
    }
    xnew[3 * (size_t)i] = o0; xnew[3 * (size_t)i + 1] = o1; xnew[3 * (size_t)i + 2] = o2;
}
template <typename T> __global__ void __launch_bounds__(128) k_advect_pos_xsph(Dev<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.N()) body_advect_pos_xsph(c, i);
}
// after positions moved on a stale grid: refresh the sweep coordinates relative to the cell each particle is STORED in
template <typename T> __global__ void __launch_bounds__(256) k_refresh_xs(Dev<T> c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    const double *x = c.x + 3 * (size_t)i;
    Vec4<T> xs = c.xs4[i];
    if (sizeof(T) == 8) { xs.x = (T)x[0]; xs.y = (T)x[1]; xs.z = (T)x[2]; }
    else {
        int cc[3];
        unflatten(c, c.gid[i], cc);
        xs.x = (T)__dsub_rn(x[0], cell_origin(c.vstart[0], c.gs, cc[0]));
        xs.y = (T)__dsub_rn(x[1], cell_origin(c.vstart[1], c.gs, cc[1]));
        xs.z = (T)__dsub_rn(x[2], cell_origin(c.vstart[2], c.gs, cc[2]));
    }
    c.xs4[i] = xs;
}
template <typename T> int advect_pos(SphCtx *c) {
    if (c->n == 0) return 0;
    const int n = (int)c->n;
    Dev<T> d = make_dev<T>(c);                  // (XSPH reads the pre-move positions: it still replays the step's lists)
    c->gnl_valid = false;                     // positions move: what follows (mu(I) regularisation, H15) walks the cells
    if (!c->p.xsph) {
        SPH_PROF(c, K_ADVECT_POS);
        k_advect_pos<T><<<blocks_for(n, 256), 256, 0, c->stream>>>(d);
        SPH_LAUNCH_CHECK(c);
    } else {
        d.xnew = (double *)(c->arena + c->f[SPH_F_X].off[1 - c->f[SPH_F_X].cur]);
        SPH_PROF(c, K_ADVECT_POS);
        k_advect_pos_xsph<T><<<blocks_for(n, 128), 128, 0, c->stream>>>(d);
        SPH_LAUNCH_CHECK(c);
        flip(c, SPH_F_X);
    }
    return 0;
}

// --------------------------------------------------------------------------- dynamic rigid bodies
// Reaction on a dynamic rigid particle j (muI:45-46, dp:164-165): the reference lets every soil particle i subtract its
// momentum term  V_j rho~_j (sigma~_j / rho~_j^2 + sigma~_i / rho~_i^2) . gradW^c_i(x_i - x_j)  from d_vel_j while it
// forms its own sum (a scatter, serial in index order).  Here j GATHERS the same terms from its soil neighbours in
// ascending index order -- the order the serial scatter reaches it -- starting from the d_vel = g of the wall loop.
template <typename T> __global__ void __launch_bounds__(128) k_rigid_reaction(Dev<T> c) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c.N()) return;
    if (rigid_body_of(c, j) < 0 || not_owned(c, j)) return;
    const T cf = c.xs4[j].w * c.vt4[j].w;
    const T *qj = c.sor + 6 * (size_t)j;
    Vec4<T> acc = c.d_vel[j];
    for_neighbors(c, j, [&](int i, T dx, T dy, T dz, T r, T Vi) {
        if (!is_soil(c.type[i])) return;
        T L[9], g[3], gc[3];
        load_L(c, i, L);
        grad_corr(c, L, kernel_dW_over_r(c, r), -dx, -dy, -dz, g, gc);      // d = x_i - x_j as the soil particle sees it
        const T *qi = c.sor + 6 * (size_t)i;
        const T mxx = cf * (qj[0] + qi[0]), myy = cf * (qj[1] + qi[1]), mzz = cf * (qj[2] + qi[2]);
        const T mxy = cf * (qj[3] + qi[3]), myz = cf * (qj[4] + qi[4]), mzx = cf * (qj[5] + qi[5]);
        T t0 = 0, t1 = 0, t2 = 0;
        t0 += mxx * gc[0]; t0 += mxy * gc[1]; t0 += mzx * gc[2];
        t1 += mxy * gc[0]; t1 += myy * gc[1]; t1 += myz * gc[2];
        t2 += mzx * gc[0]; t2 += myz * gc[1]; t2 += mzz * gc[2];
        acc.x -= t0; acc.y -= t1; acc.z -= t2;
    });
    c.d_vel[j] = acc;
}
template <typename T> static int rigid_reaction(SphCtx *c) {
    if (c->rig_n == 0 || c->n == 0) return 0;
    SPH_PROF(c, K_OTHER);
    k_rigid_reaction<T><<<blocks_for(c->n, 128), 128, 0, c->stream>>>(make_dev<T>(c));
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// Per body one block: deterministic sums over the body's particles (strided partial sums in index order, then a tree).
// what = 0: mass-weighted centre cm (calc_cm, base:501-510; also written to rest_cm when `rest`);
// what = 1: A = sum m_V rho (x - cm)(x0 - rest_cm)^T (base:483-488), then R of its polar decomposition.
__device__ void rigid_polar_rotation(const double A[9], double R[9]);
template <typename T> __global__ void __launch_bounds__(256) k_rigid_reduce(Dev<T> c, int what, int rest) {
    __shared__ double sh[256][10];
    const int body = blockIdx.x, tid = threadIdx.x;
    double *buf = c.rig_buf + (size_t)body * RIG_STRIDE;
    double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int n = c.N();
    for (int i = tid; i < n; i += 256) {
        if (rigid_body_of(c, i) != body) continue;
        const double *x = c.x + 3 * (size_t)i;
        if (what == 0) {
            const double m = (double)c.v4[i].w;
            acc[0] += m * x[0]; acc[1] += m * x[1]; acc[2] += m * x[2]; acc[3] += m;
        } else {
            const double w = (double)c.xs4[i].w * c.rho[i];
            const double *x0 = c.rig_x0 + 3 * (size_t)c.id0[i];
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) acc[3 * a + b] += w * (x[a] - buf[3 + a]) * (x0[b] - buf[b]);
        }
    }
#pragma unroll
    for (int k = 0; k < 10; k++) sh[tid][k] = acc[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s)
#pragma unroll
            for (int k = 0; k < 10; k++) sh[tid][k] += sh[tid + s][k];
        __syncthreads();
    }
    if (tid != 0) return;
    if (what == 0) {
        for (int a = 0; a < 3; a++) { buf[3 + a] = sh[0][a] / sh[0][3]; if (rest) buf[a] = buf[3 + a]; }
        buf[6] = sh[0][3];
    } else {
        double A[9], R[9];
        for (int k = 0; k < 9; k++) { A[k] = sh[0][k]; buf[7 + k] = A[k]; }
        rigid_polar_rotation(A, R);
        for (int k = 0; k < 9; k++) buf[16 + k] = R[k];
    }
}
// goal position of every particle of a dynamic rigid body: x := cm + R (x0 - rest_cm)   (base:494-498)
template <typename T> __global__ void __launch_bounds__(256) k_rigid_apply(Dev<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    const int body = rigid_body_of(c, i);
    if (body < 0) return;
    const double *buf = c.rig_buf + (size_t)body * RIG_STRIDE, *R = buf + 16;
    const double *x0 = c.rig_x0 + 3 * (size_t)c.id0[i];
    double *x = c.x + 3 * (size_t)i;
    const double q[3] = {x0[0] - buf[0], x0[1] - buf[1], x0[2] - buf[2]};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const double goal = buf[3 + a] + (R[3 * a] * q[0] + R[3 * a + 1] * q[1] + R[3 * a + 2] * q[2]);
        x[a] += (goal - x[a]) * 1.0;
    }
}
// Rotation of the polar decomposition as Taichi's polar_decompose3d defines it (U, V proper rotations of the SVD,
// R = U V^T): Jacobi eigen-decomposition of A^T A for V, u_k = A v_k / sigma_k, missing columns by cross products.
__device__ void rigid_polar_rotation(const double A[9], double R[9]) {
    double B[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) { double t = 0; for (int k = 0; k < 3; k++) t += A[3 * k + a] * A[3 * k + b]; B[3 * a + b] = t; }
    for (int sweep = 0; sweep < 30; sweep++) {
        const double off = fabs(B[1]) + fabs(B[2]) + fabs(B[5]);
        if (off < 1e-300 || off <= 1e-18 * (fabs(B[0]) + fabs(B[4]) + fabs(B[8]))) break;
        for (int pq = 0; pq < 3; pq++) {
            const int pi = pq == 2 ? 1 : 0, qi = pq == 0 ? 1 : 2;
            const double apq = B[3 * pi + qi];
            if (apq == 0.0) continue;
            const double theta = (B[3 * qi + qi] - B[3 * pi + pi]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
            for (int k = 0; k < 3; k++) { const double bp = B[3 * k + pi], bq = B[3 * k + qi]; B[3 * k + pi] = cs * bp - sn * bq; B[3 * k + qi] = sn * bp + cs * bq; }
            for (int k = 0; k < 3; k++) { const double bp = B[3 * pi + k], bq = B[3 * qi + k]; B[3 * pi + k] = cs * bp - sn * bq; B[3 * qi + k] = sn * bp + cs * bq; }
            for (int k = 0; k < 3; k++) { const double vp = V[3 * k + pi], vq = V[3 * k + qi]; V[3 * k + pi] = cs * vp - sn * vq; V[3 * k + qi] = sn * vp + cs * vq; }
        }
    }
    int ord[3] = {0, 1, 2};
    for (int a = 0; a < 2; a++)
        for (int b = a + 1; b < 3; b++) if (B[4 * ord[b]] > B[4 * ord[a]]) { const int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double v[3][3], u[3][3], sig[3];
    for (int k = 0; k < 3; k++) {
        for (int a = 0; a < 3; a++) v[k][a] = V[3 * a + ord[k]];
        sig[k] = sqrt(B[4 * ord[k]] > 0.0 ? B[4 * ord[k]] : 0.0);
    }
    v[2][0] = v[0][1] * v[1][2] - v[0][2] * v[1][1]; v[2][1] = v[0][2] * v[1][0] - v[0][0] * v[1][2]; v[2][2] = v[0][0] * v[1][1] - v[0][1] * v[1][0];
    const double tol = 1e-12 * (sig[0] > 0 ? sig[0] : 1.0);
    int rank = 0;
    for (int k = 0; k < 2; k++) {
        if (sig[k] <= tol) break;
        for (int a = 0; a < 3; a++) { double t = 0; for (int b = 0; b < 3; b++) t += A[3 * a + b] * v[k][b]; u[k][a] = t / sig[k]; }
        rank++;
    }
    if (rank == 0) { for (int a = 0; a < 9; a++) R[a] = (a % 4 == 0) ? 1.0 : 0.0; return; }   // A == 0: identity (base:491-492)
    if (rank == 1) {
        const int m = fabs(u[0][0]) < fabs(u[0][1]) ? (fabs(u[0][0]) < fabs(u[0][2]) ? 0 : 2) : (fabs(u[0][1]) < fabs(u[0][2]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[m] = 1.0;
        const double w[3] = {u[0][1] * e[2] - u[0][2] * e[1], u[0][2] * e[0] - u[0][0] * e[2], u[0][0] * e[1] - u[0][1] * e[0]};
        const double nw = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        for (int a = 0; a < 3; a++) u[1][a] = w[a] / nw;
    }
    u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1]; u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2]; u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) R[3 * a + b] = u[0][a] * v[0][b] + u[1][a] * v[1][b] + u[2][a] * v[2][b];
}
template <typename T> int init_rigid_body(SphCtx *c) {
    if (c->rig_n == 0 || c->n == 0) return 0;
    SPH_PROF(c, K_OTHER);
    k_rigid_reduce<T><<<c->rig_n, 256, 0, c->stream>>>(make_dev<T>(c), 0, 1);
    SPH_LAUNCH_CHECK(c);
    return 0;
}
template <typename T> int solve_rigid_body(SphCtx *c) {
    if (c->rig_n == 0 || c->n == 0) return 0;
    Dev<T> d = make_dev<T>(c);
    c->gnl_valid = false; c->masks_valid = false;                   // positions move
    SPH_PROF(c, K_OTHER);
    k_rigid_reduce<T><<<c->rig_n, 256, 0, c->stream>>>(d, 0, 0);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_OTHER);
    k_rigid_reduce<T><<<c->rig_n, 256, 0, c->stream>>>(d, 1, 0);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_OTHER);
    k_rigid_apply<T><<<blocks_for(c->n, 256), 256, 0, c->stream>>>(d);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// --------------------------------------------------------------------- advect_something (wc:129-132 | muI | dp)
template <typename T> __device__ __forceinline__ void chk_density(const Dev<T> &c, int i) {     // base:214-221
    double r = c.rho[i];
    if (r < c.rho0) r = c.rho0;
    c.rho[i] = r;
    Vec4<T> xs = c.xs4[i];
    xs.w = (T)((double)c.v4[i].w / r);
    c.xs4[i] = xs;
}
template <typename T> __global__ void __launch_bounds__(256) k_post_wc(Dev<T> c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    if (is_fluid(c.type[i])) chk_density(c, i);
}
template <typename T> __global__ void __launch_bounds__(256) k_post_dp(Dev<T> c) {            // dp:276-296
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    if (!is_soil(c.type[i])) return;
    chk_density(c, i);
    T s[9];
    sym_load(c.stress, (size_t)i, s);
    c.flag[i] = flag_dp(c, s);
    adapt_stress(c, s);
    sym_store(c.stress, (size_t)i, s);
    c.strain[i] += (T)c.dt * c.d_strain[i];
    c.strain_p[i] += (T)c.dt * c.d_strain_p[i];
}
template <typename T> __global__ void __launch_bounds__(256) k_post_mui_a(Dev<T> c) {         // muI:147-150
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    if (!is_soil(c.type[i])) return;
    chk_density(c, i);
    c.strain[i] += (T)c.dt * c.d_strain[i];
}
// muI:151-156 Shepard regularisation on a snapshot (stress_tmp), post-advect positions on the pre-move grid (H15)
template <typename T> __device__ __forceinline__ void body_post_mui_b(const Dev<T> &c, int i) {
    const int ti = c.type[i];
    if (!is_soil(ti)) return;
    if (not_owned(c, i)) return;
    T acc[6] = {0, 0, 0, 0, 0, 0};
    for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) {
        if (c.type[j] == ti) {
            const T w = kernel_W(c, r);
            const T *sj = c.stress_t + 6 * (size_t)j;
#pragma unroll
            for (int q = 0; q < 6; q++) acc[q] += Vj * sj[q] * w;
        }
    });
    const T f = c.cspm_f[i];
#pragma unroll
    for (int q = 0; q < 6; q++) c.stress[6 * (size_t)i + q] = acc[q] * f;
}
template <typename T> __global__ void __launch_bounds__(128) k_post_mui_b(Dev<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.N()) body_post_mui_b(c, i);
}
template <typename T> int post_step(SphCtx *c) {
    if (c->n == 0) return 0;
    const int n = (int)c->n;
    Dev<T> d = make_dev<T>(c);
    cudaStream_t st = c->stream;
    if (c->p.solver == SPH_SOLVER_WC) {
        SPH_PROF(c, K_POST);
        k_post_wc<T><<<blocks_for(n, 256), 256, 0, st>>>(d);
        SPH_LAUNCH_CHECK(c);
    } else if (c->p.solver == SPH_SOLVER_DP) {
        SPH_PROF(c, K_POST);
        k_post_dp<T><<<blocks_for(n, 256), 256, 0, st>>>(d);
        SPH_LAUNCH_CHECK(c);
    } else {
        SPH_PROF(c, K_POST);
        k_post_mui_a<T><<<blocks_for(n, 256), 256, 0, st>>>(d);
        SPH_LAUNCH_CHECK(c);
        SPH_PROF(c, K_POST);
        k_refresh_xs<T><<<blocks_for(n, 256), 256, 0, st>>>(d);
        SPH_LAUNCH_CHECK(c);
        SPH_PROF(c, K_POST_SWEEP);
        k_post_mui_b<T><<<blocks_for(n, 128), 128, 0, st>>>(d);
        SPH_LAUNCH_CHECK(c);
    }
    return 0;
}

// enforce_boundary (base:525-601), boundary mode 1: flow particles are put back inside the domain box (no lid) and lose
// (1 + c_f) of their normal velocity (simulate_collisions, c_f = 0.3)
template <typename T> __global__ void __launch_bounds__(256) k_enforce_boundary(Dev<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    if (!(rigid_body_of(c, i) >= 0 || (c.boundary == 1 && is_flow(c.type[i])))) return;      // judge_enforce_bdy (base:531-538)
    double *x = c.x + 3 * (size_t)i;
    const double rr = c.radius_d - 1e-8;
    const double p0 = x[0], p1 = x[1], p2 = x[2];
    T n0 = 0, n1 = 0, n2 = 0;
    if (p0 > c.dend[0] - rr) { n0 += (T)1; x[0] = c.dend[0] - rr; }
    if (p0 <= c.dstart[0] + rr) { n0 += (T)-1; x[0] = c.dstart[0] + rr; }
    if (p1 <= c.dstart[1] + rr) { n1 += (T)-1; x[1] = c.dstart[1] + rr; }
    if (c.dim == 3) {
        if (p2 > c.dend[2] - rr) { n2 += (T)1; x[2] = c.dend[2] - rr; }
        if (p2 <= c.dstart[2] + rr) { n2 += (T)-1; x[2] = c.dstart[2] + rr; }
    }
    const T len = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
    if (len > c.eps) {
        const T u0 = n0 / len, u1 = n1 / len, u2 = n2 / len;
        Vec4<T> v = c.v4[i];
        const T vn = v.x * u0 + v.y * u1 + v.z * u2;
        v.x -= ((T)1 + (T)0.3) * vn * u0; v.y -= ((T)1 + (T)0.3) * vn * u1; v.z -= ((T)1 + (T)0.3) * vn * u2;
        c.v4[i] = v;
    }
}
template <typename T> int enforce_boundary(SphCtx *c) {
    if (c->n == 0 || (c->p.boundary != 1 && c->rig_n == 0)) return 0;
    SPH_PROF(c, K_OTHER);
    k_enforce_boundary<T><<<blocks_for(c->n, 256), 256, 0, c->stream>>>(make_dev<T>(c));
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// advect_SE / advect_LF (base:79-87, 106-114) + advect_pos without XSPH (base:228-238) + WCSPH advect_something
// (wc:129-132) of one particle in one kernel: the same operations in the same order as k_advect(kind 0), k_advect_pos
// and k_post_wc, without writing and re-reading density, velocity and volume in between (sph_step only).
template <typename T> __global__ void __launch_bounds__(256) k_wc_finish(Dev<T> c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    const int t = c.type[i];
    if (!is_real(t)) return;
    Vec4<T> v = c.v4[i];
    const Vec4<T> dv = c.d_vel[i];
    const double dt = c.dt;
    double r = c.rho[i] + dt * (double)c.d_rho[i];
    v.x += (T)dt * dv.x; v.y += (T)dt * dv.y; v.z += (T)dt * dv.z;
    c.v4[i] = v;
    double *x = c.x + 3 * (size_t)i;
    x[0] += dt * (double)v.x; x[1] += dt * (double)v.y; x[2] += dt * (double)v.z;
    if (is_fluid(t) && r < c.rho0) r = c.rho0;                      // chk_density (base:214-221)
    c.rho[i] = r;
    Vec4<T> xs = c.xs4[i];
    xs.w = (T)((double)v.w / r);
    c.xs4[i] = xs;
}
template <typename T> int finish_step(SphCtx *c) {
    if (c->n == 0) return 0;
    c->gnl_valid = false;
    SPH_PROF(c, K_ADVECT);
    k_wc_finish<T><<<blocks_for(c->n, 256), 256, 0, c->stream>>>(make_dev<T>(c));
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// -------------------------------------------------------------------------------------------- stand-alone sweeps
template <typename T> __global__ void __launch_bounds__(128) k_neighbor_count(Dev<T> c, int *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    int cnt = 0;
    for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) { cnt++; });
    out[i] = cnt;
}
template <typename T> __global__ void __launch_bounds__(128) k_density_sum(Dev<T> c, T *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    T s = 0;
    for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) { s += c.v4[j].w * kernel_W(c, r); });
    out[i] = s;
}
// count and density in one walk (BASELINE config C5).  rest_only: the cell-tile kernel already wrote the flow particles of
// unflagged cells; this kernel completes wall particles (their masks hold flow neighbours only) and flagged cells.
template <typename T> __global__ void __launch_bounds__(128) k_density_count(Dev<T> c, int *__restrict__ cnt_out, T *__restrict__ rho_out, int rest_only) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.N(); i += gridDim.x * blockDim.x) {
        if (not_owned(c, i)) continue;                        // ghost columns of a slab belong to the neighbour rank
        if (rest_only && is_flow(c.type[i]) && !c.cellflag[c.gid[i]]) continue;
        int cnt = 0;
        T s = 0;
        for_neighbors(c, i, [&](int j, T dx, T dy, T dz, T r, T Vj) { cnt++; s += c.v4[j].w * kernel_W(c, r); });
        cnt_out[i] = cnt;
        rho_out[i] = s;
    }
}
template <typename T> int density_sweep(SphCtx *c, int32_t *count_out, void *rho_out) {
    if (c->n == 0) return 0;
    int rest_only = 0;
    if (c->fast && sizeof(T) == 4) {
        if (!c->masks_valid) { int r = tile_mask(c, false); if (r) return r; }
        int r = tile_density_sweep(c, count_out, (float *)rho_out);
        if (r) return r;
        rest_only = 1;
    }
    Dev<T> d = make_dev<T>(c);
    SPH_PROF(c, K_DENSITY_SUM);
    const int blocks = rest_only ? 148 * 16 : blocks_for(c->n, 128);
    k_density_count<T><<<blocks, 128, 0, c->stream>>>(d, count_out, (T *)rho_out, rest_only);
    SPH_LAUNCH_CHECK(c);
    return 0;
}
template <typename T> int neighbor_count(SphCtx *c, int32_t *out) {
    if (c->n == 0) return 0;
    SPH_PROF(c, K_NEIGHBOR_COUNT);
    k_neighbor_count<T><<<blocks_for(c->n, 128), 128, 0, c->stream>>>(make_dev<T>(c), out);
    SPH_LAUNCH_CHECK(c);
    return 0;
}
template <typename T> int density_sum(SphCtx *c, void *out) {
    if (c->n == 0) return 0;
    SPH_PROF(c, K_DENSITY_SUM);
    k_density_sum<T><<<blocks_for(c->n, 128), 128, 0, c->stream>>>(make_dev<T>(c), (T *)out);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

#define INST(T)                                            \
    template int calc_kernel_corr<T>(SphCtx *, bool);            \
    template int one_step<T>(SphCtx *, bool);                 \
    template int one_step_phase<T>(SphCtx *, int);                    \
    template int advect_pos<T>(SphCtx *);                  \
    template int post_step<T>(SphCtx *);                   \
    template int enforce_boundary<T>(SphCtx *);            \
    template int init_rigid_body<T>(SphCtx *);             \
    template int solve_rigid_body<T>(SphCtx *);            \
    template int finish_step<T>(SphCtx *);                   \
    template int neighbor_count<T>(SphCtx *, int32_t *);   \
    template int density_sum<T>(SphCtx *, void *);           \
    template int density_sweep<T>(SphCtx *, int32_t *, void *);
INST(float)
INST(double)

}  // namespace sph
