// sph_dev.cuh -- device-side context, math helpers and SPH kernels shared by every .cu of libtisphi_b200.
// Hand-written for sm_100a.  Semantics follow SURVEY.md Appendix A (citations into /root/reference).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tisphi_b200.h"

namespace sph {

template <typename T> struct Vec4;
template <> struct __align__(16) Vec4<float> { float x, y, z, w; };
template <> struct __align__(16) Vec4<double> { double x, y, z, w; };

// material-type predicates (ps:320-374)
__host__ __device__ __forceinline__ bool is_fluid(int t) { return t == 1; }
__host__ __device__ __forceinline__ bool is_soil(int t) { return t == 2; }
__host__ __device__ __forceinline__ bool is_flow(int t) { return t == 1 || t == 2; }
__host__ __device__ __forceinline__ bool is_real(int t) { return t > 0; }
__host__ __device__ __forceinline__ bool is_bdy(int t) { return t == -1 || t == -2; }
__host__ __device__ __forceinline__ bool is_rigid(int t) { return t == 11; }
__host__ __device__ __forceinline__ bool is_wall(int t) { return is_bdy(t) || is_rigid(t); }

__host__ __device__ __forceinline__ bool is_rep(int t) { return t == -2; }
constexpr int RIG_STRIDE = 32, RIG_MAX = 64;      // doubles per dynamic rigid body in Dev::rig_buf; bodies per scene

// rounding-exact helpers: the neighbour predicate must not be FMA-contracted
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// squared distance of the neighbour predicate.  float64: unfused (dx*dx + dy*dy) + dz*dz, the expression the reference
// evaluates (ps:268 norm()).  float32 (MIXED): fma(dz,dz, fma(dy,dy, dx*dx)), restated verbatim with fmaf() in the oracle.
__device__ __forceinline__ double dist2(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ float dist2(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// Device view of one engine instance.  Passed by value to every kernel.
template <typename T> struct Dev {
    // sizes
    int n;               // particles in the arrays (owned + ghosts); multi-GPU slabs: an upper bound, see ndev
    const int *ndev;     // multi-GPU slabs: the particle count lives on the device (migration changes it without the host
                         // ever reading it back); null on one GPU.  Kernels bound their loops with N().
    __device__ __forceinline__ int N() const { return ndev ? *ndev : n; }
    int dim, kernel, kcorr, solver, xsph, wc_fresh;
    int gn[3];
    int C;
    int own0, own1;      // owned x-columns [own0, own1): sweeps skip particles of other columns (ghosts)
    // geometry (float64: cell ids are always computed in float64 like the reference, ps:216-226)
    double vstart[3], gs;
    double dt, m_V0d;
    // engine-real constants
    T h, hinv, support, r2thr, eps, knorm, gsT;   // gsT = (T)grid_size in MIXED, 0 in F64 (coordinates are global)
    T g[3], m_V0;
    T visc_coef, rho0T, h2_001;                   // 2(dim+2) nu ; rho0 ; 0.01 h^2
    double rho0, stiff, gamma_, vsound;
    int boundary;        // 0 none, 1 enforced collision, 2 dummy, 3 repulsive, 4 dummy + repulsive
    double dstart[3], dend[3], radius_d;
    T rep_k, rep_judge, rep_ginv;                 // 0.01 vsound^2 ; particle diameter ; 1 / (0.75 h)   (base:675-689)
    T coh, mu, E, alpha, kc, G, K, eps_f, sin_dila, damp_c;   // damp_c = -5e-5 * sqrt(E)/h  (base:713-715)
    // arrays (sorted order)
    double *x;           // n x 3
    double *rho, *rho_t; // density, density_tmp
    Vec4<T> *v4;         // v.xyz, mass
    Vec4<T> *vt4;        // v_tmp.xyz, (T)density_tmp
    Vec4<T> *xs4;        // sweep coords xyz, m_V
    T *press, *pnew;
    int *type, *id0, *gid, *flag;
    T *stress, *stress_t, *strain, *strain_p;     // x6
    T *sor;                                       // x6 scratch: stress_tmp / density_tmp^2, the per-neighbour quotients of the momentum sum
    T *cspm_f, *cspm_L;
    T *d_rho; Vec4<T> *d_vel; T *d_stress, *v_grad, *d_strain, *d_strain_p;
    T *d_rho_rk; Vec4<T> *d_vel_rk; T *d_stress_rk;
    int *cell_end, *cell_cnt;
    unsigned long long *bad;                      // counter of out-of-grid particles (H7)
    // cell-tile fast path (sweeps_tile.cu); null when the fast path is not allocated
    Vec4<T> *ps4;        // sweep coords xyz, +m_V for flow particles / -m_V otherwise (tile payload A)
    Vec4<T> *pk4;        // v_tmp.xyz, pressure / density_tmp^2                        (tile payload B of the fluid pass)
    Vec4<T> *pw4;        // EOS pressure, previous pressure, 0, 0                       (tile payload C of the wall pass)
    unsigned *mask;      // neighbour bit masks, word-major: mask[word * n + i], word = neighbour cell (x-major, z fastest);
                         // bit b = the b-th particle of that cell.  Wall particles keep their FLOW neighbours only.
    unsigned *nzw;            // per particle: bitmap of its non-zero mask words
    unsigned char *cellflag;  // 1: this centre cell cannot use the tile path (a cell of its stencil holds > 32 particles ...)
    int *nflag;               // number of flagged cells (device counter)
    unsigned char *cellinfo;  // per cell, written by the mask kernel: 1 has flow particles, 2 has wall particles,
                              // 4 has wall particles AND a stencil cell with flow particles (candidate for the wall pass)
    int *worklist[4];         // work lists: footprint segments 0 occupied (stand-alone Shepard pass), 1 flow (fluid pass),
                              // 3 occupied with flow particles in reach (mask pass); 2 wall CELLS in reach of flow (wall pass)
    int *wcount;              // their lengths (4 ints), then the dynamic cursors of the persistent kernels
    T *psx, *psy, *psz, *psf; // SoA copy of the sweep coordinates + flow sign (+1 flow / -1 other): mask-kernel tiles
    unsigned char *cellflow;  // per cell: 1 when it holds a flow particle (written by the reorder kernel)
    // neighbour round lists (fast >= 2): what the first fluid pass after the masks found, replayed by the later passes
    // of the step.  For the cell whose first particle is `is` and
    // which holds nc particles: nlist[is * LIST_ROUNDS + round * nc + k] (one contiguous block per cell) = two words {cc:5 | tile index:12 | tile index:12} = the four neighbour slots
    // of that round; lrounds[cell] = rounds stored for the cell's warp, -1 when it needed more than LIST_ROUNDS.
    uint2 *nlist;
    int *lrounds;
    int flagged_only;         // generic kernels: process only particles of flagged cells
    double *xnew;             // XSPH: the alternate position buffer the advected positions go to
    // dynamic rigid bodies: per CREATION index (id0) the body a particle belongs to (-1: none / static) and
    // its rest position x0; per body 32 doubles: rest_cm[3], cm[3], mass, A[9], R[9]  (null: the scene has none)
    const int *rig_obj;
    const double *rig_x0;
    double *rig_buf;
    int rig_n;
    // per-step neighbour lists of the generic sweeps (sweeps.cu::k_build_nlist): positions are frozen between the grid build
    // and advect_pos, so the candidate walk is done ONCE per step and every sweep in between replays its result.
    // Entry k of particle i: gnl[k * gnl_stride + i] = (stencil cell << 27) | j; gnl_count[i] < 0: did not fit (walk again).
    const unsigned *gnl;      // null: no valid list (every sweep walks the cells)
    const int *gnl_count;
    int gnl_stride, gnl_cap;
};

// ps:356-362: is_rigid_dynamic / the body index of a dynamic rigid particle (-1 otherwise); pt.is_dynamic is 1 for every
// other kind of particle (ps:150-174)
template <typename T> __device__ __forceinline__ int rigid_body_of(const Dev<T> &c, int i) {
    return (c.rig_obj && c.type[i] == 11) ? c.rig_obj[c.id0[i]] : -1;
}
template <typename T> __device__ __forceinline__ bool is_dynamic(const Dev<T> &c, int i) {
    return c.type[i] != 11 || rigid_body_of(c, i) >= 0;
}

// ------------------------------------------------------------------------------------------------ cells
template <typename T>
__device__ __forceinline__ void pos_to_cell(const Dev<T> &c, const double *x, int cc[3]) {
    // ps:216-218: trunc((x - vstart) / grid_size), float64 divide, C cast
    cc[0] = (int)__ddiv_rn(__dsub_rn(x[0], c.vstart[0]), c.gs);
    cc[1] = (int)__ddiv_rn(__dsub_rn(x[1], c.vstart[1]), c.gs);
    cc[2] = (int)__ddiv_rn(__dsub_rn(x[2], c.vstart[2]), c.gs);
}
template <typename T> __device__ __forceinline__ int flatten(const Dev<T> &c, int cx, int cy, int cz) {
    return cx * c.gn[1] * c.gn[2] + cy * c.gn[2] + cz;      // ps:221-222
}
template <typename T> __device__ __forceinline__ void unflatten(const Dev<T> &c, int g, int cc[3]) {
    int nyz = c.gn[1] * c.gn[2];
    cc[0] = g / nyz;
    int r = g - cc[0] * nyz;
    cc[1] = r / c.gn[2];
    cc[2] = r - cc[1] * c.gn[2];
}
// sweep coordinate of a position stored in cell (cx,cy,cz): global in F64, cell-local float in MIXED
__device__ __forceinline__ double cell_origin(double vstart, double gs, int c) {
    return __dadd_rn(vstart, __dmul_rn((double)c, gs));
}

// --------------------------------------------------------------------------------- smoothing kernels (base:278-358)
template <typename T> __device__ __forceinline__ T kernel_W(const Dev<T> &c, T r) {
    T q = r * c.hinv, res = 0;
    if (r > c.eps && q <= (T)2) {
        if (c.kernel == 0) {
            if (q <= (T)1) res = c.knorm * ((T)0.5 * q * q * q - q * q + (T)(2.0 / 3.0));
            else { T t = (T)2 - q; res = c.knorm / (T)6 * t * t * t; }
        } else {
            T q1 = (T)1 - (T)0.5 * q, q2 = q1 * q1;
            res = c.knorm * (q2 * q2) * ((T)1 + (T)2 * q);
        }
    }
    return res;
}
// returns the scalar s with gradW = s * d  (d = x_i - x_j)
template <typename T> __device__ __forceinline__ T kernel_dW_over_r(const Dev<T> &c, T r) {
    T q = r * c.hinv, s = 0;
    if (r > c.eps && q <= (T)2) {
        if (c.kernel == 0) {
            T f = (q <= (T)1) ? c.knorm * q * ((T)1.5 * q - (T)2) : c.knorm * ((T)-0.5 * ((T)2 - q) * ((T)2 - q));
            s = f * c.hinv / r;
        } else {
            T q1 = (T)1 - (T)0.5 * q;
            s = c.knorm * (q1 * q1 * q1) * ((T)-5 * q) * c.hinv / r;      // (q / r == hinv: kept as the reference writes it)
        }
    }
    return s;
}

// ------------------------------------------------------------------ generic neighbour iteration (ps:259-269)
// Centre cell from the CURRENT master position (ps:261); 3^dim cells x-major / z-fastest; j ascending; out-of-range
// cells per axis are empty; strict r < support evaluated as r2 < r2thr (same predicate, no sqrt).
// MIXED precision: coordinates are local to the cell a particle is STORED in, and a pair is always evaluated from
// the side of the lower cell id, in the frame of the higher cell:  lo in cell A, hi in cell B > A, s = (B - A) * gs,
//   d_canon = (x_lo - s) - x_hi ,   d(lo -> hi) = d_canon ,   d(hi -> lo) = -d_canon
// so the neighbour relation and every pair geometry are EXACTLY antisymmetric (the cell-tile mask kernel relies on
// it: it evaluates a cell pair once and transposes the bit matrix).   body(j, dx, dy, dz, r, V_j)
//
// Two forms with IDENTICAL results (same neighbours, same order, same arithmetic):
//   walk    the candidate loops over the 3^dim cells.  MODE 0: the task runs inside them; 1: they only record (stencil
//           cell, j) words; 2: both -- the first sweep after the grid build, the kernel correction, builds the step's lists
//           on its way (sweeps.cu::k_corr_nlist);
//   replay  (Dev::gnl != null and the particle's list fits) the task runs over the recorded words: no candidate tests,
//           and every lane of a warp has work until its own list ends instead of idling through the ~2/3 (2D) of the
//           candidates that fail the test.
constexpr unsigned NB_IDX_BITS = 27, NB_IDX_MASK = (1u << NB_IDX_BITS) - 1u;      // lists need n_max < 2^27

// returns the number of neighbours (MODE 1, 2) or 0
template <typename T, int MODE, typename F>
__device__ __forceinline__ int for_neighbors_walk(const Dev<T> &c, int i, unsigned *out, int ostride, int ocap, F &&body) {
    int cc[3], sc[3] = {0, 0, 0};
    const double xi[3] = {c.x[3 * (size_t)i], c.x[3 * (size_t)i + 1], c.x[3 * (size_t)i + 2]};
    pos_to_cell(c, xi, cc);
    const int gi = sizeof(T) == 4 ? c.gid[i] : 0;
    if (sizeof(T) == 4) unflatten(c, gi, sc);                 // the cell xs4[i] is local to
    const Vec4<T> pi = c.xs4[i];
    const int z0 = (c.dim == 2) ? 0 : -1, z1 = (c.dim == 2) ? 0 : 1;
    int cnt = 0;
    for (int ox = -1; ox <= 1; ox++) {
        int cx = cc[0] + ox;
        if (cx < 0 || cx >= c.gn[0]) continue;
        T sx = (T)(cx - sc[0]) * c.gsT;
        for (int oy = -1; oy <= 1; oy++) {
            int cy = cc[1] + oy;
            if (cy < 0 || cy >= c.gn[1]) continue;
            T sy = (T)(cy - sc[1]) * c.gsT;
            for (int oz = z0; oz <= z1; oz++) {
                int cz = cc[2] + oz;
                if (cz < 0 || cz >= c.gn[2]) continue;
                T sz = (T)(cz - sc[2]) * c.gsT;
                int g = flatten(c, cx, cy, cz);
                int jb = g > 0 ? c.cell_end[g - 1] : 0, je = c.cell_end[g];
                const bool rev = sizeof(T) == 4 && g < gi;    // the neighbour's cell precedes mine: evaluate from its side
                const T ex = pi.x - sx, ey = pi.y - sy, ez = pi.z - sz;
                const unsigned code = (unsigned)((cx - sc[0] + 1) * 9 + (cy - sc[1] + 1) * 3 + (cz - sc[2] + 1)) << NB_IDX_BITS;
                for (int j = jb; j < je; j++) {
                    if (j == i) continue;
                    Vec4<T> pj = c.xs4[j];
                    T dx, dy, dz;
                    if (!rev) { dx = ex - pj.x; dy = ey - pj.y; dz = ez - pj.z; }
                    else { dx = -((pj.x + sx) - pi.x); dy = -((pj.y + sy) - pi.y); dz = -((pj.z + sz) - pi.z); }
                    T r2 = dist2(dx, dy, dz);
                    if (r2 < c.r2thr) {
                        if (MODE != 0) {
                            if (cnt < ocap) out[(size_t)cnt * ostride] = code | (unsigned)j;
                            cnt++;
                        }
                        if (MODE != 1) body(j, dx, dy, dz, sqrt_rn(r2), pj.w);
                    }
                }
            }
        }
    }
    return cnt;
}
// (Tried and measured slower on dp_soil, 0.35 ms: an L1 prefetch instruction per gathered member of the next neighbour, 0.40 ms;
// holding the next neighbour's v~ in registers, 92 registers -> one block less per SM, 0.40 ms.)
template <typename T, typename F> __device__ __forceinline__ void for_neighbors(const Dev<T> &c, int i, F &&body) {
    const int cnt = c.gnl ? c.gnl_count[i] : -1;
    if (cnt < 0) {
        for_neighbors_walk<T, 0>(c, i, nullptr, 0, 0, body);
        return;
    }
    // replay: the list was built from the stored cell of i (positions have not changed since), shifts relative to it
    int sc[3] = {0, 0, 0};
    const int gi = sizeof(T) == 4 ? c.gid[i] : 0;
    if (sizeof(T) == 4) unflatten(c, gi, sc);
    const Vec4<T> pi = c.xs4[i];
    // Software pipeline: the list word of neighbour k+2 and the position of neighbour k+1 are in flight while the task
    // of neighbour k runs.  Without it every neighbour costs three dependent memory round trips (list word -> position
    // -> the task's own gathers) and the soil sweeps sat at 53 % issue with 57 % of the stalls on the long scoreboard
    // (profiles/r2_ncu_soil.csv); the order of the neighbours and the arithmetic are unchanged.
    const unsigned *e = c.gnl + i;
    const size_t st = (size_t)c.gnl_stride;
    unsigned w0 = cnt > 0 ? e[0] : 0u;
    unsigned w1 = cnt > 1 ? e[st] : 0u;
    Vec4<T> p0 = c.xs4[cnt > 0 ? (int)(w0 & NB_IDX_MASK) : i];
    for (int k = 0; k < cnt; k++) {
        const unsigned w2 = k + 2 < cnt ? e[(size_t)(k + 2) * st] : 0u;
        const int jn = k + 1 < cnt ? (int)(w1 & NB_IDX_MASK) : i;
        const Vec4<T> p1 = c.xs4[jn];
        const int j = (int)(w0 & NB_IDX_MASK), code = (int)(w0 >> NB_IDX_BITS);
        const int ox = code / 9 - 1, oy = (code / 3) % 3 - 1, oz = code % 3 - 1;     // neighbour cell - stored cell, per axis
        const T sx = (T)ox * c.gsT, sy = (T)oy * c.gsT, sz = (T)oz * c.gsT;
        const bool rev = sizeof(T) == 4 && (ox < 0 || (ox == 0 && (oy < 0 || (oy == 0 && oz < 0))));
        const Vec4<T> pj = p0;
        T dx, dy, dz;
        if (!rev) { dx = (pi.x - sx) - pj.x; dy = (pi.y - sy) - pj.y; dz = (pi.z - sz) - pj.z; }
        else { dx = -((pj.x + sx) - pi.x); dy = -((pj.y + sy) - pi.y); dz = -((pj.z + sz) - pi.z); }
        body(j, dx, dy, dz, sqrt_rn(dist2(dx, dy, dz)), pj.w);
        w0 = w1; w1 = w2; p0 = p1;
    }
}

// symmetric 3x3 stored as xx,yy,zz,xy,yz,zx  <-> full row-major
template <typename T> __device__ __forceinline__ void sym_load(const T *p, size_t i, T s[9]) {
    const T *q = p + 6 * i;
    T xx = q[0], yy = q[1], zz = q[2], xy = q[3], yz = q[4], zx = q[5];
    s[0] = xx; s[1] = xy; s[2] = zx; s[3] = xy; s[4] = yy; s[5] = yz; s[6] = zx; s[7] = yz; s[8] = zz;
}
template <typename T> __device__ __forceinline__ void sym_store(T *p, size_t i, const T s[9]) {
    T *q = p + 6 * i;
    q[0] = s[0]; q[1] = s[4]; q[2] = s[8]; q[3] = s[1]; q[4] = s[5]; q[5] = s[2];
}

}  // namespace sph
