// sweeps_tile.cu -- cell-major neighbour sweeps for the WCSPH hot loop (MIXED precision), hand-written for sm_100a.
//
// Work decomposition.  A thread block owns ZB consecutive cells along the fastest grid axis (z in 3D, y in 2D) of
// one cell column; warp w owns cell f0 + w and lane l owns the l-th particle of that cell.  Because the flattened
// cell id is fastest-axis-major (ps:221-222), the particles of the 3 x ... x (ZB + 2) cells a block needs form nR
// contiguous spans of the sorted arrays (nR = 9 in 3D, 3 in 2D): each span is brought into shared memory with ONE
// 1-D TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx), raw, with no transformation.
//
// Neighbour predicate.  The own coordinate is moved into the frame of each neighbour cell once per cell
// (e = x_i - shift), then d = e - x_j, r2 = fma(dz,dz, fma(dy,dy, dx*dx)), r2 < r2thr: the expression of
// sph_dev.cuh::for_neighbors in float32, so the tile kernels and the generic kernels select identical pairs in
// identical order (cells x-major / z-fastest, j ascending) -- summation order is preserved.
//
// Positions are frozen inside a step (they only change in advect_pos), so the predicate is evaluated ONCE per step
// (k_tile_mask) into one 32-bit word per (particle, neighbour cell); the wall pass and the fluid pass of both
// one_steps then only visit set bits.  Cells that cannot be represented (more than 32 particles in a stencil cell,
// or more than TILE_CAP particles in the block's tile) are flagged and processed by the generic kernels of sweeps.cu.
//
// Replaces, for this configuration: calc_CSPM_f (base:386-398), WCSPH one_step loops A and B (wc:82-126).
#include "sph_host.h"

namespace sph {

constexpr int ZB = 4;                 // cells (warps) per block along the fastest axis
constexpr int BT = ZB * 32;           // threads per block
constexpr int TILE_CAP = 1664;        // particles per block tile (3D rest lattice: 9 * 6 * 27 = 1458)
constexpr int NRMAX = 9;
constexpr int CBW = ZB + 3;           // cell boundaries per run

typedef Vec4<float> F4;
typedef Dev<float> DevF;

struct TileShared {
    F4 A[TILE_CAP];                   // ps4 spans
    F4 B[TILE_CAP];                   // second payload (vt4 or pk4) spans
    unsigned long long bar;           // mbarrier
    int cb[NRMAX * CBW];              // tile index of the first particle of each (run, cell)
    int gdelta[NRMAX];                // global index = tile index + gdelta[run]
    int total, overflow, any;
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// ------------------------------------------------------------------------------------------------ geometry
struct TileGeom {
    int nF, nR, nseg, ny, nslow0, nslow1;   // 3D: slow axes (x, y); 2D: slow axis x only (nslow1 = 1)
};
__host__ __device__ inline TileGeom make_geom(int dim, const int gn[3]) {
    TileGeom g;
    if (dim == 3) { g.nF = gn[2]; g.nR = 9; g.nslow0 = gn[0]; g.nslow1 = gn[1]; }
    else { g.nF = gn[1]; g.nR = 3; g.nslow0 = gn[0]; g.nslow1 = 1; }
    g.ny = g.nslow1;
    g.nseg = (g.nF + ZB - 1) / ZB;
    return g;
}
// shift of neighbour cell (run r, fast offset dzi in 0..2) relative to the centre cell, in units of the cell edge
__device__ __forceinline__ void cell_shift(const DevF &c, const TileGeom &g, int r, int dzi, float &sx, float &sy, float &sz) {
    if (g.nR == 9) { sx = (float)(r / 3 - 1) * c.gsT; sy = (float)(r % 3 - 1) * c.gsT; sz = (float)(dzi - 1) * c.gsT; }
    else { sx = (float)(r - 1) * c.gsT; sy = (float)(dzi - 1) * c.gsT; sz = 0.f; }
}

__device__ __forceinline__ int cell_start(const int *cell_end, int g) { return g > 0 ? cell_end[g - 1] : 0; }

// Computes spans and cell boundaries, issues the TMA copies (payload A always, payload B when srcB != null) and
// waits for them.  Returns false (uniformly) when the tile does not fit.  col/f0 identify the block's cells.
__device__ __forceinline__ bool tile_setup(const DevF &c, const TileGeom &g, TileShared &sh, int col, int f0,
                                           const F4 *srcA, const F4 *srcB) {
    const int tid = threadIdx.x;
    const int f_lo = max(f0 - 1, 0), f_hi = min(f0 + ZB, g.nF - 1);
    if (tid < 32) {
        int len = 0, S = 0, gb = 0;
        bool valid = false;
        if (tid < g.nR) {
            int s0 = col / g.ny, s1 = col - s0 * g.ny;          // slow coordinates of the block's column
            int n0 = s0, n1 = s1;
            if (g.nR == 9) { n0 += tid / 3 - 1; n1 += tid % 3 - 1; }
            else { n0 += tid - 1; }
            valid = n0 >= 0 && n0 < g.nslow0 && n1 >= 0 && n1 < g.nslow1;
            if (valid) {
                gb = (n0 * g.ny + n1) * g.nF;
                S = cell_start(c.cell_end, gb + f_lo);
                len = c.cell_end[gb + f_hi] - S;
            }
        }
        int inc = len;                                            // inclusive scan over the first nR lanes
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += t;
        }
        const int roff = inc - len;
        if (tid < g.nR) {
            sh.gdelta[tid] = S - roff;
            for (int k = 0; k < CBW; k++) {
                const int f = f0 - 1 + k;
                int v;
                if (!valid || f < f_lo) v = roff;
                else if (f > f_hi) v = roff + len;
                else v = roff + cell_start(c.cell_end, gb + f) - S;
                sh.cb[tid * CBW + k] = v;
            }
        }
        const int total = __shfl_sync(0xffffffffu, inc, g.nR - 1);
        if (tid == 0) {
            sh.total = total;
            sh.overflow = total > TILE_CAP;
            mbar_init(&sh.bar, 1);
        }
        __syncwarp();
        if (total <= TILE_CAP) {
            if (tid == 0) mbar_expect_tx(&sh.bar, (unsigned)(total * 16 * (srcB ? 2 : 1)));
            __syncwarp();
            if (tid < g.nR && len > 0) {
                tma_load_1d(&sh.A[roff], srcA + S, (unsigned)(len * 16), &sh.bar);
                if (srcB) tma_load_1d(&sh.B[roff], srcB + S, (unsigned)(len * 16), &sh.bar);
            }
        }
    }
    __syncthreads();
    if (sh.overflow) return false;
    mbar_wait(&sh.bar, 0);
    return true;
}

// float32 forms of the smoothing kernels (base:278-358) without divisions: W(r) and s with gradW = s * d
__device__ __forceinline__ float tile_W(const DevF &c, float r) {
    return kernel_W(c, r);
}
__device__ __forceinline__ float tile_dW(const DevF &c, float r) {
    return kernel_dW_over_r(c, r);
}

// ------------------------------------------------------------------------------------------------ pass 0: masks
// One launch per step, right after the grid build: neighbour masks, flow-neighbour counts and the Shepard factor
// CSPM_f (base:386-398) of every particle.
__global__ void __launch_bounds__(BT) k_tile_mask(DevF c, TileGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileShared &sh = *reinterpret_cast<TileShared *>(smem_raw);
    const int col = blockIdx.x / g.nseg, seg = blockIdx.x - col * g.nseg;
    const int f0 = seg * ZB, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = f0 + w;
    const int gcell = col * g.nF + f;
    int is = 0, nc = 0;
    if (f < g.nF) { is = cell_start(c.cell_end, gcell); nc = c.cell_end[gcell] - is; }
    if (!__syncthreads_or(nc > 0)) return;                         // empty segment
    const bool ok = tile_setup(c, g, sh, col, f0, c.ps4, nullptr);
    if (nc == 0) return;
    // can this cell be represented?  (uniform per warp)
    bool flagged = !ok || nc > 32;
    const int NW = g.nR * 3;
    if (!flagged) {
        for (int cc = lane; cc < NW; cc += 32) {
            const int r = cc / 3, dzi = cc - 3 * r;
            if (sh.cb[r * CBW + w + dzi + 1] - sh.cb[r * CBW + w + dzi] > 32) flagged = true;
        }
        flagged = __any_sync(0xffffffffu, flagged);
    }
    if (flagged) {
        if (lane == 0) { c.cellflag[gcell] = 1; atomicAdd(c.nflag, 1); }
        return;
    }
    if (lane >= nc) return;
    const int i = is + lane;
    const int rc = g.nR / 2;                                       // centre run
    const F4 pi = sh.A[sh.cb[rc * CBW + w + 1] + lane];
    float ssum = 0.f;
    int nflow = 0;
    for (int cc = 0; cc < NW; cc++) {
        const int r = cc / 3, dzi = cc - 3 * r;
        const int a = sh.cb[r * CBW + w + dzi], nb = sh.cb[r * CBW + w + dzi + 1] - a;
        float sx, sy, sz;
        cell_shift(c, g, r, dzi, sx, sy, sz);
        const float ex = pi.x - sx, ey = pi.y - sy, ez = pi.z - sz;
        unsigned m = 0, bit = 1;
        for (int t = 0; t < nb; t++, bit <<= 1) {
            const F4 pj = sh.A[a + t];                            // broadcast read
            const float dx = ex - pj.x, dy = ey - pj.y, dz = ez - pj.z;
            if (dist2(dx, dy, dz) < c.r2thr) m |= bit;
        }
        if (cc == NW / 2) m &= ~(1u << lane);                      // i != j
        c.mask[(size_t)cc * c.n + i] = m;
        while (m) {                                                // Shepard sum over flow neighbours, j ascending
            const int t = __ffs(m) - 1;
            m &= m - 1;
            const F4 pj = sh.A[a + t];
            if (pj.w > 0.f) {
                const float dx = ex - pj.x, dy = ey - pj.y, dz = ez - pj.z;
                ssum += pj.w * tile_W(c, sqrt_rn(dist2(dx, dy, dz)));
                nflow++;
            }
        }
    }
    c.cspm_f[i] = (ssum != 0.f) ? 1.f / ssum : 1.f;
    c.nflow[i] = (unsigned char)min(nflow, 255);
}

// ------------------------------------------------------------------------------------------------ prep (pointwise)
// wc:87-88 EOS in float64 into the NEW pressure buffer; signed volume of the tile payload; fluid half of pk4.
__global__ void __launch_bounds__(256) k_tile_prep(DevF c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    const int t = c.type[i];
    const F4 xs = c.xs4[i];
    F4 ps = xs;
    ps.w = is_flow(t) ? xs.w : -xs.w;
    c.ps4[i] = ps;
    if (is_fluid(t)) {
        const double rt = c.rho_t[i];
        double v = c.stiff * (pow(rt / c.rho0, c.gamma_) - 1.0);
        v = v > 0.0 ? v : 0.0;
        const float p = (float)v;
        c.pnew[i] = p;
        F4 pk = c.vt4[i];
        pk.w = p / (pk.w * pk.w);
        c.pk4[i] = pk;
    } else if (!is_wall(t)) {
        c.pnew[i] = c.press[i];
    }
}

// ------------------------------------------------------------------------------------------------ pass A: walls
// wc:90-103 for dummy-wall particles: v~ = 2v - f sum V v~ W, rho~ = rho0, p = max(f sum V (p_j + rho~_j g_y dy) W, 0).
__global__ void __launch_bounds__(BT) k_tile_wall(DevF c, TileGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileShared &sh = *reinterpret_cast<TileShared *>(smem_raw);
    const int col = blockIdx.x / g.nseg, seg = blockIdx.x - col * g.nseg;
    const int f0 = seg * ZB, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = f0 + w;
    const int gcell = col * g.nF + f;
    int is = 0, nc = 0;
    if (f < g.nF) { is = cell_start(c.cell_end, gcell); nc = c.cell_end[gcell] - is; }
    if (nc > 0 && c.cellflag[gcell]) nc = 0;                        // flagged cells belong to the generic kernels
    const int i = is + lane;
    bool wall = false, work = false;
    if (lane < nc) {
        wall = c.ps4[i].w < 0.f;
        work = wall && c.nflow[i] > 0;
        if (wall && !work) {                                        // no flow neighbour: the sums are empty
            const F4 v = c.v4[i];
            F4 vt; vt.x = 2.f * v.x; vt.y = 2.f * v.y; vt.z = 2.f * v.z; vt.w = c.rho0T;
            c.vt4[i] = vt;
            c.rho_t[i] = c.rho0;
            c.pnew[i] = 0.f;
            F4 pk = vt; pk.w = 0.f;
            c.pk4[i] = pk;
        }
    }
    if (!__syncthreads_or(work)) return;
    if (!tile_setup(c, g, sh, col, f0, c.ps4, c.vt4)) return;      // cannot happen for unflagged cells
    if (!work) return;
    const int rc = g.nR / 2, NW = g.nR * 3;
    const F4 pi = sh.A[sh.cb[rc * CBW + w + 1] + lane];
    float Sv0 = 0.f, Sv1 = 0.f, Sv2 = 0.f, Sp = 0.f;
    for (int cc = 0; cc < NW; cc++) {
        unsigned m = c.mask[(size_t)cc * c.n + i];
        if (!m) continue;
        const int r = cc / 3, dzi = cc - 3 * r;
        const int a = sh.cb[r * CBW + w + dzi];
        const int gd = sh.gdelta[r];
        float sx, sy, sz;
        cell_shift(c, g, r, dzi, sx, sy, sz);
        const float ex = pi.x - sx, ey = pi.y - sy, ez = pi.z - sz;
        while (m) {
            const int t = __ffs(m) - 1;
            m &= m - 1;
            const F4 pj = sh.A[a + t];
            if (pj.w > 0.f) {
                const float dx = ex - pj.x, dy = ey - pj.y, dz = ez - pj.z;
                const float wgt = tile_W(c, sqrt_rn(dist2(dx, dy, dz)));
                const F4 vj = sh.B[a + t];
                const int j = a + t + gd;
                Sv0 += pj.w * vj.x * wgt; Sv1 += pj.w * vj.y * wgt; Sv2 += pj.w * vj.z * wgt;
                const float pjv = (c.wc_fresh || j < i) ? c.pnew[j] : c.press[j];
                Sp += pj.w * (pjv + vj.w * c.g[1] * dy) * wgt;
            }
        }
    }
    const float fi = c.cspm_f[i];
    const F4 v = c.v4[i];
    F4 vt;
    vt.x = 2.f * v.x - Sv0 * fi; vt.y = 2.f * v.y - Sv1 * fi; vt.z = 2.f * v.z - Sv2 * fi; vt.w = c.rho0T;
    c.vt4[i] = vt;
    c.rho_t[i] = c.rho0;
    const float p = Sp * fi;
    const float pc = p > 0.f ? p : 0.f;
    c.pnew[i] = pc;
    F4 pk = vt; pk.w = pc / (c.rho0T * c.rho0T);
    c.pk4[i] = pk;
}

// ------------------------------------------------------------------------------------------------ pass B: fluid
// wc:108-126 for fluid particles: continuity + viscosity + pressure in one visit of the set bits.
__global__ void __launch_bounds__(BT) k_tile_fluid(DevF c, TileGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileShared &sh = *reinterpret_cast<TileShared *>(smem_raw);
    unsigned *smask = reinterpret_cast<unsigned *>(smem_raw + sizeof(TileShared));   // [NW][BT]
    const int col = blockIdx.x / g.nseg, seg = blockIdx.x - col * g.nseg;
    const int f0 = seg * ZB, w = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int f = f0 + w;
    const int gcell = col * g.nF + f;
    int is = 0, nc = 0;
    if (f < g.nF) { is = cell_start(c.cell_end, gcell); nc = c.cell_end[gcell] - is; }
    if (nc > 0 && c.cellflag[gcell]) nc = 0;
    const int i = is + lane;
    const bool work = lane < nc && c.type[i] == 1;
    if (!__syncthreads_or(work)) return;
    const int NW = g.nR * 3;
    if (work) {
        for (int cc = 0; cc < NW; cc++) smask[cc * BT + tid] = c.mask[(size_t)cc * c.n + i];
    }
    if (!tile_setup(c, g, sh, col, f0, c.ps4, c.pk4)) return;
    if (!work) return;
    const int rc = g.nR / 2;
    const int ci = sh.cb[rc * CBW + w + 1] + lane;
    const F4 pi = sh.A[ci], qi = sh.B[ci];                         // qi = v~_i, p_i / rho~_i^2
    const float rhoi = c.vt4[i].w;
    const float wallfac = c.rho0T / rhoi;                           // wc:43-44 factor for wall neighbours
    float dd = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
    int cc = -1, a = 0;
    unsigned m = 0;
    float ex = 0.f, ey = 0.f, ez = 0.f;
    while (true) {
        while (m == 0) {                                            // advance to the next non-empty neighbour cell
            if (++cc >= NW) break;
            m = smask[cc * BT + tid];
            if (m) {
                const int r = cc / 3, dzi = cc - 3 * r;
                a = sh.cb[r * CBW + w + dzi];
                float sx, sy, sz;
                cell_shift(c, g, r, dzi, sx, sy, sz);
                ex = pi.x - sx; ey = pi.y - sy; ez = pi.z - sz;
            }
        }
        if (cc >= NW) break;
        const int t = __ffs(m) - 1;
        m &= m - 1;
        const F4 pj = sh.A[a + t], qj = sh.B[a + t];
        const float dx = ex - pj.x, dy = ey - pj.y, dz = ez - pj.z;
        const float r2 = dist2(dx, dy, dz);
        const float s = tile_dW(c, sqrt_rn(r2));
        const float Vj = fabsf(pj.w);
        const float ux = qi.x - qj.x, uy = qi.y - qj.y, uz = qi.z - qj.z;
        const float gx = s * dx, gy = s * dy, gz = s * dz;
        dd += Vj * ux * gx + Vj * uy * gy + Vj * uz * gz;
        const float vx = ux * dx + uy * dy + uz * dz;
        const float mn = vx < 0.f ? vx : 0.f;
        float visc = c.visc_coef * Vj;
        if (pj.w < 0.f) visc = visc * c.rho0T / rhoi;
        visc = visc * mn / (r2 + c.h2_001);
        (void)wallfac;
        const float pres = -c.rho0T * Vj * (qi.w + qj.w);
        a0 += visc * gx + pres * gx; a1 += visc * gy + pres * gy; a2 += visc * gz + pres * gz;
    }
    c.d_rho[i] = dd * rhoi;
    F4 dv; dv.x = a0 + c.g[0]; dv.y = a1 + c.g[1]; dv.z = a2 + c.g[2]; dv.w = 0.f;
    c.d_vel[i] = dv;
}

// ------------------------------------------------------------------------------------------------ host side
static size_t tile_smem(bool with_mask, int NW) { return sizeof(TileShared) + (with_mask ? (size_t)NW * BT * 4 : 0); }

int tile_mask(SphCtx *c) {
    DevF d = make_dev<float>(c);
    const TileGeom g = make_geom(c->p.dim, d.gn);
    const int ncol = g.nslow0 * g.nslow1;
    static bool attr_done = false;
    if (!attr_done) {
        SPH_CHECK(c, cudaFuncSetAttribute(k_tile_mask, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem(false, 27)));
        SPH_CHECK(c, cudaFuncSetAttribute(k_tile_wall, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem(false, 27)));
        SPH_CHECK(c, cudaFuncSetAttribute(k_tile_fluid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem(true, 27)));
        attr_done = true;
    }
    SPH_CHECK(c, cudaMemsetAsync(d.cellflag, 0, (size_t)c->C, c->stream));
    SPH_CHECK(c, cudaMemsetAsync(d.nflag, 0, 4, c->stream));
    SPH_PROF(c, K_TILE_MASK);
    k_tile_mask<<<ncol * g.nseg, BT, tile_smem(false, g.nR * 3), c->stream>>>(d, g);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// WCSPH one_step (wc:82-126) on the tile path; flagged cells are completed by the generic kernels (flagged_only).

int tile_wc_prep_and_wall(SphCtx *c) {
    DevF d = make_dev<float>(c);
    const TileGeom g = make_geom(c->p.dim, d.gn);
    const int ncol = g.nslow0 * g.nslow1, n = (int)c->n;
    SPH_PROF(c, K_WC_EOS);
    k_tile_prep<<<blocks_for(n, 256), 256, 0, c->stream>>>(d);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_TILE_WALL);
    k_tile_wall<<<ncol * g.nseg, BT, tile_smem(false, g.nR * 3), c->stream>>>(d, g);
    SPH_LAUNCH_CHECK(c);
    return 0;
}
int tile_wc_fluid(SphCtx *c) {
    DevF d = make_dev<float>(c);
    const TileGeom g = make_geom(c->p.dim, d.gn);
    const int ncol = g.nslow0 * g.nslow1;
    SPH_PROF(c, K_TILE_FLUID);
    k_tile_fluid<<<ncol * g.nseg, BT, tile_smem(true, g.nR * 3), c->stream>>>(d, g);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

}  // namespace sph
