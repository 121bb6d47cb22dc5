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
constexpr int TILE_CAP = 1536;        // particles per block tile (3D rest lattice: 9 * 6 * 27 = 1458)
constexpr int NRMAX = 9;
constexpr int CBW = ZB + 3;           // cell boundaries per run

typedef Vec4<float> F4;
typedef Dev<float> DevF;

// NPAY payload arrays of TILE_CAP float4 (+8 entries of slack: the chunked test loop may read past a cell's end)
template <int NPAY> struct TileShared {
    F4 P[NPAY][TILE_CAP + 8];
    unsigned long long bar;           // mbarrier
    int cb[NRMAX * CBW];              // tile index of the first particle of each (run, cell)
    int gdelta[NRMAX];                // global index = tile index + gdelta[run]
    int total, overflow;
    F4 ctab[ZB * 27];                 // per (warp, neighbour cell): shift xyz, tile index of the cell's first particle (int bits)
    int cgd[ZB * 27];                 // per (warp, neighbour cell): global index - tile index
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
// shift of neighbour cell (run r, fast offset dzi in 0..2) relative to the centre cell
__device__ __forceinline__ void cell_shift(const DevF &c, const TileGeom &g, int r, int dzi, float &sx, float &sy, float &sz) {
    if (g.nR == 9) {
        const int rx = r / 3;
        sx = (float)(rx - 1) * c.gsT; sy = (float)(r - 3 * rx - 1) * c.gsT; sz = (float)(dzi - 1) * c.gsT;
    } else { sx = (float)(r - 1) * c.gsT; sy = (float)(dzi - 1) * c.gsT; sz = 0.f; }
}

// per-warp table of the stencil cells, so that advancing to the next cell costs one shared load
template <int NPAY>
__device__ __forceinline__ void build_ctab(const DevF &c, const TileGeom &g, TileShared<NPAY> &sh, int w, int lane) {
    const int NW = g.nR * 3;
    if (lane < NW) {
        const int r = lane / 3, dzi = lane - 3 * r;
        F4 t;
        if (g.nR == 9) {
            const int rx = r / 3;
            t.x = (float)(rx - 1) * c.gsT; t.y = (float)(r - 3 * rx - 1) * c.gsT; t.z = (float)(dzi - 1) * c.gsT;
        } else { t.x = (float)(r - 1) * c.gsT; t.y = (float)(dzi - 1) * c.gsT; t.z = 0.f; }
        t.w = __int_as_float(sh.cb[r * CBW + w + dzi]);
        sh.ctab[w * 27 + lane] = t;
        sh.cgd[w * 27 + lane] = sh.gdelta[r];
    }
    __syncwarp();
}

__device__ __forceinline__ int cell_start(const int *cell_end, int g) { return g > 0 ? cell_end[g - 1] : 0; }

// Computes spans and cell boundaries, issues the TMA copies of NPAY payload arrays and waits for them.
// Returns false (uniformly) when the tile does not fit.
template <int NPAY>
__device__ __forceinline__ bool tile_setup(const DevF &c, const TileGeom &g, TileShared<NPAY> &sh, int col, int f0,
                                           const F4 *src0, const F4 *src1, const F4 *src2 = nullptr) {
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
            if (tid == 0) mbar_expect_tx(&sh.bar, (unsigned)(total * 16 * NPAY));
            __syncwarp();
            if (tid < g.nR && len > 0) {
                tma_load_1d(&sh.P[0][roff], src0 + S, (unsigned)(len * 16), &sh.bar);
                if (NPAY > 1) tma_load_1d(&sh.P[1][roff], src1 + S, (unsigned)(len * 16), &sh.bar);
                if (NPAY > 2) tma_load_1d(&sh.P[NPAY - 1][roff], src2 + S, (unsigned)(len * 16), &sh.bar);
            }
        }
    }
    __syncthreads();
    if (sh.overflow) return false;
    mbar_wait(&sh.bar, 0);
    return true;
}

// ------------------------------------------------------------------------------------------------ float32 kernels
// Division-free float32 forms of base:278-358 on the squared distance (MUFU.RSQ instead of IEEE sqrt/div).
// Guards: r > 1e-8 and q <= 2 as in the reference.  KERNEL: 0 cubic spline, 1 Wendland C2.
struct KernConst { float hinv, eps2, knorm, c_grad; };   // c_grad = -5 knorm / h^2 (Wendland) | knorm / h^2 (cubic)
__device__ __forceinline__ KernConst kern_const(const DevF &c) {
    KernConst k;
    k.hinv = c.hinv; k.eps2 = c.eps * c.eps; k.knorm = c.knorm;
    k.c_grad = (c.kernel == 0 ? 1.f : -5.f) * c.knorm * c.hinv * c.hinv;
    return k;
}
template <int KERNEL> __device__ __forceinline__ float fastW(const KernConst &k, float r2) {
    const float rinv = rsqrtf(r2), q = r2 * rinv * k.hinv;
    float w;
    if (KERNEL == 1) { const float q1 = fmaf(-0.5f, q, 1.f), q2 = q1 * q1; w = k.knorm * (q2 * q2) * fmaf(2.f, q, 1.f); }
    else {
        const float t = 2.f - q;
        w = q <= 1.f ? k.knorm * (q * q * fmaf(0.5f, q, -1.f) + (float)(2.0 / 3.0)) : k.knorm * (1.f / 6.f) * t * t * t;
    }
    return (r2 > k.eps2 && q <= 2.f) ? w : 0.f;
}
// s with gradW = s * d
template <int KERNEL> __device__ __forceinline__ float fastdW(const KernConst &k, float r2) {
    const float rinv = rsqrtf(r2), q = r2 * rinv * k.hinv;
    float s;
    if (KERNEL == 1) { const float q1 = fmaf(-0.5f, q, 1.f); s = k.c_grad * (q1 * q1 * q1); }
    else {
        const float t = 2.f - q;
        s = q <= 1.f ? k.c_grad * fmaf(1.5f, q, -2.f) : -0.5f * k.c_grad * t * t * (rinv / k.hinv);
    }
    return (r2 > k.eps2 && q <= 2.f) ? s : 0.f;
}

// ------------------------------------------------------------------------------------------------ bit iteration
// Per-lane cursor over the set bits of the neighbour masks in stencil order.  A round takes up to four neighbours:
// two from the current cell, then (after an optional jump to the next non-empty cell) two more.  Slots that find no
// bit are marked invalid and contribute exactly zero, so the summation order (cells x-major, j ascending) is kept.
struct Cursor {
    unsigned m, nz;        // remaining bits of the current cell; remaining non-empty cells
    int a;                 // tile index of the current cell's first particle
    float ex, ey, ez;      // own coordinates in the current cell's frame
};
__device__ __forceinline__ void cursor_jump(Cursor &k, const unsigned *smask, const F4 *ct, const F4 &pi) {
    if (k.m == 0 && k.nz != 0) {
        const int cc = __ffs(k.nz) - 1;
        k.nz &= k.nz - 1;
        k.m = smask[cc * BT + threadIdx.x];
        const F4 t = ct[cc];
        k.a = __float_as_int(t.w);
        k.ex = pi.x - t.x; k.ey = pi.y - t.y; k.ez = pi.z - t.z;
    }
}
// takes up to two bits of the current cell: tile indices i0, i1 and validity of the second (the first is valid iff m != 0)
__device__ __forceinline__ void cursor_take2(Cursor &k, int &i0, int &i1, bool &v0, bool &v1) {
    v0 = k.m != 0;
    const int t0 = v0 ? __ffs(k.m) - 1 : 0;
    k.m &= k.m - 1;
    v1 = k.m != 0;
    const int t1 = v1 ? __ffs(k.m) - 1 : t0;
    k.m &= k.m - 1;
    i0 = k.a + t0; i1 = k.a + t1;
}

// ------------------------------------------------------------------------------------------------ pass 0: masks
// One launch per step, right after the grid build: neighbour masks, flow-neighbour counts and the Shepard factor
// CSPM_f (base:386-398) of every particle.
template <int KERNEL>
__global__ void __launch_bounds__(BT) k_tile_mask(DevF c, TileGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileShared<1> &sh = *reinterpret_cast<TileShared<1> *>(smem_raw);
    unsigned *smask = reinterpret_cast<unsigned *>(smem_raw + sizeof(TileShared<1>));   // [NW][BT]
    const int col = blockIdx.x / g.nseg, seg = blockIdx.x - col * g.nseg;
    if (col / g.ny < c.own0 || col / g.ny >= c.own1) return;        // ghost column of a slab (block-uniform)
    const int f0 = seg * ZB, w = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int f = f0 + w;
    const int gcell = col * g.nF + f;
    int is = 0, nc = 0;
    if (f < g.nF) { is = cell_start(c.cell_end, gcell); nc = c.cell_end[gcell] - is; }
    if (!__syncthreads_or(nc > 0)) return;                         // empty segment
    const bool ok = tile_setup<1>(c, g, sh, col, f0, c.ps4, nullptr);
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
    const bool mine = lane < nc;
    const int i = is + lane;
    const F4 *A = sh.P[0];
    const int rc = g.nR / 2;                                       // centre run
    const F4 pi = A[sh.cb[rc * CBW + w + 1] + (mine ? lane : 0)];
    const float thr = c.r2thr;
    unsigned nz = 0;                                               // which stencil cells hold at least one neighbour
    // ---- predicate: every lane tests every candidate of the 3^dim stencil (broadcast reads, 8-way ILP)
    for (int cc = 0; cc < NW; cc++) {
        const int r = cc / 3, dzi = cc - 3 * r;
        const int a = sh.cb[r * CBW + w + dzi], nb = sh.cb[r * CBW + w + dzi + 1] - a;
        float sx, sy, sz;
        cell_shift(c, g, r, dzi, sx, sy, sz);
        const float ex = pi.x - sx, ey = pi.y - sy, ez = pi.z - sz;
        unsigned m = 0;
        for (int t0 = 0; t0 < nb; t0 += 8) {
            unsigned cm = 0;
            F4 pj[8];
#pragma unroll
            for (int u = 0; u < 8; u++) pj[u] = A[a + t0 + u];     // may run past the cell: masked below
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float dx = ex - pj[u].x, dy = ey - pj[u].y, dz = ez - pj[u].z;
                if (dist2(dx, dy, dz) < thr) cm |= 1u << u;
            }
            m |= cm << t0;
        }
        m &= nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
        if (cc == NW / 2) m &= ~(1u << lane);                      // i != j
        if (!mine) m = 0;
        smask[cc * BT + tid] = m;
        if (m) nz |= 1u << cc;
        if (mine) c.mask[(size_t)cc * c.n + i] = m;
    }
    // ---- Shepard sum over flow neighbours in stencil order (cells x-major, j ascending): warp-uniform rounds
    const KernConst kc = kern_const(c);
    build_ctab<1>(c, g, sh, w, lane);
    const F4 *ct = sh.ctab + w * 27;
    float ssum = 0.f;
    int nflow = 0;
    Cursor k;
    k.m = 0; k.nz = mine ? nz : 0u; k.a = 0; k.ex = k.ey = k.ez = 0.f;
    while (true) {
        cursor_jump(k, smask, ct, pi);
        if (!__any_sync(0xffffffffu, k.m != 0)) break;              // warp-uniform round
        int i0, i1, i2, i3;
        bool v0, v1, v2, v3;
        cursor_take2(k, i0, i1, v0, v1);
        const float e0x = k.ex, e0y = k.ey, e0z = k.ez;
        cursor_jump(k, smask, ct, pi);
        cursor_take2(k, i2, i3, v2, v3);
        const F4 p0 = A[i0], p1 = A[i1], p2 = A[i2], p3 = A[i3];
        const float w0 = fastW<KERNEL>(kc, dist2(e0x - p0.x, e0y - p0.y, e0z - p0.z));
        const float w1 = fastW<KERNEL>(kc, dist2(e0x - p1.x, e0y - p1.y, e0z - p1.z));
        const float w2 = fastW<KERNEL>(kc, dist2(k.ex - p2.x, k.ey - p2.y, k.ez - p2.z));
        const float w3 = fastW<KERNEL>(kc, dist2(k.ex - p3.x, k.ey - p3.y, k.ez - p3.z));
        const bool f0 = v0 && p0.w > 0.f, f1 = v1 && p1.w > 0.f, f2 = v2 && p2.w > 0.f, f3 = v3 && p3.w > 0.f;
        ssum += f0 ? p0.w * w0 : 0.f;
        ssum += f1 ? p1.w * w1 : 0.f;
        ssum += f2 ? p2.w * w2 : 0.f;
        ssum += f3 ? p3.w * w3 : 0.f;
        nflow += (int)f0 + (int)f1 + (int)f2 + (int)f3;
    }
    if (mine) {
        c.cspm_f[i] = (ssum != 0.f) ? 1.f / ssum : 1.f;
        c.nflow[i] = (unsigned char)min(nflow, 255);
    }
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
        F4 pw; pw.x = p; pw.y = c.press[i]; pw.z = 0.f; pw.w = 0.f;   // EOS pressure, previous pressure (wc:86-103 race)
        c.pw4[i] = pw;
    } else if (!is_wall(t)) {
        const float po = c.press[i];
        c.pnew[i] = po;
        F4 pw; pw.x = po; pw.y = po; pw.z = 0.f; pw.w = 0.f;
        c.pw4[i] = pw;
    }
}

// ------------------------------------------------------------------------------------------------ pass A: walls
// loads the NW mask words of this thread's particle into shared memory (coalesced, all in flight at once) and
// returns the bitmap of non-empty stencil cells
__device__ __forceinline__ unsigned load_masks(const DevF &c, unsigned *smask, int NW, int i, bool work) {
    unsigned nz = 0;
    if (work) {
        const unsigned *mp = c.mask + i;
#pragma unroll 9
        for (int cc = 0; cc < NW; cc++) {
            const unsigned m = mp[(size_t)cc * c.n];
            smask[cc * BT + threadIdx.x] = m;
            if (m) nz |= 1u << cc;
        }
    }
    return nz;
}

// wc:90-103 for dummy-wall particles: v~ = 2v - f sum V v~ W, rho~ = rho0, p = max(f sum V (p_j + rho~_j g_y dy) W, 0).
// Tile payloads: ps4 (coords, signed volume), vt4 (v~, rho~), pw4 (EOS pressure, previous pressure).
template <int KERNEL>
__device__ __forceinline__ void wall_pair(const KernConst &kc, float ex, float ey, float ez, const F4 pj, const F4 vj, float pjv,
                                          bool valid, float gy, float &vw, float &pterm) {
    const float dx = ex - pj.x, dy = ey - pj.y, dz = ez - pj.z;
    const float wgt = fastW<KERNEL>(kc, dist2(dx, dy, dz));
    vw = (valid && pj.w > 0.f) ? pj.w * wgt : 0.f;               // flow neighbours only (base:654-663)
    pterm = fmaf(vj.w * gy, dy, pjv);
}

template <int KERNEL>
__global__ void __launch_bounds__(BT) k_tile_wall(DevF c, TileGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileShared<3> &sh = *reinterpret_cast<TileShared<3> *>(smem_raw);
    unsigned *smask = reinterpret_cast<unsigned *>(smem_raw + sizeof(TileShared<3>));   // [NW][BT]
    const int col = blockIdx.x / g.nseg, seg = blockIdx.x - col * g.nseg;
    if (col / g.ny < c.own0 || col / g.ny >= c.own1) return;        // ghost column of a slab (block-uniform)
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
    const int NW = g.nR * 3;
    const unsigned nz = load_masks(c, smask, NW, i, work);
    if (!tile_setup<3>(c, g, sh, col, f0, c.ps4, c.vt4, c.pw4)) return;   // cannot happen for unflagged cells
    build_ctab<3>(c, g, sh, w, lane);
    const F4 *A = sh.P[0], *B = sh.P[1], *Pw = sh.P[2];
    const F4 *ct = sh.ctab + w * 27;
    const int rc = g.nR / 2;
    const int ci = sh.cb[rc * CBW + w + 1];                         // tile index of this cell's first particle
    const F4 pi = A[ci + (work ? lane : 0)];
    const KernConst kc = kern_const(c);
    const float gy = c.g[1];
    const bool fresh = c.wc_fresh != 0;
    const int self = ci + lane;                                     // "j < i" in sorted order == tile index below mine (same run)
    float Sv0 = 0.f, Sv1 = 0.f, Sv2 = 0.f, Sp = 0.f;
    Cursor k;
    k.m = 0; k.nz = nz; k.a = 0; k.ex = k.ey = k.ez = 0.f;
    const int c_lo = sh.cb[rc * CBW + w + 1], c_hi = sh.cb[rc * CBW + w + 2];   // my own cell inside the tile
    const int centre = NW / 2;
    (void)centre;
    while (true) {
        cursor_jump(k, smask, ct, pi);
        if (!__any_sync(0xffffffffu, k.m != 0)) break;
        int i0, i1, i2, i3;
        bool v0, v1, v2, v3;
        // sorted order: every particle of a stencil cell with a smaller cell id precedes i, every one with a larger id
        // follows it; inside my own cell the tile index decides.  before0/before1 = "this cell precedes mine".
        cursor_take2(k, i0, i1, v0, v1);
        const float e0x = k.ex, e0y = k.ey, e0z = k.ez;
        const float s0x = pi.x - k.ex, s0y = pi.y - k.ey, s0z = pi.z - k.ez;       // the cell's shift
        cursor_jump(k, smask, ct, pi);
        cursor_take2(k, i2, i3, v2, v3);
        const float s1x = pi.x - k.ex, s1y = pi.y - k.ey, s1z = pi.z - k.ez;
        const F4 p0 = A[i0], p1 = A[i1], p2 = A[i2], p3 = A[i3];
        const F4 u0 = B[i0], u1 = B[i1], u2 = B[i2], u3 = B[i3];
        const F4 w0 = Pw[i0], w1 = Pw[i1], w2 = Pw[i2], w3 = Pw[i3];
        // stencil order is x-major / z-fastest == ascending cell id: shift (sx, sy, sz) lexicographically negative <=> cell precedes
        const bool b0 = s0x < 0.f || (s0x == 0.f && (s0y < 0.f || (s0y == 0.f && s0z < 0.f)));
        const bool b1 = s1x < 0.f || (s1x == 0.f && (s1y < 0.f || (s1y == 0.f && s1z < 0.f)));
        const bool same0 = s0x == 0.f && s0y == 0.f && s0z == 0.f, same1 = s1x == 0.f && s1y == 0.f && s1z == 0.f;
        const float q0 = (fresh || b0 || (same0 && i0 < self)) ? w0.x : w0.y;
        const float q1 = (fresh || b0 || (same0 && i1 < self)) ? w1.x : w1.y;
        const float q2 = (fresh || b1 || (same1 && i2 < self)) ? w2.x : w2.y;
        const float q3 = (fresh || b1 || (same1 && i3 < self)) ? w3.x : w3.y;
        float vw0, pt0, vw1, pt1, vw2, pt2, vw3, pt3;
        wall_pair<KERNEL>(kc, e0x, e0y, e0z, p0, u0, q0, v0, gy, vw0, pt0);
        wall_pair<KERNEL>(kc, e0x, e0y, e0z, p1, u1, q1, v1, gy, vw1, pt1);
        wall_pair<KERNEL>(kc, k.ex, k.ey, k.ez, p2, u2, q2, v2, gy, vw2, pt2);
        wall_pair<KERNEL>(kc, k.ex, k.ey, k.ez, p3, u3, q3, v3, gy, vw3, pt3);
        Sv0 = fmaf(vw0, u0.x, Sv0); Sv1 = fmaf(vw0, u0.y, Sv1); Sv2 = fmaf(vw0, u0.z, Sv2); Sp = fmaf(vw0, pt0, Sp);
        Sv0 = fmaf(vw1, u1.x, Sv0); Sv1 = fmaf(vw1, u1.y, Sv1); Sv2 = fmaf(vw1, u1.z, Sv2); Sp = fmaf(vw1, pt1, Sp);
        Sv0 = fmaf(vw2, u2.x, Sv0); Sv1 = fmaf(vw2, u2.y, Sv1); Sv2 = fmaf(vw2, u2.z, Sv2); Sp = fmaf(vw2, pt2, Sp);
        Sv0 = fmaf(vw3, u3.x, Sv0); Sv1 = fmaf(vw3, u3.y, Sv1); Sv2 = fmaf(vw3, u3.z, Sv2); Sp = fmaf(vw3, pt3, Sp);
    }
    (void)c_lo; (void)c_hi;
    if (!work) return;
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
//   d_rho_i = rho~_i sum_j V_j (v~_i - v~_j) . gradW_ij
//   d_v_i   = g + sum_j [ 2(dim+2) nu V_j min(v_ij . x_ij, 0) / (r^2 + 0.01 h^2) {1 | rho0 / rho~_i} - rho0 V_j (p_i/rho~_i^2 + p_j/rho~_j^2) ] gradW_ij
struct FluidI { float vx, vy, vz, pr, visc_f, visc_w, h2, nrho0; };
// one pair: ddc = V_j s (v_ij . x_ij), and the coefficient cf with  d_v += cf * d
template <int KERNEL>
__device__ __forceinline__ void fluid_pair(const KernConst &kc, const FluidI &I, float ex, float ey, float ez, const F4 pj,
                                           const F4 qj, bool valid, float &ddc, float &cf, float &dx, float &dy, float &dz) {
    dx = ex - pj.x; dy = ey - pj.y; dz = ez - pj.z;
    const float r2 = dist2(dx, dy, dz);
    const float s = fastdW<KERNEL>(kc, r2);
    const float ux = I.vx - qj.x, uy = I.vy - qj.y, uz = I.vz - qj.z;
    const float vx = fmaf(uz, dz, fmaf(uy, dy, ux * dx));          // v_ij . x_ij
    const float Vs = valid ? fabsf(pj.w) * s : 0.f;
    ddc = Vs * vx;
    const float visc = (pj.w < 0.f ? I.visc_w : I.visc_f) * fminf(vx, 0.f) * __fdividef(1.f, r2 + I.h2);
    cf = Vs * fmaf(I.nrho0, I.pr + qj.w, visc);
}

template <int KERNEL>
__global__ void __launch_bounds__(BT) k_tile_fluid(DevF c, TileGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileShared<2> &sh = *reinterpret_cast<TileShared<2> *>(smem_raw);
    unsigned *smask = reinterpret_cast<unsigned *>(smem_raw + sizeof(TileShared<2>));   // [NW][BT]
    const int col = blockIdx.x / g.nseg, seg = blockIdx.x - col * g.nseg;
    if (col / g.ny < c.own0 || col / g.ny >= c.own1) return;        // ghost column of a slab (block-uniform)
    const int f0 = seg * ZB, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = f0 + w;
    const int gcell = col * g.nF + f;
    int is = 0, nc = 0;
    if (f < g.nF) { is = cell_start(c.cell_end, gcell); nc = c.cell_end[gcell] - is; }
    if (nc > 0 && c.cellflag[gcell]) nc = 0;
    const int i = is + lane;
    const bool work = lane < nc && c.type[i] == 1;
    if (!__syncthreads_or(work)) return;
    const int NW = g.nR * 3;
    const unsigned nz = load_masks(c, smask, NW, i, work);
    if (!tile_setup<2>(c, g, sh, col, f0, c.ps4, c.pk4)) return;
    build_ctab<2>(c, g, sh, w, lane);
    const F4 *A = sh.P[0], *B = sh.P[1];
    const F4 *ct = sh.ctab + w * 27;
    const int rc = g.nR / 2;
    const int ci = sh.cb[rc * CBW + w + 1] + (work ? lane : 0);
    const F4 pi = A[ci], qi = B[ci];                               // qi = v~_i, p_i / rho~_i^2
    const float rhoi = work ? c.vt4[i].w : 1.f;
    const KernConst kc = kern_const(c);
    FluidI I;
    I.vx = qi.x; I.vy = qi.y; I.vz = qi.z; I.pr = qi.w;
    I.visc_f = c.visc_coef; I.visc_w = c.visc_coef * c.rho0T / rhoi;   // wc:41-44
    I.h2 = c.h2_001; I.nrho0 = -c.rho0T;
    float dd = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
    Cursor k;
    k.m = 0; k.nz = nz; k.a = 0; k.ex = k.ey = k.ez = 0.f;
    while (true) {
        cursor_jump(k, smask, ct, pi);
        if (!__any_sync(0xffffffffu, k.m != 0)) break;              // warp-uniform round: lanes reconverge here
        int i0, i1, i2, i3;
        bool v0, v1, v2, v3;
        cursor_take2(k, i0, i1, v0, v1);
        const float e0x = k.ex, e0y = k.ey, e0z = k.ez;
        cursor_jump(k, smask, ct, pi);
        cursor_take2(k, i2, i3, v2, v3);
        const F4 p0 = A[i0], q0 = B[i0], p1 = A[i1], q1 = B[i1], p2 = A[i2], q2 = B[i2], p3 = A[i3], q3 = B[i3];
        float d0, c0, x0, y0, z0, d1, c1, x1, y1, z1, d2, c2, x2, y2, z2, d3, c3, x3, y3, z3;
        fluid_pair<KERNEL>(kc, I, e0x, e0y, e0z, p0, q0, v0, d0, c0, x0, y0, z0);
        fluid_pair<KERNEL>(kc, I, e0x, e0y, e0z, p1, q1, v1, d1, c1, x1, y1, z1);
        fluid_pair<KERNEL>(kc, I, k.ex, k.ey, k.ez, p2, q2, v2, d2, c2, x2, y2, z2);
        fluid_pair<KERNEL>(kc, I, k.ex, k.ey, k.ez, p3, q3, v3, d3, c3, x3, y3, z3);
        dd += d0; a0 = fmaf(c0, x0, a0); a1 = fmaf(c0, y0, a1); a2 = fmaf(c0, z0, a2);
        dd += d1; a0 = fmaf(c1, x1, a0); a1 = fmaf(c1, y1, a1); a2 = fmaf(c1, z1, a2);
        dd += d2; a0 = fmaf(c2, x2, a0); a1 = fmaf(c2, y2, a1); a2 = fmaf(c2, z2, a2);
        dd += d3; a0 = fmaf(c3, x3, a0); a1 = fmaf(c3, y3, a1); a2 = fmaf(c3, z3, a2);
    }
    if (!work) return;
    c.d_rho[i] = dd * rhoi;
    F4 dv; dv.x = a0 + c.g[0]; dv.y = a1 + c.g[1]; dv.z = a2 + c.g[2]; dv.w = 0.f;
    c.d_vel[i] = dv;
}

// ------------------------------------------------------------------------------------------------ host side
static size_t smem_mask(int NW) { return sizeof(TileShared<1>) + (size_t)NW * BT * 4; }
static size_t smem_pass(int NW = 27) { return sizeof(TileShared<2>) + (size_t)NW * BT * 4; }
static size_t smem_wall(int NW = 27) { return sizeof(TileShared<3>) + (size_t)NW * BT * 4; }

template <int KERNEL> static int set_attrs(SphCtx *c) {
    SPH_CHECK(c, cudaFuncSetAttribute(k_tile_mask<KERNEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mask(27)));
    SPH_CHECK(c, cudaFuncSetAttribute(k_tile_wall<KERNEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wall()));
    SPH_CHECK(c, cudaFuncSetAttribute(k_tile_fluid<KERNEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pass()));
    return 0;
}
static int ensure_attrs(SphCtx *c) {
    static bool done = false;
    if (done) return 0;
    int r = set_attrs<0>(c);
    if (!r) r = set_attrs<1>(c);
    done = r == 0;
    return r;
}

int tile_mask(SphCtx *c) {
    DevF d = make_dev<float>(c);
    const TileGeom g = make_geom(c->p.dim, d.gn);
    const int nblk = g.nslow0 * g.nslow1 * g.nseg;
    int r = ensure_attrs(c);
    if (r) return r;
    SPH_CHECK(c, cudaMemsetAsync(d.cellflag, 0, (size_t)c->C, c->stream));
    SPH_CHECK(c, cudaMemsetAsync(d.nflag, 0, 4, c->stream));
    SPH_PROF(c, K_TILE_MASK);
    if (c->p.kernel == 0) k_tile_mask<0><<<nblk, BT, smem_mask(g.nR * 3), c->stream>>>(d, g);
    else k_tile_mask<1><<<nblk, BT, smem_mask(g.nR * 3), c->stream>>>(d, g);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// WCSPH one_step (wc:82-126) on the tile path; flagged cells are completed by the generic kernels (flagged_only).
int tile_wc_prep_and_wall(SphCtx *c) {
    DevF d = make_dev<float>(c);
    const TileGeom g = make_geom(c->p.dim, d.gn);
    const int nblk = g.nslow0 * g.nslow1 * g.nseg, n = (int)c->n;
    SPH_PROF(c, K_WC_EOS);
    k_tile_prep<<<blocks_for(n, 256), 256, 0, c->stream>>>(d);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_TILE_WALL);
    if (c->p.kernel == 0) k_tile_wall<0><<<nblk, BT, smem_wall(g.nR * 3), c->stream>>>(d, g);
    else k_tile_wall<1><<<nblk, BT, smem_wall(g.nR * 3), c->stream>>>(d, g);
    SPH_LAUNCH_CHECK(c);
    return 0;
}
int tile_wc_fluid(SphCtx *c) {
    DevF d = make_dev<float>(c);
    const TileGeom g = make_geom(c->p.dim, d.gn);
    const int nblk = g.nslow0 * g.nslow1 * g.nseg;
    SPH_PROF(c, K_TILE_FLUID);
    if (c->p.kernel == 0) k_tile_fluid<0><<<nblk, BT, smem_pass(g.nR * 3), c->stream>>>(d, g);
    else k_tile_fluid<1><<<nblk, BT, smem_pass(g.nR * 3), c->stream>>>(d, g);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

}  // namespace sph
