// sweeps_tile.cu -- cell-major neighbour sweeps for the WCSPH hot loop (MIXED precision), hand-written for sm_100a.
//
// Work decomposition.  A thread block owns a footprint of BX x BY cell columns (3D; BX columns in 2D) times ZB
// consecutive cells along the fastest grid axis (z in 3D, y in 2D); warp w owns ONE cell of the footprint and lane l
// owns the l-th particle of that cell.  Because the flattened cell id is fastest-axis-major (ps:221-222), the
// particles of the (BX+2) x (BY+2) x (ZB+2) cells a block needs form NR = (BX+2)(BY+2) contiguous spans ("runs") of
// the sorted arrays: each run is brought into shared memory with ONE 1-D TMA bulk copy per payload array
// (cp.async.bulk ... mbarrier::complete_tx), raw, with no transformation.  Persistent blocks pull footprint segments
// that hold work from per-step work lists with an atomic cursor.
//
// Kernels, in the order of a step:
//   k_tile_worklist / k_wall_cells   which footprint segments / wall cells have work (compacted lists)
//   k_tile_mask                      the neighbour predicate, ONCE per step (positions are frozen inside a step), into one
//                                    32-bit word per (particle, stencil cell): packed float32 pairs (FADD2/FMUL2/FFMA2)
//                                    over SoA tiles, every cell pair evaluated once + warp bit-matrix transpose, double-
//                                    buffered tiles.  Wall particles keep only their FLOW neighbours (all a wall sum ever
//                                    reads, base:647-669): dry walls end up with empty masks and cost nothing.
//   k_tile_prep                      EOS, tile payloads, dry walls (+ advect_LF_half inside sph_step)          pointwise
//   k_wall_gather                    wall pass (wc:90-103): one warp per wall cell in reach of flow, payloads gathered via L1
//   k_tile_fluid                     fluid pass (wc:108-126): set bits only, four neighbours per round in stencil order
//                                    (cells x-major / z-fastest, j ascending), mask words streamed one cell ahead;
//                                    optionally records / replays neighbour round lists
//   k_tile_shepard                   calc_CSPM_f alone (the stand-alone API call; inside a step the first wall / fluid
//                                    pass forms the same sums)
// Cells that cannot be represented (more than 32 particles in a stencil cell, or a tile that overflows) are flagged
// together with their stencil neighbours and processed by the generic kernels of sweeps.cu.
//
// Replaces, for this configuration: calc_CSPM_f (base:386-398), WCSPH one_step loops A and B (wc:82-126).
#include "sph_host.h"

namespace sph {

constexpr int ZB = 4;                 // cells per footprint along the fastest axis
typedef Vec4<float> F4;
typedef Dev<float> DevF;

template <bool D3, int BX_, int BY_> struct Foot {
    static constexpr bool d3 = D3;
    static constexpr int BX = BX_, BY = D3 ? BY_ : 1;
    static constexpr int NRX = BX + 2, NRY = D3 ? BY + 2 : 1, NR = NRX * NRY;       // runs
    static constexpr int NWARP = BX * BY * ZB, BT = NWARP * 32;
    static constexpr int NW = D3 ? 27 : 9, CENTRE = NW / 2;                          // stencil cells
    static constexpr int CBW = ZB + 3;                                               // cell boundaries per run
    static constexpr int REST = NR * (ZB + 2) * (D3 ? 27 : 9);                       // rest-lattice tile population
    static constexpr int CAP = (REST + REST / 6 + 63) / 64 * 64;                     // + ~17 % head-room
    static constexpr int SENT = CAP;                                                 // all-zero sentinel entry
};

// NP payload arrays of CAP float4 (+8 entries of slack: the sentinel, and the chunked test loop may read past a cell)
template <class FT, int NP> struct TileShared {
    F4 P[NP][FT::CAP + 8];
    unsigned long long bar;           // mbarrier
    int cb[FT::NR * FT::CBW];         // tile index of the first particle of each (run, cell)
    int gdelta[FT::NR];               // global index = tile index + gdelta[run]
    int total, overflow, item;
    F4 ctab[FT::NWARP * FT::NW];      // per (warp, stencil cell): shift xyz, tile index of the cell's first particle | flags
    // pipelined loop (tile_stage / tile_issue): spans of the NEXT work item, computed while the current tile is in flight
    int ncb[FT::NR * FT::CBW], ngdelta[FT::NR], nS[FT::NR], nlen[FT::NR], nroff[FT::NR], ntotal;
    int blkq[2];                      // work items of iterations k and k+1 (by parity; -1 = the list is exhausted)
    int claim;                        // 0 until a warp has taken the staging duty of this iteration
    unsigned flagbits, nflagbits;     // flagged own cells (bit = warp of the footprint): active / staged
};
constexpr unsigned CT_BEFORE = 1u << 30, CT_SAME = 1u << 29, CT_IDX = (1u << 24) - 1, CT_CC = 0x1fu << 24;   // CT_CC: the stencil cell

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
__device__ __forceinline__ float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ------------------------------------------------------------------------------------------------ geometry
struct TileGeom { int nF, n0, n1, nb0, nb1, nseg; };   // fast axis cells; slow axes (x, y | x, 1); blocks per slow axis; segments
template <class FT> __host__ inline TileGeom make_geom(const int gn[3]) {
    TileGeom g;
    if (FT::d3) { g.nF = gn[2]; g.n0 = gn[0]; g.n1 = gn[1]; }
    else { g.nF = gn[1]; g.n0 = gn[0]; g.n1 = 1; }
    g.nb0 = (g.n0 + FT::BX - 1) / FT::BX;
    g.nb1 = (g.n1 + FT::BY - 1) / FT::BY;
    g.nseg = (g.nF + ZB - 1) / ZB;
    return g;
}
__device__ __forceinline__ int cell_start(const int *cell_end, int g) { return g > 0 ? cell_end[g - 1] : 0; }

// what a warp knows about its cell
struct WarpCell { int b0, b1, f0, wx, wy, wz, cx, cy, f, gcell, is, nc; };
template <class FT> __device__ __forceinline__ WarpCell warp_cell(const DevF &c, const TileGeom &g, int blk) {
    WarpCell w;
    const int seg = blk % g.nseg, t = blk / g.nseg;
    w.b1 = t % g.nb1; w.b0 = t / g.nb1; w.f0 = seg * ZB;
    const int wi = threadIdx.x >> 5;
    w.wz = wi % ZB; w.wy = (wi / ZB) % FT::BY; w.wx = wi / (ZB * FT::BY);
    w.cx = w.b0 * FT::BX + w.wx; w.cy = w.b1 * FT::BY + w.wy; w.f = w.f0 + w.wz;
    w.gcell = 0; w.is = 0; w.nc = 0;
    if (w.cx < g.n0 && w.cy < g.n1 && w.f < g.nF && w.cx >= c.own0 && w.cx < c.own1) {     // ghost columns of a slab idle
        w.gcell = (w.cx * g.n1 + w.cy) * g.nF + w.f;
        w.is = cell_start(c.cell_end, w.gcell);
        w.nc = c.cell_end[w.gcell] - w.is;
    }
    return w;
}
// stencil cell cc of a warp -> (run, boundary index) and the offsets
template <class FT> __device__ __forceinline__ void stencil(int cc, int &ox, int &oy, int &of) {
    if (FT::d3) { ox = cc / 9 - 1; oy = (cc / 3) % 3 - 1; of = cc % 3 - 1; }
    else { ox = cc / 3 - 1; oy = 0; of = cc % 3 - 1; }
}
template <class FT> __device__ __forceinline__ int stencil_cb(const WarpCell &w, int ox, int oy, int of) {
    const int r = (w.wx + ox + 1) * FT::NRY + (FT::d3 ? w.wy + oy + 1 : 0);
    return r * FT::CBW + (w.wz + of + 1);
}

// Computes spans and cell boundaries, issues the TMA copies of the NP payload arrays and waits for them.
// Returns false (uniformly) when the tile does not fit.
// once per block: the mbarrier (reused with alternating parity by every work item) and the sentinel entries
template <class FT, int NP> __device__ __forceinline__ void tile_init(TileShared<FT, NP> &sh) {
    if (threadIdx.x == 0) {
        mbar_init(&sh.bar, 1);
        F4 z; z.x = z.y = z.z = z.w = 0.f;
#pragma unroll
        for (int p = 0; p < NP; p++) sh.P[p][FT::SENT] = z;
    }
    __syncthreads();
}
template <class FT, int NP>
__device__ __forceinline__ bool tile_setup(const DevF &c, const TileGeom &g, TileShared<FT, NP> &sh, const WarpCell &w,
                                           unsigned parity, const F4 *src0, const F4 *src1 = nullptr, const F4 *src2 = nullptr) {
    const int tid = threadIdx.x;
    const int f0 = w.f0, f_lo = max(f0 - 1, 0), f_hi = min(f0 + ZB, g.nF - 1);
    if (tid < 32) {
        int len = 0, S = 0, gb = 0;
        bool valid = false;
        if (tid < FT::NR) {
            const int n0 = w.b0 * FT::BX + tid / FT::NRY - 1;
            const int n1 = FT::d3 ? w.b1 * FT::BY + tid % FT::NRY - 1 : 0;
            valid = n0 >= 0 && n0 < g.n0 && n1 >= 0 && n1 < g.n1;
            if (valid) {
                gb = (n0 * g.n1 + n1) * g.nF;
                S = cell_start(c.cell_end, gb + f_lo);
                len = c.cell_end[gb + f_hi] - S;
            }
        }
        int inc = len;                                            // inclusive scan over the first NR lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += t;
        }
        const int roff = inc - len;
        if (tid < FT::NR) {
            sh.gdelta[tid] = S - roff;
            for (int k = 0; k < FT::CBW; k++) {
                const int f = f0 - 1 + k;
                int v;
                if (!valid || f < f_lo) v = roff;
                else if (f > f_hi) v = roff + len;
                else v = roff + cell_start(c.cell_end, gb + f) - S;
                sh.cb[tid * FT::CBW + k] = v;
            }
        }
        const int total = __shfl_sync(0xffffffffu, inc, FT::NR - 1);
        if (tid == 0) {
            sh.total = total;
            sh.overflow = total > FT::CAP;
        }
        __syncwarp();
        if (total > FT::CAP) {                                    // nothing is loaded: complete the phase so that the parity still flips
            if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sh.bar)) : "memory");
        } else {
            if (tid == 0) mbar_expect_tx(&sh.bar, (unsigned)(total * 16 * NP));
            __syncwarp();
            if (tid < FT::NR && len > 0) {
                tma_load_1d(&sh.P[0][roff], src0 + S, (unsigned)(len * 16), &sh.bar);
                if (NP > 1) tma_load_1d(&sh.P[1][roff], src1 + S, (unsigned)(len * 16), &sh.bar);
                if (NP > 2) tma_load_1d(&sh.P[NP - 1][roff], src2 + S, (unsigned)(len * 16), &sh.bar);
            }
        }
    }
    __syncthreads();
    if (sh.overflow) return false;
    mbar_wait(&sh.bar, parity);
    return true;
}
// ---- pipelined variant: the global-memory latency of a work item's set-up is taken off the critical path ----
// Measured with clock64 on C4 before this existed: of the 82 k cycles a block spent per work item, 14 k were set-up
// (atomic cursor -> list -> cell_end -> spans -> TMA, plus every warp's own cell_end / flag / nzw look-ups) and 9 k the
// wait for the slowest warp at the end.  Now:
//   tile_stage  (the FIRST warp that finishes its cell; it would idle at the end barrier anyway): fetches the next work
//               item (atomic cursor + list) and computes its spans, cell boundaries and flagged-cell bits into the
//               staging members -- all the dependent global reads happen here;
//   tile_issue  (warp 0, right after the barrier that frees the tile): staging -> active members and the TMA copies;
//               shared-memory work only;
//   tile_begin  (all): barrier; a warp reads its cell's first particle / count from the active boundaries (no global
//               read), issues its nzw load, then waits for the copies (mbarrier, one phase per work item).
template <class FT, int NP>
__device__ __forceinline__ void tile_stage(const DevF &c, const TileGeom &g, TileShared<FT, NP> &sh, int blk) {
    const int lane = threadIdx.x & 31;
    const int seg = blk % g.nseg, tq = blk / g.nseg, b1 = tq % g.nb1, b0 = tq / g.nb1;
    const int f0 = seg * ZB, f_lo = max(f0 - 1, 0), f_hi = min(f0 + ZB, g.nF - 1);
    int len = 0, S = 0;
    int e[FT::CBW];                                               // cell_end of cells f0-2 .. f0+ZB (clamped), one round trip
    bool valid = false;
    if (lane < FT::NR) {
        const int n0 = b0 * FT::BX + lane / FT::NRY - 1;
        const int n1 = FT::d3 ? b1 * FT::BY + lane % FT::NRY - 1 : 0;
        valid = n0 >= 0 && n0 < g.n0 && n1 >= 0 && n1 < g.n1;
        const int gb = valid ? (n0 * g.n1 + n1) * g.nF : 0;
#pragma unroll
        for (int k = 0; k < FT::CBW; k++) {                       // e[k] = start of cell f0 - 1 + k = cell_end of the cell before
            const int f = min(max(f0 - 1 + k, f_lo), f_hi + 1);   // clamped: below f_lo -> start of f_lo, above -> end of f_hi
            const int gi = gb + f - 1;
            e[k] = (valid && gi >= 0) ? c.cell_end[gi] : 0;
        }
        if (valid) { S = e[f_lo - (f0 - 1)]; len = e[min(f_hi + 1, f0 + ZB + 1) - (f0 - 1)] - S; }
    }
    int inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const int roff = inc - len;
    if (lane < FT::NR) {
        sh.ngdelta[lane] = S - roff;
        sh.nS[lane] = S; sh.nlen[lane] = len; sh.nroff[lane] = roff;
#pragma unroll
        for (int k = 0; k < FT::CBW; k++) sh.ncb[lane * FT::CBW + k] = valid ? roff + e[k] - S : roff;
    }
    // flagged own cells (handled by the generic kernels): one bit per warp of the footprint
    bool fl = false;
    if (lane < FT::NWARP) {
        const int wz = lane % ZB, wy = (lane / ZB) % FT::BY, wx = lane / (ZB * FT::BY);
        const int cx = b0 * FT::BX + wx, cy = b1 * FT::BY + wy, f = f0 + wz;
        if (cx < g.n0 && cy < g.n1 && f < g.nF) fl = c.cellflag[(cx * g.n1 + cy) * g.nF + f] != 0;
    }
    const unsigned flb = __ballot_sync(0xffffffffu, fl);
    const int total = __shfl_sync(0xffffffffu, inc, FT::NR - 1);
    if (lane == 0) { sh.ntotal = total; sh.nflagbits = flb; }
    __syncwarp();
}
template <class FT, int NP>
__device__ __forceinline__ void tile_issue(TileShared<FT, NP> &sh, const F4 *src0, const F4 *src1 = nullptr, const F4 *src2 = nullptr) {
    const int tid = threadIdx.x;                                  // < 32
    for (int k = tid; k < FT::NR * FT::CBW; k += 32) sh.cb[k] = sh.ncb[k];
    const int total = sh.ntotal;
    int S = 0, len = 0, roff = 0;
    if (tid < FT::NR) { sh.gdelta[tid] = sh.ngdelta[tid]; S = sh.nS[tid]; len = sh.nlen[tid]; roff = sh.nroff[tid]; }
    if (tid == 0) { sh.total = total; sh.overflow = total > FT::CAP; sh.flagbits = sh.nflagbits; sh.claim = 0; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the block's reads of the old tile precede the bulk copies
    __syncwarp();
    if (total > FT::CAP) {
        if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sh.bar)) : "memory");
    } else {
        if (tid == 0) mbar_expect_tx(&sh.bar, (unsigned)(total * 16 * NP));
        __syncwarp();
        if (tid < FT::NR && len > 0) {
            tma_load_1d(&sh.P[0][roff], src0 + S, (unsigned)(len * 16), &sh.bar);
            if (NP > 1) tma_load_1d(&sh.P[1][roff], src1 + S, (unsigned)(len * 16), &sh.bar);
            if (NP > 2) tma_load_1d(&sh.P[NP - 1][roff], src2 + S, (unsigned)(len * 16), &sh.bar);
        }
    }
}
// the warp's cell from the work item number alone (arithmetic), then -- after the barrier -- first particle and count
// from the active cell boundaries.  Cells outside the grid, in ghost columns of a slab or flagged get nc = 0.
template <class FT, int NP>
__device__ __forceinline__ WarpCell tile_begin(const DevF &c, const TileGeom &g, TileShared<FT, NP> &sh, int blk) {
    WarpCell w;
    const int seg = blk % g.nseg, t = blk / g.nseg;
    w.b1 = t % g.nb1; w.b0 = t / g.nb1; w.f0 = seg * ZB;
    const int wi = threadIdx.x >> 5;
    w.wz = wi % ZB; w.wy = (wi / ZB) % FT::BY; w.wx = wi / (ZB * FT::BY);
    w.cx = w.b0 * FT::BX + w.wx; w.cy = w.b1 * FT::BY + w.wy; w.f = w.f0 + w.wz;
    w.gcell = 0; w.is = 0; w.nc = 0;
    const bool inside = w.cx < g.n0 && w.cy < g.n1 && w.f < g.nF && w.cx >= c.own0 && w.cx < c.own1;
    if (inside) w.gcell = (w.cx * g.n1 + w.cy) * g.nF + w.f;
    __syncthreads();                                               // the active members of this work item are visible
    if (inside && !((sh.flagbits >> wi) & 1u)) {
        const int q = stencil_cb<FT>(w, 0, 0, 0);
        const int a = sh.cb[q];
        w.is = a + sh.gdelta[q / FT::CBW];
        w.nc = sh.cb[q + 1] - a;
    }
    return w;
}
template <class FT, int NP> __device__ __forceinline__ bool tile_wait(TileShared<FT, NP> &sh, unsigned parity) {
    if (sh.overflow) return false;
    mbar_wait(&sh.bar, parity);
    return true;
}
// end of a warp's work on the item: the first warp to get here fetches the next work item and stages it.  A block
// thus holds ONE item beyond the one it is working on -- fetching two ahead (one more round trip hidden) cost 2-4 % of
// the pass on an 8-GPU slab, where a block sees only ~10 items and reserved items lengthen the tail.
template <class FT, int NP>
__device__ __forceinline__ void tile_finish(const DevF &c, const TileGeom &g, TileShared<FT, NP> &sh, const int *list, int items,
                                            int *cursor, unsigned k) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    int mine = 0;
    if (lane == 0) mine = atomicExch(&sh.claim, 1) == 0;
    mine = __shfl_sync(0xffffffffu, mine, 0);
    if (!mine) return;
    int nb = -1;
    if (lane == 0) {
        const int a = atomicAdd(cursor, 1);
        nb = a < items ? list[a] : -1;
    }
    nb = __shfl_sync(0xffffffffu, nb, 0);
    if (nb >= 0) tile_stage<FT, NP>(c, g, sh, nb);
    if (lane == 0) sh.blkq[(k + 1u) & 1u] = nb;
}
// The loop.  BODY(blk, parity, k_) must call tile_begin once and tile_finish once on every path.
#define TILE_PIPELINED_LOOP(FT_, NP_, SH, LIST, COUNT, CURSOR, SRC0, SRC1, CALL)                       \
    {                                                                                                     \
        const int items_ = *(COUNT);                                                                      \
        const int *list_ = (LIST);                                                                        \
        int *cursor_ = (CURSOR);                                                                          \
        if (threadIdx.x == 0) {                                                                           \
            const int a0_ = atomicAdd(cursor_, 1);                                                        \
            (SH).blkq[0] = a0_ < items_ ? list_[a0_] : -1;                                                \
        }                                                                                                 \
        __syncthreads();                                                                                  \
        if (threadIdx.x < 32 && (SH).blkq[0] >= 0) tile_stage<FT_, NP_>(c, g, (SH), (SH).blkq[0]);        \
        TT_DECL                                                                                           \
        for (unsigned k_ = 0;; k_++) {                                                                    \
            TT_A                                                                                          \
            const int blk = (SH).blkq[k_ & 1u];                                                           \
            if (blk < 0) break;                                                                           \
            if (threadIdx.x < 32) tile_issue<FT_, NP_>((SH), (SRC0), (SRC1));                             \
            const unsigned parity = k_ & 1u;                                                              \
            CALL;                                                                                         \
            TT_END                                                                                        \
            __syncthreads();                                                                              \
            TT_D                                                                                          \
        }                                                                                                 \
        TT_PRINT                                                                                          \
    }

// Persistent blocks walk a work list of footprint segments.  BODY(blk, parity) returns true when it used the tile
// (block-uniform), which flips the mbarrier parity for the next item.
#ifdef TILE_TIMING
__device__ unsigned long long g_tt[8];
#define TT_DECL long long tt_a_ = 0, tt_setup_ = 0, tt_end_ = 0;
#define TT_A tt_a_ = clock64();
#define TT_END tt_end_ = clock64();
#define TT_D                                                                                              \
    if ((threadIdx.x == 0 || threadIdx.x == blockDim.x - 32) && tt_setup_ != 0) {                         \
        const long long td_ = clock64();                                                                  \
        const int o_ = threadIdx.x == 0 ? 0 : 3;                                                          \
        atomicAdd(&g_tt[o_], (unsigned long long)(tt_setup_ - tt_a_));                                    \
        atomicAdd(&g_tt[o_ + 1], (unsigned long long)(tt_end_ - tt_setup_));                              \
        atomicAdd(&g_tt[o_ + 2], (unsigned long long)(td_ - tt_end_));                                    \
        if (threadIdx.x == 0) atomicAdd(&g_tt[6], 1ull);                                                  \
    }                                                                                                     \
    tt_setup_ = 0;
#define TT_PRINT                                                                                          \
    if (blockIdx.x == 0 && threadIdx.x == 0)                                                              \
        printf("TT items %llu  w0: pro %llu cmp %llu bar %llu   w15: pro %llu cmp %llu bar %llu\n", g_tt[6], g_tt[0], g_tt[1], g_tt[2], g_tt[3], g_tt[4], g_tt[5]);
#define TT_ARG , &tt_setup_
#else
#define TT_DECL
#define TT_A
#define TT_END
#define TT_D
#define TT_PRINT
#define TT_ARG
#endif
#define TILE_PERSISTENT_LOOP(SH, LIST, COUNT, CURSOR, CALL)                           \
    {                                                                                 \
        unsigned uses_ = 0;                                                           \
        const int items_ = *(COUNT);                                                  \
        TT_DECL                                                                       \
        while (true) {                                                                \
            TT_A                                                                      \
            if (threadIdx.x == 0) (SH).item = atomicAdd((CURSOR), 1);                 \
            __syncthreads();                                                          \
            const int it_ = (SH).item;                                                \
            if (it_ >= items_) break;                                                 \
            const int blk = (LIST)[it_];                                              \
            const unsigned parity = uses_ & 1u;                                       \
            if (CALL) uses_++;                                                        \
            TT_END                                                                    \
            __syncthreads();                                                          \
            TT_D                                                                      \
        }                                                                             \
        TT_PRINT                                                                      \
    }

// per-warp table of the stencil cells, so that advancing to the next cell costs one shared load
template <class FT, int NP>
__device__ __forceinline__ void build_ctab(const DevF &c, TileShared<FT, NP> &sh, const WarpCell &w, int lane) {
    if (lane < FT::NW) {
        int ox, oy, of;
        stencil<FT>(lane, ox, oy, of);
        F4 t;
        if (FT::d3) { t.x = (float)ox * c.gsT; t.y = (float)oy * c.gsT; t.z = (float)of * c.gsT; }
        else { t.x = (float)ox * c.gsT; t.y = (float)of * c.gsT; t.z = 0.f; }
        unsigned v = (unsigned)sh.cb[stencil_cb<FT>(w, ox, oy, of)] | ((unsigned)lane << 24);
        if (lane < FT::CENTRE) v |= CT_BEFORE;                 // stencil order == ascending cell id
        if (lane == FT::CENTRE) v |= CT_SAME;
        t.w = __uint_as_float(v);
        sh.ctab[(threadIdx.x >> 5) * FT::NW + lane] = t;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------ float32 kernels
// Division-free float32 forms of base:278-358 on the squared distance (MUFU.RSQ instead of IEEE sqrt/div).
// The reference's guards (r > 1e-8, q <= 2) become clamps: r2 is clamped from below before the rsqrt and (1 - q/2)
// resp. (2 - q) from below at 0 -- pairs selected by the mask have r < support = 2h up to rounding.
struct KernConst { float hinv, eps2, knorm, c_grad; };   // c_grad = -5 knorm / h^2 (Wendland) | knorm / h^2 (cubic)
__device__ __forceinline__ KernConst kern_const(const DevF &c) {
    KernConst k;
    k.hinv = c.hinv; k.eps2 = c.eps * c.eps; k.knorm = c.knorm;
    k.c_grad = (c.kernel == 0 ? 1.f : -5.f) * c.knorm * c.hinv * c.hinv;
    return k;
}
template <int KERNEL> __device__ __forceinline__ float fastW(const KernConst &k, float r2) {
    const float r2c = fmaxf(r2, k.eps2), r = r2c * rsqrt_fast(r2c);
    if (KERNEL == 1) {
        const float q1 = fmaxf(fmaf(-0.5f * k.hinv, r, 1.f), 0.f), q2 = q1 * q1;
        return (k.knorm * q2) * (q2 * fmaf(2.f * k.hinv, r, 1.f));
    }
    const float q = r * k.hinv, t = fmaxf(2.f - q, 0.f);
    return q <= 1.f ? k.knorm * (q * q * fmaf(0.5f, q, -1.f) + (float)(2.0 / 3.0)) : k.knorm * (1.f / 6.f) * t * t * t;
}

// One term of the Shepard sum (calc_CSPM_f): S += max(V_j signed, 0) * W.  Written with explicit roundings so that the
// stand-alone kernel and the sums fused into the wall / fluid passes give bit-identical CSPM_f (no FMA contraction).
__device__ __forceinline__ float shep_add(float s, float vsigned, float w) { return __fadd_rn(s, __fmul_rn(fmaxf(vsigned, 0.f), w)); }

// ------------------------------------------------------------------------------------------------ bit iteration
// Per-lane cursor over the set bits of the neighbour masks in stencil order.  A round takes up to four neighbours:
// two from the current cell, then (after an optional jump to the next non-empty cell) two more.  Slots that find no
// bit read the all-zero sentinel entry (volume 0) and contribute exactly zero, so the summation order (cells
// x-major, j ascending) is kept.  The mask word of the NEXT non-empty cell is always in flight from global memory.
// Mask storage: the words of a cell's particles form one contiguous block.  For the cell whose first particle is `is`
// and which holds nc particles, the word of its k-th particle for stencil cell cc is mask[is * NW + cc * nc + k]: a
// warp reads / writes nc consecutive words per stencil cell, and all offsets inside a block fit 32 bits.
__device__ __forceinline__ unsigned *mask_row(unsigned *mask, int is, int nw, int k) { return mask + ((size_t)is * nw + k); }
struct Cursor {
    unsigned m, nz, mnext;   // remaining bits of the current cell; remaining non-empty cells; prefetched word
    unsigned flags;          // CT_BEFORE / CT_SAME of the current cell
    int a;                   // tile index of the current cell's first particle
    float ex, ey, ez;        // own coordinates in the current cell's frame
};
// Masks and the non-empty-cell bitmaps are stored BIT-REVERSED (particle / cell b at bit 31 - b): the next element in
// ascending order is then one FLO (count-leading-zeros) instead of BREV + FLO, which halves the load on the XU pipe.
__device__ __forceinline__ unsigned rbit(int b) { return 0x80000000u >> b; }
// mrow = mask_row(...) of the lane's particle; n = particles in its cell (the stride between the words of a particle)
__device__ __forceinline__ void cursor_init(Cursor &k, const unsigned *mrow, unsigned n, unsigned nz) {
    k.m = 0; k.nz = nz; k.mnext = 0; k.flags = 0; k.a = 0; k.ex = k.ey = k.ez = 0.f;
    if (nz) k.mnext = __ldg(mrow + (unsigned)__clz(nz) * n);
}
// L2 prefetch of every mask line this lane will read (issued once per work item, long before the first use)
__device__ __forceinline__ void cursor_prefetch(const unsigned *mrow, unsigned n, unsigned nz) {
    while (nz) {
        const int cc = __clz(nz);
        nz &= ~rbit(cc);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(mrow + (unsigned)cc * n));
    }
}
__device__ __forceinline__ void cursor_jump(Cursor &k, const unsigned *mrow, unsigned n, const F4 *ct, const F4 &pi) {
    const bool jump = k.m == 0 && k.nz != 0;
    if (jump) {
        const int cc = __clz(k.nz);
        k.nz &= ~rbit(cc);
        k.m = k.mnext;
        const F4 t = ct[cc];
        const unsigned v = __float_as_uint(t.w);
        k.a = (int)(v & CT_IDX); k.flags = v;
        k.ex = pi.x - t.x; k.ey = pi.y - t.y; k.ez = pi.z - t.z;
    }
    // The word of the NEXT non-empty cell is requested here, outside the divergent region and as a predicated load
    // INTO the loop-carried register: a plain assignment inside the branch makes ptxas load into a temporary and
    // move it at the reconvergence point, which waits for the load at once and exposes the whole memory latency.
    const unsigned go = jump ? k.nz : 0u;
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p ld.global.nc.u32 %0, [%1];\n}"
                 : "+r"(k.mnext) : "l"(mrow + (unsigned)(__clz(go) & 31) * n), "r"(go));
}
// takes up to two bits of the current cell: tile indices (the sentinel when there is no bit)
template <int SENT> __device__ __forceinline__ void cursor_take2(Cursor &k, int &i0, int &i1) {
    const unsigned m0 = k.m;
    const int t0 = __clz(m0);
    const unsigned m1 = m0 & ~rbit(t0 & 31);
    const int t1 = __clz(m1);
    k.m = m1 & ~rbit(t1 & 31);
    i0 = m0 ? k.a + t0 : SENT;
    i1 = m1 ? k.a + t1 : SENT;
}

// 32 x 32 bit-matrix transpose across the lanes of a warp: in: lane i holds row i, out: lane j holds column j
__device__ __forceinline__ unsigned warp_transpose32(unsigned a, int lane) {
    const unsigned mk[5] = {0x0000FFFFu, 0x00FF00FFu, 0x0F0F0F0Fu, 0x33333333u, 0x55555555u};
#pragma unroll
    for (int q = 0; q < 5; q++) {
        const int s = 16 >> q;
        const unsigned o = __shfl_xor_sync(0xffffffffu, a, s);
        a = (lane & s) ? ((a & ~mk[q]) | ((o >> s) & mk[q])) : ((a & mk[q]) | ((o << s) & ~mk[q]));
    }
    return a;
}

// ------------------------------------------------------------------------------------------------ pass 0: masks
// One launch per step, right after the grid build.  nzw (bitmap of non-zero words per particle) must be zero on entry.
//
// The predicate loop is the hot spot of the step (729 candidates per particle in 3D, halved by symmetry), and it is
// bound by instruction issue, not by the FP32 pipe.  It therefore runs on PACKED float32 pairs (sub / mul / fma
// .f32x2 -> FADD2 / FMUL2 / FFMA2): one instruction evaluates one lane's particle against TWO candidates.  Each half
// of a packed operation is the IEEE round-to-nearest scalar operation, so the predicate keeps the exact expression
// of sph_dev.cuh::dist2 / for_neighbors.  To get candidate pairs into aligned 64-bit registers the mask kernel
// stages a structure-of-arrays copy of the sweep coordinates (psx, psy, psz, flow sign psf): one LDS.128 brings the x
// (or y, z) of four consecutive candidates.  TMA needs 16-byte aligned sources, so a run is copied from the 4-aligned
// particle index below its start (the `lead` entries in front belong to another cell and are masked by position).
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// Two more candidates against one particle, pushed into the chunk word from the bottom (the caller walks a chunk from its
// last pair to its first, so candidate k of the chunk ends up at bit k).  The predicate dist2(d) < thr is read off the
// SIGN of the packed difference r2 - thr: for finite operands the rounded difference is negative exactly when r2 < thr
// (a nonzero exact difference never rounds to zero, and r2 == thr gives +0), so the bit is the same as FSETP.LT's --
// but it costs one packed subtraction on the FMA pipe per two candidates and ONE funnel shift per candidate on the
// half-rate ALU pipe, where the compare + predicated OR took two.
__device__ __forceinline__ unsigned push2(unsigned cm, u64 dx, u64 dy, u64 dz, u64 thr2) {
    float lo, hi;
    up2(sub2(fma2(dz, dz, fma2(dy, dy, mul2(dx, dx))), thr2), lo, hi);     // dist2(dx, dy, dz) - thr in each half
    cm = __funnelshift_l(__float_as_uint(hi), cm, 1);
    return __funnelshift_l(__float_as_uint(lo), cm, 1);
}
__device__ __forceinline__ u64 chunk_bits(unsigned cm, int t0) { return (u64)cm << t0; }

// Two tile buffers: while the block evaluates the cell pairs of one work item, warp 0 has already fetched the next
// item, computed its spans and issued its TMA copies into the other buffer (the tile is only 49 KB, two blocks of two
// buffers fit an SM), so the copy latency and the cell_end look-ups are off the critical path.
template <class FT> struct __align__(128) MaskTile {
    static constexpr int CAPS = FT::CAP + 8;          // the chunked test loop may read up to 7 entries past a cell
    float X[CAPS], Y[CAPS], Z[CAPS], F[CAPS];
    int cb[FT::NR * FT::CBW];         // tile index of the first particle of each (run, cell)
    int gdelta[FT::NR];               // global index = tile index + gdelta[run]
    int total, overflow;
};
template <class FT> struct MaskShared {
    MaskTile<FT> t[2];
    unsigned long long bar[2];
    int item[2];
    int4 tab[FT::NWARP][FT::NW];      // per (warp, stencil cell): tile index of its first particle, particle count,
                                      // global - tile index, code = (ox+1) | (oy+1) << 2 | (of+1) << 4 | owned << 6 | has flow << 7
};
// Warp 0 only.  Spans, cell boundaries and the TMA copies of the four SoA arrays of work item `blk` into `tl`.  Every
// run starts at a multiple of four entries in the tile and is copied from the multiple of four particles at or
// below its first particle.  Always completes exactly one phase of `bar`.
template <class FT>
__device__ __forceinline__ void mask_tile_issue(const DevF &c, const TileGeom &g, MaskTile<FT> &tl, unsigned long long *bar, int blk) {
    const int tid = threadIdx.x;                                  // < 32
    const int seg = blk % g.nseg, tq = blk / g.nseg, b1 = tq % g.nb1, b0 = tq / g.nb1;
    const int f0 = seg * ZB, f_lo = max(f0 - 1, 0), f_hi = min(f0 + ZB, g.nF - 1);
    int len = 0, S = 0, gb = 0, lead = 0, len4 = 0;
    bool valid = false;
    if (tid < FT::NR) {
        const int n0 = b0 * FT::BX + tid / FT::NRY - 1;
        const int n1 = FT::d3 ? b1 * FT::BY + tid % FT::NRY - 1 : 0;
        valid = n0 >= 0 && n0 < g.n0 && n1 >= 0 && n1 < g.n1;
        if (valid) {
            gb = (n0 * g.n1 + n1) * g.nF;
            S = cell_start(c.cell_end, gb + f_lo);
            len = c.cell_end[gb + f_hi] - S;
            if (len > 0) { lead = S & 3; len4 = (lead + len + 3) & ~3; }
        }
    }
    int inc = len4;                                               // inclusive scan over the first NR lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
    }
    const int roff = inc - len4, first = roff + lead;             // tile index of the copy / of the run's first particle
    if (tid < FT::NR) {
        tl.gdelta[tid] = S - first;
        for (int k = 0; k < FT::CBW; k++) {
            const int f = f0 - 1 + k;
            int v;
            if (!valid || f < f_lo) v = first;
            else if (f > f_hi) v = first + len;
            else v = first + cell_start(c.cell_end, gb + f) - S;
            tl.cb[tid * FT::CBW + k] = v;
        }
    }
    const int total = __shfl_sync(0xffffffffu, inc, FT::NR - 1);
    if (tid == 0) {
        tl.total = total;
        tl.overflow = total > FT::CAP;
    }
    __syncwarp();
    if (total > FT::CAP) {                                        // nothing is loaded: complete the phase so that the parity still flips
        if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    } else {
        if (tid == 0) mbar_expect_tx(bar, (unsigned)(total * 16));
        __syncwarp();
        if (tid < FT::NR && len4 > 0) {
            const int S4 = S - lead;
            const unsigned bytes = (unsigned)(len4 * 4);
            tma_load_1d(&tl.X[roff], c.psx + S4, bytes, bar);
            tma_load_1d(&tl.Y[roff], c.psy + S4, bytes, bar);
            tma_load_1d(&tl.Z[roff], c.psz + S4, bytes, bar);
            tma_load_1d(&tl.F[roff], c.psf + S4, bytes, bar);
        }
    }
}

template <class FT>
__device__ __forceinline__ void mask_body(const DevF &c, const TileGeom &g, MaskShared<FT> &sm, const MaskTile<FT> &sh,
                                          unsigned long long *bar, int blk, unsigned parity) {
    const WarpCell w = warp_cell<FT>(c, g, blk);
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const bool ok = !sh.overflow;
    if (ok) mbar_wait(bar, parity);
    if (w.nc == 0) return;
    // what this warp needs to know about its 27 (9) stencil cells, computed once by lane = stencil cell
    int ox, oy, of;
    stencil<FT>(min(lane, FT::NW - 1), ox, oy, of);
    int4 te = make_int4(0, 0, 0, 0);
    bool bowned = false, bflowc = false;
    if (lane < FT::NW) {
        const int q = stencil_cb<FT>(w, ox, oy, of);
        te.x = sh.cb[q];
        te.y = sh.cb[q + 1] - te.x;
        te.z = sh.gdelta[q / FT::CBW];
        const int ncx = w.cx + ox;
        bowned = ncx >= c.own0 && ncx < c.own1;                    // B's warp runs on this rank
        if (te.y > 0) bflowc = c.cellflow[(ncx * g.n1 + (w.cy + oy)) * g.nF + (w.f + of)] != 0;
        te.w = (ox + 1) | ((oy + 1) << 2) | ((of + 1) << 4) | (bowned ? 64 : 0) | (bflowc ? 128 : 0);
        sm.tab[wi][lane] = te;
    }
    // can this cell be represented?  (uniform per warp)
    const bool flagged = __any_sync(0xffffffffu, te.y > 32) || !ok || w.nc > 32;
    if (flagged) {                                                 // my neighbours wait for words only I can write: flag them too
        if (lane < FT::NW) {
            const int nx = w.cx + ox, ny = w.cy + oy, nf = w.f + of;
            if (nx >= 0 && nx < g.n0 && ny >= 0 && ny < g.n1 && nf >= 0 && nf < g.nF) c.cellflag[(nx * g.n1 + ny) * g.nF + nf] = 1;
        }
        if (lane == 0) atomicAdd(c.nflag, 1);
        return;
    }
    const bool mine = lane < w.nc;
    const int i = w.is + lane;
    const float *X = sh.X, *Y = sh.Y, *Z = sh.Z, *F = sh.F;
    const int own = sh.cb[stencil_cb<FT>(w, 0, 0, 0)] + (mine ? lane : 0);
    const float pix = X[own], piy = Y[own], piz = Z[own];
    const bool myflow = F[own] > 0.f;
    const unsigned flowA = __ballot_sync(0xffffffffu, mine && myflow);
    const bool has_wall = __any_sync(0xffffffffu, mine && !myflow);
    const bool near_flow = flowA != 0 || __any_sync(0xffffffffu, bflowc);
    // Cell pairs this warp evaluates: itself and the cells after it (their warps get the transposed words), plus
    // the cells before it that lie in a ghost column.  A pair of cells without any flow particle has empty words
    // (walls keep flow neighbours only).  Empty words are never stored: readers only follow the bits of nzw.
    unsigned todo = __ballot_sync(0xffffffffu, te.y > 0 && (lane >= FT::CENTRE || !bowned) && (flowA != 0 || bflowc));
    const u64 thr2 = pk2(c.r2thr, c.r2thr);
    // (mask words are addressed inside the cell blocks: mask_row)
    unsigned *mrow = mask_row(c.mask, w.is, FT::NW, mine ? lane : 0);
    unsigned nz = 0;
    __syncwarp();
    while (todo) {
        const int cc = __ffs(todo) - 1;
        todo &= todo - 1;
        const int4 t = sm.tab[wi][cc];
        const int a = t.x, nb = t.y;
        const bool same = cc == FT::CENTRE, upper = cc > FT::CENTRE, b_owned = (t.w & 64) != 0;
        const float bx = (float)((t.w & 3) - 1), by = (float)(((t.w >> 2) & 3) - 1), bf = (float)(((t.w >> 4) & 3) - 1);
        float sx, sy, sz;
        if (FT::d3) { sx = bx * c.gsT; sy = by * c.gsT; sz = bf * c.gsT; }
        else { sx = bx * c.gsT; sy = bf * c.gsT; sz = 0.f; }
        const bool bflow = lane < nb && F[a + lane] > 0.f;
        const unsigned flowB = __ballot_sync(0xffffffffu, bflow);
        const int a4 = a & ~3, lead = a - a4, nslot = lead + nb;   // aligned chunk origin; candidate k sits at slot lead + k
        u64 m64 = 0;
        if (upper || same) {                                       // lower side: d = (x_i - s) - x_j
            const float ex = pix - sx, ey = piy - sy, ez = piz - sz;
            const u64 Ex = pk2(ex, ex), Ey = pk2(ey, ey), Ez = pk2(ez, ez);
            for (int t0 = 0; t0 < nslot; t0 += 8) {
                const ulonglong2 x0 = *reinterpret_cast<const ulonglong2 *>(X + a4 + t0), x1 = *reinterpret_cast<const ulonglong2 *>(X + a4 + t0 + 4);
                const ulonglong2 y0 = *reinterpret_cast<const ulonglong2 *>(Y + a4 + t0), y1 = *reinterpret_cast<const ulonglong2 *>(Y + a4 + t0 + 4);
                const ulonglong2 z0 = *reinterpret_cast<const ulonglong2 *>(Z + a4 + t0), z1 = *reinterpret_cast<const ulonglong2 *>(Z + a4 + t0 + 4);
                unsigned cm = push2(0u, sub2(Ex, x1.y), sub2(Ey, y1.y), sub2(Ez, z1.y), thr2);
                cm = push2(cm, sub2(Ex, x1.x), sub2(Ey, y1.x), sub2(Ez, z1.x), thr2);
                cm = push2(cm, sub2(Ex, x0.y), sub2(Ey, y0.y), sub2(Ez, z0.y), thr2);
                cm = push2(cm, sub2(Ex, x0.x), sub2(Ey, y0.x), sub2(Ez, z0.x), thr2);
                m64 |= chunk_bits(cm, t0);
            }
        } else {                                                   // B precedes A and is a ghost column: d' = (x_j + s) - x_i
            const u64 Sx = pk2(sx, sx), Sy = pk2(sy, sy), Sz = pk2(sz, sz);
            const u64 Px = pk2(pix, pix), Py = pk2(piy, piy), Pz = pk2(piz, piz);
            for (int t0 = 0; t0 < nslot; t0 += 8) {
                const ulonglong2 x0 = *reinterpret_cast<const ulonglong2 *>(X + a4 + t0), x1 = *reinterpret_cast<const ulonglong2 *>(X + a4 + t0 + 4);
                const ulonglong2 y0 = *reinterpret_cast<const ulonglong2 *>(Y + a4 + t0), y1 = *reinterpret_cast<const ulonglong2 *>(Y + a4 + t0 + 4);
                const ulonglong2 z0 = *reinterpret_cast<const ulonglong2 *>(Z + a4 + t0), z1 = *reinterpret_cast<const ulonglong2 *>(Z + a4 + t0 + 4);
                unsigned cm = push2(0u, sub2(add2(x1.y, Sx), Px), sub2(add2(y1.y, Sy), Py), sub2(add2(z1.y, Sz), Pz), thr2);
                cm = push2(cm, sub2(add2(x1.x, Sx), Px), sub2(add2(y1.x, Sy), Py), sub2(add2(z1.x, Sz), Pz), thr2);
                cm = push2(cm, sub2(add2(x0.y, Sx), Px), sub2(add2(y0.y, Sy), Py), sub2(add2(z0.y, Sz), Pz), thr2);
                cm = push2(cm, sub2(add2(x0.x, Sx), Px), sub2(add2(y0.x, Sy), Py), sub2(add2(z0.x, Sz), Pz), thr2);
                m64 |= chunk_bits(cm, t0);
            }
        }
        unsigned m = (unsigned)(m64 >> lead);
        m &= nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
        if (same) m &= ~(1u << lane);                              // i != j
        if (!mine) m = 0;
        const unsigned mi = myflow ? m : (m & flowB);              // walls keep their flow neighbours only
        if (mi) {
            mrow[cc * w.nc] = __brev(mi);
            nz |= rbit(cc);
        }
        if (upper && b_owned) {                                    // the same pairs seen from B: transpose
            const unsigned tr = warp_transpose32(m, lane);
            const unsigned tj = bflow ? tr : (tr & flowA);
            if (lane < nb && tj) {
                const int j = a + lane + t.z;
                const int ccm = FT::NW - 1 - cc;
                mask_row(c.mask, a + t.z, FT::NW, lane)[ccm * nb] = __brev(tj);
                atomicOr(&c.nzw[j], rbit(ccm));
            }
        }
    }
    if (mine && nz) atomicOr(&c.nzw[i], nz);
    if (lane == 0) c.cellinfo[w.gcell] = (unsigned char)((flowA ? 1 : 0) | (has_wall ? 2 : 0) | (has_wall && near_flow ? 4 : 0));
}
template <class FT> __global__ void __launch_bounds__(FT::BT, 2) k_tile_mask(DevF c, TileGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MaskShared<FT> &sh = *reinterpret_cast<MaskShared<FT> *>(smem_raw);
    const int tid = threadIdx.x;
    const int items = c.wcount[3];
    const int *list = c.worklist[3];
    int *cursor = c.wcount + 4;
    if (tid == 0) {
        mbar_init(&sh.bar[0], 1);
        mbar_init(&sh.bar[1], 1);
        sh.item[0] = atomicAdd(cursor, 1);
    }
    __syncthreads();
    if (tid < 32 && sh.item[0] < items) mask_tile_issue<FT>(c, g, sh.t[0], &sh.bar[0], list[sh.item[0]]);
    __syncthreads();
    unsigned par0 = 0, par1 = 0;
    int b = 0;
    while (true) {
        const int it = sh.item[b];
        if (it >= items) break;
        if (tid < 32) {                                            // next item: fetch, spans, TMA into the other buffer
            if (tid == 0) sh.item[b ^ 1] = atomicAdd(cursor, 1);
            __syncwarp();
            const int nit = sh.item[b ^ 1];
            if (nit < items) mask_tile_issue<FT>(c, g, sh.t[b ^ 1], &sh.bar[b ^ 1], list[nit]);
        }
        mask_body<FT>(c, g, sh, sh.t[b], &sh.bar[b], list[it], b ? par1 : par0);
        if (b) par1 ^= 1u; else par0 ^= 1u;
        __syncthreads();                                           // tile b may be refilled; item / spans of the other buffer are visible
        b ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------ work lists
// footprint segments that hold work: mode 0: any own (and owned) cell is occupied; mode 1: any own cell has flow
// particles; mode 2: any own cell has wall particles next to flow particles (cellinfo, written by the mask kernel);
// mode 3 (the mask kernel's list): occupied AND a flow particle somewhere in the footprint or its one-cell halo --
// without one every mask word of the footprint is empty (walls keep flow neighbours only) and nzw stays zero.
template <class FT> __global__ void __launch_bounds__(256) k_tile_worklist(DevF c, TileGeom g, int nblk, int mode) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const int seg = b % g.nseg, t = b / g.nseg, b1 = t % g.nb1, b0 = t / g.nb1;
    bool any = false;
    for (int wx = 0; wx < FT::BX; wx++)
        for (int wy = 0; wy < FT::BY; wy++) {
            const int cx = b0 * FT::BX + wx, cy = b1 * FT::BY + wy;
            if (cx >= g.n0 || cy >= g.n1 || cx < c.own0 || cx >= c.own1) continue;
            const int base = (cx * g.n1 + cy) * g.nF, f0 = seg * ZB, f1 = min(f0 + ZB, g.nF);
            if (mode == 0 || mode == 3) any = any || c.cell_end[base + f1 - 1] > cell_start(c.cell_end, base + f0);
            else for (int f = f0; f < f1; f++) any = any || (c.cellinfo[base + f] & (mode == 1 ? 1 : 4)) != 0;
        }
    if (any && mode == 3) {
        bool flow = false;
        const int f0 = max(seg * ZB - 1, 0), f1 = min(seg * ZB + ZB + 1, g.nF);
        for (int cx = max(b0 * FT::BX - 1, 0); cx < min(b0 * FT::BX + FT::BX + 1, g.n0) && !flow; cx++)
            for (int cy = max(b1 * FT::BY - 1, 0); cy < min(b1 * FT::BY + FT::BY + 1, g.n1) && !flow; cy++) {
                const int base = (cx * g.n1 + cy) * g.nF;
                for (int f = f0; f < f1; f++) flow = flow || c.cellflow[base + f] != 0;
            }
        any = flow;
    }
    if (any) c.worklist[mode][atomicAdd(c.wcount + mode, 1)] = b;
}

// ------------------------------------------------------------------------------------------------ Shepard factor alone
// calc_CSPM_f (base:386-398) for every particle of unflagged cells, when sph_calc_kernel_corr is called on its own;
// inside sph_step the same sums are formed by the wall pass and the first fluid pass.
template <int KERNEL, class FT>
__device__ __forceinline__ bool shepard_body(const DevF &c, const TileGeom &g, TileShared<FT, 1> &sh, int blk, unsigned parity) {
    WarpCell w = warp_cell<FT>(c, g, blk);
    const int lane = threadIdx.x & 31;
    if (w.nc > 0 && c.cellflag[w.gcell]) w.nc = 0;
    if (!__syncthreads_or(w.nc > 0)) return false;
    if (!tile_setup<FT, 1>(c, g, sh, w, parity, c.ps4)) return true;
    build_ctab<FT, 1>(c, sh, w, lane);
    if (w.nc == 0) return true;
    const bool mine = lane < w.nc;
    const int i = w.is + (mine ? lane : 0);
    const F4 *A = sh.P[0];
    const F4 *ct = sh.ctab + (threadIdx.x >> 5) * FT::NW;
    const F4 pi = A[sh.cb[stencil_cb<FT>(w, 0, 0, 0)] + (mine ? lane : 0)];
    const KernConst kc = kern_const(c);
    const unsigned n = (unsigned)w.nc;
    const unsigned *mrow = mask_row(c.mask, w.is, FT::NW, mine ? lane : 0);
    Cursor k;
    cursor_init(k, mrow, n, mine ? c.nzw[i] : 0u);
    float ssum = 0.f;
    while (true) {
        cursor_jump(k, mrow, n, ct, pi);
        if (!__any_sync(0xffffffffu, k.m != 0)) break;
        int i0, i1, i2, i3;
        cursor_take2<FT::SENT>(k, i0, i1);
        const float e0x = k.ex, e0y = k.ey, e0z = k.ez;
        cursor_jump(k, mrow, n, ct, pi);
        cursor_take2<FT::SENT>(k, i2, i3);
        const F4 p0 = A[i0], p1 = A[i1], p2 = A[i2], p3 = A[i3];
        ssum = shep_add(ssum, p0.w, fastW<KERNEL>(kc, dist2(e0x - p0.x, e0y - p0.y, e0z - p0.z)));
        ssum = shep_add(ssum, p1.w, fastW<KERNEL>(kc, dist2(e0x - p1.x, e0y - p1.y, e0z - p1.z)));
        ssum = shep_add(ssum, p2.w, fastW<KERNEL>(kc, dist2(k.ex - p2.x, k.ey - p2.y, k.ez - p2.z)));
        ssum = shep_add(ssum, p3.w, fastW<KERNEL>(kc, dist2(k.ex - p3.x, k.ey - p3.y, k.ez - p3.z)));
    }
    if (mine) c.cspm_f[i] = (ssum != 0.f) ? 1.f / ssum : 1.f;
    return true;
}
template <int KERNEL, class FT> __global__ void __launch_bounds__(FT::BT) k_tile_shepard(DevF c, TileGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileShared<FT, 1> &sh = *reinterpret_cast<TileShared<FT, 1> *>(smem_raw);
    tile_init<FT, 1>(sh);
    TILE_PERSISTENT_LOOP(sh, c.worklist[0], c.wcount + 0, c.wcount + 5, (shepard_body<KERNEL, FT>(c, g, sh, blk, parity)))
}

// ------------------------------------------------------------------------------------------------ count + density sweep
// BASELINE config C5: the bare for_all_neighbors iteration (ps:259-269) with the density task (wc:30-31) -- per FLOW
// particle of an unflagged cell the neighbour count (popcount of its mask words: exact) and sum_j mass_j W_ij over the
// set bits, in ONE walk of the masks.  Payload: ONE float4 per neighbour (coordinates, mass), written by k_density_payload
// into the pk4 array right before -- a second shared-memory gather for the mass alone made the kernel 1.5x slower (the
// pass is bound by shared-memory gather wavefronts, profiles/r2_ncu_evidence.json).  Everything else (wall particles, whose
// masks hold flow neighbours only, and flagged cells) is left to the generic kernel.
__global__ void __launch_bounds__(256) k_density_payload(DevF c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    F4 p = c.xs4[i];
    p.w = c.v4[i].w;
    c.pk4[i] = p;
}
template <int KERNEL, class FT>
__device__ __forceinline__ bool density_body(const DevF &c, const TileGeom &g, TileShared<FT, 1> &sh1, int blk, unsigned parity,
                                             int *__restrict__ count_out, float *__restrict__ rho_out) {
    WarpCell w = warp_cell<FT>(c, g, blk);
    TileShared<FT, 1> &sh = sh1;
    const int lane = threadIdx.x & 31;
    if (w.nc > 0 && c.cellflag[w.gcell]) w.nc = 0;
    const int i = w.is + lane;
    const bool work = lane < w.nc && c.ps4[i].w > 0.f;
    const unsigned nz = work ? c.nzw[i] : 0u;
    if (!__syncthreads_or(work)) return false;
    cursor_prefetch(mask_row(c.mask, w.is, FT::NW, work ? lane : 0), (unsigned)w.nc, nz);
    if (!tile_setup<FT, 1>(c, g, sh, w, parity, c.pk4)) return true;
    build_ctab<FT, 1>(c, sh, w, lane);
    if (!__any_sync(0xffffffffu, work)) return true;
    const F4 *A = sh.P[0];
    const F4 *ct = sh.ctab + (threadIdx.x >> 5) * FT::NW;
    const F4 pi = A[sh.cb[stencil_cb<FT>(w, 0, 0, 0)] + (work ? lane : 0)];
    const KernConst kc = kern_const(c);
    const unsigned n = (unsigned)w.nc;
    const unsigned *mrow = mask_row(c.mask, w.is, FT::NW, work ? lane : 0);
    Cursor k;
    cursor_init(k, mrow, n, nz);
    float s0 = 0.f, s1 = 0.f;
    int cnt = 0;
    while (true) {
        const unsigned before = k.nz;
        cursor_jump(k, mrow, n, ct, pi);
        if (before != k.nz) cnt += __popc(k.m);                 // a new cell's word was taken: its bits are neighbours
        if (!__any_sync(0xffffffffu, k.m != 0)) break;
        int i0, i1, i2, i3;
        cursor_take2<FT::SENT>(k, i0, i1);
        const float e0x = k.ex, e0y = k.ey, e0z = k.ez;
        const unsigned before2 = k.nz;
        cursor_jump(k, mrow, n, ct, pi);
        if (before2 != k.nz) cnt += __popc(k.m);
        cursor_take2<FT::SENT>(k, i2, i3);
        const F4 p0 = A[i0], p1 = A[i1], p2 = A[i2], p3 = A[i3];                // .w = mass; sentinel: 0
        s0 = fmaf(p0.w, fastW<KERNEL>(kc, dist2(e0x - p0.x, e0y - p0.y, e0z - p0.z)), s0);
        s1 = fmaf(p1.w, fastW<KERNEL>(kc, dist2(e0x - p1.x, e0y - p1.y, e0z - p1.z)), s1);
        s0 = fmaf(p2.w, fastW<KERNEL>(kc, dist2(k.ex - p2.x, k.ey - p2.y, k.ez - p2.z)), s0);
        s1 = fmaf(p3.w, fastW<KERNEL>(kc, dist2(k.ex - p3.x, k.ey - p3.y, k.ez - p3.z)), s1);
    }
    if (work) { count_out[i] = cnt; rho_out[i] = s0 + s1; }
    return true;
}
template <int KERNEL, class FT>
__global__ void __launch_bounds__(FT::BT, 2) k_tile_density(DevF c, TileGeom g, int *__restrict__ count_out, float *__restrict__ rho_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileShared<FT, 1> &sh = *reinterpret_cast<TileShared<FT, 1> *>(smem_raw);
    tile_init<FT, 1>(sh);
    TILE_PERSISTENT_LOOP(sh, c.worklist[1], c.wcount + 1, c.wcount + 6, (density_body<KERNEL, FT>(c, g, sh, blk, parity, count_out, rho_out)))
}

// ------------------------------------------------------------------------------------------------ prep (pointwise)
// wc:87-88 EOS in float64 into the NEW pressure buffer; signed volume of the tile payload; fluid half of pk4; the
// constant result of dry wall particles (no flow neighbour: empty masks, they never reach the wall pass).
// ADV (inside sph_step only): the "LF" half-step update advect_LF_half (base:96-104, = k_advect kind 1) of the same
// particle first, so that the stage state is written once and not read back by a second kernel.
template <bool ADV> __global__ void __launch_bounds__(256) k_tile_prep(DevF c, int shep) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    const int t = c.type[i];
    F4 xs = c.xs4[i];
    F4 vt;
    double rt = 0.0;
    bool have = false;
    if (ADV && is_real(t)) {
        vt = c.vt4[i];
        const F4 dv = c.d_vel[i];
        rt = c.rho_t[i] + 0.5 * c.dt * (double)c.d_rho[i];
        c.rho_t[i] = rt;
        xs.w = (float)((double)c.v4[i].w / rt);
        c.xs4[i] = xs;
        const float hdt = (float)(0.5 * c.dt);
        vt.x += hdt * dv.x; vt.y += hdt * dv.y; vt.z += hdt * dv.z; vt.w = (float)rt;
        c.vt4[i] = vt;
        have = true;
    }
    F4 ps = xs;
    const bool fl = is_flow(t);
    ps.w = fl ? xs.w : -xs.w;
    c.ps4[i] = ps;
    if (is_fluid(t)) {
        if (!have) { rt = c.rho_t[i]; vt = c.vt4[i]; }
        double v = c.stiff * (pow(rt / c.rho0, c.gamma_) - 1.0);
        v = v > 0.0 ? v : 0.0;
        const float p = (float)v;
        c.pnew[i] = p;
        F4 pk = vt;
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
    if (is_rigid(t)) {                                              // static rigid: d_vel = 0, d_density = 0 (wc:125-126)
        c.d_rho[i] = 0.f;
        F4 z; z.x = z.y = z.z = z.w = 0.f;
        c.d_vel[i] = z;
    }
    if (!fl && c.nzw[i] == 0u && !c.cellflag[c.gid[i]]) {           // dry: v~ = 2 v, rho~ = rho0, p = 0, f = 1
        const F4 v = c.v4[i];
        F4 wt; wt.x = 2.f * v.x; wt.y = 2.f * v.y; wt.z = 2.f * v.z; wt.w = c.rho0T;
        c.vt4[i] = wt;
        c.rho_t[i] = c.rho0;
        c.pnew[i] = 0.f;
        if (shep) c.cspm_f[i] = 1.f;
        F4 pk = wt; pk.w = 0.f;
        c.pk4[i] = pk;
    }
}

// ------------------------------------------------------------------------------------------------ pass A: walls
// wc:90-103 for dummy-wall particles: v~ = 2v - f sum V v~ W, rho~ = rho0, p = max(f sum V (p_j + rho~_j g_y dy) W, 0)
// SHEP (the first wall pass after the masks were built): f = 1 / sum V W (calc_CSPM_f) is formed in the same visit and
// stored; later passes of the step reuse the stored f like the reference does (m_V changes with the stage density).
// Masks of wall particles hold flow neighbours only.
// Payloads per neighbour: ps4 (coords, signed volume), vt4 (v~, rho~), pw4 (EOS pressure, previous pressure).
template <int KERNEL>
__device__ __forceinline__ void wall_pair(const KernConst &kc, float ex, float ey, float ez, const F4 pj, const F4 vj, float pjv,
                                          float gy, float &vw, float &pterm) {
    const float dx = ex - pj.x, dy = ey - pj.y, dz = ez - pj.z;
    vw = __fmul_rn(fmaxf(pj.w, 0.f), fastW<KERNEL>(kc, dist2(dx, dy, dz)));   // flow neighbour: w = +V; sentinel: 0
    pterm = fmaf(vj.w * gy, dy, pjv);
}

// ------------------------------------------------------------------------------------------------ pass A: walls, gathered
// The wall pass without a tile (the first version staged 1x1x4-cell footprints like the fluid pass): one warp per wall cell that has flow particles in reach (a compacted CELL list),
// lane = particle, the mask bits walked in the same order with the same arithmetic, but the neighbour payloads are
// gathered from global memory through L1 (the lanes of a warp share most of their neighbours).  Only ~5 % of the
// particles take part in this pass and each has few (flow-only) neighbours, so staging 54 cells x 3 payloads per four
// cells left the SMs idle behind TMA latency and block barriers (ncu: 12 % issue slots used, 9.6 warp-cycles of
// barrier stall per issue).  Warps are independent here: no block barrier, 32 warps per SM.
template <bool D3> __global__ void __launch_bounds__(256) k_wall_cells(DevF c, int cell0, int cell1) {
    const int gcell = cell0 + blockIdx.x * blockDim.x + threadIdx.x;      // (a slab only looks at the cells of its own columns)
    if (gcell >= cell1) return;
    if (!(c.cellinfo[gcell] & 4) || c.cellflag[gcell]) return;
    const int nF = D3 ? c.gn[2] : c.gn[1], n1 = D3 ? c.gn[1] : 1;
    const int cx = gcell / (nF * n1);
    if (cx < c.own0 || cx >= c.own1) return;                        // ghost columns of a slab idle
    c.worklist[2][atomicAdd(c.wcount + 2, 1)] = gcell;
}
constexpr int WG_WARPS = 8;
template <int KERNEL, bool D3, bool SHEP, bool DYN> __global__ void __launch_bounds__(WG_WARPS * 32, 4) k_wall_gather(DevF c) {
    constexpr int NW = D3 ? 27 : 9, CENTRE = NW / 2;
    __shared__ F4 s_ct[WG_WARPS][NW];        // per warp and stencil cell: shift xyz, CT_BEFORE / CT_SAME
    __shared__ int s_start[WG_WARPS][NW];    // global index of the stencil cell's first particle
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int items = c.wcount[2];
    const int nF = D3 ? c.gn[2] : c.gn[1], n1 = D3 ? c.gn[1] : 1, n0 = c.gn[0];
    const KernConst kc = kern_const(c);
    const float gy = c.g[1];
    const bool fresh = c.wc_fresh != 0;
    // (the stride between a particle's mask words is its cell's particle count)
    const F4 *ct = s_ct[wi];
    const int *cs = s_start[wi];
    for (int it = blockIdx.x * WG_WARPS + wi;; it += gridDim.x * WG_WARPS) {
        if (DYN) {                                                 // wall cells differ a lot in cost: warps pull them from a cursor
            if (lane == 0) it = atomicAdd(c.wcount + 8, 1);
            it = __shfl_sync(0xffffffffu, it, 0);
        }
        if (it >= items) break;
        const int gcell = c.worklist[2][it];
        const int f = gcell % nF, t = gcell / nF, cy = t % n1, cx = t / n1;
        const int is = cell_start(c.cell_end, gcell), nc = c.cell_end[gcell] - is;
        __syncwarp();
        if (lane < NW) {
            int ox, oy, of;
            if (D3) { ox = lane / 9 - 1; oy = (lane / 3) % 3 - 1; of = lane % 3 - 1; }
            else { ox = lane / 3 - 1; oy = 0; of = lane % 3 - 1; }
            const int nx = cx + ox, ny = cy + oy, nf = f + of;
            int st = 0;
            if (nx >= 0 && nx < n0 && ny >= 0 && ny < n1 && nf >= 0 && nf < nF) st = cell_start(c.cell_end, (nx * n1 + ny) * nF + nf);
            F4 tt;
            if (D3) { tt.x = (float)ox * c.gsT; tt.y = (float)oy * c.gsT; tt.z = (float)of * c.gsT; }
            else { tt.x = (float)ox * c.gsT; tt.y = (float)of * c.gsT; tt.z = 0.f; }
            unsigned v = 0;
            if (lane < CENTRE) v |= CT_BEFORE;                     // stencil order == ascending cell id
            if (lane == CENTRE) v |= CT_SAME;
            tt.w = __uint_as_float(v);
            s_ct[wi][lane] = tt;
            s_start[wi][lane] = st;
        }
        __syncwarp();
        const int i = is + lane;
        bool work = false;
        unsigned nz = 0;
        F4 pi; pi.x = pi.y = pi.z = pi.w = 0.f;
        if (lane < nc) {
            pi = c.ps4[i];
            if (pi.w < 0.f) { nz = c.nzw[i]; work = nz != 0; }      // dry walls: k_tile_prep
        }
        if (!work) nz = 0;
        if (!__any_sync(0xffffffffu, work)) continue;
        const int safe = work ? i : is;                             // what an empty slot loads (its volume is zeroed)
        const unsigned n = (unsigned)nc;
        const unsigned *mrow = mask_row(c.mask, is, NW, work ? lane : 0);
        float Sv0 = 0.f, Sv1 = 0.f, Sv2 = 0.f, Sp = 0.f, Sw = 0.f;
        // cursor state (see struct Cursor); the base of the current cell is a GLOBAL particle index here
        unsigned m = 0, mnext = 0, flags = 0;
        int a = 0;
        float ex = 0.f, ey = 0.f, ez = 0.f;
        if (nz) mnext = __ldg(mrow + (unsigned)__clz(nz) * n);
        auto jump = [&]() {
            const bool jmp = m == 0 && nz != 0;
            if (jmp) {
                const int cc = __clz(nz);
                nz &= ~rbit(cc);
                m = mnext;
                const F4 tt = ct[cc];
                flags = __float_as_uint(tt.w);
                a = cs[cc];
                ex = pi.x - tt.x; ey = pi.y - tt.y; ez = pi.z - tt.z;
            }
            const unsigned go = jmp ? nz : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p ld.global.nc.u32 %0, [%1];\n}"
                         : "+r"(mnext) : "l"(mrow + (unsigned)(__clz(go) & 31) * n), "r"(go));
        };
        auto take2 = [&](int &j0, int &j1) {
            const unsigned m0 = m;
            const int t0 = __clz(m0);
            const unsigned m1 = m0 & ~rbit(t0 & 31);
            const int t1 = __clz(m1);
            m = m1 & ~rbit(t1 & 31);
            j0 = m0 ? a + t0 : -1;
            j1 = m1 ? a + t1 : -1;
        };
        while (true) {
            jump();
            if (!__any_sync(0xffffffffu, m != 0)) break;
            int j0, j1, j2, j3;
            take2(j0, j1);
            const float e0x = ex, e0y = ey, e0z = ez;
            const bool b0 = fresh || (flags & CT_BEFORE), s0 = (flags & CT_SAME) != 0;
            jump();
            take2(j2, j3);
            const bool b1 = fresh || (flags & CT_BEFORE), s1 = (flags & CT_SAME) != 0;
            const int g0 = j0 >= 0 ? j0 : safe, g1 = j1 >= 0 ? j1 : safe, g2 = j2 >= 0 ? j2 : safe, g3 = j3 >= 0 ? j3 : safe;
            F4 p0 = c.ps4[g0], p1 = c.ps4[g1], p2 = c.ps4[g2], p3 = c.ps4[g3];
            const F4 u0 = c.vt4[g0], u1 = c.vt4[g1], u2 = c.vt4[g2], u3 = c.vt4[g3];
            const F4 w0 = c.pw4[g0], w1 = c.pw4[g1], w2 = c.pw4[g2], w3 = c.pw4[g3];
            if (j0 < 0) p0.w = 0.f;
            if (j1 < 0) p1.w = 0.f;
            if (j2 < 0) p2.w = 0.f;
            if (j3 < 0) p3.w = 0.f;
            const float q0 = (b0 || (s0 && g0 < i)) ? w0.x : w0.y;
            const float q1 = (b0 || (s0 && g1 < i)) ? w1.x : w1.y;
            const float q2 = (b1 || (s1 && g2 < i)) ? w2.x : w2.y;
            const float q3 = (b1 || (s1 && g3 < i)) ? w3.x : w3.y;
            float vw0, pt0, vw1, pt1, vw2, pt2, vw3, pt3;
            wall_pair<KERNEL>(kc, e0x, e0y, e0z, p0, u0, q0, gy, vw0, pt0);
            wall_pair<KERNEL>(kc, e0x, e0y, e0z, p1, u1, q1, gy, vw1, pt1);
            wall_pair<KERNEL>(kc, ex, ey, ez, p2, u2, q2, gy, vw2, pt2);
            wall_pair<KERNEL>(kc, ex, ey, ez, p3, u3, q3, gy, vw3, pt3);
            Sw = __fadd_rn(Sw, vw0); Sv0 = fmaf(vw0, u0.x, Sv0); Sv1 = fmaf(vw0, u0.y, Sv1); Sv2 = fmaf(vw0, u0.z, Sv2); Sp = fmaf(vw0, pt0, Sp);
            Sw = __fadd_rn(Sw, vw1); Sv0 = fmaf(vw1, u1.x, Sv0); Sv1 = fmaf(vw1, u1.y, Sv1); Sv2 = fmaf(vw1, u1.z, Sv2); Sp = fmaf(vw1, pt1, Sp);
            Sw = __fadd_rn(Sw, vw2); Sv0 = fmaf(vw2, u2.x, Sv0); Sv1 = fmaf(vw2, u2.y, Sv1); Sv2 = fmaf(vw2, u2.z, Sv2); Sp = fmaf(vw2, pt2, Sp);
            Sw = __fadd_rn(Sw, vw3); Sv0 = fmaf(vw3, u3.x, Sv0); Sv1 = fmaf(vw3, u3.y, Sv1); Sv2 = fmaf(vw3, u3.z, Sv2); Sp = fmaf(vw3, pt3, Sp);
        }
        // outputs are written after every lane of the warp has finished reading (vt4 of a wall particle is never a
        // flow neighbour's payload, and pnew / pk4 / cspm_f are not read by this pass)
        if (!work) continue;
        float fi;
        if (SHEP) { fi = (Sw != 0.f) ? 1.f / Sw : 1.f; c.cspm_f[i] = fi; }
        else fi = c.cspm_f[i];
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
}

// ------------------------------------------------------------------------------------------------ pass B: fluid
// wc:108-126 for fluid particles: continuity + viscosity + pressure in one visit of the set bits.
//   d_rho_i = rho~_i sum_j V_j (v~_i - v~_j) . gradW_ij
//   d_v_i   = g + sum_j [ 2(dim+2) nu V_j min(v_ij . x_ij, 0) / (r^2 + 0.01 h^2) {1 | rho0 / rho~_i} - rho0 V_j (p_i/rho~_i^2 + p_j/rho~_j^2) ] gradW_ij
// SHEP: also the Shepard sum over flow neighbours (calc_CSPM_f, base:386-398) -- the first fluid pass of a step.  The
// volumes read there belong to the FIRST one_step, whose m_V is still the step-start value calc_CSPM_f would see.
struct FluidI { float vx, vy, vz, npr, visc_f, visc_w, h2, nrho0; };      // npr = -rho0 p_i / rho~_i^2
template <int KERNEL, bool SHEP>
__device__ __forceinline__ void fluid_pair(const KernConst &kc, const FluidI &I, float ex, float ey, float ez, const F4 pj,
                                           const F4 qj, float &dd, float &a0, float &a1, float &a2, float &ssum) {
    const float dx = ex - pj.x, dy = ey - pj.y, dz = ez - pj.z;
    const float r2 = dist2(dx, dy, dz);
    const float r2c = fmaxf(r2, kc.eps2), rinv = rsqrt_fast(r2c), r = r2c * rinv;
    float s;
    if (KERNEL == 1) {
        const float q1 = fmaxf(fmaf(-0.5f * kc.hinv, r, 1.f), 0.f), q2 = q1 * q1;
        s = (kc.c_grad * q1) * q2;
        if (SHEP) ssum = shep_add(ssum, pj.w, (kc.knorm * q2) * (q2 * fmaf(2.f * kc.hinv, r, 1.f)));   // == fastW<1>
    } else {
        const float q = r * kc.hinv, t = fmaxf(2.f - q, 0.f);
        s = q <= 1.f ? kc.c_grad * fmaf(1.5f, q, -2.f) : -0.5f * kc.c_grad * t * t * (rinv / kc.hinv);
        if (SHEP) ssum = shep_add(ssum, pj.w, fastW<0>(kc, r2));
    }
    const float ux = I.vx - qj.x, uy = I.vy - qj.y, uz = I.vz - qj.z;
    const float vx = fmaf(uz, dz, fmaf(uy, dy, ux * dx));          // v_ij . x_ij
    const float Vs = fabsf(pj.w) * s;                              // sentinel: 0
    dd = fmaf(Vs, vx, dd);
    const float visc = ((pj.w < 0.f ? I.visc_w : I.visc_f) * fminf(vx, 0.f)) * rcp_fast(r2 + I.h2);
    const float cf = Vs * (fmaf(I.nrho0, qj.w, I.npr) + visc);
    a0 = fmaf(cf, dx, a0); a1 = fmaf(cf, dy, a1); a2 = fmaf(cf, dz, a2);
}

// LIST (neighbour round lists): 0 none; 1 this pass walks the mask bits and RECORDS the four tile indices of
// every round (the first fluid pass after the masks were built); 2 this pass REPLAYS the recorded rounds -- same
// neighbours, same order, same arithmetic, without the bit cursor (cells whose list overflowed walk the bits again).
template <int KERNEL, class FT, bool SHEP, int LIST>
__device__ __forceinline__ void fluid_work(const DevF &c, const TileGeom &g, TileShared<FT, 2> &sh, const WarpCell &w, int lane, int i, bool work, unsigned nz) {
    const int rounds = (LIST == 2 && w.nc > 0) ? c.lrounds[w.gcell] : -1;     // warp-uniform
    if (rounds < 0) cursor_prefetch(mask_row(c.mask, w.is, FT::NW, work ? lane : 0), (unsigned)w.nc, nz);
    build_ctab<FT, 2>(c, sh, w, lane);
    const F4 *A = sh.P[0], *B = sh.P[1];
    const F4 *ct = sh.ctab + (threadIdx.x >> 5) * FT::NW;
    const int ci = sh.cb[stencil_cb<FT>(w, 0, 0, 0)] + (work ? lane : 0);
    const F4 pi = A[ci], qi = B[ci];                               // qi = v~_i, p_i / rho~_i^2
    const float rhoi = work ? c.vt4[i].w : 1.f;
    const KernConst kc = kern_const(c);
    FluidI I;
    I.vx = qi.x; I.vy = qi.y; I.vz = qi.z; I.nrho0 = -c.rho0T; I.npr = I.nrho0 * qi.w;
    I.visc_f = c.visc_coef; I.visc_w = c.visc_coef * c.rho0T / rhoi;   // wc:41-44
    I.h2 = c.h2_001;
    float dd = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, ssum = 0.f;
    const unsigned n = (unsigned)w.nc;
    constexpr unsigned SW = ((unsigned)FT::SENT << 12) | (unsigned)FT::SENT;   // a half round of two empty slots
    if (LIST == 2 && rounds >= 0) {
        const uint2 *lrow = c.nlist + ((size_t)w.is * LIST_ROUNDS + (work ? lane : 0));   // cell block: [round][particle of the cell]
        uint2 cur = make_uint2(SW, SW);
        if (work && rounds > 0) cur = __ldg(lrow);
#pragma unroll 1
        for (int r = 0; r < rounds; r++) {
            lrow += w.nc;
            uint2 nxt = make_uint2(SW, SW);
            if (work && r + 1 < rounds) nxt = __ldg(lrow);          // one round ahead
            const F4 t0 = ct[cur.x >> 24], t1 = ct[cur.y >> 24];
            const int i0 = (cur.x >> 12) & 0xfff, i1 = cur.x & 0xfff, i2 = (cur.y >> 12) & 0xfff, i3 = cur.y & 0xfff;
            const F4 p0 = A[i0], q0 = B[i0], p1 = A[i1], q1 = B[i1], p2 = A[i2], q2 = B[i2], p3 = A[i3], q3 = B[i3];
            const float e0x = pi.x - t0.x, e0y = pi.y - t0.y, e0z = pi.z - t0.z;
            const float e1x = pi.x - t1.x, e1y = pi.y - t1.y, e1z = pi.z - t1.z;
            fluid_pair<KERNEL, SHEP>(kc, I, e0x, e0y, e0z, p0, q0, dd, a0, a1, a2, ssum);
            fluid_pair<KERNEL, SHEP>(kc, I, e0x, e0y, e0z, p1, q1, dd, a0, a1, a2, ssum);
            fluid_pair<KERNEL, SHEP>(kc, I, e1x, e1y, e1z, p2, q2, dd, a0, a1, a2, ssum);
            fluid_pair<KERNEL, SHEP>(kc, I, e1x, e1y, e1z, p3, q3, dd, a0, a1, a2, ssum);
            cur = nxt;
        }
    } else {
        const unsigned *mrow = mask_row(c.mask, w.is, FT::NW, work ? lane : 0);
        // record: 32-bit offset of this lane's entry of the current round inside nlist, and where the cell's block ends
        unsigned loff = (unsigned)w.is * LIST_ROUNDS + (work ? lane : 0);
        const unsigned lend = ((unsigned)w.is + (unsigned)w.nc) * LIST_ROUNDS;
        int r = 0;
        Cursor k;
        cursor_init(k, mrow, n, nz);
        while (true) {
            cursor_jump(k, mrow, n, ct, pi);
            if (!__any_sync(0xffffffffu, k.m != 0)) break;              // warp-uniform round: lanes reconverge here
            int i0, i1, i2, i3;
            cursor_take2<FT::SENT>(k, i0, i1);
            const float e0x = k.ex, e0y = k.ey, e0z = k.ez;
            const unsigned w0 = (k.flags & CT_CC) | ((unsigned)i0 << 12) | (unsigned)i1;
            cursor_jump(k, mrow, n, ct, pi);
            cursor_take2<FT::SENT>(k, i2, i3);
            if (LIST == 1) {
                const unsigned w1 = (k.flags & CT_CC) | ((unsigned)i2 << 12) | (unsigned)i3;
                if (work && loff < lend) c.nlist[loff] = make_uint2(w0, w1);
                loff += (unsigned)w.nc;
                r++;
            }
            const F4 p0 = A[i0], q0 = B[i0], p1 = A[i1], q1 = B[i1], p2 = A[i2], q2 = B[i2], p3 = A[i3], q3 = B[i3];
            fluid_pair<KERNEL, SHEP>(kc, I, e0x, e0y, e0z, p0, q0, dd, a0, a1, a2, ssum);
            fluid_pair<KERNEL, SHEP>(kc, I, e0x, e0y, e0z, p1, q1, dd, a0, a1, a2, ssum);
            fluid_pair<KERNEL, SHEP>(kc, I, k.ex, k.ey, k.ez, p2, q2, dd, a0, a1, a2, ssum);
            fluid_pair<KERNEL, SHEP>(kc, I, k.ex, k.ey, k.ez, p3, q3, dd, a0, a1, a2, ssum);
        }
        if (LIST == 1 && lane == 0) c.lrounds[w.gcell] = r <= LIST_ROUNDS ? r : -1;
    }
    if (!work) return;
    c.d_rho[i] = dd * rhoi;
    F4 dv; dv.x = a0 + c.g[0]; dv.y = a1 + c.g[1]; dv.z = a2 + c.g[2]; dv.w = 0.f;
    c.d_vel[i] = dv;
    if (SHEP) c.cspm_f[i] = (ssum != 0.f) ? 1.f / ssum : 1.f;
}
// One work item of the pipelined loop: cell from the active boundaries, nzw in flight during the wait for the copies,
// the flow flag from the tile itself; the first warp to finish stages the next item (tile_finish).
template <int KERNEL, class FT, bool SHEP, int LIST>
__device__ __forceinline__ void fluid_body(const DevF &c, const TileGeom &g, TileShared<FT, 2> &sh, int blk, unsigned parity,
                                           unsigned k, const int *list, int items, int *cursor, long long *tt = nullptr) {
    const int lane = threadIdx.x & 31;
    const WarpCell w = tile_begin<FT, 2>(c, g, sh, blk);
    const int i = w.is + lane;
    unsigned nz = lane < w.nc ? c.nzw[i] : 0u;
    if (tile_wait<FT, 2>(sh, parity)) {
#ifdef TILE_TIMING
        if (tt) *tt = clock64();
#endif
        const int ci = sh.cb[stencil_cb<FT>(w, 0, 0, 0)] + (lane < w.nc ? lane : 0);
        const bool work = lane < w.nc && sh.P[0][ci].w > 0.f;          // flow particle (fluid: the only flow type of WCSPH)
        if (!work) nz = 0u;
        if (__any_sync(0xffffffffu, work)) fluid_work<KERNEL, FT, SHEP, LIST>(c, g, sh, w, lane, i, work, nz);
    }
    tile_finish<FT, 2>(c, g, sh, list, items, cursor, k);
}
// The same work item in the classic loop (cursor -> barrier -> look-ups -> spans -> TMA -> barrier -> wait): a block
// holds no item beyond the one it works on, which matters when it sees only a handful of them (slabs of a multi-GPU run).
template <int KERNEL, class FT, bool SHEP, int LIST>
__device__ __forceinline__ bool fluid_body_classic(const DevF &c, const TileGeom &g, TileShared<FT, 2> &sh, int blk, unsigned parity) {
    WarpCell w = warp_cell<FT>(c, g, blk);
    const int lane = threadIdx.x & 31;
    if (w.nc > 0 && c.cellflag[w.gcell]) w.nc = 0;
    const int i = w.is + lane;
    const bool work = lane < w.nc && c.ps4[i].w > 0.f;
    const unsigned nz = work ? c.nzw[i] : 0u;
    if (!__syncthreads_or(work)) return false;
    if (!tile_setup<FT, 2>(c, g, sh, w, parity, c.ps4, c.pk4)) return true;
    if (__any_sync(0xffffffffu, work)) fluid_work<KERNEL, FT, SHEP, LIST>(c, g, sh, w, lane, i, work, nz);
    return true;
}
template <int KERNEL, class FT, bool SHEP, int LIST, bool PIPE> __global__ void __launch_bounds__(FT::BT, 2) k_tile_fluid(DevF c, TileGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileShared<FT, 2> &sh = *reinterpret_cast<TileShared<FT, 2> *>(smem_raw);
    tile_init<FT, 2>(sh);
    if (PIPE) {
        TILE_PIPELINED_LOOP(FT, 2, sh, c.worklist[1], c.wcount + 1, c.wcount + 7, c.ps4, c.pk4, (fluid_body<KERNEL, FT, SHEP, LIST>(c, g, sh, blk, parity, k_, list_, items_, cursor_ TT_ARG)))
    } else {
        TILE_PERSISTENT_LOOP(sh, c.worklist[1], c.wcount + 1, c.wcount + 7, (fluid_body_classic<KERNEL, FT, SHEP, LIST>(c, g, sh, blk, parity)))
    }
}

// ------------------------------------------------------------------------------------------------ mask-based count
// number of neighbours per FLOW particle of unflagged cells from the masks (parity probe of k_tile_mask); -1 elsewhere
__global__ void __launch_bounds__(256) k_mask_count(DevF c, int nw, int *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    int cnt = -1;
    if (c.ps4[i].w > 0.f && !c.cellflag[c.gid[i]]) {
        cnt = 0;
        const unsigned nz = c.nzw[i];                  // empty words are never stored
        const int g = c.gid[i], is = cell_start(c.cell_end, g), nc = c.cell_end[g] - is;
        const unsigned *mrow = mask_row(c.mask, is, nw, i - is);
        for (int cc = 0; cc < nw; cc++) if (nz & rbit(cc)) cnt += __popc(mrow[cc * nc]);
    }
    out[i] = cnt;
}

// ------------------------------------------------------------------------------------------------ host side
typedef Foot<true, 2, 2> F3M;      // 3D: masks / Shepard / fluid pass -- 16 warps, 16 runs
typedef Foot<false, 4, 1> F2M;     // 2D: 16 warps, 6 runs

template <class FT, int NP> static size_t smem_of() { return sizeof(TileShared<FT, NP>); }

template <typename K> static int set_smem(SphCtx *c, K kern, size_t bytes) {
    SPH_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}
template <int KERNEL, class FT> static int set_fluid_attrs(SphCtx *c) {
    int r = 0;
    if (!r) r = set_smem(c, k_tile_fluid<KERNEL, FT, false, 0, true>, smem_of<FT, 2>());
    if (!r) r = set_smem(c, k_tile_fluid<KERNEL, FT, true, 0, true>, smem_of<FT, 2>());
    if (!r) r = set_smem(c, k_tile_fluid<KERNEL, FT, false, 1, true>, smem_of<FT, 2>());
    if (!r) r = set_smem(c, k_tile_fluid<KERNEL, FT, true, 1, true>, smem_of<FT, 2>());
    if (!r) r = set_smem(c, k_tile_fluid<KERNEL, FT, false, 2, true>, smem_of<FT, 2>());
    if (!r) r = set_smem(c, k_tile_fluid<KERNEL, FT, false, 0, false>, smem_of<FT, 2>());
    if (!r) r = set_smem(c, k_tile_fluid<KERNEL, FT, true, 0, false>, smem_of<FT, 2>());
    if (!r) r = set_smem(c, k_tile_fluid<KERNEL, FT, false, 1, false>, smem_of<FT, 2>());
    if (!r) r = set_smem(c, k_tile_fluid<KERNEL, FT, true, 1, false>, smem_of<FT, 2>());
    if (!r) r = set_smem(c, k_tile_fluid<KERNEL, FT, false, 2, false>, smem_of<FT, 2>());
    if (!r) r = set_smem(c, k_tile_density<KERNEL, FT>, smem_of<FT, 1>());
    return r;
}
template <int KERNEL> static int set_attrs(SphCtx *c) {
    int r = 0;
    if (!r) r = set_smem(c, k_tile_shepard<KERNEL, F3M>, smem_of<F3M, 1>());
    if (!r) r = set_smem(c, k_tile_shepard<KERNEL, F2M>, smem_of<F2M, 1>());
    if (!r) r = set_fluid_attrs<KERNEL, F3M>(c);
    if (!r) r = set_fluid_attrs<KERNEL, F2M>(c);
    return r;
}
static int ensure_attrs(SphCtx *c) {
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || done[dev]) return 0;
    int r = set_attrs<0>(c);
    if (!r) r = set_attrs<1>(c);
    if (!r) r = set_smem(c, k_tile_mask<F3M>, sizeof(MaskShared<F3M>));
    if (!r) r = set_smem(c, k_tile_mask<F2M>, sizeof(MaskShared<F2M>));
    done[dev] = r == 0;
    return r;
}
template <class FT> static int nblocks(const TileGeom &g) { return g.nb0 * g.nb1 * g.nseg; }
// persistent grid: the blocks that are resident at once (occupancy x SMs), never more than there are segments; the
// work items are handed out dynamically (atomic cursor), so a bigger grid would only add a tail
template <typename K> static int pgrid(K kern, int threads, size_t smem, int nb) {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return nb < per_sm * sms ? nb : per_sm * sms;
}
// launches a persistent tile kernel over its work list; `cursor` = index of the dynamic cursor in wcount
#define TILE_LAUNCH_SMEM(KERN, FT, SMEM, CURSOR)                                                                  \
    do {                                                                                                          \
        const TileGeom g_ = make_geom<FT>(d.gn);                                                                  \
        cudaMemsetAsync(d.wcount + (CURSOR), 0, 4, c->stream);                                                     \
        KERN<<<pgrid(KERN, FT::BT, (SMEM), nblocks<FT>(g_)), FT::BT, (SMEM), c->stream>>>(d, g_);                  \
    } while (0)
#define TILE_LAUNCH(KERN, FT, NP, CURSOR) TILE_LAUNCH_SMEM(KERN, FT, (smem_of<FT, NP>()), CURSOR)
template <class FT> static int build_worklist(SphCtx *c, const DevF &d, int mode) {
    const TileGeom g = make_geom<FT>(d.gn);
    const int nb = nblocks<FT>(g);
    SPH_CHECK(c, cudaMemsetAsync(d.wcount + mode, 0, 4, c->stream));
    SPH_PROF(c, K_OTHER);
    k_tile_worklist<FT><<<blocks_for(nb, 256), 256, 0, c->stream>>>(d, g, nb, mode);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// masks (+ the Shepard factor of every particle when `shepard`: the stand-alone sph_calc_kernel_corr)
int tile_mask(SphCtx *c, bool shepard) {
    DevF d = make_dev<float>(c);
    int r = ensure_attrs(c);
    if (r) return r;
    const bool d3 = c->p.dim == 3;
    SPH_CHECK(c, cudaMemsetAsync(d.cellflag, 0, (size_t)c->C, c->stream));
    SPH_CHECK(c, cudaMemsetAsync(d.nflag, 0, 4, c->stream));
    SPH_CHECK(c, cudaMemsetAsync(d.nzw, 0, (size_t)c->n * 4, c->stream));
    SPH_CHECK(c, cudaMemsetAsync(d.cellinfo, 0, (size_t)c->C, c->stream));
    if ((r = d3 ? build_worklist<F3M>(c, d, 3) : build_worklist<F2M>(c, d, 3))) return r;
    SPH_PROF(c, K_TILE_MASK);
    if (d3) TILE_LAUNCH_SMEM(k_tile_mask<F3M>, F3M, sizeof(MaskShared<F3M>), 4);
    else TILE_LAUNCH_SMEM(k_tile_mask<F2M>, F2M, sizeof(MaskShared<F2M>), 4);
    SPH_LAUNCH_CHECK(c);
    if ((r = d3 ? build_worklist<F3M>(c, d, 1) : build_worklist<F2M>(c, d, 1))) return r;
    SPH_CHECK(c, cudaMemsetAsync(d.wcount + 2, 0, 4, c->stream));           // wall CELLS with flow particles in reach
    SPH_PROF(c, K_OTHER);
    const int nyz_ = c->p.gn[1] * (d3 ? c->p.gn[2] : 1), wc0 = c->own0 * nyz_, wc1 = c->own1 * nyz_;
    if (d3) k_wall_cells<true><<<blocks_for(wc1 - wc0, 256), 256, 0, c->stream>>>(d, wc0, wc1);
    else k_wall_cells<false><<<blocks_for(wc1 - wc0, 256), 256, 0, c->stream>>>(d, wc0, wc1);
    SPH_LAUNCH_CHECK(c);
    c->shep_pending = c->shep_wall_pending = !shepard;
    c->list_valid = false;
    c->masks_valid = true;
    if (!shepard) return 0;
    if ((r = d3 ? build_worklist<F3M>(c, d, 0) : build_worklist<F2M>(c, d, 0))) return r;
    SPH_PROF(c, K_CSPM_F);
    if (d3) {
        if (c->p.kernel == 0) TILE_LAUNCH((k_tile_shepard<0, F3M>), F3M, 1, 5); else TILE_LAUNCH((k_tile_shepard<1, F3M>), F3M, 1, 5);
    } else {
        if (c->p.kernel == 0) TILE_LAUNCH((k_tile_shepard<0, F2M>), F2M, 1, 5); else TILE_LAUNCH((k_tile_shepard<1, F2M>), F2M, 1, 5);
    }
    SPH_LAUNCH_CHECK(c);
    return 0;
}

static int tune_env(const char *name, int dflt) {                      // A/B switches of the tile path (tools/gpu_ab.sh)
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}
template <int KERNEL, bool D3> static void launch_wall_gather(SphCtx *c, const DevF &d, bool shep) {
    const int grid = 148 * 4;                                       // 4 blocks of 8 independent warps per SM (64 registers: all resident)
    static const int dyn_env = tune_env("TISPHI_WALL_DYN", -1);     // -1: cursor on one GPU, static stride on slabs
    const bool dyn = dyn_env < 0 ? c->slab == nullptr : dyn_env != 0;
    if (dyn) {
        cudaMemsetAsync(d.wcount + 8, 0, 4, c->stream);             // the cursor the warps pull wall cells from
        if (shep) k_wall_gather<KERNEL, D3, true, true><<<grid, WG_WARPS * 32, 0, c->stream>>>(d);
        else k_wall_gather<KERNEL, D3, false, true><<<grid, WG_WARPS * 32, 0, c->stream>>>(d);
    } else {
        if (shep) k_wall_gather<KERNEL, D3, true, false><<<grid, WG_WARPS * 32, 0, c->stream>>>(d);
        else k_wall_gather<KERNEL, D3, false, false><<<grid, WG_WARPS * 32, 0, c->stream>>>(d);
    }
}
// WCSPH one_step (wc:82-126) on the tile path; flagged cells are completed by the generic kernels (flagged_only).
static int need_masks(SphCtx *c) {
    if (c->masks_valid) return 0;
    snprintf(c->err, sizeof(c->err), "the neighbour masks are stale: call sph_calc_kernel_corr after sph_grid_build / an upload and before sph_one_step");
    return -3;
}
int tile_wc_prep_and_wall(SphCtx *c) {
    if (need_masks(c)) return -3;
    DevF d = make_dev<float>(c);
    const int n = (int)c->n;
    const bool shep = c->shep_wall_pending;
    c->shep_wall_pending = false;
    const bool adv = c->fuse_half;                                  // set by sph_step: the half-step update rides along
    c->fuse_half = false;
    SPH_PROF(c, K_WC_EOS);
    if (adv) k_tile_prep<true><<<blocks_for(n, 256), 256, 0, c->stream>>>(d, shep ? 1 : 0);
    else k_tile_prep<false><<<blocks_for(n, 256), 256, 0, c->stream>>>(d, shep ? 1 : 0);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_TILE_WALL);
    if (c->p.dim == 3) {
        if (c->p.kernel == 0) launch_wall_gather<0, true>(c, d, shep); else launch_wall_gather<1, true>(c, d, shep);
    } else {
        if (c->p.kernel == 0) launch_wall_gather<0, false>(c, d, shep); else launch_wall_gather<1, false>(c, d, shep);
    }
    SPH_LAUNCH_CHECK(c);
    return 0;
}
// list: 0 no lists, 1 record, 2 replay (never together with the Shepard sums: those belong to the first pass)
template <int KERNEL, class FT, bool PIPE> static void launch_fluid_p(SphCtx *c, const DevF &d, bool shep, int list) {
    if (list == 2) TILE_LAUNCH((k_tile_fluid<KERNEL, FT, false, 2, PIPE>), FT, 2, 7);
    else if (list == 1) {
        if (shep) TILE_LAUNCH((k_tile_fluid<KERNEL, FT, true, 1, PIPE>), FT, 2, 7);
        else TILE_LAUNCH((k_tile_fluid<KERNEL, FT, false, 1, PIPE>), FT, 2, 7);
    } else {
        if (shep) TILE_LAUNCH((k_tile_fluid<KERNEL, FT, true, 0, PIPE>), FT, 2, 7);
        else TILE_LAUNCH((k_tile_fluid<KERNEL, FT, false, 0, PIPE>), FT, 2, 7);
    }
}
template <int KERNEL, class FT> static void launch_fluid(SphCtx *c, const DevF &d, bool shep, int list) {
    static const int pipe_env = tune_env("TISPHI_FLUID_PIPE", -1);     // -1: pipelined loop on one GPU, classic loop on slabs
    const bool pipe = pipe_env < 0 ? c->slab == nullptr : pipe_env != 0;
    if (pipe) launch_fluid_p<KERNEL, FT, true>(c, d, shep, list); else launch_fluid_p<KERNEL, FT, false>(c, d, shep, list);
}
int tile_wc_fluid(SphCtx *c) {
    if (need_masks(c)) return -3;
    DevF d = make_dev<float>(c);
    const bool shep = c->shep_pending;
    c->shep_pending = false;
    const int list = !c->use_list ? 0 : ((c->list_valid && !shep) ? 2 : 1);
    c->list_valid = c->use_list;
    SPH_PROF(c, K_TILE_FLUID);
    if (c->p.dim == 3) {
        if (c->p.kernel == 0) launch_fluid<0, F3M>(c, d, shep, list); else launch_fluid<1, F3M>(c, d, shep, list);
    } else {
        if (c->p.kernel == 0) launch_fluid<0, F2M>(c, d, shep, list); else launch_fluid<1, F2M>(c, d, shep, list);
    }
    SPH_LAUNCH_CHECK(c);
    return 0;
}
// count + density of the flow particles of unflagged cells (masks must be current)
template <int KERNEL, class FT> static void launch_density(SphCtx *c, const DevF &d, int *count_out, float *rho_out) {
    const TileGeom g_ = make_geom<FT>(d.gn);
    cudaMemsetAsync(d.wcount + 6, 0, 4, c->stream);
    k_tile_density<KERNEL, FT><<<pgrid(k_tile_density<KERNEL, FT>, FT::BT, smem_of<FT, 1>(), nblocks<FT>(g_)), FT::BT, smem_of<FT, 1>(), c->stream>>>(d, g_, count_out, rho_out);
}
int tile_density_sweep(SphCtx *c, int32_t *count_out, float *rho_out) {
    if (need_masks(c)) return -3;
    DevF d = make_dev<float>(c);
    SPH_PROF(c, K_OTHER);
    k_density_payload<<<blocks_for(c->n, 256), 256, 0, c->stream>>>(d);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_C5);
    if (c->p.dim == 3) {
        if (c->p.kernel == 0) launch_density<0, F3M>(c, d, count_out, rho_out); else launch_density<1, F3M>(c, d, count_out, rho_out);
    } else {
        if (c->p.kernel == 0) launch_density<0, F2M>(c, d, count_out, rho_out); else launch_density<1, F2M>(c, d, count_out, rho_out);
    }
    SPH_LAUNCH_CHECK(c);
    return 0;
}
int tile_mask_count(SphCtx *c, int32_t *out) {
    DevF d = make_dev<float>(c);
    SPH_PROF(c, K_NEIGHBOR_COUNT);
    k_mask_count<<<blocks_for(c->n, 256), 256, 0, c->stream>>>(d, c->p.dim == 3 ? 27 : 9, out);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

}  // namespace sph
