// grid.cu -- uniform-grid build: cell-id hash, histogram, inclusive scan, STABLE counting sort, coalesced reorder.
// Replaces ParticleSystem.initialize_particle_system (eng/particle_system.py:229-257):
//   update_grid_id (ps:229-236)  ->  k_cell_id        (cell id in float64, atomic histogram)
//   PrefixSumExecutor.run (ps:256) -> k_scan_*         (warp-shuffle block scan, 3 phases)
//   counting_sort (ps:239-252)   ->  k_scatter_index, k_rank, k_reorder
// The serial reference iterates I = N-1..0 with atomic_sub, which yields the STABLE order (ties keep their previous
// relative order).  On the GPU the atomic slot order inside a cell is arbitrary, so the rank inside the cell is
// recomputed deterministically as "number of particles of my cell with a smaller previous index".
// The reference moves its whole 800-byte record twice; here only the carried members move, once, as SoA streams.
#include "sph_host.h"

namespace sph {

// ------------------------------------------------------------------------------------------------ cell ids
// After the first step the arrays are already almost sorted, so the lanes of a warp mostly fall into one or two cells:
// the histogram update is aggregated per warp (one atomicAdd per distinct cell) with match.any.  Slots inside a cell
// are arbitrary anyway -- k_rank recomputes the stable order.
template <typename T>
__global__ void __launch_bounds__(256) k_cell_id(Dev<T> c, int *__restrict__ gid_out, int *__restrict__ slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < c.N();
    long long g = -1 - lane;                    // idle lanes: distinct keys that match nobody
    if (valid) {
        const double x[3] = {c.x[3 * (size_t)i], c.x[3 * (size_t)i + 1], c.x[3 * (size_t)i + 2]};
        int cc[3];
        pos_to_cell(c, x, cc);
        g = (long long)cc[0] * c.gn[1] * c.gn[2] + (long long)cc[1] * c.gn[2] + cc[2];
        if (g < 0 || g >= c.C) {                // the reference has no check; we clamp and count
            atomicAdd(c.bad, 1ull);
            g = g < 0 ? 0 : c.C - 1;
        }
        gid_out[i] = (int)g;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, (int)g);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (valid && lane == leader) base = atomicAdd(&c.cell_cnt[(int)g], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (valid) slot[i] = base + __popc(peers & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------------------------------------ scan
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
// inclusive scan across the block of one value per thread; returns inclusive value, *total = block sum
__device__ __forceinline__ int block_incl_scan(int v, int *total) {
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = warp_incl_scan(v);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < nw ? wsum[lane] : 0;
        s = warp_incl_scan(s);
        wsum[lane] = s;
    }
    __syncthreads();
    int off = w > 0 ? wsum[w - 1] : 0;
    *total = wsum[nw - 1];
    __syncthreads();
    return inc + off;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int *__restrict__ in, int n, int *__restrict__ tile_sum) {
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) s += (base + k < n) ? in[base + k] : 0;
    int tot;
    block_incl_scan(s, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}
// exclusive scan of the tile sums, one block, chunked with a running carry
__global__ void __launch_bounds__(1024) k_scan_tiles(int *__restrict__ tile_sum, int nt) {
    int carry = 0;
    for (int base = 0; base < nt; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nt ? tile_sum[i] : 0;
        int tot;
        int inc = block_incl_scan(v, &tot);
        if (i < nt) tile_sum[i] = carry + inc - v;
        carry += tot;
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int *__restrict__ in, int n, const int *__restrict__ tile_off,
                                                             int *__restrict__ out) {
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    int v[SCAN_IPT], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
    int tot;
    int inc = block_incl_scan(s, &tot);
    int run = tile_off[blockIdx.x] + inc - s;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) { run += v[k]; if (base + k < n) out[base + k] = run; }
}

// ------------------------------------------------------------------------------------------------ stable rank
__global__ void __launch_bounds__(256) k_scatter_index(int n, const int *__restrict__ ndev, const int *__restrict__ gid,
                                                       const int *__restrict__ slot, const int *__restrict__ cell_end,
                                                       int *__restrict__ tmpidx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (ndev) n = *ndev;
    if (i >= n) return;
    int g = gid[i];
    int start = g > 0 ? cell_end[g - 1] : 0;
    tmpidx[start + slot[i]] = i;
}
__global__ void __launch_bounds__(256) k_rank(int n, const int *__restrict__ ndev, const int *__restrict__ gid,
                                              const int *__restrict__ cell_end, const int *__restrict__ tmpidx,
                                              int *__restrict__ perm, int *__restrict__ id_new) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (ndev) n = *ndev;
    if (i >= n) return;
    int g = gid[i];
    int start = g > 0 ? cell_end[g - 1] : 0, end = cell_end[g];
    int rank = 0;
    for (int k = start; k < end; k++) rank += (tmpidx[k] < i);
    perm[start + rank] = i;       // new slot -> previous index
    id_new[i] = start + rank;     // pt.id_new (ps:245)
}

// ------------------------------------------------------------------------------------------------ reorder
// One thread per DESTINATION slot: writes are fully coalesced; reads follow perm, which is near-identity between
// consecutive steps (particles move much less than a cell per step), so they are near-coalesced too.
template <typename T>
__global__ void __launch_bounds__(256) k_reorder(Dev<T> a, Dev<T> b, const int *__restrict__ perm,
                                                 const int *__restrict__ gid_unsorted, int soil, int init_tmp) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.N()) return;
    const int s = perm[k];
    const size_t k3 = 3 * (size_t)k, s3 = 3 * (size_t)s;
    const double x0 = a.x[s3], x1 = a.x[s3 + 1], x2 = a.x[s3 + 2];
    b.x[k3] = x0; b.x[k3 + 1] = x1; b.x[k3 + 2] = x2;
    const int g = gid_unsorted[s];
    a.gid[k] = g;                                   // grid_ids has a single (sorted) buffer
    // sweep coordinates: global in F64; local to the particle's cell in MIXED
    Vec4<T> xs;
    if (sizeof(T) == 8) { xs.x = (T)x0; xs.y = (T)x1; xs.z = (T)x2; }
    else {
        int cc[3];
        unflatten(a, g, cc);
        xs.x = (T)__dsub_rn(x0, cell_origin(a.vstart[0], a.gs, cc[0]));
        xs.y = (T)__dsub_rn(x1, cell_origin(a.vstart[1], a.gs, cc[1]));
        xs.z = (T)__dsub_rn(x2, cell_origin(a.vstart[2], a.gs, cc[2]));
    }
    xs.w = a.xs4[s].w;                              // m_V travels with the particle
    b.xs4[k] = xs;
    const int ty = a.type[s];
    if (a.ps4) {                                    // cell-tile payloads: AoS for the passes, SoA for the mask kernel
        const bool fl = is_flow(ty);
        Vec4<T> ps = xs; ps.w = fl ? xs.w : -xs.w; a.ps4[k] = ps;
        a.psx[k] = xs.x; a.psy[k] = xs.y; a.psz[k] = xs.z; a.psf[k] = fl ? (T)1 : (T)-1;
        if (fl) a.cellflow[g] = 1;
    }
    const Vec4<T> v = a.v4[s];
    const double rho = a.rho[s];
    b.v4[k] = v;
    if (init_tmp && is_real(ty)) {                  // init_real2tmp (base:67-74) of the WCSPH step: tmp := real
        Vec4<T> vt = v; vt.w = (T)rho;
        b.vt4[k] = vt;
        a.rho_t[k] = rho;                           // density_tmp is not carried: one buffer, sorted order
    } else b.vt4[k] = a.vt4[s];
    b.rho[k] = rho;
    b.press[k] = a.press[s];
    b.type[k] = ty;
    b.id0[k] = a.id0[s];
    if (soil) {
        const size_t k6 = 6 * (size_t)k, s6 = 6 * (size_t)s;
#pragma unroll
        for (int q = 0; q < 6; q++) b.stress[k6 + q] = a.stress[s6 + q];
        b.strain[k] = a.strain[s];
        b.strain_p[k] = a.strain_p[s];
        b.flag[k] = a.flag[s];
    }
}

// ------------------------------------------------------------------------------------------------ column selection
// Multi-GPU migration (tisphi_b200/parallel.py): the particles of [first, first + count) whose NEW cell column lies in
// [cx_lo, cx_hi], as a STABLE index list (previous order kept -- what the receiver's stable counting sort needs).
template <typename T>
__global__ void __launch_bounds__(256) k_flag_columns(Dev<T> c, int first, int count, int lo, int hi, int *__restrict__ flags) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const size_t i = (size_t)first + t;
    const int cx = (int)__ddiv_rn(__dsub_rn(c.x[3 * i], c.vstart[0]), c.gs);
    flags[t] = (cx >= lo && cx <= hi) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_compact(int first, int count, const int *__restrict__ flags, const int *__restrict__ incl,
                                                 int *__restrict__ idx, int *__restrict__ total) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    if (flags[t]) idx[incl[t] - 1] = first + t;
    if (t == count - 1) *total = incl[t];
}
template <typename T> int select_columns(SphCtx *c, int which, int64_t first, int64_t count, int lo, int hi) {
    int *total = (int *)(c->arena + c->off_bad + 16) + which;
    cudaStream_t st = c->stream;
    if (count == 0) { SPH_CHECK(c, cudaMemsetAsync(total, 0, 4, st)); return 0; }
    Dev<T> d = make_dev<T>(c);
    int *flags = (int *)(c->arena + c->off_perm), *incl = (int *)(c->arena + c->off_tmpidx);
    int *idx = (int *)(c->arena + (which == 0 ? c->off_slot : c->off_gid_unsorted));
    int *tiles = (int *)(c->arena + c->off_scan_tiles);
    const int n = (int)count, nt = (n + SCAN_TILE - 1) / SCAN_TILE;
    SPH_PROF(c, K_HALO);
    k_flag_columns<T><<<blocks_for(n, 256), 256, 0, st>>>(d, (int)first, n, lo, hi, flags);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_SCAN);
    k_scan_reduce<<<nt, SCAN_THREADS, 0, st>>>(flags, n, tiles);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_SCAN);
    k_scan_tiles<<<1, 1024, 0, st>>>(tiles, nt);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_SCAN);
    k_scan_apply<<<nt, SCAN_THREADS, 0, st>>>(flags, n, tiles, incl);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_HALO);
    k_compact<<<blocks_for(n, 256), 256, 0, st>>>((int)first, n, flags, incl, idx, total);
    SPH_LAUNCH_CHECK(c);
    return 0;
}
template int select_columns<float>(SphCtx *, int, int64_t, int64_t, int, int);
template int select_columns<double>(SphCtx *, int, int64_t, int64_t, int, int);

template <typename T> int grid_build(SphCtx *c) {
    const int n = (int)c->n;
    if (n == 0) return 0;
    c->masks_valid = false;                       // masks, work lists and round lists index the previous order
    Dev<T> a = make_dev<T>(c, -1), b = make_dev<T>(c, 1);
    int *gid_u = (int *)(c->arena + c->off_gid_unsorted), *slot = (int *)(c->arena + c->off_slot);
    int *perm = (int *)(c->arena + c->off_perm), *tmpidx = (int *)(c->arena + c->off_tmpidx);
    int *tiles = (int *)(c->arena + c->off_scan_tiles);
    int *id_new = (int *)(c->arena + c->f[SPH_F_ID_NEW].off[0]);
    cudaStream_t st = c->stream;
    SPH_CHECK(c, cudaMemsetAsync(a.cell_cnt, 0, sizeof(int) * (size_t)c->C, st));
    if (c->fast) SPH_CHECK(c, cudaMemsetAsync(a.cellflow, 0, (size_t)c->C, st));
    SPH_PROF(c, K_CELL_ID);
    k_cell_id<T><<<blocks_for(n, 256), 256, 0, st>>>(a, gid_u, slot);
    SPH_LAUNCH_CHECK(c);
    const int nt = (c->C + SCAN_TILE - 1) / SCAN_TILE;
    SPH_PROF(c, K_SCAN);
    k_scan_reduce<<<nt, SCAN_THREADS, 0, st>>>(a.cell_cnt, c->C, tiles);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_SCAN);
    k_scan_tiles<<<1, 1024, 0, st>>>(tiles, nt);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_SCAN);
    k_scan_apply<<<nt, SCAN_THREADS, 0, st>>>(a.cell_cnt, c->C, tiles, a.cell_end);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_SCATTER);
    k_scatter_index<<<blocks_for(n, 256), 256, 0, st>>>(n, a.ndev, gid_u, slot, a.cell_end, tmpidx);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_RANK);
    k_rank<<<blocks_for(n, 256), 256, 0, st>>>(n, a.ndev, gid_u, a.cell_end, tmpidx, perm, id_new);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_REORDER);
    const int init_tmp = (c->fuse_init && !c->soil) ? 1 : 0;
    c->fuse_init = false;
    k_reorder<T><<<blocks_for(n, 256), 256, 0, st>>>(a, b, perm, gid_u, c->soil ? 1 : 0, init_tmp);
    SPH_LAUNCH_CHECK(c);
    static const int carried[] = {SPH_F_X, SPH_F_XS, SPH_F_V, SPH_F_V_TMP, SPH_F_DENSITY, SPH_F_PRESSURE, SPH_F_MAT_TYPE, SPH_F_ID0};
    for (int f : carried) flip(c, f);
    if (c->soil) { flip(c, SPH_F_STRESS); flip(c, SPH_F_STRAIN_EQU); flip(c, SPH_F_STRAIN_EQU_P); flip(c, SPH_F_FLAG_RETMAP); }
    return 0;
}

template int grid_build<float>(SphCtx *);
template int grid_build<double>(SphCtx *);

}  // namespace sph
