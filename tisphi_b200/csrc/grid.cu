// grid.cu -- uniform-grid build: cell-id hash, histogram, inclusive scan, STABLE counting sort, coalesced reorder.
// Replaces ParticleSystem.initialize_particle_system (eng/particle_system.py:229-257):
//   update_grid_id (ps:229-236)  ->  k_cell_id        (cell id in float64, atomic histogram)
//   PrefixSumExecutor.run (ps:256) -> k_scan_*         (warp-shuffle block scan, 3 phases)
//   counting_sort (ps:239-252)   ->  k_scatter_index, k_rank, k_reorder
// The serial reference iterates I = N-1..0 with atomic_sub, which yields the STABLE order (ties keep their previous
// relative order).  On the GPU the atomic slot order inside a cell is arbitrary, so the rank inside the cell is
// recomputed deterministically as "number of particles of my cell with a smaller previous index".
// The reference moves its whole 800-byte record twice; here only the carried members move, once, as SoA streams.
#include <string.h>
#include "sph_host.h"

namespace sph {

// ------------------------------------------------------------------------------------------------ cell ids
// After the first step the arrays are already almost sorted, so the lanes of a warp mostly fall into one or two cells:
// the histogram update is aggregated per warp (one atomicAdd per distinct cell) with match.any.  Slots inside a cell
// are arbitrary anyway -- k_rank recomputes the stable order.
// Multi-GPU slabs: element v of the sort's input is read where it lies (SLAB; see VSrc in sph_host.h)
struct VLoc { const char *msg; int idx; };      // msg == nullptr: the member's current buffer
__device__ __forceinline__ VLoc vlocate(const VSrc &s, int v) {
    const int nl = s.ctl->src_nl, own = s.ctl->src_count;      // snapshot taken by k_slab_wait_set_n (the column table of
    VLoc r;                                                    // the new order overwrites own_first / own_count meanwhile)
    if (v < nl) { r.msg = inbox_msg((char *)s.inbox, 0, s.parity, s.msg_cap); r.idx = v; }
    else if (v < nl + own) { r.msg = nullptr; r.idx = s.ctl->src_first + (v - nl); }
    else { r.msg = inbox_msg((char *)s.inbox, 1, s.parity, s.msg_cap); r.idx = v - nl - own; }
    return r;
}
template <typename E> __device__ __forceinline__ const E *vptr(const VSrc &s, const VLoc &l, int slot, const E *cur) {
    return l.msg ? (const E *)(l.msg + s.sec[slot]) : cur;
}
template <typename T, bool SLAB>
__global__ void __launch_bounds__(256) k_cell_id(Dev<T> c, int *__restrict__ gid_out, int *__restrict__ slot, VSrc vs, int cell0, int cell1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < c.N();
    long long g = -1 - lane;                    // idle lanes: distinct keys that match nobody
    if (valid) {
        const double *xp = c.x;
        size_t xi = (size_t)i;
        if (SLAB) { const VLoc l = vlocate(vs, i); xp = vptr(vs, l, 0, c.x); xi = (size_t)l.idx; }
        const double x[3] = {xp[3 * xi], xp[3 * xi + 1], xp[3 * xi + 2]};
        int cc[3];
        pos_to_cell(c, x, cc);
        g = (long long)cc[0] * c.gn[1] * c.gn[2] + (long long)cc[1] * c.gn[2] + cc[2];
        if (g < 0 || g >= c.C) {                // the reference has no check; we clamp and count
            atomicAdd(c.bad, 1ull);
            g = g < 0 ? 0 : c.C - 1;
        }
        if (SLAB && (g < cell0 || g >= cell1)) {            // outside the slab's columns and their ghosts: it moved too far
            atomicOr(&const_cast<SlabCtl *>(vs.ctl)->err, 8);   // SLAB_ERR_FAR; the clamp only keeps the arrays consistent
            g = g < cell0 ? cell0 : cell1 - 1;
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
// Column cx starts at cell_end[cx * nyz - 1] (x-major cell ids, ps:221-222): own / ghost / boundary ranges of a slab and
// the regions the next redistribution has to look at (one thread, right after the scan).
__device__ void slab_coltable(const ColTab &t, const int *__restrict__ cell_end) {
    SlabCtl *ctl = t.ctl;
    const int n = ctl->n, a = t.a, b = t.b;
    auto start = [&](int cx) { return cx <= 0 ? 0 : (cx >= t.gn0 ? n : cell_end[(long long)cx * t.nyz - 1]); };
    const int ca1 = start(a - 1), ca = start(a), cb = start(b), cb1 = start(b + 1);
    int err = 0;
    if (ca1 != 0 || cb1 != n) err |= 8;                            // SLAB_ERR_FAR: something moved more than one column in a step
    if ((!t.has0 && ca != 0) || (!t.has1 && cb != n)) err |= 8;
    if (err) atomicOr(&ctl->err, err);
    ctl->own_first = ca; ctl->own_count = cb - ca;
    ctl->ghost_first[0] = ca1; ctl->ghost_count[0] = t.has0 ? ca - ca1 : 0;
    ctl->ghost_first[1] = cb; ctl->ghost_count[1] = t.has1 ? cb1 - cb : 0;
    ctl->send_first[0] = ca; ctl->send_count[0] = t.has0 ? start(a + 1) - ca : 0;
    const int cbm1 = start(b - 1);
    ctl->send_first[1] = cbm1; ctl->send_count[1] = t.has1 ? cb - cbm1 : 0;
    // particles move less than a cell per step: only the two old columns at each face can hold leavers
    const int l_end = start(min(a + 2, b)), r_beg = start(max(b - 2, a));
    ctl->reg_first[0] = ca; ctl->reg_count[0] = t.has0 ? l_end - ca : 0;
    ctl->reg_first[1] = r_beg; ctl->reg_count[1] = t.has1 ? cb - r_beg : 0;
}
__global__ void __launch_bounds__(256) k_scatter_index(int n, const int *__restrict__ ndev, const int *__restrict__ gid,
                                                       const int *__restrict__ slot, const int *__restrict__ cell_end,
                                                       int *__restrict__ tmpidx, ColTab ct) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (ndev) n = *ndev;
    if (i == 0 && ct.ctl) slab_coltable(ct, cell_end);
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
template <typename T, bool SLAB>
__global__ void __launch_bounds__(256) k_reorder(Dev<T> a, Dev<T> b, const int *__restrict__ perm,
                                                 const int *__restrict__ gid_unsorted, int soil, int init_tmp, VSrc vs) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.N()) return;
    const int s = perm[k];
    // where element s of the input lives: the current buffers, or (slabs) a section of an inbox message
    int si = s;
    const double *px = a.x, *prho = a.rho;
    const Vec4<T> *pxs = a.xs4, *pv = a.v4, *pvt = a.vt4;
    const T *ppress = a.press, *pstress = a.stress, *pstrain = a.strain, *pstrain_p = a.strain_p;
    const int *ptype = a.type, *pid0 = a.id0, *pflag = a.flag;
    if (SLAB) {
        const VLoc l = vlocate(vs, s);
        si = l.idx;
        px = vptr(vs, l, 0, a.x); pxs = vptr(vs, l, 1, a.xs4); pv = vptr(vs, l, 2, a.v4); pvt = vptr(vs, l, 3, a.vt4);
        prho = vptr(vs, l, 4, a.rho); ppress = vptr(vs, l, 5, a.press); ptype = vptr(vs, l, 6, a.type); pid0 = vptr(vs, l, 7, a.id0);
        if (soil) { pstress = vptr(vs, l, 8, a.stress); pstrain = vptr(vs, l, 9, a.strain); pstrain_p = vptr(vs, l, 10, a.strain_p); pflag = vptr(vs, l, 11, a.flag); }
    }
    const size_t k3 = 3 * (size_t)k, s3 = 3 * (size_t)si;
    const double x0 = px[s3], x1 = px[s3 + 1], x2 = px[s3 + 2];
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
    xs.w = pxs[si].w;                               // m_V travels with the particle
    b.xs4[k] = xs;
    const int ty = ptype[si];
    if (a.ps4) {                                    // cell-tile payloads: AoS for the passes, SoA for the mask kernel
        const bool fl = is_flow(ty);
        Vec4<T> ps = xs; ps.w = fl ? xs.w : -xs.w; a.ps4[k] = ps;
        a.psx[k] = xs.x; a.psy[k] = xs.y; a.psz[k] = xs.z; a.psf[k] = fl ? (T)1 : (T)-1;
        if (fl) a.cellflow[g] = 1;
    }
    const Vec4<T> v = pv[si];
    const double rho = prho[si];
    b.v4[k] = v;
    if (init_tmp && is_real(ty)) {                  // init_real2tmp (base:67-74) of the WCSPH step: tmp := real
        Vec4<T> vt = v; vt.w = (T)rho;
        b.vt4[k] = vt;
        a.rho_t[k] = rho;                           // density_tmp is not carried: one buffer, sorted order
    } else b.vt4[k] = pvt[si];
    b.rho[k] = rho;
    b.press[k] = ppress[si];
    b.type[k] = ty;
    b.id0[k] = pid0[si];
    if (soil) {
        const size_t k6 = 6 * (size_t)k, s6 = 6 * (size_t)si;
#pragma unroll
        for (int q = 0; q < 6; q++) b.stress[k6 + q] = pstress[s6 + q];
        b.strain[k] = pstrain[si];
        b.strain_p[k] = pstrain_p[si];
        b.flag[k] = pflag[si];
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
    c->masks_valid = false; c->gnl_valid = false;                       // masks, work lists and round lists index the previous order
    Dev<T> a = make_dev<T>(c, -1), b = make_dev<T>(c, 1);
    int *gid_u = (int *)(c->arena + c->off_gid_unsorted), *slot = (int *)(c->arena + c->off_slot);
    int *perm = (int *)(c->arena + c->off_perm), *tmpidx = (int *)(c->arena + c->off_tmpidx);
    int *tiles = (int *)(c->arena + c->off_scan_tiles);
    int *id_new = (int *)(c->arena + c->f[SPH_F_ID_NEW].off[0]);
    cudaStream_t st = c->stream;
    // the sort of a slab redistribution: input = virtual concatenation, cells of the slab's columns only, column table
    const bool slab = c->slab_sort;
    VSrc vs;
    ColTab ct;
    memset(&vs, 0, sizeof(vs));
    memset(&ct, 0, sizeof(ct));
    int cell0 = 0, cell1 = c->C;
    if (slab) slab_sort_args(c, &vs, &ct, &cell0, &cell1);
    const int nc = cell1 - cell0;
    SPH_CHECK(c, cudaMemsetAsync(a.cell_cnt + cell0, 0, sizeof(int) * (size_t)nc, st));
    if (c->fast) SPH_CHECK(c, cudaMemsetAsync(a.cellflow + cell0, 0, (size_t)nc, st));
    SPH_PROF(c, K_CELL_ID);
    if (slab) k_cell_id<T, true><<<blocks_for(n, 256), 256, 0, st>>>(a, gid_u, slot, vs, cell0, cell1);
    else k_cell_id<T, false><<<blocks_for(n, 256), 256, 0, st>>>(a, gid_u, slot, vs, cell0, cell1);
    SPH_LAUNCH_CHECK(c);
    const int nt = (nc + SCAN_TILE - 1) / SCAN_TILE;
    SPH_PROF(c, K_SCAN);
    k_scan_reduce<<<nt, SCAN_THREADS, 0, st>>>(a.cell_cnt + cell0, nc, tiles);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_SCAN);
    k_scan_tiles<<<1, 1024, 0, st>>>(tiles, nt);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_SCAN);
    k_scan_apply<<<nt, SCAN_THREADS, 0, st>>>(a.cell_cnt + cell0, nc, tiles, a.cell_end + cell0);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_SCATTER);
    k_scatter_index<<<blocks_for(n, 256), 256, 0, st>>>(n, a.ndev, gid_u, slot, a.cell_end, tmpidx, ct);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_RANK);
    k_rank<<<blocks_for(n, 256), 256, 0, st>>>(n, a.ndev, gid_u, a.cell_end, tmpidx, perm, id_new);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_REORDER);
    const int init_tmp = (c->fuse_init && !c->soil) ? 1 : 0;
    c->fuse_init = false;
    if (slab) k_reorder<T, true><<<blocks_for(n, 256), 256, 0, st>>>(a, b, perm, gid_u, c->soil ? 1 : 0, init_tmp, vs);
    else k_reorder<T, false><<<blocks_for(n, 256), 256, 0, st>>>(a, b, perm, gid_u, c->soil ? 1 : 0, init_tmp, vs);
    SPH_LAUNCH_CHECK(c);
    static const int carried[] = {SPH_F_X, SPH_F_XS, SPH_F_V, SPH_F_V_TMP, SPH_F_DENSITY, SPH_F_PRESSURE, SPH_F_MAT_TYPE, SPH_F_ID0};
    for (int f : carried) flip(c, f);
    if (c->soil) { flip(c, SPH_F_STRESS); flip(c, SPH_F_STRAIN_EQU); flip(c, SPH_F_STRAIN_EQU_P); flip(c, SPH_F_FLAG_RETMAP); }
    return 0;
}

template int grid_build<float>(SphCtx *);
template int grid_build<double>(SphCtx *);

}  // namespace sph
