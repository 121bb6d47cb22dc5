// slab.cu -- the multi-GPU slab step driven entirely from the device (the reference is single-device).
//
// tisphi_b200/parallel.py::SlabDriver states the protocol (and runs it from Python over torch.distributed, which is what
// the CPU gloo tests exercise): a rank owns the x-columns [a, b) of the GLOBAL grid plus one ghost column per side; per
// step it (1) selects -- stably, without a sort -- the owned particles whose NEW column is <= a resp. >= b - 1 and sends
// them to the neighbours (migrants + boundary column in one message), (2) sorts [from L][own][from R] once, (3) refreshes
// the ghost columns after every top-level loop of one_step.  The stable counting sort then makes every rank's order the
// restriction of the single-GPU order, so results are bit-identical to one GPU.
//
// This file is the same protocol with every host round trip removed:
//   * the particle count, the column table (own / ghost / boundary ranges) and the migration counts live in a device
//     control block (SlabCtl); kernels read the count through Dev::N(), launches are sized by the capacity;
//   * a message is not packed, sent and unpacked: the pack kernel STORES it straight into the neighbour's inbox through a
//     peer mapping of that buffer (NVLink; CUDA IPC between the per-GPU processes), followed by a system-scope release
//     store of the epoch number; the receiver's stream holds a one-block kernel that spins on that flag (acquire), then
//     the copy from the inbox into the arrays.  Inboxes are double-buffered by epoch parity: a sender can only be one
//     exchange ahead of a receiver, because its next push follows its own wait for the receiver's previous push;
//   * nothing between the phases returns to the host: sph_step(ctx, nsteps) enqueues whole steps.
// Waits are bounded (SLAB_WAIT_NS): a peer that never arrives sets an error bit instead of hanging the GPU.
#include "sph_host.h"

namespace sph {

constexpr int SLAB_MAXF = 16;
constexpr long long SLAB_WAIT_NS = 30ll * 1000 * 1000 * 1000;   // a neighbour may still be building its scene
enum { SLAB_ERR_TIMEOUT = 1, SLAB_ERR_COUNT = 2, SLAB_ERR_CAPACITY = 4, SLAB_ERR_FAR = 8, SLAB_ERR_FACE = 16 };

struct SlabField { char *cur, *alt; int wpe; long long secw; };   // member buffers, 4-byte words per particle, section offset (words)
struct SlabMsg { SlabField f[SLAB_MAXF]; int n; };

struct SlabState {
    int rank, world, a, b;
    bool has[2];
    int64_t face_cap, msg_cap;
    char *inbox, *peer[2];
    unsigned long long epoch;
    bool armed;
    int64_t n_exact, own_first, own_count;
    int err;
    unsigned long long sort_epoch;          // the migration exchange the pending sort reads
    SlabMsg sort_msg;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// every thread of a push kernel calls this after its stores: the LAST block of the side publishes count and epoch
__device__ __forceinline__ void push_finish(SlabCtl *ctl, char *peer_inbox, int side, unsigned long long epoch, int count) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&ctl->done[side], 1u);
        if (prev == gridDim.x - 1) {
            ctl->done[side] = 0;
            InboxHdr *h = (InboxHdr *)peer_inbox;
            const int from = 1 - side;                          // I am the right neighbour of my left neighbour
            h->count[from][epoch & 1ull] = count;
            __threadfence_system();
            st_release_sys(&h->flag[from], epoch);
        }
    }
}

// ---------------------------------------------------------------------------------------------- stable selection
// The owned particles of region [reg_first, +reg_count) whose NEW cell column lies in [lo, hi], as an index list in
// the previous order (what the receiver's stable sort needs).  Two launches, both sides at once (blockIdx.y):
// per-chunk counts, then every chunk sums the counts in front of it (a few hundred at most) and writes its indices.
constexpr int SEL_CHUNK = 1024, SEL_THREADS = 256;
template <typename T>
__device__ __forceinline__ unsigned sel_flags(const Dev<T> &c, int first, int count, int t0, int lo, int hi) {
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int t = t0 + k;
        if (t < count) {
            const size_t i = (size_t)first + t;
            const int cx = (int)__ddiv_rn(__dsub_rn(c.x[3 * i], c.vstart[0]), c.gs);      // ps:216-218
            if (cx >= lo && cx <= hi) m |= 1u << k;
        }
    }
    return m;
}
template <typename T>
__global__ void __launch_bounds__(SEL_THREADS) k_slab_select_count(Dev<T> c, SlabCtl *ctl, int a, int b, int *cnt0, int *cnt1) {
    const int side = blockIdx.y;
    const int first = ctl->reg_first[side], count = ctl->reg_count[side];
    if ((long long)blockIdx.x * SEL_CHUNK >= count) return;
    const int lo = side == 0 ? -(1 << 30) : b - 1, hi = side == 0 ? a : (1 << 30);
    const unsigned m = sel_flags(c, first, count, blockIdx.x * SEL_CHUNK + threadIdx.x * 4, lo, hi);
    __shared__ int ws[SEL_THREADS / 32];
    int v = __popc(m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < SEL_THREADS / 32; k++) s += ws[k];
        (side == 0 ? cnt0 : cnt1)[blockIdx.x] = s;
    }
}
template <typename T>
__global__ void __launch_bounds__(SEL_THREADS) k_slab_select_write(Dev<T> c, SlabCtl *ctl, int a, int b, const int *cnt0, const int *cnt1,
                                                                   int *idx0, int *idx1, int face_cap) {
    const int side = blockIdx.y;
    const int first = ctl->reg_first[side], count = ctl->reg_count[side];
    if (count <= 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->sel_count[side] = 0;
        return;
    }
    if ((long long)blockIdx.x * SEL_CHUNK >= count) return;
    const int *cnt = side == 0 ? cnt0 : cnt1;
    int *idx = side == 0 ? idx0 : idx1;
    const int lo = side == 0 ? -(1 << 30) : b - 1, hi = side == 0 ? a : (1 << 30);
    __shared__ int ws[SEL_THREADS / 32];
    __shared__ int s_base;
    // selected particles in the chunks in front of mine
    int part = 0;
    for (int k = threadIdx.x; k < (int)blockIdx.x; k += SEL_THREADS) part += cnt[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < SEL_THREADS / 32; k++) s += ws[k];
        s_base = s;
    }
    __syncthreads();
    const int t0 = blockIdx.x * SEL_CHUNK + threadIdx.x * 4;
    const unsigned m = sel_flags(c, first, count, t0, lo, hi);
    // exclusive scan of the per-thread counts across the block
    const int mine = __popc(m), lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; k++) woff += ws[k];
    int pos = s_base + woff + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (m & (1u << k)) {
            if (pos < face_cap) idx[pos] = first + t0 + k;
            pos++;
        }
    const bool last_chunk = (long long)(blockIdx.x + 1) * SEL_CHUNK >= count;
    if (last_chunk && threadIdx.x == SEL_THREADS - 1) {
        int total = pos;                                  // the last thread's running position = everything selected
        if (total > face_cap) { atomicOr(&ctl->err, SLAB_ERR_FACE); total = face_cap; }
        ctl->sel_count[side] = total;
    }
}

// ---------------------------------------------------------------------------------------------- pushes
// selected particles (index list) of every state member -> the neighbour's inbox
__global__ void __launch_bounds__(256) k_slab_push_selected(SlabMsg msg, SlabCtl *ctl, char *peer0, char *peer1, const int *idx0,
                                                            const int *idx1, long long msg_cap, unsigned long long epoch) {
    const int side = blockIdx.y;
    char *peer = side == 0 ? peer0 : peer1;
    if (!peer) return;
    const int *idx = side == 0 ? idx0 : idx1;
    const int count = ctl->sel_count[side];
    uint32_t *dst = (uint32_t *)inbox_msg(peer, 1 - side, (unsigned)(epoch & 1ull), msg_cap);
    const long long stride = (long long)gridDim.x * blockDim.x, t00 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = 0; k < msg.n; k++) {
        const int wpe = msg.f[k].wpe;
        const uint32_t *src = (const uint32_t *)msg.f[k].cur;
        uint32_t *d = dst + msg.f[k].secw;
        const long long nw = (long long)count * wpe;
        for (long long t = t00; t < nw; t += stride) {
            const int p = (int)(t / wpe), w = (int)(t - (long long)p * wpe);
            d[t] = src[(long long)idx[p] * wpe + w];
        }
    }
    push_finish(ctl, peer, side, epoch, count);
}
// the listed members of my boundary column -> the neighbour's inbox (its ghost column)
__global__ void __launch_bounds__(256) k_slab_push_range(SlabMsg msg, SlabCtl *ctl, char *peer0, char *peer1, long long msg_cap,
                                                         unsigned long long epoch) {
    const int side = blockIdx.y;
    char *peer = side == 0 ? peer0 : peer1;
    if (!peer) return;
    const int first = ctl->send_first[side], count = ctl->send_count[side];
    uint32_t *dst = (uint32_t *)inbox_msg(peer, 1 - side, (unsigned)(epoch & 1ull), msg_cap);
    const long long stride = (long long)gridDim.x * blockDim.x, t00 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = 0; k < msg.n; k++) {
        const int wpe = msg.f[k].wpe;
        const uint32_t *src = (const uint32_t *)msg.f[k].cur + (long long)first * wpe;
        uint32_t *d = dst + msg.f[k].secw;
        const long long nw = (long long)count * wpe;
        for (long long t = t00; t < nw; t += stride) d[t] = src[t];
    }
    push_finish(ctl, peer, side, epoch, count);
}
// bounded wait until the message of `epoch` from `side` is complete (one thread); false on timeout
__device__ __forceinline__ bool wait_flag(SlabCtl *ctl, const char *inbox, int side, unsigned long long epoch) {
    const InboxHdr *h = (const InboxHdr *)inbox;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(&h->flag[side]) < epoch) {
        if ((long long)(globaltimer_ns() - t0) > SLAB_WAIT_NS) { atomicOr(&ctl->err, SLAB_ERR_TIMEOUT); return false; }
        __nanosleep(100);
    }
    return true;
}
// one block, one thread per side: wait until both neighbours' migration messages are complete, then the particle
// count of the sort's input: arrivals from L + own range + arrivals from R
__global__ void k_slab_wait_set_n(SlabCtl *ctl, char *inbox, int has0, int has1, unsigned long long epoch, int n_max) {
    const int side = threadIdx.x;
    if (side <= 1 && (side == 0 ? has0 : has1)) wait_flag(ctl, inbox, side, epoch);
    __syncwarp();
    if (threadIdx.x == 0) {
        const InboxHdr *h = (const InboxHdr *)inbox;
        const int nl = has0 ? h->count[0][epoch & 1ull] : 0, nr = has1 ? h->count[1][epoch & 1ull] : 0;
        long long n = (long long)nl + ctl->own_count + nr;
        if (n > n_max) { ctl->err |= SLAB_ERR_CAPACITY; n = n_max; }     // (the sort then drops the tail; the step is flagged)
        ctl->n = (int)n;
        ctl->src_nl = nl; ctl->src_first = ctl->own_first; ctl->src_count = ctl->own_count;
    }
}
// inbox -> the listed members of my ghost columns
__global__ void __launch_bounds__(256) k_slab_unpack_range(SlabMsg msg, SlabCtl *ctl, char *inbox, int has0, int has1, long long msg_cap,
                                                           unsigned long long epoch) {
    const int side = blockIdx.y;
    if (!(side == 0 ? has0 : has1)) return;
    if (threadIdx.x == 0) wait_flag(ctl, inbox, side, epoch);     // every block waits for the message itself (acquire)
    __syncthreads();
    const InboxHdr *h = (const InboxHdr *)inbox;
    const int first = ctl->ghost_first[side];
    int count = ctl->ghost_count[side];
    const int sent = h->count[side][epoch & 1ull];
    if (sent != count) {                                  // both sides of a face must agree on the column population
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&ctl->err, SLAB_ERR_COUNT);
        count = min(count, sent);
    }
    const uint32_t *src = (const uint32_t *)inbox_msg(inbox, side, (unsigned)(epoch & 1ull), msg_cap);
    const long long stride = (long long)gridDim.x * blockDim.x, t00 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = 0; k < msg.n; k++) {
        const int wpe = msg.f[k].wpe;
        uint32_t *d = (uint32_t *)msg.f[k].cur + (long long)first * wpe;
        const uint32_t *s = src + msg.f[k].secw;
        const long long nw = (long long)count * wpe;
        for (long long t = t00; t < nw; t += stride) d[t] = s[t];
    }
}
// ---------------------------------------------------------------------------------------------- host side
static inline long long align16(long long v) { return (v + 15) / 16 * 16; }

static int build_msg(SphCtx *c, const int *fields, int nf, SlabMsg *m) {
    SlabState *S = c->slab;
    if (nf > SLAB_MAXF) { snprintf(c->err, sizeof(c->err), "at most %d members per message", SLAB_MAXF); return -2; }
    m->n = 0;
    long long off = 0;
    for (int k = 0; k < nf; k++) {
        char *cur, *alt; int eb;
        if (!field_ref(c, fields[k], false, &cur, &eb) || !field_ref(c, fields[k], true, &alt, &eb)) {
            snprintf(c->err, sizeof(c->err), "member %d cannot travel in a message", fields[k]); return -2;
        }
        SlabField &f = m->f[m->n++];
        f.cur = cur; f.alt = alt; f.wpe = eb / 4; f.secw = off / 4;
        off += align16(S->face_cap * eb);
    }
    if (off > S->msg_cap) { snprintf(c->err, sizeof(c->err), "message of %lld bytes exceeds the inbox buffers (%lld)", off, (long long)S->msg_cap); return -2; }
    return 0;
}
static inline SlabCtl *ctl_of(SphCtx *c) { return (SlabCtl *)(c->arena + c->off_slabctl); }
static inline int copy_grid(long long words) {
    long long b = (words + 256 * 8 - 1) / (256 * 8);
    return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

void slab_disarm(SphCtx *c) {          // the particle set is replaced from outside: the host knows the count again
    if (c->slab) c->slab->armed = false;
}
bool slab_armed(const SphCtx *c) { return c->slab && c->slab->armed; }
int64_t slab_exact_n(const SphCtx *c) { return (c->slab && c->slab->armed) ? c->slab->n_exact : c->n; }
const int *slab_ndev(SphCtx *c) { return (c->slab && c->slab->armed) ? &ctl_of(c)->n : nullptr; }
void slab_free(SphCtx *c) { delete c->slab; c->slab = nullptr; }
int slab_arm(SphCtx *c) {
    SlabState *S = c->slab;
    if (!S || S->armed) return 0;
    SlabCtl h;
    memset(&h, 0, sizeof(h));
    h.n = (int)c->n; h.own_first = 0; h.own_count = (int)c->n;
    for (int s = 0; s < 2; s++) { h.reg_first[s] = 0; h.reg_count[s] = S->has[s] ? (int)c->n : 0; }
    SPH_CHECK(c, cudaMemcpyAsync(ctl_of(c), &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    SPH_CHECK(c, cudaStreamSynchronize(c->stream));               // h is on the stack
    S->n_exact = c->n; S->own_first = 0; S->own_count = c->n;
    S->armed = true;
    c->n = c->n_max;                                              // from here on c->n only sizes launches (Dev::N() is the count)
    return 0;
}

static int slab_exchange(SphCtx *c, const int *fields, int nf) {
    SlabState *S = c->slab;
    if (nf == 0) return 0;
    SlabMsg m;
    int r = build_msg(c, fields, nf, &m);
    if (r) return r;
    const unsigned long long ep = ++S->epoch;
    long long wsum = 0;
    for (int k = 0; k < m.n; k++) wsum += m.f[k].wpe;
    const int grid = copy_grid(wsum * S->face_cap / 2);
    SlabCtl *ctl = ctl_of(c);
    SPH_PROF(c, K_HALO);
    k_slab_push_range<<<dim3(grid, 2), 256, 0, c->stream>>>(m, ctl, S->peer[0], S->peer[1], S->msg_cap, ep);
    SPH_LAUNCH_CHECK(c);
    // the copy out of the inbox starts with the wait for the message (every block spins on the flag; the grid is kept
    // small so that waiting blocks never fill the device: slabs of ONE device -- the test harness -- share it)
    SPH_PROF(c, K_HALO_WAIT);
    k_slab_unpack_range<<<dim3(grid < 96 ? grid : 96, 2), 256, 0, c->stream>>>(m, ctl, S->inbox, S->has[0], S->has[1], S->msg_cap, ep);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// the members each phase of one_step writes that neighbours read (same table as parallel.py::CudaSlabEngine)
int slab_refresh(SphCtx *c, int phase, bool final_phase, bool last_one_step) {
    if (!c->slab) return 0;
    int f[SLAB_MAXF], n = 0;
    const int solver = c->p.solver;
    if (solver == SPH_SOLVER_WC) {
        if (phase == 0) { f[n++] = SPH_F_V_TMP; f[n++] = SPH_F_DENSITY_TMP; f[n++] = SPH_F_PRESSURE; if (c->fast) f[n++] = SPH_F_PK4; }
    } else if (solver == SPH_SOLVER_MUI) {
        if (phase == 0) { f[n++] = SPH_F_STRESS_TMP; f[n++] = SPH_F_PRESSURE; }
        else if (phase == 1) { f[n++] = SPH_F_V_TMP; f[n++] = SPH_F_DENSITY_TMP; f[n++] = SPH_F_STRESS_TMP; }
    } else {
        if (phase == 0) f[n++] = SPH_F_STRESS_TMP;
        else if (phase == 1) { f[n++] = SPH_F_V_TMP; f[n++] = SPH_F_DENSITY_TMP; f[n++] = SPH_F_STRESS_TMP; }
    }
    if (final_phase) {
        // XSPH and the mu(I) regularisation sweep read neighbours after the last integrator kernel
        const bool needs_final = c->p.xsph || solver == SPH_SOLVER_MUI;
        if (last_one_step && !needs_final) return 0;              // nothing reads the ghosts' derivatives any more
        f[n++] = SPH_F_D_DENSITY; f[n++] = SPH_F_D_VEL;
        if (solver == SPH_SOLVER_DP) f[n++] = SPH_F_D_STRESS;
    }
    return slab_exchange(c, f, n);
}
int slab_refresh_post(SphCtx *c) {
    if (!c->slab || !(c->p.solver == SPH_SOLVER_MUI && c->p.xsph)) return 0;     // ghosts' XSPH sums are incomplete
    const int f[1] = {SPH_F_X};
    return slab_exchange(c, f, 1);
}

template <typename T> int slab_redistribute(SphCtx *c) {
    SlabState *S = c->slab;
    int r = slab_arm(c);
    if (r) return r;
    int32_t fields[SLAB_MAXF];
    const int nf = sph_state_fields(c, fields, SLAB_MAXF);
    SlabMsg m;
    if ((r = build_msg(c, fields, nf, &m))) return r;
    const unsigned long long ep = ++S->epoch;
    SlabCtl *ctl = ctl_of(c);
    Dev<T> d = make_dev<T>(c);
    cudaStream_t st = c->stream;
    int *cnt0 = (int *)(c->arena + c->off_perm), *cnt1 = (int *)(c->arena + c->off_tmpidx);
    int *idx0 = (int *)(c->arena + c->off_slot), *idx1 = (int *)(c->arena + c->off_gid_unsorted);
    const int nchunks = (int)((c->n_max + SEL_CHUNK - 1) / SEL_CHUNK);
    SPH_PROF(c, K_HALO);
    k_slab_select_count<T><<<dim3(nchunks, 2), SEL_THREADS, 0, st>>>(d, ctl, S->a, S->b, cnt0, cnt1);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_HALO);
    k_slab_select_write<T><<<dim3(nchunks, 2), SEL_THREADS, 0, st>>>(d, ctl, S->a, S->b, cnt0, cnt1, idx0, idx1, (int)S->face_cap);
    SPH_LAUNCH_CHECK(c);
    long long wsum = 0;
    for (int k = 0; k < m.n; k++) wsum += m.f[k].wpe;
    SPH_PROF(c, K_HALO);
    k_slab_push_selected<<<dim3(copy_grid(wsum * S->face_cap / 2), 2), 256, 0, st>>>(m, ctl, S->peer[0], S->peer[1], idx0, idx1, S->msg_cap, ep);
    SPH_LAUNCH_CHECK(c);
    SPH_PROF(c, K_HALO_WAIT);
    k_slab_wait_set_n<<<1, 32, 0, st>>>(ctl, S->inbox, S->has[0], S->has[1], ep, (int)c->n_max);
    SPH_LAUNCH_CHECK(c);
    // ONE sort of [from L | own | from R], read where it lies (grid.cu, VSrc); the column table comes out of it
    S->sort_epoch = ep;
    S->sort_msg = m;
    c->slab_sort = true;
    r = grid_build<T>(c);
    c->slab_sort = false;
    return r;
}
void slab_sort_args(SphCtx *c, VSrc *vs, ColTab *ct, int *cell0, int *cell1) {
    SlabState *S = c->slab;
    const int nyz = c->p.gn[1] * (c->p.dim == 3 ? c->p.gn[2] : 1);
    vs->ctl = ctl_of(c); vs->inbox = S->inbox; vs->has0 = S->has[0]; vs->has1 = S->has[1];
    vs->parity = (unsigned)(S->sort_epoch & 1ull); vs->msg_cap = S->msg_cap;
    for (int k = 0; k < S->sort_msg.n && k < 12; k++) vs->sec[k] = S->sort_msg.f[k].secw * 4;
    ct->ctl = ctl_of(c); ct->a = S->a; ct->b = S->b; ct->gn0 = c->p.gn[0]; ct->nyz = nyz; ct->has0 = S->has[0]; ct->has1 = S->has[1];
    // Cells of the slab's columns, its ghost columns and the columns a tile footprint of an owned cell can reach (up to
    // 4 columns wide + 1 halo): no particle lives beyond the ghost columns, so their cell_end is 0 below / n above --
    // exactly what scanning them yields; everything further out is never read.
    const int lo = S->a - 5 > 0 ? S->a - 5 : 0, hi = S->b + 5 < c->p.gn[0] ? S->b + 5 : c->p.gn[0];
    *cell0 = lo * nyz; *cell1 = hi * nyz;
}
template int slab_redistribute<float>(SphCtx *);
template int slab_redistribute<double>(SphCtx *);

}  // namespace sph

using namespace sph;

extern "C" {

int64_t sph_slab_inbox_bytes(SphCtx *c, int64_t face_cap) {
    // the largest message: the state record of a particle (migration), or every member a phase can refresh
    int32_t fields[SLAB_MAXF];
    const int nf = sph_state_fields(c, fields, SLAB_MAXF);
    long long state = 0, phase = 0;
    for (int k = 0; k < nf; k++) { char *p; int eb; if (field_ref(c, fields[k], false, &p, &eb)) state += align16(face_cap * eb); }
    const int ph[] = {SPH_F_V_TMP, SPH_F_DENSITY_TMP, SPH_F_PRESSURE, SPH_F_PK4, SPH_F_STRESS_TMP, SPH_F_D_DENSITY, SPH_F_D_VEL, SPH_F_D_STRESS, SPH_F_X};
    for (int f : ph) { char *p; int eb; if (field_ref(c, f, false, &p, &eb)) phase += align16(face_cap * eb); }
    const long long msg = ((state > phase ? state : phase) + 255) / 256 * 256;
    return SLAB_HDR + 4 * msg;
}

int sph_slab_init(SphCtx *c, int32_t rank, int32_t world, int32_t cx_begin, int32_t cx_end, int64_t face_cap, void *inbox, int64_t inbox_bytes) {
    if (rank < 0 || rank >= world || world < 1 || face_cap <= 0 || !inbox) { snprintf(c->err, sizeof(c->err), "sph_slab_init: bad arguments"); return -2; }
    int r = sph_set_owned_columns(c, cx_begin, cx_end);
    if (r) return r;
    const int64_t need = sph_slab_inbox_bytes(c, face_cap);
    if (inbox_bytes < need || (((uintptr_t)inbox) & 255)) { snprintf(c->err, sizeof(c->err), "sph_slab_init: inbox of %lld bytes (256-aligned) needed", (long long)need); return -2; }
    if (c->slab) delete c->slab;
    SlabState *S = new SlabState();
    memset(S, 0, sizeof(*S));
    S->rank = rank; S->world = world; S->a = cx_begin; S->b = cx_end;
    S->has[0] = rank > 0; S->has[1] = rank < world - 1;
    S->face_cap = face_cap; S->msg_cap = (need - SLAB_HDR) / 4;
    S->inbox = (char *)inbox;
    SPH_CHECK(c, cudaMemsetAsync(inbox, 0, SLAB_HDR, c->stream));
    // the slab sort only scans the cells near its columns: everything else must read as "no particle" from the start
    SPH_CHECK(c, cudaMemsetAsync(c->arena + c->f[SPH_F_CELL_END].off[0], 0, sizeof(int) * (size_t)(c->C + 1), c->stream));
    SPH_CHECK(c, cudaMemsetAsync(c->arena + c->f[SPH_F_CELL_COUNT].off[0], 0, sizeof(int) * (size_t)(c->C + 1), c->stream));
    SPH_CHECK(c, cudaStreamSynchronize(c->stream));
    c->slab = S;
    c->masks_valid = false; c->gnl_valid = false;
    return 0;
}
// left / right: the neighbours' inboxes as device pointers valid on THIS device (peer mappings; null at the ends)
int sph_slab_connect(SphCtx *c, void *left_inbox, void *right_inbox) {
    SlabState *S = c->slab;
    if (!S) { snprintf(c->err, sizeof(c->err), "sph_slab_connect before sph_slab_init"); return -2; }
    if ((S->has[0] && !left_inbox) || (S->has[1] && !right_inbox)) { snprintf(c->err, sizeof(c->err), "sph_slab_connect: a neighbour's inbox is missing"); return -2; }
    S->peer[0] = S->has[0] ? (char *)left_inbox : nullptr;
    S->peer[1] = S->has[1] ? (char *)right_inbox : nullptr;
    return 0;
}
// reads the device control block back: particle count, owned range, sticky error bits.  Synchronises the stream.
int sph_slab_sync(SphCtx *c, int64_t *n, int64_t *own_first, int64_t *own_count, int32_t *err) {
    SlabState *S = c->slab;
    if (!S) { snprintf(c->err, sizeof(c->err), "not a slab ctx"); return -2; }
    if (S->armed) {
        SlabCtl h;
        SPH_CHECK(c, cudaMemcpyAsync(&h, ctl_of(c), sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        SPH_CHECK(c, cudaStreamSynchronize(c->stream));
        S->n_exact = h.n; S->own_first = h.own_first; S->own_count = h.own_count; S->err = h.err;
    } else { S->n_exact = c->n; S->own_first = 0; S->own_count = c->n; }
    if (n) *n = S->n_exact;
    if (own_first) *own_first = S->own_first;
    if (own_count) *own_count = S->own_count;
    if (err) *err = S->err;
    if (S->err) {
        snprintf(c->err, sizeof(c->err), "slab step failed:%s%s%s%s%s", (S->err & SLAB_ERR_TIMEOUT) ? " a neighbour's message never arrived;" : "",
                 (S->err & SLAB_ERR_COUNT) ? " ghost / boundary column sizes disagree;" : "", (S->err & SLAB_ERR_CAPACITY) ? " particle capacity exceeded;" : "",
                 (S->err & SLAB_ERR_FAR) ? " particles moved more than one column in a step or left the outermost slab;" : "",
                 (S->err & SLAB_ERR_FACE) ? " more particles on a face than the inbox holds;" : "");
        return -4;
    }
    return 0;
}
int64_t sph_slab_epoch(SphCtx *c) { return c->slab ? (int64_t)c->slab->epoch : 0; }

// CUDA IPC plumbing for one process per GPU: the inbox must be its own cudaMalloc allocation
void *sph_ipc_alloc(int64_t bytes) {
    void *p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) return nullptr;
    return p;
}
void sph_ipc_free(void *p) { if (p) cudaFree(p); }
int sph_ipc_get_handle(void *dev_ptr, void *handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, dev_ptr) != cudaSuccess) { cudaGetLastError(); return -1; }
    memcpy(handle64, &h, 64);
    return 0;
}
void *sph_ipc_open(const void *handle64) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void sph_ipc_close(void *p) { if (p) cudaIpcCloseMemHandle(p); }

}  // extern "C"
