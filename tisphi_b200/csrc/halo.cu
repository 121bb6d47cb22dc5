// halo.cu -- multi-GPU slab support: column ranges, message pack / unpack, particle-set replacement.
// The reference is single-device; this is the part of the engine that lets one rank own the x-columns
// [own0, own1) of the global grid.  Because the flattened cell id is x-major (eng/particle_system.py:221-222), a
// column is one contiguous index range of every member array after the sort: every message is a set of plain ranges
// and one launch of k_multi_copy moves all of them.
#include "sph_host.h"

namespace sph {

struct CopyDesc { const uint32_t *src; uint32_t *dst; long long nwords; int wpe; };   // wpe: 4-byte words per particle
constexpr int MAX_COPY = 16;
struct CopyBatch { CopyDesc d[MAX_COPY]; int n; long long total; };

// all ranges are 4-byte aligned (members are int32 / float32 / float64 arrays); grid-stride, coalesced both sides.
// W = uint4 when every range is 16-byte aligned and a multiple of 16 bytes (nwords then counts 16-byte units).
template <typename W> __global__ void __launch_bounds__(256) k_multi_copy(CopyBatch b) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < b.total; t += stride) {
        long long r = t;
#pragma unroll 1
        for (int k = 0; k < b.n; k++) {
            if (r < b.d[k].nwords) { ((W *)b.d[k].dst)[r] = ((const W *)b.d[k].src)[r]; break; }
            r -= b.d[k].nwords;
        }
    }
}
// gather: destination section k holds, for t < count, the particle idx[t] of member k (wpe words each)
__global__ void __launch_bounds__(256) k_multi_gather(CopyBatch b, const int *__restrict__ idx, int count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < b.total; t += stride) {
        long long r = t;
#pragma unroll 1
        for (int k = 0; k < b.n; k++) {
            if (r < b.d[k].nwords) {
                const int wpe = b.d[k].wpe, p = (int)(r / wpe), w = (int)(r - (long long)p * wpe);
                b.d[k].dst[r] = b.d[k].src[(long long)idx[p] * wpe + w];
                break;
            }
            r -= b.d[k].nwords;
        }
    }
}

static inline int64_t align16(int64_t v) { return (v + 15) / 16 * 16; }

// bytes per particle of a member and its buffer (current or alternate)
bool field_ref(SphCtx *c, int f, bool alt, char **ptr, int *elem_bytes) {
    if (f < 0 || f >= SPH_F_NUM || f == SPH_F_MASS || f == SPH_F_M_V || f == SPH_F_CELL_END || f == SPH_F_CELL_COUNT) return false;
    const FieldSlot &F = c->f[f];
    if (!F.present) return false;
    const int eb = F.kind == 0 ? 8 : (F.kind == 1 ? c->real_bytes : 4);
    *elem_bytes = F.stride * eb;
    *ptr = c->arena + F.off[alt ? 1 - F.cur : F.cur];
    return true;
}

static int launch_copy(SphCtx *c, const CopyBatch &b_in) {
    if (b_in.total == 0) return 0;
    CopyBatch b = b_in;
    bool wide = true;
    for (int k = 0; k < b.n; k++)
        wide = wide && (((uintptr_t)b.d[k].src | (uintptr_t)b.d[k].dst) & 15) == 0 && (b.d[k].nwords & 3) == 0;
    if (wide) { b.total = 0; for (int k = 0; k < b.n; k++) { b.d[k].nwords /= 4; b.total += b.d[k].nwords; } }
    long long blocks = (b.total + 256 * 4 - 1) / (256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    SPH_PROF(c, K_HALO);
    if (wide) k_multi_copy<uint4><<<(int)blocks, 256, 0, c->stream>>>(b);
    else k_multi_copy<uint32_t><<<(int)blocks, 256, 0, c->stream>>>(b);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// message <-> member ranges.  dir 0: members -> message, 1: message -> members (alt: into the alternate buffers)
static int message_copy(SphCtx *c, int nf, const int32_t *fields, int64_t first, int64_t count, char *msg, int dir, bool alt) {
    if (count == 0) return 0;
    if (nf > MAX_COPY) { snprintf(c->err, sizeof(c->err), "at most %d members per message", MAX_COPY); return -2; }
    if (first < 0 || count < 0 || first + count > c->n_max) { snprintf(c->err, sizeof(c->err), "message range outside the arrays"); return -2; }
    CopyBatch b;
    b.n = 0; b.total = 0;
    int64_t off = 0;
    for (int k = 0; k < nf; k++) {
        char *p; int eb;
        if (!field_ref(c, fields[k], alt, &p, &eb)) { snprintf(c->err, sizeof(c->err), "member %d cannot travel in a message", fields[k]); return -2; }
        char *a = p + first * eb, *m = msg + off;
        CopyDesc &d = b.d[b.n++];
        d.src = (const uint32_t *)(dir == 0 ? a : m);
        d.dst = (uint32_t *)(dir == 0 ? m : a);
        d.nwords = count * eb / 4;
        d.wpe = eb / 4;
        b.total += d.nwords;
        off += align16(count * eb);
    }
    return launch_copy(c, b);
}

static const int STATE[] = {SPH_F_X, SPH_F_XS, SPH_F_V, SPH_F_V_TMP, SPH_F_DENSITY, SPH_F_PRESSURE, SPH_F_MAT_TYPE, SPH_F_ID0};
static const int STATE_SOIL[] = {SPH_F_STRESS, SPH_F_STRAIN_EQU, SPH_F_STRAIN_EQU_P, SPH_F_FLAG_RETMAP};

}  // namespace sph

using namespace sph;

extern "C" {

int sph_state_fields(SphCtx *c, int32_t *out, int32_t capacity) {
    int n = 0;
    for (int f : STATE) { if (n < capacity) out[n] = f; n++; }
    if (c->soil) for (int f : STATE_SOIL) { if (n < capacity) out[n] = f; n++; }
    return n;
}

int64_t sph_message_bytes(SphCtx *c, int32_t nf, const int32_t *fields, int64_t count) {
    int64_t off = 0;
    for (int k = 0; k < nf; k++) {
        char *p; int eb;
        if (!field_ref(c, fields[k], false, &p, &eb)) return -1;
        off += align16(count * eb);
    }
    return off;
}

int sph_pack_fields(SphCtx *c, int32_t nf, const int32_t *fields, int64_t first, int64_t count, void *msg) {
    return message_copy(c, nf, fields, first, count, (char *)msg, 0, false);
}
int sph_unpack_fields(SphCtx *c, int32_t nf, const int32_t *fields, int64_t first, int64_t count, const void *msg) {
    return message_copy(c, nf, fields, first, count, (char *)msg, 1, false);
}

int sph_replace_particles(SphCtx *c, int64_t keep_first, int64_t keep_count, const void *left, int64_t nl, const void *right, int64_t nr) {
    int32_t fields[MAX_COPY];
    const int nf = sph_state_fields(c, fields, MAX_COPY);
    if (keep_first < 0 || keep_count < 0 || keep_first + keep_count > c->n || nl < 0 || nr < 0) {
        snprintf(c->err, sizeof(c->err), "sph_replace_particles: bad ranges"); return -2;
    }
    const int64_t total = nl + keep_count + nr;
    if (total > c->n_max) { snprintf(c->err, sizeof(c->err), "particle capacity %lld exceeded (%lld)", (long long)c->n_max, (long long)total); return -2; }
    int r;
    if ((r = message_copy(c, nf, fields, 0, nl, (char *)left, 1, true))) return r;
    if ((r = message_copy(c, nf, fields, nl + keep_count, nr, (char *)right, 1, true))) return r;
    if (keep_count > 0) {                       // current[keep_first ...] -> alternate[nl ...]
        CopyBatch b;
        b.n = 0; b.total = 0;
        for (int k = 0; k < nf; k++) {
            char *cur, *alt; int eb;
            field_ref(c, fields[k], false, &cur, &eb);
            field_ref(c, fields[k], true, &alt, &eb);
            CopyDesc &d = b.d[b.n++];
            d.src = (const uint32_t *)(cur + keep_first * eb);
            d.dst = (uint32_t *)(alt + nl * eb);
            d.nwords = keep_count * eb / 4;
            d.wpe = eb / 4;
            b.total += d.nwords;
        }
        if ((r = launch_copy(c, b))) return r;
    }
    for (int k = 0; k < nf; k++) flip(c, fields[k]);
    c->n = total;
    c->masks_valid = false; c->gnl_valid = false;
    return 0;
}

// ---- migration without a sort: stable selection of the particles whose NEW cell column lies in [cx_lo, cx_hi]
int sph_select_columns(SphCtx *c, int32_t which, int64_t first, int64_t count, int32_t cx_lo, int32_t cx_hi) {
    if (which < 0 || which > 1 || first < 0 || count < 0 || first + count > c->n) { snprintf(c->err, sizeof(c->err), "sph_select_columns: bad range"); return -2; }
    return c->p.precision == SPH_PREC_F64 ? select_columns<double>(c, which, first, count, cx_lo, cx_hi)
                                          : select_columns<float>(c, which, first, count, cx_lo, cx_hi);
}
int sph_select_counts(SphCtx *c, int64_t *n0, int64_t *n1) {
    int v[2] = {0, 0};
    SPH_CHECK(c, cudaMemcpyAsync(v, c->arena + c->off_bad + 16, 8, cudaMemcpyDeviceToHost, c->stream));
    SPH_CHECK(c, cudaStreamSynchronize(c->stream));
    *n0 = v[0]; *n1 = v[1];
    return 0;
}
int sph_pack_selected(SphCtx *c, int32_t which, int32_t nf, const int32_t *fields, int64_t count, void *msg) {
    if (count == 0) return 0;
    if (nf > MAX_COPY || which < 0 || which > 1) { snprintf(c->err, sizeof(c->err), "sph_pack_selected: bad arguments"); return -2; }
    const int *idx = (const int *)(c->arena + (which == 0 ? c->off_slot : c->off_gid_unsorted));
    CopyBatch b;
    b.n = 0; b.total = 0;
    int64_t off = 0;
    for (int k = 0; k < nf; k++) {
        char *p; int eb;
        if (!field_ref(c, fields[k], false, &p, &eb)) { snprintf(c->err, sizeof(c->err), "member %d cannot travel in a message", fields[k]); return -2; }
        CopyDesc &d = b.d[b.n++];
        d.src = (const uint32_t *)p;
        d.dst = (uint32_t *)((char *)msg + off);
        d.nwords = count * eb / 4;
        d.wpe = eb / 4;
        b.total += d.nwords;
        off += align16(count * eb);
    }
    long long blocks = (b.total + 256 * 4 - 1) / (256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    SPH_PROF(c, K_HALO);
    k_multi_gather<<<(int)blocks, 256, 0, c->stream>>>(b, idx, (int)count);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

int sph_column_starts(SphCtx *c, int32_t ncols, const int32_t *cx, int64_t *out) {
    const int64_t nyz = (int64_t)c->p.gn[1] * (c->p.dim == 3 ? c->p.gn[2] : 1);
    const int *cell_end = (const int *)(c->arena + c->f[SPH_F_CELL_END].off[0]);
    int tmp[64];
    if (ncols > 64) { snprintf(c->err, sizeof(c->err), "at most 64 columns per call"); return -2; }
    for (int k = 0; k < ncols; k++) {
        tmp[k] = 0;
        if (cx[k] <= 0) continue;
        if (cx[k] >= c->p.gn[0]) { tmp[k] = (int)c->n; continue; }
        SPH_CHECK(c, cudaMemcpyAsync(&tmp[k], cell_end + (cx[k] * nyz - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    }
    SPH_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < ncols; k++) out[k] = tmp[k];
    return 0;
}

}  // extern "C"
