/*
 * sph_oracle.c -- CPU restatement (float64) of the tiSPHi hot path.   TEST INFRASTRUCTURE ONLY.
 *
 * This file is the CHECKER for the CUDA engine in tisphi_b200/csrc.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product path never does.
 *
 * It restates, loop for loop, the algorithm of Rabmelon/tiSPHi (paths relative to /root/reference):
 *   eng/particle_system.py:216-269   grid ids, counting sort, for_all_neighbors
 *   eng/solver_sph_base.py:41-238    step skeleton, SE/LF/RK4 integrators, shared tasks, advect_pos (XSPH)
 *   eng/solver_sph_base.py:278-423   cubic / Wendland C2 kernels, CSPM_f, CSPM_L
 *   eng/solver_sph_base.py:647-669   Adami dummy-wall tasks;  :713-715 viscous damping
 *   eng/solver_sph_wc.py:7-132, eng/solver_sph_muI.py:7-156, eng/solver_sph_dp.py:7-296
 * Parity pin: tests/test_oracle_golden.py compares it with fixtures produced by executing those reference
 * sources themselves under the serial Taichi emulator (oracle/gen_golden.py).
 *
 * Race semantics.  The reference's loops read fields that the same loop writes.
 *   serial = 1 : reproduce the single-threaded reference exactly (what the fixtures contain):
 *                - WCSPH wall pressure reads p_j in place: EOS(rho~_j) for j < i, previous value for j > i
 *                - mu(I) stress regularisation and XSPH run as in-place (Gauss-Seidel) loops
 *   serial = 0 : "jacobi": wall pressure still uses the j<i rule (it is pointwise-emulable and is what the
 *                CUDA engine implements), but regularisation / XSPH read a pre-loop snapshot (GPU-reproducible).
 *   wc_fresh=1 : wall pressure reads EOS(rho~_j) for every j (race-free variant, not the reference's).
 *
 * Build:  gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared sph_oracle.c -o _build/libsph_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int dim, kernel, kcorr, ti, xsph, solver, serial, wc_fresh;
    int gn[3];
    double h, support, grid_size, vstart[3], m_V0, g[3], dt, eps;
    double rho0, visc, stiff, gamma_;                                  /* WCSPH material (wc:12-15)      */
    double coh, fric, E, poi, dila, vsound, mu, alpha, kc, G, K, eps_f; /* soil (muI:12-24, dp:12-29)     */
    /* boundary treatment (ps:18, 25): 0 none, 1 enforced collision, 2 dummy, 3 repulsive, 4 dummy + repulsive */
    int boundary, pad_;
    double radius, dstart[3], dend[3];                                 /* particleRadius, domainStart / domainEnd */
} OrcParams;

enum { F_X, F_V, F_M_V, F_DENSITY, F_MASS, F_PRESSURE, F_STRESS, F_CSPM_F, F_CSPM_L, F_D_DENSITY, F_D_VEL,
       F_D_STRESS, F_V_GRAD, F_STRAIN_EQU, F_D_STRAIN_EQU, F_STRAIN_EQU_P, F_D_STRAIN_EQU_P, F_DENSITY_TMP,
       F_V_TMP, F_STRESS_TMP, F_D_DENSITY_RK, F_D_VEL_RK, F_D_STRESS_RK, F_X0, F_NUM };
static const int F_NC[F_NUM] = {3, 3, 1, 1, 1, 1, 9, 1, 9, 1, 3, 9, 9, 1, 1, 1, 1, 1, 3, 9, 1, 3, 9, 3};
enum { I_MAT_TYPE, I_ID0, I_GRID_IDS, I_FLAG_RETMAP, I_OBJ_ID, I_IS_DYNAMIC, I_NUM };
#define MAX_OBJ 64

typedef struct {
    OrcParams p;
    int64_t n, C;
    double *f[F_NUM];
    int32_t *ia[I_NUM];
    int64_t *cell_end;      /* inclusive scan of the histogram == grid_particle_num after prefix sum (ps:256) */
    int64_t *cell_tmp;
    double *scratch;        /* 9*n doubles */
    int32_t *iscratch;
    double rest_cm[MAX_OBJ][3];   /* ps.rigid_rest_cm (ps:118), per object id */
    int dyn_obj[MAX_OBJ], n_dyn;  /* object ids of the DYNAMIC rigid bodies */
} Orc;

#define X(o, i) (&(o)->f[F_X][3 * (i)])
#define TYPE(o, i) ((o)->ia[I_MAT_TYPE][i])
static inline int is_fluid(int t) { return t == 1; }
static inline int is_soil(int t) { return t == 2; }
static inline int is_flow(int t) { return t == 1 || t == 2; }
static inline int is_real(int t) { return t > 0; }
static inline int is_bdy(int t) { return t == -1 || t == -2; }
static inline int is_rigid(int t) { return t == 11; }
#define RIGID_DYN(o, i) (TYPE(o, i) == 11 && (o)->ia[I_IS_DYNAMIC][i] != 0)      /* ps:356-362 */

/* ------------------------------------------------------------------------------------------------ API */
Orc *orc_create(const OrcParams *p, int64_t n) {
    Orc *o = (Orc *)calloc(1, sizeof(Orc));
    o->p = *p;
    o->n = n;
    o->C = (int64_t)p->gn[0] * p->gn[1] * (p->dim == 3 ? p->gn[2] : 1);
    for (int k = 0; k < F_NUM; k++) o->f[k] = (double *)calloc((size_t)n * F_NC[k] + 1, sizeof(double));
    for (int k = 0; k < I_NUM; k++) o->ia[k] = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
    o->cell_end = (int64_t *)calloc((size_t)o->C + 1, sizeof(int64_t));
    o->cell_tmp = (int64_t *)calloc((size_t)o->C + 1, sizeof(int64_t));
    o->scratch = (double *)calloc((size_t)n * 9 + 1, sizeof(double));
    o->iscratch = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
    return o;
}
void orc_destroy(Orc *o) {
    for (int k = 0; k < F_NUM; k++) free(o->f[k]);
    for (int k = 0; k < I_NUM; k++) free(o->ia[k]);
    free(o->cell_end); free(o->cell_tmp); free(o->scratch); free(o->iscratch); free(o);
}
double *orc_field(Orc *o, int k) { return o->f[k]; }
int32_t *orc_ifield(Orc *o, int k) { return o->ia[k]; }
int64_t *orc_cell_end(Orc *o) { return o->cell_end; }
int orc_field_ncomp(int k) { return F_NC[k]; }
/* threads of the parallel loops (bench.py's CPU arm sets and reports them: torchrun exports OMP_NUM_THREADS=1) */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_params(Orc *o, const OrcParams *p) { o->p = *p; }

/* --------------------------------------------------------------------------------- grid (ps:216-257) */
static inline void pos_to_index(const OrcParams *p, const double *x, int c[3]) {
    for (int a = 0; a < 3; a++) c[a] = (int)((x[a] - p->vstart[a]) / p->grid_size); /* C cast truncates (ps:218) */
}
static inline int64_t flatten(const OrcParams *p, const int c[3]) {
    return (int64_t)c[0] * p->gn[1] * p->gn[2] + (int64_t)c[1] * p->gn[2] + c[2];   /* ps:222 */
}

static void permute(Orc *o, const int32_t *newidx) {
    int64_t n = o->n;
    for (int k = 0; k < F_NUM; k++) {
        int nc = F_NC[k];
        double *src = o->f[k], *dst = o->scratch;
        for (int64_t i = 0; i < n; i++) memcpy(dst + (size_t)newidx[i] * nc, src + (size_t)i * nc, sizeof(double) * nc);
        memcpy(src, dst, sizeof(double) * n * nc);
    }
    for (int k = 0; k < I_NUM; k++) {
        int32_t *src = o->ia[k], *dst = (int32_t *)o->scratch;
        for (int64_t i = 0; i < n; i++) dst[newidx[i]] = src[i];
        memcpy(src, dst, sizeof(int32_t) * n);
    }
}

/* returns the number of particles whose cell is outside the grid (H7); those are clamped into range */
int64_t orc_grid_build(Orc *o) {
    const OrcParams *p = &o->p;
    int64_t n = o->n, C = o->C, bad = 0;
    memset(o->cell_end, 0, sizeof(int64_t) * C);
    for (int64_t i = 0; i < n; i++) {                    /* update_grid_id (ps:229-236) */
        int c[3];
        pos_to_index(p, X(o, i), c);
        int64_t g = flatten(p, c);
        if (g < 0 || g >= C) { bad++; g = g < 0 ? 0 : C - 1; }
        o->ia[I_GRID_IDS][i] = (int32_t)g;
        o->cell_end[g]++;
    }
    memcpy(o->cell_tmp, o->cell_end, sizeof(int64_t) * C);
    for (int64_t c = 1; c < C; c++) o->cell_end[c] += o->cell_end[c - 1];   /* inclusive scan (ps:256) */
    int32_t *newidx = o->iscratch;
    for (int64_t k = 0; k < n; k++) {                    /* counting_sort, serial: I descending (ps:240-245) */
        int64_t I = n - 1 - k;
        int64_t g = o->ia[I_GRID_IDS][I];
        int64_t base = (g - 1 >= 0) ? o->cell_end[g - 1] : 0;
        newidx[I] = (int32_t)((o->cell_tmp[g]--) - 1 + base);
    }
    permute(o, newidx);                                  /* ps:247-252 */
    return bad;
}

/* ------------------------------------------------------------------- neighbour iteration (ps:259-269) */
/* Out-of-range neighbour cells are defined empty per axis. BODY sees: j, d[3] = x_i - x_j, r. */
#define FOR_NEIGHBORS(o, i, ...)                                                                        \
    do {                                                                                                \
        const OrcParams *p_ = &(o)->p;                                                                  \
        int cc_[3];                                                                                     \
        const double *xi_ = X(o, i);                                                                    \
        pos_to_index(p_, xi_, cc_);                                                                     \
        for (int ox_ = -1; ox_ <= 1; ox_++)                                                             \
            for (int oy_ = -1; oy_ <= 1; oy_++)                                                         \
                for (int oz_ = -1; oz_ <= 1; oz_++) {                                                   \
                    if (p_->dim == 2 && oz_ != 0) continue;                                             \
                    int c_[3] = {cc_[0] + ox_, cc_[1] + oy_, cc_[2] + oz_};                             \
                    if (c_[0] < 0 || c_[0] >= p_->gn[0] || c_[1] < 0 || c_[1] >= p_->gn[1] ||           \
                        c_[2] < 0 || c_[2] >= p_->gn[2]) continue;                                      \
                    int64_t g_ = flatten(p_, c_);                                                       \
                    int64_t jb_ = g_ > 0 ? (o)->cell_end[g_ - 1] : 0, je_ = (o)->cell_end[g_];          \
                    for (int64_t j = jb_; j < je_; j++) {                                               \
                        if (j == (i)) continue;                                                         \
                        const double *xj_ = X(o, j);                                                    \
                        double d[3] = {xi_[0] - xj_[0], xi_[1] - xj_[1], xi_[2] - xj_[2]};              \
                        double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);                       \
                        if (r < p_->support) { __VA_ARGS__ }                                             \
                    }                                                                                   \
                }                                                                                       \
    } while (0)

/* ----------------------------------------------------------------------------- kernels (base:278-358) */
static inline double knorm(const OrcParams *p) {
    double h1 = 1.0 / p->h, k;
    if (p->kernel == 0) k = p->dim == 2 ? 15.0 / 7.0 / M_PI : 3.0 / 2.0 / M_PI;
    else k = p->dim == 2 ? 7.0 / (4.0 * M_PI) : 21.0 / (2.0 * M_PI);       /* 3D value as in the reference (H8) */
    double hp = h1;
    for (int a = 1; a < p->dim; a++) hp *= h1;
    return k * hp;
}
static inline double W(const OrcParams *p, double r) {
    double h1 = 1.0 / p->h, k = knorm(p), q = r * h1, res = 0.0;
    if (r > p->eps && q <= 2.0) {
        if (p->kernel == 0) {
            if (q <= 1.0) { double q2 = q * q, q3 = q2 * q; res = k * (0.5 * q3 - q2 + 2.0 / 3.0); }
            else res = k / 6.0 * pow(2.0 - q, 3.0);
        } else {
            double q1 = 1.0 - 0.5 * q;
            res = k * pow(q1, 4.0) * (1.0 + 2.0 * q);
        }
    }
    return res;
}
static inline void gradW(const OrcParams *p, const double d[3], double r, double out[3]) {
    double h1 = 1.0 / p->h, k = knorm(p), q = r * h1;
    out[0] = out[1] = out[2] = 0.0;
    if (r > p->eps && q <= 2.0) {
        if (p->kernel == 0) {
            double s = (q <= 1.0) ? k * q * (3.0 / 2.0 * q - 2.0) : k * (-0.5 * (2.0 - q) * (2.0 - q));
            for (int a = 0; a < 3; a++) out[a] = s * (d[a] / r * h1);
        } else {
            double q1 = 1.0 - 0.5 * q;
            double s = k * pow(q1, 3.0) * (-5.0 * q) * h1;
            for (int a = 0; a < 3; a++) out[a] = s * d[a] / r;
        }
    }
}
/* kernel_deriv_corr (base:374-383) */
static inline void gradWc(const Orc *o, int64_t i, const double d[3], double r, double out[3]) {
    const OrcParams *p = &o->p;
    if (p->kcorr == 0) { gradW(p, d, r, out); return; }
    if (p->kcorr == 1) {
        double g[3];
        gradW(p, d, r, g);
        const double *L = &o->f[F_CSPM_L][9 * i];
        for (int a = 0; a < 3; a++) out[a] = L[3 * a] * g[0] + L[3 * a + 1] * g[1] + L[3 * a + 2] * g[2];
        return;
    }
    out[0] = out[1] = out[2] = 0.0;
}

/* ------------------------------------------------------------------- kernel correction (base:386-423) */
static double det3(const double *m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
void orc_calc_kernel_corr(Orc *o) {
    const OrcParams *p = &o->p;
    int64_t n = o->n;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {                    /* calc_CSPM_f: every particle, flow neighbours */
        double s = 0.0;
        FOR_NEIGHBORS(o, i, { if (is_flow(TYPE(o, j))) s += o->f[F_M_V][j] * W(p, r); });
        o->f[F_CSPM_F][i] = (s != 0.0) ? 1.0 / s : 1.0;
    }
    if (p->kcorr != 1) return;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {                    /* calc_CSPM_L */
        double Li[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (is_flow(TYPE(o, i))) {
            double M[9] = {0};
            FOR_NEIGHBORS(o, i, {
                if (TYPE(o, j) == TYPE(o, i)) {
                    double g[3];
                    gradW(p, d, r, g);
                    double V = o->f[F_M_V][j];
                    for (int a = 0; a < 3; a++)
                        for (int b = 0; b < 3; b++) M[3 * a + b] += V * (-d[a]) * g[b];   /* (x_j - x_i) (x) gradW */
                }
            });
            if (p->dim == 2) {
                double det = M[0] * M[4] - M[1] * M[3];
                if (fabs(det) > p->eps) {
                    double inv = 1.0 / det;
                    double L2[9] = {M[4] * inv, -M[1] * inv, 0, -M[3] * inv, M[0] * inv, 0, 0, 0, 0};
                    memcpy(Li, L2, sizeof(Li));
                }
            } else {
                double det = det3(M);
                if (fabs(det) > p->eps) {
                    double inv = 1.0 / det;
                    double c[9] = {(M[4] * M[8] - M[5] * M[7]), -(M[1] * M[8] - M[2] * M[7]), (M[1] * M[5] - M[2] * M[4]),
                                   -(M[3] * M[8] - M[5] * M[6]), (M[0] * M[8] - M[2] * M[6]), -(M[0] * M[5] - M[2] * M[3]),
                                   (M[3] * M[7] - M[4] * M[6]), -(M[0] * M[7] - M[1] * M[6]), (M[0] * M[4] - M[1] * M[3])};
                    for (int a = 0; a < 9; a++) Li[a] = c[a] * inv;
                }
            }
        }
        memcpy(&o->f[F_CSPM_L][9 * i], Li, sizeof(Li));
    }
}

/* ------------------------------------------------------------------------- integrators (base:67-180) */
void orc_init_real2tmp(Orc *o) {
    for (int64_t i = 0; i < o->n; i++) {
        int t = TYPE(o, i);
        if (is_real(t)) {
            o->f[F_DENSITY_TMP][i] = o->f[F_DENSITY][i];
            memcpy(&o->f[F_V_TMP][3 * i], &o->f[F_V][3 * i], 24);
        }
        if (is_soil(t)) memcpy(&o->f[F_STRESS_TMP][9 * i], &o->f[F_STRESS][9 * i], 72);
    }
}
/* kind: 0 advect_SE/advect_LF (full dt), 1 advect_LF_half, 2 advect_RK_4, 3 init_RK, 4 update_RK(m), 5 advect_RK */
void orc_advect(Orc *o, int kind, int m) {
    double dt = o->p.dt;
    double *rho = o->f[F_DENSITY], *rhot = o->f[F_DENSITY_TMP], *dr = o->f[F_D_DENSITY], *drk = o->f[F_D_DENSITY_RK];
    double *v = o->f[F_V], *vt = o->f[F_V_TMP], *dv = o->f[F_D_VEL], *dvk = o->f[F_D_VEL_RK];
    double *s = o->f[F_STRESS], *st = o->f[F_STRESS_TMP], *ds = o->f[F_D_STRESS], *dsk = o->f[F_D_STRESS_RK];
    double *mV = o->f[F_M_V], *mass = o->f[F_MASS];
    for (int64_t i = 0; i < o->n; i++) {
        int t = TYPE(o, i), re = is_real(t), so = is_soil(t);
        switch (kind) {
        case 0:
            if (re) { rho[i] += dt * dr[i]; mV[i] = mass[i] / rho[i]; for (int a = 0; a < 3; a++) v[3 * i + a] += dt * dv[3 * i + a]; }
            if (so) for (int a = 0; a < 9; a++) s[9 * i + a] += dt * ds[9 * i + a];
            break;
        case 1:
            if (re) { rhot[i] += 0.5 * dt * dr[i]; mV[i] = mass[i] / rhot[i]; for (int a = 0; a < 3; a++) vt[3 * i + a] += 0.5 * dt * dv[3 * i + a]; }
            if (so) for (int a = 0; a < 9; a++) st[9 * i + a] += 0.5 * dt * ds[9 * i + a];
            break;
        case 2:
            if (re) { rhot[i] = 0.5 * dt * dr[i] + rho[i]; mV[i] = mass[i] / rhot[i]; for (int a = 0; a < 3; a++) vt[3 * i + a] = 0.5 * dt * dv[3 * i + a] + v[3 * i + a]; }
            if (so) for (int a = 0; a < 9; a++) st[9 * i + a] = 0.5 * dt * ds[9 * i + a] + s[9 * i + a];
            break;
        case 3:
            if (re) { drk[i] = 0.0; for (int a = 0; a < 3; a++) dvk[3 * i + a] = 0.0; }
            if (so) for (int a = 0; a < 9; a++) dsk[9 * i + a] = 0.0;
            break;
        case 4:
            if (re) { drk[i] += dr[i] * m; for (int a = 0; a < 3; a++) dvk[3 * i + a] += dv[3 * i + a] * m; }
            if (so) for (int a = 0; a < 9; a++) dsk[9 * i + a] += ds[9 * i + a] * m;
            break;
        case 5:
            if (re) { rho[i] += dt / 6.0 * drk[i]; mV[i] = mass[i] / rho[i]; for (int a = 0; a < 3; a++) v[3 * i + a] += dt / 6.0 * dvk[3 * i + a]; }
            if (so) for (int a = 0; a < 9; a++) s[9 * i + a] += dt / 6.0 * dsk[9 * i + a];
            break;
        }
    }
}

/* ------------------------------------------------------------------- shared neighbour sums (base:186-203) */
static void sum_vgrad_ddens(const Orc *o, int64_t i, double vg[9], double *dd) {
    const double *vi = &o->f[F_V_TMP][3 * i];
    double acc = 0.0;
    for (int a = 0; a < 9; a++) vg[a] = 0.0;
    FOR_NEIGHBORS(o, i, {
        double g[3];
        gradWc(o, i, d, r, g);
        const double *vj = &o->f[F_V_TMP][3 * j];
        double V = o->f[F_M_V][j];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) vg[3 * a + b] += V * (vj[a] - vi[a]) * g[b];
        acc += V * (vi[0] - vj[0]) * g[0] + V * (vi[1] - vj[1]) * g[1] + V * (vi[2] - vj[2]) * g[2];
    });
    *dd = acc;
}

/* Adami wall sums over flow neighbours (base:647-669): Sv = sum V v~ W, Sr = sum V rho~ W, Ss = sum V (sigma~ + rho~ diag(g.*dx)) W */
static void wall_sums(const Orc *o, int64_t i, double Sv[3], double *Sr, double Ss[9]) {
    const OrcParams *p = &o->p;
    Sv[0] = Sv[1] = Sv[2] = 0.0; *Sr = 0.0;
    for (int a = 0; a < 9; a++) Ss[a] = 0.0;
    FOR_NEIGHBORS(o, i, {
        if (is_flow(TYPE(o, j))) {
            double w = W(p, r), V = o->f[F_M_V][j], rj = o->f[F_DENSITY_TMP][j];
            const double *vj = &o->f[F_V_TMP][3 * j], *sj = &o->f[F_STRESS_TMP][9 * j];
            for (int a = 0; a < 3; a++) Sv[a] += V * vj[a] * w;
            *Sr += V * rj * w;
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) {
                    double e = sj[3 * a + b] + (a == b ? rj * p->g[a] * d[a] : 0.0);
                    Ss[3 * a + b] += V * e * w;
                }
        }
    });
}

/* ----------------------------------------------------------------------------------- WCSPH (wc:82-132) */
static inline double eos_wc(const OrcParams *p, double rho) {
    double v = p->stiff * (pow(rho / p->rho0, p->gamma_) - 1.0);
    return v > 0.0 ? v : 0.0;
}
/* phase: -1 = the whole one_step; 0, 1, (2) = one top-level loop only (used by the slab-decomposition tests, which
 * refresh ghost columns between the loops exactly as tisphi_b200/parallel.py does on the GPUs) */
#define PHASE(k) if (phase == -1 || phase == (k))
/* calc_repulsive_force (base:675-689) of rep particle j on i, d = x_i - x_j, added to acc */
static inline void rep_force(const OrcParams *p, const double d[3], double r, double acc[3]) {
    const double r_judge = 2.0 * p->radius;
    const double chi = (r > 0.0 && r < r_judge) ? 1.0 - r / r_judge : 0.0;
    const double gamma = r / (0.75 * p->h);
    double f = 0.0;
    if (gamma > 0 && gamma <= 2.0 / 3.0) f = 2.0 / 3.0;
    else if (gamma > 2.0 / 3.0 && gamma <= 1) f = 2 * gamma - 1.5 * gamma * gamma;
    else if (gamma > 1 && gamma < 2) f = 0.5 * (2 - gamma) * (2 - gamma);
    const double k = 0.01 * p->vsound * p->vsound * chi * f / (r * r);
    for (int a = 0; a < 3; a++) acc[a] += k * d[a];
}
static inline int has_rep(const OrcParams *p) { return p->boundary == 3 || p->boundary == 4; }

static void one_step_wc(Orc *o, int phase) {
    const OrcParams *p = &o->p;
    int64_t n = o->n;
    double *pr = o->f[F_PRESSURE], *pold = o->scratch;            /* snapshot of pressure before loop A */
    PHASE(0) {
    memcpy(pold, pr, sizeof(double) * n);
    /* loop A (wc:86-106).  In-place read of p_j emulated pointwise: j < i -> new EOS value, j > i -> old value. */
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
        int t = TYPE(o, i);
        if (is_fluid(t)) pr[i] = eos_wc(p, o->f[F_DENSITY_TMP][i]);
        if (is_bdy(t) || is_rigid(t)) {
            double Sv[3] = {0, 0, 0}, Sp = 0.0;
            const double *xi = X(o, i);
            FOR_NEIGHBORS(o, i, {
                int tj = TYPE(o, j);
                if (is_flow(tj)) {
                    double w = W(p, r), V = o->f[F_M_V][j];
                    const double *vj = &o->f[F_V_TMP][3 * j];
                    for (int a = 0; a < 3; a++) Sv[a] += V * vj[a] * w;
                    double pj;
                    if (is_fluid(tj)) pj = (p->wc_fresh || j < i) ? eos_wc(p, o->f[F_DENSITY_TMP][j]) : pold[j];
                    else pj = pold[j];
                    Sp += V * (pj + o->f[F_DENSITY_TMP][j] * p->g[1] * (xi[1] - X(o, j)[1])) * w;
                }
            });
            double f = o->f[F_CSPM_F][i];
            for (int a = 0; a < 3; a++) o->f[F_V_TMP][3 * i + a] = 2 * o->f[F_V][3 * i + a] - Sv[a] * f;
            o->f[F_DENSITY_TMP][i] = p->rho0;
            double v = Sp * f;
            pr[i] = v > 0.0 ? v : 0.0;
        }
    }
    }
    /* loop B (wc:108-126) */
    PHASE(1)
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
        if (!is_fluid(TYPE(o, i))) continue;
        const double *vi = &o->f[F_V_TMP][3 * i];
        double rhoi = o->f[F_DENSITY_TMP][i], pi = pr[i];
        double dd = 0.0, dv[3] = {0, 0, 0};
        FOR_NEIGHBORS(o, i, {
            double gc[3];
            gradWc(o, i, d, r, gc);
            const double *vj = &o->f[F_V_TMP][3 * j];
            double V = o->f[F_M_V][j];
            dd += V * (vi[0] - vj[0]) * gc[0] + V * (vi[1] - vj[1]) * gc[1] + V * (vi[2] - vj[2]) * gc[2];
        });
        FOR_NEIGHBORS(o, i, {
            double g[3];
            gradW(p, d, r, g);
            const double *vj = &o->f[F_V_TMP][3 * j];
            double V = o->f[F_M_V][j], rhoj = o->f[F_DENSITY_TMP][j];
            int tj = TYPE(o, j);
            double vx = (vi[0] - vj[0]) * d[0] + (vi[1] - vj[1]) * d[1] + (vi[2] - vj[2]) * d[2];
            double mn = vx < 0.0 ? vx : 0.0, visc = 0.0;
            if (is_fluid(tj)) visc = 2 * (p->dim + 2) * p->visc * V * mn / (r * r + 0.01 * p->h * p->h);
            else if (is_bdy(tj) || is_rigid(tj)) visc = 2 * (p->dim + 2) * p->visc * V * p->rho0 / rhoi * mn / (r * r + 0.01 * p->h * p->h);
            double pres = -p->rho0 * V * (pi / (rhoi * rhoi) + pr[j] / (rhoj * rhoj));
            for (int a = 0; a < 3; a++) dv[a] += visc * g[a] + pres * g[a];
        });
        if (has_rep(p)) FOR_NEIGHBORS(o, i, { if (TYPE(o, j) == -2) rep_force(p, d, r, dv); });      /* wc:119-121 */
        o->f[F_D_DENSITY][i] = dd * rhoi;
        for (int a = 0; a < 3; a++) o->f[F_D_VEL][3 * i + a] = dv[a] + p->g[a];
    }
}

/* ------------------------------------------------------------------- soil momentum (muI:38-46, dp:156-165) */
/* react != 0: the same term is subtracted from d_vel of a DYNAMIC rigid neighbour (muI:45-46, dp:164-165; the caller
 * runs the loop in index order then, like the serial reference) */
static void soil_momentum(Orc *o, int64_t i, double out[3], int react) {
    const double *si = &o->f[F_STRESS_TMP][9 * i];
    double rhoi = o->f[F_DENSITY_TMP][i];
    out[0] = out[1] = out[2] = 0.0;
    FOR_NEIGHBORS(o, i, {
        double g[3], t3[3];
        gradWc(o, i, d, r, g);
        const double *sj = &o->f[F_STRESS_TMP][9 * j];
        double rhoj = o->f[F_DENSITY_TMP][j], c = o->f[F_M_V][j] * rhoj;
        for (int a = 0; a < 3; a++) {
            double s = 0.0;
            for (int b = 0; b < 3; b++) s += (c * (sj[3 * a + b] / (rhoj * rhoj) + si[3 * a + b] / (rhoi * rhoi))) * g[b];
            t3[a] = s;
            out[a] += s;
        }
        if (react && RIGID_DYN(o, j)) for (int a = 0; a < 3; a++) o->f[F_D_VEL][3 * j + a] -= t3[a];
    });
}
static inline void viscous_damping(const OrcParams *p, double rho, const double *v, double out[3]) {   /* base:713-715 */
    double c = -5e-5 * sqrt(p->E / (rho * p->h * p->h));
    for (int a = 0; a < 3; a++) out[a] = c * v[a];
}
static inline double dev_component(const double *t) {    /* type_define.py:26-28 */
    double s = 0.0;
    for (int a = 0; a < 9; a++) s += t[a] * t[a];
    return sqrt(s * 2 / 3);
}

/* ------------------------------------------------------------------------------------ mu(I) (muI:62-156) */
static void one_step_mui(Orc *o, int phase) {
    const OrcParams *p = &o->p;
    int64_t n = o->n;
    PHASE(0)
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {                    /* loop 1 (muI:67-92) */
        if (!is_soil(TYPE(o, i))) continue;
        double vg[9], dd;
        sum_vgrad_ddens(o, i, vg, &dd);
        memcpy(&o->f[F_V_GRAD][9 * i], vg, 72);
        o->f[F_D_DENSITY][i] = dd * o->f[F_DENSITY_TMP][i];
        double pv = p->vsound * p->vsound * (o->f[F_DENSITY][i] - p->rho0);    /* rho, not rho~ (H17) */
        if (pv < 0.0) pv = 0.0;
        o->f[F_PRESSURE][i] = pv;
        double sr[9], s2 = 0.0, tr;
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) sr[3 * a + b] = 0.5 * (vg[3 * a + b] + vg[3 * b + a]);
        for (int a = 0; a < 9; a++) s2 += sr[a] * sr[a];
        double dbdot = sqrt(0.5 * s2) + p->eps;
        double coef = 0.0 + (p->coh + pv * p->mu) / dbdot;            /* eta_0 = 0 (muI:19) */
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) o->f[F_STRESS_TMP][9 * i + 3 * a + b] = coef * sr[3 * a + b] - (a == b ? pv : 0.0);
        tr = sr[0] + sr[4] + sr[8];
        double se[9];
        for (int a = 0; a < 9; a++) se[a] = sr[a];
        se[0] -= tr / 3.0; se[4] -= tr / 3.0; se[8] -= tr / 3.0;
        o->f[F_D_STRAIN_EQU][i] = dev_component(se);
    }
    PHASE(1)
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {                    /* loop 2 (muI:95-109) */
        int t = TYPE(o, i);
        if (!(is_bdy(t) || is_rigid(t))) continue;
        double Sv[3], Sr, Ss[9], f = o->f[F_CSPM_F][i];
        wall_sums(o, i, Sv, &Sr, Ss);
        for (int a = 0; a < 3; a++) o->f[F_V_TMP][3 * i + a] = 2 * o->f[F_V][3 * i + a] - Sv[a] * f;
        double rt = Sr * f;
        o->f[F_DENSITY_TMP][i] = rt > p->rho0 ? rt : p->rho0;
        for (int a = 0; a < 9; a++) o->f[F_STRESS_TMP][9 * i + a] = Ss[a] * f;
        if (RIGID_DYN(o, i)) for (int a = 0; a < 3; a++) o->f[F_D_VEL][3 * i + a] = p->g[a];      /* muI:111-112 */
    }
    PHASE(2)
#pragma omp parallel for schedule(dynamic, 256) if (o->n_dyn == 0)
    for (int64_t i = 0; i < n; i++) {                    /* loop 3 (muI:115-128) */
        if (!is_soil(TYPE(o, i))) continue;
        double dv[3], Fd[3];
        soil_momentum(o, i, dv, o->n_dyn > 0);
        if (has_rep(p)) FOR_NEIGHBORS(o, i, { if (TYPE(o, j) == -2) rep_force(p, d, r, dv); });      /* muI:121-123 */
        viscous_damping(p, o->f[F_DENSITY_TMP][i], &o->f[F_V_TMP][3 * i], Fd);
        for (int a = 0; a < 3; a++) o->f[F_D_VEL][3 * i + a] = dv[a] + p->g[a] + Fd[a];
    }
}

/* ------------------------------------------------------------------------- Drucker-Prager (dp:41-296) */
static void from_stress(const OrcParams *p, const double *s, double dev[9], double *I1, double *sJ2, double *f) {
    double tr = s[0] + s[4] + s[8], s2 = 0.0;
    for (int a = 0; a < 9; a++) dev[a] = s[a];
    dev[0] -= tr / 3.0; dev[4] -= tr / 3.0; dev[8] -= tr / 3.0;
    for (int a = 0; a < 9; a++) s2 += dev[a] * dev[a];
    *I1 = tr; *sJ2 = sqrt(0.5 * s2); *f = *sJ2 + p->alpha * tr - p->kc;
}
static void adapt_stress(const OrcParams *p, double *s) {     /* dp:73-96 */
    double dev[9], I1, sJ2, f;
    from_stress(p, s, dev, &I1, &sJ2, &f);
    if (f > p->eps_f) {
        if (f > sJ2) {
            double tmp = (I1 - p->kc / p->alpha) / 3.0;
            s[0] -= tmp; s[4] -= tmp; s[8] -= tmp;
        }
        from_stress(p, s, dev, &I1, &sJ2, &f);
        double rr = (-I1 * p->alpha + p->kc) / sJ2;
        for (int a = 0; a < 9; a++) s[a] = rr * dev[a];
        s[0] += 1.0 * I1 / 3.0; s[4] += 1.0 * I1 / 3.0; s[8] += 1.0 * I1 / 3.0;
    }
}
static int flag_dp(const OrcParams *p, const double *s) {      /* dp:98-109 */
    double dev[9], I1, sJ2, f;
    from_stress(p, s, dev, &I1, &sJ2, &f);
    if (f < -p->eps_f) return 0;
    if (f > p->eps_f) return (f >= sJ2) ? 3 : 2;
    return 1;
}
static void bui2008(const OrcParams *p, const double *st, const double *vg, double ds[9], double *dse, double *dsep) {  /* dp:171-208 */
    double dev[9], I1, sJ2, f, sr[9], sp[9], J[9], se[9], te[9], tg[9] = {0}, lam = 0.0;
    from_stress(p, st, dev, &I1, &sJ2, &f);
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
            sr[3 * a + b] = 0.5 * (vg[3 * a + b] + vg[3 * b + a]);
            sp[3 * a + b] = 0.5 * (vg[3 * a + b] - vg[3 * b + a]);
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            J[3 * i + j] = st[3 * i + 0] * sp[3 * j + 0] + st[3 * i + 1] * sp[3 * j + 1] + st[3 * i + 2] * sp[3 * j + 2] +
                           st[0 + j] * sp[3 * i + 0] + st[3 + j] * sp[3 * i + 1] + st[6 + j] * sp[3 * i + 2];
    double tr = sr[0] + sr[4] + sr[8];
    for (int a = 0; a < 9; a++) se[a] = sr[a];
    se[0] -= tr / 3.0; se[4] -= tr / 3.0; se[8] -= tr / 3.0;
    for (int a = 0; a < 9; a++) te[a] = 2.0 * p->G * se[a];
    te[0] += p->K * tr; te[4] += p->K * tr; te[8] += p->K * tr;
    int plastic = (f >= -p->eps_f && sJ2 > p->eps);
    if (plastic) {
        double ss = 0.0;
        for (int a = 0; a < 9; a++) ss += dev[a] * sr[a];
        lam = (3.0 * p->alpha * p->K * tr + (p->G / sJ2) * ss) / (27.0 * p->alpha * p->K * sin(p->dila) + p->G);
        for (int a = 0; a < 9; a++) tg[a] = lam * (p->G / sJ2 * dev[a]);
        tg[0] = lam * (9.0 * p->K * sin(p->dila) + p->G / sJ2 * dev[0]);
        tg[4] = lam * (9.0 * p->K * sin(p->dila) + p->G / sJ2 * dev[4]);
        tg[8] = lam * (9.0 * p->K * sin(p->dila) + p->G / sJ2 * dev[8]);
    }
    for (int a = 0; a < 9; a++) ds[a] = J[a] + te[a] - tg[a];
    *dse = dev_component(se);
    *dsep = 0.0;
    if (plastic) {
        double gp = sJ2 + 3 * I1 * sin(p->dila), ep[9], trp;
        for (int a = 0; a < 9; a++) ep[a] = (fabs(ds[a]) > p->eps ? gp / ds[a] : 0.0) * lam;   /* pti.g_p is never written: 0 */
        trp = ep[0] + ep[4] + ep[8];
        ep[0] -= trp / 3.0; ep[4] -= trp / 3.0; ep[8] -= trp / 3.0;
        *dsep = dev_component(ep);
    }
}
static void one_step_dp(Orc *o, int phase) {
    const OrcParams *p = &o->p;
    int64_t n = o->n;
    PHASE(0)
    for (int64_t i = 0; i < n; i++)                      /* loop 1 (dp:215-217) */
        if (is_soil(TYPE(o, i))) adapt_stress(p, &o->f[F_STRESS_TMP][9 * i]);
    PHASE(1)
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {                    /* loop 2 (dp:220-231) */
        int t = TYPE(o, i);
        if (!(is_bdy(t) || is_rigid(t))) continue;
        double Sv[3], Sr, Ss[9], f = o->f[F_CSPM_F][i];
        wall_sums(o, i, Sv, &Sr, Ss);
        for (int a = 0; a < 3; a++) o->f[F_V_TMP][3 * i + a] = 2 * o->f[F_V][3 * i + a] - Sv[a] * f;
        o->f[F_DENSITY_TMP][i] = p->rho0;
        for (int a = 0; a < 9; a++) o->f[F_STRESS_TMP][9 * i + a] = Ss[a] * f;
        if (RIGID_DYN(o, i)) for (int a = 0; a < 3; a++) o->f[F_D_VEL][3 * i + a] = p->g[a];      /* dp:233-234 */
    }
    PHASE(2)
#pragma omp parallel for schedule(dynamic, 256) if (o->n_dyn == 0)
    for (int64_t i = 0; i < n; i++) {                    /* loop 3 (dp:237-270) */
        if (!is_soil(TYPE(o, i))) continue;
        double vg[9], dd, dv[3], Fd[3];
        sum_vgrad_ddens(o, i, vg, &dd);
        memcpy(&o->f[F_V_GRAD][9 * i], vg, 72);
        o->f[F_D_DENSITY][i] = dd * o->f[F_DENSITY_TMP][i];
        bui2008(p, &o->f[F_STRESS_TMP][9 * i], vg, &o->f[F_D_STRESS][9 * i], &o->f[F_D_STRAIN_EQU][i], &o->f[F_D_STRAIN_EQU_P][i]);
        soil_momentum(o, i, dv, o->n_dyn > 0);
        viscous_damping(p, o->f[F_DENSITY_TMP][i], &o->f[F_V_TMP][3 * i], Fd);
        for (int a = 0; a < 3; a++) o->f[F_D_VEL][3 * i + a] = dv[a] + p->g[a] + Fd[a];
    }
}

void orc_one_step_phase(Orc *o, int phase) {
    if (o->p.solver == 1) one_step_wc(o, phase);
    else if (o->p.solver == 2) one_step_mui(o, phase);
    else one_step_dp(o, phase);
}
void orc_one_step(Orc *o) { orc_one_step_phase(o, -1); }

/* DP constructor's init_stress (base:249-260, dp:35) */
void orc_init_stress(Orc *o) {
    const OrcParams *p = &o->p;
    double ymax = -INFINITY;
    for (int64_t i = 0; i < o->n; i++) if (is_soil(TYPE(o, i)) && X(o, i)[1] > ymax) ymax = X(o, i)[1];
    double K0 = 1.0 - sin(p->fric);
    for (int64_t i = 0; i < o->n; i++) if (is_soil(TYPE(o, i))) {
        double ver = p->rho0 * p->g[1] * (ymax - X(o, i)[1]);
        o->f[F_STRESS][9 * i + 0] = K0 * ver; o->f[F_STRESS][9 * i + 4] = ver; o->f[F_STRESS][9 * i + 8] = K0 * ver;
    }
}

/* ------------------------------------------------------------------------- advect_pos (base:228-238) */
void orc_advect_pos(Orc *o) {
    const OrcParams *p = &o->p;
    int64_t n = o->n;
    double dt = p->dt;
    if (!p->xsph) {
        for (int64_t i = 0; i < n; i++) if (is_real(TYPE(o, i))) for (int a = 0; a < 3; a++) X(o, i)[a] += dt * o->f[F_V][3 * i + a];
        return;
    }
    if (p->serial) {                                     /* in-place, index order: neighbours j < i have already moved */
        for (int64_t i = 0; i < n; i++) {
            if (!is_real(TYPE(o, i))) continue;
            double s[3] = {0, 0, 0};
            const double *vi = &o->f[F_V][3 * i];
            if (o->ia[I_IS_DYNAMIC][i])                       /* base:234: XSPH only moves dynamic particles */
            FOR_NEIGHBORS(o, i, {
                if (TYPE(o, j) == TYPE(o, i)) {
                    double w = W(p, r), V = o->f[F_M_V][j];
                    for (int a = 0; a < 3; a++) s[a] += V * (o->f[F_V][3 * j + a] - vi[a]) * w;
                }
            });
            for (int a = 0; a < 3; a++) X(o, i)[a] += dt * (vi[a] + 0.5 * s[a]);
        }
        return;
    }
    double *xn = o->scratch;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
        for (int a = 0; a < 3; a++) xn[3 * i + a] = X(o, i)[a];
        if (!is_real(TYPE(o, i))) continue;
        double s[3] = {0, 0, 0};
        const double *vi = &o->f[F_V][3 * i];
        if (o->ia[I_IS_DYNAMIC][i])
        FOR_NEIGHBORS(o, i, {
            if (TYPE(o, j) == TYPE(o, i)) {
                double w = W(p, r), V = o->f[F_M_V][j];
                for (int a = 0; a < 3; a++) s[a] += V * (o->f[F_V][3 * j + a] - vi[a]) * w;
            }
        });
        for (int a = 0; a < 3; a++) xn[3 * i + a] = X(o, i)[a] + dt * (vi[a] + 0.5 * s[a]);
    }
    memcpy(o->f[F_X], xn, sizeof(double) * 3 * n);
}

/* --------------------------------------------------- advect_something (wc:129-132, muI:134-156, dp:276-296) */
static inline void chk_density(Orc *o, int64_t i) {      /* base:214-221 */
    if (o->f[F_DENSITY][i] < o->p.rho0) o->f[F_DENSITY][i] = o->p.rho0;
    o->f[F_M_V][i] = o->f[F_MASS][i] / o->f[F_DENSITY][i];
}
void orc_post_step(Orc *o) {
    const OrcParams *p = &o->p;
    int64_t n = o->n;
    if (p->solver == 1) {
        for (int64_t i = 0; i < n; i++) if (is_fluid(TYPE(o, i))) chk_density(o, i);
    } else if (p->solver == 3) {
        for (int64_t i = 0; i < n; i++) if (is_soil(TYPE(o, i))) {
            chk_density(o, i);
            o->ia[I_FLAG_RETMAP][i] = flag_dp(p, &o->f[F_STRESS][9 * i]);
            adapt_stress(p, &o->f[F_STRESS][9 * i]);
            o->f[F_STRAIN_EQU][i] += p->dt * o->f[F_D_STRAIN_EQU][i];
            o->f[F_STRAIN_EQU_P][i] += p->dt * o->f[F_D_STRAIN_EQU_P][i];
        }
    } else {
        /* muI:134-156.  The loop interleaves, per i: clamp, strain, stress_i = stress_tmp_i, then the in-place
         * Shepard sum over stress_j.  Uses post-advect positions on the pre-move grid (H15). */
        double *S = o->f[F_STRESS], *St = o->f[F_STRESS_TMP];
        if (p->serial) {
            for (int64_t i = 0; i < n; i++) {
                if (!is_soil(TYPE(o, i))) continue;
                chk_density(o, i);
                o->f[F_STRAIN_EQU][i] += p->dt * o->f[F_D_STRAIN_EQU][i];
                memcpy(&S[9 * i], &St[9 * i], 72);
                double acc[9] = {0};
                FOR_NEIGHBORS(o, i, {
                    if (TYPE(o, j) == TYPE(o, i)) {
                        double w = W(p, r), V = o->f[F_M_V][j];
                        for (int a = 0; a < 9; a++) acc[a] += V * S[9 * j + a] * w;
                    }
                });
                for (int a = 0; a < 9; a++) S[9 * i + a] = acc[a] * o->f[F_CSPM_F][i];
            }
        } else {
            for (int64_t i = 0; i < n; i++) if (is_soil(TYPE(o, i))) {
                chk_density(o, i);
                o->f[F_STRAIN_EQU][i] += p->dt * o->f[F_D_STRAIN_EQU][i];
            }
#pragma omp parallel for schedule(dynamic, 256)
            for (int64_t i = 0; i < n; i++) {
                if (!is_soil(TYPE(o, i))) continue;
                double acc[9] = {0};
                FOR_NEIGHBORS(o, i, {
                    if (TYPE(o, j) == TYPE(o, i)) {
                        double w = W(p, r), V = o->f[F_M_V][j];
                        for (int a = 0; a < 9; a++) acc[a] += V * St[9 * j + a] * w;
                    }
                });
                for (int a = 0; a < 9; a++) S[9 * i + a] = acc[a] * o->f[F_CSPM_F][i];
            }
        }
    }
}

/* ---------------------------------------------------------------------------- SPHBase.step (base:41-61) */
/* ------------------------------------------------------------------- dynamic rigid bodies (base:467-518) */
static void rigid_cm(const Orc *o, int obj, double cm[3]) {                 /* calc_cm, base:501-510 */
    double sm = 0.0;
    cm[0] = cm[1] = cm[2] = 0.0;
    for (int64_t i = 0; i < o->n; i++)
        if (RIGID_DYN(o, i) && o->ia[I_OBJ_ID][i] == obj) {
            const double m = o->f[F_MASS][i];
            for (int a = 0; a < 3; a++) cm[a] += m * X(o, i)[a];
            sm += m;
        }
    for (int a = 0; a < 3; a++) cm[a] /= sm;
}
/* init_rigid_body (base:467-470), called by the solver constructors (base:28): rest centre of mass per dynamic object */
void orc_init_rigid_body(Orc *o) {
    o->n_dyn = 0;
    for (int obj = 0; obj < MAX_OBJ; obj++) {
        int has = 0;
        for (int64_t i = 0; i < o->n && !has; i++) has = RIGID_DYN(o, i) && o->ia[I_OBJ_ID][i] == obj;
        if (!has) continue;
        o->dyn_obj[o->n_dyn++] = obj;
        rigid_cm(o, obj, o->rest_cm[obj]);
    }
}
/* Rotation of the polar decomposition A = R S as Taichi's polar_decompose3d gives it: U, sig, V = svd(A) with U and V
 * PROPER rotations (the sign goes into the last singular value), R = U V^T.  Here: Jacobi eigen-decomposition of A^T A
 * for V, u_k = A v_k / sigma_k, missing columns completed by cross products.  (Third-party arithmetic of the reference,
 * any converged SVD agrees to ~1e-15.) */
static void polar_rotation(const double A[9], double R[9]) {
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
            for (int k = 0; k < 3; k++) {                    /* B <- J^T B J, V <- V J */
                const double bkp = B[3 * k + pi], bkq = B[3 * k + qi];
                B[3 * k + pi] = cs * bkp - sn * bkq; B[3 * k + qi] = sn * bkp + cs * bkq;
            }
            for (int k = 0; k < 3; k++) {
                const double bpk = B[3 * pi + k], bqk = B[3 * qi + k];
                B[3 * pi + k] = cs * bpk - sn * bqk; B[3 * qi + k] = sn * bpk + cs * bqk;
            }
            for (int k = 0; k < 3; k++) {
                const double vkp = V[3 * k + pi], vkq = V[3 * k + qi];
                V[3 * k + pi] = cs * vkp - sn * vkq; V[3 * k + qi] = sn * vkp + cs * vkq;
            }
        }
    }
    int ord[3] = {0, 1, 2};                              /* eigenvalues descending */
    for (int a = 0; a < 2; a++)
        for (int b = a + 1; b < 3; b++) if (B[4 * ord[b]] > B[4 * ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double v[3][3], u[3][3], sig[3];
    for (int k = 0; k < 3; k++) {
        for (int a = 0; a < 3; a++) v[k][a] = V[3 * a + ord[k]];
        sig[k] = sqrt(B[4 * ord[k]] > 0.0 ? B[4 * ord[k]] : 0.0);
    }
    /* V proper: v2 = v0 x v1 */
    v[2][0] = v[0][1] * v[1][2] - v[0][2] * v[1][1]; v[2][1] = v[0][2] * v[1][0] - v[0][0] * v[1][2]; v[2][2] = v[0][0] * v[1][1] - v[0][1] * v[1][0];
    const double tol = 1e-12 * (sig[0] > 0 ? sig[0] : 1.0);
    int rank = 0;
    for (int k = 0; k < 2; k++) {
        if (sig[k] <= tol) break;
        for (int a = 0; a < 3; a++) { double t = 0; for (int b = 0; b < 3; b++) t += A[3 * a + b] * v[k][b]; u[k][a] = t / sig[k]; }
        rank++;
    }
    if (rank == 0) { for (int a = 0; a < 9; a++) R[a] = (a % 4 == 0) ? 1.0 : 0.0; return; }   /* A == 0: identity (base:491-492) */
    if (rank == 1) {                                     /* a line of particles: any unit vector normal to u0 */
        const int m = fabs(u[0][0]) < fabs(u[0][1]) ? (fabs(u[0][0]) < fabs(u[0][2]) ? 0 : 2) : (fabs(u[0][1]) < fabs(u[0][2]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[m] = 1.0;
        double w[3] = {u[0][1] * e[2] - u[0][2] * e[1], u[0][2] * e[0] - u[0][0] * e[2], u[0][0] * e[1] - u[0][1] * e[0]};
        const double nw = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        for (int a = 0; a < 3; a++) u[1][a] = w[a] / nw;
    }
    u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1]; u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2]; u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) R[3 * a + b] = u[0][a] * v[0][b] + u[1][a] * v[1][b] + u[2][a] * v[2][b];
}
void orc_polar_rotation(const double *A, double *R) { polar_rotation(A, R); }      /* (exported for its unit test) */
/* solve_rigid_body / solve_constraints (base:472-499): shape matching of every dynamic rigid object */
void orc_solve_rigid_body(Orc *o) {
    for (int q = 0; q < o->n_dyn; q++) {
        const int obj = o->dyn_obj[q];
        double cm[3], A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, R[9];
        rigid_cm(o, obj, cm);
        for (int64_t i = 0; i < o->n; i++)
            if (RIGID_DYN(o, i) && o->ia[I_OBJ_ID][i] == obj) {
                const double w = o->f[F_M_V][i] * o->f[F_DENSITY][i];
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 3; b++)
                        A[3 * a + b] += w * (X(o, i)[a] - cm[a]) * (o->f[F_X0][3 * i + b] - o->rest_cm[obj][b]);
            }
        polar_rotation(A, R);
        for (int64_t i = 0; i < o->n; i++)
            if (RIGID_DYN(o, i) && o->ia[I_OBJ_ID][i] == obj)
                for (int a = 0; a < 3; a++) {
                    double g = cm[a];
                    for (int b = 0; b < 3; b++) g += R[3 * a + b] * (o->f[F_X0][3 * i + b] - o->rest_cm[obj][b]);
                    X(o, i)[a] += (g - X(o, i)[a]) * 1.0;
                }
    }
}

/* enforce_boundary (base:525-601): dynamic rigid particles, and with boundary == 1 flow particles too, are put back
 * inside the domain box (no lid) and lose (1 + c_f) of their normal velocity. */
void orc_enforce_boundary(Orc *o) {
    const OrcParams *p = &o->p;
    if (p->boundary != 1 && o->n_dyn == 0) return;
    const double rr = p->radius - p->eps;
    for (int64_t i = 0; i < o->n; i++) {
        if (!(RIGID_DYN(o, i) || (p->boundary == 1 && is_flow(TYPE(o, i))))) continue;
        double *x = X(o, i), *v = &o->f[F_V][3 * i], nrm[3] = {0, 0, 0};
        const double pos[3] = {x[0], x[1], x[2]};
        if (pos[0] > p->dend[0] - rr) { nrm[0] += 1.0; x[0] = p->dend[0] - rr; }
        if (pos[0] <= p->dstart[0] + rr) { nrm[0] += -1.0; x[0] = p->dstart[0] + rr; }
        if (pos[1] <= p->dstart[1] + rr) { nrm[1] += -1.0; x[1] = p->dstart[1] + rr; }
        if (p->dim == 3) {
            if (pos[2] > p->dend[2] - rr) { nrm[2] += 1.0; x[2] = p->dend[2] - rr; }
            if (pos[2] <= p->dstart[2] + rr) { nrm[2] += -1.0; x[2] = p->dstart[2] + rr; }
        }
        const double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
        if (len > p->eps) {                                /* simulate_collisions, c_f = 0.3 (base:598-601) */
            const double u[3] = {nrm[0] / len, nrm[1] / len, nrm[2] / len};
            const double vn = v[0] * u[0] + v[1] * u[1] + v[2] * u[2];
            for (int a = 0; a < 3; a++) v[a] -= (1.0 + 0.3) * vn * u[a];
        }
    }
}

int64_t orc_step(Orc *o) {
    int64_t bad = orc_grid_build(o);
    orc_calc_kernel_corr(o);
    orc_init_real2tmp(o);
    switch (o->p.ti) {
    case 1: orc_one_step(o); orc_advect(o, 0, 0); break;
    case 2: orc_one_step(o); orc_advect(o, 1, 0); orc_one_step(o); orc_advect(o, 0, 0); break;
    case 4: {
        static const int m[4] = {1, 2, 2, 1};
        orc_advect(o, 3, 0);
        for (int s = 0; s < 4; s++) { orc_one_step(o); orc_advect(o, 4, m[s]); if (s < 3) orc_advect(o, 2, 0); }
        orc_advect(o, 5, 0);
    } break;
    default: return -1;                                  /* timeIntegration 3 is broken in the reference (H18) */
    }
    orc_advect_pos(o);
    orc_post_step(o);
    orc_solve_rigid_body(o);
    orc_enforce_boundary(o);
    return bad;
}

/* -------------------------------------------------------------- stand-alone sweeps (config C5 and counts) */
/* neighbour count with the float64 predicate of ps:268 */
void orc_neighbor_count(Orc *o, int32_t *out) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < o->n; i++) {
        int32_t c = 0;
        FOR_NEIGHBORS(o, i, { (void)d; c++; });
        out[i] = c;
    }
}
/* float32 predicate: positions rounded to float, r2 = fl(fl(dx*dx + dy*dy) + dz*dz), sqrtf(r2) < (float)support.
 * Cell lookup still uses the float64 positions (the engine always bins in float64). */
void orc_neighbor_count_f32(Orc *o, int32_t *out) {
    const OrcParams *p = &o->p;
    const float sup = (float)p->support;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < o->n; i++) {
        int cc[3];
        pos_to_index(p, X(o, i), cc);
        float xi[3] = {(float)X(o, i)[0], (float)X(o, i)[1], (float)X(o, i)[2]};
        int32_t c = 0;
        for (int ox = -1; ox <= 1; ox++) for (int oy = -1; oy <= 1; oy++) for (int oz = -1; oz <= 1; oz++) {
            if (p->dim == 2 && oz != 0) continue;
            int cl[3] = {cc[0] + ox, cc[1] + oy, cc[2] + oz};
            if (cl[0] < 0 || cl[0] >= p->gn[0] || cl[1] < 0 || cl[1] >= p->gn[1] || cl[2] < 0 || cl[2] >= p->gn[2]) continue;
            int64_t g = flatten(p, cl), jb = g > 0 ? o->cell_end[g - 1] : 0, je = o->cell_end[g];
            for (int64_t j = jb; j < je; j++) {
                if (j == i) continue;
                float dx = xi[0] - (float)X(o, j)[0], dy = xi[1] - (float)X(o, j)[1], dz = xi[2] - (float)X(o, j)[2];
                float r2 = dx * dx + dy * dy;
                r2 = r2 + dz * dz;
                if (sqrtf(r2) < sup) c++;
            }
        }
        out[i] = c;
    }
}

/* float32 predicate of the MIXED engine: coordinates local to the cell each particle is stored in (grid_ids),
 * xs = (float)(x - (vstart + cell*gs)) with unfused float64 ops.  A pair is evaluated from the side of the LOWER cell
 * id in the frame of the higher cell: lo in cell A, hi in cell B > A, s = (float)(B - A) * (float)gs,
 * d = (xs_lo - s) - xs_hi (same cell: d = xs_i - xs_j), r2 = fma(dz,dz, fma(dy,dy, fl(dx*dx))), sqrtf(r2) < (float)support.
 * The relation is exactly symmetric, which the CUDA mask kernel exploits (one evaluation per cell pair + transpose). */
static inline void unflatten(const OrcParams *p, int64_t g, int c[3]) {
    int64_t nyz = (int64_t)p->gn[1] * p->gn[2];
    c[0] = (int)(g / nyz);
    int64_t r = g - c[0] * nyz;
    c[1] = (int)(r / p->gn[2]);
    c[2] = (int)(r - (int64_t)c[1] * p->gn[2]);
}
static inline void local_xs(const Orc *o, int64_t i, const int sc[3], float xs[3]) {
    const OrcParams *p = &o->p;
    for (int a = 0; a < 3; a++) {
        double org = p->vstart[a] + (double)sc[a] * p->grid_size;
        xs[a] = (float)(X(o, i)[a] - org);
    }
}
void orc_neighbor_count_f32local(Orc *o, int32_t *out) {
    const OrcParams *p = &o->p;
    const float sup = (float)p->support, gsf = (float)p->grid_size;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < o->n; i++) {
        int cc[3], sc[3];
        pos_to_index(p, X(o, i), cc);
        unflatten(p, o->ia[I_GRID_IDS][i], sc);
        float xi[3];
        local_xs(o, i, sc, xi);
        int32_t c = 0;
        for (int ox = -1; ox <= 1; ox++) for (int oy = -1; oy <= 1; oy++) for (int oz = -1; oz <= 1; oz++) {
            if (p->dim == 2 && oz != 0) continue;
            int cl[3] = {cc[0] + ox, cc[1] + oy, cc[2] + oz};
            if (cl[0] < 0 || cl[0] >= p->gn[0] || cl[1] < 0 || cl[1] >= p->gn[1] || cl[2] < 0 || cl[2] >= p->gn[2]) continue;
            float sh[3];
            for (int a = 0; a < 3; a++) sh[a] = (float)(cl[a] - sc[a]) * gsf;
            int64_t g = flatten(p, cl), jb = g > 0 ? o->cell_end[g - 1] : 0, je = o->cell_end[g];
            for (int64_t j = jb; j < je; j++) {
                if (j == i) continue;
                int sj[3];
                float xj[3];
                unflatten(p, o->ia[I_GRID_IDS][j], sj);
                local_xs(o, j, sj, xj);
                float dx, dy, dz;
                if (g >= o->ia[I_GRID_IDS][i]) {
                    float ex = xi[0] - sh[0], ey = xi[1] - sh[1], ez = xi[2] - sh[2];
                    dx = ex - xj[0]; dy = ey - xj[1]; dz = ez - xj[2];
                } else {                           /* j's cell precedes i's: evaluate from j's side (shift negated) */
                    float ex = xj[0] + sh[0], ey = xj[1] + sh[1], ez = xj[2] + sh[2];
                    dx = ex - xi[0]; dy = ey - xi[1]; dz = ez - xi[2];
                }
                float xx = dx * dx;
                float r2 = fmaf(dz, dz, fmaf(dy, dy, xx));
                if (sqrtf(r2) < sup) c++;
            }
        }
        out[i] = c;
    }
}
/* C5: rho_i = sum_j mass_j W_ij (wc:30-31 calc_density_task; self excluded like every for_all_neighbors sum) */
void orc_density_sum(Orc *o, double *out) {
    const OrcParams *p = &o->p;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < o->n; i++) {
        double s = 0.0;
        FOR_NEIGHBORS(o, i, { (void)d; s += o->f[F_MASS][j] * W(p, r); });
        out[i] = s;
    }
}
/* same sum but with positions rounded to float32 first (what a float32 engine sees), accumulated in float64 */
void orc_density_sum_f32pos(Orc *o, double *out) {
    const OrcParams *p = &o->p;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < o->n; i++) {
        int cc[3];
        pos_to_index(p, X(o, i), cc);
        float xi[3] = {(float)X(o, i)[0], (float)X(o, i)[1], (float)X(o, i)[2]};
        const float sup = (float)p->support;
        double s = 0.0;
        for (int ox = -1; ox <= 1; ox++) for (int oy = -1; oy <= 1; oy++) for (int oz = -1; oz <= 1; oz++) {
            if (p->dim == 2 && oz != 0) continue;
            int cl[3] = {cc[0] + ox, cc[1] + oy, cc[2] + oz};
            if (cl[0] < 0 || cl[0] >= p->gn[0] || cl[1] < 0 || cl[1] >= p->gn[1] || cl[2] < 0 || cl[2] >= p->gn[2]) continue;
            int64_t g = flatten(p, cl), jb = g > 0 ? o->cell_end[g - 1] : 0, je = o->cell_end[g];
            for (int64_t j = jb; j < je; j++) {
                if (j == i) continue;
                float dx = xi[0] - (float)X(o, j)[0], dy = xi[1] - (float)X(o, j)[1], dz = xi[2] - (float)X(o, j)[2];
                float r2 = dx * dx + dy * dy;
                r2 = r2 + dz * dz;
                if (sqrtf(r2) < sup) s += o->f[F_MASS][j] * W(p, sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz));
            }
        }
        out[i] = s;
    }
}
/* the squared-distance threshold equivalent to sqrtf(r2) < sup:  smallest float T with sqrtf(T) >= sup */
float orc_r2_threshold_f32(float sup) {
    float t = sup * sup;
    while (sqrtf(t) >= sup) t = nextafterf(t, 0.0f);
    while (sqrtf(t) < sup) t = nextafterf(t, INFINITY);
    return t;
}
double orc_r2_threshold_f64(double sup) {
    double t = sup * sup;
    while (sqrt(t) >= sup) t = nextafter(t, 0.0);
    while (sqrt(t) < sup) t = nextafter(t, INFINITY);
    return t;
}
