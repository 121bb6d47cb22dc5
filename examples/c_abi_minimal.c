/* Minimal host program against the C ABI of libtisphi_b200.so (include/tisphi_b200.h): a 3D block of water between
 * dummy walls is uploaded from host arrays, stepped with "LF" WCSPH and read back.  No Python, no torch: the arena is a
 * plain cudaMalloc.  This is what a non-Python host (or the cgo / JNI stub of another front end) would do.
 *
 *   nvcc -o c_abi_minimal examples/c_abi_minimal.c -Iinclude -Ltisphi_b200 -ltisphi_b200 -Xlinker -rpath=$PWD/tisphi_b200
 *   ./c_abi_minimal [state.bin]     (state.bin: n, then id0[n], x[3n], v[4n] float, density[n] -- what tests/ compare
 *                                    with the CPU oracle on the same particles)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>
#include "tisphi_b200.h"

#define CHECK(call) do { int rc_ = (call); if (rc_) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, sph_last_error(ctx)); return 1; } } while (0)

int main(int argc, char **argv) {
    /* scene constants exactly as ParticleSystem.__init__ derives them (eng/particle_system.py:32-59) */
    const double d = 0.02, h = 1.5 * d, support = 2.0 * h, gs = ceil(2.0 * 1.5) * d;
    const int nx = 20, ny = 15, nz = 10, layers = 3;              /* fluid block and wall thickness in particles */
    const double size[3] = {1.0, 0.6, 0.4};                       /* domain */
    SphParams p;
    memset(&p, 0, sizeof p);
    if (sph_params_size() != (int64_t)sizeof p) { fprintf(stderr, "header / library mismatch\n"); return 1; }
    p.dim = 3; p.kernel = 1; p.kcorr = 0; p.ti = 2; p.xsph = 0; p.solver = SPH_SOLVER_WC; p.precision = SPH_PREC_MIXED;
    p.fast = 1;
    for (int a = 0; a < 3; a++) { p.vstart[a] = -gs; p.gn[a] = (int)ceil((size[a] + 2 * gs) / gs); }
    p.h = h; p.support = support; p.grid_size = gs; p.m_V0 = d * d * d; p.eps = 1e-8;
    p.g[1] = -9.81; p.rho0 = 1000.0; p.visc = 0.01; p.stiff = 5e5; p.gamma_ = 7.0; p.vsound = 60.0;
    p.boundary = 2; p.radius = d / 2;                             /* dummy-particle walls (ps:18) */
    for (int a = 0; a < 3; a++) { p.dstart[a] = 0.0; p.dend[a] = size[a]; }
    p.dt = 0.2 * h / p.vsound;                                    /* calc_dt_CFL, base:209-212 (before the dt_min rounding) */

    /* particles: fluid lattice + a floor of dummy particles (three layers below y = 0) */
    const int n_fluid = nx * ny * nz, fx = (int)(size[0] / d), fz = (int)(size[2] / d), n_wall = fx * layers * fz;
    const int64_t n = n_fluid + n_wall;
    double *x = malloc(n * 24), *v = calloc(n, 24), *rho = malloc(n * 8);
    int32_t *typ = malloc(n * 4);
    int64_t k = 0;
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) for (int l = 0; l < nz; l++, k++) {
        x[3 * k] = (i + 0.5) * d; x[3 * k + 1] = (j + 0.5) * d; x[3 * k + 2] = (l + 0.5) * d; rho[k] = 1000.0; typ[k] = 1;
    }
    for (int i = 0; i < fx; i++) for (int j = 0; j < layers; j++) for (int l = 0; l < fz; l++, k++) {
        x[3 * k] = (i + 0.5) * d; x[3 * k + 1] = -(j + 0.5) * d; x[3 * k + 2] = (l + 0.5) * d; rho[k] = 0.0; typ[k] = -1;
    }

    /* caller-owned arena and stream */
    const int64_t bytes = sph_arena_bytes(&p, n);
    void *arena = NULL;
    cudaStream_t stream;
    if (cudaMalloc(&arena, (size_t)bytes) != cudaSuccess || cudaStreamCreate(&stream) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    SphCtx *ctx = sph_create(&p, n, arena, bytes, stream);
    if (!ctx) { fprintf(stderr, "sph_create failed\n"); return 1; }
    CHECK(sph_add_particles(ctx, n, x, v, rho, typ));
    CHECK(sph_step(ctx, 10));                                     /* SPHBase.step x 10, enqueued without host round trips */
    float *vout = malloc(n * 16);                                 /* MIXED engine: v comes back as n x 4 float (xyz, mass) */
    double *rout = malloc(n * 8);
    int32_t *id0 = malloc(n * 4);
    if (sph_real_bytes(ctx) != 4) { fprintf(stderr, "expected a float32 engine\n"); return 1; }
    CHECK(sph_read_state(ctx, x, vout, rout, NULL, id0));
    double vy = 0.0, rmax = 0.0;
    for (int64_t i = 0; i < n; i++) if (id0[i] < n_fluid) { vy += vout[4 * i + 1]; if (rout[i] > rmax) rmax = rout[i]; }
    printf("n = %lld, arena = %.1f MB, launches = %lld, mean v_y of the fluid after 10 steps = %.6f m/s, max density = %.3f, "
           "particles outside the grid = %lld\n", (long long)n, bytes / 1e6, (long long)sph_launch_count(ctx), vy / n_fluid, rmax,
           (long long)sph_read_bad_cells(ctx));
    if (argc > 1) {
        FILE *f = fopen(argv[1], "wb");
        if (!f) { fprintf(stderr, "cannot write %s\n", argv[1]); return 1; }
        fwrite(&n, 8, 1, f); fwrite(id0, 4, n, f); fwrite(x, 8, 3 * n, f); fwrite(vout, 4, 4 * n, f); fwrite(rout, 8, n, f);
        fclose(f);
    }
    sph_destroy(ctx);
    cudaFree(arena);
    return 0;
}
