// integrate.cu -- pointwise kernels: particle creation, init_real2tmp, SE / "LF" / RK4 updates, init_stress.
// Replaces eng/solver_sph_base.py:67-180 (integrators), :249-260 (init_stress) and eng/particle_system.py:274-314.
#include <string.h>
#include "sph_host.h"

namespace sph {

// ps:274-287 add_particle + ps:208-211 set_id0.  v arrives as float64 n x 3 in a staging buffer.
template <typename T>
__global__ void __launch_bounds__(256) k_add_finish(Dev<T> c, const double *__restrict__ vstage, int first, int count) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int i = first + k;
    const double rho = c.rho[i];
    Vec4<T> v;
    v.x = (T)vstage[3 * (size_t)i]; v.y = (T)vstage[3 * (size_t)i + 1]; v.z = (T)vstage[3 * (size_t)i + 2];
    v.w = (T)(c.m_V0d * rho);                           // mass = m_V0 * density (ps:282)
    c.v4[i] = v;
    Vec4<T> z; z.x = z.y = z.z = z.w = 0;
    c.vt4[i] = z;
    Vec4<T> xs; xs.x = xs.y = xs.z = 0; xs.w = c.m_V0;  // m_V = m_V0 (ps:281)
    c.xs4[i] = xs;
    c.rho_t[i] = 0.0;
    c.press[i] = 0;
    c.id0[i] = i;
}

template <typename T> int add_particles_finish(SphCtx *c, int64_t first, int64_t count) {
    Dev<T> d = make_dev<T>(c);
    const double *vstage = (const double *)(c->arena + c->f[SPH_F_X].off[1 - c->f[SPH_F_X].cur]);
    SPH_PROF(c, K_OTHER);
    k_add_finish<T><<<blocks_for(count, 256), 256, 0, c->stream>>>(d, vstage, (int)first, (int)count);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// base:67-74
template <typename T> __global__ void __launch_bounds__(256) k_init_real2tmp(Dev<T> c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    const int t = c.type[i];
    if (is_real(t)) {
        const double r = c.rho[i];
        c.rho_t[i] = r;
        Vec4<T> v = c.v4[i];
        v.w = (T)r;
        c.vt4[i] = v;
    }
    if (is_soil(t) && c.stress) {              // (a WCSPH engine has no stress arrays: a soil block in a water scene is inert)
#pragma unroll
        for (int q = 0; q < 6; q++) c.stress_t[6 * (size_t)i + q] = c.stress[6 * (size_t)i + q];
    }
}
template <typename T> int init_real2tmp(SphCtx *c) {
    if (c->n == 0) return 0;
    SPH_PROF(c, K_INIT_TMP);
    k_init_real2tmp<T><<<blocks_for(c->n, 256), 256, 0, c->stream>>>(make_dev<T>(c));
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// kind: 0 advect_SE/advect_LF (base:79-87,106-114), 1 advect_LF_half (base:96-104), 2 advect_RK_4 (base:134-142),
//       3 init_RK (base:144-151), 4 update_RK(m) (base:153-160), 5 advect_RK (base:162-170)
template <typename T> __global__ void __launch_bounds__(256) k_advect(Dev<T> c, int kind, T m) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    const int t = c.type[i];
    const bool re = is_real(t), so = is_soil(t) && c.stress != nullptr;
    if (!re && !so) return;
    const double dt = c.dt;
    const size_t i6 = 6 * (size_t)i;
    switch (kind) {
    case 0:
        if (re) {
            Vec4<T> v = c.v4[i], dv = c.d_vel[i];
            const double r = c.rho[i] + dt * (double)c.d_rho[i];
            c.rho[i] = r;
            Vec4<T> xs = c.xs4[i]; xs.w = (T)((double)v.w / r); c.xs4[i] = xs;
            v.x += (T)dt * dv.x; v.y += (T)dt * dv.y; v.z += (T)dt * dv.z;
            c.v4[i] = v;
        }
        if (so) for (int q = 0; q < 6; q++) c.stress[i6 + q] += (T)dt * c.d_stress[i6 + q];
        break;
    case 1:
        if (re) {
            Vec4<T> vt = c.vt4[i], dv = c.d_vel[i];
            const double r = c.rho_t[i] + 0.5 * dt * (double)c.d_rho[i];
            c.rho_t[i] = r;
            Vec4<T> xs = c.xs4[i]; xs.w = (T)((double)c.v4[i].w / r); c.xs4[i] = xs;
            const T hdt = (T)(0.5 * dt);
            vt.x += hdt * dv.x; vt.y += hdt * dv.y; vt.z += hdt * dv.z; vt.w = (T)r;
            c.vt4[i] = vt;
        }
        if (so) for (int q = 0; q < 6; q++) c.stress_t[i6 + q] += (T)(0.5 * dt) * c.d_stress[i6 + q];
        break;
    case 2:
        if (re) {
            Vec4<T> v = c.v4[i], dv = c.d_vel[i], vt;
            const double r = 0.5 * dt * (double)c.d_rho[i] + c.rho[i];
            c.rho_t[i] = r;
            Vec4<T> xs = c.xs4[i]; xs.w = (T)((double)v.w / r); c.xs4[i] = xs;
            const T hdt = (T)(0.5 * dt);
            vt.x = hdt * dv.x + v.x; vt.y = hdt * dv.y + v.y; vt.z = hdt * dv.z + v.z; vt.w = (T)r;
            c.vt4[i] = vt;
        }
        if (so) for (int q = 0; q < 6; q++) c.stress_t[i6 + q] = (T)(0.5 * dt) * c.d_stress[i6 + q] + c.stress[i6 + q];
        break;
    case 3:
        if (re) { c.d_rho_rk[i] = 0; Vec4<T> z; z.x = z.y = z.z = z.w = 0; c.d_vel_rk[i] = z; }
        if (so) for (int q = 0; q < 6; q++) c.d_stress_rk[i6 + q] = 0;
        break;
    case 4:
        if (re) {
            c.d_rho_rk[i] += c.d_rho[i] * m;
            Vec4<T> a = c.d_vel_rk[i], dv = c.d_vel[i];
            a.x += dv.x * m; a.y += dv.y * m; a.z += dv.z * m;
            c.d_vel_rk[i] = a;
        }
        if (so) for (int q = 0; q < 6; q++) c.d_stress_rk[i6 + q] += c.d_stress[i6 + q] * m;
        break;
    case 5:
        if (re) {
            Vec4<T> v = c.v4[i], a = c.d_vel_rk[i];
            const double r = c.rho[i] + dt / 6.0 * (double)c.d_rho_rk[i];
            c.rho[i] = r;
            Vec4<T> xs = c.xs4[i]; xs.w = (T)((double)v.w / r); c.xs4[i] = xs;
            const T s = (T)(dt / 6.0);
            v.x += s * a.x; v.y += s * a.y; v.z += s * a.z;
            c.v4[i] = v;
        }
        if (so) for (int q = 0; q < 6; q++) c.stress[i6 + q] += (T)(dt / 6.0) * c.d_stress_rk[i6 + q];
        break;
    }
}
// RK4 inside sph_step: the pointwise kernels between two one_steps in ONE pass over the particle -- update_RK(m)
// (base:153-160; the first stage also is init_RK, base:144-151: 0 + m D == m D) followed by advect_RK_4 (base:134-142)
// or, after the last stage, advect_RK (base:162-170).  The same operations in the same order per particle as
// k_advect kinds 3, 4, 2 / 5, without writing the accumulators and reading them back between launches.
template <typename T> __global__ void __launch_bounds__(256) k_rk_stage(Dev<T> c, T m, int first, int last) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    const int t = c.type[i];
    const bool re = is_real(t), so = is_soil(t) && c.stress != nullptr;
    if (!re && !so) return;
    const double dt = c.dt;
    const size_t i6 = 6 * (size_t)i;
    if (re) {
        const Vec4<T> dv = c.d_vel[i];
        const T dr = c.d_rho[i];
        T ar = first ? (T)0 : c.d_rho_rk[i];
        Vec4<T> a;
        if (first) { a.x = a.y = a.z = a.w = 0; } else a = c.d_vel_rk[i];
        ar += dr * m;
        a.x += dv.x * m; a.y += dv.y * m; a.z += dv.z * m;
        c.d_rho_rk[i] = ar;
        c.d_vel_rk[i] = a;
        Vec4<T> v = c.v4[i];
        Vec4<T> xs = c.xs4[i];
        if (!last) {
            Vec4<T> vt;
            const double r = 0.5 * dt * (double)dr + c.rho[i];
            c.rho_t[i] = r;
            xs.w = (T)((double)v.w / r); c.xs4[i] = xs;
            const T hdt = (T)(0.5 * dt);
            vt.x = hdt * dv.x + v.x; vt.y = hdt * dv.y + v.y; vt.z = hdt * dv.z + v.z; vt.w = (T)r;
            c.vt4[i] = vt;
        } else {
            const double r = c.rho[i] + dt / 6.0 * (double)ar;
            c.rho[i] = r;
            xs.w = (T)((double)v.w / r); c.xs4[i] = xs;
            const T s = (T)(dt / 6.0);
            v.x += s * a.x; v.y += s * a.y; v.z += s * a.z;
            c.v4[i] = v;
        }
    }
    if (so) {
        for (int q = 0; q < 6; q++) {
            const T ds = c.d_stress[i6 + q];
            T as = first ? (T)0 : c.d_stress_rk[i6 + q];
            as += ds * m;
            c.d_stress_rk[i6 + q] = as;
            if (!last) c.stress_t[i6 + q] = (T)(0.5 * dt) * ds + c.stress[i6 + q];
            else c.stress[i6 + q] += (T)(dt / 6.0) * as;
        }
    }
}
template <typename T> int rk_stage(SphCtx *c, int m, bool first, bool last) {
    if (c->n == 0) return 0;
    SPH_PROF(c, K_ADVECT);
    k_rk_stage<T><<<blocks_for(c->n, 256), 256, 0, c->stream>>>(make_dev<T>(c), (T)m, first ? 1 : 0, last ? 1 : 0);
    SPH_LAUNCH_CHECK(c);
    return 0;
}
template int rk_stage<float>(SphCtx *, int, bool, bool);
template int rk_stage<double>(SphCtx *, int, bool, bool);

template <typename T> int advect(SphCtx *c, int kind, int m) {
    if (c->n == 0) return 0;
    if (kind >= 3 && !c->rk) { snprintf(c->err, sizeof(c->err), "RK buffers exist only when timeIntegration == 4"); return -2; }
    SPH_PROF(c, K_ADVECT);
    k_advect<T><<<blocks_for(c->n, 256), 256, 0, c->stream>>>(make_dev<T>(c), kind, (T)m);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

// base:249-260: y_max over soil particles, then K0 hydrostatic stress
__device__ __forceinline__ unsigned long long enc_f64(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_f64(unsigned long long u) {
    u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}
template <typename T> __global__ void __launch_bounds__(256) k_ymax(Dev<T> c, unsigned long long *ymax) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    if (is_soil(c.type[i])) atomicMax(ymax, enc_f64(c.x[3 * (size_t)i + 1]));
}
template <typename T> __global__ void __launch_bounds__(256) k_init_stress(Dev<T> c, const unsigned long long *ymax, double K0, double gy) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N()) return;
    if (!is_soil(c.type[i])) return;
    const double ver = c.rho0 * gy * (dec_f64(*ymax) - c.x[3 * (size_t)i + 1]);
    T *s = c.stress + 6 * (size_t)i;
    s[0] = (T)(K0 * ver); s[1] = (T)ver; s[2] = (T)(K0 * ver);
}
static unsigned long long enc_f64_host(double v) {
    unsigned long long u;
    memcpy(&u, &v, 8);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
// ymax_ext: the y of the highest soil particle of the WHOLE scene when this ctx only holds a slab of it (a rank-local
// maximum would give every slab its own geostatic stress); null: the maximum over this ctx's particles (base:251-255)
template <typename T> int init_stress(SphCtx *c, const double *ymax_ext) {
    if (c->n == 0) return 0;
    if (!c->soil) { snprintf(c->err, sizeof(c->err), "init_stress needs a soil solver"); return -2; }
    Dev<T> d = make_dev<T>(c);
    unsigned long long *ymax = (unsigned long long *)(c->arena + c->off_bad + 8);
    if (ymax_ext) {
        const unsigned long long e = enc_f64_host(*ymax_ext);
        SPH_CHECK(c, cudaMemcpyAsync(ymax, &e, 8, cudaMemcpyHostToDevice, c->stream));
        SPH_CHECK(c, cudaStreamSynchronize(c->stream));
    } else {
        SPH_CHECK(c, cudaMemsetAsync(ymax, 0, 8, c->stream));
        SPH_PROF(c, K_OTHER);
        k_ymax<T><<<blocks_for(c->n, 256), 256, 0, c->stream>>>(d, ymax);
        SPH_LAUNCH_CHECK(c);
    }
    SPH_PROF(c, K_OTHER);
    k_init_stress<T><<<blocks_for(c->n, 256), 256, 0, c->stream>>>(d, ymax, 1.0 - sin(c->p.fric), c->p.g[1]);
    SPH_LAUNCH_CHECK(c);
    return 0;
}

template int add_particles_finish<float>(SphCtx *, int64_t, int64_t);
template int add_particles_finish<double>(SphCtx *, int64_t, int64_t);
template int init_real2tmp<float>(SphCtx *);
template int init_real2tmp<double>(SphCtx *);
template int advect<float>(SphCtx *, int, int);
template int advect<double>(SphCtx *, int, int);
template int init_stress<float>(SphCtx *, const double *);
template int init_stress<double>(SphCtx *, const double *);

}  // namespace sph
