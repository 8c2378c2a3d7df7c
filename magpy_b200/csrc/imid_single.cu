#include "common.cuh"
#include "launch.h"

namespace mb {

// ---------------------------------------------------------------------------------
// K3: implicit midpoint, single particle (lib/integrators.cpp:576-651 +
// lib/optimisation.cpp:81-149).  The quasi-Newton iteration is reproduced as the
// reference runs it: clamped increments, Euler-midpoint initial guess, Jacobian
// J = I - a'/2 - (B'.w)/2 with no dt on a', tolerance eps*||guess|| fixed before the
// loop, stop on ||delta||_2 <= tol or after 1000 iterations.
// ---------------------------------------------------------------------------------
template <bool AXIS_Z, bool EXACT>
__device__ __forceinline__ V3 imid_single_step(const V3& x0, const V3& e, const double kred, const double alpha,
                                               const double sr, const double dt, const double clampA,
                                               const double sqrt_dt, const double eps, const V3& w,
                                               const double hz_t, const double hz_mid, NewtonCount& nc) {
    const V3 wm{fmax(-clampA, fmin(clampA, w.x)) * sqrt_dt, fmax(-clampA, fmin(clampA, w.y)) * sqrt_dt,
                fmax(-clampA, fmin(clampA, w.z)) * sqrt_dt};
    const V3 sw{sr * wm.x, sr * wm.y, sr * wm.z};
    const V3 ke{kred * e.x, kred * e.y, kred * e.z};   // field-Jacobian block k e e^T = ke e^T
    const V3 nhsw{-0.5 * sw.x, -0.5 * sw.y, -0.5 * sw.z};
    // Euler half step as the initial guess of (x0 + x1)/2
    V3 X;
    {
        const double s = kred * dot(x0, e);
        const V3 g{fma(s * e.x, dt, sw.x), fma(s * e.y, dt, sw.y), fma(fma(s, e.z, hz_t), dt, sw.z)};
        const V3 f = llg_f(x0, g, alpha);
        // the reference's guess is (x0 + f) / 2 (lib/integrators.cpp:605-614); the midpoint of an Euler step is
        // x0 + f / 2, which the exact-Newton mode starts from (one iteration fewer)
        X = EXACT ? V3{fma(0.5, f.x, x0.x), fma(0.5, f.y, x0.y), fma(0.5, f.z, x0.z)}
                  : V3{(f.x + x0.x) / 2, (f.y + x0.y) / 2, (f.z + x0.z) / 2};
    }
    // err > tol is tested on the squares (no square root inside the loop): tol^2 = eps^2 |X|^2
    const double tol2 = (eps * eps) * dot(X, X);
    double err2 = 4 * tol2;
    int iter = 1000;
    unsigned long long done = 0;
    bool singular = false;
    while ((err2 > tol2) && (iter-- > 0)) {
        double A[9], d[3];
        V3 g;
        if (EXACT) {
            const double s = kred * dot(X, e);
            g = V3{fma(s * e.x, dt, sw.x), fma(s * e.y, dt, sw.y), fma(fma(s, e.z, hz_mid), dt, sw.z)};
            const V3 pg = cross(X, g);
            const V3 u{fma(alpha, pg.x, g.x), fma(alpha, pg.y, g.y), fma(alpha, pg.z, g.z)};
            newton_matrix_exact(A, X, alpha, g, u, dt * kred, e);
        } else if (AXIS_Z) {   // easy axis = z: h = (0, 0, k X_z + h_app)
            const double hz = fma(kred, X.z, hz_mid);
            g = V3{sw.x, sw.y, fma(hz, dt, sw.z)};
            newton_matrix_axis_z(A, X, alpha, hz, sw, nhsw, kred);
        } else {
            const double s = kred * dot(X, e);
            const V3 h{s * e.x, s * e.y, fma(s, e.z, hz_mid)};
            g = V3{fma(h.x, dt, sw.x), fma(h.y, dt, sw.y), fma(h.z, dt, sw.z)};
            newton_matrix(A, X, alpha, h, sw, ke, e);
        }
        const V3 f = llg_f(X, g, alpha);
        double b[3] = {-(X.x - x0.x - 0.5 * f.x), -(X.y - x0.y - 0.5 * f.y), -(X.z - x0.z - 0.5 * f.z)};
        ++done;
        if (!solve3_adjugate(A, b, d)) {
            // dgesv info > 0: the reference returns with x_root = -F (lib/optimisation.cpp:136-137)
            X = V3{b[0], b[1], b[2]};
            singular = true;
            break;
        }
        err2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        X.x += d[0]; X.y += d[1]; X.z += d[2];
    }
    nc.total += done;
    nc.worst = done > nc.worst ? done : nc.worst;
    nc.fails += (singular || iter == -1) ? 1ull : 0ull;
    return V3{2 * X.x - x0.x, 2 * X.y - x0.y, 2 * X.z - x0.z};
}

// MP = per-member material parameters (see heun_single.cu): alpha, dt, the clamp, the field scale and the sampling
// schedule are per-thread values.
template <int NOISE, bool FIELD_TAB, bool AXIS_Z, bool EXACT, bool MP = false>
__global__ void __launch_bounds__(SINGLE_THREADS) imid_single_kernel(const __grid_constant__ RunParams P) {
    __shared__ double red[(SINGLE_THREADS / 32) * 4];
    const uint64_t r_raw = (uint64_t)blockIdx.x * SINGLE_THREADS + threadIdx.x;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;

    V3 m{P.state[r], P.state[P.R + r], P.state[2 * P.R + r]};
    const V3 e{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
    const double kred = P.k_red[0], sr = P.sig[r * P.sig_rs];   // per-member sigma when the radii differ
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = member_id(P, r);
    const bool renorm = P.renorm != 0;
    NewtonCount nc{0ull, 0ull, 0ull};
    const double alpha = MP ? P.mp_alpha[r] : P.alpha, dt = MP ? P.mp_dt[r] : P.dt;
    const double sqrt_dt = MP ? sqrt(dt) : P.sqrt_dt;
    const double clampA = MP ? sqrt(2 * 1000.0 * fabs(log(dt))) : P.clampA;   // lib/integrators.cpp:598-599
    const double mp_h0 = MP ? P.mp_h0[r] : 1.0, mp_Ts = MP ? P.mp_Ts[r] : 0.0;

    uint64_t j = MP ? (uint64_t)P.member_j[r] : P.j0;
    const uint64_t tab0 = MP ? P.tab_j0 : P.j0;
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        // MP: every member stops at ITS OWN state index of sample k; a launch that records samples ends on the last of
        // them, a pure-advance launch (k0 == k1) never steps past the member's next sample
        const uint64_t tgt = !MP ? ((k < P.k1) ? P.target[k] : P.j1)
                             : (k < P.k1) ? member_target(k, dt, mp_Ts)
                             : (P.k1 > P.k0) ? j : min(P.j1, member_target(P.k0, dt, mp_Ts));
        for (; j < tgt; ++j) {
            const V3 w = draw_noise<NOISE>(P, key0, key1, j, 0u, member, r);
            double hz0 = MP ? mp_h0 : P.h_const, hz1 = hz0;
            if (FIELD_TAB) {
                const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (j - tab0));
                hz0 = h.x; hz1 = h.y;
                if (MP) { hz0 *= mp_h0; hz1 *= mp_h0; }   // the table holds the unit waveform
            }
            m = imid_single_step<AXIS_Z, EXACT>(m, e, kred, alpha, sr, dt, clampA, sqrt_dt, P.eps, w, hz0, hz1, nc);
            if (renorm) renormalise(m);
        }
        if (k < P.k1) {
            if (P.traj != nullptr && live) {
                double* t = P.traj + (uint64_t)k * 3 * P.R + r;
                t[0] = m.x; t[P.R] = m.y; t[2 * P.R] = m.z;
            }
            if (P.partial != nullptr) {
                const double z = live ? m.z : 0.0;
                cta_partial_sums<SINGLE_THREADS / 32>(live ? m.x : 0.0, live ? m.y : 0.0, z, z * z, red,
                                                      P.partial + ((uint64_t)(k - P.k0) * gridDim.x + blockIdx.x) * 4);
            }
        }
    }
    if (live) {
        P.state[r] = m.x; P.state[P.R + r] = m.y; P.state[2 * P.R + r] = m.z;
        if (MP) P.member_j[r] = (uint32_t)j;
    }
    newton_flush(P, nc, live);
}

template <int NOISE>
static void launch_is_mp(bool tab, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(SINGLE_THREADS);
    if (P.newton_exact) {
        if (tab) imid_single_kernel<NOISE, true, false, true, true><<<g, b, 0, s>>>(P);
        else imid_single_kernel<NOISE, false, false, true, true><<<g, b, 0, s>>>(P);
    } else {
        if (tab) imid_single_kernel<NOISE, true, false, false, true><<<g, b, 0, s>>>(P);
        else imid_single_kernel<NOISE, false, false, false, true><<<g, b, 0, s>>>(P);
    }
}

template <int NOISE>
static void launch_is(bool tab, bool axis_z, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(SINGLE_THREADS);
    if (P.newton_exact) {   // opt-in Newton with the exact Jacobian (general-axis arithmetic for every axis)
        if (tab) imid_single_kernel<NOISE, true, false, true><<<g, b, 0, s>>>(P);
        else imid_single_kernel<NOISE, false, false, true><<<g, b, 0, s>>>(P);
    } else if (tab) {
        if (axis_z) imid_single_kernel<NOISE, true, true, false><<<g, b, 0, s>>>(P);
        else imid_single_kernel<NOISE, true, false, false><<<g, b, 0, s>>>(P);
    } else {
        if (axis_z) imid_single_kernel<NOISE, false, true, false><<<g, b, 0, s>>>(P);
        else imid_single_kernel<NOISE, false, false, false><<<g, b, 0, s>>>(P);
    }
}

cudaError_t launch_imid_single(int noise, bool tab, bool axis_z, unsigned grid, cudaStream_t s, const RunParams& P) {
    if (P.mp_dt != nullptr) {   // per-member material parameters
        if (noise == NOISE_INJECTED) launch_is_mp<NOISE_INJECTED>(tab, grid, s, P);
        else launch_is_mp<NOISE_PHILOX_PACKED>(tab, grid, s, P);
        return cudaGetLastError();
    }
    switch (noise) {
        case NOISE_PHILOX_F32: launch_is<NOISE_PHILOX_F32>(tab, axis_z, grid, s, P); break;
        case NOISE_PHILOX_F64: launch_is<NOISE_PHILOX_F64>(tab, axis_z, grid, s, P); break;
        case NOISE_INJECTED: launch_is<NOISE_INJECTED>(tab, axis_z, grid, s, P); break;
        case NOISE_PHILOX_COARSE: launch_is<NOISE_PHILOX_COARSE>(tab, axis_z, grid, s, P); break;
        default: launch_is<NOISE_PHILOX_PACKED>(tab, axis_z, grid, s, P); break;
    }
    return cudaGetLastError();
}

}  // namespace mb
