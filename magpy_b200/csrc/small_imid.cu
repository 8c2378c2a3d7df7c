#include "small.cuh"

namespace mb {

// ---------------------------------------------------------------------------------
// implicit midpoint
// ---------------------------------------------------------------------------------
template <int NOISE, bool FIELD_TAB, int N>
__global__ void __launch_bounds__(SMALL_THREADS) imid_small_kernel(const __grid_constant__ RunParams P) {
    __shared__ double red[(SMALL_THREADS / 32) * 4];
    __shared__ __align__(32) double sd[N * N * 4];
    stage_pair_table<N>(sd, P, 1.0);
    const uint64_t r_raw = (uint64_t)blockIdx.x * SMALL_THREADS + threadIdx.x;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    const double alpha = P.alpha, dt = P.dt, clampA = P.clampA, sqrt_dt = P.sqrt_dt;
    const bool inter = P.interactions != 0, renorm = P.renorm != 0;

    V3 m[N], e[N], zero[N];
    double kred[N], sr[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint64_t c0 = 3ull * i;
        m[i] = V3{P.state[c0 * P.R + r], P.state[(c0 + 1) * P.R + r], P.state[(c0 + 2) * P.R + r]};
        e[i] = V3{P.axis[c0 * P.axis_cs + r * P.axis_rs], P.axis[(c0 + 1) * P.axis_cs + r * P.axis_rs],
                  P.axis[(c0 + 2) * P.axis_cs + r * P.axis_rs]};
        kred[i] = P.k_red[i];
        sr[i] = P.sig[i];
        zero[i] = V3{0.0, 0.0, 0.0};
    }
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = (uint32_t)(r + P.stream_offset);
    NewtonCount nc{0ull, 0ull, 0ull};

    uint64_t j = P.j0;
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        const uint64_t tgt = (k < P.k1) ? P.target[k] : P.j1;
        for (; j < tgt; ++j) {
            double hz0 = P.h_const, hz1 = P.h_const;
            if (FIELD_TAB) {
                const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (j - P.j0));
                hz0 = h.x; hz1 = h.y;
            }
            V3 wm[N], sw[N], X[N], h[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const V3 w = draw_noise<NOISE>(P, key0, key1, j, (uint32_t)i, member, r);
                wm[i] = V3{fmax(-clampA, fmin(clampA, w.x)) * sqrt_dt, fmax(-clampA, fmin(clampA, w.y)) * sqrt_dt,
                           fmax(-clampA, fmin(clampA, w.z)) * sqrt_dt};
                sw[i] = V3{sr[i] * wm[i].x, sr[i] * wm[i].y, sr[i] * wm[i].z};
            }
            // Euler half step as the initial guess of (x0 + x1)/2 (lib/integrators.cpp:605-614)
            small_fields<N>(h, m, e, kred, hz0, sd, inter, zero);
            double nrm = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const V3 g{fma(h[i].x, dt, sw[i].x), fma(h[i].y, dt, sw[i].y), fma(h[i].z, dt, sw[i].z)};
                const V3 f = llg_f(m[i], g, alpha);
                X[i] = V3{(f.x + m[i].x) / 2, (f.y + m[i].y) / 2, (f.z + m[i].z) / 2};
                nrm += dot(X[i], X[i]);
            }
            const double tol = P.eps * sqrt(nrm);
            double err = 2 * tol;
            int iter = 1000;
            unsigned long long done = 0;
            bool singular = false;
            while ((err > tol) && (iter-- > 0)) {
                small_fields<N>(h, X, e, kred, hz1, sd, inter, zero);
                V3 dl[N], b[N];
                bool ok = true;
                double e2 = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const V3 g{fma(h[i].x, dt, sw[i].x), fma(h[i].y, dt, sw[i].y), fma(h[i].z, dt, sw[i].z)};
                    const V3 f = llg_f(X[i], g, alpha);
                    double bb[3] = {-(X[i].x - m[i].x - 0.5 * f.x), -(X[i].y - m[i].y - 0.5 * f.y),
                                    -(X[i].z - m[i].z - 0.5 * f.z)};
                    b[i] = V3{bb[0], bb[1], bb[2]};
                    double A[9], d[3];
                    newton_matrix(A, X[i], alpha, h[i], sw[i], quirk_u(N, i, e[0], kred[0]), e[0]);
                    if (!solve3_adjugate(A, bb, d)) { ok = false; d[0] = d[1] = d[2] = 0.0; }
                    dl[i] = V3{d[0], d[1], d[2]};
                    e2 += d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                }
                ++done;
                if (!ok) {
                    // dgesv info > 0: the reference returns with x_root = -F (lib/optimisation.cpp:134-137)
#pragma unroll
                    for (int i = 0; i < N; ++i) X[i] = b[i];
                    singular = true;
                    break;
                }
                err = sqrt(e2);
#pragma unroll
                for (int i = 0; i < N; ++i) { X[i].x += dl[i].x; X[i].y += dl[i].y; X[i].z += dl[i].z; }
            }
            nc.total += done;
            nc.worst = done > nc.worst ? done : nc.worst;
            nc.fails += (singular || iter == -1) ? 1ull : 0ull;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                m[i] = V3{2 * X[i].x - m[i].x, 2 * X[i].y - m[i].y, 2 * X[i].z - m[i].z};
                if (renorm) renormalise(m[i]);
            }
        }
        if (k < P.k1) sample_outputs<N>(P, m, k, r, live, red);
    }
    if (live) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint64_t c0 = 3ull * i;
            P.state[c0 * P.R + r] = m[i].x; P.state[(c0 + 1) * P.R + r] = m[i].y; P.state[(c0 + 2) * P.R + r] = m[i].z;
        }
    }
    newton_flush(P, nc, live);
}

template <int NOISE, bool TAB>
static cudaError_t launch_ism(unsigned N, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(SMALL_THREADS);
    switch (N) {
        case 2: imid_small_kernel<NOISE, TAB, 2><<<g, b, 0, s>>>(P); break;
        case 3: imid_small_kernel<NOISE, TAB, 3><<<g, b, 0, s>>>(P); break;
        case 4: imid_small_kernel<NOISE, TAB, 4><<<g, b, 0, s>>>(P); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_imid_small(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P) {
    MB_NOISE_TAB_DISPATCH(launch_ism, n_particles, grid, s, P)
}


}  // namespace mb
