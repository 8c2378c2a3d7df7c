// small.cu — interacting clusters of 2..4 particles: ONE THREAD PER CLUSTER.
//
// For a handful of particles the all-pairs dipolar sum is a few dozen FMAs, so the moments of the
// whole cluster live in the fp64 registers of one thread exactly like the single-particle kernels
// (K1/K3): no shared-memory staging, no barrier inside a step, and the quasi-Newton loop of the
// implicit scheme needs no CTA-wide convergence vote.  The static pair table {sqrt(3) r_hat_ij, c_ij}
// (N(N-1) entries, the same for every member) is staged once per CTA in shared memory and read as
// warp-uniform broadcasts.
//
//   heun_small_kernel   lib/integrators.cpp:372-405 over lib/llg.cpp:332-348 with the field of
//                       lib/simulation.cpp:271-290 (anisotropy + applied + all-pairs dipolar)
//   imid_small_kernel   lib/integrators.cpp:576-651 + lib/optimisation.cpp:81-149; the reference's
//                       3N x 3N Jacobian is block diagonal (lib/llg.cpp:378-427 only writes the 3x3 /
//                       3x3x3 diagonal blocks of zero-filled arrays and the dipolar field has no
//                       Jacobian, lib/simulation.cpp:292-303), so dgesv is N pivoted 3x3 solves
#include "common.cuh"
#include "launch.h"

namespace mb {

constexpr int SMALL_THREADS = 128;

// sd[(i*N + j)*4 + {0,1,2}] = sqrt(3) r_hat_ij,  [3] = cscale * c_ij   (diagonal: zeros)
template <int N>
__device__ __forceinline__ void stage_pair_table(double* sd, const RunParams& P, const double cscale) {
    for (int q = threadIdx.x; q < N * N; q += blockDim.x) {
        sd[4 * q + 0] = P.dip[4 * q + 0];   // already sqrt(3) r_hat (host pair table)
        sd[4 * q + 1] = P.dip[4 * q + 1];
        sd[4 * q + 2] = P.dip[4 * q + 2];
        sd[4 * q + 3] = cscale * P.dip[4 * q + 3];
    }
    __syncthreads();
}

// out_i = ks_i (m_i . e_i) e_i + hz z + sum_{j != i} c_ij ((m_j . t_ij) t_ij - m_j) + add_i
//   (t = sqrt(3) r_hat, so (m.t) t = 3 (m.r_hat) r_hat: lib/field.cpp:217-225 in 9 fp64 operations per pair)
template <int N>
__device__ __forceinline__ void small_fields(V3 (&out)[N], const V3 (&m)[N], const V3 (&e)[N], const double (&ks)[N],
                                             const double hz, const double* sd, const bool inter, const V3 (&add)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double s = dot(m[i], e[i]) * ks[i];
        V3 h{fma(s, e[i].x, add[i].x), fma(s, e[i].y, add[i].y), fma(s, e[i].z, add[i].z + hz)};
        if (inter) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (j == i) continue;
                const double4 t = *reinterpret_cast<const double4*>(sd + (i * N + j) * 4);
                const double d = m[j].x * t.x + m[j].y * t.y + m[j].z * t.z;
                h.x = fma(t.w, fma(d, t.x, -m[j].x), h.x);
                h.y = fma(t.w, fma(d, t.y, -m[j].y), h.y);
                h.z = fma(t.w, fma(d, t.z, -m[j].z), h.z);
            }
        }
        out[i] = h;
    }
}

template <int N>
__device__ __forceinline__ void sample_outputs(const RunParams& P, const V3 (&m)[N], const uint32_t k, const uint64_t r,
                                               const bool live, double* red) {
    if (P.traj != nullptr && live) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double* t = P.traj + ((uint64_t)k * 3 * N + 3 * i) * P.R + r;
            t[0] = m[i].x; t[P.R] = m[i].y; t[2 * P.R] = m[i].z;
        }
    }
    if (P.partial != nullptr) {
        double sx = 0, sy = 0, sz = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) { sx += m[i].x; sy += m[i].y; sz += m[i].z; }
        if (!live) { sx = 0; sy = 0; sz = 0; }
        cta_partial_sums<SMALL_THREADS / 32>(sx, sy, sz, sz * sz, red,
                                             P.partial + ((uint64_t)(k - P.k0) * gridDim.x + blockIdx.x) * 4);
    }
}

// ---------------------------------------------------------------------------------
// Heun
// ---------------------------------------------------------------------------------
template <int NOISE, bool FIELD_TAB, int N, bool RENORM>
__global__ void __launch_bounds__(SMALL_THREADS) heun_small_kernel(const __grid_constant__ RunParams P) {
    __shared__ double red[(SMALL_THREADS / 32) * 4];
    __shared__ __align__(32) double sd[N * N * 4];
    const double alpha = P.alpha, dt = P.dt;
    stage_pair_table<N>(sd, P, dt);   // g = h dt + c w: the dipolar prefactor carries the dt
    const uint64_t r_raw = (uint64_t)blockIdx.x * SMALL_THREADS + threadIdx.x;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    const bool inter = P.interactions != 0;

    V3 m[N], e[N];
    double kdt[N], c[N];
    float bm[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint64_t c0 = 3ull * i;
        m[i] = V3{P.state[c0 * P.R + r], P.state[(c0 + 1) * P.R + r], P.state[(c0 + 2) * P.R + r]};
        e[i] = V3{P.axis[c0 * P.axis_cs + r * P.axis_rs], P.axis[(c0 + 1) * P.axis_cs + r * P.axis_rs],
                  P.axis[(c0 + 2) * P.axis_cs + r * P.axis_rs]};
        kdt[i] = P.k_red[i] * dt;
        c[i] = P.sig[i] * P.sqrt_dt;
        bm[i] = scale_to_bm(c[i]);
    }
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = (uint32_t)(r + P.stream_offset);

    // one Heun step of the whole cluster from the scaled increments cw (same fused form as K1:
    // f(m,g) = -m x (g + alpha m x g), predictor/corrector adds folded into the last cross product)
    auto advance = [&](const V3 (&cw)[N], const uint64_t jj) {
        double hz0 = P.h_const * dt, hz1 = hz0;
        if (FIELD_TAB) {
            const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (jj - P.j0));
            hz0 = h.x * dt; hz1 = h.y * dt;
        }
        V3 g[N], mt[N];
        small_fields<N>(g, m, e, kdt, hz0, sd, inter, cw);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const V3 p = cross(m[i], g[i]);
            const V3 u{fma(alpha, p.x, g[i].x), fma(alpha, p.y, g[i].y), fma(alpha, p.z, g[i].z)};
            mt[i] = V3{fma(-m[i].y, u.z, fma(m[i].z, u.y, m[i].x)), fma(-m[i].z, u.x, fma(m[i].x, u.z, m[i].y)),
                       fma(-m[i].x, u.y, fma(m[i].y, u.x, m[i].z))};
        }
        small_fields<N>(g, mt, e, kdt, hz1, sd, inter, cw);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const V3 p = cross(mt[i], g[i]);
            const V3 u{fma(alpha, p.x, g[i].x), fma(alpha, p.y, g[i].y), fma(alpha, p.z, g[i].z)};
            const V3 hm{0.5 * mt[i].x, 0.5 * mt[i].y, 0.5 * mt[i].z};
            const V3 h{fma(0.5, m[i].x, hm.x), fma(0.5, m[i].y, hm.y), fma(0.5, m[i].z, hm.z)};
            m[i] = V3{fma(-hm.y, u.z, fma(hm.z, u.y, h.x)), fma(-hm.z, u.x, fma(hm.x, u.z, h.y)),
                      fma(-hm.x, u.y, fma(hm.y, u.x, h.z))};
            if (RENORM) renormalise(m[i]);
        }
    };

    uint64_t j = P.j0;
    // Packed noise: g6[i] holds the six increments of particle i from Philox block `gblk` (steps 2 gblk and
    // 2 gblk + 1) while `have` is set; it is carried across sample boundaries.  For N <= 2 the pair loop is
    // software pipelined like K1 (the next block is generated inside the body that integrates the current one).
    float g6[N][6];
    uint32_t gblk = 0;
    bool have = false;
    auto generate = [&](const uint32_t blk, float (&out)[N][6]) {
#pragma unroll
        for (int i = 0; i < N; ++i) philox_gauss6_f32(key0, key1, blk, (uint32_t)i, member, bm[i], out[i]);
    };
    auto need = [&](const uint32_t blk) {
        if (!have || gblk != blk) {
            generate(blk, g6);
            gblk = blk;
            have = true;
        }
    };
    auto half_step = [&](const int hlf, const uint64_t jj) {
        V3 cw[N];
#pragma unroll
        for (int i = 0; i < N; ++i)
            cw[i] = V3{widen_f32(g6[i][3 * hlf]), widen_f32(g6[i][3 * hlf + 1]), widen_f32(g6[i][3 * hlf + 2])};
        advance(cw, jj);
    };
    constexpr bool PIPELINED = N <= 2;
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        const uint64_t tgt = (k < P.k1) ? P.target[k] : P.j1;
        if (NOISE == NOISE_PHILOX_PACKED) {
            if ((j & 1) && j < tgt) {
                need((uint32_t)(j >> 1));
                half_step(1, j);
                ++j;
            }
            const uint32_t pairs = (uint32_t)((tgt - j) >> 1);
            uint32_t blk = (uint32_t)(j >> 1);
            if (pairs != 0) need(blk);
            for (uint32_t i = pairs; i != 0; --i, j += 2) {
                if (PIPELINED) {
                    float gn[N][6];
                    generate(++blk, gn);
                    half_step(0, j);
                    half_step(1, j + 1);
#pragma unroll
                    for (int p = 0; p < N; ++p)
#pragma unroll
                        for (int q = 0; q < 6; ++q) g6[p][q] = gn[p][q];
                    gblk = blk;
                } else {
                    half_step(0, j);
                    half_step(1, j + 1);
                    ++blk;
                    if (i > 1) { generate(blk, g6); gblk = blk; }
                    else have = false;
                }
            }
            if (j < tgt) {
                need((uint32_t)(j >> 1));
                half_step(0, j);
                ++j;
            }
        } else {
            for (; j < tgt; ++j) {
                V3 cw[N];
#pragma unroll
                for (int i = 0; i < N; ++i)
                    cw[i] = draw_scaled<NOISE>(P, key0, key1, j, (uint32_t)i, member, r, c[i], bm[i]);
                advance(cw, j);
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
}

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

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
template <int NOISE, bool TAB>
static cudaError_t launch_hsm(unsigned N, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(SMALL_THREADS);
    const bool renorm = P.renorm != 0;
    switch (N) {
        case 2:
            if (renorm) heun_small_kernel<NOISE, TAB, 2, true><<<g, b, 0, s>>>(P);
            else heun_small_kernel<NOISE, TAB, 2, false><<<g, b, 0, s>>>(P);
            break;
        case 3:
            if (renorm) heun_small_kernel<NOISE, TAB, 3, true><<<g, b, 0, s>>>(P);
            else heun_small_kernel<NOISE, TAB, 3, false><<<g, b, 0, s>>>(P);
            break;
        case 4:
            if (renorm) heun_small_kernel<NOISE, TAB, 4, true><<<g, b, 0, s>>>(P);
            else heun_small_kernel<NOISE, TAB, 4, false><<<g, b, 0, s>>>(P);
            break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
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

#define MB_NOISE_TAB_DISPATCH(fn, ...)                                                              \
    switch (noise) {                                                                                \
        case NOISE_PHILOX_F32: return tab ? fn<NOISE_PHILOX_F32, true>(__VA_ARGS__) : fn<NOISE_PHILOX_F32, false>(__VA_ARGS__); \
        case NOISE_PHILOX_F64: return tab ? fn<NOISE_PHILOX_F64, true>(__VA_ARGS__) : fn<NOISE_PHILOX_F64, false>(__VA_ARGS__); \
        case NOISE_INJECTED: return tab ? fn<NOISE_INJECTED, true>(__VA_ARGS__) : fn<NOISE_INJECTED, false>(__VA_ARGS__);       \
        default: return tab ? fn<NOISE_PHILOX_PACKED, true>(__VA_ARGS__) : fn<NOISE_PHILOX_PACKED, false>(__VA_ARGS__);        \
    }

cudaError_t launch_heun_small(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P) {
    MB_NOISE_TAB_DISPATCH(launch_hsm, n_particles, grid, s, P)
}

cudaError_t launch_imid_small(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P) {
    MB_NOISE_TAB_DISPATCH(launch_ism, n_particles, grid, s, P)
}

}  // namespace mb
