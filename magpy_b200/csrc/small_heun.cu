#include "small.cuh"

namespace mb {

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
    const uint32_t member = member_id(P, r);

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
        case 5:
            if (renorm) heun_small_kernel<NOISE, TAB, 5, true><<<g, b, 0, s>>>(P);
            else heun_small_kernel<NOISE, TAB, 5, false><<<g, b, 0, s>>>(P);
            break;
        case 6:
            if (renorm) heun_small_kernel<NOISE, TAB, 6, true><<<g, b, 0, s>>>(P);
            else heun_small_kernel<NOISE, TAB, 6, false><<<g, b, 0, s>>>(P);
            break;
        case 7:
            if (renorm) heun_small_kernel<NOISE, TAB, 7, true><<<g, b, 0, s>>>(P);
            else heun_small_kernel<NOISE, TAB, 7, false><<<g, b, 0, s>>>(P);
            break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_heun_small(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P) {
    MB_NOISE_TAB_DISPATCH(launch_hsm, n_particles, grid, s, P)
}


}  // namespace mb
