// cluster_mma.cu — K2m: Heun for interacting clusters with the dipolar field as a matrix product on the
// FP64 MMA path (DMMA.8x8x4) of sm_100a.
//
// The all-pairs dipolar field (lib/field.cpp:187-225) of ONE member is a matrix-vector product with a matrix
// that depends on the geometry only,
//     H[(i,a)] = sum_{(j,b)} D[(i,a),(j,b)] M[(j,b)],   D = c_dip / cube_ij (3 r_a r_b - delta_ab),  M_j = v_red,j m_j,
// and the geometry is shared by every member of the ensemble, so for a CTA's members it is the product
// H[3N x members] = D[3N x 3N] . M[3N x members]: the same 9 FMA per ordered pair as the scalar loop of cluster.cu,
// but issued as one DMMA per 256 FMA (instead of one DFMA per 32) and fed with 5 shared-memory operand loads per
// 6 DMMA (instead of 11 per 36 DFMA).  Measured on B200 (profiles/r01_dmma_microbench.txt): DMMA.8x8x4 runs at
// 16 cycles per SM sub-partition = the full rate of the FP64 pipe (37.1 TFLOP/s), also when fed from shared memory.
//
// Index order of the rows of D and of M: (particle group pg = p / 8, component a, g = p % 8) -> 24 pg + 8 a + g.
// A warp owns one particle group (three 8-row tiles = x, y, z of 8 particles) and one 16-member half of the CTA's
// members (two 8-column tiles), so that after the product thread (g, t) holds all three field components of
// particle 8 pg + g for its four members 16 mh + 8 j + 2 t + e — exactly what the per-particle Heun update needs,
// with no exchange between threads.
//
// Clusters of 65..128 particles use the same kernel with the blocks of D read from global memory through the read-only
// path (template parameter DG): the matrix (up to 612 KB) is shared by every CTA and stays in L2 — measured 30 TFLOP/s
// by W_alg at N = 128 against 14.8 for the scalar kernel.
//
// D is symmetric (v_red is folded into M), so only the 24 x 24 blocks with pg <= kg are kept in shared memory
// (N = 64: 36 blocks = 162 KB instead of 288 KB); a warp reads the blocks left of its diagonal transposed.  Inside
// a block, element (i, c) sits at i * 24 + (c ^ 4 ((i >> 1) & 1)), which makes both the direct and the
// transposed fragment loads bank-conflict free.
//
// Every group of warps that shares a 16-member half is an independent problem and synchronises on its own named
// barrier.  (DFMA and DMMA share the FP64 datapath — scripts/micro/dmma.cu — so running one group's update under
// another's product hides latency only; a forced phase offset between the groups measured no gain.)
#include <type_traits>
#include "common.cuh"
#include "launch.h"
#include "mma.cuh"

namespace mb {

// TAIL = false: a CTA of the whole waves (all column tiles live, the fast path); TAIL = true: a CTA of the partial last
// wave, launched separately (P.cta_offset), whose warps may have 2, 1 or 0 live column tiles.
// DG = true: the matrix stays in global memory (N = 65..128), two moment buffers.
template <int NOISE, bool FIELD_TAB, bool ONE_BUF, bool TAIL, bool DG>
__global__ void __launch_bounds__(512, 1) heun_cluster_mma_kernel(const __grid_constant__ RunParams P) {
    extern __shared__ double smem[];
    const int N = (int)P.N, G = (int)P.G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int n_warps = blockDim.x >> 5, MH = n_warps / G, MB = 16 * MH, LD = MB + 4;
    const int pg = warp % G, mh = warp / G;
    const int n_blk = G * (G + 1) / 2;
    double* sm_d = smem;                                              // [n_blk][576] (not staged when DG)
    double* sm_m = sm_d + (DG ? 0 : (size_t)n_blk * MMA_BLK);         // [24 G][LD] current moments (times v_red)
    double* sm_t = ONE_BUF ? sm_m : sm_m + (size_t)24 * G * LD;       // [24 G][LD] predictor moments
    double* sm_red = sm_t + (size_t)24 * G * LD;                      // [G][3][MB] sample reduction
    {
        const int nd = n_blk * MMA_BLK;
        if (!DG)
            for (int q = threadIdx.x; q < nd; q += blockDim.x) sm_d[q] = P.dmat[q];
        const int nm = 24 * G * LD * (ONE_BUF ? 1 : 2);
        for (int q = threadIdx.x; q < nm; q += blockDim.x) sm_m[q] = 0.0;   // rows of padding particles stay zero
    }
    __syncthreads();

    const int p_raw = 8 * pg + g;
    const bool valid = p_raw < N;
    const uint32_t pid = valid ? (uint32_t)p_raw : 0u;
    const double alpha = P.alpha, dt = P.dt;
    const bool renorm = P.renorm != 0, inter = P.interactions != 0;
    const double kred = __ldg(P.k_red + pid), vred = __ldg(P.v_red + pid);
    const double csig = __ldg(P.sig + pid) * P.sqrt_dt;
    const float bm = scale_to_bm(csig);
    const int bar_id = 1 + mh, bar_n = 32 * G;
    const bool mono = P.mma_mono != 0;         // every particle has the mean volume (v_red = 1)
    const bool shared_axis = P.axis_rs == 0;   // one set of easy axes for all members

    // the thread's four members: q = 2 j + e  ->  local column 16 mh + 8 j + 2 t + e
    // Members of this CTA: the first P.mma_full CTAs (whole waves) take MB each; the CTAs of the partial last wave
    // take P.mma_tail (a multiple of 8 <= MB) so that the wave's work is spread over all SMs in column tiles of 8.
    const uint32_t cta = blockIdx.x + P.cta_offset;
    const uint64_t cta_start = (!TAIL || cta < P.mma_full)
                                   ? (uint64_t)cta * MB
                                   : (uint64_t)P.mma_full * MB + (uint64_t)(cta - P.mma_full) * P.mma_tail;
    const int cta_cnt = (!TAIL || cta < P.mma_full) ? MB : (int)P.mma_tail;
    const uint64_t r_base = cta_start + 16 * mh + 2 * t;
    auto is_live = [&](const int q) {
        const int o = 8 * (q >> 1) + (q & 1);
        return (!TAIL || 16 * mh + 2 * t + o < cta_cnt) && r_base + o < P.R;
    };
    auto member = [&](const int q) {   // dead columns repeat the last member (computed, never stored)
        const uint64_t r_raw = r_base + 8 * (q >> 1) + (q & 1);
        return r_raw < P.R ? r_raw : P.R - 1;
    };
    // live column tiles of this warp (the same for all warps of a member half): tile j = columns 16 mh + 8 j .. + 7
    const int n_tiles = !TAIL ? 2
                        : (16 * mh < cta_cnt && cta_start + 16 * mh < P.R)
                            ? ((16 * mh + 8 < cta_cnt && cta_start + 16 * mh + 8 < P.R) ? 2 : 1) : 0;
    V3 m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint64_t c0 = 3ull * pid, rq = member(q);
        m[q] = V3{P.state[c0 * P.R + rq], P.state[(c0 + 1) * P.R + rq], P.state[(c0 + 2) * P.R + rq]};
    }
    // own rows of the moment buffers: row(comp a) = 24 pg + 8 a + g, columns 16 mh + 8 j + 2 t + {0, 1}
    const int own_off = (24 * pg + g) * LD + 16 * mh + 2 * t;
    // NT (live column tiles of the warp, a compile-time constant of the step loop) bounds every per-member loop
    auto put = [&](auto nt, double* buf, const V3 (&x)[4]) {
        constexpr int NT = decltype(nt)::value;
        if (!valid) return;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            double* d = buf + own_off + 8 * j;
            if (mono) {   // all reduced volumes are 1: the moments go in as they are
                *reinterpret_cast<double2*>(d) = make_double2(x[2 * j].x, x[2 * j + 1].x);
                *reinterpret_cast<double2*>(d + 8 * LD) = make_double2(x[2 * j].y, x[2 * j + 1].y);
                *reinterpret_cast<double2*>(d + 16 * LD) = make_double2(x[2 * j].z, x[2 * j + 1].z);
            } else {
                *reinterpret_cast<double2*>(d) = make_double2(vred * x[2 * j].x, vred * x[2 * j + 1].x);
                *reinterpret_cast<double2*>(d + 8 * LD) = make_double2(vred * x[2 * j].y, vred * x[2 * j + 1].y);
                *reinterpret_cast<double2*>(d + 16 * LD) = make_double2(vred * x[2 * j].z, vred * x[2 * j + 1].z);
            }
        }
    };
    put(std::integral_constant<int, 2>{}, sm_m, m);
    __syncthreads();
    // this thread's B-fragment row of the two moment buffers, as indices into smem
    const int b_m = (int)(sm_m - smem) + t * LD + 16 * mh + g;
    const int b_t = (int)(sm_t - smem) + t * LD + 16 * mh + g;

    constexpr bool PACKED = NOISE == NOISE_PHILOX_PACKED;
    // scaled increments of the current step: kept as the fp32 the packed generator produces (widened at use) so that
    // they cost 12 registers, not 24, across the two matrix products
    struct Inc {
        typename std::conditional<PACKED, float, double>::type x, y, z;
    };
    const double inv_v = 1.0 / vred;
    auto get = [&](auto nt, const double* buf, V3 (&x)[4]) {   // own moments back from a moment buffer
        constexpr int NT = decltype(nt)::value;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const double* d = buf + own_off + 8 * j;
            const double2 vx = *reinterpret_cast<const double2*>(d), vy = *reinterpret_cast<const double2*>(d + 8 * LD),
                          vz = *reinterpret_cast<const double2*>(d + 16 * LD);
            if (mono) {
                x[2 * j] = V3{vx.x, vy.x, vz.x};
                x[2 * j + 1] = V3{vx.y, vy.y, vz.y};
            } else {
                x[2 * j] = V3{vx.x * inv_v, vy.x * inv_v, vz.x * inv_v};
                x[2 * j + 1] = V3{vx.y * inv_v, vy.y * inv_v, vz.y * inv_v};
            }
        }
    };

    // g_q = dt h_q + cw_q with h = anisotropy + applied + dipolar (acc)
    auto stage_g = [&](auto nt, V3 (&gq)[4], const V3 (&x)[4], const double (&acc)[3][2][2], const double hz,
                       const Inc (&cw)[4]) {
        constexpr int NQ = 2 * decltype(nt)::value;
        const uint64_t c0 = 3ull * pid;
        V3 e{0.0, 0.0, 0.0};
        if (shared_axis) e = V3{__ldg(P.axis + c0), __ldg(P.axis + c0 + 1), __ldg(P.axis + c0 + 2)};
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (!shared_axis) {
                const uint64_t rq = member(q);
                e = V3{__ldg(P.axis + c0 * P.axis_cs + rq), __ldg(P.axis + (c0 + 1) * P.axis_cs + rq),
                       __ldg(P.axis + (c0 + 2) * P.axis_cs + rq)};
            }
            const double s = dot(x[q], e) * kred;
            const V3 h{fma(s, e.x, acc[0][q >> 1][q & 1]), fma(s, e.y, acc[1][q >> 1][q & 1]),
                       fma(s, e.z, hz) + acc[2][q >> 1][q & 1]};
            gq[q] = V3{fma(h.x, dt, (double)cw[q].x), fma(h.y, dt, (double)cw[q].y), fma(h.z, dt, (double)cw[q].z)};
        }
    };

    auto advance = [&](auto nt, const Inc (&cw)[4], const uint64_t jj) {
        constexpr int NT = decltype(nt)::value, NQ = 2 * NT;
        double hz0 = P.h_const, hz1 = P.h_const;
        if (FIELD_TAB) {
            const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (jj - P.j0));
            hz0 = h.x; hz1 = h.y;
        }
        double acc[3][2][2];
        V3 gq[4];
        if (inter) dipolar_mma<NT, DG>(acc, smem, P.dmat, b_m, G, pg, LD, g, t);
        else {
#pragma unroll
            for (int a = 0; a < 3; ++a) acc[a][0][0] = acc[a][0][1] = acc[a][1][0] = acc[a][1][1] = 0.0;
        }
        stage_g(nt, gq, m, acc, hz0, cw);
        if (ONE_BUF) group_barrier(bar_id, bar_n);   // every warp of the group has read the current moments
        {
            V3 mt[4];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const V3 pc = cross(m[q], gq[q]);
                const V3 u{fma(alpha, pc.x, gq[q].x), fma(alpha, pc.y, gq[q].y), fma(alpha, pc.z, gq[q].z)};
                mt[q] = V3{fma(-m[q].y, u.z, fma(m[q].z, u.y, m[q].x)), fma(-m[q].z, u.x, fma(m[q].x, u.z, m[q].y)),
                           fma(-m[q].x, u.y, fma(m[q].y, u.x, m[q].z))};
            }
            put(nt, sm_t, mt);
        }
        group_barrier(bar_id, bar_n);
        if (inter) dipolar_mma<NT, DG>(acc, smem, P.dmat, b_t, G, pg, LD, g, t);
        // the predictor moments are not kept in registers across the second product: the thread reads its own back
        // from shared memory (exact when v_red = 1; otherwise one rounding of v (1/v), 12 orders below the noise)
        V3 mt[4];
        get(nt, sm_t, mt);
        stage_g(nt, gq, mt, acc, hz1, cw);
        if (ONE_BUF) group_barrier(bar_id, bar_n);   // every warp of the group has read the predictor moments
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const V3 pc = cross(mt[q], gq[q]);
            const V3 u{fma(alpha, pc.x, gq[q].x), fma(alpha, pc.y, gq[q].y), fma(alpha, pc.z, gq[q].z)};
            const V3 hm{0.5 * mt[q].x, 0.5 * mt[q].y, 0.5 * mt[q].z};
            const V3 hh{fma(0.5, m[q].x, hm.x), fma(0.5, m[q].y, hm.y), fma(0.5, m[q].z, hm.z)};
            m[q] = V3{fma(-hm.y, u.z, fma(hm.z, u.y, hh.x)), fma(-hm.z, u.x, fma(hm.x, u.z, hh.y)),
                      fma(-hm.x, u.y, fma(hm.y, u.x, hh.z))};
            if (renorm) renormalise(m[q]);
        }
        put(nt, sm_m, m);
        group_barrier(bar_id, bar_n);
    };

    float carry[PACKED ? 4 : 1][3];
    bool have_carry = false;
    uint64_t j = P.j0;
    // all steps up to state index tgt for a warp with NT live column tiles
    auto steps_to = [&](auto nt, const uint64_t tgt) {
        constexpr int NQ = 2 * decltype(nt)::value;
        for (; j < tgt; ++j) {
            Inc cw[4];
            if constexpr (PACKED) {
                const bool odd = (j & 1) != 0;
                if (odd && have_carry) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) cw[q] = Inc{carry[q][0], carry[q][1], carry[q][2]};
                    have_carry = false;
                } else {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const uint64_t rq = member(q), seed = (uint64_t)__ldg(P.seeds + rq);
                        float g6[6];
                        philox_gauss6_f32((uint32_t)seed, (uint32_t)(seed >> 32), j >> 1, pid, member_id(P, rq),
                                          bm, g6);
                        cw[q] = odd ? Inc{g6[3], g6[4], g6[5]} : Inc{g6[0], g6[1], g6[2]};
                        carry[q][0] = g6[3]; carry[q][1] = g6[4]; carry[q][2] = g6[5];
                    }
                    have_carry = !odd;
                }
            } else {
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const uint64_t rq = member(q), seed = (uint64_t)__ldg(P.seeds + rq);
                    const V3 w = draw_scaled<NOISE>(P, (uint32_t)seed, (uint32_t)(seed >> 32), j, pid,
                                                    member_id(P, rq), rq, csig, bm);
                    cw[q] = Inc{w.x, w.y, w.z};
                }
            }
            advance(nt, cw, j);
        }
    };
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        const uint64_t tgt = (k < P.k1) ? P.target[k] : P.j1;
        if (!TAIL || n_tiles == 2) steps_to(std::integral_constant<int, 2>{}, tgt);
        else if (n_tiles == 1) steps_to(std::integral_constant<int, 1>{}, tgt);
        else j = tgt;   // a member half without members only joins the CTA-wide sample reduction
        if (k < P.k1) {
            if (P.traj != nullptr && valid) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (!is_live(q)) continue;
                    double* o = P.traj + ((uint64_t)k * 3 * N + 3ull * pid) * P.R + member(q);
                    o[0] = m[q].x; o[P.R] = m[q].y; o[2 * P.R] = m[q].z;
                }
            }
            if (P.partial != nullptr) {
                // cluster magnetisation of each member: particles of a group by shuffles over g, groups in fixed order
                __syncthreads();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    double sx = valid ? m[q].x : 0.0, sy = valid ? m[q].y : 0.0, sz = valid ? m[q].z : 0.0;
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        sx += __shfl_xor_sync(0xffffffffu, sx, o);
                        sy += __shfl_xor_sync(0xffffffffu, sy, o);
                        sz += __shfl_xor_sync(0xffffffffu, sz, o);
                    }
                    if (g == 0) {
                        const int col = 16 * mh + 8 * (q >> 1) + 2 * t + (q & 1);
                        double* rr = sm_red + (size_t)pg * 3 * MB + col;
                        const bool lv = is_live(q);
                        rr[0] = lv ? sx : 0.0; rr[MB] = lv ? sy : 0.0; rr[2 * MB] = lv ? sz : 0.0;
                    }
                }
                __syncthreads();
                if (warp == 0) {
                    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                    for (int col = lane; col < MB; col += 32) {
                        double Mx = 0, My = 0, Mz = 0;
                        for (int s2 = 0; s2 < G; ++s2) {
                            const double* q2 = sm_red + (size_t)s2 * 3 * MB + col;
                            Mx += q2[0]; My += q2[MB]; Mz += q2[2 * MB];
                        }
                        v0 += Mx; v1 += My; v2 += Mz; v3 += Mz * Mz;
                    }
                    v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
                    if (lane == 0) {
                        double* o = P.partial + ((uint64_t)(k - P.k0) * P.cta_total + cta) * 4;
                        o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
                    }
                }
                __syncthreads();
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!is_live(q)) continue;
            const uint64_t c0 = 3ull * pid, rq = member(q);
            P.state[c0 * P.R + rq] = m[q].x; P.state[(c0 + 1) * P.R + rq] = m[q].y; P.state[(c0 + 2) * P.R + rq] = m[q].z;
        }
    }
}

template <int NOISE, bool TAB>
static cudaError_t launch_hm(bool one_buf, unsigned full_ctas, unsigned tail_ctas, unsigned threads, size_t smem, cudaStream_t s,
                             RunParams P) {
    auto go = [&](auto kernel, unsigned grid, unsigned offset) -> cudaError_t {
        if (grid == 0) return cudaSuccess;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        P.cta_offset = offset;
        kernel<<<grid, threads, smem, s>>>(P);
        return cudaGetLastError();
    };
    P.cta_total = full_ctas + tail_ctas;
    if (P.mma_dglobal) {   // N = 65..128: matrix in global memory, always two moment buffers
        cudaError_t e = go(heun_cluster_mma_kernel<NOISE, TAB, false, false, true>, full_ctas, 0);
        if (e != cudaSuccess) return e;
        return go(heun_cluster_mma_kernel<NOISE, TAB, false, true, true>, tail_ctas, full_ctas);
    }
    cudaError_t e = one_buf ? go(heun_cluster_mma_kernel<NOISE, TAB, true, false, false>, full_ctas, 0)
                            : go(heun_cluster_mma_kernel<NOISE, TAB, false, false, false>, full_ctas, 0);
    if (e != cudaSuccess) return e;
    return one_buf ? go(heun_cluster_mma_kernel<NOISE, TAB, true, true, false>, tail_ctas, full_ctas)
                   : go(heun_cluster_mma_kernel<NOISE, TAB, false, true, false>, tail_ctas, full_ctas);
}

cudaError_t launch_heun_cluster_mma(int noise, bool tab, bool one_buf, unsigned full_ctas, unsigned tail_ctas, unsigned threads,
                                    size_t smem, cudaStream_t s, const RunParams& P) {
    switch (noise) {
        case NOISE_PHILOX_F32: return tab ? launch_hm<NOISE_PHILOX_F32, true>(one_buf, full_ctas, tail_ctas, threads, smem, s, P)
                                          : launch_hm<NOISE_PHILOX_F32, false>(one_buf, full_ctas, tail_ctas, threads, smem, s, P);
        case NOISE_PHILOX_F64: return tab ? launch_hm<NOISE_PHILOX_F64, true>(one_buf, full_ctas, tail_ctas, threads, smem, s, P)
                                          : launch_hm<NOISE_PHILOX_F64, false>(one_buf, full_ctas, tail_ctas, threads, smem, s, P);
        case NOISE_INJECTED: return tab ? launch_hm<NOISE_INJECTED, true>(one_buf, full_ctas, tail_ctas, threads, smem, s, P)
                                        : launch_hm<NOISE_INJECTED, false>(one_buf, full_ctas, tail_ctas, threads, smem, s, P);
        default: return tab ? launch_hm<NOISE_PHILOX_PACKED, true>(one_buf, full_ctas, tail_ctas, threads, smem, s, P)
                            : launch_hm<NOISE_PHILOX_PACKED, false>(one_buf, full_ctas, tail_ctas, threads, smem, s, P);
    }
}

}  // namespace mb
