// cluster_mma_imid.cu — K4m: implicit midpoint for interacting clusters (lib/integrators.cpp:576-651 +
// lib/optimisation.cpp:81-149 over the N-particle LLG system of lib/llg.cpp:453-481) with the dipolar field of every
// quasi-Newton iteration evaluated as the matrix product of cluster_mma.cu / mma.cuh on DMMA.8x8x4.
//
// The reference's Newton matrix is block diagonal (a' and B' only ever hold 3x3 / 3x3x3 diagonal blocks and the dipolar
// field has no Jacobian, lib/simulation.cpp:292-303): what couples the particles of a member inside an iteration is
// (i) the dipolar field in the residual F — ONE product D . X per iteration — and (ii) the two 3N-wide norms
// (tolerance eps |X_0|, error |delta|).  Mapping: a warp owns one particle group of 8 (three 8-row tiles of the product)
// and ONE 8-member column tile, i.e. thread (g, t) holds particle 8 pg + g of the two members 8 ct + 2 t + {0, 1}: all of
// a member's per-particle Newton work (residual, 3x3 matrix, adjugate solve) stays in registers (two members per thread =
// two independent dependent chains), the iterate X lives in ONE shared-memory moment buffer (the step's initial state
// stays in registers; it is multiplied once, for the initial guess), and the G warps that share a column tile are an
// independent problem with their own named barrier — 8 members vote on convergence instead of the 32 of the scalar
// kernel (cluster.cu), and a vote costs nothing: every warp of the group sums the same shared-memory partials in the
// same order, so all of them take the same decision without exchanging it.
// Iterates, tolerance test and iteration counts are those of the reference (identical counts vs the oracle,
// tests/test_parity_gpu.py); `implicit_newton = exact` swaps the matrix as in the other implicit kernels.
#include "common.cuh"
#include "launch.h"
#include "mma.cuh"

namespace mb {

// DG: the packed matrix is read from global memory / L2 (65..128 particles)
template <int NOISE, bool FIELD_TAB, bool DG>
__global__ void __launch_bounds__(512, 1) imid_cluster_mma_kernel(const __grid_constant__ RunParams P) {
    extern __shared__ double smem[];
    const int N = (int)P.N, G = (int)P.G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int n_warps = blockDim.x >> 5, CT = n_warps / G, MB = 8 * CT, LD = MB + 4;
    const int pg = warp % G, ct = warp / G;
    const int n_blk = G * (G + 1) / 2;
    double* sm_d = smem;                                              // [n_blk][576] (not staged when DG)
    double* sm_x = sm_d + (DG ? 0 : (size_t)n_blk * MMA_BLK);         // [24 G][LD] moments entering the product (times v_red)
    double* sm_red = sm_x + (size_t)24 * G * LD;                      // [3][G][MB] norm partials / flags / sample sums
    {
        const int nd = n_blk * MMA_BLK;
        if (!DG)
            for (int q = threadIdx.x; q < nd; q += blockDim.x) sm_d[q] = P.dmat[q];
        const int nm = 24 * G * LD;
        for (int q = threadIdx.x; q < nm; q += blockDim.x) sm_x[q] = 0.0;   // rows of padding particles stay zero
    }
    __syncthreads();

    const int p_raw = 8 * pg + g;
    const bool valid = p_raw < N;
    const uint32_t pid = valid ? (uint32_t)p_raw : 0u;
    const double alpha = P.alpha, dt = P.dt, clampA = P.clampA, sqrt_dt = P.sqrt_dt, eps2 = P.eps * P.eps;
    const bool renorm = P.renorm != 0, inter = P.interactions != 0, exact = P.newton_exact != 0, zero_u = P.quirk_zero != 0;
    const double kred = __ldg(P.k_red + pid), vred = __ldg(P.v_red + pid), sr = __ldg(P.sig + pid), k0 = __ldg(P.k_red);
    const bool mono = P.mma_mono != 0;
    const int bar_id = 1 + ct, bar_n = 32 * G;

    // the thread's two members: local column 8 ct + 2 t + e; columns past the ensemble repeat the last member (never stored)
    const uint64_t r_first = (uint64_t)blockIdx.x * MB + 8 * ct + 2 * t;
    uint64_t rr[2];
    bool live[2];
    uint32_t key0[2], key1[2], mid[2];
    V3 m[2];
    // easy axes are re-read (L1 hits) where they are used instead of living in 36 registers across the iteration: the
    // own particle's axis and particle 0's, whose rank-one "field Jacobian" block the reference reads (quirk_u)
    auto axis_of = [&](const uint32_t particle, const int e) {
        const uint64_t c0 = 3ull * particle, ar = rr[e] * P.axis_rs;
        return V3{__ldg(P.axis + c0 * P.axis_cs + ar), __ldg(P.axis + (c0 + 1) * P.axis_cs + ar),
                  __ldg(P.axis + (c0 + 2) * P.axis_cs + ar)};
    };
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        live[e] = r_first + e < P.R;
        rr[e] = live[e] ? r_first + e : P.R - 1;
        const uint64_t seed = (uint64_t)__ldg(P.seeds + rr[e]);
        key0[e] = (uint32_t)seed; key1[e] = (uint32_t)(seed >> 32);
        mid[e] = member_id(P, rr[e]);
        const uint64_t c0 = 3ull * pid;
        m[e] = valid ? V3{P.state[c0 * P.R + rr[e]], P.state[(c0 + 1) * P.R + rr[e]], P.state[(c0 + 2) * P.R + rr[e]]}
                     : V3{0.0, 0.0, 0.0};
    }
    const int own_off = (24 * pg + g) * LD + 8 * ct + 2 * t;
    const int b_idx = (int)(sm_x - smem) + t * LD + 8 * ct + g;
    auto put = [&](const V3 (&x)[2]) {
        if (!valid) return;
        double* d = sm_x + own_off;
        const double s = mono ? 1.0 : vred;
        *reinterpret_cast<double2*>(d) = make_double2(s * x[0].x, s * x[1].x);
        *reinterpret_cast<double2*>(d + 8 * LD) = make_double2(s * x[0].y, s * x[1].y);
        *reinterpret_cast<double2*>(d + 16 * LD) = make_double2(s * x[0].z, s * x[1].z);
    };
    // sum over the N particles of a member of one value per (particle, member): the 8 particles of a warp by shuffles
    // over g, the G warps of the group through shared memory (fixed order -> every warp of the group gets the same bits)
    auto group_sum2 = [&](double (&v)[2], double* slab) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) v[e] += __shfl_xor_sync(0xffffffffu, v[e], o);
        if (g == 0) *reinterpret_cast<double2*>(slab + pg * MB + 8 * ct + 2 * t) = make_double2(v[0], v[1]);
    };
    auto group_read2 = [&](double (&v)[2], const double* slab) {
        v[0] = v[1] = 0.0;
        for (int s2 = 0; s2 < G; ++s2) {
            const double2 w = *reinterpret_cast<const double2*>(slab + s2 * MB + 8 * ct + 2 * t);
            v[0] += w.x; v[1] += w.y;
        }
    };
    double* red_a = sm_red;
    double* red_b = sm_red + (size_t)G * MB;
    double* red_c = sm_red + (size_t)2 * G * MB;   // the tolerance norm has its own slab: its readers are not fenced off
                                                   // from the first iteration's writers of red_a by a barrier
    // field of the own particle of member e at moments x: anisotropy + applied + dipolar (acc)
    auto field = [&](const int e, const V3& ax, const V3& x, const double (&acc)[3][2][2], const double hz) {
        const double s = dot(x, ax) * kred;
        return V3{fma(s, ax.x, acc[0][0][e]), fma(s, ax.y, acc[1][0][e]), fma(s, ax.z, hz) + acc[2][0][e]};
    };

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
            V3 sw[2], X[2];
            double acc[3][2][2];
            // ---- initial guess: Euler step from x0 (lib/integrators.cpp:605-614) ----
            put(m);
            group_barrier(bar_id, bar_n);
            if (inter) dipolar_mma<1, DG>(acc, smem, P.dmat, b_idx, G, pg, LD, g, t);
            else {
#pragma unroll
                for (int a = 0; a < 3; ++a) acc[a][0][0] = acc[a][0][1] = 0.0;
            }
            double part[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                V3 w{0.0, 0.0, 0.0};
                if (valid) w = draw_noise<NOISE>(P, key0[e], key1[e], j, pid, mid[e], rr[e]);
                const V3 wm{fmax(-clampA, fmin(clampA, w.x)) * sqrt_dt, fmax(-clampA, fmin(clampA, w.y)) * sqrt_dt,
                            fmax(-clampA, fmin(clampA, w.z)) * sqrt_dt};
                sw[e] = V3{sr * wm.x, sr * wm.y, sr * wm.z};
                const V3 h = field(e, axis_of(pid, e), m[e], acc, hz0);
                const V3 gg{fma(h.x, dt, sw[e].x), fma(h.y, dt, sw[e].y), fma(h.z, dt, sw[e].z)};
                const V3 f = llg_f(m[e], gg, alpha);
                X[e] = exact ? V3{fma(0.5, f.x, m[e].x), fma(0.5, f.y, m[e].y), fma(0.5, f.z, m[e].z)}
                             : V3{(f.x + m[e].x) / 2, (f.y + m[e].y) / 2, (f.z + m[e].z) / 2};
                part[e] = valid ? dot(X[e], X[e]) : 0.0;
            }
            group_barrier(bar_id, bar_n);          // every warp of the group has multiplied x0
            put(X);
            group_sum2(part, red_c);
            group_barrier(bar_id, bar_n);          // X and the norm partials are visible
            double tol[2], err[2];
            group_read2(tol, red_c);
            // err > tol is tested on the squares (no square root in the dependent chain of an iteration)
            int iter[2] = {1000, 1000};
            unsigned long long done[2] = {0ull, 0ull};
            bool singular[2] = {false, false};
#pragma unroll
            for (int e = 0; e < 2; ++e) { tol[e] *= eps2; err[e] = 4 * tol[e]; }
            // ---- quasi-Newton iteration (lib/optimisation.cpp:81-149), all members of the column tile in lock step ----
            while (true) {
                bool active[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    active[e] = (err[e] > tol[e]) && !singular[e];
                    if (active[e]) { active[e] = iter[e] > 0; --iter[e]; }
                }
                // identical in every warp of the group: all of them read the same partial sums
                if (!__any_sync(0xffffffffu, active[0] || active[1])) break;
                if (inter) dipolar_mma<1, DG>(acc, smem, P.dmat, b_idx, G, pg, LD, g, t);
                V3 dl[2];
                double bad[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const V3 e_own = axis_of(pid, e);
                    const V3 h = field(e, e_own, X[e], acc, hz1);
                    const V3 gg{fma(h.x, dt, sw[e].x), fma(h.y, dt, sw[e].y), fma(h.z, dt, sw[e].z)};
                    const V3 f = llg_f(X[e], gg, alpha);
                    double b[3] = {-(X[e].x - m[e].x - 0.5 * f.x), -(X[e].y - m[e].y - 0.5 * f.y),
                                   -(X[e].z - m[e].z - 0.5 * f.z)};
                    double A[9], d[3];
                    if (exact) {   // opt-in: each particle's exact own Jacobian; the dipolar coupling stays out of the matrix
                        const V3 pc = cross(X[e], gg);
                        const V3 u{fma(alpha, pc.x, gg.x), fma(alpha, pc.y, gg.y), fma(alpha, pc.z, gg.z)};
                        newton_matrix_exact(A, X[e], alpha, gg, u, dt * kred, e_own);
                    } else {
                        const V3 e_zero = axis_of(0u, e);
                        newton_matrix(A, X[e], alpha, h, sw[e], quirk_u((unsigned)N, pid, e_zero, k0), e_zero, zero_u);
                    }
                    bool ok = true;
                    if (!solve3_adjugate(A, b, d)) { ok = false; d[0] = d[1] = d[2] = 0.0; }
                    dl[e] = V3{d[0], d[1], d[2]};
                    part[e] = valid ? d[0] * d[0] + d[1] * d[1] + d[2] * d[2] : 0.0;
                    bad[e] = (valid && !ok) ? 1.0 : 0.0;
                }
                group_sum2(part, red_a);
                group_sum2(bad, red_b);
                group_barrier(bar_id, bar_n);      // products done, partials visible
                double e2[2], nb[2];
                group_read2(e2, red_a);
                group_read2(nb, red_b);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (!active[e]) continue;
                    ++done[e];
                    if (nb[e] != 0.0) {
                        singular[e] = true;
                    } else {
                        err[e] = e2[e];
                        X[e].x += dl[e].x; X[e].y += dl[e].y; X[e].z += dl[e].z;
                    }
                }
                put(X);
                group_barrier(bar_id, bar_n);      // the new iterate is visible; the partials may be overwritten
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (pg == 0 && g == 0 && live[e]) {
                    nc.total += done[e];
                    nc.worst = done[e] > nc.worst ? done[e] : nc.worst;
                    nc.fails += (singular[e] || iter[e] == -1) ? 1ull : 0ull;
                }
                m[e] = V3{2 * X[e].x - m[e].x, 2 * X[e].y - m[e].y, 2 * X[e].z - m[e].z};
                if (renorm && valid) renormalise(m[e]);
            }
        }
        if (k < P.k1) {
            if (P.traj != nullptr && valid) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (!live[e]) continue;
                    double* o = P.traj + ((uint64_t)k * 3 * N + 3ull * pid) * P.R + rr[e];
                    o[0] = m[e].x; o[P.R] = m[e].y; o[2 * P.R] = m[e].z;
                }
            }
            if (P.partial != nullptr) {
                // cluster magnetisation of each member: particles of a group by shuffles over g, groups in fixed order
                __syncthreads();
                double sx[2], sy[2], sz[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool c = valid && live[e];
                    sx[e] = c ? m[e].x : 0.0; sy[e] = c ? m[e].y : 0.0; sz[e] = c ? m[e].z : 0.0;
                }
                group_sum2(sx, sm_red);
                group_sum2(sy, sm_red + (size_t)G * MB);
                group_sum2(sz, sm_red + (size_t)2 * G * MB);
                __syncthreads();
                if (warp == 0) {
                    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                    for (int col = lane; col < MB; col += 32) {
                        double Mx = 0, My = 0, Mz = 0;
                        for (int s2 = 0; s2 < G; ++s2) {
                            Mx += sm_red[s2 * MB + col];
                            My += sm_red[(size_t)(G + s2) * MB + col];
                            Mz += sm_red[(size_t)(2 * G + s2) * MB + col];
                        }
                        v0 += Mx; v1 += My; v2 += Mz; v3 += Mz * Mz;
                    }
                    v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
                    if (lane == 0) {
                        double* o = P.partial + ((uint64_t)(k - P.k0) * gridDim.x + blockIdx.x) * 4;
                        o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
                    }
                }
                __syncthreads();
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if (!live[e]) continue;
            const uint64_t c0 = 3ull * pid;
            P.state[c0 * P.R + rr[e]] = m[e].x; P.state[(c0 + 1) * P.R + rr[e]] = m[e].y; P.state[(c0 + 2) * P.R + rr[e]] = m[e].z;
        }
    }
    newton_flush(P, nc, true);
}

template <int NOISE, bool TAB>
static cudaError_t launch_im(unsigned grid, unsigned threads, size_t smem, cudaStream_t s, const RunParams& P) {
    auto go = [&](auto kernel) -> cudaError_t {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kernel<<<grid, threads, smem, s>>>(P);
        return cudaGetLastError();
    };
    return P.mma_dglobal ? go(imid_cluster_mma_kernel<NOISE, TAB, true>) : go(imid_cluster_mma_kernel<NOISE, TAB, false>);
}

cudaError_t launch_imid_cluster_mma(int noise, bool tab, unsigned grid, unsigned threads, size_t smem, cudaStream_t s,
                                    const RunParams& P) {
    switch (noise) {
        case NOISE_PHILOX_F32: return tab ? launch_im<NOISE_PHILOX_F32, true>(grid, threads, smem, s, P)
                                          : launch_im<NOISE_PHILOX_F32, false>(grid, threads, smem, s, P);
        case NOISE_PHILOX_F64: return tab ? launch_im<NOISE_PHILOX_F64, true>(grid, threads, smem, s, P)
                                          : launch_im<NOISE_PHILOX_F64, false>(grid, threads, smem, s, P);
        case NOISE_INJECTED: return tab ? launch_im<NOISE_INJECTED, true>(grid, threads, smem, s, P)
                                        : launch_im<NOISE_INJECTED, false>(grid, threads, smem, s, P);
        default: return tab ? launch_im<NOISE_PHILOX_PACKED, true>(grid, threads, smem, s, P)
                            : launch_im<NOISE_PHILOX_PACKED, false>(grid, threads, smem, s, P);
    }
}

}  // namespace mb
