#include "common.cuh"
#include "launch.h"

#ifndef MB_DIP_UNROLL
#define MB_DIP_UNROLL 2   // unroll factor of the j loop of the dipolar sum (tuning knob)
#endif
#define MB_PRAGMA_STR(x) #x
#define MB_PRAGMA(x) _Pragma(MB_PRAGMA_STR(x))
#define MB_DIP_UNROLL_PRAGMA MB_PRAGMA(unroll MB_DIP_UNROLL)

namespace mb {

// ---------------------------------------------------------------------------------
// K2 / K4: interacting clusters.  blockDim = (32 members, PS particle slots); thread
// (lane, slot) owns particles slot, slot+PS, ... (NP of them, compile time).  The
// moments of all N particles of the CTA's 32 members live in shared memory as
// [N][3][32] so the j-loop of the dipolar sum reads conflict-free rows, while the
// static pair table {r_hat, c_ij} is the same address for the whole warp (one
// broadcast transaction).
// ---------------------------------------------------------------------------------

struct Own {  // per-thread, per-owned-particle constants (implicit kernel)
    V3 e;
    double kred, sr;
    uint32_t p;
    bool valid;
};

// Dipolar field on the NP own particles of this thread from the moments of all N particles in
// shared memory (lib/field.cpp:187-225).  j is the OUTER loop so that m_j is read once per thread and
// reused for all own particles; the static pair table {sqrt(3) r_hat_ij, c_ij} is a warp-uniform
// read-only load (one sector per warp), its diagonal is zero so no j == i branch is needed, and
// (m.t) t with t = sqrt(3) r_hat is 3 (m.r_hat) r_hat: 9 fp64 operations per ordered pair.
template <int NP, bool TAB_SMEM = false>
__device__ __forceinline__ void add_dipolar(V3 (&h)[NP], const uint32_t (&p)[NP], const double* sm, const double* dip,
                                            const uint32_t N, const int lane) {
    const double2* row[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) row[q] = reinterpret_cast<const double2*>(dip) + (uint64_t)p[q] * N * 2;
    MB_DIP_UNROLL_PRAGMA
    for (uint32_t jq = 0; jq < N; ++jq) {
        const double* mj = sm + (uint64_t)jq * 3 * CL_LANES + lane;
        const double mx = mj[0], my = mj[CL_LANES], mz = mj[2 * CL_LANES];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            double2 t0, t1;
            if (TAB_SMEM) { t0 = row[q][2 * jq]; t1 = row[q][2 * jq + 1]; }          // LDS.128 broadcast
            else { t0 = __ldg(row[q] + 2 * jq); t1 = __ldg(row[q] + 2 * jq + 1); }
            const double d = mx * t0.x + my * t0.y + mz * t1.x;
            h[q].x = fma(t1.y, fma(d, t0.x, -mx), h[q].x);
            h[q].y = fma(t1.y, fma(d, t0.y, -my), h[q].y);
            h[q].z = fma(t1.y, fma(d, t1.x, -mz), h[q].z);
        }
    }
}

// effective field of one own particle given all moments in shared memory (lib/simulation.cpp:271-290)
// (`dip_smem` = the pair table staged in shared memory)
__device__ __forceinline__ V3 cluster_field(const RunParams& P, const double* sm /*[N][3][32]*/, const double* dip_smem,
                                            const Own& o, const V3& m, const double hz, const int lane) {
    const double s = dot(m, o.e) * o.kred;
    V3 h[1] = {V3{s * o.e.x, s * o.e.y, fma(s, o.e.z, hz)}};
    if (P.interactions) {
        const uint32_t p[1] = {o.p};
        add_dipolar<1, true>(h, p, sm, dip_smem, P.N, lane);
    }
    return h[0];
}

// effective fields of ALL own particles of the implicit kernel: anisotropy + applied, then one pass of the
// j-outer dipolar sum that reads every shared-memory moment once for all of them
template <int NP>
__device__ __forceinline__ void own_fields(V3 (&h)[NP], const RunParams& P, const double* sm, const double* dip_smem,
                                           const Own (&own)[NP], const V3 (&x)[NP], const double hz, const int lane) {
    uint32_t pid[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const double s = dot(x[q], own[q].e) * own[q].kred;
        h[q] = V3{s * own[q].e.x, s * own[q].e.y, fma(s, own[q].e.z, hz)};
        pid[q] = own[q].p;
    }
    if (P.interactions) add_dipolar<NP, true>(h, pid, sm, dip_smem, P.N, lane);
}

// K2: Heun.  Per step: fields of the own particles from sm_m -> predictor moments into sm_t -> barrier ->
// fields from sm_t -> corrected moments into sm_m -> barrier.  Same fused arithmetic as K1.
//
// LAYOUT 0: pair table read from global memory (N > 64), two moment buffers
// LAYOUT 1: pair table staged in shared memory, two moment buffers
// LAYOUT 2: pair table staged in shared memory, ONE moment buffer (predictor overwrites the current
//           moments between two extra barriers; every thread keeps its own moments in registers) — this
//           is what lets the 128 KB table of a 64-particle cluster sit next to the moments in 227 KB
template <int NOISE, bool FIELD_TAB, int NP, int LAYOUT>
__global__ void __launch_bounds__(512) heun_cluster_kernel(const __grid_constant__ RunParams P) {
    extern __shared__ double smem[];
    const uint32_t N = P.N;
    constexpr bool TAB_SMEM = LAYOUT != 0, ONE_BUF = LAYOUT == 2;
    const int lane = threadIdx.x, slot = threadIdx.y, PS = blockDim.y;
    double* sm_m = smem;                                                    // [N][3][32] current moments
    double* sm_t = ONE_BUF ? sm_m : smem + (uint64_t)N * 3 * CL_LANES;      // [N][3][32] predictor moments
    double* sm_red = sm_t + (uint64_t)N * 3 * CL_LANES;                     // [PS][3][32] sample reduction
    double* sm_tab = sm_red + (uint64_t)PS * 3 * CL_LANES;                  // [N][N][4] pair table
    if (TAB_SMEM) {
        const uint32_t n4 = N * N * 4, tid = slot * CL_LANES + lane, nt = PS * CL_LANES;
        for (uint32_t q = tid; q < n4; q += nt) sm_tab[q] = P.dip[q];
    }
    const double* dip = TAB_SMEM ? sm_tab : P.dip;
    const uint64_t r_raw = (uint64_t)blockIdx.x * CL_LANES + lane;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    const double alpha = P.alpha, dt = P.dt;
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = member_id(P, r);
    const bool renorm = P.renorm != 0, inter = P.interactions != 0;
    // consume the packed noise stream two steps per Philox block when few particles are owned; with many
    // owned particles the dipolar sum dwarfs the generator and the carry registers are worth more
    constexpr bool PAIRWISE = NOISE == NOISE_PHILOX_PACKED && NP <= 2;

    uint32_t pid[NP];
    bool valid[NP];
    V3 m[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const uint32_t p = slot + q * PS;
        valid[q] = p < N;
        pid[q] = valid[q] ? p : 0;
        const uint64_t c0 = 3ull * pid[q];
        m[q] = V3{P.state[c0 * P.R + r], P.state[(c0 + 1) * P.R + r], P.state[(c0 + 2) * P.R + r]};
        if (valid[q]) {
            double* d = sm_m + c0 * CL_LANES + lane;
            d[0] = m[q].x; d[CL_LANES] = m[q].y; d[2 * CL_LANES] = m[q].z;
        }
    }
    __syncthreads();

    // g_q = dt h_q(all moments in `sm`, own moment x_q) + cw_q
    auto stage_g = [&](V3 (&g)[NP], const V3 (&x)[NP], const double* sm, const double hz, const V3 (&cw)[NP]) {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const uint64_t c0 = 3ull * pid[q];
            const V3 e{__ldg(P.axis + c0 * P.axis_cs + r * P.axis_rs), __ldg(P.axis + (c0 + 1) * P.axis_cs + r * P.axis_rs),
                       __ldg(P.axis + (c0 + 2) * P.axis_cs + r * P.axis_rs)};
            const double s = dot(x[q], e) * __ldg(P.k_red + pid[q]);
            g[q] = V3{s * e.x, s * e.y, fma(s, e.z, hz)};
        }
        if (inter) add_dipolar<NP, TAB_SMEM>(g, pid, sm, dip, N, lane);
#pragma unroll
        for (int q = 0; q < NP; ++q)
            g[q] = V3{fma(g[q].x, dt, cw[q].x), fma(g[q].y, dt, cw[q].y), fma(g[q].z, dt, cw[q].z)};
    };

    auto advance = [&](const V3 (&cw)[NP], const uint64_t jj) {
        double hz0 = P.h_const, hz1 = P.h_const;
        if (FIELD_TAB) {
            const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (jj - P.j0));
            hz0 = h.x; hz1 = h.y;
        }
        V3 g[NP], mt[NP];
        stage_g(g, m, sm_m, hz0, cw);
        if (ONE_BUF) __syncthreads();   // every thread has read the current moments
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const V3 pc = cross(m[q], g[q]);
            const V3 u{fma(alpha, pc.x, g[q].x), fma(alpha, pc.y, g[q].y), fma(alpha, pc.z, g[q].z)};
            mt[q] = V3{fma(-m[q].y, u.z, fma(m[q].z, u.y, m[q].x)), fma(-m[q].z, u.x, fma(m[q].x, u.z, m[q].y)),
                       fma(-m[q].x, u.y, fma(m[q].y, u.x, m[q].z))};
            if (valid[q]) {
                double* d = sm_t + 3ull * pid[q] * CL_LANES + lane;
                d[0] = mt[q].x; d[CL_LANES] = mt[q].y; d[2 * CL_LANES] = mt[q].z;
            }
        }
        __syncthreads();
        stage_g(g, mt, sm_t, hz1, cw);
        if (ONE_BUF) __syncthreads();   // every thread has read the predictor moments
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const V3 pc = cross(mt[q], g[q]);
            const V3 u{fma(alpha, pc.x, g[q].x), fma(alpha, pc.y, g[q].y), fma(alpha, pc.z, g[q].z)};
            const V3 hm{0.5 * mt[q].x, 0.5 * mt[q].y, 0.5 * mt[q].z};
            const V3 hh{fma(0.5, m[q].x, hm.x), fma(0.5, m[q].y, hm.y), fma(0.5, m[q].z, hm.z)};
            m[q] = V3{fma(-hm.y, u.z, fma(hm.z, u.y, hh.x)), fma(-hm.z, u.x, fma(hm.x, u.z, hh.y)),
                      fma(-hm.x, u.y, fma(hm.y, u.x, hh.z))};
            if (renorm) renormalise(m[q]);
            if (valid[q]) {
                double* d = sm_m + 3ull * pid[q] * CL_LANES + lane;
                d[0] = m[q].x; d[CL_LANES] = m[q].y; d[2 * CL_LANES] = m[q].z;
            }
        }
        __syncthreads();
    };

    auto bm_of = [&](const int q) { return scale_to_bm(__ldg(P.sig + pid[q]) * P.sqrt_dt); };

    uint64_t j = P.j0;
    float carry[PAIRWISE ? NP : 1][3];
    if (PAIRWISE && (j & 1)) {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            float g6[6];
            philox_gauss6_f32(key0, key1, j >> 1, pid[q], member, bm_of(q), g6);
            carry[q][0] = g6[3]; carry[q][1] = g6[4]; carry[q][2] = g6[5];
        }
    }
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        const uint64_t tgt = (k < P.k1) ? P.target[k] : P.j1;
        if (PAIRWISE) {
            V3 cw[NP];
            if ((j & 1) && j < tgt) {
#pragma unroll
                for (int q = 0; q < NP; ++q)
                    cw[q] = V3{widen_f32(carry[q][0]), widen_f32(carry[q][1]), widen_f32(carry[q][2])};
                advance(cw, j);
                ++j;
            }
            for (; j + 2 <= tgt; j += 2) {
                float g6[NP][6];
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    philox_gauss6_f32(key0, key1, j >> 1, pid[q], member, bm_of(q), g6[q]);
                    cw[q] = V3{widen_f32(g6[q][0]), widen_f32(g6[q][1]), widen_f32(g6[q][2])};
                }
                advance(cw, j);
#pragma unroll
                for (int q = 0; q < NP; ++q) cw[q] = V3{widen_f32(g6[q][3]), widen_f32(g6[q][4]), widen_f32(g6[q][5])};
                advance(cw, j + 1);
            }
            if (j < tgt) {
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    float g6[6];
                    philox_gauss6_f32(key0, key1, j >> 1, pid[q], member, bm_of(q), g6);
                    cw[q] = V3{widen_f32(g6[0]), widen_f32(g6[1]), widen_f32(g6[2])};
                    carry[q][0] = g6[3]; carry[q][1] = g6[4]; carry[q][2] = g6[5];
                }
                advance(cw, j);
                ++j;
            }
        } else {
            for (; j < tgt; ++j) {
                V3 cw[NP];
#pragma unroll
                for (int q = 0; q < NP; ++q)
                    cw[q] = draw_scaled<NOISE>(P, key0, key1, j, pid[q], member, r, __ldg(P.sig + pid[q]) * P.sqrt_dt, bm_of(q));
                advance(cw, j);
            }
        }
        if (k < P.k1) {
            double sx = 0, sy = 0, sz = 0;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (!valid[q]) continue;
                if (P.traj != nullptr && live) {
                    double* t = P.traj + ((uint64_t)k * 3 * N + 3ull * pid[q]) * P.R + r;
                    t[0] = m[q].x; t[P.R] = m[q].y; t[2 * P.R] = m[q].z;
                }
                sx += m[q].x; sy += m[q].y; sz += m[q].z;
            }
            if (P.partial != nullptr) {
                // cluster magnetisation of each member: fixed-order sum over the particle slots
                double* rr = sm_red + (uint64_t)slot * 3 * CL_LANES + lane;
                rr[0] = sx; rr[CL_LANES] = sy; rr[2 * CL_LANES] = sz;
                __syncthreads();
                if (slot == 0) {
                    double Mx = 0, My = 0, Mz = 0;
                    for (int s2 = 0; s2 < PS; ++s2) {
                        const double* q2 = sm_red + (uint64_t)s2 * 3 * CL_LANES + lane;
                        Mx += q2[0]; My += q2[CL_LANES]; Mz += q2[2 * CL_LANES];
                    }
                    if (!live) { Mx = 0; My = 0; Mz = 0; }
                    const double v0 = warp_sum(Mx), v1 = warp_sum(My), v2 = warp_sum(Mz), v3 = warp_sum(Mz * Mz);
                    if (lane == 0) {
                        double* o = P.partial + ((uint64_t)(k - P.k0) * gridDim.x + blockIdx.x) * 4;
                        o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
                    }
                }
                __syncthreads();
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        if (!valid[q] || !live) continue;
        const uint64_t c0 = 3ull * pid[q];
        P.state[c0 * P.R + r] = m[q].x; P.state[(c0 + 1) * P.R + r] = m[q].y; P.state[(c0 + 2) * P.R + r] = m[q].z;
    }
}

// K4: implicit midpoint for clusters.  The reference's J is block diagonal (a' and B' are
// only ever written on the 3x3 / 3x3x3 diagonal blocks of zero-filled arrays,
// lib/simulation.cpp:189-195, lib/llg.cpp:378-427, and the dipolar field has no Jacobian,
// lib/simulation.cpp:292-303), so dgesv on the 3N system is N independent pivoted 3x3
// solves; what couples the particles is the dipolar field inside F and the two 3N-wide
// norms (tolerance and error).
template <int NOISE, bool FIELD_TAB, int NP>
__global__ void __launch_bounds__(256) imid_cluster_kernel(const __grid_constant__ RunParams P) {
    extern __shared__ double smem[];
    const uint32_t N = P.N;
    double* sm_m = smem;                                 // [N][3][32] x0
    double* sm_x = smem + (uint64_t)N * 3 * CL_LANES;    // [N][3][32] midpoint iterate X
    double* sm_red = sm_x + (uint64_t)N * 3 * CL_LANES;  // [PS][3][32]
    const int lane = threadIdx.x, slot = threadIdx.y, PS = blockDim.y;
    double* sm_tab = sm_red + (uint64_t)PS * 3 * CL_LANES;   // [N][N][4] pair table (N <= 64: at most 128 KB)
    for (uint32_t q = slot * CL_LANES + lane; q < N * N * 4; q += PS * CL_LANES) sm_tab[q] = P.dip[q];
    const uint64_t r_raw = (uint64_t)blockIdx.x * CL_LANES + lane;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    const double alpha = P.alpha, dt = P.dt, clampA = P.clampA, sqrt_dt = P.sqrt_dt;
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = member_id(P, r);
    const bool renorm = P.renorm != 0, exact = P.newton_exact != 0, zero_u = P.quirk_zero != 0;
    NewtonCount nc{0ull, 0ull, 0ull};

    Own own[NP];
    V3 m[NP];
    V3 qu[NP];
    const V3 e0{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
    const double k0 = P.k_red[0];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const uint32_t p = slot + q * PS;
        own[q].valid = p < N;
        own[q].p = own[q].valid ? p : 0;
        const uint64_t c0 = 3ull * own[q].p;
        own[q].e = V3{P.axis[c0 * P.axis_cs + r * P.axis_rs], P.axis[(c0 + 1) * P.axis_cs + r * P.axis_rs],
                      P.axis[(c0 + 2) * P.axis_cs + r * P.axis_rs]};
        own[q].kred = P.k_red[own[q].p];
        own[q].sr = P.sig[own[q].p];
        m[q] = V3{P.state[c0 * P.R + r], P.state[(c0 + 1) * P.R + r], P.state[(c0 + 2) * P.R + r]};
        if (own[q].valid) {
            double* d = sm_m + c0 * CL_LANES + lane;
            d[0] = m[q].x; d[CL_LANES] = m[q].y; d[2 * CL_LANES] = m[q].z;
        }
        qu[q] = quirk_u(N, own[q].p, e0, k0);   // rank-one field-Jacobian block the reference reads (llg_math.cuh)
    }
    __syncthreads();

    uint64_t j = P.j0;
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        const uint64_t tgt = (k < P.k1) ? P.target[k] : P.j1;
        for (; j < tgt; ++j) {
            double hz0 = P.h_const, hz1 = P.h_const;
            if (FIELD_TAB) {
                const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (j - P.j0));
                hz0 = h.x; hz1 = h.y;
            }
            V3 X[NP], wm[NP], sw[NP];
            double part = 0.0;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (!own[q].valid) { X[q] = m[q]; continue; }   // padding slot: defined, never stored
                const V3 w = draw_noise<NOISE>(P, key0, key1, j, own[q].p, member, r);
                wm[q] = V3{fmax(-clampA, fmin(clampA, w.x)) * sqrt_dt, fmax(-clampA, fmin(clampA, w.y)) * sqrt_dt,
                           fmax(-clampA, fmin(clampA, w.z)) * sqrt_dt};
                sw[q] = V3{own[q].sr * wm[q].x, own[q].sr * wm[q].y, own[q].sr * wm[q].z};
                const V3 h = cluster_field(P, sm_m, sm_tab, own[q], m[q], hz0, lane);
                const V3 g{fma(h.x, dt, sw[q].x), fma(h.y, dt, sw[q].y), fma(h.z, dt, sw[q].z)};
                const V3 f = llg_f(m[q], g, alpha);
                X[q] = exact ? V3{fma(0.5, f.x, m[q].x), fma(0.5, f.y, m[q].y), fma(0.5, f.z, m[q].z)}
                             : V3{(f.x + m[q].x) / 2, (f.y + m[q].y) / 2, (f.z + m[q].z) / 2};
                double* d = sm_x + 3ull * own[q].p * CL_LANES + lane;
                d[0] = X[q].x; d[CL_LANES] = X[q].y; d[2 * CL_LANES] = X[q].z;
                part += dot(X[q], X[q]);
            }
            sm_red[slot * CL_LANES + lane] = part;
            __syncthreads();
            double nrm = 0.0;
            for (int s2 = 0; s2 < PS; ++s2) nrm += sm_red[s2 * CL_LANES + lane];
            // err > tol is tested on the squares (no square root in the dependent chain of an iteration)
            const double tol = (P.eps * P.eps) * nrm;
            double err = 4 * tol;
            int iter = 1000;
            unsigned long long done = 0;
            bool singular = false;
            while (true) {
                bool active = (err > tol) && !singular;
                if (active) { active = iter > 0; --iter; }
                // barrier + vote: also orders the previous iteration's sm_x / sm_red traffic
                if (!__syncthreads_or(active ? 1 : 0)) break;
                V3 dl[NP], hq[NP];
                own_fields<NP>(hq, P, sm_x, sm_tab, own, X, hz1, lane);
                bool ok = true;
                part = 0.0;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (!own[q].valid) continue;
                    const V3 h = hq[q];
                    const V3 g{fma(h.x, dt, sw[q].x), fma(h.y, dt, sw[q].y), fma(h.z, dt, sw[q].z)};
                    const V3 f = llg_f(X[q], g, alpha);
                    double b[3] = {-(X[q].x - m[q].x - 0.5 * f.x), -(X[q].y - m[q].y - 0.5 * f.y),
                                   -(X[q].z - m[q].z - 0.5 * f.z)};
                    double A[9], d[3];
                    if (exact) {   // opt-in: each particle's exact own Jacobian (llg_math.cuh), dipolar coupling left out as in the reference
                        const V3 pg = cross(X[q], g);
                        const V3 u{fma(alpha, pg.x, g.x), fma(alpha, pg.y, g.y), fma(alpha, pg.z, g.z)};
                        newton_matrix_exact(A, X[q], alpha, g, u, dt * own[q].kred, own[q].e);
                    } else {
                        newton_matrix(A, X[q], alpha, h, sw[q], qu[q], e0, zero_u);
                    }
                    if (!solve3_adjugate(A, b, d)) { ok = false; d[0] = d[1] = d[2] = 0.0; }
                    dl[q] = V3{d[0], d[1], d[2]};
                    part += d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                }
                sm_red[slot * CL_LANES + lane] = part;
                sm_red[(PS + slot) * CL_LANES + lane] = ok ? 0.0 : 1.0;
                __syncthreads();
                double e2 = 0.0, bad = 0.0;
                for (int s2 = 0; s2 < PS; ++s2) {
                    e2 += sm_red[s2 * CL_LANES + lane];
                    bad += sm_red[(PS + s2) * CL_LANES + lane];
                }
                if (active) {
                    ++done;
                    if (bad != 0.0) {
                        singular = true;
                    } else {
                        err = e2;
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            if (!own[q].valid) continue;
                            X[q].x += dl[q].x; X[q].y += dl[q].y; X[q].z += dl[q].z;
                            double* d = sm_x + 3ull * own[q].p * CL_LANES + lane;
                            d[0] = X[q].x; d[CL_LANES] = X[q].y; d[2 * CL_LANES] = X[q].z;
                        }
                    }
                }
            }
            if (slot == 0) {
                nc.total += done;
                nc.worst = done > nc.worst ? done : nc.worst;
                nc.fails += (singular || iter == -1) ? 1ull : 0ull;
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (!own[q].valid) continue;
                m[q] = V3{2 * X[q].x - m[q].x, 2 * X[q].y - m[q].y, 2 * X[q].z - m[q].z};
                if (renorm) renormalise(m[q]);
                double* d = sm_m + 3ull * own[q].p * CL_LANES + lane;
                d[0] = m[q].x; d[CL_LANES] = m[q].y; d[2 * CL_LANES] = m[q].z;
            }
            __syncthreads();
        }
        if (k < P.k1) {
            double sx = 0, sy = 0, sz = 0;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (!own[q].valid) continue;
                if (P.traj != nullptr && live) {
                    double* t = P.traj + ((uint64_t)k * 3 * N + 3ull * own[q].p) * P.R + r;
                    t[0] = m[q].x; t[P.R] = m[q].y; t[2 * P.R] = m[q].z;
                }
                sx += m[q].x; sy += m[q].y; sz += m[q].z;
            }
            if (P.partial != nullptr) {
                double* rr = sm_red + (uint64_t)slot * 3 * CL_LANES + lane;
                rr[0] = sx; rr[CL_LANES] = sy; rr[2 * CL_LANES] = sz;
                __syncthreads();
                if (slot == 0) {
                    double Mx = 0, My = 0, Mz = 0;
                    for (int s2 = 0; s2 < PS; ++s2) {
                        const double* q2 = sm_red + (uint64_t)s2 * 3 * CL_LANES + lane;
                        Mx += q2[0]; My += q2[CL_LANES]; Mz += q2[2 * CL_LANES];
                    }
                    if (!live) { Mx = 0; My = 0; Mz = 0; }
                    const double v0 = warp_sum(Mx), v1 = warp_sum(My), v2 = warp_sum(Mz), v3 = warp_sum(Mz * Mz);
                    if (lane == 0) {
                        double* o = P.partial + ((uint64_t)(k - P.k0) * gridDim.x + blockIdx.x) * 4;
                        o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
                    }
                }
                __syncthreads();
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        if (!own[q].valid || !live) continue;
        const uint64_t c0 = 3ull * own[q].p;
        P.state[c0 * P.R + r] = m[q].x; P.state[(c0 + 1) * P.R + r] = m[q].y; P.state[(c0 + 2) * P.R + r] = m[q].z;
    }
    if (slot == 0) newton_flush(P, nc, live);
}

template <class K>
static cudaError_t launch_with_smem(K kernel, dim3 g, dim3 b, size_t smem, cudaStream_t s, const RunParams& P) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<g, b, smem, s>>>(P);
    return cudaGetLastError();
}

// np / layout pairs the host may ask for (magpy_b200.cu, launch geometry): N <= 32: np 2, layout 1;
// 33..64: np 4, layout 1 or 2; 65..128: np 8, layout 0
template <int NOISE, bool TAB>
static cudaError_t launch_hc(int np, int layout, dim3 g, dim3 b, size_t smem, cudaStream_t s, const RunParams& P) {
    if (np == 2 && layout == 1) return launch_with_smem(heun_cluster_kernel<NOISE, TAB, 2, 1>, g, b, smem, s, P);
    if (np == 4 && layout == 1) return launch_with_smem(heun_cluster_kernel<NOISE, TAB, 4, 1>, g, b, smem, s, P);
    if (np == 4 && layout == 2) return launch_with_smem(heun_cluster_kernel<NOISE, TAB, 4, 2>, g, b, smem, s, P);
    if (np == 8 && layout == 0) return launch_with_smem(heun_cluster_kernel<NOISE, TAB, 8, 0>, g, b, smem, s, P);
    return cudaErrorInvalidValue;
}

template <int NOISE, bool TAB>
static cudaError_t launch_ic(int np, dim3 g, dim3 b, size_t smem, cudaStream_t s, const RunParams& P) {
    switch (np) {
        case 1: return launch_with_smem(imid_cluster_kernel<NOISE, TAB, 1>, g, b, smem, s, P);
        case 2: return launch_with_smem(imid_cluster_kernel<NOISE, TAB, 2>, g, b, smem, s, P);
        case 4: return launch_with_smem(imid_cluster_kernel<NOISE, TAB, 4>, g, b, smem, s, P);
        default: return launch_with_smem(imid_cluster_kernel<NOISE, TAB, 8>, g, b, smem, s, P);   // 33..64 particles
    }
}

#define MB_NOISE_TAB_DISPATCH(fn, ...)                                                              \
    switch (noise) {                                                                                \
        case NOISE_PHILOX_F32: return tab ? fn<NOISE_PHILOX_F32, true>(__VA_ARGS__) : fn<NOISE_PHILOX_F32, false>(__VA_ARGS__); \
        case NOISE_PHILOX_F64: return tab ? fn<NOISE_PHILOX_F64, true>(__VA_ARGS__) : fn<NOISE_PHILOX_F64, false>(__VA_ARGS__); \
        case NOISE_INJECTED: return tab ? fn<NOISE_INJECTED, true>(__VA_ARGS__) : fn<NOISE_INJECTED, false>(__VA_ARGS__);       \
        default: return tab ? fn<NOISE_PHILOX_PACKED, true>(__VA_ARGS__) : fn<NOISE_PHILOX_PACKED, false>(__VA_ARGS__);        \
    }

cudaError_t launch_heun_cluster(int noise, bool tab, int np, int layout, dim3 grid, dim3 block, size_t smem,
                                cudaStream_t s, const RunParams& P) {
    MB_NOISE_TAB_DISPATCH(launch_hc, np, layout, grid, block, smem, s, P)
}

cudaError_t launch_imid_cluster(int noise, bool tab, int np, dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                const RunParams& P) {
    MB_NOISE_TAB_DISPATCH(launch_ic, np, grid, block, smem, s, P)
}

}  // namespace mb
