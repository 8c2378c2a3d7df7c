// kernels.cuh — sm_100a kernels of the ensemble sLLG integrator.
//
// Data layout in HBM (R = members on this device, N = particles per cluster, n = 3N):
//   state     [n][R]        fp64, member index fastest -> every load/store is coalesced
//   axis      [n][R] or [n] anisotropy axes (per member or shared)
//   traj      [S][n][R]     sampled trajectory points (optional)
//   partial   [Sc][grid][4] per-CTA partial ensemble sums of one chunk of samples
//   sums      [S][4]        ensemble sums {Mx, My, Mz, Mz^2} (cluster-summed, reduced units)
//   field_tab [steps][2]    applied field at the two evaluation times of every step
//   dW        [steps][n][R] injected unit-variance increments (parity mode only)
//
// K1 heun_single      one thread per member, N = 1, state in fp64 registers
// K3 imid_single      same mapping, implicit midpoint with the reference's quasi-Newton
// K2 heun_cluster     one warp-wide CTA row per particle slot: lanes = 32 members,
//                     threadIdx.y = particle slot; moments staged in shared memory for
//                     the all-pairs dipolar sum
// K4 imid_cluster     same mapping; block-diagonal quasi-Newton, CTA-wide convergence
// K5 ensemble sums    fused into K1-K4 (warp shuffle -> smem -> per-CTA partial) +
//                     reduce_partials (fixed-order, deterministic)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "llg_math.cuh"
#include "rng.cuh"

namespace mb {

enum { NOISE_PHILOX_F32 = 0, NOISE_PHILOX_F64 = 1, NOISE_INJECTED = 2, NOISE_PHILOX_PACKED = 3 };

struct RunParams {
    uint64_t R;       // members on this device
    uint32_t N;       // particles per cluster
    int renorm;       // divide each moment by its 2-norm after every step
    int interactions; // all-pairs dipolar field
    double alpha, dt, sqrt_dt;
    double eps, clampA;  // implicit: tolerance, Ah = sqrt(2*1000*|ln dt|)
    double h_const;      // applied field when no table is used (reduced units)
    const double* k_red; // [N]
    const double* sig;   // [N] thermal field strength sigma_i
    const double* dip;   // [N][N][4] {rx, ry, rz, c_dip * v_j / cube_ij}; diagonal zero
    const double* axis;  // see layout
    uint64_t axis_cs, axis_rs;  // component stride, member stride
    const int64_t* seeds;       // [R]
    uint64_t stream_offset;
    double* state;              // [n][R]
    const uint64_t* target;     // [S] state index stored by sample k
    uint64_t j0, j1;            // advance the state from index j0 to j1
    uint32_t k0, k1;            // samples recorded by this launch
    const double* field_tab;    // [(j1-j0)][2] or nullptr
    const double* dW;           // injected noise or nullptr
    uint64_t dW_j0;             // step index of dW row 0
    double* traj;               // nullptr or [S][n][R]
    double* partial;            // nullptr or [(k1-k0)][gridDim.x][4]
    unsigned long long* newton; // [3] total / max / failures
};

// ---------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// unit-variance draws (implicit kernels clamp them before scaling, lib/integrators.cpp:598-602)
template <int NOISE>
__device__ __forceinline__ V3 draw_noise(const RunParams& P, uint32_t k0, uint32_t k1, uint64_t j, uint32_t particle,
                                         uint32_t member, uint64_t r) {
    if (NOISE == NOISE_INJECTED) {
        const uint64_t n = 3ull * P.N;
        const double* row = P.dW + ((j - P.dW_j0) * n + 3ull * particle) * P.R + r;
        return V3{row[0], row[P.R], row[2 * P.R]};
    } else {
        const Gauss3 g = philox_gauss3<NOISE>(k0, k1, j + 1, particle, member);
        return V3{g.x, g.y, g.z};
    }
}

// Heun kernels: the scaled increment c*w with c = sigma*sqrt(dt).  In the fp32 Gaussian mode the
// scale is folded into the Box-Muller radius (neg2ln2_c2 = -2 ln2 c^2), so no fp64 multiply is
// spent on the noise; the injected and fp64 modes multiply in fp64.
template <int NOISE>
__device__ __forceinline__ V3 draw_scaled(const RunParams& P, uint32_t k0, uint32_t k1, uint64_t j, uint32_t particle,
                                          uint32_t member, uint64_t r, double c, float neg2ln2_c2) {
    if (NOISE == NOISE_PHILOX_F32) {
        float x, y, z;
        philox_gauss3_f32(k0, k1, j + 1, particle, member, neg2ln2_c2, x, y, z);
        return V3{widen_f32(x), widen_f32(y), widen_f32(z)};
    } else {
        const V3 w = draw_noise<NOISE>(P, k0, k1, j, particle, member, r);
        return V3{c * w.x, c * w.y, c * w.z};
    }
}

__device__ __forceinline__ float scale_to_bm(double c) { return (float)(-1.3862943611198906 * c * c); }

// CTA-level sum of 4 values per thread over a 1-D block of NW warps into partial[slot][4]
template <int NW>
__device__ __forceinline__ void cta_partial_sums(double v0, double v1, double v2, double v3, double* smem /*NW*4*/,
                                                 double* out4) {
    v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        smem[warp * 4 + 0] = v0; smem[warp * 4 + 1] = v1; smem[warp * 4 + 2] = v2; smem[warp * 4 + 3] = v3;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += smem[w * 4 + threadIdx.x];
        out4[threadIdx.x] = s;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------
// K1: explicit Heun, single particle (lib/integrators.cpp:372-405 over the LLG SDE of
// lib/llg.cpp:332-348 with the field of lib/simulation.cpp:271-290, N = 1)
// ---------------------------------------------------------------------------------
constexpr int SINGLE_THREADS = 128;

// One Heun step of a single macrospin in 40 fp64 instructions (47 with a general easy axis).
// With u = g + alpha (m x g) the LLG increment is f(m,g) = -m x u, so the predictor
// x~ = m + f(m,g) and the corrector m' = (m + x~)/2 + f(x~,g~)/2 are accumulated directly in the
// FMAs of the second cross product (no separate adds, and f1 is never materialised).
template <bool AXIS_Z>
__device__ __forceinline__ V3 heun_single_step(const V3& m, const V3& e, const V3& edt, const double alpha,
                                               const double dt, const V3& cw, const double hz0, const double hz1) {
    // stage 1: g = h(m,t) dt + sigma sqrt(dt) w
    V3 g;
    if (AXIS_Z) {  // easy axis = z: h = (k m_z + h_app) z, two fp64 ops instead of seven
        g = V3{cw.x, cw.y, fma(m.z, edt.z, fma(hz0, dt, cw.z))};
    } else {
        const double s = dot(m, e);
        g = V3{fma(s, edt.x, cw.x), fma(s, edt.y, cw.y), fma(s, edt.z, fma(hz0, dt, cw.z))};
    }
    V3 p = cross(m, g);
    V3 u{fma(alpha, p.x, g.x), fma(alpha, p.y, g.y), fma(alpha, p.z, g.z)};
    const V3 mt{fma(-m.y, u.z, fma(m.z, u.y, m.x)), fma(-m.z, u.x, fma(m.x, u.z, m.y)),
                fma(-m.x, u.y, fma(m.y, u.x, m.z))};
    // stage 2 at (x~, t+dt), same Wiener increment
    if (AXIS_Z) {
        g = V3{cw.x, cw.y, fma(mt.z, edt.z, fma(hz1, dt, cw.z))};
    } else {
        const double s = dot(mt, e);
        g = V3{fma(s, edt.x, cw.x), fma(s, edt.y, cw.y), fma(s, edt.z, fma(hz1, dt, cw.z))};
    }
    p = cross(mt, g);
    u = V3{fma(alpha, p.x, g.x), fma(alpha, p.y, g.y), fma(alpha, p.z, g.z)};
    const V3 hm{0.5 * mt.x, 0.5 * mt.y, 0.5 * mt.z};
    const V3 h{fma(0.5, m.x, hm.x), fma(0.5, m.y, hm.y), fma(0.5, m.z, hm.z)};
    return V3{fma(-hm.y, u.z, fma(hm.z, u.y, h.x)), fma(-hm.z, u.x, fma(hm.x, u.z, h.y)),
              fma(-hm.x, u.y, fma(hm.y, u.x, h.z))};
}

__device__ __forceinline__ void renormalise(V3& m) {
    const double inv = 1.0 / sqrt(dot(m, m));
    m.x *= inv; m.y *= inv; m.z *= inv;
}

template <int NOISE, bool FIELD_TAB, bool AXIS_Z>
__global__ void __launch_bounds__(SINGLE_THREADS) heun_single_kernel(const __grid_constant__ RunParams P) {
    __shared__ double red[(SINGLE_THREADS / 32) * 4];
    const uint64_t r_raw = (uint64_t)blockIdx.x * SINGLE_THREADS + threadIdx.x;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;

    V3 m{P.state[r], P.state[P.R + r], P.state[2 * P.R + r]};
    V3 e{0.0, 0.0, 1.0};
    if (!AXIS_Z)
        e = V3{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
    const double alpha = P.alpha, dt = P.dt;
    const double kdt = P.k_red[0] * dt;
    const V3 edt{e.x * kdt, e.y * kdt, e.z * kdt};
    const double c = P.sig[0] * P.sqrt_dt;
    const float bm_scale = scale_to_bm(c);
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = (uint32_t)(r + P.stream_offset);
    const bool renorm = P.renorm != 0;

    // one Heun step from the scaled increment cw; j is the 0-based step index
    auto advance = [&](const V3& cw, const uint64_t jj) {
        double hz0 = P.h_const, hz1 = P.h_const;
        if (FIELD_TAB) {
            const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (jj - P.j0));
            hz0 = h.x; hz1 = h.y;
        }
        m = heun_single_step<AXIS_Z>(m, e, edt, alpha, dt, cw, hz0, hz1);
        if (renorm) renormalise(m);
    };

    uint64_t j = P.j0;
    float carry[3] = {0.f, 0.f, 0.f};   // packed mode: increments of the odd step of the current Philox block
    if (NOISE == NOISE_PHILOX_PACKED && (j & 1)) {
        float g[6];
        philox_gauss6_f32(key0, key1, j >> 1, 0u, member, bm_scale, g);
        carry[0] = g[3]; carry[1] = g[4]; carry[2] = g[5];
    }
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        const uint64_t tgt = (k < P.k1) ? P.target[k] : P.j1;
        if (NOISE == NOISE_PHILOX_PACKED) {
            // invariant: when j is odd, `carry` holds the second half of block j >> 1
            if ((j & 1) && j < tgt) {
                advance(V3{widen_f32(carry[0]), widen_f32(carry[1]), widen_f32(carry[2])}, j);
                ++j;
            }
            for (; j + 2 <= tgt; j += 2) {
                float g[6];
                philox_gauss6_f32(key0, key1, j >> 1, 0u, member, bm_scale, g);
                advance(V3{widen_f32(g[0]), widen_f32(g[1]), widen_f32(g[2])}, j);
                advance(V3{widen_f32(g[3]), widen_f32(g[4]), widen_f32(g[5])}, j + 1);
            }
            if (j < tgt) {
                float g[6];
                philox_gauss6_f32(key0, key1, j >> 1, 0u, member, bm_scale, g);
                advance(V3{widen_f32(g[0]), widen_f32(g[1]), widen_f32(g[2])}, j);
                carry[0] = g[3]; carry[1] = g[4]; carry[2] = g[5];
                ++j;
            }
        } else {
#pragma unroll 2
            for (; j < tgt; ++j) advance(draw_scaled<NOISE>(P, key0, key1, j, 0u, member, r, c, bm_scale), j);
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
    }
}

// ---------------------------------------------------------------------------------
// K3: implicit midpoint, single particle (lib/integrators.cpp:576-651 +
// lib/optimisation.cpp:81-149).  The quasi-Newton iteration is reproduced as the
// reference runs it: clamped increments, Euler-midpoint initial guess, Jacobian
// J = I - a'/2 - (B'.w)/2 with no dt on a', tolerance eps*||guess|| fixed before the
// loop, stop on ||delta||_2 <= tol or after 1000 iterations.
// ---------------------------------------------------------------------------------
struct NewtonCount {
    unsigned long long total, worst, fails;
};

__device__ __forceinline__ V3 imid_single_step(const V3& x0, const V3& e, const double kred, const double alpha,
                                               const double sr, const double dt, const double clampA,
                                               const double sqrt_dt, const double eps, const V3& w,
                                               const double hz_t, const double hz_mid, const double hj[9],
                                               NewtonCount& nc) {
    const V3 wm{fmax(-clampA, fmin(clampA, w.x)) * sqrt_dt, fmax(-clampA, fmin(clampA, w.y)) * sqrt_dt,
                fmax(-clampA, fmin(clampA, w.z)) * sqrt_dt};
    const V3 sw{sr * wm.x, sr * wm.y, sr * wm.z};
    // Euler half step as the initial guess of (x0 + x1)/2
    V3 X;
    {
        const double s = kred * dot(x0, e);
        const V3 g{fma(s * e.x, dt, sw.x), fma(s * e.y, dt, sw.y), fma(fma(s, e.z, hz_t), dt, sw.z)};
        const V3 f = llg_f(x0, g, alpha);
        X = V3{(f.x + x0.x) / 2, (f.y + x0.y) / 2, (f.z + x0.z) / 2};
    }
    const double tol = eps * sqrt(dot(X, X));
    double err = 2 * tol;
    int iter = 1000;
    unsigned long long done = 0;
    bool singular = false;
    while ((err > tol) && (iter-- > 0)) {
        const double s = kred * dot(X, e);
        const V3 h{s * e.x, s * e.y, fma(s, e.z, hz_mid)};
        const V3 g{fma(h.x, dt, sw.x), fma(h.y, dt, sw.y), fma(h.z, dt, sw.z)};
        const V3 f = llg_f(X, g, alpha);
        double b[3] = {-(X.x - x0.x - 0.5 * f.x), -(X.y - x0.y - 0.5 * f.y), -(X.z - x0.z - 0.5 * f.z)};
        double A[9], D[9], d[3];
        drift_jacobian(A, X, alpha, h, hj);
        diffusion_jacobian_dot(D, X, sr, alpha, wm);
#pragma unroll
        for (int i = 0; i < 9; ++i) A[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * A[i] - 0.5 * D[i];
        ++done;
        if (!solve3(A, b, d)) {
            // dgesv info > 0: the reference returns with x_root = -F (lib/optimisation.cpp:136-137)
            X = V3{b[0], b[1], b[2]};
            singular = true;
            break;
        }
        err = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        X.x += d[0]; X.y += d[1]; X.z += d[2];
    }
    nc.total += done;
    nc.worst = done > nc.worst ? done : nc.worst;
    nc.fails += (singular || iter == -1) ? 1ull : 0ull;
    return V3{2 * X.x - x0.x, 2 * X.y - x0.y, 2 * X.z - x0.z};
}

__device__ __forceinline__ void newton_flush(const RunParams& P, const NewtonCount& nc, bool live) {
    unsigned long long t = live ? nc.total : 0ull, w = live ? nc.worst : 0ull, f = live ? nc.fails : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t += __shfl_xor_sync(0xffffffffu, t, o);
        f += __shfl_xor_sync(0xffffffffu, f, o);
        const unsigned long long ow = __shfl_xor_sync(0xffffffffu, w, o);
        w = ow > w ? ow : w;
    }
    if ((threadIdx.x & 31) == 0 && P.newton != nullptr) {
        atomicAdd(P.newton + 0, t);
        atomicMax(P.newton + 1, w);
        atomicAdd(P.newton + 2, f);
    }
}

template <int NOISE, bool FIELD_TAB>
__global__ void __launch_bounds__(SINGLE_THREADS) imid_single_kernel(const __grid_constant__ RunParams P) {
    __shared__ double red[(SINGLE_THREADS / 32) * 4];
    const uint64_t r_raw = (uint64_t)blockIdx.x * SINGLE_THREADS + threadIdx.x;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;

    V3 m{P.state[r], P.state[P.R + r], P.state[2 * P.R + r]};
    const V3 e{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
    const double kred = P.k_red[0], sr = P.sig[0];
    double hj[9];
    {
        const double ev[3] = {e.x, e.y, e.z};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jx = 0; jx < 3; ++jx) hj[3 * i + jx] = kred * ev[i] * ev[jx];  // lib/field.cpp:159-174
    }
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = (uint32_t)(r + P.stream_offset);
    const bool renorm = P.renorm != 0;
    NewtonCount nc{0ull, 0ull, 0ull};

    uint64_t j = P.j0;
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        const uint64_t tgt = (k < P.k1) ? P.target[k] : P.j1;
        for (; j < tgt; ++j) {
            const V3 w = draw_noise<NOISE>(P, key0, key1, j, 0u, member, r);
            double hz0 = P.h_const, hz1 = P.h_const;
            if (FIELD_TAB) {
                const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (j - P.j0));
                hz0 = h.x; hz1 = h.y;
            }
            m = imid_single_step(m, e, kred, P.alpha, sr, P.dt, P.clampA, P.sqrt_dt, P.eps, w, hz0, hz1, hj, nc);
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
    }
    newton_flush(P, nc, live);
}

// ---------------------------------------------------------------------------------
// K2 / K4: interacting clusters.  blockDim = (32 members, PS particle slots); thread
// (lane, slot) owns particles slot, slot+PS, ... (NP of them, compile time).  The
// moments of all N particles of the CTA's 32 members live in shared memory as
// [N][3][32] so the j-loop of the dipolar sum reads conflict-free rows, while the
// static pair table {r_hat, c_ij} is the same address for the whole warp (one
// broadcast transaction).
// ---------------------------------------------------------------------------------
constexpr int CL_LANES = 32;

struct Own {  // per-thread, per-owned-particle constants
    V3 e;
    double kred, sr;
    float bm;  // -2 ln2 (sigma sqrt(dt))^2: Box-Muller scale of this particle's increments
    uint32_t p;
    bool valid;
};

// effective field of particle p given all moments in shared memory (lib/simulation.cpp:271-290)
__device__ __forceinline__ V3 cluster_field(const RunParams& P, const double* sm /*[N][3][32]*/, const Own& o,
                                            const V3& m, const double hz, const int lane) {
    double s = dot(m, o.e) * o.kred;
    V3 h{s * o.e.x, s * o.e.y, fma(s, o.e.z, hz)};
    if (P.interactions) {
        const double2* tab = reinterpret_cast<const double2*>(P.dip) + (uint64_t)o.p * P.N * 2;
        for (uint32_t jq = 0; jq < P.N; ++jq) {
            if (jq == o.p) continue;
            const double2 t0 = __ldg(tab + 2 * jq), t1 = __ldg(tab + 2 * jq + 1);
            const double* mj = sm + (uint64_t)jq * 3 * CL_LANES + lane;
            const double mx = mj[0], my = mj[CL_LANES], mz = mj[2 * CL_LANES];
            const double t3 = 3.0 * (mx * t0.x + my * t0.y + mz * t1.x);
            h.x = fma(t1.y, fma(t3, t0.x, -mx), h.x);
            h.y = fma(t1.y, fma(t3, t0.y, -my), h.y);
            h.z = fma(t1.y, fma(t3, t1.x, -mz), h.z);
        }
    }
    return h;
}

template <int NOISE, bool FIELD_TAB, int NP>
__global__ void __launch_bounds__(512) heun_cluster_kernel(const __grid_constant__ RunParams P) {
    extern __shared__ double smem[];
    const uint32_t N = P.N;
    double* sm_m = smem;                                  // [N][3][32] current moments
    double* sm_t = smem + (uint64_t)N * 3 * CL_LANES;     // [N][3][32] predictor moments
    double* sm_red = sm_t + (uint64_t)N * 3 * CL_LANES;   // [PS][3][32] sample reduction
    const int lane = threadIdx.x, slot = threadIdx.y, PS = blockDim.y;
    const uint64_t r_raw = (uint64_t)blockIdx.x * CL_LANES + lane;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    const double alpha = P.alpha, dt = P.dt;
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = (uint32_t)(r + P.stream_offset);
    const bool renorm = P.renorm != 0;

    Own own[NP];
    V3 m[NP];
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
        own[q].bm = scale_to_bm(own[q].sr * P.sqrt_dt);
        m[q] = V3{P.state[c0 * P.R + r], P.state[(c0 + 1) * P.R + r], P.state[(c0 + 2) * P.R + r]};
        if (own[q].valid) {
            double* d = sm_m + c0 * CL_LANES + lane;
            d[0] = m[q].x; d[CL_LANES] = m[q].y; d[2 * CL_LANES] = m[q].z;
        }
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
            V3 f1[NP], cw[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (!own[q].valid) continue;
                cw[q] = draw_scaled<NOISE>(P, key0, key1, j, own[q].p, member, r, own[q].sr * P.sqrt_dt, own[q].bm);
                const V3 h = cluster_field(P, sm_m, own[q], m[q], hz0, lane);
                const V3 g{fma(h.x, dt, cw[q].x), fma(h.y, dt, cw[q].y), fma(h.z, dt, cw[q].z)};
                f1[q] = llg_f(m[q], g, alpha);
                double* d = sm_t + 3ull * own[q].p * CL_LANES + lane;
                d[0] = m[q].x + f1[q].x; d[CL_LANES] = m[q].y + f1[q].y; d[2 * CL_LANES] = m[q].z + f1[q].z;
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (!own[q].valid) continue;
                const V3 mt{m[q].x + f1[q].x, m[q].y + f1[q].y, m[q].z + f1[q].z};
                const V3 h = cluster_field(P, sm_t, own[q], mt, hz1, lane);
                const V3 g{fma(h.x, dt, cw[q].x), fma(h.y, dt, cw[q].y), fma(h.z, dt, cw[q].z)};
                const V3 f2 = llg_f(mt, g, alpha);
                m[q] = V3{fma(0.5, f1[q].x + f2.x, m[q].x), fma(0.5, f1[q].y + f2.y, m[q].y),
                          fma(0.5, f1[q].z + f2.z, m[q].z)};
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
        if (!own[q].valid || !live) continue;
        const uint64_t c0 = 3ull * own[q].p;
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
    const uint64_t r_raw = (uint64_t)blockIdx.x * CL_LANES + lane;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    const double alpha = P.alpha, dt = P.dt, clampA = P.clampA, sqrt_dt = P.sqrt_dt;
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = (uint32_t)(r + P.stream_offset);
    const bool renorm = P.renorm != 0;
    NewtonCount nc{0ull, 0ull, 0ull};

    Own own[NP];
    V3 m[NP];
    double hj[NP][9];
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
        // The 9 doubles the reference reads as this particle's field Jacobian: flat offsets
        // 3p..3p+8 of the dense row-major (3N)^2 anisotropy Jacobian (lib/llg.cpp:387,
        // lib/field.cpp:159-174) — the true block only for N = 1.
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const uint32_t idx = 3u * own[q].p + i, row = idx / (3u * N), col = idx % (3u * N);
            double v = 0.0;
            if (row / 3 == col / 3 && row < 3u * N)
                v = P.k_red[row / 3] * P.axis[(uint64_t)row * P.axis_cs + r * P.axis_rs] *
                    P.axis[(uint64_t)col * P.axis_cs + r * P.axis_rs];
            hj[q][i] = v;
        }
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
                if (!own[q].valid) continue;
                const V3 w = draw_noise<NOISE>(P, key0, key1, j, own[q].p, member, r);
                wm[q] = V3{fmax(-clampA, fmin(clampA, w.x)) * sqrt_dt, fmax(-clampA, fmin(clampA, w.y)) * sqrt_dt,
                           fmax(-clampA, fmin(clampA, w.z)) * sqrt_dt};
                sw[q] = V3{own[q].sr * wm[q].x, own[q].sr * wm[q].y, own[q].sr * wm[q].z};
                const V3 h = cluster_field(P, sm_m, own[q], m[q], hz0, lane);
                const V3 g{fma(h.x, dt, sw[q].x), fma(h.y, dt, sw[q].y), fma(h.z, dt, sw[q].z)};
                const V3 f = llg_f(m[q], g, alpha);
                X[q] = V3{(f.x + m[q].x) / 2, (f.y + m[q].y) / 2, (f.z + m[q].z) / 2};
                double* d = sm_x + 3ull * own[q].p * CL_LANES + lane;
                d[0] = X[q].x; d[CL_LANES] = X[q].y; d[2 * CL_LANES] = X[q].z;
                part += dot(X[q], X[q]);
            }
            sm_red[slot * CL_LANES + lane] = part;
            __syncthreads();
            double nrm = 0.0;
            for (int s2 = 0; s2 < PS; ++s2) nrm += sm_red[s2 * CL_LANES + lane];
            const double tol = P.eps * sqrt(nrm);
            double err = 2 * tol;
            int iter = 1000;
            unsigned long long done = 0;
            bool singular = false;
            while (true) {
                bool active = (err > tol) && !singular;
                if (active) { active = iter > 0; --iter; }
                // barrier + vote: also orders the previous iteration's sm_x / sm_red traffic
                if (!__syncthreads_or(active ? 1 : 0)) break;
                V3 dl[NP];
                bool ok = true;
                part = 0.0;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (!own[q].valid) continue;
                    const V3 h = cluster_field(P, sm_x, own[q], X[q], hz1, lane);
                    const V3 g{fma(h.x, dt, sw[q].x), fma(h.y, dt, sw[q].y), fma(h.z, dt, sw[q].z)};
                    const V3 f = llg_f(X[q], g, alpha);
                    double b[3] = {-(X[q].x - m[q].x - 0.5 * f.x), -(X[q].y - m[q].y - 0.5 * f.y),
                                   -(X[q].z - m[q].z - 0.5 * f.z)};
                    double A[9], D[9], d[3];
                    drift_jacobian(A, X[q], alpha, h, hj[q]);
                    diffusion_jacobian_dot(D, X[q], own[q].sr, alpha, wm[q]);
#pragma unroll
                    for (int i = 0; i < 9; ++i) A[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * A[i] - 0.5 * D[i];
                    if (!solve3(A, b, d)) { ok = false; d[0] = d[1] = d[2] = 0.0; }
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
                        err = sqrt(e2);
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

// ---------------------------------------------------------------------------------
// small service kernels
// ---------------------------------------------------------------------------------
// applied field at the two evaluation times of each step (lib/field.cpp:37-54,
// lib/simulation.cpp:350-355): step index s = j+1, t = s*dt; Heun evaluates at t and t+dt
// (lib/integrators.cpp:386,394), implicit midpoint at t and t+dt/2 (:605,621).
__global__ void field_table_kernel(double* tab, uint64_t j0, uint64_t n_steps, double dt, double second_offset,
                                   int shape, double h0, double f_red) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_steps) return;
    const double t = (double)(unsigned int)(j0 + i + 1) * dt;  // `step` is unsigned int in the reference
    const double t2 = t + second_offset;
    double a, b;
    if (shape == 0) {
        a = h0 * sin(2 * 3.14159265358979323846 * f_red * t);
        b = h0 * sin(2 * 3.14159265358979323846 * f_red * t2);
    } else {
        a = h0 * (((int)(t * f_red * 2)) % 2 ? -1 : 1);
        b = h0 * (((int)(t2 * f_red * 2)) % 2 ? -1 : 1);
    }
    tab[2 * i] = a;
    tab[2 * i + 1] = b;
}

// sums[k0+kk][q] = sum over CTAs of partial[kk][cta][q], fixed order -> deterministic
__global__ void reduce_partials_kernel(const double* partial, double* sums, uint32_t k0, uint32_t n_cta) {
    __shared__ double sh[256 * 4];
    const uint32_t kk = blockIdx.x;
    const double* p = partial + (uint64_t)kk * n_cta * 4;
    double a[4] = {0, 0, 0, 0};
    for (uint32_t c = threadIdx.x; c < n_cta; c += blockDim.x) {
        const double4 v = *reinterpret_cast<const double4*>(p + (uint64_t)c * 4);
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) sh[threadIdx.x * 4 + q] = a[q];
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
#pragma unroll
            for (int q = 0; q < 4; ++q) sh[threadIdx.x * 4 + q] += sh[(threadIdx.x + s) * 4 + q];
        __syncthreads();
    }
    if (threadIdx.x < 4) sums[(uint64_t)(k0 + kk) * 4 + threadIdx.x] = sh[threadIdx.x];
}

// batched strided 2-D transpose: element (b,row,col) at in[b*in_bs + row*in_rs + col] goes to
// out[b*out_bs + col*out_rs + row], times `scale`.  Both sides are walked along their
// contiguous index, so loads and stores are coalesced.
__global__ void transpose_kernel(const double* in, double* out, uint64_t rows, uint64_t cols, uint64_t in_bs,
                                 uint64_t in_rs, uint64_t out_bs, uint64_t out_rs, double scale) {
    __shared__ double tile[32][33];
    const double* src = in + (uint64_t)blockIdx.z * in_bs;
    double* dst = out + (uint64_t)blockIdx.z * out_bs;
    const uint64_t c0 = (uint64_t)blockIdx.x * 32, r0 = (uint64_t)blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const uint64_t rr = r0 + i, cc = c0 + threadIdx.x;
        if (rr < rows && cc < cols) tile[i][threadIdx.x] = src[rr * in_rs + cc];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const uint64_t cc = c0 + i, rr = r0 + threadIdx.x;
        if (rr < rows && cc < cols) dst[cc * out_rs + rr] = tile[threadIdx.x][i] * scale;
    }
}

// out[q][r] = in[q] for q < n, r < R (shared initial state replicated over the members)
__global__ void broadcast_rows_kernel(const double* in, double* out, uint64_t n, uint64_t R) {
    const uint64_t total = n * R;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = in[i / R];
}

__global__ void scale_kernel(double* a, uint64_t n, double s) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] *= s;
}

// dependent-free DFMA chains: the measured FP64 roofline denominator
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(uint64_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void philox_words_kernel(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                    uint32_t* out) {
    philox4x32_10(c0, c1, c2, c3, k0, k1);
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <int GAUSS_MODE>
__global__ void gaussians_kernel(uint64_t seed, uint32_t member, uint32_t particle, uint64_t first_step,
                                 uint64_t n_steps, double* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_steps) return;
    const Gauss3 g = philox_gauss3<GAUSS_MODE>((uint32_t)seed, (uint32_t)(seed >> 32), first_step + i, particle, member);
    out[3 * i] = g.x; out[3 * i + 1] = g.y; out[3 * i + 2] = g.z;
}

}  // namespace mb
