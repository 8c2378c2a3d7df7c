// small.cuh — interacting clusters of a few particles: ONE THREAD PER CLUSTER
// (Heun: 2..7 particles, small_heun.cu; implicit midpoint: 2..4, small_imid.cu).
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
#pragma once
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


#define MB_NOISE_TAB_DISPATCH(fn, ...)                                                              \
    switch (noise) {                                                                                \
        case NOISE_PHILOX_F32: return tab ? fn<NOISE_PHILOX_F32, true>(__VA_ARGS__) : fn<NOISE_PHILOX_F32, false>(__VA_ARGS__); \
        case NOISE_PHILOX_F64: return tab ? fn<NOISE_PHILOX_F64, true>(__VA_ARGS__) : fn<NOISE_PHILOX_F64, false>(__VA_ARGS__); \
        case NOISE_INJECTED: return tab ? fn<NOISE_INJECTED, true>(__VA_ARGS__) : fn<NOISE_INJECTED, false>(__VA_ARGS__);       \
        default: return tab ? fn<NOISE_PHILOX_PACKED, true>(__VA_ARGS__) : fn<NOISE_PHILOX_PACKED, false>(__VA_ARGS__);        \
    }

}  // namespace mb
