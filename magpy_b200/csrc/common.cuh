// common.cuh — shared declarations of the sm_100a kernels of the ensemble sLLG integrator
// (kernels: heun_single.cu, imid_single.cu, small_heun.cu, small_imid.cu, cluster.cu, service.cu; launchers: launch.h).
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
//   dip       [N][N][4]     static pair table {sqrt(3) r_hat_ij, c_dip v_j / cube_ij}, zero diagonal
//
// K1 heun_single  (heun_single.cu)  one thread per member, N = 1, state in fp64 registers
// K3 imid_single  (imid_single.cu)  same mapping, implicit midpoint with the reference's quasi-Newton
// K2s/K4s *_small (small_*.cu)      Heun N = 2..7, implicit N = 2..4: one thread per cluster, moments in registers
// K2 heun_cluster (cluster.cu)      N = 8..128: CTA = 32 members (lanes) x particle slots, 2/4/8 own particles
//                                   per thread, moments (and the pair table, N <= 64) in shared memory
// K2m heun_cluster_mma (cluster_mma.cu) N = 8..128 with a dense enough 8-particle grouping: the dipolar field as
//                                   D[3N x 3N] . M[3N x members] on DMMA.8x8x4, D packed in shared memory
// K4 imid_cluster (cluster.cu)      N = 5..64: same mapping; block-diagonal quasi-Newton, CTA-wide convergence
// K5 ensemble sums                  fused into all of them (warp shuffle -> smem -> per-CTA partial) +
//                                   reduce_partials (fixed order, deterministic)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "llg_math.cuh"
#include "rng.cuh"

namespace mb {

enum { NOISE_PHILOX_F32 = 0, NOISE_PHILOX_F64 = 1, NOISE_INJECTED = 2, NOISE_PHILOX_PACKED = 3, NOISE_PHILOX_COARSE = 4 };

struct RunParams {
    uint64_t R;       // members on this device
    uint32_t N;       // particles per cluster
    int renorm;       // divide each moment by its 2-norm after every step
    int interactions; // all-pairs dipolar field
    double alpha, dt, sqrt_dt;
    double eps, clampA;  // implicit: tolerance, Ah = sqrt(2*1000*|ln dt|)
    int quirk_zero;      // implicit, N >= 2: the first easy axis (shared by all members) has no x and no y component, so the
                         // "field Jacobian" block the reference reads is all zeros (llg_math.cuh: quirk_u, newton_matrix)
    int newton_exact;    // implicit: 0 = the reference's quasi-Newton iteration (parity), 1 = Newton with the exact Jacobian
    double h_const;      // applied field when no table is used (reduced units)
    const double* k_red; // [N]
    double half_kdt0, half_dt;  // k_red[0] dt / 2 and dt / 2 from the host: kernel-parameter constants, so that the single-particle
                         // Heun kernels read them as uniform-register operands (llg_math.cuh: heun_single_step works in half units;
                         // a DFMA with three REGISTER operands holds the issue port for 3 cycles instead of 2)
    const double* sig;   // [N] thermal field strength sigma_i; N = 1 with per-member radii: [R] (sig_rs = 1)
    uint64_t sig_rs;     // member stride of `sig` (0 or 1)
    const double* dip;   // [N][N][4] {sqrt(3) r_hat_ij (3), c_dip * v_j / cube_ij}; diagonal zero
    const double* dmat;  // K2m: packed upper 24x24 blocks of the symmetric dipolar matrix (cluster_mma.cu) or nullptr
    const double* v_red; // [N] reduced volumes (K2m folds them into the moments)
    uint32_t G;          // K2m: particle groups of 8 (= ceil(N / 8))
    uint32_t mma_full, mma_tail;  // K2m: CTAs that take a full set of members; members per CTA of the partial last wave
    uint32_t mma_dglobal;            // K2m: the matrix is read from global memory (N = 65..128)
    uint32_t mma_mono;               // K2m: all reduced volumes are exactly 1 (no scaling of the moments)
    uint32_t cta_offset, cta_total;  // K2m: first CTA index of this launch, CTAs of both launches together
    const double* axis;  // see layout
    uint64_t axis_cs, axis_rs;  // component stride, member stride
    const int64_t* seeds;       // [R]
    uint64_t stream_offset;
    const uint32_t* member_idx; // [R] global member index of the Philox counter, or nullptr: stream_offset + r
    // per-member material parameters (N = 1, `MP` instantiations of the single-particle kernels): every thread carries
    // its own reduced time scale, hence its own dt, noise amplitude, field scale and zero-order-hold schedule
    const double* mp_alpha;     // [R] damping
    const double* mp_dt;        // [R] reduced time step dt * tau_r
    const double* mp_h0;        // [R] reduced field amplitude H0_r / H_k,r (the field table holds the unit waveform)
    const double* mp_Ts;        // [R] reduced sampling interval T_r / (S - 1)
    uint32_t* member_j;         // [R] state index reached by each member (carried between the launches of a run)
    uint64_t tab_j0;            // step index of field-table row 0 in the MP instantiations (j0 minus a margin)
    uint32_t philox_m0, philox_m1;  // the two Philox multipliers, passed at run time for the split multiply of rng.cuh
    uint32_t bm_mask_r, bm_mask_a;  // 0x007fffff / 0x007fffe0: the field masks of the packed Gaussian stream, passed at run time so that
                                    // (w & mask) | exponent stays ONE LOP3 (rng.cuh: philox_gauss6_f32)
    uint32_t coarsen_log2;      // NOISE_PHILOX_COARSE: step s sums the packed stream's fine steps s 2^L .. (s+1) 2^L - 1
    double* state;              // [n][R]
    double* state_t;            // [n][R] predictor moments of the capacity-free Heun kernel (cluster_big.cu) or nullptr
    double* state_u;            // [n][R] second midpoint-iterate buffer of the capacity-free implicit kernel or nullptr
    const uint64_t* target;     // [S] state index stored by sample k
    uint64_t j0, j1;            // advance the state from index j0 to j1
    uint32_t k0, k1;            // samples recorded by this launch
    const double* field_tab;    // [(j1-j0)][2] or nullptr
    const double* dW;           // injected noise or nullptr
    uint64_t dW_j0;             // step index of dW row 0
    double* traj;               // nullptr or [S][n][R]
    double* partial;            // nullptr or [(k1-k0)][gridDim.x][4]
    unsigned long long* newton; // [3] total / max / failures
    // K1b (heun_single_balanced.cu): persistent kernel over (time segment, block of 128 members) tasks
    uint32_t bal_vctas, bal_segments;   // member blocks ("virtual CTAs") and time segments of this launch
    unsigned int* bal_counter;          // task counter (zeroed before the launch)
    unsigned int* bal_flags;            // [bal_vctas] segments completed by each member block
    const uint32_t* bal_seg_k;          // [bal_segments] first sample of the launch with state index >= the segment's first step
};

// ---------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Philox counter word 1: the member's GLOBAL index in the ensemble — does not depend on how the ensemble is cut into
// parameter groups, shards or devices (explicit list), or stream_offset + local index for a contiguous slice
__device__ __forceinline__ uint32_t member_id(const RunParams& P, uint64_t r) {
    return P.member_idx ? __ldg(P.member_idx + r) : (uint32_t)(r + P.stream_offset);
}

// unit-variance draws (implicit kernels clamp them before scaling, lib/integrators.cpp:598-602)
template <int NOISE>
__device__ __forceinline__ V3 draw_noise(const RunParams& P, uint32_t k0, uint32_t k1, uint64_t j, uint32_t particle,
                                         uint32_t member, uint64_t r) {
    if (NOISE == NOISE_PHILOX_COARSE) {
        // L >= 1: the 2^L fine steps of coarse step j are the 2^(L-1) whole blocks j 2^(L-1) ... of the packed stream
        const uint32_t L = P.coarsen_log2;
        const uint64_t nb = 1ull << (L - 1), b0 = j << (L - 1);
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (uint64_t b = 0; b < nb; ++b) {
            float g[6];
            philox_gauss6_f32(k0, k1, b0 + b, particle, member, MB_NEG_2LN2, g);
            ax += (double)g[0] + (double)g[3];
            ay += (double)g[1] + (double)g[4];
            az += (double)g[2] + (double)g[5];
        }
        const double sc = exp2(-0.5 * (double)L);
        return V3{sc * ax, sc * ay, sc * az};
    } else if (NOISE == NOISE_INJECTED) {
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

constexpr int SINGLE_THREADS = 128;
constexpr int CL_LANES = 32;

// m / |m| (lib/simulation.cpp:379-387).  After a step from a unit vector |m|^2 = 1 + d with a tiny d (Heun: ~|f|^4 / 4,
// 6e-7 at the bench's step), so 1 / sqrt(1 + d) = 1 - d/2 + 3 d^2/8 - 5 d^3/16 to < 1.1e-15 for |d| < 2.5e-4: four
// dependent FMAs instead of MUFU.RSQ64H + Newton steps + the fix-up code of rsqrt() (K1 with renorm: 2.05e11 -> see
// profiles/r02_probe_renorm.log).  Anything further from 1 (first step of an unnormalised start, huge steps) takes
// rsqrt(), which is branch free and accurate to 1 ulp; the literal 1/sqrt() costs a DSQRT and a DDIV with their
// slow-path calls.
__device__ __forceinline__ void renormalise(V3& m) {
    const double d = dot(m, m) - 1.0;
    double inv = fma(d, fma(d, fma(d, -0.3125, 0.375), -0.5), 1.0);
    if (fabs(d) >= 2.5e-4) inv = rsqrt(d + 1.0);
    m.x *= inv; m.y *= inv; m.z *= inv;
}

// State index stored by sample k >= 1 of ONE member: the zero-order-hold schedule of lib/simulation.cpp:342-355 in the
// member's own reduced units — the smallest step count s with fl(s * dt) > fl(k * Ts), minus one (the state before the
// step that breaches the sampling time).  Same fp64 operations as build_schedule on the host (magpy_b200.cu); the
// products are formed with __dmul_rn so that no FMA contraction changes a comparison at an exact tie (Ts / dt an
// integer is the common case).
__device__ __forceinline__ uint64_t member_target(const uint32_t k, const double dt, const double Ts) {
    if (k == 0) return 0;
    const double lim = __dmul_rn((double)k, Ts);
    const double est = floor(__ddiv_rn(lim, dt));
    uint64_t s = est > 2.0 ? (uint64_t)est - 2 : 0;
    while (__dmul_rn((double)s, dt) <= lim) ++s;
    while (s > 0 && __dmul_rn((double)(s - 1), dt) > lim) --s;
    return s - 1;
}

struct NewtonCount {
    unsigned long long total, worst, fails;
};

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

}  // namespace mb
