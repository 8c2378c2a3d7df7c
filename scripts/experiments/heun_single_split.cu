// heun_single_split.cu — K1s: explicit Heun, single particle, for ensembles too small to hide latency with warps.
//
// Below one warp per SM sub-partition (R < ~19k members; BASELINE config 1 has 1000) K1 is bound by the in-order
// instruction stream of ONE warp-step: ~89 instructions, of which the integrator's 40 DFMA form a 12-deep dependent
// chain and the rest is the noise generator (Philox, Box-Muller on the SFU, float->double) — 191 cycles per step,
// 9.7 ms for 1e5 steps whatever the ensemble size.  Nothing can be gained across members there, so the step itself is
// split across warps: in a CTA of 96 threads warp 0 integrates 32 members and warps 1 and 2 generate their Wiener
// increments one batch ahead (alternating batches of 8 steps) into a shared-memory ring; producer and consumer meet
// on named barriers (bar.arrive / bar.sync pairs, one "full" and one "empty" barrier per ring slot).  The three warps
// of a CTA sit on three different sub-partitions, so the generator's IMAD.WIDE / F2F no longer queue in front of the
// integrator's DFMAs.  Same Philox counters, same fp32 Box-Muller, same fused Heun arithmetic as K1: the output is
// bit-identical to heun_single_kernel's (tests/test_parity_gpu.py).
#include "common.cuh"
#include "launch.h"

namespace mb {

constexpr int SPLIT_BATCH = 8;     // steps per ring slot
constexpr int SPLIT_SLOTS = 4;     // ring depth (two per producer warp)
constexpr int SPLIT_PRODUCERS = 2;

__device__ __forceinline__ void bar_sync(const int id, const int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(const int id, const int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <bool FIELD_TAB, bool AXIS_Z, bool RENORM>
__global__ void __launch_bounds__(32 * (1 + SPLIT_PRODUCERS)) heun_single_split_kernel(const __grid_constant__ RunParams P) {
    __shared__ double ring[SPLIT_SLOTS][SPLIT_BATCH][3][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t r_raw = (uint64_t)blockIdx.x * 32 + lane;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    const uint64_t n_steps = P.j1 - P.j0;
    const uint64_t n_batches = (n_steps + SPLIT_BATCH - 1) / SPLIT_BATCH;
    // barrier ids: full[s] = 1 + s, empty[s] = 1 + SPLIT_SLOTS + s; 64 = the consumer warp + the producer warp of the slot
    if (warp > 0) {
        // ---- producer: batches b = warp - 1, warp - 1 + PRODUCERS, ... into ring slot b % SLOTS ----
        const double c = P.sig[r * P.sig_rs] * P.sqrt_dt;
        const float bm_scale = scale_to_bm(c);
        const uint64_t seed = (uint64_t)P.seeds[r];
        const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
        const uint32_t member = member_id(P, r);
        uint64_t use = 0;   // how many times this producer has filled each of its slots so far (per slot round)
        for (uint64_t b = warp - 1; b < n_batches; b += SPLIT_PRODUCERS, ++use) {
            const int slot = (int)(b % SPLIT_SLOTS);
            if (b >= SPLIT_SLOTS) bar_sync(1 + SPLIT_SLOTS + slot, 64);      // the consumer has drained this slot
            const uint64_t jb = P.j0 + b * SPLIT_BATCH;
            const int n = (int)min((uint64_t)SPLIT_BATCH, P.j1 - jb);
            for (int s = 0; s < n;) {
                const uint64_t j = jb + s;
                float g[6];
                philox_gauss6_f32<0>(key0, key1, (uint32_t)(j >> 1), 0u, member, bm_scale, g, P.philox_m0, P.philox_m1);
                if ((j & 1) == 0) {
                    ring[slot][s][0][lane] = widen_f32(g[0]); ring[slot][s][1][lane] = widen_f32(g[1]); ring[slot][s][2][lane] = widen_f32(g[2]);
                    if (s + 1 < n) {
                        ring[slot][s + 1][0][lane] = widen_f32(g[3]); ring[slot][s + 1][1][lane] = widen_f32(g[4]);
                        ring[slot][s + 1][2][lane] = widen_f32(g[5]);
                    }
                    s += 2;
                } else {
                    ring[slot][s][0][lane] = widen_f32(g[3]); ring[slot][s][1][lane] = widen_f32(g[4]); ring[slot][s][2][lane] = widen_f32(g[5]);
                    s += 1;
                }
            }
            __threadfence_block();
            bar_arrive(1 + slot, 64);                                          // slot is full
        }
        return;
    }
    // ---- consumer: the integrator ----
    V3 m{P.state[r], P.state[P.R + r], P.state[2 * P.R + r]};
    V3 e{0.0, 0.0, 1.0};
    if (!AXIS_Z)
        e = V3{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
    const double alpha = P.alpha, dt = P.dt;
    const double kdt = P.k_red[0] * dt;
    const V3 edt{e.x * kdt, e.y * kdt, e.z * kdt};
    const double2* tab = reinterpret_cast<const double2*>(P.field_tab);
    uint32_t k = P.k0;
    auto record = [&]() {   // sample k holds the current state
        if (P.traj != nullptr && live) {
            double* t = P.traj + (uint64_t)k * 3 * P.R + r;
            t[0] = m.x; t[P.R] = m.y; t[2 * P.R] = m.z;
        }
        if (P.partial != nullptr) {
            const double z = live ? m.z : 0.0;
            const double v0 = warp_sum(live ? m.x : 0.0), v1 = warp_sum(live ? m.y : 0.0), v2 = warp_sum(z), v3 = warp_sum(z * z);
            if (lane == 0) {
                double* o = P.partial + ((uint64_t)(k - P.k0) * gridDim.x + blockIdx.x) * 4;
                o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
            }
        }
        ++k;
    };
    uint64_t next = k < P.k1 ? P.target[k] : ~0ull;
    for (uint64_t b = 0; b < n_batches; ++b) {
        const int slot = (int)(b % SPLIT_SLOTS);
        bar_sync(1 + slot, 64);                                                // wait until the slot is full
        const uint64_t jb = P.j0 + b * SPLIT_BATCH;
        const int n = (int)min((uint64_t)SPLIT_BATCH, P.j1 - jb);
#pragma unroll 2
        for (int s = 0; s < n; ++s) {
            const uint64_t j = jb + s;
            while (j == next) {
                record();
                next = k < P.k1 ? P.target[k] : ~0ull;
            }
            const V3 cw{ring[slot][s][0][lane], ring[slot][s][1][lane], ring[slot][s][2][lane]};
            double hz0 = P.h_const, hz1 = P.h_const;
            if (FIELD_TAB) {
                const double2 h = __ldg(tab + (j - P.j0));
                hz0 = h.x; hz1 = h.y;
            }
            m = heun_single_step<AXIS_Z>(m, e, edt, alpha, dt, cw, hz0, hz1);
            if (RENORM) renormalise(m);
        }
        if (b + SPLIT_SLOTS < n_batches) bar_arrive(1 + SPLIT_SLOTS + slot, 64);   // slot drained (only if someone waits for it)
    }
    while (k < P.k1 && P.target[k] == P.j1) record();
    if (live) {
        P.state[r] = m.x; P.state[P.R + r] = m.y; P.state[2 * P.R + r] = m.z;
    }
}

cudaError_t launch_heun_single_split(bool tab, bool axis_z, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(32 * (1 + SPLIT_PRODUCERS));
    const bool renorm = P.renorm != 0;
#define MB_HSS(T, A)                                                            \
    if (renorm) heun_single_split_kernel<T, A, true><<<g, b, 0, s>>>(P);        \
    else heun_single_split_kernel<T, A, false><<<g, b, 0, s>>>(P)
    if (tab) { if (axis_z) { MB_HSS(true, true); } else { MB_HSS(true, false); } }
    else { if (axis_z) { MB_HSS(false, true); } else { MB_HSS(false, false); } }
#undef MB_HSS
    return cudaGetLastError();
}

}  // namespace mb
