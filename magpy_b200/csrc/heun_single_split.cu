// heun_single_split.cu — K1s: explicit Heun, single particle, for ensembles too small to give every SM sub-partition
// more than one warp (up to 64 members per SM: BASELINE config 1 has 1000 members).
//
// A warp that has its sub-partition to itself is bound by its own in-order instruction stream: one Heun step is 11
// dependent levels of FP64 instructions (8.1 cycles of latency each, scripts/micro/dfma_operands.cu) whose 2-3 members issue
// one after the other at 2-3 cycles each, and next to them the ~48 generator and loop instructions of the step take an
// issue slot each: 187 cycles per step measured for K1, whatever the ensemble size below one warp per sub-partition.  A
// small ensemble leaves most of an SM's four sub-partitions idle, so everything that does not depend on the state is
// moved there: in a CTA of 128 threads warp 0 (the consumer) integrates 32 members and warps 1-3 (the producers, on the
// other three sub-partitions, alternating batches) generate their Wiener increments — Philox, Box-Muller, float -> double
// — read the applied-field table and form the two z-field constants of the Heun stages, t0 = h_app(t) dt/2 + cwh.z and
// t1 = h_app(t + dt) dt/2 + cwh.z, a batch of SPLIT_B steps ahead, into a shared-memory ring.  The consumer's step is then
// 35 FP64 instructions and two LDS.128; ring slots are handed over on named barriers (one "full" and one "empty" barrier
// per slot, bar.arrive on one side, bar.sync on the other).  A hand-over costs ~100 cycles, hence batches of 32 steps
// (with 8 the kernel ran at 172 cycles per step, with 16 at 152, with 32 at 146).
// Measured (profiles/r02_probe_c1_split_v2.log, r02_probe_k1s_shapes.log): config 1 (1000 members x 1e5 steps) 9.52 -> 7.44 ms
// = 146 cycles per step — the 11 levels x (8.1 cycles + the serial issue of the level's members) of the step's own chain,
// i.e. what one member's step costs on one in-order warp; the bare 89-cycle chain is not reachable in order.  The same
// ensemble in a sine field 10.63 -> 7.76 ms (the table fetch leaves the integrator's stream too: one coalesced load per batch
// by the producer), 9472 members (two CTAs per SM, one producer each) 9.49 -> 7.80 ms, in a sine field 11.5 -> 8.0 ms; with
// renorm and / or a general easy axis the gain is 1-50 %.
// Same Philox counters, same fp32 Box-Muller, same fused arithmetic in the same order (llg_math.cuh: heun_single_core) as
// heun_single_kernel: per-member output is bit-identical (tests/test_parity_gpu.py); the ensemble sums are formed per
// 32 members instead of per 128, i.e. in another (equally fixed) order.
// A first version (earlier in round 2, scripts/experiments/README.md) moved only the generator, one Philox block at a time
// (a latency-bound producer that starved the consumer), and was slower than K1.
#include "common.cuh"
#include "launch.h"

namespace mb {

constexpr int SPLIT_B = 32;       // steps per ring slot (whole Philox blocks: slot boundaries sit at even step indices)
constexpr int SPLIT_SLOTS = 3;      // 3 x 32 steps x 1 KB = 96 KB of dynamic shared memory (one CTA per SM)
// NPROD generator warps per integrator warp: 3 (one on each of the SM's other three sub-partitions, alternating batches) when
// there is at most one CTA per SM, 1 (CTAs of 64 threads) when two CTAs share an SM — one generator keeps up with its
// integrator (~50 against 146 issue cycles per step), and two CTAs of two warps put every warp on its own sub-partition

__device__ __forceinline__ void bar_sync(const int id, const int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(const int id, const int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <bool FIELD_TAB, bool AXIS_Z, bool RENORM, int NPROD>
__global__ void __launch_bounds__(32 * (1 + NPROD)) heun_single_split_kernel(const __grid_constant__ RunParams P) {
    // ring[slot][step][0][lane] = {cwh.x, cwh.y}, ring[slot][step][1][lane] = {t0, t1}: every access is one conflict-free
    // 128-bit shared-memory transaction per lane
    extern __shared__ double2 ring_raw[];
    double2 (*ring)[SPLIT_B][2][32] = reinterpret_cast<double2 (*)[SPLIT_B][2][32]>(ring_raw);   // [SPLIT_SLOTS]
    __shared__ double2 htab[NPROD][32];    // per producer warp: the applied field of the batch being generated
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t r_raw = (uint64_t)blockIdx.x * 32 + lane;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    // batch b = the steps [base + b B, base + (b + 1) B) that lie in [j0, j1), base = j0 rounded down to an even index
    const uint64_t base = P.j0 & ~1ull;
    const uint64_t n_batches = (P.j1 - base + SPLIT_B - 1) / SPLIT_B;
    // named barriers: full[s] = 1 + s, empty[s] = 1 + SPLIT_SLOTS + s (barrier 0 is __syncthreads); 64 = the consumer warp
    // and the one producer warp that fills the slot this time round
    if (warp > 0) {
        // ---- producer `warp - 1`: batches warp - 1, warp - 1 + PRODUCERS, ... ----
        const double ch = 0.5 * (P.sig[r * P.sig_rs] * P.sqrt_dt);
        const double dth = P.half_dt;
        const float bm_scale = scale_to_bm(ch);
        const uint64_t seed = (uint64_t)P.seeds[r];
        const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
        const uint32_t member = member_id(P, r);
        const double2* tab = reinterpret_cast<const double2*>(P.field_tab);
        for (uint64_t b = warp - 1; b < n_batches; b += NPROD) {
            const int slot = (int)(b % SPLIT_SLOTS);
            if (b >= SPLIT_SLOTS) bar_sync(1 + SPLIT_SLOTS + slot, 64);      // the consumer has drained this slot
            const uint64_t jb = base + b * SPLIT_B;
            // the applied field of the batch's SPLIT_B steps: ONE coalesced load (lane i fetches step jb + i), parked in shared
            // memory and read back as broadcasts — a load per step at its point of use would put an L2 round trip into the
            // chain of every Philox block and, with one producer per consumer, starve the consumer
            if (FIELD_TAB) {
                static_assert(SPLIT_B == 32, "one table entry per lane");
                const uint64_t jl = jb + lane;
                htab[warp - 1][lane] = (jl >= P.j0 && jl < P.j1) ? __ldg(tab + (jl - P.j0)) : make_double2(0.0, 0.0);   // steps outside the launch are never consumed
                __syncwarp();
            }
            // the B / 2 Philox blocks of the batch, independent of each other: unrolled so that their multiply and
            // MUFU chains overlap (one block after the other is latency bound and would starve the consumer)
#pragma unroll 4
            for (int i = 0; i < SPLIT_B / 2; ++i) {
                float g[6];
                philox_gauss6_f32<0>(key0, key1, (uint32_t)((jb >> 1) + i), 0u, member, bm_scale, g, P.philox_m0, P.philox_m1,
                                     P.bm_mask_r, P.bm_mask_a);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double hz0 = P.h_const, hz1 = P.h_const;
                    if (FIELD_TAB) {
                        const double2 t = htab[warp - 1][2 * i + h];
                        hz0 = t.x; hz1 = t.y;
                    }
                    const double cz = widen_f32(g[3 * h + 2]);
                    ring[slot][2 * i + h][0][lane] = make_double2(widen_f32(g[3 * h]), widen_f32(g[3 * h + 1]));
                    ring[slot][2 * i + h][1][lane] = make_double2(fma(hz0, dth, cz), fma(hz1, dth, cz));
                }
            }
            if (FIELD_TAB) __syncwarp();      // every lane has read the table slab before the next batch overwrites it
            __threadfence_block();
            bar_arrive(1 + slot, 64);                                          // the slot is full
        }
        return;
    }
    // ---- consumer: the integrator ----
    V3 m{P.state[r], P.state[P.R + r], P.state[2 * P.R + r]};
    V3 e{0.0, 0.0, 1.0};
    if (!AXIS_Z)
        e = V3{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
    const double alpha = P.alpha, hkdt = P.half_kdt0;
    const V3 eh{e.x * hkdt, e.y * hkdt, e.z * hkdt};
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
    // state index at which sample k is recorded (zero-order hold); a launch records samples k0 .. k1 - 1
    uint64_t next = k < P.k1 ? P.target[k] : ~0ull;
    uint64_t j = P.j0;
    for (uint64_t b = 0; b < n_batches; ++b) {
        const int slot = (int)(b % SPLIT_SLOTS);
        bar_sync(1 + slot, 64);                                                // wait until the slot is full
        const uint64_t jb = base + b * SPLIT_B;
        const int n = (int)(min(jb + SPLIT_B, P.j1) - jb);                     // slot entries [s, n) hold steps of this launch
        const double2* sp = &ring[slot][0][0][lane];
        int s = (int)(j - jb);
        while (s < n) {
            while (j == next) {
                record();
                next = k < P.k1 ? P.target[k] : ~0ull;
            }
            const int run = (int)min((uint64_t)(n - s), next - j);      // steps before the next sample (or the end of the slot)
#pragma unroll 4
            for (int i = 0; i < run; ++i) {
                const double2 c = sp[(s + i) * 64], t = sp[(s + i) * 64 + 32];
                m = heun_single_core<AXIS_Z>(m, e, eh, alpha, c.x, c.y, t.x, t.y);
                if (RENORM) renormalise(m);
            }
            s += run;
            j += (uint64_t)run;
        }
        if (b + SPLIT_SLOTS < n_batches) bar_arrive(1 + SPLIT_SLOTS + slot, 64);   // slot drained (only if someone waits for it)
    }
    while (k < P.k1 && P.target[k] == P.j1) record();
    if (live) {
        P.state[r] = m.x; P.state[P.R + r] = m.y; P.state[2 * P.R + r] = m.z;
    }
}

// grid = ceil(R / 32) CTAs of 32 (1 + producers) threads; packed-noise production mode only
cudaError_t launch_heun_single_split(bool tab, bool axis_z, int producers, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(32 * (1 + producers));
    const bool renorm = P.renorm != 0;
    constexpr size_t smem = sizeof(double2) * SPLIT_SLOTS * SPLIT_B * 2 * 32;
    // the ring is beyond the 48 KB static limit: opt in before every launch (the attribute is per device, and a process may
    // drive several devices)
#define MB_HSS2(T, A, RN, NP)                                                                                         \
    do {                                                                                                              \
        const cudaError_t e = cudaFuncSetAttribute(heun_single_split_kernel<T, A, RN, NP>,                            \
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
        if (e != cudaSuccess) return e;                                                                               \
        heun_single_split_kernel<T, A, RN, NP><<<g, b, smem, s>>>(P);                                                 \
    } while (0)
#define MB_HSS(T, A, RN)                                             \
    do {                                                             \
        if (producers == 3) MB_HSS2(T, A, RN, 3);                    \
        else if (producers == 1) MB_HSS2(T, A, RN, 1);               \
        else return cudaErrorInvalidValue;                           \
    } while (0)
    if (tab) {
        if (axis_z) { if (renorm) MB_HSS(true, true, true); else MB_HSS(true, true, false); }
        else { if (renorm) MB_HSS(true, false, true); else MB_HSS(true, false, false); }
    } else {
        if (axis_z) { if (renorm) MB_HSS(false, true, true); else MB_HSS(false, true, false); }
        else { if (renorm) MB_HSS(false, false, true); else MB_HSS(false, false, false); }
    }
#undef MB_HSS
#undef MB_HSS2
    return cudaGetLastError();
}

}  // namespace mb
