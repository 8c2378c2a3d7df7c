// heun_single_balanced.cu — K1b: heun_single_kernel as a persistent kernel over (time segment, member block) tasks.
//
// K1 is bound by the issue time of a warp-step, so the time of a launch is (warps per SM sub-partition,
// rounded UP) x steps x 145 cycles (151 with the round-1 instruction stream these measurements were taken with): a shard of 125,000 members (1M members over 8 GPUs) is 6.6 warps per sub-partition
// and pays for 7 (0.92 measured, profiles/r02_probe_k1_variants.log); 250,000 members pay 14 for 13.2.  Members are
// independent but a member's steps are sequential, so the only way to hand a sub-partition a FRACTION of a warp's work
// is to cut the time axis: the launch's step range is cut into segments, the ensemble into blocks of 128 members
// ("virtual CTAs"), and a grid of exactly the resident CTAs (6 per SM) pulls (segment, block) tasks from a counter in
// segment-major order.  A block's state is parked in HBM between its segments (24 B per member per segment — the same
// hand-over the host-side chunking uses) and a per-block flag orders a segment after its predecessor (release / acquire;
// the predecessor was handed out earlier to a CTA that is running, so the wait cannot deadlock).
// Measured (profiles/r02_probe_k1_balance*.log, 100,000 steps in the bench's sine field) with 6 CTAs per SM: 250,000 members
// 0.925 -> 0.99 of the 1M rate, 500,000 0.97 -> 1.01, 1M +1.8 %; 125,000 members 0.78-0.88 -> 0.92.  What kept the last
// from 0.99 was traced (profiles/r02_probe_k1_trace_fixed_segments.log): the sm_100a warp scheduler serves the oldest warps first and a
// single warp already takes 45 % of the FP64 pipe, so of the 6 CTAs resident on an SM the youngest two crawl — task
// durations per physical CTA differ by 9.5-19x on the SAME SM — a block that lands on one is held for milliseconds, and
// with only 1.1 blocks per CTA the launch ends in a ~10 ms tail of sequential stragglers (slot utilisation 0.85; 0.994
// at 1M members).  The remedy that works is to leave the starved CTAs out — FOUR CTAs per SM, see below: 125,000 members
// 0.98.  Two others were tried and did not pay: a FIFO of ready blocks instead of the in-order hand-out
// (same times), and slices sized per CTA for equal duration (profiles/r02_probe_k1_trace_adaptive_slices.log: utilisation
// 0.89, but 60 % more tasks and their overhead: 57.3 ms against 54.9, and 3 % slower at 1M).
// Same Philox counters and arithmetic as K1: bit-identical results (tests/test_parity_gpu.py).
#include "common.cuh"
#include "launch.h"

namespace mb {

// The register allocation is the plain kernel's (6 CTAs per SM with the easy axis along z: 76 registers; 5 with a general
// axis: 86 — forcing the latter into 80 costs 4 %), but only FOUR CTAs per SM are launched
// (profiles/r02_probe_k1_bal_ctas.log): the FP64 pipe is saturated by the four oldest warps of a sub-partition, the fifth
// and sixth CTA of an SM only hold blocks back — 125,000 members x 1e5 steps: 54.1 ms with 6 CTAs per SM, 52.3 with 5,
// 51.1 with 4, 52.4 with 3 (1M members: 403.3 / 401.1 / 402.6 / 417.7).  Spending the freed registers on a prefetch of
// the next step pair's field-table entries was measured and is slower at full load (1M: 417 ms,
// profiles/r02_probe_k1_bal_ctas_prefetch.log).
#ifndef MB_K1B_MIN_BLOCKS
#define MB_K1B_MIN_BLOCKS (AXIS_Z ? 6 : 5)   // tuning knob (scripts/build_variant.sh): a free allocation (82 registers) is 1.9 % slower
#endif
template <bool FIELD_TAB, bool AXIS_Z, bool RENORM>
__global__ void __launch_bounds__(SINGLE_THREADS, MB_K1B_MIN_BLOCKS) heun_single_balanced_kernel(const __grid_constant__ RunParams P) {
    __shared__ double red[(SINGLE_THREADS / 32) * 4];
    __shared__ unsigned int s_task;
    const uint32_t n_vcta = P.bal_vctas, n_seg = P.bal_segments;
    const uint64_t n_tasks = (uint64_t)n_vcta * n_seg;
    const double alpha = P.alpha, hkdt = P.half_kdt0, hdt = P.half_dt;   // half units: llg_math.cuh, heun_single_step
    const double2* tab = reinterpret_cast<const double2*>(P.field_tab);

    for (;;) {
        if (threadIdx.x == 0) s_task = atomicAdd(P.bal_counter, 1u);
        __syncthreads();
        const uint64_t task = s_task;
        if (task >= n_tasks) break;
        const uint32_t seg = (uint32_t)(task / n_vcta), vcta = (uint32_t)(task % n_vcta);
        // steps [ja, jb) of this launch's [j0, j1); the samples this segment records are those whose state index lies in
        // [ja, jb), plus — last segment — those at j1 itself
        const uint64_t span = P.j1 - P.j0;
        const uint64_t ja = P.j0 + span * seg / n_seg, jb = P.j0 + span * (seg + 1) / n_seg;
        const bool last = seg + 1 == n_seg;
        if (threadIdx.x == 0 && seg > 0) {     // the block's previous segment has parked its state
            unsigned int done, spins = 0;
            do {
                asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(done) : "l"(P.bal_flags + vcta));
                if (done < seg) {
                    __nanosleep(200);
                    // the predecessor was handed out earlier to a running CTA, so this wait is bounded by one segment's run
                    // time (milliseconds); a minute of waiting is a bug: fail the launch instead of hanging the device
                    if (++spins > (1u << 28)) __trap();
                }
            } while (done < seg);
        }
        __syncthreads();

        const uint64_t r_raw = (uint64_t)vcta * SINGLE_THREADS + threadIdx.x;
        const bool live = r_raw < P.R;
        const uint64_t r = live ? r_raw : P.R - 1;
        V3 m{__ldcg(P.state + r), __ldcg(P.state + P.R + r), __ldcg(P.state + 2 * P.R + r)};   // written by another SM: bypass L1
        V3 e{0.0, 0.0, 1.0};
        if (!AXIS_Z)
            e = V3{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
        const V3 eh{e.x * hkdt, e.y * hkdt, e.z * hkdt};
        const double ch = 0.5 * (P.sig[r * P.sig_rs] * P.sqrt_dt);
        const float bm_scale = scale_to_bm(ch);
        const uint64_t seed = (uint64_t)P.seeds[r];
        const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
        const uint32_t member = member_id(P, r);
        auto advance = [&](const V3& cw, const double2* tp) {
            double hz0 = P.h_const, hz1 = P.h_const;
            if (FIELD_TAB) {
                const double2 h = __ldg(tp);
                hz0 = h.x; hz1 = h.y;
            }
            m = heun_single_step<AXIS_Z>(m, e, eh, alpha, hdt, cw, hz0, hz1);
            if (RENORM) renormalise(m);
        };

        float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint32_t gblk = 0;
        bool have = false;
        auto need = [&](const uint32_t blk) {
            if (!have || gblk != blk) {
                philox_gauss6_f32<0>(key0, key1, blk, 0u, member, bm_scale, g, P.philox_m0, P.philox_m1, P.bm_mask_r, P.bm_mask_a);
                gblk = blk;
                have = true;
            }
        };
        uint64_t j = ja;
        // first sample of the launch whose state index is >= ja (P.bal_seg_k[seg], found on the host)
        for (uint32_t k = P.bal_seg_k[seg];; ++k) {
            const bool sample = k < P.k1 && (P.target[k] < jb || (last && P.target[k] == jb));
            const uint64_t tgt = sample ? P.target[k] : jb;
            if ((j & 1) && j < tgt) {
                need((uint32_t)(j >> 1));
                advance(V3{widen_f32(g[3]), widen_f32(g[4]), widen_f32(g[5])}, tab + (j - P.j0));
                ++j;
            }
            const uint32_t pairs = (uint32_t)((tgt - j) >> 1);
            uint32_t blk = (uint32_t)(j >> 1);
            const double2* tp = tab + (j - P.j0);
            if (pairs != 0) need(blk);
            uint32_t i = pairs;
            // two step pairs per trip: the generator alternates between the two blocks of increments, no copies
            for (; i >= 2; i -= 2, tp += 4) {
                float gn[6];
                philox_gauss6_f32<0>(key0, key1, ++blk, 0u, member, bm_scale, gn, P.philox_m0, P.philox_m1, P.bm_mask_r, P.bm_mask_a);
                advance(V3{widen_f32(g[0]), widen_f32(g[1]), widen_f32(g[2])}, tp);
                advance(V3{widen_f32(g[3]), widen_f32(g[4]), widen_f32(g[5])}, tp + 1);
                philox_gauss6_f32<0>(key0, key1, ++blk, 0u, member, bm_scale, g, P.philox_m0, P.philox_m1, P.bm_mask_r, P.bm_mask_a);
                advance(V3{widen_f32(gn[0]), widen_f32(gn[1]), widen_f32(gn[2])}, tp + 2);
                advance(V3{widen_f32(gn[3]), widen_f32(gn[4]), widen_f32(gn[5])}, tp + 3);
                gblk = blk;
            }
            for (; i != 0; --i, tp += 2) {
                float gn[6];
                philox_gauss6_f32<0>(key0, key1, ++blk, 0u, member, bm_scale, gn, P.philox_m0, P.philox_m1, P.bm_mask_r, P.bm_mask_a);
                advance(V3{widen_f32(g[0]), widen_f32(g[1]), widen_f32(g[2])}, tp);
                advance(V3{widen_f32(g[3]), widen_f32(g[4]), widen_f32(g[5])}, tp + 1);
#pragma unroll
                for (int q = 0; q < 6; ++q) g[q] = gn[q];
                gblk = blk;
            }
            j += 2ull * pairs;
            if (j < tgt) {
                need((uint32_t)(j >> 1));
                advance(V3{widen_f32(g[0]), widen_f32(g[1]), widen_f32(g[2])}, tab + (j - P.j0));
                ++j;
            }
            if (!sample) break;
            if (P.traj != nullptr && live) {
                double* t = P.traj + (uint64_t)k * 3 * P.R + r;
                t[0] = m.x; t[P.R] = m.y; t[2 * P.R] = m.z;
            }
            if (P.partial != nullptr) {
                const double z = live ? m.z : 0.0;
                cta_partial_sums<SINGLE_THREADS / 32>(live ? m.x : 0.0, live ? m.y : 0.0, z, z * z, red,
                                                      P.partial + ((uint64_t)(k - P.k0) * n_vcta + vcta) * 4);
            }
        }
        if (live) {
            __stcg(P.state + r, m.x); __stcg(P.state + P.R + r, m.y); __stcg(P.state + 2 * P.R + r, m.z);
        }
        __syncthreads();                       // every thread's state is stored before the flag is raised
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(P.bal_flags + vcta), "r"(seg + 1) : "memory");
        }
    }
}

cudaError_t launch_heun_single_balanced(bool tab, bool axis_z, unsigned phys_grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(phys_grid), b(SINGLE_THREADS);
    const bool renorm = P.renorm != 0;
#define MB_HSB(T, A)                                                              \
    if (renorm) heun_single_balanced_kernel<T, A, true><<<g, b, 0, s>>>(P);       \
    else heun_single_balanced_kernel<T, A, false><<<g, b, 0, s>>>(P)
    if (tab) { if (axis_z) { MB_HSB(true, true); } else { MB_HSB(true, false); } }
    else { if (axis_z) { MB_HSB(false, true); } else { MB_HSB(false, false); } }
#undef MB_HSB
    return cudaGetLastError();
}

int heun_single_balanced_resident_ctas(bool tab, bool axis_z, bool renorm) {
    int n = 0;
    auto q = [&](auto kernel) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, SINGLE_THREADS, 0); };
    if (tab) {
        if (axis_z) { if (renorm) q(heun_single_balanced_kernel<true, true, true>); else q(heun_single_balanced_kernel<true, true, false>); }
        else { if (renorm) q(heun_single_balanced_kernel<true, false, true>); else q(heun_single_balanced_kernel<true, false, false>); }
    } else {
        if (axis_z) { if (renorm) q(heun_single_balanced_kernel<false, true, true>); else q(heun_single_balanced_kernel<false, true, false>); }
        else { if (renorm) q(heun_single_balanced_kernel<false, false, true>); else q(heun_single_balanced_kernel<false, false, false>); }
    }
    return n;
}

}  // namespace mb
