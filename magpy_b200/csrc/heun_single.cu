#include "common.cuh"
#include "launch.h"

namespace mb {

// ---------------------------------------------------------------------------------
// K1: explicit Heun, single particle (lib/integrators.cpp:372-405 over the LLG SDE of
// lib/llg.cpp:332-348 with the field of lib/simulation.cpp:271-290, N = 1)
// ---------------------------------------------------------------------------------

#ifndef MB_PHILOX_SPLIT
#define MB_PHILOX_SPLIT 0   // leading Philox rounds whose wide multiplies are issued as IMAD.HI + IMAD (tuning knob, rng.cuh)
#endif
// MINB = resident CTAs per SM the register allocation must allow.  1: ptxas is free (76 registers, 6 CTAs = 24 warps
// per SM); 7: at most 72 registers (66 used, no spills).  The host picks 7 when the shard fits in one wave of 7 CTAs per
// SM but not in one of 6 (magpy_b200.cu: choose_k1_variant; measured in profiles/r02_probe_k1_variants.log).
// MP = per-member material parameters (anisotropy, damping, field amplitude next to radius / temperature): alpha, dt,
// the noise amplitude, the field scale and the sampling schedule are per-thread values (RunParams::mp_*).
// MINB = K1_LATENCY: the variant for ensembles too small to hide latency with warps — the applied-field table entries of
// the NEXT step pair are fetched while the current pair is integrated (8 more registers; free allocation).
template <int NOISE, bool FIELD_TAB, bool AXIS_Z, bool RENORM, int MINB, bool MP = false>
__global__ void __launch_bounds__(SINGLE_THREADS, MINB == 7 ? 7 : 1) heun_single_kernel(const __grid_constant__ RunParams P) {
    constexpr bool PREFETCH_TAB = MINB == K1_LATENCY && FIELD_TAB;
    __shared__ double red[(SINGLE_THREADS / 32) * 4];
    const uint64_t r_raw = (uint64_t)blockIdx.x * SINGLE_THREADS + threadIdx.x;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;

    V3 m{P.state[r], P.state[P.R + r], P.state[2 * P.R + r]};
    V3 e{0.0, 0.0, 1.0};
    if (!AXIS_Z)
        e = V3{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
    const double alpha = MP ? P.mp_alpha[r] : P.alpha, dt = MP ? P.mp_dt[r] : P.dt;
    // heun_single_step works in HALF units (llg_math.cuh): k dt e / 2, dt / 2, sigma sqrt(dt) w / 2.  The halves of the
    // ensemble-wide constants come from the host as kernel parameters (uniform-register operands)
    const double hkdt = MP ? 0.5 * (P.k_red[0] * dt) : P.half_kdt0;
    const V3 eh{e.x * hkdt, e.y * hkdt, e.z * hkdt};
    // per-member sigma when the radii differ between members; ch = sigma sqrt(dt) / 2
    const double ch = 0.5 * (P.sig[r * P.sig_rs] * (MP ? sqrt(dt) : P.sqrt_dt));
    const double mp_h0 = MP ? P.mp_h0[r] : 0.0, mp_Ts = MP ? P.mp_Ts[r] : 0.0;
    const double hdth = MP ? 0.5 * (mp_h0 * dt) : P.half_dt;   // multiplies the applied field (MP: unit waveform in the table)
    const float bm_scale = scale_to_bm(ch);
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = member_id(P, r);

    // one Heun step from the HALVED scaled increment cw; `tp` = this step's entry of the field table
    auto advance = [&](const V3& cw, const double2* tp) {
        double hz0 = MP ? 1.0 : P.h_const, hz1 = hz0;
        if (FIELD_TAB) {
            const double2 h = __ldg(tp);
            hz0 = h.x; hz1 = h.y;
        }
        m = heun_single_step<AXIS_Z>(m, e, eh, alpha, hdth, cw, hz0, hz1);
        if (RENORM) renormalise(m);
    };
    auto advance_h = [&](const V3& cw, const double2 h) {   // the same with the table entry already in registers
        m = heun_single_step<AXIS_Z>(m, e, eh, alpha, hdth, cw, h.x, h.y);
        if (RENORM) renormalise(m);
    };

    // All loop state of the inner loops is 32-bit (a launch covers at most 2^32 steps and the Philox
    // counter word is the low 32 bits of the step / step-pair index anyway) plus one table pointer.
    uint64_t j = MP ? (uint64_t)P.member_j[r] : P.j0;
    const uint64_t tab0 = MP ? P.tab_j0 : P.j0;
    const double2* tab = reinterpret_cast<const double2*>(P.field_tab);
    // Packed noise: `g` holds the six increments of Philox block `gblk` (steps 2 gblk and 2 gblk + 1) while
    // `have` is set.  The pair loop is software pipelined — the block of the NEXT step pair is generated inside
    // the loop body that integrates the CURRENT pair, so that the generator's IMAD.WIDE / MUFU / F2F instructions
    // are interleaved with the integrator's DFMA chain in program order — and the prefetched block is carried
    // across sample boundaries, so nothing is generated twice or thrown away (except once per launch).
    float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t gblk = 0;
    bool have = false;
    auto need = [&](const uint32_t blk) {
        if (!have || gblk != blk) {
            philox_gauss6_f32<MB_PHILOX_SPLIT>(key0, key1, blk, 0u, member, bm_scale, g, P.philox_m0, P.philox_m1, P.bm_mask_r, P.bm_mask_a);
            gblk = blk;
            have = true;
        }
    };
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        // MP: every member stops at ITS OWN state index of sample k; a launch that records samples ends on the last of
        // them, a pure-advance launch (k0 == k1) never steps past the member's next sample
        const uint64_t tgt = !MP ? ((k < P.k1) ? P.target[k] : P.j1)
                             : (k < P.k1) ? member_target(k, dt, mp_Ts)
                             : (P.k1 > P.k0) ? j : min(P.j1, member_target(P.k0, dt, mp_Ts));
        if (NOISE == NOISE_PHILOX_PACKED) {
            if ((j & 1) && j < tgt) {          // odd step: second half of its block
                need((uint32_t)(j >> 1));
                advance(V3{widen_f32(g[3]), widen_f32(g[4]), widen_f32(g[5])}, tab + (j - tab0));
                ++j;
            }
            const uint32_t pairs = (!MP || j < tgt) ? (uint32_t)((tgt - j) >> 1) : 0u;   // MP: a member may already be past P.j1
            uint32_t blk = (uint32_t)(j >> 1);
            const double2* tp = tab + (j - tab0);
            if (pairs != 0) need(blk);
            double2 ha = make_double2(0.0, 0.0), hb = ha;
            if (PREFETCH_TAB && pairs != 0) { ha = __ldg(tp); hb = __ldg(tp + 1); }
            for (uint32_t i = pairs; i != 0; --i, tp += 2) {
                float gn[6];
                double2 na, nb;
                if (PREFETCH_TAB) { na = __ldg(tp + 2); nb = __ldg(tp + 3); }   // the table is allocated with a margin of rows
                philox_gauss6_f32<MB_PHILOX_SPLIT>(key0, key1, ++blk, 0u, member, bm_scale, gn, P.philox_m0, P.philox_m1, P.bm_mask_r, P.bm_mask_a);
                if (PREFETCH_TAB) {
                    advance_h(V3{widen_f32(g[0]), widen_f32(g[1]), widen_f32(g[2])}, ha);
                    advance_h(V3{widen_f32(g[3]), widen_f32(g[4]), widen_f32(g[5])}, hb);
                    ha = na; hb = nb;
                } else {
                    advance(V3{widen_f32(g[0]), widen_f32(g[1]), widen_f32(g[2])}, tp);
                    advance(V3{widen_f32(g[3]), widen_f32(g[4]), widen_f32(g[5])}, tp + 1);
                }
#pragma unroll
                for (int q = 0; q < 6; ++q) g[q] = gn[q];
                gblk = blk;
            }
            j += 2ull * pairs;
            if (j < tgt) {                     // one more (even) step before the sample: first half of its block
                need((uint32_t)(j >> 1));
                advance(V3{widen_f32(g[0]), widen_f32(g[1]), widen_f32(g[2])}, tab + (j - tab0));
                ++j;
            }
        } else {
            const double2* tp = tab + (j - tab0);
#pragma unroll 2
            for (; j < tgt; ++j, ++tp) advance(draw_scaled<NOISE>(P, key0, key1, j, 0u, member, r, ch, bm_scale), tp);
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
        if (MP) P.member_j[r] = (uint32_t)j;
    }
}

template <int NOISE, bool TAB, bool AXIS_Z, int MINB>
static void launch_hs2(bool renorm, dim3 g, dim3 b, cudaStream_t s, const RunParams& P) {
    if (renorm) heun_single_kernel<NOISE, TAB, AXIS_Z, true, MINB><<<g, b, 0, s>>>(P);
    else heun_single_kernel<NOISE, TAB, AXIS_Z, false, MINB><<<g, b, 0, s>>>(P);
}

template <int NOISE, int MINB>
static void launch_hs(bool tab, bool axis_z, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(SINGLE_THREADS);
    const bool renorm = P.renorm != 0;
    if (tab) {
        if (axis_z) launch_hs2<NOISE, true, true, MINB>(renorm, g, b, s, P);
        else launch_hs2<NOISE, true, false, MINB>(renorm, g, b, s, P);
    } else {
        if (axis_z) launch_hs2<NOISE, false, true, MINB>(renorm, g, b, s, P);
        else launch_hs2<NOISE, false, false, MINB>(renorm, g, b, s, P);
    }
}

// per-member material parameters: general-axis arithmetic, packed Philox or injected increments
template <int NOISE>
static void launch_hs_mp(bool tab, dim3 g, dim3 b, cudaStream_t s, const RunParams& P) {
    const bool renorm = P.renorm != 0;
    if (tab) {
        if (renorm) heun_single_kernel<NOISE, true, false, true, 1, true><<<g, b, 0, s>>>(P);
        else heun_single_kernel<NOISE, true, false, false, 1, true><<<g, b, 0, s>>>(P);
    } else {
        if (renorm) heun_single_kernel<NOISE, false, false, true, 1, true><<<g, b, 0, s>>>(P);
        else heun_single_kernel<NOISE, false, false, false, 1, true><<<g, b, 0, s>>>(P);
    }
}

// min_blocks = 1, 7 or K1_LATENCY (production noise mode only; the other modes have the one free-allocation instantiation)
cudaError_t launch_heun_single(int noise, bool tab, bool axis_z, int min_blocks, unsigned grid, cudaStream_t s,
                               const RunParams& P) {
    if (P.mp_dt != nullptr) {
        if (noise == NOISE_INJECTED) launch_hs_mp<NOISE_INJECTED>(tab, dim3(grid), dim3(SINGLE_THREADS), s, P);
        else launch_hs_mp<NOISE_PHILOX_PACKED>(tab, dim3(grid), dim3(SINGLE_THREADS), s, P);
        return cudaGetLastError();
    }
    switch (noise) {
        case NOISE_PHILOX_F32: launch_hs<NOISE_PHILOX_F32, 1>(tab, axis_z, grid, s, P); break;
        case NOISE_PHILOX_F64: launch_hs<NOISE_PHILOX_F64, 1>(tab, axis_z, grid, s, P); break;
        case NOISE_INJECTED: launch_hs<NOISE_INJECTED, 1>(tab, axis_z, grid, s, P); break;
        case NOISE_PHILOX_COARSE: launch_hs<NOISE_PHILOX_COARSE, 1>(tab, axis_z, grid, s, P); break;
        default:
            if (min_blocks == 7) launch_hs<NOISE_PHILOX_PACKED, 7>(tab, axis_z, grid, s, P);
            else if (min_blocks == K1_LATENCY && tab) launch_hs<NOISE_PHILOX_PACKED, K1_LATENCY>(true, axis_z, grid, s, P);
            else launch_hs<NOISE_PHILOX_PACKED, 1>(tab, axis_z, grid, s, P);
            break;
    }
    return cudaGetLastError();
}

// resident CTAs per SM of the production instantiation <PACKED, tab, axis_z, renorm, min_blocks> on the current device
int heun_single_resident_ctas(bool tab, bool axis_z, bool renorm, int min_blocks) {
    int n = 0;
    auto q = [&](auto kernel) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, SINGLE_THREADS, 0); };
#define MB_Q(T, A, RN) \
    (min_blocks == 7 ? q(heun_single_kernel<NOISE_PHILOX_PACKED, T, A, RN, 7>) \
                     : q(heun_single_kernel<NOISE_PHILOX_PACKED, T, A, RN, 1>))
    if (tab) {
        if (axis_z) { if (renorm) MB_Q(true, true, true); else MB_Q(true, true, false); }
        else { if (renorm) MB_Q(true, false, true); else MB_Q(true, false, false); }
    } else {
        if (axis_z) { if (renorm) MB_Q(false, true, true); else MB_Q(false, true, false); }
        else { if (renorm) MB_Q(false, false, true); else MB_Q(false, false, false); }
    }
#undef MB_Q
    return n;
}

}  // namespace mb
