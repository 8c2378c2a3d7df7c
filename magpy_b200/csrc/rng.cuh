// rng.cuh — in-kernel counter-based Gaussian stream (replaces lib/rng.cpp:14-24,
// std::mt19937_64 + std::normal_distribution, which cannot be reproduced on the device
// at speed).  Philox4x32-10 (Salmon et al. SC'11) is used as a pure function of
//   counter = (step index — or step-pair index in the packed mode — [32 bit, like the reference's step
//              counter], global member index, seed low word, seed high word)
//              -- the member's seed of magpy/model.py:202-203
//   key     = (particle | mode/block tag << 24, MB_PHILOX_KEY1)
// so the Wiener increment of (member, particle, step) does not depend on how the ensemble is
// chunked in time, laid out over CTAs, or sharded over GPUs.  Everything that varies per thread
// sits in the counter: the key schedule (20 adds per block) is warp-uniform and, for the
// single-particle kernels, a compile-time constant.
//
// Three Gaussian transforms of the Philox words:
//   GAUSS_F32: Box-Muller in fp32 on the otherwise idle FP32/SFU pipes, widened to fp64
//              (keeps the FP64 pipe for the integrator).  One Philox call yields the 3 draws of
//              a particle-step.
//   GAUSS_F64: Box-Muller in fp64 from 53-bit uniforms (two Philox calls per particle-step).
//   GAUSS_F32_PACKED (production default): fp32 Box-Muller with 22-bit (cell-midpoint) radius / 18-bit angle
//              uniforms, so that one Philox block feeds TWO particle-steps (the wide integer
//              multiplies of Philox take FP64-pipe time on sm_100a — profiles/README.md — so halving
//              them is what raises the FP64 pipe's share of the cycle).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mb {

// 32x32 -> 64 bit product as one IMAD.WIDE.U32 whose halves are read straight from the
// register pair (the plain C++ form makes ptxas add a zero carry word to every high half)
__host__ __device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    unsigned long long p;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
#else
    const uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                       uint32_t k0, uint32_t k1) {
#ifndef MB_ABL_ROUNDS
#define MB_ABL_ROUNDS 10   // timing ablation only (scripts/experiments/probe_k1_ablation.py): anything but 10 is not Philox4x32-10
#endif
#pragma unroll
    for (int r = 0; r < MB_ABL_ROUNDS; ++r) {
        uint32_t h0, l0, h1, l1;
        mulhilo32(0xD2511F53u, c0, h0, l0);
        mulhilo32(0xCD9E8D57u, c2, h1, l1);
        const uint32_t n0 = h1 ^ c1 ^ k0;
        const uint32_t n2 = h0 ^ c3 ^ k1;
        c1 = l1;
        c3 = l0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// The same block with the wide multiplies of the first SPLIT rounds issued as IMAD.HI (immediate multiplier) + IMAD
// (multiplier from a register, m0r / m1r = the same constants passed at run time so that ptxas cannot fuse the pair
// back into one IMAD.WIDE).  On sm_100a IMAD.WIDE takes FP64-pipe time, IMAD.HI and IMAD do not (profiles/README.md):
// splitting SOME rounds moves work from the saturated FP64 side to the 32-bit integer pipe.
template <int SPLIT>
__device__ __forceinline__ void philox4x32_10_split(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                                    uint32_t k1, const uint32_t m0r, const uint32_t m1r) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, l0, h1, l1;
        if (r < SPLIT) {
            h0 = __umulhi(c0, 0xD2511F53u); l0 = c0 * m0r;
            h1 = __umulhi(c2, 0xCD9E8D57u); l1 = c2 * m1r;
        } else {
            mulhilo32(0xD2511F53u, c0, h0, l0);
            mulhilo32(0xCD9E8D57u, c2, h1, l1);
        }
        const uint32_t n0 = h1 ^ c1 ^ k0;
        const uint32_t n2 = h0 ^ c3 ^ k1;
        c1 = l1;
        c3 = l0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// float -> double widening of a Gaussian increment.  Measured on B200 (profiles/README.md): in the
// issue-mix micro-benchmark one F2F.F64.F32 takes ~3.5 cycles of the FP64 pipe away from a DFMA stream,
// but assembling the double from the float's bits (LOP3, LEA.HI, LOP3, SHF: exponent re-biased by 896,
// mantissa shifted by 3) makes the Heun kernel 9 % slower end to end (2.22e11 vs 2.44e11 particle-steps/s):
// the kernel has no issue slots to spare for 24 more instructions per step pair.  F2F stays.
__device__ __forceinline__ double widen_f32(float f) {
#ifdef MB_ABL_NOF2F   // timing ablation: a reinterpretation instead of the conversion (wrong values, same dependencies)
    return __hiloint2double((int)(__float_as_uint(f) >> 4), 0);
#elif defined(MB_WIDEN_INT)
    const uint32_t b = __float_as_uint(f);
    const uint32_t hi = (((b & 0x7fffffffu) >> 3) + 0x38000000u) | (b & 0x80000000u);
    const uint32_t lo = b << 29;
    return __hiloint2double((int)hi, (int)lo);
#else
    return (double)f;
#endif
}

__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {   // one MUFU.SQRT; sqrt(0) = 0
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Box-Muller radius times `amp`:  amp * sqrt(-2 ln u),  u in (0,1) from 32 bits.
// neg2ln2_amp2 = -2 ln(2) amp^2 folds the scale into the one multiply that follows lg2, so a
// scaled draw costs the same as a unit one.  The integer is converted toward zero so that u
// never rounds to 1.  Branch free: I2FP, FFMA, MUFU.LG2, FMUL, MUFU.SQRT.
__device__ __forceinline__ float bm_radius_f32(uint32_t x, float neg2ln2_amp2) {
    const float u = fmaf(__uint2float_rz(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    return sqrt_approx(lg2_approx(u) * neg2ln2_amp2);
}

#define MB_TWO_PI_2M32 1.4629180792671596e-9f /* 2 pi 2^-32 */
#define MB_PHILOX_KEY1 0xB2005EEDu
#define MB_NEG_2LN2 (-1.3862943611198906f)

// three draws of N(0, amp^2) in fp32 from one Philox block: (member, particle, step) -> words
// (w0,w1,w2,w3); pair 1 = radius(w0), angle(w1) gives x = r cos, y = r sin; pair 2 = radius(w2),
// angle(w3) gives z = r cos.
__device__ __forceinline__ void philox_gauss3_f32(uint32_t seed_lo, uint32_t seed_hi, uint64_t step,
                                                  uint32_t particle, uint32_t member, float neg2ln2_amp2,
                                                  float& gx, float& gy, float& gz) {
    uint32_t c0 = (uint32_t)step, c1 = member, c2 = seed_lo, c3 = seed_hi;
    philox4x32_10(c0, c1, c2, c3, particle, MB_PHILOX_KEY1);
    const float r1 = bm_radius_f32(c0, neg2ln2_amp2), r2 = bm_radius_f32(c2, neg2ln2_amp2);
    const float a1 = (float)c1 * MB_TWO_PI_2M32, a2 = (float)c3 * MB_TWO_PI_2M32;
    gx = r1 * __cosf(a1);
    gy = r1 * __sinf(a1);
    gz = r2 * __cosf(a2);
}

// Packed mode: SIX draws of N(0, amp^2) from one Philox block, i.e. the increments of the two
// steps s = 2b and s = 2b + 1 (s = 0-based step index) of one (member, particle).  The 128 bits are cut into three
// (23-bit radius field, 18-bit angle) pairs (layout in philox_gauss6_f32); the top 22 bits of each radius field are used,
// as cell midpoints (see bm_pair_packed; radius up to 5.65 sigma); 2^18 equidistant directions reproduce every circular
// moment below order 2^18.
//   pair 0 = (w0[22:0], {w1:w0}[40:23])   -> g[0], g[1]
//   pair 1 = (w1[31:9], {w3:w2}[40:23])   -> g[2], g[3]
//   pair 2 = (w2[22:0], w3[31:14])        -> g[4], g[5]
// g[0..2] belong to the even step, g[3..5] to the odd one.
#define MB_PACKED_KEY_TAG (2u << 24)

// Both uniforms are turned into floats in [1, 2) by OR-ing the bits into the mantissa (one LOP3, no
// integer-to-float conversion) and the angle needs no offset at all because sin/cos have period one revolution.
// The radius uniform is u = 2 - f with the lowest mantissa bit of f forced to one: u = (k + 1/2) 2^-22, k = 0 ... 2^22 - 1
// — the MIDPOINTS of a 22-bit grid on (0, 1), never 0 or 1.  (With all 23 bits, u = k 2^-23 samples the right end of
// every cell: a first-order bias that leaves P(|z| > 5) 10 % short — 2.4 standard errors at the 1e9 draws of the tail
// test; midpoints make the discretisation error second order: < 1 % at 5 sigma.  Same instruction count.)  Largest
// radius sqrt(2 ln 2^23) = 5.65: the density beyond it (1.6e-8 per draw) is the stream's truncation.
__device__ __forceinline__ void bm_pair_packed(uint32_t r23_bits /* 23 bits, already in mantissa position */,
                                               uint32_t a18_bits /* 18 bits, in the top of the mantissa */,
                                               float neg2ln2_amp2, float& c, float& s) {
    const float u = 2.0f - __uint_as_float(r23_bits | 0x3f800001u);          // (k + 1/2) 2^-22 in (0, 1)
#ifdef MB_ABL_NOMUFU   // timing ablation: the four MUFU of a pair replaced by FMULs
    const float r = (u * neg2ln2_amp2) * 0.7f;
    const float a = __uint_as_float(a18_bits | 0x3f800000u) * 6.283185307179586f;
    c = r * (a * 0.3f);
    s = r * (a * 0.4f);
#else
    const float r = sqrt_approx(lg2_approx(u) * neg2ln2_amp2);
    const float a = __uint_as_float(a18_bits | 0x3f800000u) * 6.283185307179586f;   // 2 pi (1 + k 2^-18)
    c = r * __cosf(a);
    s = r * __sinf(a);
#endif
}

template <int SPLIT = 0>
__device__ __forceinline__ void philox_gauss6_f32(uint32_t seed_lo, uint32_t seed_hi, uint64_t pair_index,
                                                  uint32_t particle, uint32_t member, float neg2ln2_amp2,
                                                  float (&g)[6], const uint32_t m0r = 0xD2511F53u,
                                                  const uint32_t m1r = 0xCD9E8D57u, const uint32_t mask_r = 0x007fffffu,
                                                  const uint32_t mask_a = 0x007fffe0u) {
    uint32_t w0 = (uint32_t)pair_index, w1 = member, w2 = seed_lo, w3 = seed_hi;
    if (SPLIT > 0) philox4x32_10_split<SPLIT>(w0, w1, w2, w3, particle | MB_PACKED_KEY_TAG, MB_PHILOX_KEY1, m0r, m1r);
    else philox4x32_10(w0, w1, w2, w3, particle | MB_PACKED_KEY_TAG, MB_PHILOX_KEY1);
    // Field layout: each 64-bit half {w1:w0}, {w3:w2} is [23-bit radius field in place | 18-bit angle | 23 bits].  A field
    // that already sits in mantissa position costs ONE LOP3 ((w & mask) | exponent — the masks arrive in registers in the
    // single-particle kernels so that ptxas keeps it one instruction), one that straddles the words a funnel shift more:
    // 10 instructions per block against the 18 of the round-1 layout (most-significant-first string of 24 + 18 bit pairs).
    // The kernel is bound by instruction issue, so these count.
    const uint32_t r0 = w0 & mask_r;                                   // {w1:w0} bits  0..22
    const uint32_t a0 = __funnelshift_r(w0, w1, 18) & mask_a;          // {w1:w0} bits 23..40 -> mantissa bits 5..22
    const uint32_t r1 = w1 >> 9;                                       // {w1:w0} bits 41..63
    const uint32_t r2 = w2 & mask_r;                                   // {w3:w2} bits  0..22
    const uint32_t a1 = __funnelshift_r(w2, w3, 18) & mask_a;          // {w3:w2} bits 23..40
    const uint32_t a2 = (w3 >> 9) & mask_a;                            // {w3:w2} bits 46..63 (41..45 unused)
    bm_pair_packed(r0, a0, neg2ln2_amp2, g[0], g[1]);
    bm_pair_packed(r1, a1, neg2ln2_amp2, g[2], g[3]);
    bm_pair_packed(r2, a2, neg2ln2_amp2, g[4], g[5]);
}

struct Gauss3 {
    double x, y, z;
};

// unit-variance draws in fp64 (validation mode and the implicit kernels' unscaled draws)
template <int GAUSS_MODE>
__device__ __forceinline__ Gauss3 philox_gauss3(uint32_t seed_lo, uint32_t seed_hi, uint64_t step,
                                                uint32_t particle, uint32_t member) {
    Gauss3 g;
    if (GAUSS_MODE == 3) {   // packed (NOISE_PHILOX_PACKED): `step` is 1-based (the reference's counter), blocks pair steps (0,1), (2,3), ...
        float v[6];
        philox_gauss6_f32(seed_lo, seed_hi, (step - 1) >> 1, particle, member, MB_NEG_2LN2, v);
        const bool odd = ((step - 1) & 1) != 0;
        g.x = widen_f32(odd ? v[3] : v[0]);
        g.y = widen_f32(odd ? v[4] : v[1]);
        g.z = widen_f32(odd ? v[5] : v[2]);
    } else if (GAUSS_MODE == 0) {
        float x, y, z;
        philox_gauss3_f32(seed_lo, seed_hi, step, particle, member, MB_NEG_2LN2, x, y, z);
        g.x = widen_f32(x);
        g.y = widen_f32(y);
        g.z = widen_f32(z);
    } else {
        uint32_t c0 = (uint32_t)step, c1 = member, c2 = seed_lo, c3 = seed_hi;
        philox4x32_10(c0, c1, c2, c3, particle, MB_PHILOX_KEY1);
        const double two53 = 1.1102230246251565e-16;  // 2^-53
        double u1 = ((double)((((uint64_t)c1 << 32) | c0) >> 11) + 0.5) * two53;
        double u2 = ((double)((((uint64_t)c3 << 32) | c2) >> 11) + 0.5) * two53;
        double r = sqrt(-2.0 * log(u1)), s, c;
        sincospi(2.0 * u2, &s, &c);
        g.x = r * c;
        g.y = r * s;
        c0 = (uint32_t)step; c1 = member; c2 = seed_lo; c3 = seed_hi;
        philox4x32_10(c0, c1, c2, c3, particle | (1u << 24), MB_PHILOX_KEY1);
        u1 = ((double)((((uint64_t)c1 << 32) | c0) >> 11) + 0.5) * two53;
        u2 = ((double)((((uint64_t)c3 << 32) | c2) >> 11) + 0.5) * two53;
        r = sqrt(-2.0 * log(u1));
        sincospi(2.0 * u2, &s, &c);
        g.z = r * c;
    }
    return g;
}

}  // namespace mb
