// rng.cuh — in-kernel counter-based Gaussian stream (replaces lib/rng.cpp:14-24,
// std::mt19937_64 + std::normal_distribution, which cannot be reproduced on the device
// at speed).  Philox4x32-10 (Salmon et al. SC'11) is a pure function of
//   key     = 64-bit seed of the ensemble member (magpy/model.py:202-203 seeds)
//   counter = (step index, particle | block<<24, global member index)
// so the Wiener increment of (member, particle, step) does not depend on how the
// ensemble is chunked in time, laid out over CTAs, or sharded over GPUs.
//
// Two Gaussian transforms of the Philox words:
//   GAUSS_F32: Box-Muller in fp32 on the otherwise idle FP32/SFU pipes, widened to fp64
//              by integer bit manipulation (keeps the FP64 pipe for the integrator).
//              One Philox call yields the 3 draws of a particle-step.
//   GAUSS_F64: Box-Muller in fp64 from 53-bit uniforms (two Philox calls per particle-step).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mb {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                       uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// exact float -> double widening without the (quarter-rate) F2F.F64.F32 conversion:
// rebias the exponent (+896) and shift the mantissa.  Exact for normal floats; a zero or
// denormal input (probability < 2^-100 per Gaussian) maps to a value below 2^-126.
__device__ __forceinline__ double widen_f32(float f) {
    const uint32_t b = __float_as_uint(f);
    const uint32_t hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
    const uint32_t lo = b << 29;
    return __hiloint2double((int)hi, (int)lo);
}

// (0,1] uniform from 32 bits, then r = sqrt(-2 ln u)
__device__ __forceinline__ float bm_radius_f32(uint32_t x) {
    const float u = fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    return sqrtf(-2.0f * __logf(u));
}

struct Gauss3 {
    double x, y, z;
};

template <int GAUSS_MODE>
__device__ __forceinline__ Gauss3 philox_gauss3(uint32_t k0, uint32_t k1, uint64_t step, uint32_t particle,
                                                uint32_t member) {
    Gauss3 g;
    if (GAUSS_MODE == 0) {
        uint32_t c0 = (uint32_t)step, c1 = (uint32_t)(step >> 32), c2 = particle, c3 = member;
        philox4x32_10(c0, c1, c2, c3, k0, k1);
        const float r1 = bm_radius_f32(c0), r2 = bm_radius_f32(c2);
        float s1, co1;
        __sincosf(6.2831853071795865f * ((float)c1 * 2.3283064365386963e-10f), &s1, &co1);
        const float co2 = __cosf(6.2831853071795865f * ((float)c3 * 2.3283064365386963e-10f));
        g.x = widen_f32(r1 * co1);
        g.y = widen_f32(r1 * s1);
        g.z = widen_f32(r2 * co2);
    } else {
        uint32_t c0 = (uint32_t)step, c1 = (uint32_t)(step >> 32), c2 = particle, c3 = member;
        philox4x32_10(c0, c1, c2, c3, k0, k1);
        const double two53 = 1.1102230246251565e-16;  // 2^-53
        double u1 = ((double)((((uint64_t)c1 << 32) | c0) >> 11) + 0.5) * two53;
        double u2 = ((double)((((uint64_t)c3 << 32) | c2) >> 11) + 0.5) * two53;
        double r = sqrt(-2.0 * log(u1)), s, c;
        sincospi(2.0 * u2, &s, &c);
        g.x = r * c;
        g.y = r * s;
        c0 = (uint32_t)step; c1 = (uint32_t)(step >> 32); c2 = particle | (1u << 24); c3 = member;
        philox4x32_10(c0, c1, c2, c3, k0, k1);
        u1 = ((double)((((uint64_t)c1 << 32) | c0) >> 11) + 0.5) * two53;
        u2 = ((double)((((uint64_t)c3 << 32) | c2) >> 11) + 0.5) * two53;
        r = sqrt(-2.0 * log(u1));
        sincospi(2.0 * u2, &s, &c);
        g.z = r * c;
    }
    return g;
}

}  // namespace mb
