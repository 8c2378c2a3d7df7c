#include <cstdlib>
#include "common.cuh"
#include "launch.h"

namespace mb {

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

// Trajectory fetch: the kernels store sampled points as in[s][q][r] (member index fastest: one coalesced 256 B line per
// warp and component); the API wants out[r][q][s] (include/magpy_b200.h, out_trajectories).  For a member the n S
// values are one contiguous run of the output, so a CTA takes 32 members x TRAJ_TILE_K consecutive output entries
// k' = q S + s: every load is a full 256 B line of 32 members, every store a full 256 B line of a member's run (the
// generic 32 x 32 tile transpose writes ragged rows whenever S is not a multiple of 32: 2.95 TB/s at S = 101).
template <int TRAJ_TILE_K>
__global__ void __launch_bounds__(256) traj_fetch_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                         const uint64_t R, const uint32_t n, const uint32_t S,
                                                         const double scale) {
    __shared__ double tile[TRAJ_TILE_K][33];
    const uint32_t K = n * S;
    const uint64_t r0 = (uint64_t)blockIdx.x * 32;
    const uint32_t k0 = blockIdx.y * TRAJ_TILE_K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t kk = warp; kk < TRAJ_TILE_K; kk += 8) {
        const uint32_t k = k0 + kk;
        if (k < K && r0 + lane < R) {
            const uint32_t q = k / S, sidx = k - q * S;
            tile[kk][lane] = in[((uint64_t)sidx * n + q) * R + r0 + lane];
        }
    }
    __syncthreads();
    const uint32_t kn = K - k0 < TRAJ_TILE_K ? K - k0 : TRAJ_TILE_K;   // valid entries of this tile
    for (uint32_t rr = warp; rr < 32; rr += 8) {
        if (r0 + rr >= R) break;
        double* dst = out + (r0 + rr) * K + k0;
        for (uint32_t kk = lane; kk < kn; kk += 32) dst[kk] = tile[kk][rr] * scale;
    }
}

// The same for short runs (n S <= 800 doubles): the CTA stages ALL n S entries of its 32 members (dynamic shared
// memory), whose output is then one contiguous, 256 B-aligned region written linearly.
__global__ void __launch_bounds__(512) traj_fetch_full_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                              const uint64_t R, const uint32_t n, const uint32_t S,
                                                              const double scale) {
    extern __shared__ double ftile[];   // [K][33]
    const uint32_t K = n * S;
    const uint64_t r0 = (uint64_t)blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t k = warp; k < K; k += 16) {
        if (r0 + lane < R) {
            const uint32_t q = k / S, sidx = k - q * S;
            ftile[k * 33 + lane] = in[((uint64_t)sidx * n + q) * R + r0 + lane];
        }
    }
    __syncthreads();
    const uint32_t members = R - r0 < 32 ? (uint32_t)(R - r0) : 32u;
    const uint32_t total = members * K;
    double* dst = out + r0 * K;
    for (uint32_t e = threadIdx.x; e < total; e += 512) {
        const uint32_t rr = e / K, k = e - rr * K;
        dst[e] = ftile[k * 33 + rr] * scale;
    }
}

// out[q][r] = in[q] for q < n, r < R (shared initial state replicated over the members)
__global__ void broadcast_rows_kernel(const double* in, double* out, uint64_t n, uint64_t R) {
    const uint64_t total = n * R;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = in[i / R];
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

// 12 independent DMMA.8x8x4 accumulator chains per warp (512 flop per warp instruction)
__global__ void fp64_mma_peak_kernel(double* out, int iters, double a0, double b0) {
    double c[12][2];
    const double a = a0 + threadIdx.x * 1e-9, b = b0 + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < 12; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 12; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += c[i][0] + c[i][1];
    out[(uint64_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
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

// Moments and histograms of the Gaussian stream, accumulated on the device so that the statistical tests can look at
// 1e9+ draws (tail probabilities, Kolmogorov-Smirnov, uniformity of the Box-Muller angle): thread = member, loop over
// steps, 3 draws per step exactly as the integration kernels consume them.
constexpr int GS_BINS = 4096, GS_ANGLE_BINS = 1024;
template <int GAUSS_MODE>
__global__ void __launch_bounds__(256) gauss_stats_kernel(uint64_t seed, uint64_t first_member, uint64_t n_members,
                                                          uint64_t n_steps, unsigned long long* hist,
                                                          unsigned long long* angle_hist, double* moments) {
    __shared__ unsigned int sh[GS_BINS], sa[GS_ANGLE_BINS];
    __shared__ double red[8 * 5];
    for (int i = threadIdx.x; i < GS_BINS; i += blockDim.x) sh[i] = 0;
    for (int i = threadIdx.x; i < GS_ANGLE_BINS; i += blockDim.x) sa[i] = 0;
    __syncthreads();
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0, mx = 0;
    if (t < n_members) {
        const uint32_t member = (uint32_t)(first_member + t);
        for (uint64_t step = 1; step <= n_steps; ++step) {
            const Gauss3 g = philox_gauss3<GAUSS_MODE>((uint32_t)seed, (uint32_t)(seed >> 32), step, 0u, member);
            const double v[3] = {g.x, g.y, g.z};
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const double z = v[q], z2 = z * z;
                s1 += z; s2 += z2; s3 += z2 * z; s4 += z2 * z2;
                mx = fmax(mx, fabs(z));
                int b = (int)floor((z + 8.0) * 256.0);
                b = b < 0 ? 0 : (b >= GS_BINS ? GS_BINS - 1 : b);
                atomicAdd(&sh[b], 1u);
            }
            // (x, y) of one Box-Muller pair: every step in the 32-bit / fp64 modes, the even (0-based) steps of the packed one
            if (GAUSS_MODE != 3 || ((step - 1) & 1) == 0) {
                // fraction of a revolution, shifted by half a step of the packed mode's 2^18-direction grid so that no grid
                // direction sits on a bin edge (every bin then holds exactly 256 of them)
                double a = atan2(g.y, g.x) * 0.15915494309189535;
                a = a - floor(a) + 1.9073486328125e-06;
                const int b = ((int)(a * GS_ANGLE_BINS)) & (GS_ANGLE_BINS - 1);
                atomicAdd(&sa[b], 1u);
            }
        }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3); s4 = warp_sum(s4);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[warp * 5] = s1; red[warp * 5 + 1] = s2; red[warp * 5 + 2] = s3; red[warp * 5 + 3] = s4; red[warp * 5 + 4] = mx; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double a = 0;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) a += red[w * 5 + threadIdx.x];
        atomicAdd(moments + threadIdx.x, a);
    } else if (threadIdx.x == 4) {
        double a = 0;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) a = fmax(a, red[w * 5 + 4]);
        // max of non-negative doubles = max of their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(moments + 4), (unsigned long long)__double_as_longlong(a));
    }
    for (int i = threadIdx.x; i < GS_BINS; i += blockDim.x) if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
    for (int i = threadIdx.x; i < GS_ANGLE_BINS; i += blockDim.x) if (sa[i]) atomicAdd(angle_hist + i, (unsigned long long)sa[i]);
}

cudaError_t launch_gauss_stats(int noise, uint64_t seed, uint64_t first_member, uint64_t n_members, uint64_t n_steps,
                               unsigned long long* hist, unsigned long long* angle_hist, double* moments) {
    const unsigned g = (unsigned)((n_members + 255) / 256);
    if (noise == NOISE_PHILOX_PACKED) gauss_stats_kernel<NOISE_PHILOX_PACKED><<<g, 256>>>(seed, first_member, n_members, n_steps, hist, angle_hist, moments);
    else if (noise == NOISE_PHILOX_F64) gauss_stats_kernel<NOISE_PHILOX_F64><<<g, 256>>>(seed, first_member, n_members, n_steps, hist, angle_hist, moments);
    else gauss_stats_kernel<NOISE_PHILOX_F32><<<g, 256>>>(seed, first_member, n_members, n_steps, hist, angle_hist, moments);
    return cudaGetLastError();
}

// device self-test of the 3x3 solve of the quasi-Newton update (llg_math.cuh: solve3_adjugate) on caller matrices
__global__ void solve3_kernel(const double* A, const double* b, double* x, int* ok, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[9], rhs[3], d[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 9; ++q) a[q] = A[9 * i + q];
#pragma unroll
    for (int q = 0; q < 3; ++q) rhs[q] = b[3 * i + q];
    ok[i] = solve3_adjugate(a, rhs, d) ? 1 : 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) x[3 * i + q] = d[q];
}

cudaError_t launch_solve3(const double* A, const double* b, double* x, int* ok, uint64_t n) {
    solve3_kernel<<<(unsigned)((n + 127) / 128), 128>>>(A, b, x, ok, n);
    return cudaGetLastError();
}

cudaError_t launch_field_table(double* tab, uint64_t j0, uint64_t n_steps, double dt, double second_offset, int shape,
                               double h0, double f_red, cudaStream_t s) {
    field_table_kernel<<<(unsigned)((n_steps + 255) / 256), 256, 0, s>>>(tab, j0, n_steps, dt, second_offset, shape, h0, f_red);
    return cudaGetLastError();
}

cudaError_t launch_reduce_partials(const double* partial, double* sums, uint32_t k0, uint32_t n_samples, uint32_t n_cta,
                                   cudaStream_t s) {
    reduce_partials_kernel<<<n_samples, 256, 0, s>>>(partial, sums, k0, n_cta);
    return cudaGetLastError();
}

cudaError_t launch_transpose(const double* in, double* out, uint64_t rows, uint64_t cols, uint64_t batches, uint64_t in_bs,
                             uint64_t in_rs, uint64_t out_bs, uint64_t out_rs, double scale, cudaStream_t s) {
    const dim3 g((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)batches), b(32, 8);
    transpose_kernel<<<g, b, 0, s>>>(in, out, rows, cols, in_bs, in_rs, out_bs, out_rs, scale);
    return cudaGetLastError();
}

cudaError_t launch_traj_fetch(const double* in, double* out, uint64_t R, uint32_t n, uint32_t S, double scale,
                              cudaStream_t s) {
    const uint64_t gx = (R + 31) / 32;
    if ((uint64_t)n * S <= 800 && gx <= 0x7FFFFFFFull && !std::getenv("MAGPY_B200_TRAJ_TILE")) {
        const size_t smem = (size_t)n * S * 33 * sizeof(double);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(traj_fetch_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        traj_fetch_full_kernel<<<(unsigned)gx, 512, smem, s>>>(in, out, R, n, S, scale);
        return cudaGetLastError();
    }
    int tk = 128;
    if (const char* env = std::getenv("MAGPY_B200_TRAJ_TILE")) tk = std::atoi(env);   // tuning knob: 32, 64 or 128
    const uint32_t gy = (n * S + tk - 1) / tk;
    if (gx > 0x7FFFFFFFull || gy > 65535) return cudaErrorInvalidValue;
    const dim3 g((unsigned)gx, gy);
    if (tk == 64) traj_fetch_kernel<64><<<g, 256, 0, s>>>(in, out, R, n, S, scale);
    else if (tk == 32) traj_fetch_kernel<32><<<g, 256, 0, s>>>(in, out, R, n, S, scale);
    else if (tk == 128) traj_fetch_kernel<128><<<g, 256, 0, s>>>(in, out, R, n, S, scale);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_broadcast_rows(const double* in, double* out, uint64_t n, uint64_t R, cudaStream_t s) {
    const uint64_t total = n * R;
    const unsigned g = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    broadcast_rows_kernel<<<g, 256, 0, s>>>(in, out, n, R);
    return cudaGetLastError();
}

cudaError_t launch_fp64_peak(double* out, int blocks, int threads, int iters, cudaStream_t s) {
    fp64_peak_kernel<<<blocks, threads, 0, s>>>(out, iters, 0.999999, 1e-7);
    return cudaGetLastError();
}

cudaError_t launch_fp64_mma_peak(double* out, int blocks, int threads, int iters, cudaStream_t s) {
    fp64_mma_peak_kernel<<<blocks, threads, 0, s>>>(out, iters, 1.0000001, 1e-9);
    return cudaGetLastError();
}

cudaError_t launch_philox_words(const uint32_t ctr[4], const uint32_t key[2], uint32_t* out) {
    philox_words_kernel<<<1, 1>>>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
    return cudaGetLastError();
}

cudaError_t launch_gaussians(int noise, uint64_t seed, uint32_t member, uint32_t particle, uint64_t first_step,
                             uint64_t n_steps, double* out) {
    const unsigned g = (unsigned)((n_steps + 255) / 256);
    if (noise == NOISE_PHILOX_PACKED) gaussians_kernel<NOISE_PHILOX_PACKED><<<g, 256>>>(seed, member, particle, first_step, n_steps, out);
    else if (noise == NOISE_PHILOX_F64) gaussians_kernel<NOISE_PHILOX_F64><<<g, 256>>>(seed, member, particle, first_step, n_steps, out);
    else gaussians_kernel<NOISE_PHILOX_F32><<<g, 256>>>(seed, member, particle, first_step, n_steps, out);
    return cudaGetLastError();
}

}  // namespace mb
