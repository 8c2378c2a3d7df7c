// Micro-benchmark: does the fp64 warp-level MMA (DMMA, mma.sync.aligned.m8n8k4.f64) of sm_100a run at
// the rate of the FP64 vector pipe, faster, or slower — and does it overlap with DFMA?  The dipolar field
// of a cluster is H[3N x members] = D[3N x 3N] . M[3N x members] with a static D shared by all members:
// if DMMA is at least as fast as DFMA it does the same 9 FMA per ordered pair with 8x fewer issued
// instructions and far fewer shared-memory operand loads.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

#if defined(DMMA_BIG)
// sm_90+ shapes
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
#endif

// NACC independent accumulator tiles per warp, NDF DFMA chains per DMMA slot
template <int NACC, int NDF>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
    double x[8];
    const double a = a0 + threadIdx.x * 1e-9, b = b0 + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
                for (int d = 0; d < NDF; ++d) x[(i * NDF + d) & 7] = fma(x[(i * NDF + d) & 7], a0, b0);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the operand pattern of the cluster kernel: per k-step 3 A fragments + 4 B fragments from shared memory
// feed 12 DMMAs (a 24-row x 32-member output tile per warp)
__global__ void dmma_lds_kernel(double* out, int ksteps, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 24 * 196 + 192 * 36; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    const double* A = sm;              // [24][196]
    const double* B = sm + 24 * 196;   // [192][36]
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double c[3][4][2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int ks = 0; ks < ksteps; ++ks) {
            double a[3], b[4];
#pragma unroll
            for (int i = 0; i < 3; ++i) a[i] = A[(8 * i + g) * 196 + 4 * ks + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = B[(4 * ks + t) * 36 + 8 * j + g];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, int NDF>
static void run(const char* name, int warps_per_smsp, double* out, int sms, double ghz) {
    const int threads = 128 * warps_per_smsp;   // 4 SMSPs x warps x 32
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dmma_kernel<NACC, NDF><<<sms, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dmma_kernel<NACC, NDF><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n_mma = (double)iters * 4 * NACC * warps_per_smsp * 4 * sms;
    const double flop = n_mma * 512 + n_mma * NDF * 64;
    const double cyc_per_slot = ms * 1e-3 * ghz * 1e9 / ((double)iters * 4 * NACC * warps_per_smsp);
    printf("%-34s warps/SMSP=%d acc=%2d dfma/slot=%d: %8.3f ms  %6.2f TFLOP/s  %6.2f cycles per slot per SMSP  (%s)\n", name,
           warps_per_smsp, NACC, NDF, ms, flop / ms / 1e9, cyc_per_slot, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 1024 * 4);
    const int sms = p.multiProcessorCount;
    run<12, 0>("DMMA.884 alone", 1, out, sms, ghz);
    run<12, 0>("DMMA.884 alone", 2, out, sms, ghz);
    run<12, 0>("DMMA.884 alone", 4, out, sms, ghz);
    run<4, 0>("DMMA.884 alone", 4, out, sms, ghz);
    run<2, 0>("DMMA.884 alone (latency bound)", 1, out, sms, ghz);
    run<1, 0>("DMMA.884 alone (latency bound)", 1, out, sms, ghz);
    run<12, 1>("DMMA.884 + 1 DFMA", 2, out, sms, ghz);
    run<12, 2>("DMMA.884 + 2 DFMA", 2, out, sms, ghz);
    run<12, 4>("DMMA.884 + 4 DFMA", 2, out, sms, ghz);
    run<12, 8>("DMMA.884 + 8 DFMA", 2, out, sms, ghz);
    // operand-fed version
    for (int w = 1; w <= 4; w *= 2) {
        const int threads = 128 * w, iters = 200, ks = 48;
        const size_t smem = sizeof(double) * (24 * 196 + 192 * 36);
        cudaFuncSetAttribute(dmma_lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        dmma_lds_kernel<<<sms, threads, smem>>>(out, ks, 2);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        dmma_lds_kernel<<<sms, threads, smem>>>(out, ks, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double n_mma = (double)iters * ks * 12 * w * 4 * sms;
        printf("DMMA.884 fed from shared memory (7 LDS.64 per 12 DMMA) warps/SMSP=%d: %8.3f ms  %6.2f TFLOP/s (%s)\n", w, ms,
               n_mma * 512 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
