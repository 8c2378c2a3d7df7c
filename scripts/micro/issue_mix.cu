// Micro-benchmark: how many issue cycles does a DFMA cost next to FP32 / INT work on sm_100a?
// Each kernel runs K independent DFMA chains plus `MIX` independent non-FP64 ops per DFMA.
#include <cstdio>
#include <cuda_runtime.h>

template <int MIX, int KIND>
__global__ void mix_kernel(double* out, float* fout, int iters, double a, double b, float fa, unsigned ia) {
    double x[8];
    float y[16];
    unsigned z[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) { y[i] = threadIdx.x * 0.5f + i; z[i] = threadIdx.x * 7 + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = fma(x[i], a, b);
#pragma unroll
                for (int m = 0; m < MIX; ++m) {
                    const int q = (i * MIX + m) & 15;
                    if (KIND == 0) y[q] = fmaf(y[q], fa, 1.0f);
                    else if (KIND == 1) z[q] = z[q] * ia + 12345u;          // IMAD
                    else z[q] = (z[q] ^ ia) + (z[q] >> 3);                   // LOP3/SHF/IADD mix
                }
            }
        }
    }
    double s = 0; float t = 0; unsigned w = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) { t += y[i]; w += z[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    fout[blockIdx.x * blockDim.x + threadIdx.x] = t + (float)w;
}

template <int MIX, int KIND>
void run(const char* name, int warps_per_smsp) {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int threads = 128, blocks = p.multiProcessorCount * warps_per_smsp, iters = 4096;
    double* d; float* f;
    cudaMalloc(&d, sizeof(double) * blocks * threads); cudaMalloc(&f, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        mix_kernel<MIX, KIND><<<blocks, threads>>>(d, f, iters, 0.999999, 1e-7, 0.999f, 2654435761u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    // warp-level DFMAs per SMSP: each SMSP has warps_per_smsp warps (blocks*4 warps over SMs*4 SMSPs)
    const double dfma_per_warp = (double)iters * 4 * 8;
    const double cycles = best * 1e-3 * khz * 1e3;
    printf("%-28s warps/SMSP=%2d  %.3f ms  cycles per DFMA-slot (per SMSP) = %.3f  [%d other ops per DFMA]\n", name,
           warps_per_smsp, best, cycles / (dfma_per_warp * warps_per_smsp), MIX);
    cudaFree(d); cudaFree(f);
}

int main() {
    for (int w : {4, 8}) {
        run<0, 0>("DFMA only", w);
        run<1, 0>("DFMA + 1 FFMA", w);
        run<2, 0>("DFMA + 2 FFMA", w);
        run<3, 0>("DFMA + 3 FFMA", w);
        run<1, 1>("DFMA + 1 IMAD", w);
        run<2, 1>("DFMA + 2 IMAD", w);
        run<1, 2>("DFMA + 1 (LOP3,SHF,IADD)", w);
    }
    return 0;
}
