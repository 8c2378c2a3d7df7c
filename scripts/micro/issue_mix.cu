// Micro-benchmark: what does one DFMA cost in issue/dispatch cycles on sm_100a next to the other
// instruction classes of the Heun kernel (Philox = IMAD.WIDE + LOP3, Box-Muller = MUFU + FMUL,
// widening = F2F.F64.F32)?  Every kernel runs 8 independent DFMA chains per thread (NDF = 1) or none
// (NDF = 0) plus MIX independent "other" ops per DFMA slot.  Reported: SM cycles per slot per SMSP.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

enum { K_FFMA = 0, K_IMAD = 1, K_LOP = 2, K_IMADW = 3, K_MUFU = 4, K_F2F = 5, K_I2F = 6, K_HILO = 7, K_HI = 8, K_PRMT = 9, K_HILO_SPLIT = 10, K_WIDEN_INT = 11, K_I2F_F64 = 12 };

template <int NDF, int MIX, int KIND>
__global__ void mix_kernel(double* out, float* fout, int iters, double a, double b, float fa, unsigned ia) {
    double x[8];
    float y[16];
    unsigned z[16];
    double dacc = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) { y[i] = threadIdx.x * 0.5f + i + 1.0f; z[i] = threadIdx.x * 7 + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int d = 0; d < NDF; ++d) x[i] = fma(x[i], a, b + d);
#pragma unroll
                for (int m = 0; m < MIX; ++m) {
                    const int q = (i * MIX + m) & 15;
                    if (KIND == K_FFMA) y[q] = fmaf(y[q], fa, 1.0f);
                    else if (KIND == K_IMAD) z[q] = z[q] * ia + 12345u;
                    else if (KIND == K_LOP) z[q] = (z[q] ^ ia) + (z[q] >> 3);   // LOP3 / SHF / IADD mix (3 instr)
                    else if (KIND == K_IMADW) {
                        unsigned long long p;
                        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(z[q]), "r"(ia));
                        unsigned lo, hi;
                        asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
                        z[q] = hi ^ lo;                                           // IMAD.WIDE + LOP3
                    } else if (KIND == K_MUFU) {
                        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y[q]) : "f"(y[q]));
                    } else if (KIND == K_F2F) {
                        z[q] ^= (unsigned)__double2hiint((double)y[q]);           // F2F.F64.F32 + LOP3 + FADD
                        y[q] += 1.0f;
                    } else if (KIND == K_I2F) {
                        y[q] += __uint2float_rz(z[q]);                            // I2FP + FADD + IADD
                        z[q] += 3u;
                    } else if (KIND == K_HILO) {
                        z[q] = __umulhi(z[q], ia) ^ (z[q] * ia);                  // IMAD.HI.U32 + IMAD + LOP3
                    } else if (KIND == K_HILO_SPLIT) {
                        // high word against the immediate, low word against the same value held in a register:
                        // ptxas cannot prove the multipliers equal, so it cannot fuse the two into one IMAD.WIDE
                        const unsigned hi = __umulhi(z[q], 2654435761u);
                        z[q] = hi ^ (z[q] * ia);
                    } else if (KIND == K_WIDEN_INT) {
                        // float -> double by integer re-biasing (5 ALU instructions, no XU)
                        const unsigned bits = __float_as_uint(y[q]);
                        const unsigned hi = (bits & 0x80000000u) | (((bits & 0x7fffffffu) >> 3) + 0x38000000u);
                        const unsigned lo = bits << 29;
                        z[q] ^= hi ^ lo;
                        y[q] += 1.0f;
                    } else if (KIND == K_HI) {
                        z[q] = __umulhi(z[q], ia);                                // IMAD.HI.U32
                    } else if (KIND == K_PRMT) {
                        z[q] = __byte_perm(z[q], ia, 0x2103);                     // PRMT
                    }
                }
            }
        }
    }
    double s = dacc; float t = 0; unsigned w = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) { t += y[i]; w += z[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    fout[blockIdx.x * blockDim.x + threadIdx.x] = t + (float)w;
}

template <int NDF, int MIX, int KIND>
void run(const char* name, int warps_per_smsp) {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int threads = 128, blocks = p.multiProcessorCount * warps_per_smsp, iters = 2048;
    double* d; float* f;
    cudaMalloc(&d, sizeof(double) * blocks * threads); cudaMalloc(&f, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        mix_kernel<NDF, MIX, KIND><<<blocks, threads>>>(d, f, iters, 0.999999, 1e-7, 0.999f, 2654435761u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double slots_per_warp = (double)iters * 4 * 8;
    const double cycles = best * 1e-3 * khz * 1e3;
    printf("%-34s warps/SMSP=%2d  %8.3f ms  cycles per slot per SMSP = %6.3f  [%d DFMA + %d other per slot]\n", name,
           warps_per_smsp, best, cycles / (slots_per_warp * warps_per_smsp), NDF, MIX);
    cudaFree(d); cudaFree(f);
}

int main() {
    for (int w : {8}) {
        run<1, 0, K_FFMA>("DFMA only", w);
        run<1, 1, K_FFMA>("DFMA + 1 FFMA", w);
        run<1, 2, K_FFMA>("DFMA + 2 FFMA", w);
        run<1, 3, K_FFMA>("DFMA + 3 FFMA", w);
        run<0, 2, K_FFMA>("2 FFMA", w);
        run<1, 1, K_IMAD>("DFMA + 1 IMAD", w);
        run<1, 2, K_IMAD>("DFMA + 2 IMAD", w);
        run<0, 2, K_IMAD>("2 IMAD", w);
        run<1, 1, K_IMADW>("DFMA + 1 (IMAD.WIDE,LOP3)", w);
        run<1, 2, K_IMADW>("DFMA + 2 (IMAD.WIDE,LOP3)", w);
        run<0, 2, K_IMADW>("2 (IMAD.WIDE,LOP3)", w);
        run<1, 1, K_LOP>("DFMA + 1 (LOP3,SHF,IADD)", w);
        run<0, 1, K_LOP>("1 (LOP3,SHF,IADD)", w);
        run<1, 1, K_MUFU>("DFMA + 1 MUFU", w);
        run<0, 1, K_MUFU>("1 MUFU", w);
        run<1, 1, K_F2F>("DFMA + 1 (F2F.F64.F32,LOP3,FADD)", w);
        run<0, 1, K_F2F>("1 (F2F.F64.F32,LOP3,FADD)", w);
        run<1, 1, K_I2F>("DFMA + 1 (I2FP,FADD,IADD)", w);
        run<0, 1, K_I2F>("1 (I2FP,FADD,IADD)", w);
        run<1, 1, K_HILO>("DFMA + 1 (IMAD.HI,IMAD,LOP3)", w);
        run<1, 2, K_HILO>("DFMA + 2 (IMAD.HI,IMAD,LOP3)", w);
        run<0, 1, K_HILO>("1 (IMAD.HI,IMAD,LOP3)", w);
        run<1, 1, K_HILO_SPLIT>("DFMA + 1 (IMAD.HI imm,IMAD reg,LOP3)", w);
        run<1, 2, K_HILO_SPLIT>("DFMA + 2 (IMAD.HI imm,IMAD reg,LOP3)", w);
        run<0, 1, K_HILO_SPLIT>("1 (IMAD.HI imm,IMAD reg,LOP3)", w);
        run<4, 1, K_IMADW>("4 DFMA + 1 (IMAD.WIDE,LOP3)", w);
        run<4, 1, K_HILO_SPLIT>("4 DFMA + 1 (IMAD.HI imm,IMAD reg,LOP3)", w);
        run<4, 0, K_FFMA>("4 DFMA", w);
        run<4, 1, K_F2F>("4 DFMA + 1 (F2F.F64.F32,LOP3,FADD)", w);
        run<8, 1, K_F2F>("8 DFMA + 1 (F2F.F64.F32,LOP3,FADD)", w);
        run<4, 1, K_WIDEN_INT>("4 DFMA + 1 (int widen ~6 ALU,FADD)", w);
        run<8, 1, K_WIDEN_INT>("8 DFMA + 1 (int widen ~6 ALU,FADD)", w);
        run<8, 0, K_FFMA>("8 DFMA", w);
        run<8, 1, K_MUFU>("8 DFMA + 1 MUFU", w);
        run<8, 2, K_MUFU>("8 DFMA + 2 MUFU", w);
        run<8, 1, K_I2F>("8 DFMA + 1 (I2FP,FADD,IADD)", w);
        run<4, 1, K_MUFU>("4 DFMA + 1 MUFU", w);
        run<4, 2, K_FFMA>("4 DFMA + 2 FFMA", w);
        run<4, 4, K_FFMA>("4 DFMA + 4 FFMA", w);
        run<4, 2, K_PRMT>("4 DFMA + 2 PRMT", w);
        run<4, 4, K_PRMT>("4 DFMA + 4 PRMT", w);
        run<1, 1, K_HI>("DFMA + 1 IMAD.HI", w);
        run<1, 2, K_HI>("DFMA + 2 IMAD.HI", w);
        run<0, 2, K_HI>("2 IMAD.HI", w);
        run<0, 2, K_PRMT>("2 PRMT", w);
        run<1, 2, K_PRMT>("DFMA + 2 PRMT", w);
    }
    return 0;
}
