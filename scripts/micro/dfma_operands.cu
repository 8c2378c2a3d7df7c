// Micro-benchmark: does the rate of DFMA on sm_100a depend on where its operands come from?  The peak kernel and
// issue_mix.cu run x = fma(x, a, b) with a, b shared by all chains (one live register operand per instruction); the Heun
// kernel's DFMAs have three distinct register operands (cross products).  Variants, 16 independent chains per thread:
//   V1  x[i] = fma(x[i], a, b)            a, b uniform
//   V2  x[i] = fma(x[i], p[i], b)         two per-thread register operands
//   V3  x[i] = fma(x[i], p[i], q[i])      hmm: x*p + q — three distinct register operands, accumulator in slot A
//   V4  x[i] = fma(p[i], q[j], x[i])      three distinct, accumulator in slot C, q rotating
//   V5  like V4 but consecutive instructions share p (the operand-reuse cache can serve slot A)
// and each of them with MIX independent LOP3 per DFMA.  Reported: SM cycles per DFMA per sub-partition.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_operands dfma_operands.cu && ./dfma_operands
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int V, int MIX>
__global__ void k(double* out, const double* in, int iters, double a, double b, unsigned ia) {
    double x[16], p[8], q[8];
    unsigned z[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i] = in[threadIdx.x + 32 * (16 + i)]; q[i] = in[threadIdx.x + 32 * (24 + i)]; z[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (V == 1) x[i] = fma(x[i], a, b);
                else if (V == 2) x[i] = fma(x[i], p[i & 7], b);
                else if (V == 3) x[i] = fma(x[i], p[i & 7], q[(i + u) & 7]);
                else if (V == 4) x[i] = fma(p[i & 7], q[(i + u) & 7], x[i]);
                else if (V == 5) x[i] = fma(p[(i >> 2) + 4 * (u & 1)], q[(i + u) & 7], x[i]);
#pragma unroll
                for (int m = 0; m < MIX; ++m) {   // MIX independent 2-input LOP3 per DFMA, 8 chains
                    const int q = (i * MIX + m) & 7;
                    z[q] = (z[q] ^ ia) | (z[q] >> 31 << 30);   // compiles to SHF + LOP3 or one LOP3 — see the SASS count printed by the script
                }
            }
        }
    }
    double s = 0; unsigned w = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) w += z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + w;
}

// dependent-issue latency: ONE chain per thread, one warp per sub-partition; CH > 1 = that many independent chains
template <int CH, bool THREE_REG>
__global__ void chain(double* out, const double* in, int iters, double a, double b) {
    double x[CH], p = in[threadIdx.x + 32 * 16], q = in[threadIdx.x + 32 * 24];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = in[threadIdx.x + 32 * i];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 64 / CH; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) x[i] = THREE_REG ? fma(x[i], p, q) : fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH, bool THREE_REG>
void run_chain(const double* in, double* out) {
    int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        chain<CH, THREE_REG><<<pr.multiProcessorCount, 128>>>(out, in, iters, 0.999999, 1e-7);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    printf("%d dependent chain(s) per warp, one warp per SMSP, %s: %6.2f cycles per DFMA of a chain\n", CH,
           THREE_REG ? "three register operands" : "uniform multiplier / addend", best * 1e-3 * khz * 1e3 / (iters * 64.0 / CH));
}

template <int V, int MIX>
void run(const char* name, int warps_per_smsp, const double* in, double* out) {
    int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    const int threads = 128, blocks = pr.multiProcessorCount * warps_per_smsp, iters = 1024;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<V, MIX><<<blocks, threads>>>(out, in, iters, 0.999999, 1e-7, 2654435761u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double n = (double)iters * 4 * 16;
    printf("%-44s warps/SMSP=%d  cycles per DFMA per SMSP = %6.3f   [+%d x (SHF, LOP3, LOP3) each]\n", name, warps_per_smsp,
           best * 1e-3 * khz * 1e3 / (n * warps_per_smsp), MIX);
}

int main() {
    double *in, *out;
    cudaMalloc(&in, 8 * 32 * 32);
    {   // NON-ZERO operands: multipliers near 1, addends near 1e-3 (an all-zero buffer runs up to 4x faster: the FP64 pipe of
        // sm_100a is data dependent — found the hard way, profiles/r02_dfma_operands_zero_operands.txt)
        double h[32 * 32];
        unsigned long long st = 88172645463325252ull;
        for (int i = 0; i < 32 * 32; ++i) {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            const double u = (double)(st >> 11) * (1.0 / 9007199254740992.0);
            const int row = i / 32;
            h[i] = getenv("DFMA_ZERO") ? 0.0 : (row < 16 ? 0.5 + u : row < 24 ? 0.999 + 1e-4 * u : 1e-3 * (0.5 + u));
        }
        cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    } cudaMalloc(&out, 8 * 148 * 8 * 128);
    run_chain<1, false>(in, out); run_chain<1, true>(in, out); run_chain<2, true>(in, out); run_chain<4, true>(in, out); run_chain<8, true>(in, out);
    for (int w : {1, 2, 4, 8}) {
        run<1, 0>("V1 fma(x, a, b)  uniform a b", w, in, out);
        run<2, 0>("V2 fma(x, p[i], b)", w, in, out);
        run<3, 0>("V3 fma(x, p[i], q[j])", w, in, out);
        run<4, 0>("V4 fma(p[i], q[j], x)", w, in, out);
        run<5, 0>("V5 fma(p[i/4], q[j], x)  shared slot A", w, in, out);
    }
    for (int w : {4, 8}) {
        run<0, 2>("no DFMA, 2 int ops", w, in, out);
        run<1, 1>("V1 + 1", w, in, out);
        run<1, 2>("V1 + 2", w, in, out);
        run<1, 4>("V1 + 4", w, in, out);
        run<4, 1>("V4 + 1", w, in, out);
        run<4, 2>("V4 + 2", w, in, out);
        run<4, 4>("V4 + 4", w, in, out);
    }
    return 0;
}
