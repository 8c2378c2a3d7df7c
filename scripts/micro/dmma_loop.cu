// Micro-benchmark: the matrix-product loop of cluster_mma.cu (dipolar_mma) alone, G = 8, 16 warps per CTA.
// (extracted from ../../magpy_b200/csrc/cluster_mma.cu by the command in profiles/README.md)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void group_barrier(const int id, const int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

constexpr int MMA_BLK = 576;   // doubles per 24 x 24 block of D

// Operand addressing of one 24 x 24 block for this thread.  Fragment element (row 8 a + g, column 4 ks + t) of the
// warp's A operand sits at  a_base + ks * sks + a * sa -/+ dsw  (the swizzle of a directly read block moves the
// even k-steps up and the odd ones down by 4 sg doubles; a transposed block needs no correction).
struct BlockPtr {
    const double* a;
    const double* b;
    int sks, sa, dsw;
};

__device__ __forceinline__ void load_frags(double (&af)[3], double (&bf)[2], const BlockPtr& bp, const int ks, const int LD4) {
    const double* ap = bp.a + ks * bp.sks + ((ks & 1) ? -bp.dsw : bp.dsw);
#pragma unroll
    for (int a = 0; a < 3; ++a) af[a] = ap[a * bp.sa];
    const double* bq = bp.b + ks * LD4;
    bf[0] = bq[0];
    bf[1] = bq[8];
}

// acc[a][j][e] = H_a(particle 8 pg + g, member 16 mh + 8 j + 2 t + e).  The operand fragments of the next k-step
// (also across block boundaries) are loaded before the six DMMAs of the current one are issued.
__device__ __forceinline__ void dipolar_mma(double (&acc)[3][2][2], const double* __restrict__ sm_d,
                                            const double* __restrict__ sm_b /* + 16 mh + g + t * LD */, const int G,
                                            const int pg, const int LD, const int g, const int t) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[a][j][0] = acc[a][j][1] = 0.0;
    const int dsw = ((g >> 1) & 1) << 2, st4 = ((t >> 1) & 1) << 2;
    const int off_direct = 24 * g + t, off_transp = 24 * t + (g ^ st4);
    const int LD4 = 4 * LD, LD24 = 24 * LD;
    auto block_ptr = [&](const int kg) {
        BlockPtr bp;
        bp.b = sm_b + kg * LD24;
        if (kg >= pg) {
            bp.a = sm_d + (pg * G - (pg * (pg - 1)) / 2 + (kg - pg)) * MMA_BLK + off_direct;
            bp.sks = 4; bp.sa = 192; bp.dsw = dsw;
        } else {
            bp.a = sm_d + (kg * G - (kg * (kg - 1)) / 2 + (pg - kg)) * MMA_BLK + off_transp;
            bp.sks = 96; bp.sa = 8; bp.dsw = 0;
        }
        return bp;
    };
    BlockPtr cur = block_ptr(0);
    double af[3], bf[2];
    load_frags(af, bf, cur, 0, LD4);
    for (int kg = 0; kg < G; ++kg) {
        const BlockPtr nxt = block_ptr(kg + 1 < G ? kg + 1 : kg);   // the last prefetch re-reads a valid block
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) {
            double an[3], bn[2];
            if (ks < 5) load_frags(an, bn, cur, ks + 1, LD4);
            else load_frags(an, bn, nxt, 0, LD4);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma884(acc[a][j][0], acc[a][j][1], af[a], bf[j]);
#pragma unroll
            for (int a = 0; a < 3; ++a) af[a] = an[a];
            bf[0] = bn[0]; bf[1] = bn[1];
        }
        cur = nxt;
    }
}


__global__ void __launch_bounds__(512, 1) loop_kernel(double* out, int G, int iters, int nbar) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int n_warps = blockDim.x >> 5, MH = n_warps / G, MB = 16 * MH, LD = MB + 4;
    const int pg = warp % G, mh = warp / G;
    const int n_blk = G * (G + 1) / 2;
    double* sm_d = smem;
    double* sm_m = sm_d + (size_t)n_blk * MMA_BLK;
    for (int q = threadIdx.x; q < n_blk * MMA_BLK + 24 * G * LD; q += blockDim.x) smem[q] = 1e-3 * (q % 89);
    __syncthreads();
    const double* b_m = sm_m + t * LD + 16 * mh + g;
    double s = 0;
    for (int it = 0; it < iters; ++it) {
        double acc[3][2][2];
        dipolar_mma(acc, sm_d, b_m, G, pg, LD, g, t);
#pragma unroll
        for (int a = 0; a < 3; ++a) s += acc[a][0][0] + acc[a][0][1] + acc[a][1][0] + acc[a][1][1];
        if (nbar) asm volatile("bar.sync %0, %1;" ::"r"(1 + mh), "r"(32 * G) : "memory");
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 512);
    for (int nbar = 0; nbar < 2; ++nbar)
    for (int G = 4; G <= 8; G += 4) {
        const int MH = 16 / G, MB = 16 * MH, LD = MB + 4, iters = 2000;
        const size_t smem = sizeof(double) * ((size_t)G * (G + 1) / 2 * 576 + 24 * G * LD);
        cudaFuncSetAttribute(loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        loop_kernel<<<p.multiProcessorCount, 512, smem>>>(out, G, 10, nbar);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        loop_kernel<<<p.multiProcessorCount, 512, smem>>>(out, G, iters, nbar);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double n_mma = (double)iters * G * 36 * 16 * p.multiProcessorCount;
        printf("dipolar_mma loop G=%d (16 warps, %s): %8.3f ms  %6.2f TFLOP/s  %.0f cycles per product per SMSP (ideal %d) (%s)\n", G,
               nbar ? "group barrier after each product" : "no barrier", ms, n_mma * 512 / ms / 1e9,
               ms * 1e-3 * 1.965e9 / iters, 4 * G * 36 * 16, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
