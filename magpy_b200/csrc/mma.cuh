// mma.cuh — the dipolar field of a CTA's members as a matrix product on the FP64 MMA path (DMMA.8x8x4) of sm_100a:
// shared by the Heun (cluster_mma.cu) and implicit-midpoint (cluster_mma_imid.cu) cluster kernels.  See
// cluster_mma.cu for the data layout (row order, packed symmetric blocks, swizzle).
#pragma once
#include "common.cuh"

namespace mb {

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void group_barrier(const int id, const int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

constexpr int MMA_BLK = 576;   // doubles per 24 x 24 block of D

// Operand addressing of one 24 x 24 block for this thread, as indices into the CTA's shared memory (32-bit: the
// generic-pointer version of this loop cost 8 more registers).  Fragment element (row 8 a + g, column 4 ks + t) of
// the warp's A operand sits at  a_idx + ks * sks + a * sa -/+ dsw  with (sks, sa) = (4, 192) for a block read directly
// — where the swizzle moves the even k-steps up and the odd ones down by dsw = 4 sg doubles — and (96, 8), no
// correction, for a block read transposed.
// DG: the blocks of D are read from global memory (`dg`, read-only path) instead of the CTA's shared memory — clusters
// of 65..128 particles, whose packed matrix (up to 612 KB) does not fit next to the moments.
template <bool DG, int NT = 2>
__device__ __forceinline__ void load_frags(double (&af)[3], double (&bf)[2], const double* __restrict__ sm,
                                           const double* __restrict__ dg, const int a_idx, const bool direct, const int dsw,
                                           const int b_idx, const int ks, const int LD4) {
    const int sks = direct ? 4 : 96, sa = direct ? 192 : 8, d = direct ? dsw : 0;
    const double* ap = (DG ? dg : sm) + a_idx + ks * sks + ((ks & 1) ? -d : d);
#pragma unroll
    for (int a = 0; a < 3; ++a) af[a] = DG ? __ldg(ap + a * sa) : ap[a * sa];
    const double* bq = sm + b_idx + ks * LD4;
    bf[0] = bq[0];
    if (NT > 1) bf[1] = bq[8];
    else bf[1] = 0.0;
}

// acc[a][j][e] = H_a(particle 8 pg + g, member 16 mh + 8 j + 2 t + e).  The operand fragments of the next k-step
// (also across block boundaries) are loaded before the six DMMAs of the current one are issued.  NT = live column
// tiles of this warp (2, or 1 in a CTA of the partial last wave whose second tile holds no member).
// sm = the CTA's shared memory: D blocks from index 0, the moment buffer row of this thread's B fragment at b0.
template <int NT, bool DG>
__device__ __forceinline__ void dipolar_mma(double (&acc)[3][2][2], const double* __restrict__ sm,
                                            const double* __restrict__ dg, const int b0, const int G, const int pg,
                                            const int LD, const int g, const int t) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[a][j][0] = acc[a][j][1] = 0.0;
    const int dsw = ((g >> 1) & 1) << 2, st4 = ((t >> 1) & 1) << 2;
    const int off_direct = 24 * g + t, off_transp = 24 * t + (g ^ st4);
    const int LD4 = 4 * LD, LD24 = 24 * LD;
    auto a_index = [&](const int kg) {
        return kg >= pg ? (pg * G - (pg * (pg - 1)) / 2 + (kg - pg)) * MMA_BLK + off_direct
                        : (kg * G - (kg * (kg - 1)) / 2 + (pg - kg)) * MMA_BLK + off_transp;
    };
    int a_cur = a_index(0), b_cur = b0;
    double af[3], bf[2];
    load_frags<DG, NT>(af, bf, sm, dg, a_cur, 0 >= pg, dsw, b_cur, 0, LD4);
    for (int kg = 0; kg < G; ++kg) {
        const int kn = kg + 1 < G ? kg + 1 : kg;   // the last prefetch re-reads a valid block
        const int a_nxt = a_index(kn), b_nxt = b0 + kn * LD24;
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) {
            double an[3], bn[2];
            if (ks < 5) load_frags<DG, NT>(an, bn, sm, dg, a_cur, kg >= pg, dsw, b_cur, ks + 1, LD4);
            else load_frags<DG, NT>(an, bn, sm, dg, a_nxt, kn >= pg, dsw, b_nxt, 0, LD4);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[a][j][0], acc[a][j][1], af[a], bf[j]);
#pragma unroll
            for (int a = 0; a < 3; ++a) af[a] = an[a];
            bf[0] = bn[0]; bf[1] = bn[1];
        }
        a_cur = a_nxt;
        b_cur = b_nxt;
    }
}

}  // namespace mb
