// llg_math.cuh — per-particle arithmetic of the reduced-unit stochastic LLG equation.
//
// Everything the integrators need about one macrospin, written once for device code.
// The drift and the diffusion of the reference are the same linear map
//     f(m, g) = -m x g - alpha m x (m x g)
// applied to g = h (drift, lib/llg.cpp:14-29) and g = sigma*w (B(m).w, lib/llg.cpp:92-106),
// so one Heun stage is a single f(m, h dt + sigma sqrt(dt) w).  The implicit scheme
// additionally needs the reference's Jacobian tables *as tabulated* (lib/llg.cpp:68-81,
// 118-158), including the two entries of the diffusion Jacobian that are not the
// analytic derivative — the quasi-Newton iterate sequence depends on them.
#pragma once
#include <cuda_runtime.h>

namespace mb {

struct V3 {
    double x, y, z;
};

__device__ __forceinline__ V3 cross(const V3& a, const V3& b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// f(m,g) = -m x g - alpha m x (m x g)            [24 flop: 2 cross products + combine]
__device__ __forceinline__ V3 llg_f(const V3& m, const V3& g, const double alpha) {
    const V3 p = cross(m, g);
    const V3 q = cross(m, p);
    return V3{-fma(alpha, q.x, p.x), -fma(alpha, q.y, p.y), -fma(alpha, q.z, p.z)};
}

// d a_i / d m_j as the reference tabulates it (lib/llg.cpp:68-81); hj = 3x3 row-major
// "field Jacobian" block handed in by the caller (which reproduces the block the
// reference actually reads, lib/llg.cpp:387).
__device__ __forceinline__ void drift_jacobian(double J[9], const V3& m, const double a, const V3& h,
                                               const double hj[9]) {
    const double m0 = m.x, m1 = m.y, m2 = m.z, h0 = h.x, h1 = h.y, h2 = h.z;
    const double s12 = m1 * m1 + m2 * m2, s02 = m0 * m0 + m2 * m2, s01 = m0 * m0 + m1 * m1;
    J[0] = m2 * hj[3] - m1 * hj[6] + a * (-m1 * h1 - m2 * h2 + s12 * hj[0] - m0 * (m1 * hj[3] + m2 * hj[6]));
    J[1] = -h2 + m2 * hj[4] - m1 * hj[7] + a * (2 * m1 * h0 + s12 * hj[1] - m0 * (h1 + m1 * hj[4] + m2 * hj[7]));
    J[2] = h1 + m2 * hj[5] - m1 * hj[8] + a * (2 * m2 * h0 + s12 * hj[2] - m0 * (h2 + m1 * hj[5] + m2 * hj[8]));
    J[3] = h2 - m2 * hj[0] + m0 * hj[6] + a * (2 * m0 * h1 + s02 * hj[3] - m1 * (h0 + m0 * hj[0] + m2 * hj[6]));
    J[4] = -m2 * hj[1] + m0 * hj[7] + a * (-m0 * h0 - m2 * h2 + s02 * hj[4] - m1 * (m0 * hj[1] + m2 * hj[7]));
    J[5] = -h0 - m2 * hj[2] + m0 * hj[8] + a * (2 * m2 * h1 + s02 * hj[5] - m1 * (h2 + m0 * hj[2] + m2 * hj[8]));
    J[6] = -h1 + m1 * hj[0] - m0 * hj[3] + a * (2 * m0 * h2 + s01 * hj[6] - m2 * (h0 + m0 * hj[0] + m1 * hj[3]));
    J[7] = h0 + m1 * hj[1] - m0 * hj[4] + a * (2 * m1 * h2 + s01 * hj[7] - m2 * (h1 + m0 * hj[1] + m1 * hj[4]));
    J[8] = m1 * hj[2] - m0 * hj[5] + a * (-m0 * h0 - m1 * h1 + s01 * hj[8] - m2 * (m0 * hj[2] + m1 * hj[5]));
}

// Contribution of the diffusion Jacobian to the quasi-Newton matrix:
//   D[3i+j] = sum_k T[i][k][j] * w_k   with T = lib/llg.cpp:118-158 (index [x][y][z] = 9x+3y+z).
// Entries T[0][1][1] and T[2][2][1] use m2 where the analytic derivative has m0 / m1;
// they are kept as the reference has them.
__device__ __forceinline__ void diffusion_jacobian_dot(double D[9], const V3& m, const double sr,
                                                       const double alpha, const V3& w) {
    const double as = alpha * sr;
    const double m0 = m.x, m1 = m.y, m2 = m.z, w0 = w.x, w1 = w.y, w2 = w.z;
    // i = 0 : T[0][k][j]
    D[0] = /*k0*/ 0.0 + /*k1*/ (-as * m1) * w1 + /*k2*/ (-as * m2) * w2;
    D[1] = (2 * as * m1) * w0 + (-as * m2) * w1 + (-sr) * w2;
    D[2] = (2 * as * m2) * w0 + (sr)*w1 + (-as * m0) * w2;
    // i = 1 : T[1][k][j]
    D[3] = (-as * m1) * w0 + (2 * as * m0) * w1 + (sr)*w2;
    D[4] = (-as * m0) * w0 + 0.0 + (-as * m2) * w2;
    D[5] = (-sr) * w0 + (2 * as * m2) * w1 + (-as * m1) * w2;
    // i = 2 : T[2][k][j]
    D[6] = (-as * m2) * w0 + (-sr) * w1 + (2 * as * m0) * w2;
    D[7] = (sr)*w0 + (-as * m2) * w1 + (2 * as * m2) * w2;
    D[8] = (-as * m0) * w0 + (-as * m1) * w1 + 0.0;
}

// Solve the 3x3 system A d = b in registers: Gaussian elimination with row partial
// pivoting (first largest |entry| in the column), i.e. what dgesv does to the 3x3
// diagonal block the reference's block-diagonal J reduces to (lib/optimisation.cpp:134).
// Returns false when a pivot is exactly zero (dgesv info > 0).
__device__ __forceinline__ bool solve3(double A[9], double b[3], double d[3]) {
    // column 0
    {
        const double a0 = fabs(A[0]), a1 = fabs(A[3]), a2 = fabs(A[6]);
        int p = 0;
        double best = a0;
        if (a1 > best) { best = a1; p = 1; }
        if (a2 > best) { best = a2; p = 2; }
        if (p == 1) {
            double t;
            t = A[0]; A[0] = A[3]; A[3] = t;
            t = A[1]; A[1] = A[4]; A[4] = t;
            t = A[2]; A[2] = A[5]; A[5] = t;
            t = b[0]; b[0] = b[1]; b[1] = t;
        } else if (p == 2) {
            double t;
            t = A[0]; A[0] = A[6]; A[6] = t;
            t = A[1]; A[1] = A[7]; A[7] = t;
            t = A[2]; A[2] = A[8]; A[8] = t;
            t = b[0]; b[0] = b[2]; b[2] = t;
        }
        if (A[0] == 0.0) return false;
        const double inv = 1.0 / A[0];
        const double l1 = A[3] * inv, l2 = A[6] * inv;
        A[4] -= l1 * A[1]; A[5] -= l1 * A[2]; b[1] -= l1 * b[0];
        A[7] -= l2 * A[1]; A[8] -= l2 * A[2]; b[2] -= l2 * b[0];
    }
    // column 1
    {
        if (fabs(A[7]) > fabs(A[4])) {
            double t;
            t = A[4]; A[4] = A[7]; A[7] = t;
            t = A[5]; A[5] = A[8]; A[8] = t;
            t = b[1]; b[1] = b[2]; b[2] = t;
        }
        if (A[4] == 0.0) return false;
        const double l = A[7] * (1.0 / A[4]);
        A[8] -= l * A[5];
        b[2] -= l * b[1];
    }
    if (A[8] == 0.0) return false;
    d[2] = b[2] / A[8];
    d[1] = (b[1] - A[5] * d[2]) / A[4];
    d[0] = (b[0] - A[1] * d[1] - A[2] * d[2]) / A[0];
    return true;
}

}  // namespace mb
