// llg_math.cuh — per-particle arithmetic of the reduced-unit stochastic LLG equation.
//
// Everything the integrators need about one macrospin, written once for device code.
// The drift and the diffusion of the reference are the same linear map
//     f(m, g) = -m x g - alpha m x (m x g)
// applied to g = h (drift, lib/llg.cpp:14-29) and g = sigma*w (B(m).w, lib/llg.cpp:92-106),
// so one Heun stage is a single f(m, h dt + sigma sqrt(dt) w).  The implicit scheme
// additionally needs the reference's Jacobian tables *as tabulated* (lib/llg.cpp:68-81,
// 118-158), including the two entries of the diffusion Jacobian that are not the
// analytic derivative — the quasi-Newton iterate sequence depends on them (newton_matrix).
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

// One Heun step of a single macrospin in 37 fp64 instructions (44 with a general easy axis), everything in HALF units.
// With u = g + alpha (m x g) the LLG increment is f(m,g) = -m x u.  Fed with HALF the stage-1 argument, g1/2, the second
// cross product yields the midpoint h = (m + x~)/2 = m - m x u1/2 directly in its FMAs; the predictor is x~ = 2 h - m
// (one FMA per component) and the corrector m' = h - x~ x u2/2 is accumulated in the FMAs of the last cross product.
// The round-1 form (predictor from the full-size u1, then 0.5 x~ and 0.5 m + 0.5 x~) spent three more multiplies.
// Arguments: eh = k dt e / 2, dth = dt / 2 (times the field amplitude in the MP instantiations, whose table holds the
// unit waveform), cwh = sigma sqrt(dt) w / 2 (the Box-Muller scale carries the 1/2), hz0 / hz1 = applied field at t, t + dt.
// On sm_100a an FP64 instruction holds its sub-partition's issue port for 2 cycles, 3 when all three operands are
// distinct registers (scripts/micro/dfma_operands.cu), and the kernel is bound by exactly that:
// what counts is the number of FP64 instructions and of three-register DFMAs, not the flop count.
// heun_single_core: the step proper, from the two z-field constants of its stages t0 = h_app(t) dt/2 + cwh.z and
// t1 = h_app(t+dt) dt/2 + cwh.z (the producer warps of heun_single_split.cu hand exactly these over).
template <bool AXIS_Z>
__device__ __forceinline__ V3 heun_single_core(const V3& m, const V3& e, const V3& eh, const double alpha, const double cx,
                                               const double cy, const double t0, const double t1) {
    // stage 1: g/2 = (h(m,t) dt + sigma sqrt(dt) w) / 2
    V3 g;
    if (AXIS_Z) {  // easy axis = z: h = (k m_z + h_app) z, two fp64 ops instead of seven
        g = V3{cx, cy, fma(m.z, eh.z, t0)};
    } else {
        const double s = dot(m, e);
        g = V3{fma(s, eh.x, cx), fma(s, eh.y, cy), fma(s, eh.z, t0)};
    }
    V3 p = cross(m, g);
    V3 u{fma(alpha, p.x, g.x), fma(alpha, p.y, g.y), fma(alpha, p.z, g.z)};
    const V3 h{fma(-m.y, u.z, fma(m.z, u.y, m.x)), fma(-m.z, u.x, fma(m.x, u.z, m.y)),
               fma(-m.x, u.y, fma(m.y, u.x, m.z))};
    const V3 mt{fma(2.0, h.x, -m.x), fma(2.0, h.y, -m.y), fma(2.0, h.z, -m.z)};
    // stage 2 at (x~, t+dt), same Wiener increment
    if (AXIS_Z) {
        g = V3{cx, cy, fma(mt.z, eh.z, t1)};
    } else {
        const double s = dot(mt, e);
        g = V3{fma(s, eh.x, cx), fma(s, eh.y, cy), fma(s, eh.z, t1)};
    }
    p = cross(mt, g);
    u = V3{fma(alpha, p.x, g.x), fma(alpha, p.y, g.y), fma(alpha, p.z, g.z)};
    return V3{fma(-mt.y, u.z, fma(mt.z, u.y, h.x)), fma(-mt.z, u.x, fma(mt.x, u.z, h.y)),
              fma(-mt.x, u.y, fma(mt.y, u.x, h.z))};
}

template <bool AXIS_Z>
__device__ __forceinline__ V3 heun_single_step(const V3& m, const V3& e, const V3& eh, const double alpha,
                                               const double dth, const V3& cwh, const double hz0, const double hz1) {
    return heun_single_core<AXIS_Z>(m, e, eh, alpha, cwh.x, cwh.y, fma(hz0, dth, cwh.z), fma(hz1, dth, cwh.z));
}

// Solve the 3x3 system A d = b of one particle's quasi-Newton update in registers.  The reference hands its
// block-diagonal 3N x 3N matrix to dgesv (lib/optimisation.cpp:134), i.e. pivoted elimination of each 3x3 block;
// here the block is solved by its adjugate (Cramer's rule): 9 cofactors, the determinant, ONE reciprocal — about
// 40 fp64 operations against ~55 plus three reciprocals and the pivot selects of an in-register elimination
// (measured: +17 % on the single-particle implicit kernel, +22..38 % on the 2..4-particle ones).  The matrices of
// the implicit-midpoint iteration are I + O(|h| + |sigma w|), diagonally dominant at every step size at which the
// scheme is usable, where the adjugate matches pivoted elimination to a few ulp: trajectories stay within 1e-10 of
// the reference and the iteration counts stay identical (tests/test_parity_gpu.py).  det == 0 is reported the way
// dgesv reports a zero pivot (info > 0).
// 1 / x for the determinant of a quasi-Newton block: MUFU.RCP64H seed (20 bits) + two Newton steps, <= 1 ulp for the
// well-scaled determinants (1 + O(dt)) of this iteration, without the exponent range checks and the slow-path call
// of the IEEE division (which sit in the dependent chain of every iteration).
__device__ __forceinline__ double rcp_newton(const double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}

__device__ __forceinline__ bool solve3_adjugate(const double A[9], const double b[3], double d[3]) {
    const double c00 = fma(A[4], A[8], -A[5] * A[7]), c01 = fma(A[5], A[6], -A[3] * A[8]), c02 = fma(A[3], A[7], -A[4] * A[6]);
    const double det = fma(A[0], c00, fma(A[1], c01, A[2] * c02));
    if (det == 0.0) return false;
    const double c10 = fma(A[2], A[7], -A[1] * A[8]), c11 = fma(A[0], A[8], -A[2] * A[6]), c12 = fma(A[1], A[6], -A[0] * A[7]);
    const double c20 = fma(A[1], A[5], -A[2] * A[4]), c21 = fma(A[2], A[3], -A[0] * A[5]), c22 = fma(A[0], A[4], -A[1] * A[3]);
    const double inv = rcp_newton(det);
    // d = adj(A) b / det,  adj(A)[i][j] = cofactor[j][i]
    d[0] = inv * fma(c00, b[0], fma(c10, b[1], c20 * b[2]));
    d[1] = inv * fma(c01, b[0], fma(c11, b[1], c21 * b[2]));
    d[2] = inv * fma(c02, b[0], fma(c12, b[1], c22 * b[2]));
    return true;
}

// Quasi-Newton matrix of one particle:   A = I - (a' + B'.w)/2   with the reference's tables
// (lib/llg.cpp:68-81, 118-158, lib/integrators.cpp:630-636), built from their structure instead of entry
// by entry.  Both tables are the same linear map of a "field" v,
//     L(v) = [v]x - alpha ((m.v) I + m v^T - 2 v m^T),
// a' = L(h) + G HJ with G = da/dh and HJ the 3x3 "field Jacobian" block the reference reads, and
// B'.w = L(sigma w) + C, where C holds the two entries in which the reference's diffusion table differs
// from the analytic derivative (T[0][1][1] and T[2][2][1] use m2).  HJ is always rank one, HJ = u eb^T
// (see quirk_u), and G u = f(m,u), so
//     A = I + L(-(h + sigma w)/2) - f(m,u) eb^T / 2 - C/2          59 fp64 operations instead of 164.
// N = 1: u = k e, eb = e (the true anisotropy block k e e^T).
// zero_u (warp-uniform): the caller knows u = 0 — every cluster of two or more particles whose first easy axis has no x and
// no y component (quirk_u: the block the reference reads then holds only zeros) — so f(m,u) and the rank-one term are
// skipped: 35 fp64 operations instead of 59, the same bits (the skipped FMAs would add +-0).
__device__ __forceinline__ void newton_matrix(double A[9], const V3& m, const double alpha, const V3& h,
                                              const V3& sw /* sigma * w */, const V3& u, const V3& e,
                                              const bool zero_u = false) {
    const V3 vh{-0.5 * (h.x + sw.x), -0.5 * (h.y + sw.y), -0.5 * (h.z + sw.z)};
    const V3 am{alpha * m.x, alpha * m.y, alpha * m.z};
    const V3 tv{2.0 * vh.x, 2.0 * vh.y, 2.0 * vh.z};
    const double base = fma(-alpha, dot(m, vh), 1.0);
    A[0] = fma(am.x, vh.x, base);
    A[4] = fma(am.y, vh.y, base);
    A[8] = fma(am.z, vh.z, base);
    A[1] = fma(-am.x, vh.y, fma(tv.x, am.y, -vh.z));
    A[2] = fma(-am.x, vh.z, fma(tv.x, am.z, vh.y));
    A[3] = fma(-am.y, vh.x, fma(tv.y, am.x, vh.z));
    A[5] = fma(-am.y, vh.z, fma(tv.y, am.z, -vh.x));
    A[6] = fma(-am.z, vh.x, fma(tv.z, am.x, -vh.y));
    A[7] = fma(-am.z, vh.y, fma(tv.z, am.y, vh.x));
    if (!zero_u) {
        const V3 f = llg_f(m, u, alpha);
        const V3 fh{-0.5 * f.x, -0.5 * f.y, -0.5 * f.z};
        A[0] = fma(fh.x, e.x, A[0]); A[1] = fma(fh.x, e.y, A[1]); A[2] = fma(fh.x, e.z, A[2]);
        A[3] = fma(fh.y, e.x, A[3]); A[4] = fma(fh.y, e.y, A[4]); A[5] = fma(fh.y, e.z, A[5]);
        A[6] = fma(fh.z, e.x, A[6]); A[7] = fma(fh.z, e.y, A[7]); A[8] = fma(fh.z, e.z, A[8]);
    }
    // the reference's two non-analytic diffusion-table entries
    A[1] = fma(0.5 * (m.z - m.x), alpha * sw.y, A[1]);
    A[7] = fma(-(m.z - m.y), alpha * sw.z, A[7]);
}

// newton_matrix for a single particle whose easy axis is +z (u = k z, e = z): h = (0, 0, h_z), the rank-one
// term only touches the third column and f(m, k z) = -k (m_y + alpha m_x m_z, -m_x + alpha m_y m_z,
// -alpha (m_x^2 + m_y^2)).  `nhsw` = -sigma w / 2 (constant over the iteration).
__device__ __forceinline__ void newton_matrix_axis_z(double A[9], const V3& m, const double alpha, const double hz,
                                                     const V3& sw, const V3& nhsw, const double kred) {
    const V3 vh{nhsw.x, nhsw.y, fma(-0.5, hz, nhsw.z)};
    const V3 am{alpha * m.x, alpha * m.y, alpha * m.z};
    const V3 tv{2.0 * vh.x, 2.0 * vh.y, 2.0 * vh.z};
    const double base = fma(-alpha, dot(m, vh), 1.0);
    const double kx = kred * m.x, ky = kred * m.y;
    // -f(m, k z) / 2
    const V3 fh{0.5 * fma(am.z, kx, ky), 0.5 * fma(am.z, ky, -kx), -0.5 * fma(am.x, kx, am.y * ky)};
    A[0] = fma(am.x, vh.x, base);
    A[4] = fma(am.y, vh.y, base);
    A[8] = fma(am.z, vh.z, base) + fh.z;
    A[1] = fma(-am.x, vh.y, fma(tv.x, am.y, -vh.z));
    A[2] = fma(-am.x, vh.z, fma(tv.x, am.z, vh.y)) + fh.x;
    A[3] = fma(-am.y, vh.x, fma(tv.y, am.x, vh.z));
    A[5] = fma(-am.y, vh.z, fma(tv.y, am.z, -vh.x)) + fh.y;
    A[6] = fma(-am.z, vh.x, fma(tv.z, am.x, -vh.y));
    A[7] = fma(-am.z, vh.y, fma(tv.z, am.y, vh.x));
    A[1] = fma(0.5 * (m.z - m.x), alpha * sw.y, A[1]);
    A[7] = fma(-(m.z - m.y), alpha * sw.z, A[7]);
}

// Exact Jacobian of the implicit-midpoint residual F(X) = X - x0 - f(X, g(X)) / 2 of one particle,
//     g(X) = dt h(X) + sigma w,  h(X) = k (X.e) e + (applied, dipolar: no X dependence kept),  f(X, g) = -X x u,  u = g + alpha X x g:
//     J = (1 + alpha (X.g) / 2) I + ( C(u) + r e^T - alpha g X^T ) / 2,
//     C(u) c = c x u,   r = dt k (X x e + alpha X x (X x e)).
// This is the opt-in `implicit_newton = exact` mode (hard part 1): Newton's method proper, quadratic
// convergence (3 iterations to 1e-9 instead of the ~20 of the reference's quasi-Newton, whose matrix lacks the dt on
// the field Jacobian, lib/integrators.cpp:633).  It solves the SAME implicit-midpoint equation to a tighter residual,
// so trajectories differ from the reference's truncated iterates at the 1e-9 level per step — validated
// statistically and on closed forms, never used for pathwise parity.  `u` = g + alpha X x g from the caller.
__device__ __forceinline__ void newton_matrix_exact(double A[9], const V3& X, const double alpha, const V3& g, const V3& u,
                                                    const double dtk, const V3& e) {
    const V3 p = cross(X, e);
    const V3 q = cross(X, p);
    const V3 r{0.5 * dtk * fma(alpha, q.x, p.x), 0.5 * dtk * fma(alpha, q.y, p.y), 0.5 * dtk * fma(alpha, q.z, p.z)};
    const double diag = fma(0.5 * alpha, dot(X, g), 1.0);
    const V3 ag{-0.5 * alpha * g.x, -0.5 * alpha * g.y, -0.5 * alpha * g.z};
    const V3 hu{0.5 * u.x, 0.5 * u.y, 0.5 * u.z};
    A[0] = fma(r.x, e.x, fma(ag.x, X.x, diag));
    A[4] = fma(r.y, e.y, fma(ag.y, X.y, diag));
    A[8] = fma(r.z, e.z, fma(ag.z, X.z, diag));
    A[1] = fma(r.x, e.y, fma(ag.x, X.y, hu.z));
    A[2] = fma(r.x, e.z, fma(ag.x, X.z, -hu.y));
    A[3] = fma(r.y, e.x, fma(ag.y, X.x, -hu.z));
    A[5] = fma(r.y, e.z, fma(ag.y, X.z, hu.x));
    A[6] = fma(r.z, e.x, fma(ag.z, X.x, hu.y));
    A[7] = fma(r.z, e.y, fma(ag.z, X.y, -hu.x));
}

// The 3x3 block the reference hands to llg::drift_jacobian for particle p of an N-particle cluster is
// the 9 doubles at flat offsets 3p..3p+8 of the dense row-major (3N)^2 anisotropy Jacobian
// (lib/llg.cpp:387 reads hj+(3*n); lib/field.cpp:159-174 fills the diagonal blocks with k e e^T) — the
// true block only for N = 1.  Its row i is the 3-column segment number g = p + i of the dense matrix:
// dense row r = g / N, block column g % N, non-zero only on the block diagonal (g % N == r / 3), where
// it equals k_b e_b[r % 3] e_b^T with b = r / 3.  For N >= 2, g <= N + 1 means r <= 1 and b = 0: every
// block is u e_0^T with u_i = k_0 e_0[r] when g is 0 or N, else 0.  (N = 1: g = i, u = k_0 e_0.)
__device__ __forceinline__ V3 quirk_u(const unsigned N, const unsigned p, const V3& e0, const double k0) {
    double u[3];
#pragma unroll
    for (unsigned i = 0; i < 3; ++i) {
        const unsigned g = p + i, r = g / N;
        const double comp = (r % 3 == 0) ? e0.x : (r % 3 == 1) ? e0.y : e0.z;
        u[i] = (g % N == r / 3) ? k0 * comp : 0.0;
    }
    return V3{u[0], u[1], u[2]};
}

}  // namespace mb
