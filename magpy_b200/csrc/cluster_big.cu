// cluster_big.cu — Heun (and, second kernel, implicit midpoint) for clusters beyond the register / shared-memory kernels (N > 128): any cluster size the
// reference accepts (lib/simulation.cpp:182-200 allocates by n_particles).  Nothing is held on chip across a step: the
// moments stay in HBM / L2 in the [3N][R] state layout (member index fastest, so every load of a warp is one coalesced
// line) next to a predictor buffer of the same shape, the static pair table {sqrt(3) r_hat_ij, c_ij} is read through the
// read-only path (warp-uniform addresses), and the Wiener increment of a (particle, step) is regenerated for the
// corrector instead of being stored (counter-based stream).  CTA = 32 members x 16 particle slots; thread (lane, slot)
// walks the particles slot, slot + 16, ...; two CTA barriers per step separate predictor and corrector.  This is the
// scalar arithmetic of cluster.cu (9 fp64 operations per ordered pair) without its capacity limits — a functional path
// (L2-bandwidth bound), not a tuned one: the reference's own dense (3N)^3 work arrays make such clusters impractical
// there (1.5 GB per member at N = 192).
#include "common.cuh"
#include "launch.h"

namespace mb {

constexpr int BIG_SLOTS = 16;

template <int NOISE, bool FIELD_TAB>
__global__ void __launch_bounds__(CL_LANES * BIG_SLOTS) heun_cluster_big_kernel(const __grid_constant__ RunParams P) {
    __shared__ double sm_red[BIG_SLOTS * 3 * CL_LANES];
    const uint32_t N = P.N;
    const int lane = threadIdx.x, slot = threadIdx.y;
    const uint64_t r_raw = (uint64_t)blockIdx.x * CL_LANES + lane, R = P.R;
    const bool live = r_raw < R;
    const uint64_t r = live ? r_raw : R - 1;
    const double alpha = P.alpha, dt = P.dt;
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = member_id(P, r);
    const bool renorm = P.renorm != 0, inter = P.interactions != 0;
    const double* cur = P.state;
    double* pre = P.state_t;

    // effective field on particle p from the moments in `buf` (lib/simulation.cpp:271-290, lib/field.cpp:187-225)
    auto field = [&](const uint32_t p, const V3& mp, const double* buf, const double hz) {
        const uint64_t c0 = 3ull * p;
        const V3 e{__ldg(P.axis + c0 * P.axis_cs + r * P.axis_rs), __ldg(P.axis + (c0 + 1) * P.axis_cs + r * P.axis_rs),
                   __ldg(P.axis + (c0 + 2) * P.axis_cs + r * P.axis_rs)};
        const double s = dot(mp, e) * __ldg(P.k_red + p);
        V3 h{s * e.x, s * e.y, fma(s, e.z, hz)};
        if (inter) {
            const double2* row = reinterpret_cast<const double2*>(P.dip) + (uint64_t)p * N * 2;
#pragma unroll 4
            for (uint32_t jq = 0; jq < N; ++jq) {
                const double* mj = buf + (uint64_t)3 * jq * R + r;
                const double mx = mj[0], my = mj[R], mz = mj[2 * R];
                const double2 t0 = __ldg(row + 2 * jq), t1 = __ldg(row + 2 * jq + 1);   // zero diagonal: no j == p branch
                const double d = mx * t0.x + my * t0.y + mz * t1.x;
                h.x = fma(t1.y, fma(d, t0.x, -mx), h.x);
                h.y = fma(t1.y, fma(d, t0.y, -my), h.y);
                h.z = fma(t1.y, fma(d, t1.x, -mz), h.z);
            }
        }
        return h;
    };
    auto load = [&](const double* buf, const uint32_t p) {
        const double* q = buf + (uint64_t)3 * p * R + r;
        return V3{q[0], q[R], q[2 * R]};
    };

    uint64_t j = P.j0;
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        const uint64_t tgt = (k < P.k1) ? P.target[k] : P.j1;
        for (; j < tgt; ++j) {
            double hz0 = P.h_const, hz1 = P.h_const;
            if (FIELD_TAB) {
                const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (j - P.j0));
                hz0 = h.x; hz1 = h.y;
            }
            // predictor: x~ = m + f(m, h(m) dt + c w)
            for (uint32_t p = slot; p < N; p += BIG_SLOTS) {
                const V3 m = load(cur, p);
                const double c = __ldg(P.sig + p) * P.sqrt_dt;
                const V3 cw = draw_scaled<NOISE>(P, key0, key1, j, p, member, r, c, scale_to_bm(c));
                const V3 h = field(p, m, cur, hz0);
                const V3 g{fma(h.x, dt, cw.x), fma(h.y, dt, cw.y), fma(h.z, dt, cw.z)};
                const V3 f = llg_f(m, g, alpha);
                double* q = pre + (uint64_t)3 * p * R + r;
                if (live) { q[0] = m.x + f.x; q[R] = m.y + f.y; q[2 * R] = m.z + f.z; }
            }
            __syncthreads();
            // corrector: m' = m + (f(m, g) + f(x~, g~)) / 2 = (m + x~) / 2 + f(x~, g~) / 2, same Wiener increment
            for (uint32_t p = slot; p < N; p += BIG_SLOTS) {
                const V3 m = load(cur, p), mt = load(pre, p);
                const double c = __ldg(P.sig + p) * P.sqrt_dt;
                const V3 cw = draw_scaled<NOISE>(P, key0, key1, j, p, member, r, c, scale_to_bm(c));
                const V3 h = field(p, mt, pre, hz1);
                const V3 g{fma(h.x, dt, cw.x), fma(h.y, dt, cw.y), fma(h.z, dt, cw.z)};
                const V3 f = llg_f(mt, g, alpha);
                V3 mn{0.5 * (m.x + mt.x + f.x), 0.5 * (m.y + mt.y + f.y), 0.5 * (m.z + mt.z + f.z)};
                if (renorm) renormalise(mn);
                double* q = P.state + (uint64_t)3 * p * R + r;
                if (live) { q[0] = mn.x; q[R] = mn.y; q[2 * R] = mn.z; }
            }
            __syncthreads();
        }
        if (k < P.k1) {
            double sx = 0, sy = 0, sz = 0;
            for (uint32_t p = slot; p < N; p += BIG_SLOTS) {
                const V3 m = load(cur, p);
                if (P.traj != nullptr && live) {
                    double* o = P.traj + ((uint64_t)k * 3 * N + 3ull * p) * R + r;
                    o[0] = m.x; o[R] = m.y; o[2 * R] = m.z;
                }
                sx += m.x; sy += m.y; sz += m.z;
            }
            if (P.partial != nullptr) {
                double* rr = sm_red + (uint64_t)slot * 3 * CL_LANES + lane;
                rr[0] = sx; rr[CL_LANES] = sy; rr[2 * CL_LANES] = sz;
                __syncthreads();
                if (slot == 0) {
                    double Mx = 0, My = 0, Mz = 0;
                    for (int s2 = 0; s2 < BIG_SLOTS; ++s2) {
                        const double* q2 = sm_red + (uint64_t)s2 * 3 * CL_LANES + lane;
                        Mx += q2[0]; My += q2[CL_LANES]; Mz += q2[2 * CL_LANES];
                    }
                    if (!live) { Mx = 0; My = 0; Mz = 0; }
                    const double v0 = warp_sum(Mx), v1 = warp_sum(My), v2 = warp_sum(Mz), v3 = warp_sum(Mz * Mz);
                    if (lane == 0) {
                        double* o = P.partial + ((uint64_t)(k - P.k0) * gridDim.x + blockIdx.x) * 4;
                        o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
                    }
                }
                __syncthreads();
            }
        }
    }
}

// Implicit midpoint for clusters beyond the shared-memory kernels (N > 128): the reference's quasi-Newton iteration
// (lib/integrators.cpp:576-651, lib/optimisation.cpp:81-149) with the same capacity-free layout.  The midpoint iterate X
// of every particle lives in TWO global buffers [3N][R] (P.state_t, P.state_u): an iteration reads all particles' X from
// one (the dipolar field enters the residual of every particle, lib/simulation.cpp:292-303) and writes X + delta to the
// other; which of the two holds the current iterate is a per-member bit, because members of a CTA converge after
// different numbers of iterations and a finished member must keep its iterate while the CTA goes on (the loop itself is
// CTA-uniform: __syncthreads_or over "some member still iterates").  The quasi-Newton matrix is block diagonal — N
// independent 3x3 solves, as in cluster.cu — and what couples the particles besides the field are the two 3N-wide norms
// (tolerance and error), reduced over the particle slots through shared memory in fixed order.  The Wiener increment of a
// particle is regenerated in every iteration instead of being stored (a Philox block against the N-term dipolar sum).
template <int NOISE, bool FIELD_TAB>
__global__ void __launch_bounds__(CL_LANES * BIG_SLOTS) imid_cluster_big_kernel(const __grid_constant__ RunParams P) {
    __shared__ double sm_red[2 * BIG_SLOTS * 3 * CL_LANES];
    const uint32_t N = P.N;
    const int lane = threadIdx.x, slot = threadIdx.y;
    const uint64_t r_raw = (uint64_t)blockIdx.x * CL_LANES + lane, R = P.R;
    const bool live = r_raw < R;
    const uint64_t r = live ? r_raw : R - 1;
    const double alpha = P.alpha, dt = P.dt, clampA = P.clampA, sqrt_dt = P.sqrt_dt;
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = member_id(P, r);
    const bool renorm = P.renorm != 0, inter = P.interactions != 0, exact = P.newton_exact != 0, zero_u = P.quirk_zero != 0;
    const V3 e0{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
    const double k0 = P.k_red[0];
    NewtonCount nc{0ull, 0ull, 0ull};

    auto axis_of = [&](const uint32_t p) {
        const uint64_t c0 = 3ull * p;
        return V3{__ldg(P.axis + c0 * P.axis_cs + r * P.axis_rs), __ldg(P.axis + (c0 + 1) * P.axis_cs + r * P.axis_rs),
                  __ldg(P.axis + (c0 + 2) * P.axis_cs + r * P.axis_rs)};
    };
    // effective field on particle p (moment mp, easy axis e) from the moments in `buf`
    auto field = [&](const uint32_t p, const V3& mp, const V3& e, const double* buf, const double hz) {
        const double s = dot(mp, e) * __ldg(P.k_red + p);
        V3 h{s * e.x, s * e.y, fma(s, e.z, hz)};
        if (inter) {
            const double2* row = reinterpret_cast<const double2*>(P.dip) + (uint64_t)p * N * 2;
#pragma unroll 4
            for (uint32_t jq = 0; jq < N; ++jq) {
                const double* mj = buf + (uint64_t)3 * jq * R + r;
                const double mx = mj[0], my = mj[R], mz = mj[2 * R];
                const double2 t0 = __ldg(row + 2 * jq), t1 = __ldg(row + 2 * jq + 1);   // zero diagonal: no j == p branch
                const double d = mx * t0.x + my * t0.y + mz * t1.x;
                h.x = fma(t1.y, fma(d, t0.x, -mx), h.x);
                h.y = fma(t1.y, fma(d, t0.y, -my), h.y);
                h.z = fma(t1.y, fma(d, t1.x, -mz), h.z);
            }
        }
        return h;
    };
    auto load = [&](const double* buf, const uint32_t p) {
        const double* q = buf + (uint64_t)3 * p * R + r;
        return V3{q[0], q[R], q[2 * R]};
    };
    auto store = [&](double* buf, const uint32_t p, const V3& v) {
        double* q = buf + (uint64_t)3 * p * R + r;
        if (live) { q[0] = v.x; q[R] = v.y; q[2 * R] = v.z; }
    };
    // sigma_p * clamp(w) * sqrt(dt) (lib/integrators.cpp:598-602)
    auto noise = [&](const uint32_t p, const uint64_t j) {
        const V3 w = draw_noise<NOISE>(P, key0, key1, j, p, member, r);
        const double sr = __ldg(P.sig + p);
        return V3{sr * (fmax(-clampA, fmin(clampA, w.x)) * sqrt_dt), sr * (fmax(-clampA, fmin(clampA, w.y)) * sqrt_dt),
                  sr * (fmax(-clampA, fmin(clampA, w.z)) * sqrt_dt)};
    };
    // sum of one value per (slot, lane) over the slots, in slot order: every thread of a member gets the same bits
    auto slot_sum = [&](const double v, const int which) {
        sm_red[(which * BIG_SLOTS + slot) * CL_LANES + lane] = v;
        __syncthreads();
        double t = 0.0;
        for (int s2 = 0; s2 < BIG_SLOTS; ++s2) t += sm_red[(which * BIG_SLOTS + s2) * CL_LANES + lane];
        return t;
    };

    uint64_t j = P.j0;
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        const uint64_t tgt = (k < P.k1) ? P.target[k] : P.j1;
        for (; j < tgt; ++j) {
            double hz0 = P.h_const, hz1 = P.h_const;
            if (FIELD_TAB) {
                const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (j - P.j0));
                hz0 = h.x; hz1 = h.y;
            }
            // Euler half step as the initial guess of (x0 + x1)/2 (lib/integrators.cpp:605-614)
            int sel = 0;
            double part = 0.0;
            for (uint32_t p = slot; p < N; p += BIG_SLOTS) {
                const V3 m = load(P.state, p), sw = noise(p, j);
                const V3 h = field(p, m, axis_of(p), P.state, hz0);
                const V3 g{fma(h.x, dt, sw.x), fma(h.y, dt, sw.y), fma(h.z, dt, sw.z)};
                const V3 f = llg_f(m, g, alpha);
                const V3 X = exact ? V3{fma(0.5, f.x, m.x), fma(0.5, f.y, m.y), fma(0.5, f.z, m.z)}
                                   : V3{(f.x + m.x) / 2, (f.y + m.y) / 2, (f.z + m.z) / 2};
                store(P.state_t, p, X);
                part += dot(X, X);
            }
            const double nrm = slot_sum(part, 0);   // its barrier also publishes the guess
            const double tol = (P.eps * P.eps) * nrm;     // err > tol is tested on the squares
            double err = 4 * tol;
            int iter = 1000;
            unsigned long long done = 0;
            bool singular = false;
            while (true) {
                bool active = (err > tol) && !singular;
                if (active) { active = iter > 0; --iter; }
                // barrier + vote: also orders the previous iteration's reads of sm_red and writes of the iterate
                if (!__syncthreads_or(active ? 1 : 0)) break;
                const double* src = sel ? P.state_u : P.state_t;
                double* dst = sel ? P.state_t : P.state_u;
                bool ok = true;
                part = 0.0;
                if (active) {
                    for (uint32_t p = slot; p < N; p += BIG_SLOTS) {
                        const V3 m = load(P.state, p), X = load(src, p), sw = noise(p, j), e = axis_of(p);
                        const V3 h = field(p, X, e, src, hz1);
                        const V3 g{fma(h.x, dt, sw.x), fma(h.y, dt, sw.y), fma(h.z, dt, sw.z)};
                        const V3 f = llg_f(X, g, alpha);
                        double b[3] = {-(X.x - m.x - 0.5 * f.x), -(X.y - m.y - 0.5 * f.y), -(X.z - m.z - 0.5 * f.z)};
                        double A[9], d[3];
                        if (exact) {   // opt-in: each particle's exact own Jacobian (llg_math.cuh), dipolar coupling left out as in the reference
                            const V3 pg = cross(X, g);
                            const V3 u{fma(alpha, pg.x, g.x), fma(alpha, pg.y, g.y), fma(alpha, pg.z, g.z)};
                            newton_matrix_exact(A, X, alpha, g, u, dt * __ldg(P.k_red + p), e);
                        } else {
                            newton_matrix(A, X, alpha, h, sw, quirk_u(N, p, e0, k0), e0, zero_u);
                        }
                        if (!solve3_adjugate(A, b, d)) { ok = false; d[0] = d[1] = d[2] = 0.0; }
                        store(dst, p, V3{X.x + d[0], X.y + d[1], X.z + d[2]});
                        part += d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                    }
                }
                sm_red[(BIG_SLOTS + slot) * CL_LANES + lane] = ok ? 0.0 : 1.0;
                const double e2 = slot_sum(part, 0);
                double bad = 0.0;
                for (int s2 = 0; s2 < BIG_SLOTS; ++s2) bad += sm_red[(BIG_SLOTS + s2) * CL_LANES + lane];
                if (active) {
                    ++done;
                    if (bad != 0.0) {
                        singular = true;   // dgesv info > 0: the iteration stops, the iterate stays (as in cluster.cu)
                    } else {
                        err = e2;
                        sel ^= 1;          // the member's iterate is now in the other buffer
                    }
                }
            }
            if (slot == 0) {
                nc.total += done;
                nc.worst = done > nc.worst ? done : nc.worst;
                nc.fails += (singular || iter == -1) ? 1ull : 0ull;
            }
            const double* fin = sel ? P.state_u : P.state_t;
            for (uint32_t p = slot; p < N; p += BIG_SLOTS) {
                const V3 m = load(P.state, p), X = load(fin, p);
                V3 mn{2 * X.x - m.x, 2 * X.y - m.y, 2 * X.z - m.z};
                if (renorm) renormalise(mn);
                store(P.state, p, mn);
            }
            __syncthreads();
        }
        if (k < P.k1) {
            double sx = 0, sy = 0, sz = 0;
            for (uint32_t p = slot; p < N; p += BIG_SLOTS) {
                const V3 m = load(P.state, p);
                if (P.traj != nullptr && live) {
                    double* o = P.traj + ((uint64_t)k * 3 * N + 3ull * p) * R + r;
                    o[0] = m.x; o[R] = m.y; o[2 * R] = m.z;
                }
                sx += m.x; sy += m.y; sz += m.z;
            }
            if (P.partial != nullptr) {
                double* rr = sm_red + (uint64_t)slot * 3 * CL_LANES + lane;
                rr[0] = sx; rr[CL_LANES] = sy; rr[2 * CL_LANES] = sz;
                __syncthreads();
                if (slot == 0) {
                    double Mx = 0, My = 0, Mz = 0;
                    for (int s2 = 0; s2 < BIG_SLOTS; ++s2) {
                        const double* q2 = sm_red + (uint64_t)s2 * 3 * CL_LANES + lane;
                        Mx += q2[0]; My += q2[CL_LANES]; Mz += q2[2 * CL_LANES];
                    }
                    if (!live) { Mx = 0; My = 0; Mz = 0; }
                    const double v0 = warp_sum(Mx), v1 = warp_sum(My), v2 = warp_sum(Mz), v3 = warp_sum(Mz * Mz);
                    if (lane == 0) {
                        double* o = P.partial + ((uint64_t)(k - P.k0) * gridDim.x + blockIdx.x) * 4;
                        o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
                    }
                }
                __syncthreads();
            }
        }
    }
    newton_flush(P, nc, live && slot == 0);
}

cudaError_t launch_heun_cluster_big(int noise, bool tab, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(CL_LANES, BIG_SLOTS);
#define MB_BIG(NOISE)                                                             \
    if (tab) heun_cluster_big_kernel<NOISE, true><<<g, b, 0, s>>>(P);             \
    else heun_cluster_big_kernel<NOISE, false><<<g, b, 0, s>>>(P)
    switch (noise) {
        case NOISE_PHILOX_F32: MB_BIG(NOISE_PHILOX_F32); break;
        case NOISE_PHILOX_F64: MB_BIG(NOISE_PHILOX_F64); break;
        case NOISE_INJECTED: MB_BIG(NOISE_INJECTED); break;
        default: MB_BIG(NOISE_PHILOX_PACKED); break;
    }
#undef MB_BIG
    return cudaGetLastError();
}

cudaError_t launch_imid_cluster_big(int noise, bool tab, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(CL_LANES, BIG_SLOTS);
#define MB_BIGI(NOISE)                                                            \
    if (tab) imid_cluster_big_kernel<NOISE, true><<<g, b, 0, s>>>(P);             \
    else imid_cluster_big_kernel<NOISE, false><<<g, b, 0, s>>>(P)
    switch (noise) {
        case NOISE_PHILOX_F32: MB_BIGI(NOISE_PHILOX_F32); break;
        case NOISE_PHILOX_F64: MB_BIGI(NOISE_PHILOX_F64); break;
        case NOISE_INJECTED: MB_BIGI(NOISE_INJECTED); break;
        default: MB_BIGI(NOISE_PHILOX_PACKED); break;
    }
#undef MB_BIGI
    return cudaGetLastError();
}

}  // namespace mb
