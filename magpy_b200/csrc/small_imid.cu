#include "small.cuh"

namespace mb {

// ---------------------------------------------------------------------------------
// implicit midpoint
// ---------------------------------------------------------------------------------
// EXACT: Newton with each particle's exact own-particle Jacobian (llg_math.cuh: newton_matrix_exact; the dipolar coupling
// between particles stays out of the matrix, as in the reference) instead of the reference's quasi-Newton matrix.
template <int NOISE, bool FIELD_TAB, int N, bool EXACT>
__global__ void __launch_bounds__(SMALL_THREADS) imid_small_kernel(const __grid_constant__ RunParams P) {
    __shared__ double red[(SMALL_THREADS / 32) * 4];
    __shared__ __align__(32) double sd[N * N * 4];
    stage_pair_table<N>(sd, P, 1.0);
    const uint64_t r_raw = (uint64_t)blockIdx.x * SMALL_THREADS + threadIdx.x;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    const double alpha = P.alpha, dt = P.dt, clampA = P.clampA, sqrt_dt = P.sqrt_dt;
    const bool inter = P.interactions != 0, renorm = P.renorm != 0, zero_u = P.quirk_zero != 0;

    V3 m[N], e[N], zero[N];
    double kred[N], sr[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint64_t c0 = 3ull * i;
        m[i] = V3{P.state[c0 * P.R + r], P.state[(c0 + 1) * P.R + r], P.state[(c0 + 2) * P.R + r]};
        e[i] = V3{P.axis[c0 * P.axis_cs + r * P.axis_rs], P.axis[(c0 + 1) * P.axis_cs + r * P.axis_rs],
                  P.axis[(c0 + 2) * P.axis_cs + r * P.axis_rs]};
        kred[i] = P.k_red[i];
        sr[i] = P.sig[i];
        zero[i] = V3{0.0, 0.0, 0.0};
    }
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = member_id(P, r);
    NewtonCount nc{0ull, 0ull, 0ull};

    uint64_t j = P.j0;
    for (uint32_t k = P.k0; k <= P.k1; ++k) {
        const uint64_t tgt = (k < P.k1) ? P.target[k] : P.j1;
        for (; j < tgt; ++j) {
            double hz0 = P.h_const, hz1 = P.h_const;
            if (FIELD_TAB) {
                const double2 h = __ldg(reinterpret_cast<const double2*>(P.field_tab) + (j - P.j0));
                hz0 = h.x; hz1 = h.y;
            }
            V3 wm[N], sw[N], X[N], h[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const V3 w = draw_noise<NOISE>(P, key0, key1, j, (uint32_t)i, member, r);
                wm[i] = V3{fmax(-clampA, fmin(clampA, w.x)) * sqrt_dt, fmax(-clampA, fmin(clampA, w.y)) * sqrt_dt,
                           fmax(-clampA, fmin(clampA, w.z)) * sqrt_dt};
                sw[i] = V3{sr[i] * wm[i].x, sr[i] * wm[i].y, sr[i] * wm[i].z};
            }
            // Euler half step as the initial guess of (x0 + x1)/2 (lib/integrators.cpp:605-614)
            small_fields<N>(h, m, e, kred, hz0, sd, inter, zero);
            double nrm = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const V3 g{fma(h[i].x, dt, sw[i].x), fma(h[i].y, dt, sw[i].y), fma(h[i].z, dt, sw[i].z)};
                const V3 f = llg_f(m[i], g, alpha);
                X[i] = EXACT ? V3{fma(0.5, f.x, m[i].x), fma(0.5, f.y, m[i].y), fma(0.5, f.z, m[i].z)}
                             : V3{(f.x + m[i].x) / 2, (f.y + m[i].y) / 2, (f.z + m[i].z) / 2};
                nrm += dot(X[i], X[i]);
            }
            // err > tol is tested on the squares (no square root in the dependent chain of an iteration)
            const double tol2 = (P.eps * P.eps) * nrm;
            double err2 = 4 * tol2;
            int iter = 1000;
            unsigned long long done = 0;
            bool singular = false;
            while ((err2 > tol2) && (iter-- > 0)) {
                small_fields<N>(h, X, e, kred, hz1, sd, inter, zero);
                V3 dl[N], b[N];
                bool ok = true;
                double e2 = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const V3 g{fma(h[i].x, dt, sw[i].x), fma(h[i].y, dt, sw[i].y), fma(h[i].z, dt, sw[i].z)};
                    const V3 f = llg_f(X[i], g, alpha);
                    double bb[3] = {-(X[i].x - m[i].x - 0.5 * f.x), -(X[i].y - m[i].y - 0.5 * f.y),
                                    -(X[i].z - m[i].z - 0.5 * f.z)};
                    b[i] = V3{bb[0], bb[1], bb[2]};
                    double A[9], d[3];
                    if (EXACT) {
                        const V3 pg = cross(X[i], g);
                        const V3 u{fma(alpha, pg.x, g.x), fma(alpha, pg.y, g.y), fma(alpha, pg.z, g.z)};
                        newton_matrix_exact(A, X[i], alpha, g, u, dt * kred[i], e[i]);
                    } else {
                        newton_matrix(A, X[i], alpha, h[i], sw[i], quirk_u(N, i, e[0], kred[0]), e[0], zero_u);
                    }
                    if (!solve3_adjugate(A, bb, d)) { ok = false; d[0] = d[1] = d[2] = 0.0; }
                    dl[i] = V3{d[0], d[1], d[2]};
                    e2 += d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                }
                ++done;
                if (!ok) {
                    // dgesv info > 0: the reference returns with x_root = -F (lib/optimisation.cpp:134-137)
#pragma unroll
                    for (int i = 0; i < N; ++i) X[i] = b[i];
                    singular = true;
                    break;
                }
                err2 = e2;
#pragma unroll
                for (int i = 0; i < N; ++i) { X[i].x += dl[i].x; X[i].y += dl[i].y; X[i].z += dl[i].z; }
            }
            nc.total += done;
            nc.worst = done > nc.worst ? done : nc.worst;
            nc.fails += (singular || iter == -1) ? 1ull : 0ull;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                m[i] = V3{2 * X[i].x - m[i].x, 2 * X[i].y - m[i].y, 2 * X[i].z - m[i].z};
                if (renorm) renormalise(m[i]);
            }
        }
        if (k < P.k1) sample_outputs<N>(P, m, k, r, live, red);
    }
    if (live) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint64_t c0 = 3ull * i;
            P.state[c0 * P.R + r] = m[i].x; P.state[(c0 + 1) * P.R + r] = m[i].y; P.state[(c0 + 2) * P.R + r] = m[i].z;
        }
    }
    newton_flush(P, nc, live);
}

// ---------------------------------------------------------------------------------
// implicit midpoint, ONE LANE PER PARTICLE (N = 2 or 4 adjacent lanes form a cluster)
//
// For small ensembles the thread-per-cluster kernel is latency bound: 10,000 dimers are 313 warps on 592 warp
// schedulers, each walking the N particles of its cluster one after the other through ~20 quasi-Newton iterations
// per step.  Here the particles of a cluster sit on adjacent lanes and exchange the midpoint iterate with shuffles
// (no shared memory, no barrier), which divides the dependent instruction chain of an iteration by N.  Same
// arithmetic per particle as imid_small_kernel; the two norms are summed by a butterfly over the cluster's lanes,
// which gives every lane the same value (each level adds the same two numbers in either order) and, for N = 2,
// the same value as the sequential sum.  The host picks this kernel when the ensemble cannot fill the GPU with
// one thread per cluster.
// ---------------------------------------------------------------------------------
template <int NOISE, bool FIELD_TAB, int N>
__global__ void __launch_bounds__(SMALL_THREADS) imid_split_kernel(const __grid_constant__ RunParams P) {
    static_assert(N == 2 || N == 4, "lanes per cluster");
    __shared__ double red[(SMALL_THREADS / 32) * 4];
    __shared__ __align__(32) double sd[N * N * 4];
    stage_pair_table<N>(sd, P, 1.0);
    const uint32_t tid = blockIdx.x * SMALL_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, p = lane & (N - 1), base = lane & ~(N - 1);
    const unsigned cmask = ((1u << N) - 1u) << base;   // the lanes of this cluster
    const uint64_t r_raw = tid / N;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    const double alpha = P.alpha, dt = P.dt, clampA = P.clampA, sqrt_dt = P.sqrt_dt;
    const bool inter = P.interactions != 0, renorm = P.renorm != 0, exact = P.newton_exact != 0, zero_u = P.quirk_zero != 0;

    const uint64_t c0 = 3ull * p;
    V3 m{P.state[c0 * P.R + r], P.state[(c0 + 1) * P.R + r], P.state[(c0 + 2) * P.R + r]};
    const V3 e{P.axis[c0 * P.axis_cs + r * P.axis_rs], P.axis[(c0 + 1) * P.axis_cs + r * P.axis_rs],
               P.axis[(c0 + 2) * P.axis_cs + r * P.axis_rs]};
    const V3 e0{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
    const double kred = P.k_red[p], sr = P.sig[p];
    const V3 qu = quirk_u(N, p, e0, P.k_red[0]);
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = member_id(P, r);
    NewtonCount nc{0ull, 0ull, 0ull};

    // effective field of the own particle from the cluster's moments (own = x, the others by shuffle)
    auto field = [&](const V3& x, const double hz) {
        const double s = dot(x, e) * kred;
        V3 h{s * e.x, s * e.y, fma(s, e.z, hz)};
#pragma unroll
        for (int o = 1; o < N; ++o) {   // partner lanes p ^ o: every lane of the cluster executes every shuffle
            const int jp = p ^ o;
            const V3 xj{__shfl_xor_sync(cmask, x.x, o), __shfl_xor_sync(cmask, x.y, o), __shfl_xor_sync(cmask, x.z, o)};
            if (inter) {
                const double4 tt = *reinterpret_cast<const double4*>(sd + (p * N + jp) * 4);
                const double d = xj.x * tt.x + xj.y * tt.y + xj.z * tt.z;
                h.x = fma(tt.w, fma(d, tt.x, -xj.x), h.x);
                h.y = fma(tt.w, fma(d, tt.y, -xj.y), h.y);
                h.z = fma(tt.w, fma(d, tt.z, -xj.z), h.z);
            }
        }
        return h;
    };
    auto cluster_sum = [&](double v) {
#pragma unroll
        for (int o = 1; o < N; o <<= 1) v += __shfl_xor_sync(cmask, v, o);
        return v;
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
            const V3 w = draw_noise<NOISE>(P, key0, key1, j, (uint32_t)p, member, r);
            const V3 wm{fmax(-clampA, fmin(clampA, w.x)) * sqrt_dt, fmax(-clampA, fmin(clampA, w.y)) * sqrt_dt,
                        fmax(-clampA, fmin(clampA, w.z)) * sqrt_dt};
            const V3 sw{sr * wm.x, sr * wm.y, sr * wm.z};
            V3 X;
            {
                const V3 h = field(m, hz0);
                const V3 g{fma(h.x, dt, sw.x), fma(h.y, dt, sw.y), fma(h.z, dt, sw.z)};
                const V3 f = llg_f(m, g, alpha);
                X = exact ? V3{fma(0.5, f.x, m.x), fma(0.5, f.y, m.y), fma(0.5, f.z, m.z)}
                          : V3{(f.x + m.x) / 2, (f.y + m.y) / 2, (f.z + m.z) / 2};
            }
            const double tol2 = (P.eps * P.eps) * cluster_sum(dot(X, X));
            double err2 = 4 * tol2;
            int iter = 1000;
            unsigned long long done = 0;
            bool singular = false;
            while ((err2 > tol2) && (iter-- > 0)) {   // err2, tol2 and iter are the same on every lane of the cluster
                const V3 h = field(X, hz1);
                const V3 g{fma(h.x, dt, sw.x), fma(h.y, dt, sw.y), fma(h.z, dt, sw.z)};
                const V3 f = llg_f(X, g, alpha);
                double bb[3] = {-(X.x - m.x - 0.5 * f.x), -(X.y - m.y - 0.5 * f.y), -(X.z - m.z - 0.5 * f.z)};
                double A[9], d[3];
                if (exact) {
                    const V3 pg = cross(X, g);
                    const V3 u{fma(alpha, pg.x, g.x), fma(alpha, pg.y, g.y), fma(alpha, pg.z, g.z)};
                    newton_matrix_exact(A, X, alpha, g, u, dt * kred, e);
                } else {
                    newton_matrix(A, X, alpha, h, sw, qu, e0, zero_u);
                }
                const bool ok = solve3_adjugate(A, bb, d);
                if (!ok) d[0] = d[1] = d[2] = 0.0;
                const double e2 = cluster_sum(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                const bool all_ok = cluster_sum(ok ? 0.0 : 1.0) == 0.0;
                ++done;
                if (!all_ok) {   // dgesv info > 0: the reference returns with x_root = -F (lib/optimisation.cpp:134-137)
                    X = V3{bb[0], bb[1], bb[2]};
                    singular = true;
                    break;
                }
                err2 = e2;
                X.x += d[0]; X.y += d[1]; X.z += d[2];
            }
            if (p == 0) {
                nc.total += done;
                nc.worst = done > nc.worst ? done : nc.worst;
                nc.fails += (singular || iter == -1) ? 1ull : 0ull;
            }
            m = V3{2 * X.x - m.x, 2 * X.y - m.y, 2 * X.z - m.z};
            if (renorm) renormalise(m);
        }
        if (k < P.k1) {
            if (P.traj != nullptr && live) {
                double* t = P.traj + ((uint64_t)k * 3 * N + 3 * p) * P.R + r;
                t[0] = m.x; t[P.R] = m.y; t[2 * P.R] = m.z;
            }
            if (P.partial != nullptr) {
                const double Mz = cluster_sum(m.z);   // cluster magnetisation, identical on the cluster's lanes
                const double sx = live ? m.x : 0.0, sy = live ? m.y : 0.0, sz = live ? m.z : 0.0;
                cta_partial_sums<SMALL_THREADS / 32>(sx, sy, sz, (live && p == 0) ? Mz * Mz : 0.0, red,
                                                     P.partial + ((uint64_t)(k - P.k0) * gridDim.x + blockIdx.x) * 4);
            }
        }
    }
    if (live) {
        P.state[c0 * P.R + r] = m.x; P.state[(c0 + 1) * P.R + r] = m.y; P.state[(c0 + 2) * P.R + r] = m.z;
    }
    newton_flush(P, nc, live && p == 0);
}

// ---------------------------------------------------------------------------------
// implicit midpoint, ONE WARP PER PARTICLE (CTA = N warps, lane = member; N = 2, 3 or 4)
//
// A warp that has its sub-partition to itself is bound by its own in-order instruction stream, and one thread per
// cluster puts all N particles of a quasi-Newton iteration into that one stream (measured: a lone warp of tetramers takes
// twice as long per iteration as a lone warp of dimers).  For ensembles that leave sub-partitions idle (BASELINE config 2:
// 10,000 dimers = 313 warps on 592 sub-partitions) the particles of a cluster are given to DIFFERENT warps of a CTA, i.e.
// to different sub-partitions: warp p integrates particle p of 32 members, the midpoint iterates are exchanged through a
// double-buffered shared-memory slab and ONE CTA barrier per iteration, and the two cluster-wide norms are summed from the
// same slab in particle order by every warp — the same values, hence the same branch, in all of them (the loop itself is
// uniform per warp through __any_sync: all warps hold the same 32 members).  Same arithmetic per particle and the same
// summation order of the norms as imid_small_kernel: identical iterates and iteration counts.
// ---------------------------------------------------------------------------------
template <int NOISE, bool FIELD_TAB, int N>
__global__ void __launch_bounds__(32 * N) imid_warps_kernel(const __grid_constant__ RunParams P) {
    __shared__ double xs[2][N][3][32];     // midpoint iterates (moments at the start of a step), double buffered
    __shared__ double rs[2][N][2][32];     // per particle: |delta|^2 (or |X0|^2), solve-failed flag
    __shared__ __align__(32) double sd[N * N * 4];
    stage_pair_table<N>(sd, P, 1.0);
    const int p = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t r_raw = (uint64_t)blockIdx.x * 32 + lane;
    const bool live = r_raw < P.R;
    const uint64_t r = live ? r_raw : P.R - 1;
    const double alpha = P.alpha, dt = P.dt, clampA = P.clampA, sqrt_dt = P.sqrt_dt;
    const bool inter = P.interactions != 0, renorm = P.renorm != 0, exact = P.newton_exact != 0, zero_u = P.quirk_zero != 0;

    const uint64_t c0 = 3ull * p;
    V3 m{P.state[c0 * P.R + r], P.state[(c0 + 1) * P.R + r], P.state[(c0 + 2) * P.R + r]};
    const V3 e{P.axis[c0 * P.axis_cs + r * P.axis_rs], P.axis[(c0 + 1) * P.axis_cs + r * P.axis_rs],
               P.axis[(c0 + 2) * P.axis_cs + r * P.axis_rs]};
    const V3 e0{P.axis[r * P.axis_rs], P.axis[P.axis_cs + r * P.axis_rs], P.axis[2 * P.axis_cs + r * P.axis_rs]};
    const double kred = P.k_red[p], sr = P.sig[p];
    const V3 qu = quirk_u(N, p, e0, P.k_red[0]);
    const uint64_t seed = (uint64_t)P.seeds[r];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint32_t member = member_id(P, r);
    NewtonCount nc{0ull, 0ull, 0ull};

    // effective field of the own particle: own moment x, the others from slab `b`
    auto field = [&](const V3& x, const int b, const double hz) {
        const double s = dot(x, e) * kred;
        V3 h{s * e.x, s * e.y, fma(s, e.z, hz)};
        if (inter) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (j == p) continue;
                const V3 xj{xs[b][j][0][lane], xs[b][j][1][lane], xs[b][j][2][lane]};
                const double4 tt = *reinterpret_cast<const double4*>(sd + (p * N + j) * 4);
                const double d = xj.x * tt.x + xj.y * tt.y + xj.z * tt.z;
                h.x = fma(tt.w, fma(d, tt.x, -xj.x), h.x);
                h.y = fma(tt.w, fma(d, tt.y, -xj.y), h.y);
                h.z = fma(tt.w, fma(d, tt.z, -xj.z), h.z);
            }
        }
        return h;
    };
    auto publish = [&](const int b, const V3& x, const double v0, const double v1) {
        xs[b][p][0][lane] = x.x; xs[b][p][1][lane] = x.y; xs[b][p][2][lane] = x.z;
        rs[b][p][0][lane] = v0; rs[b][p][1][lane] = v1;
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
            const V3 w = draw_noise<NOISE>(P, key0, key1, j, (uint32_t)p, member, r);
            const V3 wm{fmax(-clampA, fmin(clampA, w.x)) * sqrt_dt, fmax(-clampA, fmin(clampA, w.y)) * sqrt_dt,
                        fmax(-clampA, fmin(clampA, w.z)) * sqrt_dt};
            const V3 sw{sr * wm.x, sr * wm.y, sr * wm.z};
            publish(0, m, 0.0, 0.0);
            __syncthreads();
            V3 X;
            {
                const V3 h = field(m, 0, hz0);
                const V3 g{fma(h.x, dt, sw.x), fma(h.y, dt, sw.y), fma(h.z, dt, sw.z)};
                const V3 f = llg_f(m, g, alpha);
                X = exact ? V3{fma(0.5, f.x, m.x), fma(0.5, f.y, m.y), fma(0.5, f.z, m.z)}
                          : V3{(f.x + m.x) / 2, (f.y + m.y) / 2, (f.z + m.z) / 2};
            }
            publish(1, X, dot(X, X), 0.0);
            __syncthreads();
            double nrm = 0.0;
#pragma unroll
            for (int q = 0; q < N; ++q) nrm += rs[1][q][0][lane];
            const double tol2 = (P.eps * P.eps) * nrm;
            double err2 = 4 * tol2;
            int iter = 1000, cur = 1;      // slab that holds the current iterate of every particle
            unsigned long long done = 0;
            bool singular = false;
            while (true) {
                bool active = (err2 > tol2) && !singular;
                if (active) { active = iter > 0; --iter; }
                if (!__any_sync(0xffffffffu, active)) break;      // the same 32 members, hence the same vote, in every warp
                V3 Xn = X;
                double e2p = 0.0, badp = 0.0;
                if (active) {
                    const V3 h = field(X, cur, hz1);
                    const V3 g{fma(h.x, dt, sw.x), fma(h.y, dt, sw.y), fma(h.z, dt, sw.z)};
                    const V3 f = llg_f(X, g, alpha);
                    double bb[3] = {-(X.x - m.x - 0.5 * f.x), -(X.y - m.y - 0.5 * f.y), -(X.z - m.z - 0.5 * f.z)};
                    double A[9], d[3];
                    if (exact) {
                        const V3 pg = cross(X, g);
                        const V3 u{fma(alpha, pg.x, g.x), fma(alpha, pg.y, g.y), fma(alpha, pg.z, g.z)};
                        newton_matrix_exact(A, X, alpha, g, u, dt * kred, e);
                    } else {
                        newton_matrix(A, X, alpha, h, sw, qu, e0, zero_u);
                    }
                    if (!solve3_adjugate(A, bb, d)) { badp = 1.0; d[0] = d[1] = d[2] = 0.0; }
                    e2p = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                    Xn = V3{X.x + d[0], X.y + d[1], X.z + d[2]};
                }
                publish(cur ^ 1, Xn, e2p, badp);
                __syncthreads();
                double e2 = 0.0, bad = 0.0;
#pragma unroll
                for (int q = 0; q < N; ++q) { e2 += rs[cur ^ 1][q][0][lane]; bad += rs[cur ^ 1][q][1][lane]; }
                if (active) {
                    ++done;
                    if (bad != 0.0) {      // dgesv info > 0: the iteration stops, the member keeps its iterate (as in cluster.cu)
                        singular = true;
                        // the iterate slab is read again only after the next barrier; the flags are left alone (the other warps
                        // may still be summing them)
                        xs[cur ^ 1][p][0][lane] = X.x; xs[cur ^ 1][p][1][lane] = X.y; xs[cur ^ 1][p][2][lane] = X.z;
                    } else {
                        err2 = e2;
                        X = Xn;
                    }
                }
                cur ^= 1;
            }
            if (p == 0) {
                nc.total += done;
                nc.worst = done > nc.worst ? done : nc.worst;
                nc.fails += (singular || iter == -1) ? 1ull : 0ull;
            }
            m = V3{2 * X.x - m.x, 2 * X.y - m.y, 2 * X.z - m.z};
            if (renorm) renormalise(m);
            __syncthreads();      // every warp has left the loop (no reads of the slabs pending) before the next step publishes
        }
        if (k < P.k1) {
            if (P.traj != nullptr && live) {
                double* t = P.traj + ((uint64_t)k * 3 * N + 3 * p) * P.R + r;
                t[0] = m.x; t[P.R] = m.y; t[2 * P.R] = m.z;
            }
            if (P.partial != nullptr) {
                publish(0, m, 0.0, 0.0);
                __syncthreads();
                if (p == 0) {
                    double Mx = 0, My = 0, Mz = 0;
#pragma unroll
                    for (int q = 0; q < N; ++q) { Mx += xs[0][q][0][lane]; My += xs[0][q][1][lane]; Mz += xs[0][q][2][lane]; }
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
    if (live) {
        P.state[c0 * P.R + r] = m.x; P.state[(c0 + 1) * P.R + r] = m.y; P.state[(c0 + 2) * P.R + r] = m.z;
    }
    newton_flush(P, nc, live && p == 0);
}

template <int NOISE, bool TAB>
static cudaError_t launch_ism(unsigned N, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(SMALL_THREADS);
    if (P.newton_exact) {
        switch (N) {
            case 2: imid_small_kernel<NOISE, TAB, 2, true><<<g, b, 0, s>>>(P); break;
            case 3: imid_small_kernel<NOISE, TAB, 3, true><<<g, b, 0, s>>>(P); break;
            case 4: imid_small_kernel<NOISE, TAB, 4, true><<<g, b, 0, s>>>(P); break;
            default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
    switch (N) {
        case 2: imid_small_kernel<NOISE, TAB, 2, false><<<g, b, 0, s>>>(P); break;
        case 3: imid_small_kernel<NOISE, TAB, 3, false><<<g, b, 0, s>>>(P); break;
        case 4: imid_small_kernel<NOISE, TAB, 4, false><<<g, b, 0, s>>>(P); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_imid_small(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P) {
    MB_NOISE_TAB_DISPATCH(launch_ism, n_particles, grid, s, P)
}

template <int NOISE, bool TAB>
static cudaError_t launch_isp(unsigned N, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid), b(SMALL_THREADS);
    switch (N) {
        case 2: imid_split_kernel<NOISE, TAB, 2><<<g, b, 0, s>>>(P); break;
        case 4: imid_split_kernel<NOISE, TAB, 4><<<g, b, 0, s>>>(P); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// one lane per particle (N = 2 or 4): grid = ceil(R N / SMALL_THREADS)
cudaError_t launch_imid_split(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P) {
    MB_NOISE_TAB_DISPATCH(launch_isp, n_particles, grid, s, P)
}

template <int NOISE, bool TAB>
static cudaError_t launch_iw(unsigned N, unsigned grid, cudaStream_t s, const RunParams& P) {
    const dim3 g(grid);
    switch (N) {
        case 2: imid_warps_kernel<NOISE, TAB, 2><<<g, dim3(64), 0, s>>>(P); break;
        case 3: imid_warps_kernel<NOISE, TAB, 3><<<g, dim3(96), 0, s>>>(P); break;
        case 4: imid_warps_kernel<NOISE, TAB, 4><<<g, dim3(128), 0, s>>>(P); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// one warp per particle (N = 2..4): grid = ceil(R / 32) CTAs of 32 N threads
cudaError_t launch_imid_warps(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P) {
    MB_NOISE_TAB_DISPATCH(launch_iw, n_particles, grid, s, P)
}

}  // namespace mb
