/* sllg_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's (owlas/magpy) ensemble stochastic
 * Landau-Lifshitz-Gilbert path, written from the behaviour of the reference
 * sources; every function cites the reference file:line it restates.  It keeps
 * the reference's dense 3N x 3N data layout and literal association order so
 * it can be compared with the compiled reference (oracle/_ref) at ~1 ulp.
 *
 * PARITY PINNED: yes — checked in tests/test_oracle_cpu.py against (i) every
 * known-answer value the reference's own unit tests hold for this path
 * (test/tests.cpp, restated as tests/golden/reference_kat.json), (ii) the
 * reference itself compiled from /root/reference (oracle/_ref/libmagpy_ref.so,
 * when present) and (iii) committed golden trajectories generated from that
 * build (tests/golden npz files, generator tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this file's shared object.  The product (magpy_b200) never does.
 *
 * Third-party arithmetic on the path that is NOT under /root/reference and is
 * restated here from its published algorithm:
 *   - LAPACK dgesv (OpenBLAS, version unpinned by the reference's
 *     environment.yml:6): LU with row partial pivoting + two triangular solves;
 *   - BLAS dnrm2: Euclidean norm;
 *   - libstdc++ std::mt19937_64 and std::normal_distribution<double>
 *     (Marsaglia polar method, second variate cached), pinned by the reference's
 *     test/tests.cpp:346-355.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* include/constants.hpp:10-12 — digit for digit (MU0 is the truncated value). */
#define ORC_KB 1.38064852e-23
#define ORC_MU0 1.25663706e-6
#define ORC_GYROMAG 1.76086e11

enum { ORC_SUCCESS = 0, ORC_MAX_ITER = 1, ORC_LAPACK = 2 }; /* include/optimisation.hpp:22-33 */
enum { ORC_SINE = 0, ORC_SQUARE = 1, ORC_CONSTANT = 2 };    /* include/field.hpp:95-97 */

/* ------------------------------------------------------------------------- *
 * Reference noise stream: std::mt19937_64 + std::normal_distribution<double>
 * (lib/rng.cpp:14-24; libstdc++ bits/random.tcc)                            *
 * ------------------------------------------------------------------------- */
typedef struct {
    uint64_t mt[312];
    int idx;
    int has_saved;
    double saved;
} orc_mtnorm;

static void mt_seed(orc_mtnorm* g, uint64_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 312; ++i)
        g->mt[i] = 6364136223846793005ULL * (g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) + (uint64_t)i;
    g->idx = 312;
    g->has_saved = 0;
    g->saved = 0.0;
}

static uint64_t mt_next(orc_mtnorm* g) {
    if (g->idx >= 312) {
        const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL, A = 0xB5026F5AA96619E9ULL;
        for (int i = 0; i < 312; ++i) {
            uint64_t x = (g->mt[i] & UM) | (g->mt[(i + 1) % 312] & LM);
            g->mt[i] = g->mt[(i + 156) % 312] ^ (x >> 1) ^ ((x & 1ULL) ? A : 0ULL);
        }
        g->idx = 0;
    }
    uint64_t y = g->mt[g->idx++];
    y ^= (y >> 29) & 0x5555555555555555ULL;
    y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
    y ^= (y << 37) & 0xFFF7EEE000000000ULL;
    y ^= (y >> 43);
    return y;
}

/* std::generate_canonical<double,53>(mt19937_64): one 64-bit draw / 2^64, clamped below 1 */
static double mt_canonical(orc_mtnorm* g) {
    double r = (double)mt_next(g) / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

static double mtnorm_get(orc_mtnorm* g, double std) {
    if (g->has_saved) {
        g->has_saved = 0;
        return g->saved * std;
    }
    double x, y, r2;
    do {
        x = 2.0 * mt_canonical(g) - 1.0;
        y = 2.0 * mt_canonical(g) - 1.0;
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0.0);
    const double mult = sqrt(-2.0 * log(r2) / r2);
    g->saved = x * mult;
    g->has_saved = 1;
    return y * mult * std;
}

void orc_rng_normal(uint64_t seed, double std, size_t n, double* out) {
    orc_mtnorm g;
    mt_seed(&g, seed);
    for (size_t i = 0; i < n; ++i) out[i] = mtnorm_get(&g, std);
}

/* ------------------------------------------------------------------------- *
 * Philox4x32-10 (Salmon et al., SC'11; Random123) — the counter-based stream *
 * the product generates in-kernel.  Restated here so the device words can be *
 * checked bit-for-bit.                                                       *
 * ------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* The three unit-variance draws the product uses for (seed, member, particle, step):
 * restatement of magpy_b200/csrc/rng.cuh (philox_gauss3) for the parity tests.
 * counter = (step, member, seed_lo, seed_hi), key = (particle | block<<24, 0xB2005EED).
 * mode 0: fp32 Box-Muller of one Philox block (device uses SFU approximations, so compare
 *         with a ~1e-5 tolerance); mode 1: fp64 Box-Muller from 53-bit uniforms, two blocks;
 * mode 2: packed fp32 Box-Muller — ONE block, counter word 0 = (step-1)>>1, key word 0 =
 *         particle | 2<<24, feeds the two steps 2b+1, 2b+2: the 128 bits are cut into three
 *         (23-bit radius field of which the top 22 bits are used as cell midpoints, 18-bit angle)
 *         pairs -> six draws, the first three for the odd `step` (1-based), the last three
 *         for the even one. */
static float u32_to_float_rz(uint32_t x) { /* cvt.rz.f32.u32 */
    if (x >= (1u << 24)) {
        int sh = 8 - __builtin_clz(x);
        x &= ~((1u << sh) - 1u);
    }
    return (float)x;
}

static void bm_pair_packed(uint32_t r23, uint32_t a18, float* c, float* s) {
    const float u = 2.0f - (1.0f + (float)(r23 | 1u) * 1.1920928955078125e-07f); /* (k + 1/2) 2^-22: cell midpoints */
    const float r = sqrtf(log2f(u) * -1.3862943611198906f);
    const float a = (1.0f + (float)a18 * 3.814697265625e-06f) * 6.283185307179586f; /* 2 pi (1 + k 2^-18) */
    *c = r * cosf(a);
    *s = r * sinf(a);
}

void orc_philox_gauss3(uint64_t seed, uint32_t member, uint32_t particle, uint64_t step, int mode, double out[3]) {
    if (mode == 2) {
        const uint32_t pkey[2] = {particle | (2u << 24), 0xB2005EEDu};
        const uint32_t pctr[4] = {(uint32_t)((step - 1) >> 1), member, (uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t v[4];
        float g[6];
        orc_philox4x32_10(pctr, pkey, v);
        /* field layout of magpy_b200/csrc/rng.cuh (philox_gauss6_f32): each 64-bit half {v1:v0}, {v3:v2} is
         * [23-bit radius field | 18-bit angle | 23 bits] from the least significant bit up */
        const uint64_t h0 = ((uint64_t)v[1] << 32) | v[0], h1 = ((uint64_t)v[3] << 32) | v[2];
        bm_pair_packed(v[0] & 0x7fffffu, (uint32_t)(h0 >> 23) & 0x3ffffu, &g[0], &g[1]);
        bm_pair_packed(v[1] >> 9, (uint32_t)(h1 >> 23) & 0x3ffffu, &g[2], &g[3]);
        bm_pair_packed(v[2] & 0x7fffffu, v[3] >> 14, &g[4], &g[5]);
        const int odd = (int)((step - 1) & 1);
        for (int i = 0; i < 3; ++i) out[i] = (double)g[3 * odd + i];
        return;
    }
    uint32_t key[2] = {particle, 0xB2005EEDu};
    const uint32_t ctr[4] = {(uint32_t)step, member, (uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t w[4];
    orc_philox4x32_10(ctr, key, w);
    if (mode == 0) {
        const float K = -1.3862943611198906f, A = 1.4629180792671596e-9f;
        const float u1 = fmaf(u32_to_float_rz(w[0]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        const float u2 = fmaf(u32_to_float_rz(w[2]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        const float r1 = sqrtf(log2f(u1) * K), r2 = sqrtf(log2f(u2) * K);
        const float a1 = (float)w[1] * A, a2 = (float)w[3] * A;
        out[0] = (double)(r1 * cosf(a1));
        out[1] = (double)(r1 * sinf(a1));
        out[2] = (double)(r2 * cosf(a2));
    } else {
        const double two53 = 1.1102230246251565e-16;
        double u1 = ((double)(((((uint64_t)w[1]) << 32) | w[0]) >> 11) + 0.5) * two53;
        double u2 = ((double)(((((uint64_t)w[3]) << 32) | w[2]) >> 11) + 0.5) * two53;
        double r = sqrt(-2.0 * log(u1));
        out[0] = r * cos(2.0 * M_PI * u2);
        out[1] = r * sin(2.0 * M_PI * u2);
        key[0] = particle | (1u << 24);
        orc_philox4x32_10(ctr, key, w);
        u1 = ((double)(((((uint64_t)w[1]) << 32) | w[0]) >> 11) + 0.5) * two53;
        u2 = ((double)(((((uint64_t)w[3]) << 32) | w[2]) >> 11) + 0.5) * two53;
        r = sqrt(-2.0 * log(u1));
        out[2] = r * cos(2.0 * M_PI * u2);
    }
}

/* ------------------------------------------------------------------------- *
 * Leaf functions                                                             *
 * ------------------------------------------------------------------------- */
/* lib/llg.cpp:14-29 */
void orc_drift(double* d, const double* m, double alpha, const double* h) {
    d[0] = m[2] * h[1] - m[1] * h[2] + alpha * (h[0] * (m[1] * m[1] + m[2] * m[2]) - m[0] * (m[1] * h[1] + m[2] * h[2]));
    d[1] = m[0] * h[2] - m[2] * h[0] + alpha * (h[1] * (m[0] * m[0] + m[2] * m[2]) - m[1] * (m[0] * h[0] + m[2] * h[2]));
    d[2] = m[1] * h[0] - m[0] * h[1] + alpha * (h[2] * (m[0] * m[0] + m[1] * m[1]) - m[2] * (m[0] * h[0] + m[1] * h[1]));
}

/* lib/llg.cpp:92-106 — row-major 3x3, B[i][j] multiplies dW_j */
void orc_diffusion(double* B, const double* m, double sr, double alpha) {
    B[0] = alpha * sr * (m[1] * m[1] + m[2] * m[2]);
    B[1] = sr * (m[2] - alpha * m[0] * m[1]);
    B[2] = -sr * (m[1] + alpha * m[0] * m[2]);
    B[3] = -sr * (m[2] + alpha * m[0] * m[1]);
    B[4] = alpha * sr * (m[0] * m[0] + m[2] * m[2]);
    B[5] = sr * (m[0] - alpha * m[1] * m[2]);
    B[6] = sr * (m[1] - alpha * m[0] * m[2]);
    B[7] = -sr * (m[0] + alpha * m[1] * m[2]);
    B[8] = alpha * sr * (m[0] * m[0] + m[1] * m[1]);
}

/* lib/llg.cpp:68-81 — J[3i+j] = d a_i / d m_j given h and hj (hj read as 9 consecutive doubles) */
void orc_drift_jacobian(double* J, const double* m, double a, const double* h, const double* hj) {
    const double m0 = m[0], m1 = m[1], m2 = m[2];
    J[0] = m2 * hj[3] - m1 * hj[6] + a * (-m1 * h[1] - m2 * h[2] + (m1 * m1 + m2 * m2) * hj[0] - m0 * (m1 * hj[3] + m2 * hj[6]));
    J[1] = -h[2] + m2 * hj[4] - m1 * hj[7] + a * (2 * m1 * h[0] + (m1 * m1 + m2 * m2) * hj[1] - m0 * (h[1] + m1 * hj[4] + m2 * hj[7]));
    J[2] = h[1] + m2 * hj[5] - m1 * hj[8] + a * (2 * m2 * h[0] + (m1 * m1 + m2 * m2) * hj[2] - m0 * (h[2] + m1 * hj[5] + m2 * hj[8]));
    J[3] = h[2] - m2 * hj[0] + m0 * hj[6] + a * (2 * m0 * h[1] + (m0 * m0 + m2 * m2) * hj[3] - m1 * (h[0] + m0 * hj[0] + m2 * hj[6]));
    J[4] = -m2 * hj[1] + m0 * hj[7] + a * (-m0 * h[0] - m2 * h[2] + (m0 * m0 + m2 * m2) * hj[4] - m1 * (m0 * hj[1] + m2 * hj[7]));
    J[5] = -h[0] - m2 * hj[2] + m0 * hj[8] + a * (2 * m2 * h[1] + (m0 * m0 + m2 * m2) * hj[5] - m1 * (h[2] + m0 * hj[2] + m2 * hj[8]));
    J[6] = -h[1] + m1 * hj[0] - m0 * hj[3] + a * (2 * m0 * h[2] + (m0 * m0 + m1 * m1) * hj[6] - m2 * (h[0] + m0 * hj[0] + m1 * hj[3]));
    J[7] = h[0] + m1 * hj[1] - m0 * hj[4] + a * (2 * m1 * h[2] + (m0 * m0 + m1 * m1) * hj[7] - m2 * (h[1] + m0 * hj[1] + m1 * hj[4]));
    J[8] = m1 * hj[2] - m0 * hj[5] + a * (-m0 * h[0] - m1 * h[1] + (m0 * m0 + m1 * m1) * hj[8] - m2 * (m0 * hj[2] + m1 * hj[5]));
}

/* lib/llg.cpp:118-158 — T[9x+3y+z] = "d B_xy / d m_z" exactly as the reference
 * tabulates it.  Two entries are NOT the analytic derivative and are kept as the
 * reference has them: T[4] uses m2 (analytic: m0) and T[25] uses m2 (analytic: m1). */
void orc_diffusion_jacobian(double* T, const double* m, double sr, double alpha) {
    const double as = alpha * sr;
    T[0] = 0;            T[1] = 2 * as * m[1];  T[2] = 2 * as * m[2];
    T[3] = -as * m[1];   T[4] = -as * m[2];     T[5] = sr;
    T[6] = -as * m[2];   T[7] = -sr;            T[8] = -as * m[0];
    T[9] = -as * m[1];   T[10] = -as * m[0];    T[11] = -sr;
    T[12] = 2 * as * m[0]; T[13] = 0;           T[14] = 2 * as * m[2];
    T[15] = sr;          T[16] = -as * m[2];    T[17] = -as * m[1];
    T[18] = -as * m[2];  T[19] = sr;            T[20] = -as * m[0];
    T[21] = -sr;         T[22] = -as * m[2];    T[23] = -as * m[1];
    T[24] = 2 * as * m[0]; T[25] = 2 * as * m[2]; T[26] = 0;
}

/* lib/field.cpp:37-40, 51-54; constant: lib/simulation.cpp:555-559 */
double orc_field_value(int shape, double t, double h, double f) {
    switch (shape) {
        case ORC_SINE: return h * sin(2 * M_PI * f * t);
        case ORC_SQUARE: return h * (((int)(t * f * 2)) % 2 ? -1 : 1);
        default: return h;
    }
}

/* lib/field.cpp:187-225 (prefactor :212-215, pair term :217-225): adds, for every i, the
 * dipolar field of all j != i, j ascending. dists [N][N][3] unit vectors, dist_cubes [N][N]. */
void orc_multi_add_dipolar(double* field, double ms, double k_av, const double* v_red, const double* mag,
                           const double* dists, const double* dist_cubes, int N) {
    const double pre = ORC_MU0 * ms * ms / 8.0 / M_PI / k_av;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            if (j == i) continue;
            const double* mj = mag + 3 * j;
            const double* d = dists + (size_t)i * N * 3 + j * 3;
            const double dotp = mj[0] * d[0] + mj[1] * d[1] + mj[2] * d[2];
            const double t1 = v_red[j] / dist_cubes[i * N + j];
            for (int c = 0; c < 3; ++c) field[3 * i + c] += pre * t1 * (3 * dotp * d[c] - mj[c]);
        }
}

/* ------------------------------------------------------------------------- *
 * Dense linear algebra the reference gets from BLAS/LAPACK                   *
 * ------------------------------------------------------------------------- */
double orc_nrm2(int n, const double* x) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += x[i] * x[i];
    return sqrt(s);
}

/* dgesv semantics on a row-major n x n matrix, one right-hand side
 * (lib/optimisation.cpp:134-135).  Returns 0, or i>0 if U(i,i) is exactly 0. */
int orc_dgesv(int n, double* A, int* ipiv, double* b) {
    int info = 0;
    for (int c = 0; c < n; ++c) {
        int p = c;
        double best = fabs(A[c * n + c]);
        for (int r = c + 1; r < n; ++r)
            if (fabs(A[r * n + c]) > best) { best = fabs(A[r * n + c]); p = r; }
        ipiv[c] = p + 1;
        if (A[p * n + c] == 0.0) { if (!info) info = c + 1; continue; }
        if (p != c) {
            for (int k = 0; k < n; ++k) { double t = A[c * n + k]; A[c * n + k] = A[p * n + k]; A[p * n + k] = t; }
            double t = b[c]; b[c] = b[p]; b[p] = t;
        }
        const double inv = 1.0 / A[c * n + c];
        for (int r = c + 1; r < n; ++r) {
            const double l = A[r * n + c] * inv;
            A[r * n + c] = l;
            for (int k = c + 1; k < n; ++k) A[r * n + k] -= l * A[c * n + k];
        }
    }
    if (info) return info;
    for (int r = 1; r < n; ++r)
        for (int c = 0; c < r; ++c) b[r] -= A[r * n + c] * b[c];
    for (int r = n - 1; r >= 0; --r) {
        for (int c = r + 1; c < n; ++c) b[r] -= A[r * n + c] * b[c];
        b[r] /= A[r * n + r];
    }
    return 0;
}

/* ------------------------------------------------------------------------- *
 * Generic integrators over a caller-supplied SDE (used by the unit tests that *
 * restate the reference's integrator tests, and by the LLG driver below)       *
 * ------------------------------------------------------------------------- */
typedef void (*orc_sde_fn)(double* a, double* B, const double* x, double t, void* ctx);
typedef void (*orc_sde_jac_fn)(double* a, double* B, double* ad, double* Bd, const double* x, double ta,
                               double tb, void* ctx);

/* lib/integrators.cpp:372-405 */
void orc_heun_step(double* next, const double* cur, const double* dW, orc_sde_fn sde, void* ctx, int n,
                   int w, double t, double dt, double* work /* 2n + 2nw */) {
    double *a = work, *at = work + n, *B = work + 2 * n, *Bt = work + 2 * n + n * w;
    sde(a, B, cur, t, ctx);
    for (int i = 0; i < n; ++i) {
        next[i] = cur[i] + dt * a[i];
        for (int j = 0; j < w; ++j) next[i] += B[j + i * w] * dW[j] * sqrt(dt);
    }
    sde(at, Bt, next, t + dt, ctx);
    for (int i = 0; i < n; ++i) {
        next[i] = cur[i] + 0.5 * dt * (at[i] + a[i]);
        for (int j = 0; j < w; ++j) next[i] += 0.5 * dW[j] * sqrt(dt) * (Bt[j + i * w] + B[j + i * w]);
    }
}

/* lib/integrators.cpp:576-651 + lib/optimisation.cpp:81-149.
 * work: dwm[w] a[n] B[nw] ad[nn] Bd[nwn] guess[n] tmp[n] J[nn]; ipiv[n].
 * Returns flag + lapack_err like the reference; *iters_out = quasi-Newton iterations done. */
int orc_implicit_midpoint_step(double* x, const double* x0, const double* dW, orc_sde_jac_fn sde, void* ctx,
                               int n, int w, double t, double dt, double eps, long max_iter, double* work,
                               int* ipiv, long* iters_out) {
    double* dwm = work;
    double* a = dwm + w;
    double* B = a + n;
    double* ad = B + n * w;
    double* Bd = ad + n * n;
    double* guess = Bd + (size_t)n * w * n;
    double* tmp = guess + n;
    double* J = tmp + n;

    const double Ah = sqrt(2 * 1000.0 * fabs(log(dt)));
    for (int i = 0; i < w; ++i) dwm[i] = fmax(-Ah, fmin(Ah, dW[i])) * sqrt(dt);

    sde(a, B, ad, Bd, x0, t, t, ctx);
    for (int i = 0; i < n; ++i) {
        guess[i] = a[i] * dt;
        for (int j = 0; j < w; ++j) guess[i] += B[i * w + j] * dwm[j];
    }
    for (int i = 0; i < n; ++i) guess[i] = (guess[i] + x0[i]) / 2;

    for (int i = 0; i < n; ++i) x[i] = guess[i];
    const double tol = eps * orc_nrm2(n, x);
    double err = 2 * tol;
    long iter = max_iter, done = 0;
    int lapack_err = 0;
    while ((err > tol) && (iter-- > 0)) {
        for (int i = 0; i < n; ++i) tmp[i] = x[i];
        sde(a, B, ad, Bd, tmp, t + dt / 2, t, ctx);
        for (int i = 0; i < n; ++i) {
            x[i] = tmp[i] - 0.5 * a[i] * dt - x0[i];
            for (int j = 0; j < w; ++j) x[i] -= 0.5 * B[i * w + j] * dwm[j];
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                J[i * n + j] = (i == j) - 0.5 * ad[i * n + j];
                for (int k = 0; k < w; ++k) J[i * n + j] -= 0.5 * Bd[(size_t)i * w * n + k * n + j] * dwm[k];
            }
        for (int i = 0; i < n; ++i) x[i] *= -1;
        lapack_err = orc_dgesv(n, J, ipiv, x);
        ++done;
        if (lapack_err != 0) {
            if (iters_out) *iters_out = done;
            for (int i = 0; i < n; ++i) x[i] = 2 * x[i] - x0[i];
            return ORC_LAPACK + lapack_err;
        }
        err = orc_nrm2(n, x);
        for (int i = 0; i < n; ++i) x[i] += tmp[i];
    }
    if (iters_out) *iters_out = done;
    for (int i = 0; i < n; ++i) x[i] = 2 * x[i] - x0[i];
    return iter == -1 ? ORC_MAX_ITER : ORC_SUCCESS;
}

/* lib/integrators.cpp:429-457 and :486-537 (multi-step drivers; states row-major (steps+1) x n) */
void orc_driver_heun(double* states, const double* x0, const double* dW, orc_sde_fn sde, void* ctx,
                     size_t n_steps, int n, int w, double dt) {
    double* work = (double*)malloc(sizeof(double) * (2 * n + 2 * n * w));
    for (int i = 0; i < n; ++i) states[i] = x0[i];
    for (unsigned int s = 0; s < n_steps; ++s)
        orc_heun_step(states + (s + 1) * n, states + s * n, dW + s * w, sde, ctx, n, w, s * dt, dt, work);
    free(work);
}

int orc_driver_implicit(double* states, const double* x0, const double* dW, orc_sde_jac_fn sde, void* ctx,
                        size_t n_steps, int n, int w, double t0, double dt, double eps, long max_iter) {
    size_t wl = (size_t)w + n + (size_t)n * w + (size_t)n * n + (size_t)n * w * n + 2 * n + (size_t)n * n;
    double* work = (double*)malloc(sizeof(double) * wl);
    int* ipiv = (int*)malloc(sizeof(int) * n);
    int worst = 0;
    double t = t0;
    for (int i = 0; i < n; ++i) states[i] = x0[i];
    for (unsigned int s = 0; s < n_steps; ++s) {
        t += dt;
        int e = orc_implicit_midpoint_step(states + (s + 1) * n, states + s * n, dW + s * w, sde, ctx, n, w, t,
                                           dt, eps, max_iter, work, ipiv, NULL);
        if (e) worst = e;
    }
    free(work);
    free(ipiv);
    return worst;
}

/* ------------------------------------------------------------------------- *
 * The N-particle LLG system in reduced units                                 *
 * ------------------------------------------------------------------------- */
typedef struct {
    int N;
    int interactions, field_shape;
    double alpha, Ms, K_av, h0, f_red;
    const double *k_red, *v_red, *sigma, *axes; /* N, N, N, 3N */
    const double *runit, *rcube;                /* N*N*3, N*N */
    double *heff, *hjac;                        /* 3N, (3N)^2 scratch */
} orc_llg;

/* lib/simulation.cpp:271-290 via lib/field.cpp:232-236, 116-132, 81-88, 187-225 */
static void llg_heff(const orc_llg* S, double* h, const double* m, double t) {
    const int N = S->N;
    for (int i = 0; i < 3 * N; ++i) h[i] = 0.0;
    for (int n = 0; n < N; ++n) {
        const double* e = S->axes + 3 * n;
        double dot = m[3 * n] * e[0] + m[3 * n + 1] * e[1] + m[3 * n + 2] * e[2];
        dot *= S->k_red[n];
        h[3 * n + 0] += dot * e[0];
        h[3 * n + 1] += dot * e[1];
        h[3 * n + 2] += dot * e[2];
    }
    const double happ = orc_field_value(S->field_shape, t, S->h0, S->f_red);
    for (int n = 0; n < N; ++n) h[3 * n + 2] += happ;
    if (S->interactions)
        orc_multi_add_dipolar(h, S->Ms, S->K_av, S->v_red, m, S->runit, S->rcube, N);
}

/* lib/llg.cpp:332-348 (drift :257-266, dense diffusion :296-324) */
static void llg_sde(double* a, double* B, const double* m, double t, void* ctx) {
    orc_llg* S = (orc_llg*)ctx;
    const int N = S->N, n3 = 3 * N;
    llg_heff(S, S->heff, m, t);
    for (int n = 0; n < N; ++n) orc_drift(a + 3 * n, m + 3 * n, S->alpha, S->heff + 3 * n);
    memset(B, 0, sizeof(double) * n3 * n3);
    for (int n = 0; n < N; ++n) {
        double b[9];
        orc_diffusion(b, m + 3 * n, S->sigma[n], S->alpha);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) B[(3 * n + r) * n3 + 3 * n + c] = b[3 * r + c];
    }
}

/* lib/llg.cpp:453-481: sde + anisotropy-only field Jacobian (lib/field.cpp:159-174,
 * lib/simulation.cpp:292-303) + quasi-Jacobian of the drift (lib/llg.cpp:378-402; the
 * field-Jacobian block is read at flat offset 3n of the dense (3N)^2 array, :387) +
 * diffusion Jacobian on the 3-d block diagonal (lib/llg.cpp:409-427).
 * ad and Bd off-block entries are left untouched (the caller zero-fills them once,
 * lib/simulation.cpp:189-195). */
static void llg_sde_jac(double* a, double* B, double* ad, double* Bd, const double* m, double ta, double tb,
                        void* ctx) {
    (void)tb;
    orc_llg* S = (orc_llg*)ctx;
    const int N = S->N, n3 = 3 * N;
    llg_sde(a, B, m, ta, ctx);
    memset(S->hjac, 0, sizeof(double) * ((size_t)n3 * n3 + 9));
    for (int i = 0; i < n3; ++i)
        for (int j = 0; j < n3; ++j)
            if (i / 3 == j / 3) S->hjac[i * n3 + j] = S->k_red[i / 3] * S->axes[i] * S->axes[j];
    for (int n = 0; n < N; ++n) {
        double J[9], T[27];
        orc_drift_jacobian(J, m + 3 * n, S->alpha, S->heff + 3 * n, S->hjac + 3 * n);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) ad[(3 * n + r) * n3 + 3 * n + c] = J[3 * r + c];
        orc_diffusion_jacobian(T, m + 3 * n, S->sigma[n], S->alpha);
        for (int x = 0; x < 3; ++x)
            for (int y = 0; y < 3; ++y)
                for (int z = 0; z < 3; ++z)
                    Bd[(size_t)(3 * n + x) * n3 * n3 + (3 * n + y) * n3 + 3 * n + z] = T[9 * x + 3 * y + z];
    }
}

/* ------------------------------------------------------------------------- *
 * SI -> reduced units (lib/simulation.cpp:498-549, 577-583)                   *
 * out_scalars: [0]=V_av [1]=K_av [2]=H_k [3]=time_factor [4]=dt_red           *
 *              [5]=T_red [6]=h0 [7]=f_red [8]=dipolar prefactor               *
 * ------------------------------------------------------------------------- */
void orc_reduce_units(int N, const double* radius, const double* anisotropy, const double* location,
                      double Ms, double alpha, double T, double dt, double t_end, double H0, double f,
                      double* k_red, double* v_red, double* sigma, double* runit, double* rcube,
                      double* out_scalars) {
    double* vol = (double*)malloc(sizeof(double) * N);
    double vsum = 0.0, ksum = 0.0;
    for (int i = 0; i < N; ++i) {
        vol[i] = 4.0 / 3.0 * M_PI * radius[i] * radius[i] * radius[i];
        vsum += vol[i];
        ksum += anisotropy[i];
    }
    const double V_av = vsum / N, K_av = ksum / N;
    for (int i = 0; i < N; ++i) {
        v_red[i] = vol[i] / V_av;
        k_red[i] = anisotropy[i] / K_av;
    }
    const double H_k = 2 * K_av / ORC_MU0 / Ms;
    const double tf = ORC_GYROMAG * ORC_MU0 * H_k / (1 + alpha * alpha);
    for (int i = 0; i < N; ++i)
        sigma[i] = sqrt(alpha * ORC_KB * T / (K_av * vol[i]) / (1 + alpha * alpha));
    const double lscale = pow(V_av, 1. / 3);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double d[3], mag;
            for (int k = 0; k < 3; ++k) d[k] = location[3 * j + k] - location[3 * i + k];
            mag = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            for (int k = 0; k < 3; ++k) runit[((size_t)i * N + j) * 3 + k] = d[k] / mag; /* NaN on the diagonal, never read */
            rcube[i * N + j] = pow(mag / lscale, 3);
        }
    out_scalars[0] = V_av;
    out_scalars[1] = K_av;
    out_scalars[2] = H_k;
    out_scalars[3] = tf;
    out_scalars[4] = dt * tf;
    out_scalars[5] = t_end * tf;
    out_scalars[6] = H0 / H_k;
    out_scalars[7] = f / tf;
    out_scalars[8] = ORC_MU0 * Ms * Ms / 8.0 / M_PI / K_av;
    free(vol);
}

/* Step/sample schedule of lib/simulation.cpp:171-174, 342-355: cum[k] = number of
 * integrator steps executed when sample k (k>=1) is stored; the stored state is the
 * one after cum[k]-1 steps (zero-order hold on the *previous* state, :392-405). */
void orc_schedule(double dt_red, double T_red, size_t S, uint64_t* cum) {
    const double Ts = T_red / (S - 1);
    unsigned int step = 0;
    double t = 0;
    cum[0] = 0;
    for (unsigned int k = 1; k < S; ++k) {
        while (t <= k * Ts) {
            step++;
            t = step * dt_red;
        }
        cum[k] = step;
    }
}

/* ------------------------------------------------------------------------- *
 * simulation::full_dynamics, SI overload + reduced overload                  *
 * (lib/simulation.cpp:476-624 and :132-429)                                   *
 * noise: dW_inject == NULL  -> RngMtNorm(seed,1.0) stream (:586)              *
 *        dW_inject != NULL  -> consumed in order, 3N values per step          *
 *                              (RngArray semantics, lib/rng.cpp:67-94)        *
 * out_m layout [particle][component][sample]; iters[3] = {total, max, count of *
 * the last executed step} quasi-Newton iterations (the last executed step is   *
 * never observed by a sample; the product skips it); returns number of steps    *
 * whose implicit solve failed.                                                  *
 * ------------------------------------------------------------------------- */
long orc_simulate(int N, const double* radius, const double* anisotropy, const double* axis, const double* m0,
                  const double* location, double Ms, double alpha, double T, int renorm, int interactions,
                  int use_implicit, double eps, double dt, double t_end, size_t S, uint64_t seed,
                  int field_shape, double H0, double f, const double* dW_inject, size_t dW_len,
                  double* out_time, double* out_field, double* out_m, long* iters) {
    const int n3 = 3 * N;
    double sc[9];
    double* k_red = (double*)malloc(sizeof(double) * N);
    double* v_red = (double*)malloc(sizeof(double) * N);
    double* sigma = (double*)malloc(sizeof(double) * N);
    double* runit = (double*)malloc(sizeof(double) * N * N * 3);
    double* rcube = (double*)malloc(sizeof(double) * N * N);
    orc_reduce_units(N, radius, anisotropy, location, Ms, alpha, T, dt, t_end, H0, f, k_red, v_red, sigma, runit,
                     rcube, sc);
    const double H_k = sc[2], tf = sc[3], dtr = sc[4], Tr = sc[5];

    orc_llg sys;
    sys.N = N; sys.interactions = interactions; sys.field_shape = field_shape;
    sys.alpha = alpha; sys.Ms = Ms; sys.K_av = sc[1]; sys.h0 = sc[6]; sys.f_red = sc[7];
    sys.k_red = k_red; sys.v_red = v_red; sys.sigma = sigma; sys.axes = axis;
    sys.runit = runit; sys.rcube = rcube;
    sys.heff = (double*)malloc(sizeof(double) * n3);
    sys.hjac = (double*)calloc((size_t)n3 * n3 + 9, sizeof(double));

    size_t wl = (size_t)n3 * 2 + (size_t)n3 * n3 * 3 + (size_t)n3 * n3 * n3 + 3 * n3 + 16;
    double* work = (double*)calloc(wl, sizeof(double)); /* zero-filled like :182-195 */
    int* ipiv = (int*)malloc(sizeof(int) * n3);
    double* p = (double*)malloc(sizeof(double) * n3);
    double* nx = (double*)malloc(sizeof(double) * n3);
    double* w = (double*)malloc(sizeof(double) * n3);
    orc_mtnorm rng;
    mt_seed(&rng, seed);

    const double Ts = Tr / (S - 1);
    for (int i = 0; i < n3; ++i) nx[i] = m0[i];
    out_time[0] = 0;
    out_field[0] = orc_field_value(field_shape, 0, sys.h0, sys.f_red);
    for (int i = 0; i < N; ++i)
        for (int c = 0; c < 3; ++c) out_m[((size_t)i * 3 + c) * S + 0] = m0[3 * i + c];

    unsigned int step = 0;
    double t = 0;
    size_t used = 0;
    long fails = 0, it_total = 0, it_max = 0, it_last = 0;
    for (unsigned int k = 1; k < S; ++k) {
        while (t <= k * Ts) {
            for (int i = 0; i < n3; ++i) p[i] = nx[i];
            step++;
            t = step * dtr;
            for (int i = 0; i < n3; ++i) {
                if (dW_inject) w[i] = used < dW_len ? dW_inject[used++] : 0.0;
                else w[i] = mtnorm_get(&rng, 1.0);
            }
            if (use_implicit) {
                long it = 0;
                int e = orc_implicit_midpoint_step(nx, p, w, llg_sde_jac, &sys, n3, n3, t, dtr, eps, 1000, work,
                                                   ipiv, &it);
                if (e) ++fails;
                it_total += it;
                it_last = it;
                if (it > it_max) it_max = it;
            } else {
                orc_heun_step(nx, p, w, llg_sde, &sys, n3, n3, t, dtr, work);
            }
            if (renorm)
                for (int o = 0; o < n3; o += 3) {
                    const double nrm = orc_nrm2(3, nx + o);
                    for (int c = 0; c < 3; ++c) nx[o + c] = nx[o + c] / nrm;
                }
        }
        out_time[k] = k * Ts;
        out_field[k] = orc_field_value(field_shape, k * Ts, sys.h0, sys.f_red);
        for (int i = 0; i < N; ++i)
            for (int c = 0; c < 3; ++c) out_m[((size_t)i * 3 + c) * S + k] = p[3 * i + c];
    }
    /* back to SI (:610-621) */
    for (size_t s = 0; s < S; ++s) {
        out_time[s] /= tf;
        out_field[s] *= H_k;
    }
    for (size_t j = 0; j < (size_t)n3 * S; ++j) out_m[j] *= Ms;
    if (iters) { iters[0] = it_total; iters[1] = it_max; iters[2] = it_last; }

    free(k_red); free(v_red); free(sigma); free(runit); free(rcube);
    free(sys.heff); free(sys.hjac); free(work); free(ipiv); free(p); free(nx); free(w);
    return fails;
}

/* ------------------------------------------------------------------------- *
 * Discrete-orientation (thermal activation) comparator, lib/dom.cpp:33-59:   *
 * the 2x2 Neel-Brown transition matrix of a uniaxial particle in a reduced   *
 * field h (|h| < 1) along its axis.  Valid for sigma (1-h)^2 >> 1 only       *
 * (lib/dom.cpp:20-22).  Pinned by test/tests.cpp:609-620.                    *
 * ------------------------------------------------------------------------- */
void orc_dom_transition_matrix(double* W, double k, double v, double T, double h, double ms, double alpha) {
    const double sigma = k * v / ORC_KB / T;
    const double taun = v * ms * (1 + alpha * alpha) / 2.0 / ORC_GYROMAG / alpha / ORC_KB / T;
    const double e1 = sigma * (1 - h) * (1 - h), e2 = sigma * (1 + h) * (1 + h);
    const double prefactor = taun * sqrt(M_PI) / pow(sigma, 1.5) / (1 - h * h);
    const double rate1 = 1.0 / prefactor * (1 - h) * exp(-e1);
    const double rate2 = 1.0 / prefactor * (1 + h) * exp(-e2);
    W[0] = -rate2; W[1] = rate1;
    W[2] = rate2;  W[3] = -rate1;
}

/* ------------------------------------------------------------------------- *
 * Discrete-orientation model, time integration:    *
 * the two-state master equation dp/dt = W(h(t)) p of lib/dom.cpp:86-101        *
 * (W p is a 2x2 dgemv, lib/stochastic_processes.cpp:18-25) integrated by the  *
 * reference's adaptive Cash-Karp RK45 (lib/integrators.cpp:152-251, table      *
 * include/integrators.hpp:128-160) under the driver of                         *
 * lib/simulation.cpp:660-766: tolerance = time_step, first step 0.01 time_step,*
 * step capped at end_time/1000 AFTER each step, samples by first-order hold    *
 * between the last two states.  field_shape: 0 sine, 1 square, 2 constant,     *
 * 3 square_fourier (lib/field.cpp:23-79); h_red = H0 / H_k.                    *
 * Returns the number of accepted RK45 steps.                                   *
 * ------------------------------------------------------------------------- */
static double orc_dom_field(int shape, double t, double h, double f, size_t ncomp) {
    switch (shape) {
        case 0: return h * sin(2 * M_PI * f * t);
        case 1: return h * ((int)(t * f * 2) % 2 ? -1 : 1);
        case 3: {
            double field = 0;
            for (unsigned int k = 1; k < ncomp + 1; k++) field += sin(2 * M_PI * (2 * k - 1) * f * t) / (2 * k - 1);
            field *= 4 / M_PI * h;
            return field;
        }
        default: return h;
    }
}

typedef struct {
    double k, v, T, ms, alpha, h, f;
    int shape;
    size_t ncomp;
} orc_dom_sys;

static void orc_dom_derivs(double* d, const double* p, double t, const orc_dom_sys* s) {
    double W[4];
    orc_dom_transition_matrix(W, s->k, s->v, s->T, orc_dom_field(s->shape, t, s->h, s->f, s->ncomp), s->ms, s->alpha);
    d[0] = W[0] * p[0] + W[1] * p[1];
    d[1] = W[2] * p[0] + W[3] * p[1];
}

/* one adaptive step, lib/integrators.cpp:152-251 (n_dims = 2) */
static void orc_rk45(double* next, double* h_ptr, double* t_ptr, const double* cur, const orc_dom_sys* s, double tol) {
    const double c11 = 0.2, c21 = 3.0 / 40.0, c22 = 9.0 / 40.0, c31 = 3.0 / 10.0, c32 = -9.0 / 10.0, c33 = 6.0 / 5.0,
                 c41 = -11.0 / 54.0, c42 = 2.5, c43 = -70.0 / 27.0, c44 = 35.0 / 27.0, c51 = 1631.0 / 55296.0,
                 c52 = 175.0 / 512.0, c53 = 575.0 / 13824.0, c54 = 44275.0 / 110592.0, c55 = 253.0 / 4096.0,
                 hc1 = 0.2, hc2 = 0.3, hc3 = 0.6, hc4 = 1.0, hc5 = 7.0 / 8.0,
                 x11 = 37.0 / 378.0, x13 = 250.0 / 621.0, x14 = 125.0 / 594.0, x16 = 512.0 / 1771.0,
                 x21 = 2825.0 / 27648.0, x23 = 18575.0 / 48384.0, x24 = 13525.0 / 55296.0, x25 = 277.0 / 14336.0, x26 = 0.25;
    double k1[2], k2[2], k3[2], k4[2], k5[2], k6[2], tmp[2];
    int ok = 0;
    double err = 0, h = *h_ptr, t = *t_ptr;
    while (!ok) {
        orc_dom_derivs(k1, cur, t, s);
        for (int i = 0; i < 2; i++) next[i] = k1[i] * h * c11 + cur[i];
        orc_dom_derivs(k2, next, t + h * hc1, s);
        for (int i = 0; i < 2; i++) next[i] = cur[i] + h * (c21 * k1[i] + c22 * k2[i]);
        orc_dom_derivs(k3, next, t + h * hc2, s);
        for (int i = 0; i < 2; i++) next[i] = cur[i] + h * (c31 * k1[i] + c32 * k2[i] + c33 * k3[i]);
        orc_dom_derivs(k4, next, t + h * hc3, s);
        for (int i = 0; i < 2; i++) next[i] = cur[i] + h * (c41 * k1[i] + c42 * k2[i] + c43 * k3[i] + c44 * k4[i]);
        orc_dom_derivs(k5, next, t + h * hc4, s);
        for (int i = 0; i < 2; i++)
            next[i] = cur[i] + h * (c51 * k1[i] + c52 * k2[i] + c53 * k3[i] + c54 * k4[i] + c55 * k5[i]);
        orc_dom_derivs(k6, next, t + h * hc5, s);
        for (int i = 0; i < 2; i++) tmp[i] = cur[i] + h * (x11 * k1[i] + x13 * k3[i] + x14 * k4[i] + x16 * k6[i]);
        for (int i = 0; i < 2; i++)
            next[i] = cur[i] + h * (x21 * k1[i] + x23 * k3[i] + x24 * k4[i] + x25 * k5[i] + x26 * k6[i]);
        err = 0;
        double mag = 0;
        for (int i = 0; i < 2; i++) mag += cur[i] * cur[i];
        mag = pow(mag, 0.5);
        for (int i = 0; i < 2; i++) err += pow(fabs(tmp[i] - next[i]), 2);
        err = pow(err, 0.5);
        err /= (2 * tol * (1 + mag));
        if (err < 1.0) ok = 1;
        else {
            double hf = 0.84 * pow(err, -0.2);
            hf = fabs(hf) < 0.1 ? 0.1 : hf;
            h *= hf;
        }
    }
    *t_ptr = t + h;
    double hf = err == 0.0 ? 5.0 : 0.84 * pow(err, -0.2);
    hf = hf > 5 ? 5.0 : hf;
    *h_ptr = hf * h;
}

uint64_t orc_dom_simulate(double volume, double anisotropy, double temperature, double ms, double alpha, int shape,
                          double h_red, double freq, size_t ncomp, const double* p0, double time_step, double end_time,
                          size_t S, double* out_time, double* out_field, double* out_mz) {
    const orc_dom_sys sys = {anisotropy, volume, temperature, ms, alpha, h_red, freq, shape, ncomp};
    double last[2], next[2] = {p0[0], p0[1]};
    const double sampling_time = end_time / (S - 1);
    out_mz[0] = next[0] - next[1];
    out_time[0] = 0;
    out_field[0] = orc_dom_field(shape, 0, h_red, freq, ncomp);
    double t = 0, t_last = 0;
    const double max_dt = end_time / 1000.0;
    double dt = 0.01 * time_step;
    uint64_t step = 0;
    const double eps = time_step;
    for (unsigned int sample = 1; sample < S; sample++) {
        while (t <= sample * sampling_time) {
            last[0] = next[0];
            last[1] = next[1];
            t_last = t;
            step++;
            orc_rk45(next, &dt, &t, last, &sys, eps);
            dt = dt > max_dt ? max_dt : dt;
        }
        const double mz_last = last[0] - last[1], mz_next = next[0] - next[1];
        const double t_sample = sample * sampling_time;
        const double beta = (mz_next - mz_last) / (t - t_last);
        out_mz[sample] = mz_last + beta * (t_sample - t_last);
        out_time[sample] = t_sample;
        out_field[sample] = orc_dom_field(shape, t_sample, h_red, freq, ncomp);
    }
    return step;
}

/* Number of N(0,1) draws orc_simulate consumes: 3N * (steps executed). */
uint64_t orc_steps_executed(double dt_red, double T_red, size_t S) {
    uint64_t* cum = (uint64_t*)malloc(sizeof(uint64_t) * S);
    orc_schedule(dt_red, T_red, S, cum);
    uint64_t r = cum[S - 1];
    free(cum);
    return r;
}

/* Ensemble over seeds, OpenMP over members (the pragma is ours; the reference's
 * fan-out is joblib, magpy/model.py:202-208).  Same outputs as ref_ensemble.
 * Used as bench.py's cpu_baseline "port" when oracle/_ref is unavailable. */
double orc_ensemble(size_t R, const uint64_t* seeds, int N, const double* radius, const double* anisotropy,
                    const double* axis, size_t axis_stride, const double* m0, size_t m0_stride,
                    const double* location, double Ms, double alpha, double T, int renorm, int interactions,
                    int use_implicit, double eps, double dt, double t_end, size_t S, int field_shape, double H0,
                    double f, double* out_sums, double* out_final) {
    if (out_sums) memset(out_sums, 0, sizeof(double) * S * 4);
#pragma omp parallel
    {
        double* tm = (double*)malloc(sizeof(double) * S);
        double* fl = (double*)malloc(sizeof(double) * S);
        double* mm = (double*)malloc(sizeof(double) * 3 * N * S);
        double* loc = (double*)calloc(S * 4, sizeof(double));
#pragma omp for schedule(dynamic)
        for (long i = 0; i < (long)R; ++i) {
            orc_simulate(N, radius, anisotropy, axis + i * axis_stride, m0 + i * m0_stride, location, Ms, alpha, T,
                         renorm, interactions, use_implicit, eps, dt, t_end, S, seeds[i], field_shape, H0, f,
                         NULL, 0, tm, fl, mm, NULL);
            for (size_t s = 0; s < S; ++s) {
                double sx = 0, sy = 0, sz = 0;
                for (int p = 0; p < N; ++p) {
                    sx += mm[((size_t)p * 3 + 0) * S + s];
                    sy += mm[((size_t)p * 3 + 1) * S + s];
                    sz += mm[((size_t)p * 3 + 2) * S + s];
                }
                loc[4 * s] += sx; loc[4 * s + 1] += sy; loc[4 * s + 2] += sz; loc[4 * s + 3] += sz * sz;
            }
            if (out_final)
                for (int p = 0; p < N; ++p)
                    for (int c = 0; c < 3; ++c)
                        out_final[((size_t)i * N + p) * 3 + c] = mm[((size_t)p * 3 + c) * S + S - 1];
        }
        if (out_sums) {
#pragma omp critical
            for (size_t j = 0; j < S * 4; ++j) out_sums[j] += loc[j];
        }
        free(tm); free(fl); free(mm); free(loc);
    }
    return 0.0;
}
