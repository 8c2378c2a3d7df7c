// dom.cu — discrete-orientation model: the two-state master equation of a uniaxial particle
// in a field along its axis, dp/dt = W(h(t)) p with the Neel-Brown rates of lib/dom.cpp:33-59, integrated by the
// reference's adaptive Cash-Karp RK45 (lib/integrators.cpp:152-251) under the driver of lib/simulation.cpp:660-766
// (tolerance = time_step, first step 0.01 time_step, step capped at end_time / 1000, first-order-hold sampling).
//
// B200 mapping: the reference integrates ONE particle per call (magpy/model.py:211-296); what a GPU adds is the batch —
// a size / anisotropy distribution, one thread per item, each running its own adaptive step sequence (a few hundred
// bytes of state in registers, no shared memory, no inter-thread dependence).  The arithmetic follows the reference
// operation by operation; the unit is compiled with -fmad=false so that, like the reference build (x86-64 without FMA),
// no multiply-add is contracted — the step-size controller amplifies rounding differences into different step
// sequences otherwise.
#include <cstdint>
#include <cuda_runtime.h>
#include "launch.h"

namespace mb {

namespace {

constexpr double DOM_KB = 1.38064852e-23;     // include/constants.hpp:10-12
constexpr double DOM_GYROMAG = 1.76086e11;
constexpr double DOM_PI = 3.14159265358979323846;

struct DomSys {
    double k, v, T, ms, alpha, h, f;
    double sigma, pre0;   // field-independent parts of lib/dom.cpp:44-53, hoisted (same operations, same order)
    int shape;
    unsigned ncomp;
};

// lib/field.cpp:23-79
__device__ double dom_field(const DomSys& s, const double t) {
    switch (s.shape) {
        case 0: return s.h * sin(2 * DOM_PI * s.f * t);
        case 1: return s.h * ((int)(t * s.f * 2) % 2 ? -1 : 1);
        case 3: {
            double field = 0;
            for (unsigned k = 1; k < s.ncomp + 1; k++) field += sin(2 * DOM_PI * (2 * k - 1) * s.f * t) / (2 * k - 1);
            field *= 4 / DOM_PI * s.h;
            return field;
        }
        default: return s.h;
    }
}

// lib/dom.cpp:33-59 then W p (lib/stochastic_processes.cpp:18-25)
__device__ void dom_derivs(double (&d)[2], const double (&p)[2], const double t, const DomSys& s) {
    const double h = dom_field(s, t);
    const double sigma = s.sigma;
    const double e1 = sigma * (1 - h) * (1 - h), e2 = sigma * (1 + h) * (1 + h);
    const double prefactor = s.pre0 / (1 - h * h);   // ((taun sqrt(pi)) / sigma^1.5) / (1 - h^2), left to right as the reference
    const double rate1 = 1.0 / prefactor * (1 - h) * exp(-e1);
    const double rate2 = 1.0 / prefactor * (1 + h) * exp(-e2);
    d[0] = -rate2 * p[0] + rate1 * p[1];
    d[1] = rate2 * p[0] + -rate1 * p[1];
}

// one adaptive step, lib/integrators.cpp:152-251 with the table of include/integrators.hpp:128-160
__device__ void dom_rk45(double (&next)[2], double& h_io, double& t_io, const double (&cur)[2], const DomSys& s,
                         const double tol) {
    constexpr double c11 = 0.2, c21 = 3.0 / 40.0, c22 = 9.0 / 40.0, c31 = 3.0 / 10.0, c32 = -9.0 / 10.0, c33 = 6.0 / 5.0,
                     c41 = -11.0 / 54.0, c42 = 2.5, c43 = -70.0 / 27.0, c44 = 35.0 / 27.0, c51 = 1631.0 / 55296.0,
                     c52 = 175.0 / 512.0, c53 = 575.0 / 13824.0, c54 = 44275.0 / 110592.0, c55 = 253.0 / 4096.0,
                     hc1 = 0.2, hc2 = 0.3, hc3 = 0.6, hc4 = 1.0, hc5 = 7.0 / 8.0,
                     x11 = 37.0 / 378.0, x13 = 250.0 / 621.0, x14 = 125.0 / 594.0, x16 = 512.0 / 1771.0,
                     x21 = 2825.0 / 27648.0, x23 = 18575.0 / 48384.0, x24 = 13525.0 / 55296.0, x25 = 277.0 / 14336.0,
                     x26 = 0.25;
    double k1[2], k2[2], k3[2], k4[2], k5[2], k6[2], tmp[2];
    bool ok = false;
    double err = 0, h = h_io;
    const double t = t_io;
    while (!ok) {
        dom_derivs(k1, cur, t, s);
        for (int i = 0; i < 2; i++) next[i] = k1[i] * h * c11 + cur[i];
        dom_derivs(k2, next, t + h * hc1, s);
        for (int i = 0; i < 2; i++) next[i] = cur[i] + h * (c21 * k1[i] + c22 * k2[i]);
        dom_derivs(k3, next, t + h * hc2, s);
        for (int i = 0; i < 2; i++) next[i] = cur[i] + h * (c31 * k1[i] + c32 * k2[i] + c33 * k3[i]);
        dom_derivs(k4, next, t + h * hc3, s);
        for (int i = 0; i < 2; i++) next[i] = cur[i] + h * (c41 * k1[i] + c42 * k2[i] + c43 * k3[i] + c44 * k4[i]);
        dom_derivs(k5, next, t + h * hc4, s);
        for (int i = 0; i < 2; i++)
            next[i] = cur[i] + h * (c51 * k1[i] + c52 * k2[i] + c53 * k3[i] + c54 * k4[i] + c55 * k5[i]);
        dom_derivs(k6, next, t + h * hc5, s);
        for (int i = 0; i < 2; i++) tmp[i] = cur[i] + h * (x11 * k1[i] + x13 * k3[i] + x14 * k4[i] + x16 * k6[i]);
        for (int i = 0; i < 2; i++)
            next[i] = cur[i] + h * (x21 * k1[i] + x23 * k3[i] + x24 * k4[i] + x25 * k5[i] + x26 * k6[i]);
        err = 0;
        double mag = 0;
        for (int i = 0; i < 2; i++) mag += cur[i] * cur[i];
        mag = sqrt(mag);
        for (int i = 0; i < 2; i++) err += fabs(tmp[i] - next[i]) * fabs(tmp[i] - next[i]);
        err = sqrt(err);
        err /= (2 * tol * (1 + mag));
        if (err < 1.0) ok = true;
        else {
            double hf = 0.84 * pow(err, -0.2);
            hf = fabs(hf) < 0.1 ? 0.1 : hf;
            h *= hf;
        }
    }
    t_io = t + h;
    double hf = err == 0.0 ? 5.0 : 0.84 * pow(err, -0.2);
    hf = hf > 5 ? 5.0 : hf;
    h_io = hf * h;
}

__global__ void dom_kernel(const DomBatch B) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B.n) return;
    DomSys s;
    s.k = B.anisotropy[b]; s.v = B.volume[b]; s.T = B.temperature; s.ms = B.magnetisation; s.alpha = B.alpha;
    const double H_k = 2.0 * s.k / B.mu0 / s.ms;              // magpy/core.pyx:224-225
    s.h = B.field_amplitude / H_k; s.f = B.field_frequency; s.shape = B.field_shape; s.ncomp = B.n_components;
    s.sigma = s.k * s.v / DOM_KB / s.T;
    {
        const double taun = s.v * s.ms * (1 + s.alpha * s.alpha) / 2.0 / DOM_GYROMAG / s.alpha / DOM_KB / s.T;
        s.pre0 = taun * sqrt(DOM_PI) / pow(s.sigma, 1.5);
    }
    double last[2], next[2] = {B.p0[2 * b], B.p0[2 * b + 1]};
    const uint64_t S = B.S;
    const double sampling_time = B.end_time / (S - 1);
    double* mz = B.out_mz + b * S;
    double* fl = B.out_field + b * S;
    mz[0] = next[0] - next[1];
    fl[0] = dom_field(s, 0) * H_k;
    double t = 0, t_last = 0;
    const double max_dt = B.end_time / 1000.0;
    double dt = 0.01 * B.time_step;
    unsigned long long step = 0;
    const double eps = B.time_step;
    for (unsigned int sample = 1; sample < S; sample++) {
        while (t <= sample * sampling_time) {
            last[0] = next[0];
            last[1] = next[1];
            t_last = t;
            step++;
            dom_rk45(next, dt, t, last, s, eps);
            dt = dt > max_dt ? max_dt : dt;
        }
        const double mz_last = last[0] - last[1], mz_next = next[0] - next[1];
        const double t_sample = sample * sampling_time;
        const double beta = (mz_next - mz_last) / (t - t_last);
        mz[sample] = mz_last + beta * (t_sample - t_last);
        fl[sample] = dom_field(s, t_sample) * H_k;
    }
    if (B.out_steps != nullptr) B.out_steps[b] = step;
}

}  // namespace

cudaError_t launch_dom(const DomBatch& B, cudaStream_t s) {
    const unsigned threads = 64;   // divergent adaptive loops: small CTAs spread a batch over more schedulers
    dom_kernel<<<(unsigned)((B.n + threads - 1) / threads), threads, 0, s>>>(B);
    return cudaGetLastError();
}

}  // namespace mb
