// mini_lapack.cpp — fallback providers of the three BLAS/LAPACK entry points the
// reference links (lib/optimisation.cpp:98,134-135,142; lib/simulation.cpp:383;
// lib/stochastic_processes.cpp:23), used ONLY when no bundled OpenBLAS can be
// linked (make BLAS=mini).  Textbook algorithms: unblocked right-looking LU with
// row partial pivoting (what dgesv specifies), scaled 2-norm, row-major gemv.
// TEST INFRASTRUCTURE ONLY.
#include <cmath>
#include <vector>
#include "shim/lapacke.h"
#include "shim/cblas.h"

extern "C" lapack_int LAPACKE_dgesv_work(int layout, lapack_int n, lapack_int nrhs, double* a,
                                         lapack_int lda, lapack_int* ipiv, double* b, lapack_int ldb) {
    if (layout != LAPACK_ROW_MAJOR) return -1;
    for (lapack_int c = 0; c < n; ++c) {
        lapack_int p = c;
        double best = std::fabs(a[c * lda + c]);
        for (lapack_int r = c + 1; r < n; ++r)
            if (std::fabs(a[r * lda + c]) > best) { best = std::fabs(a[r * lda + c]); p = r; }
        ipiv[c] = p + 1;
        if (a[p * lda + c] == 0.0) return c + 1;
        if (p != c) {
            for (lapack_int k = 0; k < n; ++k) std::swap(a[c * lda + k], a[p * lda + k]);
            for (lapack_int k = 0; k < nrhs; ++k) std::swap(b[c * ldb + k], b[p * ldb + k]);
        }
        const double inv = 1.0 / a[c * lda + c];
        for (lapack_int r = c + 1; r < n; ++r) {
            const double l = a[r * lda + c] * inv;
            a[r * lda + c] = l;
            for (lapack_int k = c + 1; k < n; ++k) a[r * lda + k] -= l * a[c * lda + k];
            for (lapack_int k = 0; k < nrhs; ++k) b[r * ldb + k] -= l * b[c * ldb + k];
        }
    }
    for (lapack_int k = 0; k < nrhs; ++k)
        for (lapack_int r = n - 1; r >= 0; --r) {
            double s = b[r * ldb + k];
            for (lapack_int c = r + 1; c < n; ++c) s -= a[r * lda + c] * b[c * ldb + k];
            b[r * ldb + k] = s / a[r * lda + r];
        }
    return 0;
}

extern "C" double cblas_dnrm2(const int n, const double* x, const int incx) {
    double scale = 0.0, ssq = 1.0;
    for (int i = 0; i < n; ++i) {
        const double v = std::fabs(x[i * incx]);
        if (v == 0.0) continue;
        if (scale < v) { ssq = 1.0 + ssq * (scale / v) * (scale / v); scale = v; }
        else ssq += (v / scale) * (v / scale);
    }
    return scale * std::sqrt(ssq);
}

extern "C" void cblas_dgemv(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE trans, const int m,
                            const int n, const double alpha, const double* a, const int lda,
                            const double* x, const int incx, const double beta, double* y,
                            const int incy) {
    const bool t = trans != CblasNoTrans;
    const int rows = t ? n : m, cols = t ? m : n;
    for (int i = 0; i < rows; ++i) {
        double s = 0.0;
        for (int j = 0; j < cols; ++j) s += (t ? a[j * lda + i] : a[i * lda + j]) * x[j * incx];
        y[i * incy] = alpha * s + beta * y[i * incy];
    }
}
