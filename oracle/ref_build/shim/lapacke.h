/* Minimal stand-in for the system <lapacke.h>, which this image does not ship.
 * Declares only what the reference's lib/optimisation.cpp uses. The symbol is
 * resolved at link time from the OpenBLAS bundled with opencv-python-headless
 * (or from mini_lapack.cpp when BLAS=mini). Test infrastructure only. */
#ifndef MAGPY_B200_SHIM_LAPACKE_H
#define MAGPY_B200_SHIM_LAPACKE_H
#include <cstddef>
#include <cstdlib>
#include <cstdio>
typedef int lapack_int;
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
extern "C" lapack_int LAPACKE_dgesv_work(int matrix_layout, lapack_int n, lapack_int nrhs,
                                         double* a, lapack_int lda, lapack_int* ipiv,
                                         double* b, lapack_int ldb);
#endif
