/* Minimal stand-in for the system <cblas.h> (absent from this image): the three
 * entry points the reference links. Test infrastructure only. */
#ifndef MAGPY_B200_SHIM_CBLAS_H
#define MAGPY_B200_SHIM_CBLAS_H
extern "C" {
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
double cblas_dnrm2(const int n, const double* x, const int incx);
void cblas_dgemv(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE trans, const int m,
                 const int n, const double alpha, const double* a, const int lda, const double* x,
                 const int incx, const double beta, double* y, const int incy);
}
#endif
