/* TEST INFRASTRUCTURE ONLY (oracle/): minimal CBLAS declarations so that the
 * reference's proj3/dnn_openblas.c and dnn_avx.c compile unmodified, from where
 * they lie under /root/reference, against the OpenBLAS that ships inside scipy
 * (its symbols carry a scipy_ prefix).  Only the three routines the reference
 * calls are declared (call sites: proj3/dnn_openblas.c:12, :28, :184).
 * Nothing under the product package includes this file. */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H

#define cblas_scopy scipy_cblas_scopy
#define cblas_saxpy scipy_cblas_saxpy
#define cblas_sgemm scipy_cblas_sgemm

typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;

void cblas_scopy(int n, const float *x, int incx, float *y, int incy);
void cblas_saxpy(int n, float alpha, const float *x, int incx, float *y, int incy);
void cblas_sgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb,
                 int m, int n, int k, float alpha, const float *a, int lda,
                 const float *b, int ldb, float beta, float *c, int ldc);

#endif
