/* TEST INFRASTRUCTURE ONLY (oracle build shim).
 *
 * Minimal stand-in for the system <cblas.h> that the reference tree expects
 * (/root/reference/framework/include/saf_externals.h:144) but this image does
 * not ship.  Only the declarations needed to *compile* the nine reference
 * translation units listed in oracle/Makefile are provided; the only BLAS
 * symbols the convolver path actually *calls* are the level-1 routines
 * cblas_scopy / cblas_ccopy / cblas_saxpy / cblas_sscal, which
 * are resolved at link time against OpenBLAS (if found) or oracle/shim/l1blas.c.
 */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum CBLAS_ORDER     { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113, CblasConjNoTrans = 114 } CBLAS_TRANSPOSE;
typedef enum CBLAS_UPLO      { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum CBLAS_DIAG      { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;
typedef enum CBLAS_SIDE      { CblasLeft = 141, CblasRight = 142 } CBLAS_SIDE;
typedef CBLAS_ORDER CBLAS_LAYOUT;

/* level 1: the ones on the convolver path */
void   cblas_scopy(const int n, const float* x, const int incx, float* y, const int incy);
void   cblas_ccopy(const int n, const void*  x, const int incx, void*  y, const int incy);
void   cblas_saxpy(const int n, const float a, const float* x, const int incx, float* y, const int incy);
void   cblas_sscal(const int n, const float a, float* x, const int incx);

/* remaining prototypes only so that saf_utility_veclib.c / _fft.c / _filters.c / _misc.c
 * compile cleanly; none of them is reachable from saf_matrixConv / saf_multiConv / saf_TVConv */
void   cblas_dcopy(const int n, const double* x, const int incx, double* y, const int incy);
void   cblas_zcopy(const int n, const void* x, const int incx, void* y, const int incy);
void   cblas_daxpy(const int n, const double a, const double* x, const int incx, double* y, const int incy);
void   cblas_caxpy(const int n, const void* a, const void* x, const int incx, void* y, const int incy);
void   cblas_zaxpy(const int n, const void* a, const void* x, const int incx, void* y, const int incy);
void   cblas_dscal(const int n, const double a, double* x, const int incx);
void   cblas_cscal(const int n, const void* a, void* x, const int incx);
void   cblas_zscal(const int n, const void* a, void* x, const int incx);
void   cblas_csscal(const int n, const float a, void* x, const int incx);
void   cblas_zdscal(const int n, const double a, void* x, const int incx);
float  cblas_sdot(const int n, const float* x, const int incx, const float* y, const int incy);
double cblas_ddot(const int n, const double* x, const int incx, const double* y, const int incy);
void   cblas_cdotu_sub(const int n, const void* x, const int incx, const void* y, const int incy, void* ret);
void   cblas_cdotc_sub(const int n, const void* x, const int incx, const void* y, const int incy, void* ret);
void   cblas_zdotu_sub(const int n, const void* x, const int incx, const void* y, const int incy, void* ret);
void   cblas_zdotc_sub(const int n, const void* x, const int incx, const void* y, const int incy, void* ret);
float  cblas_sasum(const int n, const float* x, const int incx);
float  cblas_scasum(const int n, const void* x, const int incx);
float  cblas_snrm2(const int n, const float* x, const int incx);
float  cblas_scnrm2(const int n, const void* x, const int incx);
size_t cblas_isamax(const int n, const float* x, const int incx);
size_t cblas_icamax(const int n, const void* x, const int incx);
size_t cblas_idamax(const int n, const double* x, const int incx);
size_t cblas_izamax(const int n, const void* x, const int incx);
void   cblas_sgemm(CBLAS_LAYOUT l, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, const int m, const int n, const int k,
                   const float alpha, const float* a, const int lda, const float* b, const int ldb,
                   const float beta, float* c, const int ldc);
void   cblas_dgemm(CBLAS_LAYOUT l, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, const int m, const int n, const int k,
                   const double alpha, const double* a, const int lda, const double* b, const int ldb,
                   const double beta, double* c, const int ldc);
void   cblas_cgemm(CBLAS_LAYOUT l, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, const int m, const int n, const int k,
                   const void* alpha, const void* a, const int lda, const void* b, const int ldb,
                   const void* beta, void* c, const int ldc);
void   cblas_zgemm(CBLAS_LAYOUT l, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, const int m, const int n, const int k,
                   const void* alpha, const void* a, const int lda, const void* b, const int ldb,
                   const void* beta, void* c, const int ldc);
void   cblas_sgemv(CBLAS_LAYOUT l, CBLAS_TRANSPOSE ta, const int m, const int n, const float alpha,
                   const float* a, const int lda, const float* x, const int incx, const float beta, float* y, const int incy);

#ifdef __cplusplus
}
#endif
#endif
