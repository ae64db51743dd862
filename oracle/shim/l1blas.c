/* TEST INFRASTRUCTURE ONLY (oracle build shim).
 *
 * The four level-1 BLAS routines that the reference convolver path calls
 * (/root/reference/framework/modules/saf_utilities/saf_utility_matrixConv.c:177,183,194,202,205,213,227,233
 *  and saf_utility_fft.c:751).  Used for oracle/_ref/libsaf_ref_conv_l1shim.so when no
 * OpenBLAS shared object is available.  They are pure copy / axpy / scale, so
 * results are identical to any conforming BLAS (alpha is always 1.0f for saxpy).
 */
#include <string.h>

void cblas_scopy(const int n, const float* x, const int incx, float* y, const int incy)
{
    if (incx == 1 && incy == 1) { memmove(y, x, (size_t)n * sizeof(float)); return; }
    for (int i = 0; i < n; i++) y[(size_t)i * incy] = x[(size_t)i * incx];
}

void cblas_ccopy(const int n, const void* xv, const int incx, void* yv, const int incy)
{
    const float* x = (const float*)xv; float* y = (float*)yv;
    if (incx == 1 && incy == 1) { memmove(y, x, (size_t)n * 2 * sizeof(float)); return; }
    for (int i = 0; i < n; i++) {
        y[2 * (size_t)i * incy]     = x[2 * (size_t)i * incx];
        y[2 * (size_t)i * incy + 1] = x[2 * (size_t)i * incx + 1];
    }
}

void cblas_saxpy(const int n, const float a, const float* x, const int incx, float* y, const int incy)
{
    for (int i = 0; i < n; i++) y[(size_t)i * incy] += a * x[(size_t)i * incx];
}

void cblas_sscal(const int n, const float a, float* x, const int incx)
{
    for (int i = 0; i < n; i++) x[(size_t)i * incx] *= a;
}
