/*
 * safconv_sh.cuh -- real spherical harmonics for the filter producers (safconv_producers.cu), __host__ __device__
 * so that tests/test_producers_host.py can compile the SAME code into a host program and check it without a GPU.
 *
 * Two evaluators, because the reference has two and the producers inherit their arithmetic:
 *
 *   scsh_rsh_dir()          getRSH (/root/reference/framework/modules/saf_hoa/saf_hoa.c:118-150) = getSHreal
 *                           (saf_sh.c:190-253) times sqrt(4 pi): orthonormal real SH in ACN order, N3D, no
 *                           Condon-Shortley phase; degrees in, the degree -> radian step in fp32 (saf_hoa.c:138-141),
 *                           everything else in fp64, rounded to fp32 at the end and scaled by sqrtf(4 pi) in fp32
 *                           (saf_hoa.c:133,147).  Used by the binaural decoder design.
 *   scsh_shreal_recur_dir() getSHreal_recur (saf_sh.c:255-330) with unnorm_legendreP_recur (saf_sh.c:129-182): the
 *                           all-fp32 recurrences the image-source simulator encodes every reflection with
 *                           (saf_reverb_internal.c:556-569).  Same operations in the same order; the normalisation
 *                           constants come from a table built once on the host (scsh_recur_norms).
 *
 * The reference builds the Legendre functions of getSHreal by a downward recursion in m with the Condon-Shortley
 * phase and cancels the phase afterwards (saf_sh.c:53-126, 217-226); here they come from the usual upward recurrence
 * in n for fixed m, without the phase -- both are fp64-accurate, the fp32 results agree except for occasional last-bit
 * rounding (tests: >= 99 % of the values bit-equal, the rest 1 ulp).
 */
#ifndef SAFCONV_SH_CUH_INCLUDED
#define SAFCONV_SH_CUH_INCLUDED

#include <math.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define SCSH_HD __host__ __device__ __forceinline__
#else
#define SCSH_HD static inline
#endif

#define SCSH_MAX_ORDER 10                         /* 121 channels; the decoder and the simulator reject higher orders */
#define SCSH_PI_F      3.14159265358979323846264338327950288f      /* SAF_PI  (saf_utilities.h) as a float constant */
#define SCSH_PI_D      3.14159265358979323846264338327950288
#define SCSH_SQRT4PI_F 3.544907701811032f                          /* SQRT4PI (saf_utilities.h)                     */

/* Y[q * stride], q = n^2 + n + m, for ONE direction given as [azimuth, elevation] in degrees. */
SCSH_HD void scsh_rsh_dir(int order, float azi_deg, float elev_deg, float* Y, int stride)
{
    const float azi_f  = azi_deg * SCSH_PI_F / 180.0f;
    const float incl_f = SCSH_PI_F / 2.0f - (elev_deg * SCSH_PI_F / 180.0f);
    const float scale  = sqrtf(4.0f * SCSH_PI_F);
    const double azi = (double)azi_f;
    const double x = cos((double)incl_f);
    double s2 = 1.0 - x * x;
    if (s2 < 0.0) s2 = 0.0;
    const double s = sqrt(s2);
    const double sqrt2 = sqrt(2.0);
    double pmm = 1.0;                              /* P_m^m = (2m-1)!! s^m                                          */
    double rmm = 1.0;                              /* (n-m)! / (n+m)! at n = m: 1 / (2m)!                           */
    for (int m = 0; m <= order; m++) {
        if (m > 0) { pmm *= (double)(2 * m - 1) * s; rmm /= (double)(2 * m - 1) * (double)(2 * m); }
        const double sm = (m > 0) ? sqrt2 * sin((double)m * azi) : 0.0;
        const double cm = (m > 0) ? sqrt2 * cos((double)m * azi) : 1.0;
        double p2 = 0.0, p1 = pmm, r = rmm;        /* P_{n-2}^m, P_{n-1}^m, ratio of factorials at n                */
        for (int n = m; n <= order; n++) {
            double p;
            if (n == m) p = pmm;
            else {
                p = ((double)(2 * n - 1) * x * p1 - (double)(n + m - 1) * p2) / (double)(n - m);
                r *= (double)(n - m) / (double)(n + m);
                p2 = p1; p1 = p;
            }
            const double norm = sqrt((2.0 * (double)n + 1.0) * r / (4.0 * SCSH_PI_D));
            const int q = n * n + n;
            if (m == 0) Y[(size_t)q * stride] = (float)(norm * p) * scale;
            else {
                Y[(size_t)(q - m) * stride] = (float)(norm * p * sm) * scale;
                Y[(size_t)(q + m) * stride] = (float)(norm * p * cm) * scale;
            }
        }
    }
}

/* Legendre polynomial P_n(x), fp64 (the value getMaxREweights takes from unnorm_legendreP, saf_hoa.c:248-258) */
SCSH_HD double scsh_legendre(int n, double x)
{
    double p0 = 1.0, p1 = x;
    if (n == 0) return 1.0;
    for (int k = 2; k <= n; k++) { const double p = ((double)(2 * k - 1) * x * p1 - (double)(k - 1) * p0) / (double)k; p0 = p1; p1 = p; }
    return p1;
}

/* normalisation table of the fp32 evaluator: c[n][0] = Nn0 / SQRT4PI, c[n][m] = Nnm / SQRT4PI with
 * Nn0 = sqrtf(2n + 1), Nnm = Nn0 * sqrtf(2 * (float)(n-m)! / (float)(n+m)!)   (saf_sh.c:287-313), fp32 as written. */
static inline void scsh_recur_norms(int order, float c[SCSH_MAX_ORDER + 1][SCSH_MAX_ORDER + 1])
{
    float fact[2 * SCSH_MAX_ORDER + 2];
    long double f = 1.0L;
    fact[0] = 1.0f;
    for (int i = 1; i < 2 * SCSH_MAX_ORDER + 2; i++) { f *= (long double)i; fact[i] = (float)f; }
    for (int n = 0; n <= SCSH_MAX_ORDER; n++)
        for (int m = 0; m <= SCSH_MAX_ORDER; m++) c[n][m] = 0.0f;
    for (int n = 0; n <= order; n++) {
        const float Nn0 = sqrtf(2.0f * (float)n + 1.0f);
        c[n][0] = (n == 0) ? 1.0f / SCSH_SQRT4PI_F : Nn0 / SCSH_SQRT4PI_F;
        for (int m = 1; m <= n; m++) {
            const float Nnm = Nn0 * sqrtf(2.0f * fact[n - m] / fact[n + m]);
            c[n][m] = Nnm / SCSH_SQRT4PI_F;
        }
    }
}

/* (2n-1)!! in fp32, multiplied up like saf_sh.c:160-166 */
SCSH_HD float scsh_dfact(int n)
{
    const int k = 2 * n - 1;
    float d = 1.0f;
    for (int kk = 1; kk < (k + 1) / 2 + 1; kk++) d *= (2.0f * (float)kk - 1.0f);
    return d;
}

#ifdef __cplusplus
/* all-fp32 real SH of ONE direction [azimuth, inclination] in radians: Y[q], q < (order+1)^2, NMAX >= order.
 * `c` = the table of scsh_recur_norms (row stride SCSH_MAX_ORDER + 1). */
template <int NMAX>
SCSH_HD void scsh_shreal_recur_dir(int order, float azi, float incl, const float* c, float* Y)
{
    float leg[NMAX + 1], leg1[NMAX + 1], leg2[NMAX + 1];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int m = 0; m <= NMAX; m++) { leg[m] = 0.0f; leg1[m] = 0.0f; leg2[m] = 0.0f; }
    const float x = cosf(incl);
    const float x2 = x * x;
    Y[0] = c[0];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int n = 1; n <= NMAX; n++) {
        if (n > order) break;
        if (n == 1) {
            leg[0] = x;
            leg[1] = sqrtf(1.0f - x2);
        } else if (n == 2) {
            leg[0] = (3.0f * x2 - 1.0f) / 2.0f;
            leg[1] = x * 3.0f * sqrtf(1.0f - x2);
            leg[2] = 3.0f * (1.0f - x2);
        } else {
            const float k = (float)(2 * n - 1);
            leg[n] = scsh_dfact(n) * powf(1.0f - x2, (float)n / 2.0f);
            leg[n - 1] = k * x * leg1[n - 1];
#if defined(__CUDACC__)
#pragma unroll
#endif
            for (int m = 0; m < NMAX - 1; m++)
                if (m < n - 1) leg[m] = ((k * x * leg1[m]) - ((float)(n + m - 1) * leg2[m])) / (float)(n - m);
        }
        const float* cn = c + n * (SCSH_MAX_ORDER + 1);
        const int q = n * n + n;
        Y[q] = cn[0] * leg[0];
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int m = 1; m <= NMAX; m++) {
            if (m > n) break;
            const float a = (float)m * azi;
            Y[q - m] = cn[m] * leg[m] * sinf(a);
            Y[q + m] = cn[m] * leg[m] * cosf(a);
        }
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int m = 0; m <= NMAX; m++) { leg2[m] = leg1[m]; leg1[m] = leg[m]; }
    }
}
#endif  /* __cplusplus */

#endif
