/*
 * safconv_producers.c -- C host layer of the filter PRODUCERS in front of the convolvers:
 *
 *   getBinauralAmbiDecoderFilters / getBinauralAmbiDecoderMtx
 *       /root/reference/framework/modules/saf_hoa/saf_hoa.h:401-471, saf_hoa.c:393-497; decoder designs
 *       saf_hoa_internal.c:162-228 (LS), :230-330 (LSDIFFEQ), :332-430 (SPR), :432-523 (TA), :525-623 (MAGLS)
 *   ims_shoebox_create / destroy / computeEchograms / renderRIRs / set* / add* / update* / remove*
 *       /root/reference/framework/modules/saf_reverb/saf_reverb.h:93-230, saf_reverb.c:36-295, 541-856
 *
 * Argument checks, planning, handles and launch sequencing only; every number is computed by the kernels of
 * safconv_producers.cu (no CPU compute path: without a CUDA device the calls fail with an error string).  What stays on
 * the host are O(order) / O(reflection order) TABLES, like the twiddle tables of the convolvers: the max-rE weights per SH
 * order, the fp32 SH normalisation constants, and the wall-reflection products per axis and reflection order (evaluated
 * with the same fp32 powf expressions as saf_reverb_internal.c:601-627).
 *
 * Both producers end in the exact array the convolver's create takes, so they can hand it over WITHOUT leaving the
 * device: safconv_binauralDecoder_create_matrixConv and safconv_ims_create_matrixConv call the filter transform (K0) on
 * the device-resident bank (for a configs[3]-sized bank that skips a 1.57 GB device -> host -> device round trip).
 */
#include "safconv_host_internal.h"
#include "safconv_sh.cuh"
#include "safconv_prod_core.cuh"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PROD_TRY(call, what) do { int e_ = (call); if (e_) { rc = prod_fail(SAFCONV_ERR_CUDA, what, e_); goto done; } } while (0)

static int prod_fail(int code, const char* what, int cudaErr)
{
    char buf[256];
    if (cudaErr) snprintf(buf, sizeof buf, "%s: %s", what, scdev_error_string(cudaErr));
    else         snprintf(buf, sizeof buf, "%s", what);
    sch_set_tl_error(code, "%s", buf);
    return code;
}

static int prod_device_begin(int* device, int* smCount)
{
    int ndev = 0;
    if (scdev_device_count(&ndev) != 0 || ndev < 1)
        return prod_fail(SAFCONV_ERR_NO_DEVICE, "no usable CUDA device (libsafconv_b200 has no CPU fallback)", 0);
    int dev = sch_thread_device();
    int e = 0;
    if (dev < 0) e = scdev_get_device(&dev);
    if (!e) e = scdev_set_device(dev);
    int maxSmem = 0, ccMaj = 0, ccMin = 0;
    if (!e) e = scdev_device_props(dev, smCount, &maxSmem, &ccMaj, &ccMin);
    if (e) return prod_fail(SAFCONV_ERR_CUDA, "device set-up", e);
    *device = dev;
    return SAFCONV_OK;
}

/* ================================================================================================================ */
/*  binaural Ambisonic decoder                                                                                       */
/* ================================================================================================================ */

/* methods, saf_hoa.h:131-171 */
enum { DEC_DEFAULT = 0, DEC_LS = 1, DEC_LSDIFFEQ = 2, DEC_SPR = 3, DEC_TA = 4, DEC_MAGLS = 5 };

/* ---- t-designs for the SPR decoder ------------------------------------------------------------------------------
 * getBinDecoder_SPR projects on the reference's minimum t-design of degree 2 * order (saf_hoa_internal.c:383-389:
 * __HANDLES_Tdesign_dirs_deg[2*order-1], tables of saf_utility_loudspeaker_presets.c).  Those tables are SAF's data and
 * are not copied into this library; it takes them from (1) what the host registered with safconv_register_tdesign, or
 * (2) SAF's own symbols if the process that loaded this library carries them (WEAK references: NULL otherwise). */
extern const float* __HANDLES_Tdesign_dirs_deg[21] __attribute__((weak));
extern const int    __Tdesign_nPoints_per_degree[21] __attribute__((weak));

#define TD_MAX_DEGREE 64
static struct { float* dirs; int n; } g_td[TD_MAX_DEGREE + 1];
static pthread_mutex_t g_td_lock = PTHREAD_MUTEX_INITIALIZER;

int safconv_register_tdesign(int degree, const float* dirs_deg, int nPoints)
{
    if (degree < 1 || degree > TD_MAX_DEGREE || !dirs_deg || nPoints < 1)
        return prod_fail(SAFCONV_ERR_ARG, "safconv_register_tdesign: need 1 <= degree <= 64, a direction list and nPoints >= 1", 0);
    float* c = (float*)malloc(sizeof(float) * 2 * (size_t)nPoints);
    if (!c) return prod_fail(SAFCONV_ERR_NOMEM, "out of host memory", 0);
    memcpy(c, dirs_deg, sizeof(float) * 2 * (size_t)nPoints);
    pthread_mutex_lock(&g_td_lock);
    free(g_td[degree].dirs);
    g_td[degree].dirs = c; g_td[degree].n = nPoints;
    pthread_mutex_unlock(&g_td_lock);
    return SAFCONV_OK;
}

/* copy of the t-design of `degree` (caller frees), or NULL */
static float* tdesign_lookup(int degree, int* nPoints)
{
    float* out = NULL;
    pthread_mutex_lock(&g_td_lock);
    if (degree >= 1 && degree <= TD_MAX_DEGREE && g_td[degree].dirs) {
        *nPoints = g_td[degree].n;
        out = (float*)malloc(sizeof(float) * 2 * (size_t)g_td[degree].n);
        if (out) memcpy(out, g_td[degree].dirs, sizeof(float) * 2 * (size_t)g_td[degree].n);
    }
    pthread_mutex_unlock(&g_td_lock);
    if (!out && degree >= 1 && degree <= 21 && __HANDLES_Tdesign_dirs_deg && __Tdesign_nPoints_per_degree &&
        __HANDLES_Tdesign_dirs_deg[degree - 1]) {
        *nPoints = __Tdesign_nPoints_per_degree[degree - 1];
        out = (float*)malloc(sizeof(float) * 2 * (size_t)*nPoints);
        if (out) memcpy(out, __HANDLES_Tdesign_dirs_deg[degree - 1], sizeof(float) * 2 * (size_t)*nPoints);
    }
    return out;
}

/* Decoding matrices of all bands on the device: *pD = float2 [nB][2][nSH] (caller frees). */
static int decoder_mtx_device(const void* hrtfs, const float* dirs_deg, int nD, int nB, int method, int order,
                              const float* freqVector, const float* weights, int diffCM, int maxRE,
                              void* stream, void** pD)
{
    int rc = SAFCONV_OK;
    const int n = (order + 1) * (order + 1);
    void *d_H = NULL, *d_D = NULL, *d_hm = NULL;
    float *d_dirs = NULL, *d_Y = NULL, *d_w = NULL, *d_G = NULL, *d_a = NULL;
    float *d_tdirs = NULL, *d_Ytd = NULL, *d_ws = NULL, *d_cond = NULL;
    double *d_aug = NULL, *d_M = NULL;
    int* d_flag = NULL;
    float *w = NULL, *ws = NULL, *td = NULL;
    int K = 0, nhMax = 0, nS = n;            /* SPR: t-design points, highest candidate interpolation order, its channel count */
    *pD = NULL;

    if (method < DEC_DEFAULT || method > DEC_MAGLS) method = DEC_DEFAULT;      /* the reference's `default:` label, saf_hoa.c:415 */
    if ((method == DEC_TA || method == DEC_MAGLS) && !freqVector)
        return prod_fail(SAFCONV_ERR_ARG, "getBinauralAmbiDecoderMtx: TA / MAGLS need freqVector", 0);

    /* band closest to 1.5 kHz, first minimum (saf_hoa_internal.c:465-473, 559-567) */
    int bc = 0;
    if (method == DEC_TA || method == DEC_MAGLS) {
        float minVal = 2.23e10f;
        for (int b = 0; b < nB; b++)
            if (minVal > fabsf(freqVector[b] - 1.5e3f)) { minVal = fabsf(freqVector[b] - 1.5e3f); bc = b; }
    }

    /* integration weights: the given ones or 1 / N_dirs (saf_hoa_internal.c:192-199) */
    w = (float*)malloc(sizeof(float) * (size_t)nD);
    if (!w) return prod_fail(SAFCONV_ERR_NOMEM, "out of host memory", 0);
    for (int i = 0; i < nD; i++) w[i] = weights ? weights[i] : 1.0f / (float)nD;

    if (method == DEC_SPR) {
        td = tdesign_lookup(2 * order, &K);
        if (!td) { free(w); return prod_fail(SAFCONV_ERR_ARG, "SPR: no t-design of degree 2 * order (safconv_register_tdesign)", 0); }
        /* candidate interpolation orders 0 .. min(sqrt(N_dirs) - 1, 20) (saf_hoa_internal.c:357) */
        nhMax = (int)(sqrtf((float)nD) - 1.0f);
        if (nhMax > 20) nhMax = 20;
        if (nhMax < 0) nhMax = 0;
        nS = (nhMax + 1) * (nhMax + 1);
        if (nS < n) nS = n;
        /* weights of the condition check: the caller's, or none (saf_sh.c:905-914); of the projection: weights / 4 pi or 1 / N (:347-353) */
        ws = (float*)malloc(sizeof(float) * 2 * (size_t)nD);
        if (!ws) { free(w); free(td); return prod_fail(SAFCONV_ERR_NOMEM, "out of host memory", 0); }
        for (int i = 0; i < nD; i++) { ws[i] = weights ? weights[i] : 1.0f; ws[nD + i] = weights ? weights[i] / (4.0f * SCSH_PI_F) : 1.0f / (float)nD; }
    }
    const size_t bytesH = sizeof(float) * 2 * (size_t)nB * 2 * nD;
    int e = scdev_malloc(&d_H, bytesH);
    if (!e) e = scdev_malloc((void**)&d_dirs, sizeof(float) * 2 * (size_t)nD);
    if (!e) e = scdev_malloc((void**)&d_Y, sizeof(float) * (size_t)nS * nD);
    if (!e) e = scdev_malloc((void**)&d_w, sizeof(float) * (size_t)nD);
    if (!e) e = scdev_malloc((void**)&d_G, sizeof(float) * (size_t)n * nD);
    if (!e) e = scdev_malloc((void**)&d_aug, sizeof(double) * 2 * (size_t)nS * nS);
    if (!e && method == DEC_SPR) {
        e = scdev_malloc((void**)&d_tdirs, sizeof(float) * 2 * (size_t)K);
        if (!e) e = scdev_malloc((void**)&d_Ytd, sizeof(float) * (size_t)nS * K);
        if (!e) e = scdev_malloc((void**)&d_ws, sizeof(float) * 2 * (size_t)nD);
        if (!e) e = scdev_malloc((void**)&d_cond, sizeof(float) * (size_t)(nhMax + 1));
        if (!e) e = scdev_malloc((void**)&d_M, sizeof(double) * (size_t)nS * n);
    }
    if (!e) e = scdev_malloc((void**)&d_flag, sizeof(int));
    if (!e) e = scdev_malloc(&d_D, sizeof(float) * 2 * (size_t)nB * 2 * n);
    if (!e && method == DEC_MAGLS) e = scdev_malloc(&d_hm, sizeof(float) * 2 * 2 * (size_t)nD);
    if (e) { rc = prod_fail(SAFCONV_ERR_NOMEM, "decoder design: device allocation", e); goto done; }

    PROD_TRY(scdev_memcpy_h2d_async(d_H, hrtfs, bytesH, stream), "HRTF upload");
    PROD_TRY(scdev_memcpy_h2d_async(d_dirs, dirs_deg, sizeof(float) * 2 * (size_t)nD, stream), "direction upload");
    PROD_TRY(scdev_memcpy_h2d_async(d_w, w, sizeof(float) * (size_t)nD, stream), "weight upload");
    if (method == DEC_SPR) {
        /* getBinDecoder_SPR (saf_hoa_internal.c:332-430): interpolate the HRTF set with SH of the highest well-conditioned
         * order Nh, evaluate it on the t-design of degree 2 * order, project on the SH of the decoding order -- folded
         * into ONE matrix G, so that the decoder of every band is the same product D = H G^T as the least-squares one */
        float cond[21];
        int Nh = 0;
        PROD_TRY(scdev_memcpy_h2d_async(d_ws, ws, sizeof(float) * 2 * (size_t)nD, stream), "weight upload");
        PROD_TRY(scdev_memcpy_h2d_async(d_tdirs, td, sizeof(float) * 2 * (size_t)K, stream), "t-design upload");
        PROD_TRY(scdev_prod_rsh(nhMax > order ? nhMax : order, d_dirs, nD, d_Y, stream), "SH evaluation");     /* rows of order <= N first (ACN) */
        PROD_TRY(scdev_prod_spr_cond(d_Y, d_ws, nD, nhMax, d_aug, d_cond, stream), "condition numbers");
        PROD_TRY(scdev_memcpy_d2h_async(cond, d_cond, sizeof(float) * (size_t)(nhMax + 1), stream), "condition number download");
        PROD_TRY(scdev_stream_sync(stream), "condition numbers");
        for (int i = 0; i <= nhMax; i++) Nh = (cond[i] < 100.0f) ? i : Nh;                                    /* :369-370 */
        if (Nh < order) { rc = prod_fail(SAFCONV_ERR_ARG, "SPR: input order exceeds the modal order of the spatial grid (saf_hoa_internal.c:371)", 0); goto done; }
        const int nA = (Nh + 1) * (Nh + 1);
        PROD_TRY(scdev_prod_rsh(Nh, d_tdirs, K, d_Ytd, stream), "SH evaluation (t-design)");
        PROD_TRY(scdev_prod_spr_matrix(d_Y, d_Ytd, d_ws + nD, nD, K, nA, n, d_M, d_G, stream), "SPR matrix");
    } else {
        PROD_TRY(scdev_prod_rsh(order, d_dirs, nD, d_Y, stream), "SH evaluation");
        PROD_TRY(scdev_prod_lsmatrix(d_Y, d_w, nD, n, d_aug, d_G, d_flag, stream), "least-squares matrix");
        int flag = 0;
        PROD_TRY(scdev_memcpy_d2h_async(&flag, d_flag, sizeof(int), stream), "flag download");
        PROD_TRY(scdev_stream_sync(stream), "least-squares matrix");
        if (flag) { rc = prod_fail(SAFCONV_ERR_ARG, "decoder design: the SH Gram matrix of the measurement grid is singular at this order", 0); goto done; }
    }
    PROD_TRY(scdev_prod_ls(d_H, d_G, nB, nD, n, method == DEC_TA, bc, d_D, stream), "least-squares decoder");
    if (method == DEC_LSDIFFEQ) PROD_TRY(scdev_prod_diffeq(d_H, d_Y, d_w, nB, nD, n, d_D, stream), "diffuse-field equalisation");
    if (method == DEC_MAGLS)    PROD_TRY(scdev_prod_magls(d_H, d_Y, d_G, nB, nD, n, bc, d_D, d_hm, stream), "MagLS recurrence");
    if (maxRE) {
        /* a_n = P_n(cos(137.9 deg / (order + 1.51))) per SH order (getMaxREweights, saf_hoa.c:235-266) */
        float a[(SCSH_MAX_ORDER + 1) * (SCSH_MAX_ORDER + 1)];
        const double x = (double)cosf(137.9f * (SCSH_PI_F / 180.0f) / ((float)order + 1.51f));
        for (int nn = 0; nn <= order; nn++) {
            const float v = (float)scsh_legendre(nn, x);
            for (int q = nn * nn; q < (nn + 1) * (nn + 1); q++) a[q] = v;
        }
        e = scdev_malloc((void**)&d_a, sizeof(float) * (size_t)n);
        if (e) { rc = prod_fail(SAFCONV_ERR_NOMEM, "max-rE weights", e); goto done; }
        PROD_TRY(scdev_memcpy_h2d_sync(d_a, a, sizeof(float) * (size_t)n, stream), "max-rE weight upload");
        PROD_TRY(scdev_prod_scale(d_D, d_a, n, (size_t)nB * 2 * n, stream), "max-rE weighting");
    }
    if (diffCM) PROD_TRY(scdev_prod_diffcov(d_H, d_Y, d_w, nB, nD, n, d_D, stream), "diffuse-field covariance matching");
    PROD_TRY(scdev_stream_sync(stream), "decoder design");
    *pD = d_D; d_D = NULL;
done:
    free(w); free(ws); free(td);
    scdev_free(d_H); scdev_free(d_dirs); scdev_free(d_Y); scdev_free(d_w); scdev_free(d_G); scdev_free(d_aug);
    scdev_free(d_flag); scdev_free(d_D); scdev_free(d_hm); scdev_free(d_a);
    scdev_free(d_tdirs); scdev_free(d_Ytd); scdev_free(d_ws); scdev_free(d_cond); scdev_free(d_M);
    return rc;
}

static int decoder_check(const void* hrtfs, const float* dirs, int nD, int nB, int method, int order, const void* out)
{
    if (!hrtfs || !dirs || !out || nD < 1 || nB < 1 || order < 0)
        return prod_fail(SAFCONV_ERR_ARG, "getBinauralAmbiDecoder*: invalid argument", 0);
    if (method == DEC_SPR) {
        /* the reference indexes its t-design table with 2 * order - 1 (saf_hoa_internal.c:384): order 0 reads before the table */
        if (order < 1) return prod_fail(SAFCONV_ERR_ARG, "getBinauralAmbiDecoder*: BINAURAL_DECODER_SPR needs order >= 1", 0);
        int k = 0;
        float* td = tdesign_lookup(2 * order, &k);
        if (!td)
            return prod_fail(SAFCONV_ERR_ARG, "getBinauralAmbiDecoder*: BINAURAL_DECODER_SPR needs the t-design of degree 2 * order: register it with "
                             "safconv_register_tdesign (SAF's tables are not copied into this library; a host that carries SAF's "
                             "__HANDLES_Tdesign_dirs_deg is served without)", 0);
        free(td);
    }
    if (order > SCSH_MAX_ORDER)
        return prod_fail(SAFCONV_ERR_ARG, "getBinauralAmbiDecoder*: order > 10 is not supported", 0);
    return SAFCONV_OK;
}

int safconv_getBinauralAmbiDecoderMtx(const void* hrtfs, const float* hrtf_dirs_deg, int N_dirs, int N_bands, int method,
                                      int order, const float* freqVector, const float* itd_s, const float* weights,
                                      int enableDiffCovMatching, int enableMaxReWeighting, void* decMtx)
{
    (void)itd_s;    /* read by nothing that reaches the result: the TA phase term is exp(0 * itd) (saf_hoa_internal.c:494-497) */
    sch_set_tl_error(0, "%s", "");
    int rc = decoder_check(hrtfs, hrtf_dirs_deg, N_dirs, N_bands, method, order, decMtx);
    if (rc) return rc;
    int dev = 0, sm = 0;
    if ((rc = prod_device_begin(&dev, &sm)) != 0) return rc;
    void* stream = NULL; void* d_D = NULL;
    int e = scdev_stream_create(&stream);
    if (e) return prod_fail(SAFCONV_ERR_CUDA, "cudaStreamCreate", e);
    rc = decoder_mtx_device(hrtfs, hrtf_dirs_deg, N_dirs, N_bands, method, order, freqVector, weights,
                            enableDiffCovMatching, enableMaxReWeighting, stream, &d_D);
    if (!rc) {
        const int n = (order + 1) * (order + 1);
        e = scdev_memcpy_d2h_async(decMtx, d_D, sizeof(float) * 2 * (size_t)N_bands * 2 * n, stream);
        if (!e) e = scdev_stream_sync(stream);
        if (e) rc = prod_fail(SAFCONV_ERR_CUDA, "decoder matrix download", e);
    }
    scdev_free(d_D);
    scdev_stream_destroy(stream);
    return rc;
}

/* decoder filters on the device: *pF = float [2][nSH][fftSize] (caller frees) */
static int decoder_filters_device(const void* hrtfs, const float* dirs_deg, int nD, int fftSize, float fs, int method, int order,
                                  const float* weights, int diffCM, int maxRE, void* stream, float** pF)
{
    int rc = SAFCONV_OK;
    const int nB = fftSize / 2 + 1, n = (order + 1) * (order + 1), rows = 2 * n;
    void *d_D = NULL, *d_Dt = NULL;
    float* d_F = NULL;
    float* freq = NULL;
    scdev_gfft_plan pl;
    int havePlan = 0;
    *pF = NULL;
    memset(&pl, 0, sizeof pl);
    /* uniform frequency vector (getUniformFreqVector, saf_utility_fft.c:145-155) */
    freq = (float*)malloc(sizeof(float) * (size_t)nB);
    if (!freq) return prod_fail(SAFCONV_ERR_NOMEM, "out of host memory", 0);
    for (int k = 0; k < nB; k++) freq[k] = (float)k * fs / (float)fftSize;
    rc = decoder_mtx_device(hrtfs, dirs_deg, nD, nB, method, order, freq, weights, diffCM, maxRE, stream, &d_D);
    if (rc) goto done;
    int e = scdev_malloc(&d_Dt, sizeof(float) * 2 * (size_t)rows * nB);
    if (!e) e = scdev_malloc((void**)&d_F, sizeof(float) * (size_t)rows * fftSize);
    if (e) { rc = prod_fail(SAFCONV_ERR_NOMEM, "decoder filters: device allocation", e); goto done; }
    e = scr_plan_init(&pl, fftSize, stream);
    if (!e) { havePlan = 1; e = scr_plan_reserve(&pl, rows); }
    if (e) { rc = prod_fail(e < 0 ? SAFCONV_ERR_NOMEM : SAFCONV_ERR_CUDA, "decoder filters: FFT plan", e < 0 ? 0 : e); goto done; }
    /* one inverse real FFT per (ear, SH channel) over the bins (saf_hoa.c:480-490) */
    PROD_TRY(scdev_prod_pack(d_D, nB, rows, d_Dt, stream), "decoder transpose");
    PROD_TRY(scdev_gfft_run(&pl, 1, rows, (const float*)d_Dt, d_F, stream), "inverse FFT over the bins");
    PROD_TRY(scdev_stream_sync(stream), "decoder filters");
    *pF = d_F; d_F = NULL;
done:
    free(freq);
    if (havePlan) scr_plan_free(&pl);
    scdev_free(d_D); scdev_free(d_Dt); scdev_free(d_F);
    return rc;
}

static int decoder_filters_check(const void* hrtfs, const float* dirs, int nD, int fftSize, int method, int order, const void* out)
{
    if (fftSize < 2 || (fftSize & 1))
        return prod_fail(SAFCONV_ERR_ARG, "getBinauralAmbiDecoderFilters: fftSize must be even and >= 2 (saf_utility_fft.c:542)", 0);
    return decoder_check(hrtfs, dirs, nD, fftSize / 2 + 1, method, order, out);
}

int safconv_getBinauralAmbiDecoderFilters(const void* hrtfs, const float* hrtf_dirs_deg, int N_dirs, int fftSize, float fs,
                                          int method, int order, const float* itd_s, const float* weights,
                                          int enableDiffCovMatching, int enableMaxReWeighting, float* decFilters)
{
    (void)itd_s;
    sch_set_tl_error(0, "%s", "");
    int rc = decoder_filters_check(hrtfs, hrtf_dirs_deg, N_dirs, fftSize, method, order, decFilters);
    if (rc) return rc;
    int dev = 0, sm = 0;
    if ((rc = prod_device_begin(&dev, &sm)) != 0) return rc;
    void* stream = NULL; float* d_F = NULL;
    int e = scdev_stream_create(&stream);
    if (e) return prod_fail(SAFCONV_ERR_CUDA, "cudaStreamCreate", e);
    rc = decoder_filters_device(hrtfs, hrtf_dirs_deg, N_dirs, fftSize, fs, method, order, weights,
                                enableDiffCovMatching, enableMaxReWeighting, stream, &d_F);
    if (!rc) {
        const int n = (order + 1) * (order + 1);
        e = scdev_memcpy_d2h_async(decFilters, d_F, sizeof(float) * 2 * (size_t)n * fftSize, stream);
        if (!e) e = scdev_stream_sync(stream);
        if (e) rc = prod_fail(SAFCONV_ERR_CUDA, "decoder filter download", e);
    }
    scdev_free(d_F);
    scdev_stream_destroy(stream);
    return rc;
}

int safconv_binauralDecoder_create_matrixConv(void** const phMC, int hopSize, const void* hrtfs, const float* hrtf_dirs_deg,
                                              int N_dirs, int fftSize, float fs, int method, int order, const float* weights,
                                              int enableDiffCovMatching, int enableMaxReWeighting)
{
    sch_set_tl_error(0, "%s", "");
    if (!phMC) return prod_fail(SAFCONV_ERR_ARG, "binauralDecoder_create_matrixConv: phMC is NULL", 0);
    *phMC = NULL;
    int rc = decoder_filters_check(hrtfs, hrtf_dirs_deg, N_dirs, fftSize, method, order, phMC);
    if (rc) return rc;
    int dev = 0, sm = 0;
    if ((rc = prod_device_begin(&dev, &sm)) != 0) return rc;
    void* stream = NULL; float* d_F = NULL;
    int e = scdev_stream_create(&stream);
    if (e) return prod_fail(SAFCONV_ERR_CUDA, "cudaStreamCreate", e);
    rc = decoder_filters_device(hrtfs, hrtf_dirs_deg, N_dirs, fftSize, fs, method, order, weights,
                                enableDiffCovMatching, enableMaxReWeighting, stream, &d_F);
    if (!rc) {
        /* decFilters is FLAT 2 x nSH x fftSize = nCHout x nCHin x length_h (saf_hoa.h:455, saf_utility_matrixConv.h:48) */
        safconv_matrixConv_create_device(phMC, hopSize, d_F, fftSize, (order + 1) * (order + 1), 2);
        if (!*phMC) rc = safconv_last_error(NULL) ? safconv_last_error(NULL) : SAFCONV_ERR_CUDA;
    }
    scdev_free(d_F);
    scdev_stream_destroy(stream);
    return rc;
}

/* the reference's names; weak, so that a host that also links the reference's saf_hoa.c keeps its own */
__attribute__((weak)) void getBinauralAmbiDecoderMtx(void* hrtfs, float* hrtf_dirs_deg, int N_dirs, int N_bands, int method, int order,
                                                     float* freqVector, float* itd_s, float* weights, int enableDiffCovMatching,
                                                     int enableMaxReWeighting, void* decMtx)
{
    (void)safconv_getBinauralAmbiDecoderMtx(hrtfs, hrtf_dirs_deg, N_dirs, N_bands, method, order, freqVector, itd_s, weights,
                                            enableDiffCovMatching, enableMaxReWeighting, decMtx);
}
__attribute__((weak)) void getBinauralAmbiDecoderFilters(void* hrtfs, float* hrtf_dirs_deg, int N_dirs, int fftSize, float fs, int method,
                                                         int order, float* itd_s, float* weights, int enableDiffCovMatching,
                                                         int enableMaxReWeighting, float* decFilters)
{
    (void)safconv_getBinauralAmbiDecoderFilters(hrtfs, hrtf_dirs_deg, N_dirs, fftSize, fs, method, order, itd_s, weights,
                                                enableDiffCovMatching, enableMaxReWeighting, decFilters);
}

/* ================================================================================================================ */
/*  shoebox image-source simulator                                                                                   */
/* ================================================================================================================ */
#define IMS_MAGIC 0x5AFC1A55u
#define IMS_MAX_SRC 128            /* IMS_MAX_NUM_SOURCES,   saf_reverb.h:52 */
#define IMS_MAX_REC 16             /* IMS_MAX_NUM_RECEIVERS, saf_reverb.h:55 */
#define IMS_UNASSIGNED (-1)

typedef struct ims_obj { int id; float pos[3]; int nCh; } ims_obj;

typedef struct ims_pair {
    int refreshEcho, refreshRIR;   /* refreshEchogramFLAG / refreshRIRFLAG of ims_core_workspace (saf_reverb_internal.h:155-198) */
    int haveParams, mode, maxN;    /* parameters of the last computeEchograms */
    float maxTime;
    float* d_rir;                  /* device [nCh][len] */
    float* h_rir;                  /* host copy, made on request */
    int len, nCh, nImages;
} ims_pair;

typedef struct safconv_ims {
    uint32_t magic;
    int device, smCount;
    void* stream;
    float room[3], c_ms, fs;
    int nBands;
    float* abs_wall;               /* [nBands][6] */
    ims_obj srcs[IMS_MAX_SRC], recs[IMS_MAX_REC];
    int nSources, nReceivers;
    ims_pair pair[IMS_MAX_REC][IMS_MAX_SRC];
    float* d_norms;                /* scsh_recur_norms table */
} safconv_ims;

static safconv_ims* as_ims(void* p)
{
    safconv_ims* s = (safconv_ims*)p;
    return (s && s->magic == IMS_MAGIC) ? s : NULL;
}

static void pair_clear(ims_pair* p)
{
    scdev_free(p->d_rir); free(p->h_rir);
    memset(p, 0, sizeof *p);
}

void safconv_ims_shoebox_create(void** phIms, float roomDimensions[3], float* abs_wall, float lowestOctaveBand, int nOctBands,
                                float c_ms, float fs)
{
    (void)lowestOctaveBand;   /* only names the filterbank bands, whose output the reference discards (saf_reverb_internal.c:697-702) */
    sch_set_tl_error(0, "%s", "");
    if (!phIms) return;
    *phIms = NULL;
    if (!roomDimensions || !abs_wall || !(c_ms > 0.0f) || !(fs > 0.0f) || !(roomDimensions[0] > 0.0f) || !(roomDimensions[1] > 0.0f) ||
        !(roomDimensions[2] > 0.0f)) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_create: invalid argument", 0); return; }
    int dev = 0, sm = 0;
    if (prod_device_begin(&dev, &sm)) return;
    safconv_ims* s = (safconv_ims*)calloc(1, sizeof *s);
    if (!s) { prod_fail(SAFCONV_ERR_NOMEM, "out of host memory", 0); return; }
    s->magic = IMS_MAGIC; s->device = dev; s->smCount = sm;
    s->room[0] = roomDimensions[0]; s->room[1] = roomDimensions[1]; s->room[2] = roomDimensions[2];
    s->c_ms = c_ms; s->fs = fs;
    s->nBands = nOctBands > 1 ? nOctBands : 1;                                  /* saf_reverb.c:57-72 */
    s->abs_wall = (float*)malloc(sizeof(float) * 6 * (size_t)s->nBands);
    for (int i = 0; i < IMS_MAX_SRC; i++) s->srcs[i].id = IMS_UNASSIGNED;
    for (int i = 0; i < IMS_MAX_REC; i++) s->recs[i].id = IMS_UNASSIGNED;
    float norms[SCSH_MAX_ORDER + 1][SCSH_MAX_ORDER + 1];
    scsh_recur_norms(SCSH_MAX_ORDER, norms);
    int e = s->abs_wall ? 0 : -1;
    if (!e) memcpy(s->abs_wall, abs_wall, sizeof(float) * 6 * (size_t)s->nBands);
    if (!e) e = scdev_stream_create(&s->stream);
    if (!e) e = scdev_malloc((void**)&s->d_norms, sizeof norms);
    if (!e) e = scdev_memcpy_h2d_sync(s->d_norms, norms, sizeof norms, s->stream);
    if (e) {
        prod_fail(e < 0 ? SAFCONV_ERR_NOMEM : SAFCONV_ERR_CUDA, "ims_shoebox_create", e < 0 ? 0 : e);
        scdev_free(s->d_norms); if (s->stream) scdev_stream_destroy(s->stream); free(s->abs_wall); free(s);
        return;
    }
    *phIms = s;
}

void safconv_ims_shoebox_destroy(void** phIms)
{
    if (!phIms) return;
    safconv_ims* s = as_ims(*phIms);
    if (s) {
        scdev_set_device(s->device);
        scdev_stream_sync(s->stream);
        for (int r = 0; r < IMS_MAX_REC; r++) for (int k = 0; k < IMS_MAX_SRC; k++) pair_clear(&s->pair[r][k]);
        scdev_free(s->d_norms);
        scdev_stream_destroy(s->stream);
        free(s->abs_wall);
        s->magic = 0;
        free(s);
    }
    *phIms = NULL;
}

static int find_obj(const ims_obj* o, int n, int id)
{
    if (id < 0) return -1;
    for (int i = 0; i < n; i++) if (o[i].id == id) return i;
    return -1;
}

/* slot + ID assignment of saf_reverb.c:598-647 / :649-702: first free slot; ID = 0, incremented once for every OTHER
 * slot (in slot order) that currently holds the candidate value */
static int add_obj(ims_obj* o, int n, const float xyz[3])
{
    int idx = -1;
    for (int i = 0; i < n; i++) if (o[i].id == IMS_UNASSIGNED) { idx = i; break; }
    if (idx < 0) return -1;
    o[idx].id = 0;
    for (int i = 0; i < n; i++) if (i != idx && o[i].id == o[idx].id) o[idx].id++;
    o[idx].pos[0] = xyz[0]; o[idx].pos[1] = xyz[1]; o[idx].pos[2] = xyz[2];
    return idx;
}

int safconv_ims_shoebox_addSource(void* hIms, float position_xyz[3], float** pSrc_sig)
{
    (void)pSrc_sig;       /* signal pointers belong to ims_shoebox_applyEchogramTD, which is not part of the RIR path */
    safconv_ims* s = as_ims(hIms);
    if (!s || !position_xyz) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_addSource: invalid argument", 0); return -1; }
    const int idx = add_obj(s->srcs, IMS_MAX_SRC, position_xyz);
    if (idx < 0) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_addSource: exceeded the maximum supported number of sources (128)", 0); return -1; }
    s->nSources++;
    for (int r = 0; r < IMS_MAX_REC; r++) { pair_clear(&s->pair[r][idx]); s->pair[r][idx].refreshEcho = 1; }
    return s->srcs[idx].id;
}

int safconv_ims_shoebox_addReceiverSH(void* hIms, int sh_order, float position_xyz[3], float*** pSH_sigs)
{
    (void)pSH_sigs;
    safconv_ims* s = as_ims(hIms);
    if (!s || !position_xyz || sh_order < 0) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_addReceiverSH: invalid argument", 0); return -1; }
    if (sh_order > SCSH_MAX_ORDER) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_addReceiverSH: sh_order > 10 is not supported", 0); return -1; }
    const int idx = add_obj(s->recs, IMS_MAX_REC, position_xyz);
    if (idx < 0) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_addReceiverSH: exceeded the maximum supported number of receivers (16)", 0); return -1; }
    s->recs[idx].nCh = (sh_order + 1) * (sh_order + 1);
    s->nReceivers++;
    for (int k = 0; k < IMS_MAX_SRC; k++) { pair_clear(&s->pair[idx][k]); s->pair[idx][k].refreshEcho = 1; }
    return s->recs[idx].id;
}

void safconv_ims_shoebox_updateSource(void* hIms, int sourceID, float position_xyz[3])
{
    safconv_ims* s = as_ims(hIms);
    const int k = s ? find_obj(s->srcs, IMS_MAX_SRC, sourceID) : -1;
    if (k < 0 || !position_xyz) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_updateSource: invalid sourceID", 0); return; }
    float* p = s->srcs[k].pos;
    if (p[0] != position_xyz[0] || p[1] != position_xyz[1] || p[2] != position_xyz[2]) {
        p[0] = position_xyz[0]; p[1] = position_xyz[1]; p[2] = position_xyz[2];
        for (int r = 0; r < IMS_MAX_REC; r++) s->pair[r][k].refreshEcho = 1;
    }
}

void safconv_ims_shoebox_updateReceiver(void* hIms, int receiverID, float position_xyz[3])
{
    safconv_ims* s = as_ims(hIms);
    const int r = s ? find_obj(s->recs, IMS_MAX_REC, receiverID) : -1;
    if (r < 0 || !position_xyz) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_updateReceiver: invalid receiverID", 0); return; }
    float* p = s->recs[r].pos;
    if (p[0] != position_xyz[0] || p[1] != position_xyz[1] || p[2] != position_xyz[2]) {
        p[0] = position_xyz[0]; p[1] = position_xyz[1]; p[2] = position_xyz[2];
        for (int k = 0; k < IMS_MAX_SRC; k++) s->pair[r][k].refreshEcho = 1;
    }
}

void safconv_ims_shoebox_removeSource(void* hIms, int sourceID)
{
    safconv_ims* s = as_ims(hIms);
    const int k = s ? find_obj(s->srcs, IMS_MAX_SRC, sourceID) : -1;
    if (k < 0) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_removeSource: invalid sourceID", 0); return; }
    scdev_set_device(s->device); scdev_stream_sync(s->stream);
    s->srcs[k].id = IMS_UNASSIGNED;
    for (int r = 0; r < IMS_MAX_REC; r++) pair_clear(&s->pair[r][k]);
    s->nSources--;
}

void safconv_ims_shoebox_removeReceiver(void* hIms, int receiverID)
{
    safconv_ims* s = as_ims(hIms);
    const int r = s ? find_obj(s->recs, IMS_MAX_REC, receiverID) : -1;
    if (r < 0) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_removeReceiver: invalid receiverID", 0); return; }
    scdev_set_device(s->device); scdev_stream_sync(s->stream);
    s->recs[r].id = IMS_UNASSIGNED;
    for (int k = 0; k < IMS_MAX_SRC; k++) pair_clear(&s->pair[r][k]);
    s->nReceivers--;
}

static void flag_all_pairs(safconv_ims* s)
{
    for (int r = 0; r < IMS_MAX_REC; r++)
        for (int k = 0; k < IMS_MAX_SRC; k++)
            if (s->recs[r].id != IMS_UNASSIGNED && s->srcs[k].id != IMS_UNASSIGNED) s->pair[r][k].refreshEcho = 1;
}

void safconv_ims_shoebox_setRoomDimensions(void* hIms, float new_roomDimensions[3])
{
    safconv_ims* s = as_ims(hIms);
    if (!s || !new_roomDimensions) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_setRoomDimensions: invalid argument", 0); return; }
    if (s->room[0] != new_roomDimensions[0] || s->room[1] != new_roomDimensions[1] || s->room[2] != new_roomDimensions[2]) {
        s->room[0] = new_roomDimensions[0]; s->room[1] = new_roomDimensions[1]; s->room[2] = new_roomDimensions[2];
        flag_all_pairs(s);
    }
}

void safconv_ims_shoebox_setWallAbsCoeffs(void* hIms, float* abs_wall)
{
    safconv_ims* s = as_ims(hIms);
    if (!s || !abs_wall) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_setWallAbsCoeffs: invalid argument", 0); return; }
    int changed = 0;
    for (int i = 0; i < 6 * s->nBands; i++) if (s->abs_wall[i] != abs_wall[i]) { s->abs_wall[i] = abs_wall[i]; changed = 1; }
    if (changed) flag_all_pairs(s);
}

/* Which pairs need a new echogram (saf_reverb.c:184-257).  The image sources themselves are enumerated by the render
 * kernels; this call only records the request. */
void safconv_ims_shoebox_computeEchograms(void* hIms, int maxN, float maxTime_s)
{
    safconv_ims* s = as_ims(hIms);
    sch_set_tl_error(0, "%s", "");
    if (!s) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_computeEchograms: invalid handle", 0); return; }
    /* exactly one of the two must be given (saf_reverb.c:196-197) */
    if (!((maxN < 0) != (maxTime_s < 0.0f)) || (maxN < 0 && !(maxTime_s > 0.0f))) {
        prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_computeEchograms: one of maxN / maxTime_s must be >= 0 (> 0) and the other negative", 0);
        return;
    }
    const int mode = (maxTime_s > 0.0f) ? 0 : 1;
    for (int r = 0; r < IMS_MAX_REC; r++)
        for (int k = 0; k < IMS_MAX_SRC; k++) {
            if (s->recs[r].id == IMS_UNASSIGNED || s->srcs[k].id == IMS_UNASSIGNED) continue;
            ims_pair* p = &s->pair[r][k];
            /* :224-231 -- a changed target length / order forces a refresh (in T mode the reference compares a distance
             * with a time, so it always refreshes; the result is the same) */
            if (!p->haveParams || p->mode != mode || (mode == 0 ? p->maxTime != maxTime_s : p->maxN != maxN)) p->refreshEcho = 1;
            p->haveParams = 1; p->mode = mode; p->maxN = maxN; p->maxTime = maxTime_s;
            if (p->refreshEcho) { p->refreshEcho = 0; p->refreshRIR = 1; }
        }
}

/* reflection coefficient product along one axis for `order` reflections (saf_reverb_internal.c:601-627) */
static float wall_product(float r0, float r1, int order)
{
    const float a = (float)abs(order);
    if ((order % 2) == 0) return powf(r0, a / 2.0f) * powf(r1, a / 2.0f);
    if (order > 0)        return powf(r0, ceilf((float)order / 2.0f)) * powf(r1, floorf((float)order / 2.0f));
    return powf(r0, floorf(a / 2.0f)) * powf(r1, ceilf(a / 2.0f));
}

void safconv_ims_shoebox_renderRIRs(void* hIms, int fractionalDelaysFLAG)
{
    safconv_ims* s = as_ims(hIms);
    sch_set_tl_error(0, "%s", "");
    if (!s) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_renderRIRs: invalid handle", 0); return; }
    if (fractionalDelaysFLAG) { prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_renderRIRs: fractional delays are not implemented (nor in the reference, saf_reverb_internal.c:660-663)", 0); return; }
    int rc = SAFCONV_OK;
    int nP = 0, mode = -1, maxN = 0, maxOrder = 0;
    float maxTime = 0.0f;
    int idxR[IMS_MAX_REC * IMS_MAX_SRC], idxS[IMS_MAX_REC * IMS_MAX_SRC];
    for (int r = 0; r < IMS_MAX_REC; r++)
        for (int k = 0; k < IMS_MAX_SRC; k++) {
            ims_pair* p = &s->pair[r][k];
            if (s->recs[r].id == IMS_UNASSIGNED || s->srcs[k].id == IMS_UNASSIGNED || !p->refreshRIR || !p->haveParams) continue;
            if (mode < 0) { mode = p->mode; maxN = p->maxN; maxTime = p->maxTime; }
            idxR[nP] = r; idxS[nP] = k; nP++;
        }
    if (!nP) return;
    scdev_set_device(s->device);

    ScpImsPair* hp = (ScpImsPair*)calloc((size_t)nP, sizeof *hp);
    unsigned int* stats = (unsigned int*)calloc((size_t)nP * 2, sizeof *stats);
    void *d_pairs = NULL, *d_stats = NULL, *d_ptrs = NULL;
    float* d_abs = NULL; float* absTab = NULL;
    double* d_acc = NULL;
    float** rirPtrs = (float**)calloc((size_t)nP, sizeof *rirPtrs);
    if (!hp || !stats || !rirPtrs) { rc = prod_fail(SAFCONV_ERR_NOMEM, "out of host memory", 0); goto done; }

    /* lattice (saf_reverb_internal.c:299-303 / :426) */
    int Nx, Ny, Nz; float dmax = 0.0f;
    if (mode == 0) {
        dmax = maxTime * s->c_ms;
        Nx = (int)(dmax / s->room[0] + 1.0f); Ny = (int)(dmax / s->room[1] + 1.0f); Nz = (int)(dmax / s->room[2] + 1.0f);
    } else Nx = Ny = Nz = maxN;
    const long long lengthVec = (long long)(2 * Nx + 1) * (2 * Ny + 1) * (2 * Nz + 1);
    for (int i = 0; i < nP; i++) {
        const ims_obj* rec = &s->recs[idxR[i]]; const ims_obj* src = &s->srcs[idxS[i]];
        ScpImsPair* q = &hp[i];
        /* y is flipped (saf_reverb.c:206-212), then the origin moves to the room centre (saf_reverb_internal.c:288-294) */
        const float ry = s->room[1] - rec->pos[1], sy = s->room[1] - src->pos[1];
        q->room[0] = s->room[0]; q->room[1] = s->room[1]; q->room[2] = s->room[2];
        q->so[0] = src->pos[0] - s->room[0] / 2.0f; q->so[1] = s->room[1] / 2.0f - sy; q->so[2] = src->pos[2] - s->room[2] / 2.0f;
        q->ro[0] = rec->pos[0] - s->room[0] / 2.0f; q->ro[1] = s->room[1] / 2.0f - ry; q->ro[2] = rec->pos[2] - s->room[2] / 2.0f;
        q->c_ms = s->c_ms; q->fs = s->fs; q->dmax = dmax; q->mode = mode;
        q->Nx = Nx; q->Ny = Ny; q->Nz = Nz; q->lengthVec = lengthVec;
        q->nSH = rec->nCh; q->order = (int)(sqrt((double)rec->nCh) + 0.5) - 1;
        if (q->order > maxOrder) maxOrder = q->order;
    }
    /* wall-reflection tables [axis][band][2 Nmax + 1] */
    int Nmax = Nx > Ny ? Nx : Ny; if (Nz > Nmax) Nmax = Nz;
    const int maxW = 2 * Nmax + 1;
    absTab = (float*)calloc((size_t)3 * s->nBands * maxW, sizeof(float));
    if (!absTab) { rc = prod_fail(SAFCONV_ERR_NOMEM, "out of host memory", 0); goto done; }
    for (int ax = 0; ax < 3; ax++) {
        const int N = ax == 0 ? Nx : ax == 1 ? Ny : Nz;
        for (int b = 0; b < s->nBands; b++) {
            const float r0 = sqrtf(1.0f - s->abs_wall[b * 6 + 2 * ax]), r1 = sqrtf(1.0f - s->abs_wall[b * 6 + 2 * ax + 1]);
            for (int o = -N; o <= N; o++) absTab[((size_t)ax * s->nBands + b) * maxW + o + N] = wall_product(r0, r1, o);
        }
    }
    int e = scdev_malloc(&d_pairs, sizeof(ScpImsPair) * (size_t)nP);
    if (!e) e = scdev_malloc(&d_stats, sizeof(unsigned int) * 2 * (size_t)nP);
    if (!e) e = scdev_malloc((void**)&d_abs, sizeof(float) * 3 * (size_t)s->nBands * maxW);
    if (e) { rc = prod_fail(SAFCONV_ERR_NOMEM, "ims_shoebox_renderRIRs: device allocation", e); goto done; }
    PROD_TRY(scdev_memcpy_h2d_async(d_pairs, hp, sizeof(ScpImsPair) * (size_t)nP, s->stream), "pair upload");
    PROD_TRY(scdev_memcpy_h2d_async(d_abs, absTab, sizeof(float) * 3 * (size_t)s->nBands * maxW, s->stream), "absorption table upload");
    /* pass 1: images per pair and the latest arrival -> RIR lengths */
    PROD_TRY(scdev_ims_count(d_pairs, nP, lengthVec, (unsigned int*)d_stats, s->smCount, s->stream), "image count");
    PROD_TRY(scdev_memcpy_d2h_async(stats, d_stats, sizeof(unsigned int) * 2 * (size_t)nP, s->stream), "image count download");
    PROD_TRY(scdev_stream_sync(s->stream), "image count");
    size_t total = 0;
    int maxWindows = 0, accDoubles = 0;
    for (int i = 0; i < nP; i++) {
        float dLast; memcpy(&dLast, &stats[2 * i + 1], sizeof dLast);
        if (stats[2 * i] == 0) { rc = prod_fail(SAFCONV_ERR_ARG, "ims_shoebox_renderRIRs: an echogram is empty (maxTime_s shorter than the direct path)", 0); goto done; }
        hp[i].len = scp_ims_length(&hp[i], dLast);
        hp[i].accOff = (long long)total;
        total += (size_t)hp[i].nSH * (size_t)hp[i].len;
        /* taps per window: the window's fp64 taps of all channels fit in 48 KB of shared memory */
        int tw = (6144 / hp[i].nSH) & ~31;
        if (tw < 32) tw = 32;
        if (tw > 1024) tw = 1024;
        hp[i].tw = tw;
        const int nw = (hp[i].len + tw - 1) / tw;
        if (nw > maxWindows) maxWindows = nw;
        if (hp[i].nSH * tw > accDoubles) accDoubles = hp[i].nSH * tw;
    }
    for (int i = 0; i < nP; i++) {
        ims_pair* p = &s->pair[idxR[i]][idxS[i]];
        scdev_free(p->d_rir); p->d_rir = NULL;
        free(p->h_rir); p->h_rir = NULL;
        e = scdev_malloc((void**)&p->d_rir, sizeof(float) * (size_t)hp[i].nSH * (size_t)hp[i].len);
        if (e) { rc = prod_fail(SAFCONV_ERR_NOMEM, "ims_shoebox_renderRIRs: RIR allocation", e); goto done; }
        rirPtrs[i] = p->d_rir;
    }
    PROD_TRY(scdev_memcpy_h2d_async(d_pairs, hp, sizeof(ScpImsPair) * (size_t)nP, s->stream), "pair upload");
    /* Two render kernels.  The lattice scan with fp64 atomics in HBM is the faster one on B200 (measured: 2.95 vs 6.9 ms for
     * 6.4 M images x 64 channels) but needs an fp64 accumulator of the whole bank (8 bytes per tap: 3.1 GB for the 64 x 64 x
     * 96 000 bank); the windowed kernel needs nothing beyond the fp32 RIRs.  So: the scan while its accumulator fits in a
     * quarter of the free device memory, the windows beyond; SAFCONV_IMS_WINDOWS=0 / 1 forces one.  (A lattice row must
     * fit the window kernel's hit queue: 2 Nx + 1 <= 2048.) */
    int useWindows = sch_env_int("SAFCONV_IMS_WINDOWS", -1, -1, 1);
    if (useWindows < 0) {
        size_t freeB = 0;
        useWindows = (scdev_mem_free_bytes(&freeB) == 0 && (double)total * 8.0 > 0.25 * (double)freeB) ? 1 : 0;
    }
    if (useWindows && 2 * Nx + 1 <= 2048) {
        /* pass 2, windowed: one CTA per (pair, window of taps), taps accumulated in shared memory */
        e = scdev_malloc(&d_ptrs, sizeof(float*) * (size_t)nP);
        if (e) { rc = prod_fail(SAFCONV_ERR_NOMEM, "ims_shoebox_renderRIRs: pointer table", e); goto done; }
        PROD_TRY(scdev_memcpy_h2d_async(d_ptrs, rirPtrs, sizeof(float*) * (size_t)nP, s->stream), "pointer table upload");
        PROD_TRY(scdev_ims_render_windows(d_pairs, nP, maxWindows, accDoubles, maxOrder, d_abs, s->nBands, maxW, s->d_norms,
                                          (float* const*)d_ptrs, s->stream), "image render (windows)");
    } else {
        /* pass 2, global atomics: every lattice point into fp64 taps in HBM, then fp64 -> fp32 per pair */
        e = scdev_malloc((void**)&d_acc, sizeof(double) * total);
        if (e) { rc = prod_fail(SAFCONV_ERR_NOMEM, "ims_shoebox_renderRIRs: tap accumulator", e); goto done; }
        PROD_TRY(scdev_ims_render(d_pairs, nP, lengthVec, maxOrder, d_abs, s->nBands, maxW, s->d_norms, d_acc, total, s->smCount, s->stream), "image render");
        for (int i = 0; i < nP; i++)
            PROD_TRY(scdev_ims_finish(d_acc + hp[i].accOff, rirPtrs[i], (size_t)hp[i].nSH * (size_t)hp[i].len, s->stream), "RIR conversion");
    }
    for (int i = 0; i < nP; i++) {
        ims_pair* p = &s->pair[idxR[i]][idxS[i]];
        p->len = hp[i].len; p->nCh = hp[i].nSH; p->nImages = (int)stats[2 * i];
        p->refreshRIR = 0;
    }
    PROD_TRY(scdev_stream_sync(s->stream), "image render");
done:
    if (rc && rirPtrs)           /* a failed batch leaves no half-rendered RIRs behind */
        for (int i = 0; i < nP; i++)
            if (rirPtrs[i]) { ims_pair* p = &s->pair[idxR[i]][idxS[i]]; scdev_free(p->d_rir); p->d_rir = NULL; p->len = 0; }
    free(hp); free(stats); free(absTab); free(rirPtrs);
    scdev_free(d_pairs); scdev_free(d_stats); scdev_free(d_abs); scdev_free(d_acc); scdev_free(d_ptrs);
}

/* ---- accessors (the reference keeps the RIRs inside its handle, ims_scene_data::rirs, without a getter) ---------- */
static ims_pair* rendered_pair(safconv_ims* s, int receiverID, int sourceID)
{
    const int r = s ? find_obj(s->recs, IMS_MAX_REC, receiverID) : -1;
    const int k = s ? find_obj(s->srcs, IMS_MAX_SRC, sourceID) : -1;
    if (r < 0 || k < 0) { prod_fail(SAFCONV_ERR_ARG, "ims: invalid receiverID / sourceID", 0); return NULL; }
    ims_pair* p = &s->pair[r][k];
    if (!p->d_rir) { prod_fail(SAFCONV_ERR_ARG, "ims: no RIR rendered for this pair yet (computeEchograms + renderRIRs first)", 0); return NULL; }
    return p;
}

int safconv_ims_get_rir_device(void* hIms, int receiverID, int sourceID, const float** d_data, int* length, int* nChannels)
{
    ims_pair* p = rendered_pair(as_ims(hIms), receiverID, sourceID);
    if (!p) return SAFCONV_ERR_ARG;
    if (d_data) *d_data = p->d_rir;
    if (length) *length = p->len;
    if (nChannels) *nChannels = p->nCh;
    return SAFCONV_OK;
}

int safconv_ims_get_rir(void* hIms, int receiverID, int sourceID, const float** data, int* length, int* nChannels)
{
    safconv_ims* s = as_ims(hIms);
    ims_pair* p = rendered_pair(s, receiverID, sourceID);
    if (!p) return SAFCONV_ERR_ARG;
    if (!p->h_rir) {
        const size_t bytes = sizeof(float) * (size_t)p->nCh * (size_t)p->len;
        p->h_rir = (float*)malloc(bytes);
        if (!p->h_rir) return prod_fail(SAFCONV_ERR_NOMEM, "out of host memory", 0);
        scdev_set_device(s->device);
        int e = scdev_memcpy_d2h_async(p->h_rir, p->d_rir, bytes, s->stream);
        if (!e) e = scdev_stream_sync(s->stream);
        if (e) { free(p->h_rir); p->h_rir = NULL; return prod_fail(SAFCONV_ERR_CUDA, "RIR download", e); }
    }
    if (data) *data = p->h_rir;
    if (length) *length = p->len;
    if (nChannels) *nChannels = p->nCh;
    return SAFCONV_OK;
}

int safconv_ims_get_num_images(void* hIms, int receiverID, int sourceID)
{
    ims_pair* p = rendered_pair(as_ims(hIms), receiverID, sourceID);
    return p ? p->nImages : -1;
}

/* The rendered RIRs of one receiver as a matrix convolver: input channel = source (active sources in slot order),
 * output channel = SH channel of the receiver, length_h = the longest RIR (shorter ones zero-padded).  The filter bank
 * is assembled and transformed on the device. */
int safconv_ims_create_matrixConv(void* hIms, int receiverID, int hopSize, void** const phMC)
{
    sch_set_tl_error(0, "%s", "");
    safconv_ims* s = as_ims(hIms);
    if (!phMC) return prod_fail(SAFCONV_ERR_ARG, "ims_create_matrixConv: phMC is NULL", 0);
    *phMC = NULL;
    const int r = s ? find_obj(s->recs, IMS_MAX_REC, receiverID) : -1;
    if (r < 0) return prod_fail(SAFCONV_ERR_ARG, "ims_create_matrixConv: invalid handle / receiverID", 0);
    const float* ptrs[IMS_MAX_SRC]; int lens[IMS_MAX_SRC];
    int nSrc = 0, L = 0;
    const int nCh = s->recs[r].nCh;
    for (int k = 0; k < IMS_MAX_SRC; k++) {
        if (s->srcs[k].id == IMS_UNASSIGNED) continue;
        ims_pair* p = &s->pair[r][k];
        if (!p->d_rir || p->nCh != nCh) return prod_fail(SAFCONV_ERR_ARG, "ims_create_matrixConv: a source of this receiver has no rendered RIR", 0);
        ptrs[nSrc] = p->d_rir; lens[nSrc] = p->len; if (p->len > L) L = p->len;
        nSrc++;
    }
    if (!nSrc) return prod_fail(SAFCONV_ERR_ARG, "ims_create_matrixConv: no sources", 0);
    int rc = SAFCONV_OK;
    void *d_ptrs = NULL, *d_lens = NULL; float* d_H = NULL;
    scdev_set_device(s->device);
    int e = scdev_malloc(&d_ptrs, sizeof ptrs[0] * (size_t)nSrc);
    if (!e) e = scdev_malloc(&d_lens, sizeof(int) * (size_t)nSrc);
    if (!e) e = scdev_malloc((void**)&d_H, sizeof(float) * (size_t)nCh * nSrc * L);
    if (e) { rc = prod_fail(SAFCONV_ERR_NOMEM, "ims_create_matrixConv: device allocation", e); goto done; }
    PROD_TRY(scdev_memcpy_h2d_async(d_ptrs, ptrs, sizeof ptrs[0] * (size_t)nSrc, s->stream), "pointer upload");
    PROD_TRY(scdev_memcpy_h2d_async(d_lens, lens, sizeof(int) * (size_t)nSrc, s->stream), "length upload");
    PROD_TRY(scdev_ims_bank((const float* const*)d_ptrs, (const int*)d_lens, nSrc, nCh, L, d_H, s->stream), "filter bank assembly");
    PROD_TRY(scdev_stream_sync(s->stream), "filter bank assembly");
    {
        const int prev = sch_thread_device();
        sch_set_thread_device(s->device);
        safconv_matrixConv_create_device(phMC, hopSize, d_H, L, nSrc, nCh);
        sch_set_thread_device(prev);
        if (!*phMC) rc = safconv_last_error(NULL) ? safconv_last_error(NULL) : SAFCONV_ERR_CUDA;
    }
done:
    scdev_free(d_ptrs); scdev_free(d_lens); scdev_free(d_H);
    return rc;
}

/* the reference's names (weak: a host that links the reference's saf_reverb.c keeps its own simulator) */
__attribute__((weak)) void ims_shoebox_create(void** phIms, float roomDimensions[3], float* abs_wall, float lowestOctaveBand, int nOctBands, float c_ms, float fs)
{ safconv_ims_shoebox_create(phIms, roomDimensions, abs_wall, lowestOctaveBand, nOctBands, c_ms, fs); }
__attribute__((weak)) void ims_shoebox_destroy(void** phIms) { safconv_ims_shoebox_destroy(phIms); }
__attribute__((weak)) void ims_shoebox_computeEchograms(void* hIms, int maxN, float maxTime_s) { safconv_ims_shoebox_computeEchograms(hIms, maxN, maxTime_s); }
__attribute__((weak)) void ims_shoebox_renderRIRs(void* hIms, int fractionalDelaysFLAG) { safconv_ims_shoebox_renderRIRs(hIms, fractionalDelaysFLAG); }
__attribute__((weak)) void ims_shoebox_setRoomDimensions(void* hIms, float new_roomDimensions[3]) { safconv_ims_shoebox_setRoomDimensions(hIms, new_roomDimensions); }
__attribute__((weak)) void ims_shoebox_setWallAbsCoeffs(void* hIms, float* abs_wall) { safconv_ims_shoebox_setWallAbsCoeffs(hIms, abs_wall); }
__attribute__((weak)) int  ims_shoebox_addSource(void* hIms, float position_xyz[3], float** pSrc_sig) { return safconv_ims_shoebox_addSource(hIms, position_xyz, pSrc_sig); }
__attribute__((weak)) int  ims_shoebox_addReceiverSH(void* hIms, int sh_order, float position_xyz[3], float*** pSH_sigs) { return safconv_ims_shoebox_addReceiverSH(hIms, sh_order, position_xyz, pSH_sigs); }
__attribute__((weak)) void ims_shoebox_updateSource(void* hIms, int sourceID, float position_xyz[3]) { safconv_ims_shoebox_updateSource(hIms, sourceID, position_xyz); }
__attribute__((weak)) void ims_shoebox_updateReceiver(void* hIms, int receiverID, float position_xyz[3]) { safconv_ims_shoebox_updateReceiver(hIms, receiverID, position_xyz); }
__attribute__((weak)) void ims_shoebox_removeSource(void* hIms, int sourceID) { safconv_ims_shoebox_removeSource(hIms, sourceID); }
__attribute__((weak)) void ims_shoebox_removeReceiver(void* hIms, int receiverID) { safconv_ims_shoebox_removeReceiver(hIms, receiverID); }
