/*
 * safconv_rfft.c -- C host layer of the general-size real FFT: the reference's saf_rfft API as a drop-in
 *
 *   /root/reference/framework/modules/saf_utilities/saf_utility_fft.h:240-276   saf_rfft_create / _destroy / _forward /
 *   _backward (implementation saf_utility_fft.c:531-753; default backend KissFFT: kiss_fftr.c:69-161, kiss_fft.c:93-331)
 *
 * Any even N >= 2 (saf_utility_fft.c:542), forward unscaled, backward scaled by 1/N (.c:749-752), imaginary parts of
 * the DC and Nyquist bins ignored by the backward transform (kiss_fftr.c:137-138).  A handle is a RESIDENT plan: the
 * factorisation of N/2 (4s, 2s, 3s, 5s, then the remaining primes -- the order of kf_factor, kiss_fft.c:310-331), the
 * twiddle tables (evaluated in double, stored as float, like kiss_fft.c:358-364 / kiss_fftr.c:59-65) on the device,
 * device work arrays, page-locked staging and a stream; a transform call allocates and uploads nothing.  No CUDA code
 * here (kernels: safconv_gfft.cu); no CPU compute path: without a device create() yields a NULL handle + error string.
 */
#include "safconv_host_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SAFCONV_MAGIC_RFFT 0x5AFC0FF7u

typedef struct safconv_rfft {
    uint32_t magic;
    int err;
    char errmsg[256];
    int device;
    scdev_gfft_plan pl;
    void* stream;
    int   capBatch;                 /* transforms the device / staging buffers hold */
    float *d_td, *d_fd;             /* device [capBatch][N] real, [capBatch][N/2+1] complex */
    float *h_td, *h_fd;             /* page-locked staging of the same sizes */
} safconv_rfft;

static safconv_rfft* as_rfft(void* p)
{
    safconv_rfft* h = (safconv_rfft*)p;
    return (h && h->magic == SAFCONV_MAGIC_RFFT) ? h : NULL;
}

static int r_fail(safconv_rfft* h, int code, const char* what, int cudaErr)
{
    char buf[256];
    if (cudaErr) snprintf(buf, sizeof buf, "%s: %s", what, scdev_error_string(cudaErr));
    else         snprintf(buf, sizeof buf, "%s", what);
    if (h) { h->err = code; snprintf(h->errmsg, sizeof h->errmsg, "%s", buf); }
    sch_set_tl_error(code, "%s", buf);
    return code;
}

/* radices in pass order: kf_factor's order (kiss_fft.c:310-331).  Returns the count, 0 if more than `cap`. */
int safconv_debug_fft_factors(int M, int* fac, int cap)
{
    int n = M, nf = 0, p = 4;
    if (M == 1) { if (cap < 1) return 0; fac[0] = 1; return 1; }
    const double floor_sqrt = floor(sqrt((double)n));
    do {
        while (n % p) {
            switch (p) {
                case 4: p = 2; break;
                case 2: p = 3; break;
                default: p += 2; break;
            }
            if (p > floor_sqrt) p = n;
        }
        n /= p;
        if (nf >= cap) return 0;
        fac[nf++] = p;
    } while (n > 1);
    return nf;
}

/* plan of an N-point real FFT: radices of N/2 + the two twiddle tables on the current device.  Returns 0, a CUDA error
 * code, or -1 (host memory / more than SC_GFFT_MAX_FACTORS prime factors). */
int scr_plan_init(scdev_gfft_plan* pl, int N, void* stream)
{
    const int M = N / 2;
    memset(pl, 0, sizeof *pl);
    pl->N = N; pl->M = M;
    pl->nf = safconv_debug_fft_factors(M, pl->fac, SC_GFFT_MAX_FACTORS);
    if (pl->nf < 1) return -1;
    /* tables: W_M^e (e < M) and W_N^k (k <= M/2), evaluated in double like kiss_fft.c:358-364 / kiss_fftr.c:59-65 */
    const size_t nS = (size_t)M / 2 + 1;
    float* t = (float*)malloc(sizeof(float) * 2 * ((size_t)M + nS));
    if (!t) return -1;
    const double pi = 3.141592653589793238462643383279502884;
    for (int i = 0; i < M; i++) {
        const double ph = -2.0 * pi * (double)i / (double)M;
        t[2 * i] = (float)cos(ph); t[2 * i + 1] = (float)sin(ph);
    }
    float* s = t + 2 * (size_t)M;
    for (size_t k = 0; k < nS; k++) {
        const double ph = -2.0 * pi * (double)k / (double)N;
        s[2 * k] = (float)cos(ph); s[2 * k + 1] = (float)sin(ph);
    }
    int e = scdev_malloc(&pl->tw, sizeof(float) * 2 * (size_t)M);
    if (!e) e = scdev_malloc(&pl->stw, sizeof(float) * 2 * nS);
    if (!e) e = scdev_memcpy_h2d_sync(pl->tw, t, sizeof(float) * 2 * (size_t)M, stream);
    if (!e) e = scdev_memcpy_h2d_sync(pl->stw, s, sizeof(float) * 2 * nS, stream);
    free(t);
    return e;
}

/* work arrays of the multi-launch path for batches of up to nBatch transforms (no-op for one-CTA sizes) */
int scr_plan_reserve(scdev_gfft_plan* pl, int nBatch)
{
    if (scdev_gfft_smem_ok(pl->M)) { if (nBatch > pl->maxBatch) pl->maxBatch = nBatch; return 0; }
    if (nBatch <= pl->maxBatch) return 0;
    scdev_free(pl->w0); scdev_free(pl->w1);
    pl->w0 = pl->w1 = NULL; pl->maxBatch = 0;
    int e = scdev_malloc(&pl->w0, sizeof(float) * 2 * (size_t)pl->M * nBatch);
    if (!e) e = scdev_malloc(&pl->w1, sizeof(float) * 2 * (size_t)pl->M * nBatch);
    if (!e) pl->maxBatch = nBatch;
    return e;
}

void scr_plan_free(scdev_gfft_plan* pl)
{
    scdev_free(pl->tw); scdev_free(pl->stw); scdev_free(pl->w0); scdev_free(pl->w1);
    memset(pl, 0, sizeof *pl);
}

static void rfft_free(safconv_rfft* h)
{
    if (!h) return;
    if (h->device >= 0) scdev_set_device(h->device);
    if (h->stream) scdev_stream_sync(h->stream);
    scr_plan_free(&h->pl);
    scdev_free(h->d_td); scdev_free(h->d_fd);
    scdev_host_free(h->h_td); scdev_host_free(h->h_fd);
    scdev_stream_destroy(h->stream);
    h->magic = 0;
    free(h);
}

/* (re)size the device / staging buffers for `nBatch` transforms */
static int rfft_reserve(safconv_rfft* h, int nBatch)
{
    if (nBatch <= h->capBatch) return 0;
    const size_t N = (size_t)h->pl.N, M = (size_t)h->pl.M;
    scdev_stream_sync(h->stream);
    scdev_free(h->d_td); scdev_free(h->d_fd); scdev_host_free(h->h_td); scdev_host_free(h->h_fd);
    h->d_td = h->d_fd = h->h_td = h->h_fd = NULL; h->capBatch = 0;
    int e = scdev_malloc((void**)&h->d_td, sizeof(float) * N * nBatch);
    if (!e) e = scdev_malloc((void**)&h->d_fd, sizeof(float) * 2 * (M + 1) * nBatch);
    if (!e) e = scdev_host_alloc((void**)&h->h_td, sizeof(float) * N * nBatch);
    if (!e) e = scdev_host_alloc((void**)&h->h_fd, sizeof(float) * 2 * (M + 1) * nBatch);
    if (!e) e = scr_plan_reserve(&h->pl, nBatch);
    if (e) return r_fail(h, SAFCONV_ERR_NOMEM, "rfft buffers", e);
    h->capBatch = nBatch;
    return 0;
}

static safconv_rfft* rfft_create(int N, int nBatch)
{
    sch_set_tl_error(SAFCONV_OK, "%s", "");
    if (N < 2 || (N & 1) || nBatch < 1) {
        sch_set_tl_error(SAFCONV_ERR_ARG, "rfft: only even (non zero) FFT sizes are supported%s", "");     /* reference .c:542 */
        return NULL;
    }
    int ndev = 0;
    if (scdev_device_count(&ndev) != 0 || ndev < 1) {
        sch_set_tl_error(SAFCONV_ERR_NO_DEVICE, "no usable CUDA device%s (libsafconv_b200 has no CPU fallback)", "");
        return NULL;
    }
    safconv_rfft* h = (safconv_rfft*)calloc(1, sizeof *h);
    if (!h) { sch_set_tl_error(SAFCONV_ERR_NOMEM, "out of host memory%s", ""); return NULL; }
    h->magic = SAFCONV_MAGIC_RFFT;
    h->device = -1;
    int dev = sch_thread_device();
    int e = 0;
    if (dev < 0) e = scdev_get_device(&dev);
    if (!e) e = scdev_set_device(dev);
    if (e) { r_fail(h, SAFCONV_ERR_CUDA, "cudaSetDevice", e); rfft_free(h); return NULL; }
    h->device = dev;
    e = scdev_stream_create(&h->stream);
    if (e) { r_fail(h, SAFCONV_ERR_CUDA, "cudaStreamCreate", e); rfft_free(h); return NULL; }
    e = scr_plan_init(&h->pl, N, h->stream);
    if (e) { r_fail(h, e < 0 ? SAFCONV_ERR_NOMEM : SAFCONV_ERR_CUDA, "rfft plan (factors / twiddle tables)", e < 0 ? 0 : e); rfft_free(h); return NULL; }
    if (rfft_reserve(h, nBatch)) { rfft_free(h); return NULL; }
    return h;
}

/* nBatch transforms, host pointers (page-locked buffers are used directly, small ones even by the kernel itself) */
static int rfft_run(safconv_rfft* h, int dir, int nBatch, const float* in, float* out)
{
    h->err = SAFCONV_OK; h->errmsg[0] = 0;
    int e = scdev_set_device(h->device);
    if (e) return r_fail(h, SAFCONV_ERR_CUDA, "cudaSetDevice", e);
    if (rfft_reserve(h, nBatch)) return h->err;
    const size_t N = (size_t)h->pl.N, M = (size_t)h->pl.M;
    const size_t tdBytes = sizeof(float) * N * nBatch, fdBytes = sizeof(float) * 2 * (M + 1) * nBatch;
    const size_t inBytes = dir == 0 ? tdBytes : fdBytes, outBytes = dir == 0 ? fdBytes : tdBytes;
    const int pinned = scdev_is_pinned_host(in) && scdev_is_pinned_host((const char*)in + inBytes - 1)
                    && scdev_is_pinned_host(out) && scdev_is_pinned_host((const char*)out + outBytes - 1);
    float* h_in  = dir == 0 ? h->h_td : h->h_fd;
    float* h_out = dir == 0 ? h->h_fd : h->h_td;
    float* d_in  = dir == 0 ? h->d_td : h->d_fd;
    float* d_out = dir == 0 ? h->d_fd : h->d_td;
    const float* src = pinned ? in : h_in;
    float*       dst = pinned ? out : h_out;
    if (!pinned) memcpy(h_in, in, inBytes);
    if (scdev_gfft_smem_ok(h->pl.M) && inBytes <= (1u << 20) && outBytes <= (1u << 20)) {
        /* zero-copy: the kernel reads / writes the page-locked host buffers itself -- one launch, one synchronisation */
        e = scdev_gfft_run(&h->pl, dir, nBatch, src, dst, h->stream);
    } else {
        e = scdev_memcpy_h2d_async(d_in, src, inBytes, h->stream);
        if (!e) e = scdev_gfft_run(&h->pl, dir, nBatch, d_in, d_out, h->stream);
        if (!e) e = scdev_memcpy_d2h_async(dst, d_out, outBytes, h->stream);
    }
    if (!e) e = scdev_stream_sync(h->stream); else scdev_stream_sync(h->stream);
    if (e) return r_fail(h, SAFCONV_ERR_CUDA, dir == 0 ? "rfft forward" : "rfft backward", e);
    if (!pinned) memcpy(out, h_out, outBytes);
    return SAFCONV_OK;
}

/* ------------------------------------------------------------------------------------------ */
/*  drop-in API (reference saf_utility_fft.h:240-276)                                           */
/* ------------------------------------------------------------------------------------------ */
void saf_rfft_create(void** const phFFT, int N)
{
    if (!phFFT) return;
    *phFFT = rfft_create(N, 1);
}

void saf_rfft_destroy(void** const phFFT)
{
    if (!phFFT) return;
    safconv_rfft* h = as_rfft(*phFFT);
    if (h) rfft_free(h);
    *phFFT = NULL;
}

void saf_rfft_forward(void* const hFFT, float* inputTD, void* outputFD)
{
    safconv_rfft* h = as_rfft(hFFT);
    if (!h || !inputTD || !outputFD) return;
    rfft_run(h, 0, 1, inputTD, (float*)outputFD);
}

void saf_rfft_backward(void* const hFFT, void* inputFD, float* outputTD)
{
    safconv_rfft* h = as_rfft(hFFT);
    if (!h || !inputFD || !outputTD) return;
    rfft_run(h, 1, 1, (const float*)inputFD, outputTD);
}

/* ------------------------------------------------------------------------------------------ */
/*  extension: batches, error query                                                             */
/* ------------------------------------------------------------------------------------------ */
int safconv_rfft_batch(void* hFFT, int dir, int nBatch, const float* in, float* out)
{
    safconv_rfft* h = as_rfft(hFFT);
    if (!h || !in || !out || nBatch < 1 || (dir != 0 && dir != 1)) return SAFCONV_ERR_ARG;
    return rfft_run(h, dir, nBatch, in, out);
}

int safconv_rfft_last_error(void* hFFT)
{
    safconv_rfft* h = as_rfft(hFFT);
    return h ? h->err : safconv_last_error(NULL);
}

const char* safconv_rfft_last_error_string(void* hFFT)
{
    safconv_rfft* h = as_rfft(hFFT);
    return h ? h->errmsg : safconv_last_error_string(NULL);
}

/* number of passes + radices of a handle's plan (introspection / tests) */
int safconv_rfft_get_factors(void* hFFT, int* fac, int cap)
{
    safconv_rfft* h = as_rfft(hFFT);
    if (!h) return 0;
    for (int i = 0; i < h->pl.nf && fac && i < cap; i++) fac[i] = h->pl.fac[i];
    return h->pl.nf;
}

/* stateless helpers kept from round 1: any even N now */
static int rfft_oneshot(int N, int nBatch, const float* in, float* out, int dir)
{
    if (!in || !out) { sch_set_tl_error(SAFCONV_ERR_ARG, "rfft: NULL pointer%s", ""); return SAFCONV_ERR_ARG; }
    safconv_rfft* h = rfft_create(N, nBatch);
    if (!h) return safconv_last_error(NULL) ? safconv_last_error(NULL) : SAFCONV_ERR_CUDA;
    int rc = rfft_run(h, dir, nBatch, in, out);
    if (rc) { char keep[256]; snprintf(keep, sizeof keep, "%s", h->errmsg); rfft_free(h); sch_set_tl_error(rc, "%s", keep); return rc; }
    rfft_free(h);
    return SAFCONV_OK;
}

int safconv_rfft_forward(int N, int nBatch, const float* x, float* X)  { return rfft_oneshot(N, nBatch, x, X, 0); }
int safconv_rfft_backward(int N, int nBatch, const float* X, float* x) { return rfft_oneshot(N, nBatch, X, x, 1); }
