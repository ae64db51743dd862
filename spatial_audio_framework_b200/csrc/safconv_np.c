/*
 * safconv_np.c -- TRUE non-partitioned convolver modes (usePartFLAG = 0 with the reference's big-FFT semantics)
 *
 *   /root/reference/framework/modules/saf_utilities/saf_utility_matrixConv.c:71-96   (create), :174-207 (apply)  matrix
 *                                                                             :277-298 (create), :368-386 (apply)  multi
 *
 * fftSize = numOvrlpAddBlocks * hopSize with numOvrlpAddBlocks = ceil((hop + L - 1) / hop) -- generally NOT a power of two
 * (hop 256 / L 1024 -> 1280) --, nBins = fftSize/2 + 1, ONE forward FFT per input channel and block, per-bin products
 * with the filter spectra, inverse FFTs, and a fftSize-long overlap-add buffer per output that is shifted by one hop
 * every block.  The FFTs run on the general-size device FFT (safconv_gfft.cu, plans from safconv_rfft.c).
 *
 * By default usePartFLAG = 0 is still served by the partitioned engine (same causal linear convolution, lower latency
 * and far less memory for long filters); this engine is selected with SAFCONV_TRUE_MODE0=1 (environment) or
 * safconv_set_true_mode0(1) (process-wide, read by the next create).  If fftSize is odd (odd hop with an odd block count) the
 * reference itself cannot run (saf_rfft_create asserts an even size, saf_utility_fft.c:542): the partitioned engine
 * serves the handle.
 *
 * One re-design kept from the partitioned engine: the reference inverse-transforms every (output, input) product and
 * sums in time (.c:192-195: nIn inverse FFTs per output); the inverse FFT is linear, so the products are summed per bin
 * first and each output needs ONE inverse FFT.
 */
#include "safconv_host_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SAFCONV_MAGIC_NP 0x5AFC0B30u

typedef struct safconv_np {
    uint32_t magic;
    int err;
    char errmsg[256];
    int kind, device;
    int hop, len, nIn, nOut;          /* multi: nIn == nOut == nCH */
    int F, nBins, nBlocksOA;          /* fftSize, fftSize/2+1, numOvrlpAddBlocks */
    scdev_gfft_plan pl;
    void* stream;
    size_t inBytes, outBytes;
    float *d_in, *d_out, *h_in, *h_out;
    float *xpad;                      /* [nIn][F]                                        */
    float *X;                         /* [nIn][nBins] complex                            */
    float *Hf;                        /* matrix [nOut][nIn][nBins], multi [nCH][nBins]   */
    float *Z;                         /* [nOut][nBins] complex                           */
    float *z;                         /* [nOut][F]                                       */
    float *ov[2];                     /* overlap-add buffers [nOut][F], ping-pong        */
    int cur;
    int detectPinned;
} safconv_np;

static int g_true_mode0 = -1;     /* -1: not decided yet (environment) */

int safconv_set_true_mode0(int enable) { g_true_mode0 = enable ? 1 : 0; return SAFCONV_OK; }
int scn_enabled(void)
{
    if (g_true_mode0 < 0) g_true_mode0 = sch_env_int("SAFCONV_TRUE_MODE0", 0, 0, 1);
    return g_true_mode0;
}

static safconv_np* as_np(const void* p)
{
    const safconv_np* h = (const safconv_np*)p;
    return (h && h->magic == SAFCONV_MAGIC_NP) ? (safconv_np*)h : NULL;
}

int scn_is_np(const void* p) { return as_np(p) != NULL; }

static int n_fail(safconv_np* h, int code, const char* what, int cudaErr)
{
    char buf[256];
    if (cudaErr) snprintf(buf, sizeof buf, "%s: %s", what, scdev_error_string(cudaErr));
    else         snprintf(buf, sizeof buf, "%s", what);
    if (h) { h->err = code; snprintf(h->errmsg, sizeof h->errmsg, "%s", buf); }
    sch_set_tl_error(code, "%s", buf);
    return code;
}

static void np_free(safconv_np* h)
{
    if (!h) return;
    if (h->device >= 0) scdev_set_device(h->device);
    if (h->stream) scdev_stream_sync(h->stream);
    scr_plan_free(&h->pl);
    scdev_free(h->d_in); scdev_free(h->d_out); scdev_host_free(h->h_in); scdev_host_free(h->h_out);
    scdev_free(h->xpad); scdev_free(h->X); scdev_free(h->Hf); scdev_free(h->Z); scdev_free(h->z);
    scdev_free(h->ov[0]); scdev_free(h->ov[1]);
    scdev_stream_destroy(h->stream);
    h->magic = 0;
    free(h);
}

#define NP_TRY(call, what) do { int e__ = (call); if (e__) { n_fail(h, SAFCONV_ERR_CUDA, (what), e__); goto fail; } } while (0)

/* fftSize of the reference's non-partitioned mode (.c:73-75); 0 when it is odd (the reference cannot run there) */
int safconv_debug_mode0_fft_size(int hop, int len)
{
    const int nb = (int)(ceilf((float)(hop + len - 1) / (float)hop) + 0.1f);
    const long long F = (long long)nb * hop;
    if (F < 2 || (F & 1) || F > 0x3fffffff) return 0;
    return (int)F;
}

void* scn_create(int kind, int hop, const float* H, int len, int nIn, int nOut)
{
    sch_set_tl_error(SAFCONV_OK, "%s", "");
    if (hop < 1 || len < 1 || nIn < 1 || nOut < 1 || !H) {
        sch_set_tl_error(SAFCONV_ERR_ARG, "invalid argument%s (need hopSize >= 1, length_h >= 1, channels >= 1, H != NULL)", "");
        return NULL;
    }
    const int F = safconv_debug_mode0_fft_size(hop, len);
    if (!F) return NULL;                                        /* caller falls back to the partitioned engine */
    int ndev = 0;
    if (scdev_device_count(&ndev) != 0 || ndev < 1) {
        sch_set_tl_error(SAFCONV_ERR_NO_DEVICE, "no usable CUDA device%s (libsafconv_b200 has no CPU fallback)", "");
        return NULL;
    }
    safconv_np* h = (safconv_np*)calloc(1, sizeof *h);
    if (!h) { sch_set_tl_error(SAFCONV_ERR_NOMEM, "out of host memory%s", ""); return NULL; }
    h->magic = SAFCONV_MAGIC_NP;
    h->device = -1;
    h->kind = kind; h->hop = hop; h->len = len; h->nIn = nIn; h->nOut = nOut;
    h->F = F; h->nBins = F / 2 + 1; h->nBlocksOA = F / hop;
    h->detectPinned = 1;
    int dev = sch_thread_device();
    if (dev < 0) NP_TRY(scdev_get_device(&dev), "cudaGetDevice");
    NP_TRY(scdev_set_device(dev), "cudaSetDevice");
    h->device = dev;
    NP_TRY(scdev_stream_create(&h->stream), "cudaStreamCreate");
    {
        int e = scr_plan_init(&h->pl, F, h->stream);
        if (e) { n_fail(h, e < 0 ? SAFCONV_ERR_NOMEM : SAFCONV_ERR_CUDA, "FFT plan", e < 0 ? 0 : e); goto fail; }
    }
    const size_t Fz = (size_t)F, nB = (size_t)h->nBins;
    const size_t rowsH = (kind == SC_KIND_MATRIX) ? (size_t)nOut * nIn : (size_t)nOut;
    h->inBytes = sizeof(float) * (size_t)nIn * hop;
    h->outBytes = sizeof(float) * (size_t)nOut * hop;
    /* rows of H transformed per launch group at create (bounds the zero-padded staging array to ~256 MB) */
    size_t chunk = (size_t)(256u << 20) / (Fz * sizeof(float));
    if (chunk < 1) chunk = 1;
    if (chunk > rowsH) chunk = rowsH;
    if (chunk > 32768) chunk = 32768;
    size_t maxBatch = chunk;
    if ((size_t)nIn > maxBatch) maxBatch = (size_t)nIn;
    if ((size_t)nOut > maxBatch) maxBatch = (size_t)nOut;
    if (maxBatch > 65535) { n_fail(h, SAFCONV_ERR_ARG, "non-partitioned mode: more than 65535 channels", 0); goto fail; }
    NP_TRY(scr_plan_reserve(&h->pl, (int)maxBatch), "FFT work arrays");
    NP_TRY(scdev_malloc((void**)&h->d_in, h->inBytes), "input staging");
    NP_TRY(scdev_malloc((void**)&h->d_out, h->outBytes), "output staging");
    NP_TRY(scdev_host_alloc((void**)&h->h_in, h->inBytes), "pinned input staging");
    NP_TRY(scdev_host_alloc((void**)&h->h_out, h->outBytes), "pinned output staging");
    NP_TRY(scdev_malloc((void**)&h->xpad, sizeof(float) * Fz * (chunk > (size_t)nIn ? chunk : (size_t)nIn)), "zero-padded input");
    NP_TRY(scdev_malloc((void**)&h->X, sizeof(float) * 2 * nB * nIn), "input spectra");
    NP_TRY(scdev_malloc((void**)&h->Hf, sizeof(float) * 2 * nB * rowsH), "filter spectra");
    NP_TRY(scdev_malloc((void**)&h->Z, sizeof(float) * 2 * nB * nOut), "output spectra");
    NP_TRY(scdev_malloc((void**)&h->z, sizeof(float) * Fz * nOut), "output blocks");
    for (int i = 0; i < 2; i++) {
        NP_TRY(scdev_malloc((void**)&h->ov[i], sizeof(float) * Fz * nOut), "overlap-add buffer");
        NP_TRY(scdev_memset_async(h->ov[i], 0, sizeof(float) * Fz * nOut, h->stream), "overlap-add buffer");
    }
    /* filters: upload a chunk of rows, zero-pad to fftSize, forward FFT straight into H_f (.c:88-94) */
    {
        float* d_h = NULL;
        int e = scdev_malloc((void**)&d_h, sizeof(float) * chunk * (size_t)len);
        for (size_t r0 = 0; r0 < rowsH && !e; r0 += chunk) {
            const size_t n = (rowsH - r0 < chunk) ? rowsH - r0 : chunk;
            e = scdev_memcpy_h2d_sync(d_h, H + r0 * (size_t)len, sizeof(float) * n * (size_t)len, h->stream);
            if (!e) e = scdev_np_pad(d_h, h->xpad, (int)n, len, F, (size_t)len, h->stream);
            if (!e) e = scdev_gfft_run(&h->pl, 0, (int)n, h->xpad, h->Hf + 2 * nB * r0, h->stream);
            if (!e) e = scdev_stream_sync(h->stream);
        }
        scdev_free(d_h);
        if (e) { n_fail(h, SAFCONV_ERR_CUDA, "filter transform (non-partitioned)", e); goto fail; }
    }
    return h;
fail:
    np_free(h);
    return NULL;
}

void scn_destroy(void** pp)
{
    if (!pp) return;
    safconv_np* h = as_np(*pp);
    *pp = NULL;
    if (h) np_free(h);
}

void scn_apply(void* p, int kind, const float* in, float* out)
{
    safconv_np* h = as_np(p);
    if (!h || h->kind != kind || !in || !out) return;
    h->err = SAFCONV_OK; h->errmsg[0] = 0;
    int e = scdev_set_device(h->device);
    if (e) { n_fail(h, SAFCONV_ERR_CUDA, "cudaSetDevice", e); return; }
    const int direct = h->detectPinned && scdev_is_pinned_host(in) && scdev_is_pinned_host(out)
                    && scdev_is_pinned_host((const char*)in + h->inBytes - 1) && scdev_is_pinned_host((const char*)out + h->outBytes - 1);
    const float* src = direct ? in : h->h_in;
    float*       dst = direct ? out : h->h_out;
    if (!direct) memcpy(h->h_in, in, h->inBytes);
    /* blocks of up to 1 MB: the pad / overlap-add kernels read / write the page-locked host buffers themselves */
    const int zc = h->inBytes <= (1u << 20) && h->outBytes <= (1u << 20);
    const float* kin = zc ? src : h->d_in;
    float* kout = zc ? dst : h->d_out;
    const int multi = (h->kind == SC_KIND_MULTI);
    if (!zc) e = scdev_memcpy_h2d_async(h->d_in, src, h->inBytes, h->stream);
    if (!e) e = scdev_np_pad(kin, h->xpad, h->nIn, h->hop, h->F, (size_t)h->hop, h->stream);                 /* .c:177 / :371 */
    if (!e) e = scdev_gfft_run(&h->pl, 0, h->nIn, h->xpad, h->X, h->stream);                                 /* .c:178 / :372 */
    if (!e) e = scdev_np_mac(h->Hf, h->X, h->Z, h->nOut, h->nIn, h->nBins, multi, h->stream);                /* .c:186 / :378 */
    if (!e) e = scdev_gfft_run(&h->pl, 1, h->nOut, h->Z, h->z, h->stream);                                   /* .c:193 / :380 */
    if (!e) e = scdev_np_ola(h->z, h->ov[h->cur], h->ov[h->cur ^ 1], kout, h->nOut, h->hop, h->F, h->stream); /* .c:198-205 */
    if (!e && !zc) e = scdev_memcpy_d2h_async(dst, h->d_out, h->outBytes, h->stream);
    if (!e) e = scdev_stream_sync(h->stream); else scdev_stream_sync(h->stream);
    if (e) { n_fail(h, SAFCONV_ERR_CUDA, "apply (non-partitioned)", e); return; }
    h->cur ^= 1;
    if (!direct) memcpy(out, h->h_out, h->outBytes);
}

int scn_last_error(void* p) { safconv_np* h = as_np(p); return h ? h->err : SAFCONV_ERR_ARG; }
const char* scn_last_error_string(void* p) { safconv_np* h = as_np(p); return h ? h->errmsg : ""; }

int scn_reset_state(void* p)
{
    safconv_np* h = as_np(p);
    if (!h) return SAFCONV_ERR_ARG;
    scdev_set_device(h->device);
    int e = 0;
    for (int i = 0; i < 2 && !e; i++) e = scdev_memset_async(h->ov[i], 0, sizeof(float) * (size_t)h->F * h->nOut, h->stream);
    if (!e) e = scdev_stream_sync(h->stream);
    return e ? n_fail(h, SAFCONV_ERR_CUDA, "reset_state", e) : SAFCONV_OK;
}

int scn_get_info(void* p, safconv_info* info)
{
    safconv_np* h = as_np(p);
    if (!h || !info) return SAFCONV_ERR_ARG;
    memset(info, 0, sizeof *info);
    info->kind = h->kind; info->hopSize = h->hop; info->length_h = h->len;
    info->nCHin = h->nIn; info->nCHout = h->nOut; info->nOutLocal = h->nOut;
    info->fftSize = h->F; info->nBinsPacked = h->nBins; info->numFilterBlocks = 1;
    info->maxBatch = 1; info->device = h->device;
    const double rows = (h->kind == SC_KIND_MATRIX) ? (double)h->nOut * h->nIn : (double)h->nOut;
    info->bytesFilters = (size_t)(8.0 * h->nBins * rows);
    info->bytesDelayLine = 0;
    info->macAlgBytesPerBlock = 8.0 * h->nBins * (rows + h->nIn);
    info->algBytesPerBlock = info->macAlgBytesPerBlock + 4.0 * h->hop * (h->nIn + h->nOut) + 8.0 * (double)h->F * h->nOut;
    return SAFCONV_OK;
}
