/*
 * safconv_host.c -- C host layer of libsafconv_b200.so
 *
 * Implements the reference's convolver API
 *   /root/reference/framework/modules/saf_utilities/saf_utility_matrixConv.h:55-190
 * (saf_matrixConv_*, saf_multiConv_*, saf_TVConv_*) on top of the thin C-ABI CUDA
 * layer declared in safconv_dev.h.  This file contains no CUDA code: it validates
 * arguments, derives the execution plan (FFT size, partition count, MAC tiling and
 * split-K work distribution), owns the handle and its device buffers, and sequences
 * the per-block launches.  There is no CPU compute path: if the device layer fails,
 * the error is recorded and create() leaves *phMC == NULL.
 */
#include "safconv_host_internal.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static __thread int  tl_err = 0;
static __thread char tl_msg[256] = "";
static __thread int  tl_device = -1;
static __thread int  tl_filters_on_device = 0;   /* safconv_matrixConv_create_device: `chunks[0]` already is the device-resident bank */

static void set_tl_error(int code, const char* fmt, const char* detail)
{
    tl_err = code;
    snprintf(tl_msg, sizeof tl_msg, fmt, detail ? detail : "");
}

static int h_fail(safconv_handle* h, int code, const char* what, int cudaErr)
{
    char buf[256];
    if (cudaErr) snprintf(buf, sizeof buf, "%s: %s", what, scdev_error_string(cudaErr));
    else         snprintf(buf, sizeof buf, "%s", what);
    if (h) { h->err = code; snprintf(h->errmsg, sizeof h->errmsg, "%s", buf); }
    set_tl_error(code, "%s", buf);
    return code;
}

void sch_set_tl_error(int code, const char* fmt, const char* detail) { set_tl_error(code, fmt, detail); }
int  sch_fail(safconv_handle* h, int code, const char* what, int cudaErr) { return h_fail(h, code, what, cudaErr); }

#define DEV_TRY(h, call, what) do { int e__ = (call); if (e__) { h_fail((h), SAFCONV_ERR_CUDA, (what), e__); goto fail; } } while (0)

/* "last error" semantics: every apply / extension call starts with a clean slate, so one transient failure
 * (e.g. an out-of-range irIdx) does not make later, correct calls look failed */
static void h_clear(safconv_handle* h) { h->err = SAFCONV_OK; h->errmsg[0] = 0; }

static safconv_handle* as_handle(void* p)
{
    safconv_handle* h = (safconv_handle*)p;
    return (h && h->magic == SAFCONV_MAGIC) ? h : NULL;
}
safconv_handle* sch_as_handle(void* p) { return as_handle(p); }
static void res_stop(safconv_handle* h);
/* for every entry point except the host-pointer apply: a resident latency kernel owns the stream, stop it first */
static safconv_handle* as_handle_q(void* p)
{
    safconv_handle* h = as_handle(p);
    if (h) res_stop(h);
    return h;
}
int sch_thread_device(void) { return tl_device; }
void sch_set_thread_device(int device) { tl_device = device; }

/* ------------------------------------------------------------------------------------------ */
/*  planning                                                                                    */
/* ------------------------------------------------------------------------------------------ */

static inline double now_ns(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e9 + (double)ts.tv_nsec;
}

static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

/* FFT size: the reference uses exactly 2*hop (saf_utility_matrixConv.c:100); we need a power of two,
 * so N = nextpow2(max(64, 2*hop)).  For a power-of-two hop >= 32 this is the reference's size; for any
 * other hop the transform is longer than needed, which leaves the linear convolution unchanged
 * (hop + hop - 1 <= N) -- only the first 2*hop output samples are used. */
static void plan_fft(scdev_plan* pl, int hop, int len)
{
    int N = 64;
    while (N < 2 * hop) N <<= 1;
    pl->hop = hop; pl->len = len;
    pl->N = N; pl->M = N / 2; pl->logM = ilog2(pl->M);
    pl->P = (int)ceilf((float)len / (float)hop);           /* reference .c:102 */
    if (pl->P < 1) pl->P = 1;
    int t = pl->M / 4;
    if (t < 32) t = 32;
    if (t > 256) t = 256;
    pl->fftThreads = t;
}

/* MAC tiling + even split of all pipeline stages over `grid` CTAs */
static int env_int(const char* name, int dflt, int lo, int hi)
{
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    int x = atoi(v);
    return (x < lo || x > hi) ? dflt : x;
}

int sch_env_int(const char* name, int dflt, int lo, int hi) { return env_int(name, dflt, lo, hi); }

static void plan_mac(scdev_plan* pl, int smCount)
{
    const int nOut = pl->nOutLocal;
    /* tuning knobs (benchmarks only): pipeline depth and bytes of H per stage */
    pl->macStages     = env_int("SAFCONV_MAC_STAGES", SC_MAC_NSTAGES, 2, SC_MAC_MAX_STAGES);
    pl->macStageBytes = 1024 * env_int("SAFCONV_MAC_STAGE_KB", SC_STAGE_H_BYTES / 1024, 1, 96);
    pl->nKT  = pl->M / SC_BK;
    pl->nOT  = (nOut + SC_MAX_OT - 1) / SC_MAX_OT;
    pl->OTsz = (nOut + pl->nOT - 1) / pl->nOT;
    /* warp roles: WGo warp groups over outputs (R outputs each, held in registers so that one delay-line
     * load feeds R filters), WGk = 8 / WGo groups over the input rows of a stage.  Smallest WGo with R <= 8. */
    int wgo = 1;
    while ((pl->OTsz + wgo - 1) / wgo > 8) wgo <<= 1;
    pl->WGo = wgo;
    pl->R   = (pl->OTsz + wgo - 1) / wgo;
    const int R = pl->R;
    int sni = pl->macStageBytes / (pl->OTsz * SC_BK * 8);
    if (sni < 1) sni = 1;
    if (sni > SC_MAX_SNI) sni = SC_MAX_SNI;
    if (sni > pl->nIn) sni = pl->nIn;
    pl->SNI = sni;
    pl->SPU = (pl->nIn + sni - 1) / sni;
    int wgk = SC_MAC_CWARPS / pl->WGo;
    if (wgk < 1) wgk = 1;
    if (wgk > sni) wgk = sni;
    pl->WGk = wgk;
    pl->nGroups = pl->nOT * pl->nKT;
    pl->totalStages = (long long)pl->nGroups * pl->P * pl->SPU;
    /* blocks handed over together (safconv_apply_device_blocks) share one forward-FFT launch and one
     * inverse-FFT launch; the delay line then needs maxBatch extra ring slots (the FFTs of the whole batch
     * are written before the first MAC runs).  Keep the extra memory below ~256 MB. */
    {
        int mb = env_int("SAFCONV_MAX_BATCH", 32, 1, 256);
        const double slotBytes = (double)pl->M * pl->nIn * 8.0;
        const double ztBytes   = (double)nOut * pl->hop * 8.0;
        while (mb > 1 && mb * (slotBytes + ztBytes) > 256e6) mb >>= 1;
        pl->maxBatch = mb;
        pl->RS = pl->P + mb;
    }
    long long g = smCount;
    if (g > pl->totalStages) g = pl->totalStages;
    pl->macGrid = (int)g;
    pl->macSmemBytes = pl->macStages * (pl->SNI * pl->OTsz * SC_BK * 8 + pl->SNI * SC_BK * 8)
                     + SC_MAC_CWARPS * R * 32 * 8 + 2 * pl->macStages * 8;
}

/* split-K bookkeeping: CTA c streams stages [c*T/G, (c+1)*T/G); every (ot,kt) group it touches gets
 * one partial tile.  ctaBase[c] = first partial slot of CTA c; grpStart/grpList = CSR list of the
 * partial slots that K3 sums for each group.  Returns the number of slots, or -1 on malloc failure. */
static int build_split_tables_for(long long T, long long spg, int G, int nG, int** ctaBaseOut, int** grpStartOut, int** grpListOut);

static int build_split_tables(const scdev_plan* pl, int** ctaBaseOut, int** grpStartOut, int** grpListOut)
{
    return build_split_tables_for(pl->totalStages, (long long)pl->P * pl->SPU, pl->macGrid, pl->nGroups,
                                  ctaBaseOut, grpStartOut, grpListOut);
}

/* T stages in total, spg stages per group, split evenly over G CTAs */
static int build_split_tables_for(long long T, long long spg, int G, int nG, int** ctaBaseOut, int** grpStartOut, int** grpListOut)
{
    int* ctaBase = (int*)malloc(sizeof(int) * (size_t)(G + 1));
    int* cnt     = (int*)calloc((size_t)nG + 1, sizeof(int));
    if (!ctaBase || !cnt) { free(ctaBase); free(cnt); return -1; }
    int slots = 0;
    for (int c = 0; c < G; c++) {
        const long long s0 = T * c / G, s1 = T * (c + 1) / G;
        ctaBase[c] = slots;
        if (s1 > s0) {
            const long long g0 = s0 / spg, g1 = (s1 - 1) / spg;
            for (long long g = g0; g <= g1; g++) cnt[g]++;
            slots += (int)(g1 - g0 + 1);
        }
    }
    ctaBase[G] = slots;
    int* grpStart = (int*)malloc(sizeof(int) * (size_t)(nG + 1));
    int* grpList  = (int*)malloc(sizeof(int) * (size_t)(slots > 0 ? slots : 1));
    int* fill     = (int*)calloc((size_t)nG + 1, sizeof(int));
    if (!grpStart || !grpList || !fill) { free(ctaBase); free(cnt); free(grpStart); free(grpList); free(fill); return -1; }
    grpStart[0] = 0;
    for (int g = 0; g < nG; g++) grpStart[g + 1] = grpStart[g] + cnt[g];
    for (int c = 0; c < G; c++) {
        const long long s0 = T * c / G, s1 = T * (c + 1) / G;
        if (s1 <= s0) continue;
        const long long g0 = s0 / spg, g1 = (s1 - 1) / spg;
        for (long long g = g0; g <= g1; g++) grpList[grpStart[g] + fill[g]++] = ctaBase[c] + (int)(g - g0);
    }
    free(cnt); free(fill);
    *ctaBaseOut = ctaBase; *grpStartOut = grpStart; *grpListOut = grpList;
    return slots;
}

int safconv_debug_plan_size(void) { return (int)sizeof(scdev_plan); }

/* exported for the host-logic unit tests (tests/test_plan.py): fills the plan exactly as create() does */
int safconv_debug_plan(int kind, int hop, int len, int nIn, int nOutLocal, int smCount, scdev_plan* out,
                       int* ctaBase /* >= smCount+1 */, int* grpStart, int* grpList, int cap)
{
    scdev_plan pl;
    memset(&pl, 0, sizeof pl);
    pl.kind = kind; pl.nIn = nIn; pl.nOutLocal = nOutLocal;
    plan_fft(&pl, hop, len);
    if (kind == SC_KIND_MATRIX) {
        plan_mac(&pl, smCount);
        int *a = NULL, *b = NULL, *c = NULL;
        const int slots = build_split_tables(&pl, &a, &b, &c);
        if (slots < 0) return -1;
        pl.nSlots = slots;
        if (ctaBase && grpStart && grpList && slots <= cap && pl.nGroups + 1 <= cap && pl.macGrid + 1 <= cap) {
            memcpy(ctaBase, a, sizeof(int) * (size_t)(pl.macGrid + 1));
            memcpy(grpStart, b, sizeof(int) * (size_t)(pl.nGroups + 1));
            memcpy(grpList, c, sizeof(int) * (size_t)slots);
        }
        free(a); free(b); free(c);
    }
    *out = pl;
    return 0;
}

/* host-logic unit tests: the split tables of a MAC pass over partitions [pLo, pLo + nP) exactly as make_pass() builds
 * them for the look-ahead apply (tail pass (1, P-1), head pass (0, 1)).  Returns the number of partial slots or < 0. */
int safconv_debug_pass_tables(int hop, int len, int nIn, int nOutLocal, int smCount, int pLo, int nP,
                              long long* totalStages, int* grid, int* ctaBase, int* grpStart, int* grpList, int cap)
{
    scdev_plan pl;
    memset(&pl, 0, sizeof pl);
    pl.kind = SC_KIND_MATRIX; pl.nIn = nIn; pl.nOutLocal = nOutLocal;
    plan_fft(&pl, hop, len);
    plan_mac(&pl, smCount);
    if (pLo < 0 || nP < 1 || pLo + nP > pl.P) return -2;
    const long long T = (long long)pl.nGroups * nP * pl.SPU;
    const int G = (T < smCount) ? (int)T : smCount;
    int *a = NULL, *b = NULL, *c = NULL;
    const int slots = build_split_tables_for(T, (long long)nP * pl.SPU, G, pl.nGroups, &a, &b, &c);
    if (slots < 0) return -1;
    if (slots <= cap && pl.nGroups + 1 <= cap && G + 1 <= cap) {
        memcpy(ctaBase, a, sizeof(int) * (size_t)(G + 1));
        memcpy(grpStart, b, sizeof(int) * (size_t)(pl.nGroups + 1));
        memcpy(grpList, c, sizeof(int) * (size_t)slots);
    }
    free(a); free(b); free(c);
    *totalStages = T; *grid = G;
    return slots;
}

/* ------------------------------------------------------------------------------------------ */
/*  create / destroy                                                                            */
/* ------------------------------------------------------------------------------------------ */

static void handle_free(safconv_handle* h)
{
    if (!h) return;
    if (h->hostTrace && h->htN)
        fprintf(stderr, "safconv host trace: %u zero-copy applies, us per call: checks %.2f, launch %.2f, wait %.2f\n", h->htN,
                h->htAcc[0] / h->htN * 1e-3, h->htAcc[1] / h->htN * 1e-3, h->htAcc[2] / h->htN * 1e-3);
    if (h->device >= 0) scdev_set_device(h->device);
    if (h->resActive && h->mailbox) { h->mailbox->bell = (unsigned long long)SC_RES_EXIT; __atomic_thread_fence(__ATOMIC_SEQ_CST); }   /* resident kernel: leave */
    if (h->stream) scdev_stream_sync(h->stream);
    scdev_graph_destroy(h->graphExec);
    if (h->evRing) { for (int i = 0; i < 4 * h->timingCap; i++) scdev_event_destroy(h->evRing[i]); free(h->evRing); }
    free(h->evBlocks);
    scdev_offline_free(&h->off);
    for (int i = 0; i < 4; i++) scdev_event_destroy(h->offEv[i]);
    for (int i = 0; i < 6; i++) scdev_event_destroy(h->offPipeEv[i]);
    scdev_stream_destroy(h->offIn); scdev_stream_destroy(h->offOut);
    for (int i = 0; i < 2; i++) { scdev_free(h->offStageIn[i]); scdev_free(h->offStageOut[i]); }
    scdev_free(h->b.tw); scdev_free(h->b.H); scdev_free(h->b.X); scdev_free(h->b.Zp); scdev_free(h->b.zt);
    scdev_free(h->b.tail); scdev_free(h->b.tail2); scdev_free(h->b.counters);
    scdev_free(h->b.ctaBase); scdev_free(h->b.grpStart); scdev_free(h->b.wtab);
    scdev_free(h->tailPass.ctaBase); scdev_free(h->tailPass.grpStart); scdev_free(h->tailPass.Zp);
    scdev_free(h->headPass.ctaBase); scdev_free(h->headPass.grpStart); scdev_free(h->headPass.Zp);
    if (h->streamIn) scdev_stream_sync(h->streamIn);
    if (h->streamOut) scdev_stream_sync(h->streamOut);
    scdev_event_destroy(h->evDone); scdev_event_destroy(h->evIn); scdev_event_destroy(h->evFence); scdev_event_destroy(h->evMac); scdev_event_destroy(h->evTail);
    scdev_event_destroy(h->evTailB[0]); scdev_event_destroy(h->evTailB[1]);
    if (h->tlEv) for (int i = 0; i < 8 * h->tlCap; i++) scdev_event_destroy(h->tlEv[i]);
    free(h->tlEv); free(h->tlHost); free(h->tlReg);
    scdev_stream_destroy(h->streamIn); scdev_stream_destroy(h->streamOut);
    scdev_free(h->tailPass.ZpB);
    for (int i = 0; i < 6; i++) scdev_event_destroy(h->trEv[i]);
    scdev_free(h->d_in); scdev_free(h->d_out);
    scdev_host_free(h->h_in); scdev_host_free(h->h_out); scdev_host_free((void*)h->mailbox);
    scdev_stream_destroy(h->streamOwn);
    h->magic = 0;
    free(h);
}

static int upload(safconv_handle* h, void** dptr, const void* src, size_t bytes, const char* what)
{
    int e = scdev_malloc(dptr, bytes);
    if (e) return h_fail(h, SAFCONV_ERR_NOMEM, what, e);
    e = scdev_memcpy_h2d_sync(*dptr, src, bytes, h->stream);
    if (e) return h_fail(h, SAFCONV_ERR_CUDA, what, e);
    return 0;
}

static int zalloc(safconv_handle* h, void** dptr, size_t bytes, const char* what)
{
    int e = scdev_malloc(dptr, bytes);
    if (e) return h_fail(h, SAFCONV_ERR_NOMEM, what, e);
    e = scdev_memset_async(*dptr, 0, bytes ? bytes : 16, h->stream);
    if (e) return h_fail(h, SAFCONV_ERR_CUDA, what, e);
    return 0;
}

/* split-K tables + partial-tile buffer of the MAC pass over partitions [pLo, pLo + nP) */
static int make_pass(safconv_handle* h, scdev_macpass* ps, int pLo, int nP)
{
    const scdev_plan* pl = &h->pl;
    ps->pLo = pLo; ps->nP = nP;
    ps->totalStages = (long long)pl->nGroups * nP * pl->SPU;
    ps->grid = (ps->totalStages < h->smCount) ? (int)ps->totalStages : h->smCount;
    if (pLo >= 1) {
        /* A tail pass runs while the caller comes back with the next block, whose K1 -> head pass -> K3 chain is the
         * latency of the call.  On an SM shared with a MAC CTA those kernels run 3-5x slower (issue slots, not HBM: the
         * device timeline, SAFCONV_TIMELINE, shows K3 16 us alone and 60-70 us beside a tail pass), and the tail pass is
         * HBM-bound long before it needs every SM -- so it leaves 32 SMs free (SAFCONV_TAIL_RESERVE_SMS).  Measured on
         * one GPU's share of configs[3] at 8 GPUs (64 x 8, 59 us of filter stream per block): back-to-back p50 85 -> 63 us;
         * configs[3] on one GPU: p50 0.459 -> 0.449 ms, p99 0.88 -> 0.48 ms. */
        int reserve = env_int("SAFCONV_TAIL_RESERVE_SMS", 32, 0, 64);
        if (reserve > h->smCount / 2) reserve = h->smCount / 2;
        if (reserve && ps->grid > h->smCount - reserve) ps->grid = h->smCount - reserve;
    }
    int *ctaBase = NULL, *grpStart = NULL, *grpList = NULL;
    const int slots = build_split_tables_for(ps->totalStages, (long long)nP * pl->SPU, ps->grid, pl->nGroups,
                                             &ctaBase, &grpStart, &grpList);
    if (slots < 0) return h_fail(h, SAFCONV_ERR_NOMEM, "split tables (pass)", 0);
    ps->nSlots = slots;
    int e = upload(h, (void**)&ps->ctaBase, ctaBase, sizeof(int) * (size_t)(ps->grid + 1), "ctaBase upload (pass)");
    if (!e) e = upload(h, (void**)&ps->grpStart, grpStart, sizeof(int) * (size_t)(pl->nGroups + 1), "grpStart upload (pass)");
    for (int q = 0; q < slots && !e; q++)
        if (grpList[q] != q) e = h_fail(h, SAFCONV_ERR_ARG, "internal: split-K slot order (pass)", 0);
    free(ctaBase); free(grpStart); free(grpList);
    if (e) return e;
    return zalloc(h, &ps->Zp, (size_t)slots * pl->OTsz * SC_BK * 8, "partial spectra allocation (pass)");
}

/* Common constructor.  `rows` time-domain FIRs of `len` taps are given as `nChunks` host chunks
 * (chunk i holds rowsPerChunk rows) so that TVConv's float** and the flat layouts share one path. */
static safconv_handle* conv_create(int kind, int hop, const float* const* chunks, int nChunks, size_t rowsPerChunk,
                                   int len, int nIn, int nOutLocal, int nOutTotal, int outBegin, int nIRs)
{
    tl_err = 0; tl_msg[0] = 0;
    if (hop > SC_MAX_M) {
        set_tl_error(SAFCONV_ERR_ARG, "invalid argument%s: hopSize > 8192 is not supported by this engine (one block's FFT lives in one CTA's "
                     "shared memory); the reference's own hosts clamp to 8192 (matrixconv_internal.h:40-41, tvconv_internal.h:43-44)", "");
        return NULL;
    }
    if (hop < 1 || len < 1 || nIn < 1 || nOutLocal < 1 || !chunks) {
        set_tl_error(SAFCONV_ERR_ARG, "invalid argument%s (need 1 <= hopSize <= 8192, length_h >= 1, channels >= 1, H != NULL)", "");
        return NULL;
    }
    int ndev = 0;
    if (scdev_device_count(&ndev) != 0 || ndev < 1) {
        set_tl_error(SAFCONV_ERR_NO_DEVICE, "no usable CUDA device%s (libsafconv_b200 has no CPU fallback)", "");
        return NULL;
    }
    safconv_handle* h = (safconv_handle*)calloc(1, sizeof *h);
    if (!h) { set_tl_error(SAFCONV_ERR_NOMEM, "out of host memory%s", ""); return NULL; }
    h->magic = SAFCONV_MAGIC;
    h->device = -1;
    h->nCHoutTotal = nOutTotal; h->outBegin = outBegin;

    int dev = tl_device;
    if (dev < 0) DEV_TRY(h, scdev_get_device(&dev), "cudaGetDevice");
    DEV_TRY(h, scdev_set_device(dev), "cudaSetDevice");
    h->device = dev;
    int ccMaj = 0, ccMin = 0;
    DEV_TRY(h, scdev_device_props(dev, &h->smCount, &h->maxSmem, &ccMaj, &ccMin), "device properties");
    DEV_TRY(h, scdev_stream_create(&h->streamOwn), "cudaStreamCreate");
    h->stream = h->streamOwn;

    scdev_plan* pl = &h->pl;
    pl->kind = kind; pl->nIn = nIn; pl->nOutLocal = nOutLocal; pl->nIRs = nIRs;
    plan_fft(pl, hop, len);
    pl->macHints = env_int("SAFCONV_MAC_HINTS", -1, -1, 2);   /* -1: decided below from the size of the filter set */
    h->detectPinned = 1;
    h->batching = 1;
    h->smallFused = env_int("SAFCONV_SMALL_FUSED", 1, 0, 1);
    h->hostTrace = env_int("SAFCONV_HOSTTRACE", 0, 0, 1);
    h->flagWait = env_int("SAFCONV_FLAG_WAIT", 1, 0, 1);
    h->residentUs = env_int("SAFCONV_RESIDENT_US", 0, 0, 2000000);

    const size_t M = (size_t)pl->M, P = (size_t)pl->P;
    /* twiddles W_N^j, j < M, evaluated in double like the reference's KissFFT tables (kiss_fft.c:358-364) */
    {
        float* tw = (float*)malloc(sizeof(float) * 2 * M);
        if (!tw) { h_fail(h, SAFCONV_ERR_NOMEM, "twiddle table", 0); goto fail; }
        for (size_t j = 0; j < M; j++) {
            const double ph = -2.0 * 3.141592653589793238462643383279502884 * (double)j / (double)pl->N;
            tw[2 * j] = (float)cos(ph);
            tw[2 * j + 1] = (float)sin(ph);
        }
        int e = upload(h, &h->b.tw, tw, sizeof(float) * 2 * M, "twiddle upload");
        free(tw);
        if (e) goto fail;
    }

    size_t rowsTotal;
    if (kind == SC_KIND_MATRIX) {
        plan_mac(pl, h->smCount);
        if (pl->macSmemBytes > h->maxSmem) { h_fail(h, SAFCONV_ERR_CUDA, "MAC pipeline does not fit in shared memory", 0); goto fail; }
        int *ctaBase = NULL, *grpStart = NULL, *grpList = NULL;
        const int slots = build_split_tables(pl, &ctaBase, &grpStart, &grpList);
        if (slots < 0) { h_fail(h, SAFCONV_ERR_NOMEM, "split tables", 0); goto fail; }
        pl->nSlots = slots;
        int e = upload(h, (void**)&h->b.ctaBase, ctaBase, sizeof(int) * (size_t)(pl->macGrid + 1), "ctaBase upload");
        if (!e) e = upload(h, (void**)&h->b.grpStart, grpStart, sizeof(int) * (size_t)(pl->nGroups + 1), "grpStart upload");
        /* the partial slots of a group are consecutive by construction (slots are numbered in (CTA, segment)
         * order and that order is monotone in the group index): K3 only needs grpStart */
        for (int q = 0; q < slots; q++) if (grpList[q] != q) e = -1;
        free(ctaBase); free(grpStart); free(grpList);
        if (e) { h_fail(h, SAFCONV_ERR_ARG, "internal: split-K slot order", 0); goto fail; }
        while (pl->maxBatch > 1 && (double)pl->maxBatch * slots * pl->OTsz * SC_BK * 8.0 > 256e6) pl->maxBatch >>= 1;
        pl->RS = pl->P + pl->maxBatch;
        h->bytesH  = (size_t)pl->nOT * pl->nKT * P * nIn * pl->OTsz * SC_BK * 8;
        /* L2 policy of the filter stream: a filter set that fits in L2 (126 MB) is kept there between blocks */
        if (pl->macHints < 0) pl->macHints = (h->bytesH <= ((size_t)64 << 20)) ? 2 : 1;
        h->bytesX  = (size_t)pl->nKT * pl->RS * nIn * SC_BK * 8;
        h->bytesZp = (size_t)pl->maxBatch * slots * pl->OTsz * SC_BK * 8;
        rowsTotal  = (size_t)nOutLocal * nIn;
    } else if (kind == SC_KIND_MULTI) {
        /* blocks handed over together are transformed ahead of their MAC: maxBatch extra ring slots (<= ~256 MB) */
        int mb = env_int("SAFCONV_MAX_BATCH", 32, 1, 256);
        while (mb > 1 && (double)mb * nOutLocal * ((double)M * 8.0 + (double)hop * 8.0) > 256e6) mb >>= 1;
        pl->maxBatch = mb;
        pl->RS = pl->P + mb;
        h->bytesH = (size_t)nOutLocal * P * M * 8;
        h->bytesX = (size_t)nOutLocal * pl->RS * M * 8;
        rowsTotal = (size_t)nOutLocal;
    } else {
        h->bytesH = (size_t)nIRs * nOutLocal * P * M * 8;
        h->bytesX = P * M * 8;
        rowsTotal = (size_t)nIRs * nOutLocal;
    }
    if (kind == SC_KIND_MATRIX && pl->P >= 2) {
        /* look-ahead depth D: the tail pass of block b covers partitions p >= D (it reads X(b-D) and older, so it can be
         * enqueued D calls ahead), the head pass partitions p < D.  D = 2 (SAFCONV_LA_DEPTH=2) keeps TWO tail passes queued,
         * which makes back-to-back call times very even (C4: p99 0.47 ms instead of 0.89 ms) at the price of a head pass
         * twice as long (paced p50 63 instead of 55 us); measured on both, D = 1 stays the default */
        h->laDepth = (pl->P >= 3) ? env_int("SAFCONV_LA_DEPTH", 1, 1, 2) : 1;
        if (make_pass(h, &h->tailPass, h->laDepth, pl->P - h->laDepth) || make_pass(h, &h->headPass, 0, h->laDepth)) goto fail;
        DEV_TRY(h, scdev_event_create_sync(&h->evDone), "cudaEventCreate");
        DEV_TRY(h, scdev_event_create_sync(&h->evIn), "cudaEventCreate");
        DEV_TRY(h, scdev_event_create_sync(&h->evFence), "cudaEventCreate");
        DEV_TRY(h, scdev_event_create_sync(&h->evMac), "cudaEventCreate");
        DEV_TRY(h, scdev_event_create_sync(&h->evTail), "cudaEventCreate");
        DEV_TRY(h, scdev_event_create_sync(&h->evTailB[0]), "cudaEventCreate");
        DEV_TRY(h, scdev_event_create_sync(&h->evTailB[1]), "cudaEventCreate");
        DEV_TRY(h, scdev_stream_create(&h->streamIn), "cudaStreamCreate");
        DEV_TRY(h, scdev_stream_create_high_priority(&h->streamOut), "cudaStreamCreate");
        if (zalloc(h, &h->tailPass.ZpB, (size_t)h->tailPass.nSlots * pl->OTsz * SC_BK * 8, "partial spectra allocation (tail, second buffer)")) goto fail;
        h->lookahead = env_int("SAFCONV_LOOKAHEAD", 1, 0, 1);
        h->oneStreamLatency = env_int("SAFCONV_ONE_STREAM", 0, 0, 1);   /* measured: 55.3 vs 54.9 us paced p50 at configs[3] -- the event hops are not what the call waits for; off */
        h->headInK3 = env_int("SAFCONV_HEAD_IN_K3", 0, 0, 1);   /* since K3 gathers in three load rounds, head pass + K3 (55 us) beats K3 adding the partition itself (62 us) */
        h->trace = env_int("SAFCONV_TRACE", 0, 0, 1);
        for (int i = 0; i < 6 && h->trace; i++) DEV_TRY(h, scdev_event_create(&h->trEv[i]), "cudaEventCreate");
        h->tlCap = env_int("SAFCONV_TIMELINE", 0, 0, 256);
        if (h->tlCap) {
            h->tlEv = (void**)calloc((size_t)8 * h->tlCap, sizeof(void*));
            h->tlHost = (double*)calloc((size_t)2 * h->tlCap, sizeof(double));
            h->tlReg = (char*)calloc((size_t)h->tlCap, 1);
            if (!h->tlEv || !h->tlHost || !h->tlReg) { h_fail(h, SAFCONV_ERR_NOMEM, "timeline buffers", 0); goto fail; }
            for (int i = 0; i < 8 * h->tlCap; i++) DEV_TRY(h, scdev_event_create(&h->tlEv[i]), "cudaEventCreate");
        }
    }
    h->smallOk = (kind == SC_KIND_MATRIX) ? scdev_small_fits(pl, h->maxSmem) : 0;
    if ((kind == SC_KIND_MULTI && env_int("SAFCONV_MULTI_WFFT", 1, 0, 1)) || (kind == SC_KIND_MATRIX && h->smallOk))
        DEV_TRY(h, scdev_wfft_tables(pl, &h->b, h->stream), "warp-FFT tables");     /* the latency kernel of small matrix problems uses them too */
    DEV_TRY(h, scdev_prepare(pl), "kernel attribute setup");

    if (zalloc(h, &h->b.H, h->bytesH, "filter spectra allocation")) goto fail;
    if (zalloc(h, &h->b.X, h->bytesX, "delay line allocation")) goto fail;
    if (kind == SC_KIND_MATRIX && zalloc(h, &h->b.Zp, h->bytesZp, "partial spectra allocation")) goto fail;
    if ((kind == SC_KIND_MATRIX || kind == SC_KIND_MULTI) && pl->maxBatch > 1 &&
        zalloc(h, (void**)&h->b.zt, sizeof(float) * (size_t)pl->maxBatch * nOutLocal * 2 * hop, "batched inverse-transform buffer")) goto fail;
    if (zalloc(h, (void**)&h->b.tail, sizeof(float) * (size_t)nOutLocal * hop, "overlap tails")) goto fail;
    if (kind == SC_KIND_TV && zalloc(h, (void**)&h->b.tail2, sizeof(float) * (size_t)nOutLocal * hop, "overlap tails (last)")) goto fail;
    if (kind == SC_KIND_TV && pl->M > 4096 &&       /* hop > 4096: the three-launch path keeps its inverse transforms in global memory */
        zalloc(h, (void**)&h->b.zt, sizeof(float) * 3 * (size_t)nOutLocal * 2 * hop, "TVConv inverse-transform scratch")) goto fail;
    if (zalloc(h, (void**)&h->b.counters, 4 * sizeof(unsigned int), "counters")) goto fail;

    h->inBytes  = sizeof(float) * (size_t)nIn * hop;
    h->outBytes = sizeof(float) * (size_t)nOutLocal * hop;
    if (kind == SC_KIND_TV) h->inBytes = sizeof(float) * (size_t)hop;
    if (zalloc(h, (void**)&h->d_in, h->inBytes, "input staging")) goto fail;
    if (zalloc(h, (void**)&h->d_out, h->outBytes, "output staging")) goto fail;
    DEV_TRY(h, scdev_host_alloc((void**)&h->h_in, h->inBytes), "pinned input staging");
    DEV_TRY(h, scdev_host_alloc((void**)&h->h_out, h->outBytes), "pinned output staging");
    if (h->smallOk) {
        DEV_TRY(h, scdev_host_alloc((void**)&h->mailbox, sizeof(sc_mailbox)), "completion word / mailbox");
        memset((void*)h->mailbox, 0, sizeof(sc_mailbox));
        h->doneWord = &h->mailbox->done;
    }

    /* K0: upload the time-domain filters once and transform them on the device (reference .c:116-125) */
    {
        float* d_h = NULL;
        const size_t rowBytes = sizeof(float) * (size_t)len;
        int e = 0;
        if (tl_filters_on_device) {
            /* the bank was produced on this device (safconv_producers.c): transform it where it lies */
            e = scdev_filter_transform(pl, &h->b, chunks[0], h->stream);
            if (!e) e = scdev_stream_sync(h->stream);
        } else {
            e = scdev_malloc((void**)&d_h, rowsTotal * rowBytes);
            if (e) { h_fail(h, SAFCONV_ERR_NOMEM, "time-domain filter upload buffer", e); goto fail; }
            for (int c = 0; c < nChunks && !e; c++)
                e = scdev_memcpy_h2d_sync((char*)d_h + (size_t)c * rowsPerChunk * rowBytes, chunks[c], rowsPerChunk * rowBytes, h->stream);
            if (!e) e = scdev_filter_transform(pl, &h->b, d_h, h->stream);
            if (!e) e = scdev_stream_sync(h->stream);
            scdev_free(d_h);
        }
        if (e) { h_fail(h, SAFCONV_ERR_CUDA, "filter transform", e); goto fail; }
    }
    return h;

fail:
    handle_free(h);
    return NULL;
}

safconv_handle* sch_conv_create(int kind, int hop, const float* const* chunks, int nChunks, size_t rowsPerChunk,
                                int len, int nIn, int nOutLocal, int nOutTotal, int outBegin, int nIRs)
{ return conv_create(kind, hop, chunks, nChunks, rowsPerChunk, len, nIn, nOutLocal, nOutTotal, outBegin, nIRs); }
void sch_handle_free(safconv_handle* h) { handle_free(h); }

static void conv_destroy(void** const ph)
{
    if (!ph) return;
    if (scm_is_multi(*ph)) { scm_destroy(ph); return; }
    if (scn_is_np(*ph)) { scn_destroy(ph); return; }
    safconv_handle* h = as_handle(*ph);
    if (h) handle_free(h);
    *ph = NULL;
}

/* ------------------------------------------------------------------------------------------ */
/*  per-block sequencing                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* nBlocks (<= maxBatch) consecutive blocks: d_in [nBlocks][nIn][hop] -> d_out [nBlocks][nOutLocal][hop].
 * matrix: K1 for all blocks in one launch, one MAC launch per block (each streams the filters once),
 * K3 for all blocks (one block: the fused inverse-FFT + overlap-add kernel). */
static int enqueue_blocks(safconv_handle* h, const float* d_in, float* d_out, int nBlocks)
{
    const scdev_plan* pl = &h->pl;
    void** ev = NULL;
    int e = 0;
    if (pl->kind == SC_KIND_MATRIX && nBlocks == 1 && h->smallFused && h->smallOk && !h->useGraph && !h->timingCap) {
        return scdev_small_fused(pl, &h->b, d_in, d_out, h->stream, NULL, 0, NULL);     /* one launch instead of three */
    }
    if (pl->kind == SC_KIND_MATRIX) {
        if (h->timingCap && h->timingCount < h->timingCap) {
            ev = h->evRing + 4 * (size_t)h->timingCount;
            h->evBlocks[h->timingCount++] = nBlocks;
        }
        if (ev) scdev_event_record(ev[0], h->stream);
        e = scdev_input_fft(pl, &h->b, d_in, nBlocks, h->stream);
        if (ev) scdev_event_record(ev[1], h->stream);
        if (!e) e = scdev_mac(pl, &h->b, 0, nBlocks, h->stream);
        if (ev) scdev_event_record(ev[2], h->stream);
        if (!e) e = (nBlocks == 1) ? scdev_ifft_ola(pl, &h->b, d_out, h->stream)
                                   : scdev_ifft_ola_batch(pl, &h->b, d_out, nBlocks, h->stream);
        if (ev) scdev_event_record(ev[3], h->stream);
    } else if (pl->kind == SC_KIND_MULTI && nBlocks > 1) {
        /* batch: forward FFTs of all blocks, then all (channel, block) MACs + inverse FFTs, then the overlap-add chain */
        if (h->timingCap && h->timingCount < h->timingCap) {
            ev = h->evRing + 4 * (size_t)h->timingCount;
            h->evBlocks[h->timingCount++] = nBlocks;
        }
        if (ev) scdev_event_record(ev[0], h->stream);
        e = scdev_multi_batch(pl, &h->b, d_in, d_out, nBlocks, 0, h->stream);
        if (ev) scdev_event_record(ev[1], h->stream);
        if (!e) e = scdev_multi_batch(pl, &h->b, d_in, d_out, nBlocks, 1, h->stream);
        if (ev) scdev_event_record(ev[2], h->stream);
        if (!e) e = scdev_multi_batch(pl, &h->b, d_in, d_out, nBlocks, 2, h->stream);
        if (ev) scdev_event_record(ev[3], h->stream);
    } else if (pl->kind == SC_KIND_MULTI) {
        const size_t inStride = (size_t)pl->nIn * pl->hop, outStride = (size_t)pl->nOutLocal * pl->hop;
        for (int b = 0; b < nBlocks && !e; b++) {
            ev = NULL;
            if (h->timingCap && h->timingCount < h->timingCap) {
                ev = h->evRing + 4 * (size_t)h->timingCount;
                h->evBlocks[h->timingCount++] = 1;
            }
            if (ev) scdev_event_record(ev[1], h->stream);
            e = scdev_multi_fused(pl, &h->b, d_in + b * inStride, d_out + b * outStride, h->stream);
            if (ev) scdev_event_record(ev[2], h->stream);
        }
    } else {
        return (int)SAFCONV_ERR_ARG;
    }
    return e;
}

static int enqueue_block(safconv_handle* h, const float* d_in, float* d_out) { return enqueue_blocks(h, d_in, d_out, 1); }

/* ---- look-ahead apply (matrix, P >= 2; DESIGN.md §4 "Look-ahead apply") --------------------------------------
 * Streams:  stream     tail passes (+ head pass / K3 in the latency regime)
 *           streamIn   K1 of the new block (overlaps a tail pass that is still streaming)
 *           streamOut  high priority: head pass and K3 in the throughput regime, K3 of the first block
 * State:    tailReady  a tail pass for block `count` has been enqueued (partial tiles in tail buffer count & 1)
 *           evTail     end of the most recent tail pass;  evDone  this block's output is complete */

#define LA_TRY(call) do { if (!e) e = (call); } while (0)
#define LA_TL(i, st) do { if (h->tlEv && h->tlN < h->tlCap) scdev_event_record(h->tlEv[8 * h->tlN + (i)], (st)); } while (0)
#define LA_TRACE(i, st) do { if (tr) scdev_event_record(h->trEv[i], (st)); } while (0)

/* K1 of the new block; returns with `stream` ordered behind it.  Three sources (sch_la_io): a copy-engine upload into
 * h->d_in (K1 then runs on `stream`), or K1 on the side stream straight from a page-locked host buffer / from a device
 * buffer that a foreign stream fills (evSrc). */
static int la_input(safconv_handle* h, const sch_la_io* io, int idle)
{
    const scdev_plan* pl = &h->pl;
    int e = 0;
    if (io->h2dSrc) {
        e = scdev_memcpy_h2d_async(h->d_in, io->h2dSrc, h->inBytes, h->stream);
        LA_TRY(scdev_input_fft(pl, &h->b, h->d_in, 1, h->stream));
        return e;
    }
    if (idle && !io->evSrc) {
        /* latency regime: nothing is running, so there is nothing to overlap -- K1, head pass and K3 go down ONE stream
         * (a kernel-to-kernel hand-over on a stream costs ~1.5 us, an event hop between streams 3-4 us) */
        LA_TL(0, h->stream);
        e = scdev_input_fft(pl, &h->b, io->k1src, 1, h->stream);
        LA_TL(1, h->stream);
        LA_TRY(scdev_event_record(h->evIn, h->stream));
        return e;
    }
    if (!h->tailReady) {                  /* something else may still be running on `stream`: order K1 behind it */
        e = scdev_event_record(h->evFence, h->stream);
        LA_TRY(scdev_stream_wait_event(h->streamIn, h->evFence));
    }
    if (io->evSrc) LA_TRY(scdev_stream_wait_event(h->streamIn, io->evSrc));
    LA_TL(0, h->streamIn);
    LA_TRY(scdev_input_fft(pl, &h->b, io->k1src, 1, h->streamIn));
    LA_TL(1, h->streamIn);
    LA_TRY(scdev_event_record(h->evIn, h->streamIn));
    LA_TRY(scdev_stream_wait_event(h->stream, h->evIn));
    return e;
}

/* Throughput regime (the previous tail pass is still running when the caller comes back): `stream` carries nothing
 * but tail passes, back to back.  The head pass of block c runs on the side stream with a two-stage pipeline (69 KB of
 * shared memory: its CTAs fit on the SMs beside the resident tail CTAs) while the block's own tail pass is still
 * streaming; K3 follows as soon as that tail pass is done, beside the tail pass of block c+1. */
static int la_throughput(safconv_handle* h, unsigned int c, const sch_la_io* io, int tr)
{
    const scdev_plan* pl = &h->pl;
    const int tb = (int)(c & 1u);
    scdev_macpass head2 = h->headPass;
    head2.stages = 2;
    const int zc = (io->h2dSrc == NULL);                                  /* K1 ran on the side stream */
    float* const kout = io->kout;
    int e = 0;
    if (!zc) e = scdev_event_record(h->evMac, h->stream);                 /* K1 ran on `stream`: its spectrum is ready here */
    LA_TRY(scdev_stream_wait_event(h->streamOut, zc ? h->evIn : h->evMac));
    LA_TRACE(0, h->streamOut); LA_TL(2, h->streamOut);
    LA_TRY(scdev_mac_pass(pl, &h->b, &head2, 0, 1, 0, -1, h->streamOut));
    LA_TRACE(1, h->streamOut); LA_TL(3, h->streamOut);
    LA_TRY(scdev_stream_wait_event(h->streamOut, h->evTailB[tb]));        /* tail pass of block c: K3 needs both */
    LA_TRACE(3, h->streamOut); LA_TL(4, h->streamOut);
    LA_TRY(scdev_ifft_ola_passes(pl, &h->b, &h->tailPass, tb, &h->headPass, kout, h->streamOut));
    if (io->d2hDst) LA_TRY(scdev_memcpy_d2h_async(io->d2hDst, h->d_out, h->outBytes, h->streamOut));
    LA_TRACE(4, h->streamOut); LA_TL(5, h->streamOut);
    LA_TRY(scdev_event_record(h->evDone, h->streamOut));
    return e;
}

/* Latency regime (real-time pacing: the GPU is idle when the block arrives).  With a pre-computed tail there is no
 * head-pass launch at all: K3 adds the newest partition itself while it gathers the tail's partial tiles (or, with
 * SAFCONV_HEAD_IN_K3=0, head pass then K3 on the side stream).  Without one (first block, or after any other call on the
 * handle) the full MAC runs.  The next tail pass is ordered behind K3. */
static int la_latency(safconv_handle* h, unsigned int c, int hadTail, const sch_la_io* io, int tr)
{
    const scdev_plan* pl = &h->pl;
    const int tb = (int)(c & 1u);
    float* const kout = io->kout;
    int e = 0;
    if (hadTail && h->headInK3) {
        LA_TRACE(0, h->stream); LA_TRACE(1, h->stream); LA_TRACE(3, h->stream); LA_TL(2, h->stream); LA_TL(3, h->stream); LA_TL(4, h->stream);
        LA_TRY(scdev_ifft_ola_passes(pl, &h->b, &h->tailPass, tb, NULL, kout, h->stream));
        if (io->d2hDst) LA_TRY(scdev_memcpy_d2h_async(io->d2hDst, h->d_out, h->outBytes, h->stream));
        LA_TRACE(4, h->stream); LA_TL(5, h->stream);
        LA_TRY(scdev_event_record(h->evDone, h->stream));
    } else {
        LA_TRACE(0, h->stream);
        if (hadTail) LA_TRY(scdev_mac_pass(pl, &h->b, &h->headPass, 0, 1, 0, -1, h->stream));
        else         LA_TRY(scdev_mac(pl, &h->b, 0, 1, h->stream));
        LA_TRACE(1, h->stream);
        if (h->oneStreamLatency && io->sync && !io->evSrc) {
            LA_TRACE(3, h->stream); LA_TL(2, h->stream); LA_TL(3, h->stream); LA_TL(4, h->stream);
            if (hadTail) LA_TRY(scdev_ifft_ola_passes(pl, &h->b, &h->tailPass, tb, &h->headPass, kout, h->stream));
            else         LA_TRY(scdev_ifft_ola(pl, &h->b, kout, h->stream));
            if (io->d2hDst) LA_TRY(scdev_memcpy_d2h_async(io->d2hDst, h->d_out, h->outBytes, h->stream));
            LA_TRACE(4, h->stream); LA_TL(5, h->stream);
            LA_TRY(scdev_event_record(h->evDone, h->stream));
            return e;
        }
        LA_TRY(scdev_event_record(h->evMac, h->stream));
        LA_TRY(scdev_stream_wait_event(h->streamOut, h->evMac));
        LA_TRACE(3, h->streamOut);
        if (hadTail) LA_TRY(scdev_ifft_ola_passes(pl, &h->b, &h->tailPass, tb, &h->headPass, kout, h->streamOut));
        else         LA_TRY(scdev_ifft_ola(pl, &h->b, kout, h->streamOut));
        if (io->d2hDst) LA_TRY(scdev_memcpy_d2h_async(io->d2hDst, h->d_out, h->outBytes, h->streamOut));
        LA_TRACE(4, h->streamOut);
        LA_TRY(scdev_event_record(h->evDone, h->streamOut));
        LA_TRY(scdev_stream_wait_event(h->stream, h->evDone));
    }
    return e;
}

/* SAFCONV_TRACE=1: device timeline of this call on stderr (trEv[2] = end of the tail pass enqueued by the previous call) */
static void la_trace_report(safconv_handle* h, unsigned int c, int tr)
{
    if (tr && h->count > 3) {
        float gap = 0.f, head = 0.f, k3wait = 0.f, k3 = 0.f;
        scdev_event_elapsed_ms(h->trEv[2], h->trEv[0], &gap);
        scdev_event_elapsed_ms(h->trEv[0], h->trEv[1], &head);
        scdev_event_elapsed_ms(h->trEv[1], h->trEv[3], &k3wait);
        scdev_event_elapsed_ms(h->trEv[3], h->trEv[4], &k3);
        fprintf(stderr, "[safconv trace] block %u (%s): prev tail end -> head start %.1f us, head %.1f us, head end -> K3 start %.1f us, K3 %.1f us\n",
                c, h->trRegime ? "throughput" : "latency", 1e3f * gap, 1e3f * head, 1e3f * k3wait, 1e3f * k3);
    }
    if (h->trace) scdev_event_record(h->trEv[2], h->stream);      /* end of the tail pass just enqueued: read by the next call */
}

/* SAFCONV_TIMELINE=n: the device timeline of the first n look-ahead calls, microseconds since K1 of the first one
 * (missing stages -- no head pass in the latency regime, cold start -- print as '-') */
static void la_timeline_print(safconv_handle* h)
{
    scdev_stream_sync(h->stream); scdev_stream_sync(h->streamOut); scdev_stream_sync(h->streamIn);
    static const char* name[4] = { "K1", "head", "K3", "tail" };
    fprintf(stderr, "[safconv timeline] call regime host(in..out) |");
    for (int i = 0; i < 4; i++) fprintf(stderr, " %s(start..end)", name[i]);
    fprintf(stderr, "   (us; device times relative to the first K1, host times to the first call)\n");
    for (int n = 0; n < h->tlN; n++) {
        fprintf(stderr, "[safconv timeline] %3d %c %8.1f..%-8.1f |", n, h->tlReg[n], (h->tlHost[2 * n] - h->tlHost[0]) * 1e-3,
                (h->tlHost[2 * n + 1] - h->tlHost[0]) * 1e-3);
        for (int i = 0; i < 4; i++) {
            float a = 0.f, b = 0.f;
            const int ea = scdev_event_elapsed_ms(h->tlEv[0], h->tlEv[8 * n + 2 * i], &a);
            const int eb = scdev_event_elapsed_ms(h->tlEv[0], h->tlEv[8 * n + 2 * i + 1], &b);
            if (ea || eb) { (void)scdev_last_error_clear(); fprintf(stderr, "        -        "); }
            else fprintf(stderr, " %8.1f..%-8.1f", 1e3f * a, 1e3f * b);
        }
        fprintf(stderr, "\n");
    }
}

/* One block through the look-ahead sequence described by `io`; with io->sync the output is complete on return,
 * otherwise evDone is recorded behind it. */
int sch_apply_lookahead_io(safconv_handle* h, const sch_la_io* io)
{
    const scdev_plan* pl = &h->pl;
    const unsigned int c = h->count;
    const int D = h->laDepth;
    const int hadTail = h->tailReady;                         /* tail passes of blocks c .. tailUpTo are enqueued */
    const unsigned int upTo = hadTail ? h->tailUpTo : c;      /* (c itself only counts with hadTail) */
    const int tr = h->trace && hadTail;
    const int tl = h->tlEv && h->tlN < h->tlCap;
    if (tl) h->tlHost[2 * h->tlN] = now_ns();
    /* is a tail pass still running, i.e. is the caller coming back faster than the GPU streams the filters? */
    int backToBack = 0;
    if (hadTail) {
        backToBack = scdev_event_done(h->evTailB[c & 1u]) == 0;
        if (!backToBack && upTo > c) backToBack = scdev_event_done(h->evTailB[(c + 1u) & 1u]) == 0;
    }
    int e = la_input(h, io, !backToBack && h->oneStreamLatency && io->sync);
    h->tailReady = 0;
    h->trRegime = backToBack;
    if (backToBack) LA_TRY(la_throughput(h, c, io, tr));
    else            LA_TRY(la_latency(h, c, hadTail, io, tr));
    /* the tail passes that are not queued yet, up to block c + D.  `stream` is ordered behind K1 of this block (la_input)
     * and, in the latency regime, behind K3.  The pass of block c + 2 re-uses the partial-tile buffer K3 of this block
     * reads: it waits for K3 (which ends long before the pass of block c + 1, queued in front of it, does). */
    for (unsigned int b = (hadTail ? upTo : c) + 1u; !e && b <= c + (unsigned)D; ++b) {
        if (b == c + 2u) LA_TRY(scdev_stream_wait_event(h->stream, h->evDone));
        LA_TL(6, h->stream);
        LA_TRY(scdev_mac_pass(pl, &h->b, &h->tailPass, 0, 1, (int)(b & 1u), (long long)b, h->stream));
        LA_TL(7, h->stream);
        LA_TRY(scdev_event_record(h->evTailB[b & 1u], h->stream));
    }
    if (!e) { h->tailReady = 1; h->tailUpTo = c + (unsigned)D; h->count = c + 1; if (io->sync) e = scdev_event_sync(h->evDone); }
    if (!e && io->sync) la_trace_report(h, c, tr);
    if (e) h->tailReady = 0;
    if (tl) {
        h->tlHost[2 * h->tlN + 1] = now_ns();
        h->tlReg[h->tlN] = backToBack ? 'T' : (hadTail ? 'L' : 'C');
        if (++h->tlN == h->tlCap) la_timeline_print(h);
    }
    return e;
}

static int apply_lookahead(safconv_handle* h, const float* src, float* dst)
{
    /* blocks of up to 1 MB are read / written by K1 / K3 straight from / to the page-locked host buffers */
    const int zc = h->inBytes <= (1u << 20) && h->outBytes <= (1u << 20);
    sch_la_io io;
    io.k1src = zc ? src : h->d_in;  io.h2dSrc = zc ? NULL : src;  io.evSrc = NULL;
    io.kout  = zc ? dst : h->d_out; io.d2hDst = zc ? NULL : dst;  io.sync = 1;
    return sch_apply_lookahead_io(h, &io);
}

int sch_uses_lookahead(const safconv_handle* h)
{
    return h->lookahead && h->pl.kind == SC_KIND_MATRIX && !h->useGraph && !h->timingCap && !(h->smallFused && h->smallOk);
}

/* ---- resident latency kernel (option "resident_us") ----------------------------------------------
 * While it is active the handle's stream is busy with it: every other entry point stops it first (res_stop). */
static void res_stop(safconv_handle* h)
{
    if (!h || !h->resActive) return;
    if (env_int("SAFCONV_KSTAMPS", 0, 0, 1)) {                     /* debugging: phases of the last block served */
        unsigned long long t[8];
        if (scdev_small_resident_stamps(t) == 0)
            fprintf(stderr, "resident kernel, last block (us): doorbell -> start %.2f, A %.2f, barrier %.2f, B %.2f, barrier %.2f, C %.2f\n",
                    (double)(long long)(t[0] - t[6]) * 1e-3, (double)(long long)(t[1] - t[0]) * 1e-3, (double)(long long)(t[2] - t[1]) * 1e-3,
                    (double)(long long)(t[3] - t[2]) * 1e-3, (double)(long long)(t[4] - t[3]) * 1e-3, (double)(long long)(t[5] - t[4]) * 1e-3);
    }
    h->mailbox->bell = (unsigned long long)SC_RES_EXIT;
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    scdev_stream_sync(h->stream);
    h->mailbox->bell = (unsigned long long)h->doneSeq;
    h->mailbox->alive = 0;
    h->resActive = 0;
}

/* ring the doorbell for block (src -> dst); starts the kernel if none is polling.  Returns a CUDA error code. */
static int res_post(safconv_handle* h, const float* src, float* dst)
{
    sc_mailbox* mb = h->mailbox;
    int e = 0;
    if (h->resActive && !mb->alive) { e = scdev_stream_sync(h->stream); h->resActive = 0; }     /* it left on its idle timer */
    if (!e && !h->resActive) {
        mb->bell = (unsigned long long)h->doneSeq; mb->alive = 1;
        __atomic_thread_fence(__ATOMIC_SEQ_CST);
        e = scdev_small_resident_start(&h->pl, &h->b, (void*)mb, h->doneSeq, (unsigned)h->residentUs, h->stream);
        if (e) { mb->alive = 0; return e; }
        h->resActive = 1;
    }
    /* the block's buffers: a slot of the mailbox's table (callers re-use a handful of buffers; a new pair replaces a slot
     * and bumps the generation, which makes the kernel re-read the table once) */
    int slot = -1;
    for (int i = 0; i < 8; i++)
        if (mb->buf[2 * i] == (unsigned long long)(uintptr_t)src && mb->buf[2 * i + 1] == (unsigned long long)(uintptr_t)dst && src) { slot = i; break; }
    if (slot < 0) {
        slot = (int)(h->resNext++ & 7u);
        mb->buf[2 * slot] = (unsigned long long)(uintptr_t)src; mb->buf[2 * slot + 1] = (unsigned long long)(uintptr_t)dst;
        h->resGen = (h->resGen + 1u) & 0xFFFFFFu;
    }
    ++h->doneSeq;
    if (h->doneSeq == SC_RES_EXIT) h->doneSeq = 1;
    __atomic_thread_fence(__ATOMIC_RELEASE);                   /* table and input block before the doorbell */
    mb->bell = ((unsigned long long)((h->resGen << 8) | (unsigned)slot) << 32) | (unsigned long long)h->doneSeq;
    return e;
}

/* wait for the completion word of the block posted last; the kernel may have left on its idle timer in the same instant:
 * then it is started again (it sees the pending doorbell at once) */
static int res_wait(safconv_handle* h)
{
    sc_mailbox* mb = h->mailbox;
    const unsigned int want = h->doneSeq;
    const double tEnd = now_ns() + 2.0e9;
    unsigned int spins = 0;
    while (mb->done != want) {
        __builtin_ia32_pause();
        if ((++spins & 255u) == 0) {
            if (!mb->alive) {
                int e = scdev_stream_sync(h->stream);
                h->resActive = 0;
                if (e) return e;
                if (mb->done == want) break;
                mb->alive = 1;
                __atomic_thread_fence(__ATOMIC_SEQ_CST);
                e = scdev_small_resident_start(&h->pl, &h->b, (void*)mb, want - 1u, (unsigned)h->residentUs, h->stream);     /* it sees the pending doorbell at once */
                if (e) { mb->alive = 0; return e; }
                h->resActive = 1;
            }
            if (now_ns() > tEnd) { res_stop(h); return 702 /* cudaErrorLaunchTimeout */; }
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    return 0;
}

/* Blocks of up to 1 MB without look-ahead: the kernels read / write the page-locked host buffers directly -- no
 * copy-engine round trips, one synchronisation.  Returns -1 if the handle has no such path. */
static int apply_zero_copy(safconv_handle* h, const float* src, float* dst, int irIdx, int* signalled)
{
    const scdev_plan* pl = &h->pl;
    *signalled = 0;
    const int small = h->smallFused && !h->useGraph && !h->timingCap && pl->kind == SC_KIND_MATRIX && h->smallOk;
    if (!(small && h->residentUs > 0)) res_stop(h);
    if (!h->smallFused || h->useGraph || h->timingCap) return -1;
    if (small) {                                                  /* small problem: K1 + K2 + K3 in ONE launch */
        if (h->residentUs > 0 && h->mailbox && h->b.wtab) {       /* ... or no launch at all: the resident kernel's doorbell */
            const int e = res_post(h, src, dst);
            if (!e) { *signalled = 2; return 0; }
            /* not a cluster-kernel plan (cudaErrorNotSupported), or the resident kernel could not be started: this handle
             * goes on with one launch per block */
            h->residentUs = 0;
            if (e != 801) (void)scdev_last_error_clear();
        }
        return scdev_small_fused(pl, &h->b, src, dst, h->stream, (h->flagWait && h->doneWord) ? h->doneWord : NULL, ++h->doneSeq, signalled);
    }
    if (h->inBytes > (1u << 20) || h->outBytes > (1u << 20)) return -1;
    if (pl->kind == SC_KIND_MULTI)                                /* one fused launch, one CTA per channel */
        return scdev_multi_fused(pl, &h->b, src, dst, h->stream);
    if (pl->kind == SC_KIND_TV) {
        const int e = scdev_tv_fused(pl, &h->b, src, dst, irIdx, h->tvLast, h->tvLast2, h->stream);
        h->tvLast2 = h->tvLast;                                   /* reference .c:618-619 */
        h->tvLast  = irIdx;
        return e;
    }
    if (h->lookahead) return -1;                                  /* matrix, P >= 2: apply_lookahead */
    int e = scdev_input_fft(pl, &h->b, src, 1, h->stream);        /* single partition (or look-ahead switched off) */
    LA_TRY(scdev_mac(pl, &h->b, 0, 1, h->stream));
    LA_TRY(scdev_ifft_ola(pl, &h->b, dst, h->stream));
    return e;
}

/* host-pointer apply (reference semantics: synchronous).  Caller buffers that are already page-locked
 * (cudaHostAlloc / cudaHostRegister) are used directly; ordinary malloc'd buffers go through the handle's pinned
 * staging buffers. */
static void conv_apply_host(safconv_handle* h, const float* in, float* out, int irIdx)
{
    const double t0 = h->hostTrace ? now_ns() : 0.0;
    h_clear(h);
    int e = scdev_set_device(h->device);
    if (e) { h_fail(h, SAFCONV_ERR_CUDA, "cudaSetDevice", e); return; }
    /* first AND last byte: a caller buffer that is only partly registered must take the staging path */
    const int direct = h->detectPinned && scdev_is_pinned_host(in) && scdev_is_pinned_host(out)
                    && scdev_is_pinned_host((const char*)in + h->inBytes - 1) && scdev_is_pinned_host((const char*)out + h->outBytes - 1);
    const float* src = direct ? in : h->h_in;
    float*       dst = direct ? out : h->h_out;
    if (!direct) memcpy(h->h_in, in, h->inBytes);
    if (h->hostTrace) h->htAcc[0] += now_ns() - t0;
    sch_apply_pinned(h, src, dst, irIdx);
    if (!direct && !h->err) memcpy(out, h->h_out, h->outBytes);
}

/* one block, src / dst page-locked and visible to the handle's device */
void sch_apply_pinned(safconv_handle* h, const float* src, float* dst, int irIdx)
{
    int e = scdev_set_device(h->device);
    if (e) { h_fail(h, SAFCONV_ERR_CUDA, "cudaSetDevice", e); return; }
    const double t1 = h->hostTrace ? now_ns() : 0.0;
    int signalled = 0;
    e = apply_zero_copy(h, src, dst, irIdx, &signalled);
    if (e >= 0) {
        const double t2 = h->hostTrace ? now_ns() : 0.0;
        if (!e && signalled == 2) e = res_wait(h);
        else if (!e && signalled) {
            /* the kernel writes the call's sequence number behind its last output store (system-scope fences): poll that
             * word; after 2 ms without it (device busy elsewhere, or a fault) fall back to the stream, which also reports errors */
            const unsigned int want = h->doneSeq;
            const double tEnd = now_ns() + 2.0e6;
            unsigned int spins = 0;
            while (*h->doneWord != want) {
                __builtin_ia32_pause();
                if ((++spins & 1023u) == 0 && now_ns() > tEnd) { e = scdev_stream_sync(h->stream); break; }
            }
            __atomic_thread_fence(__ATOMIC_ACQUIRE);
        } else if (!e) e = scdev_stream_sync(h->stream);
        if (h->hostTrace) { const double t3 = now_ns(); h->htAcc[1] += t2 - t1; h->htAcc[2] += t3 - t2; h->htN++; }
        if (e) { h_fail(h, SAFCONV_ERR_CUDA, "apply (zero-copy)", e); return; }
        h->count++;
    } else if (h->lookahead && h->pl.kind == SC_KIND_MATRIX && !h->useGraph && !h->timingCap) {
        e = apply_lookahead(h, src, dst);                         /* advances h->count itself */
        if (e) { h_fail(h, SAFCONV_ERR_CUDA, "apply (look-ahead)", e); return; }
    } else {
        if (h->useGraph && h->pl.kind != SC_KIND_TV) {
            if (h->graphExec && (h->graphIn != (void*)src || h->graphOut != (void*)dst)) {
                scdev_graph_destroy(h->graphExec);
                h->graphExec = NULL;
            }
            if (!h->graphExec) {
                e = scdev_graph_begin(h->stream);
                if (!e) e = scdev_memcpy_h2d_async(h->d_in, src, h->inBytes, h->stream);
                if (!e) e = enqueue_block(h, h->d_in, h->d_out);
                if (!e) e = scdev_memcpy_d2h_async(dst, h->d_out, h->outBytes, h->stream);
                int e2 = scdev_graph_end(h->stream, &h->graphExec);
                if (!e) e = e2;
                if (e) { h->graphExec = NULL; h->useGraph = 0; h_fail(h, SAFCONV_ERR_CUDA, "graph capture", e); return; }
                h->graphIn = (void*)src; h->graphOut = (void*)dst;
            }
            e = scdev_graph_launch(h->graphExec, h->stream);
        } else {
            /* copy-engine path: H2D -> kernels -> D2H */
            e = scdev_memcpy_h2d_async(h->d_in, src, h->inBytes, h->stream);
            if (!e) {
                if (h->pl.kind == SC_KIND_TV) {
                    e = scdev_tv_fused(&h->pl, &h->b, h->d_in, h->d_out, irIdx, h->tvLast, h->tvLast2, h->stream);
                    h->tvLast2 = h->tvLast;                      /* reference .c:618-619 */
                    h->tvLast  = irIdx;
                } else {
                    e = enqueue_block(h, h->d_in, h->d_out);
                }
            }
            if (!e) e = scdev_memcpy_d2h_async(dst, h->d_out, h->outBytes, h->stream);
        }
        if (!e) e = scdev_stream_sync(h->stream);
        if (e) { h_fail(h, SAFCONV_ERR_CUDA, "apply", e); return; }
        h->count++;
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  drop-in API                                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* Blocks longer than 8192 samples: the partitioned engine keeps one block's FFT inside one CTA's shared memory and
 * stops there, but the reference takes any hop (saf_utility_matrixConv.c:100: fftSize = 2 * hopSize).  Such handles are
 * served by the big-FFT engine of safconv_np.c (one numOvrlpAddBlocks * hopSize-point transform per channel and block on
 * the general-size device FFT, which runs sizes beyond shared memory as one launch per pass): the same causal linear
 * convolution for either usePartFLAG.  Not possible only if that FFT size is odd (odd hop with an odd block count). */
static void* create_big_hop(int kind, int hopSize, const float* H, int length_h, int nIn, int nOut)
{
    void* h = scn_create(kind, hopSize, H, length_h, nIn, nOut);
    if (!h && !tl_err)
        set_tl_error(SAFCONV_ERR_ARG, "invalid argument%s: hopSize > 8192 needs an even numOvrlpAddBlocks * hopSize (the big-FFT engine serves these block sizes)", "");
    return h;
}

void saf_matrixConv_create(void** const phMC, int hopSize, float* H, int length_h, int nCHin, int nCHout, int usePartFLAG)
{
    /* both reference modes compute the same linear convolution; the partitioned engine serves both unless the true
     * big-FFT semantics of mode 0 are asked for (safconv_np.c) */
    if (!phMC) return;
    if (hopSize > SC_MAX_M) { *phMC = create_big_hop(SC_KIND_MATRIX, hopSize, H, length_h, nCHin, nCHout); return; }
    if (!usePartFLAG && scn_enabled()) {
        *phMC = scn_create(SC_KIND_MATRIX, hopSize, H, length_h, nCHin, nCHout);
        if (*phMC || tl_err) return;                     /* built, or failed for a real reason; odd fftSize falls through */
    }
    if (H && getenv("SAFCONV_DEVICES") && (*phMC = scm_create_from_env(SC_KIND_MATRIX, hopSize, H, length_h, nCHin, nCHout)) != NULL) return;
    const float* chunk = H;
    *phMC = conv_create(SC_KIND_MATRIX, hopSize, H ? &chunk : NULL, 1,
                        (size_t)(nCHout > 0 ? nCHout : 0) * (size_t)(nCHin > 0 ? nCHin : 0),
                        length_h, nCHin, nCHout, nCHout, 0, 0);
}

void safconv_matrixConv_create_device(void** const phMC, int hopSize, const float* d_H, int length_h, int nCHin, int nCHout)
{
    if (!phMC) return;
    if (hopSize > SC_MAX_M) {
        set_tl_error(SAFCONV_ERR_ARG, "invalid argument%s: safconv_matrixConv_create_device serves hopSize <= 8192", "");
        *phMC = NULL;
        return;
    }
    const float* chunk = d_H;
    tl_filters_on_device = 1;
    *phMC = conv_create(SC_KIND_MATRIX, hopSize, d_H ? &chunk : NULL, 1,
                        (size_t)(nCHout > 0 ? nCHout : 0) * (size_t)(nCHin > 0 ? nCHin : 0),
                        length_h, nCHin, nCHout, nCHout, 0, 0);
    tl_filters_on_device = 0;
}

void safconv_matrixConv_create_shard(void** const phMC, int hopSize, const float* H, int length_h,
                                     int nCHin, int nCHout, int outBegin, int outCount)
{
    if (!phMC) return;
    if (!H || outBegin < 0 || outCount < 1 || outBegin + outCount > nCHout || nCHin < 1 || length_h < 1) {
        set_tl_error(SAFCONV_ERR_ARG, "invalid shard%s", "");
        *phMC = NULL;
        return;
    }
    const float* chunk = H + (size_t)outBegin * nCHin * length_h;
    *phMC = conv_create(SC_KIND_MATRIX, hopSize, &chunk, 1, (size_t)outCount * nCHin,
                        length_h, nCHin, outCount, nCHout, outBegin, 0);
}

void safconv_matrixConv_create_from_shard(void** const phMC, int hopSize, const float* Hshard, int length_h,
                                          int nCHin, int nCHout, int outBegin, int outCount)
{
    if (!phMC) return;
    if (!Hshard || outBegin < 0 || outCount < 1 || outBegin + outCount > nCHout || nCHin < 1 || length_h < 1) {
        set_tl_error(SAFCONV_ERR_ARG, "invalid shard%s", "");
        *phMC = NULL;
        return;
    }
    const float* chunk = Hshard;
    *phMC = conv_create(SC_KIND_MATRIX, hopSize, &chunk, 1, (size_t)outCount * nCHin,
                        length_h, nCHin, outCount, nCHout, outBegin, 0);
}

void saf_matrixConv_destroy(void** const phMC) { conv_destroy(phMC); }

void saf_matrixConv_apply(void* const hMC, float* inputSigs, float* outputSigs)
{
    if (scm_is_multi(hMC)) { scm_apply(hMC, SC_KIND_MATRIX, inputSigs, outputSigs); return; }
    if (scn_is_np(hMC)) { scn_apply(hMC, SC_KIND_MATRIX, inputSigs, outputSigs); return; }
    safconv_handle* h = as_handle(hMC);
    if (!h || h->pl.kind != SC_KIND_MATRIX || !inputSigs || !outputSigs) return;
    conv_apply_host(h, inputSigs, outputSigs, 0);
}

void saf_multiConv_create(void** const phMC, int hopSize, float* H, int length_h, int nCH, int usePartFLAG)
{
    if (!phMC) return;
    if (hopSize > SC_MAX_M) { *phMC = create_big_hop(SC_KIND_MULTI, hopSize, H, length_h, nCH, nCH); return; }
    if (!usePartFLAG && scn_enabled()) {
        *phMC = scn_create(SC_KIND_MULTI, hopSize, H, length_h, nCH, nCH);
        if (*phMC || tl_err) return;
    }
    if (H && getenv("SAFCONV_DEVICES") && (*phMC = scm_create_from_env(SC_KIND_MULTI, hopSize, H, length_h, nCH, nCH)) != NULL) return;
    const float* chunk = H;
    *phMC = conv_create(SC_KIND_MULTI, hopSize, H ? &chunk : NULL, 1, (size_t)(nCH > 0 ? nCH : 0),
                        length_h, nCH, nCH, nCH, 0, 0);
}

void safconv_multiConv_create_shard(void** const phMC, int hopSize, const float* H, int length_h,
                                    int nCH, int chBegin, int chCount)
{
    if (!phMC) return;
    if (!H || chBegin < 0 || chCount < 1 || chBegin + chCount > nCH || length_h < 1) {
        set_tl_error(SAFCONV_ERR_ARG, "invalid shard%s", "");
        *phMC = NULL;
        return;
    }
    const float* chunk = H + (size_t)chBegin * length_h;
    *phMC = conv_create(SC_KIND_MULTI, hopSize, &chunk, 1, (size_t)chCount, length_h, chCount, chCount, nCH, chBegin, 0);
}

void saf_multiConv_destroy(void** const phMC) { conv_destroy(phMC); }

void saf_multiConv_apply(void* const hMC, float* inputSigs, float* outputSigs)
{
    if (scm_is_multi(hMC)) { scm_apply(hMC, SC_KIND_MULTI, inputSigs, outputSigs); return; }
    if (scn_is_np(hMC)) { scn_apply(hMC, SC_KIND_MULTI, inputSigs, outputSigs); return; }
    safconv_handle* h = as_handle(hMC);
    if (!h || h->pl.kind != SC_KIND_MULTI || !inputSigs || !outputSigs) return;
    conv_apply_host(h, inputSigs, outputSigs, 0);
}

void saf_TVConv_create(void** const phTVC, int hopSize, float** H, int length_h, int nIRs, int nCHout, int initIdx)
{
    if (!phTVC) return;
    if (!H || nIRs < 1 || nCHout < 1) { set_tl_error(SAFCONV_ERR_ARG, "invalid argument%s", ""); *phTVC = NULL; return; }
    for (int i = 0; i < nIRs; i++)
        if (!H[i]) { set_tl_error(SAFCONV_ERR_ARG, "H[i] is NULL%s", ""); *phTVC = NULL; return; }
    safconv_handle* h = conv_create(SC_KIND_TV, hopSize, (const float* const*)H, nIRs, (size_t)nCHout,
                                    length_h, 1, nCHout, nCHout, 0, nIRs);
    if (h) h->tvLast = h->tvLast2 = (initIdx >= 0 && initIdx < nIRs) ? initIdx : 0;   /* reference .c:459-465 */
    *phTVC = h;
}

void saf_TVConv_destroy(void** const phTVC) { conv_destroy(phTVC); }

void saf_TVConv_apply(void* const hTVC, float* inputSigs, float* outputSigs, int irIdx)
{
    safconv_handle* h = as_handle(hTVC);
    if (!h || h->pl.kind != SC_KIND_TV || !inputSigs || !outputSigs) return;
    h_clear(h);
    if (irIdx < 0 || irIdx >= h->pl.nIRs) { h_fail(h, SAFCONV_ERR_ARG, "irIdx out of range", 0); return; }
    conv_apply_host(h, inputSigs, outputSigs, irIdx);
}

/* ------------------------------------------------------------------------------------------ */
/*  extension API                                                                               */
/* ------------------------------------------------------------------------------------------ */

int safconv_last_error(void* hp)
{
    if (scm_is_multi(hp)) return scm_last_error(hp);
    if (scn_is_np(hp)) return scn_last_error(hp);
    safconv_handle* h = as_handle(hp);
    return h ? h->err : tl_err;
}

const char* safconv_last_error_string(void* hp)
{
    if (scm_is_multi(hp)) return scm_last_error_string(hp);
    if (scn_is_np(hp)) return scn_last_error_string(hp);
    safconv_handle* h = as_handle(hp);
    return h ? h->errmsg : tl_msg;
}

const char* safconv_version(void) { return SAFCONV_VERSION_STRING; }

int safconv_set_device(int device)
{
    int n = 0;
    if (scdev_device_count(&n) != 0 || device < 0 || device >= n) return SAFCONV_ERR_NO_DEVICE;
    tl_device = device;
    return SAFCONV_OK;
}

int safconv_apply_device(void* hp, const float* d_in, float* d_out)
{
    return safconv_apply_device_blocks(hp, d_in, d_out, 1);
}

int safconv_apply_device_blocks(void* hp, const float* d_in, float* d_out, int nBlocks)
{
    safconv_handle* h = as_handle_q(hp);
    if (!h || !d_in || !d_out || nBlocks < 1 || h->pl.kind == SC_KIND_TV) return SAFCONV_ERR_ARG;
    h_clear(h);
    int e = scdev_set_device(h->device);
    h->tailReady = 0;          /* a pre-computed tail belongs to the block counter it was enqueued for */
    const size_t inStride = (size_t)h->pl.nIn * h->pl.hop, outStride = (size_t)h->pl.nOutLocal * h->pl.hop;
    const int mb = (h->batching && h->pl.maxBatch > 1) ? h->pl.maxBatch : 1;
    for (int b = 0; b < nBlocks && !e; b += mb) {
        const int n = (nBlocks - b < mb) ? nBlocks - b : mb;
        e = enqueue_blocks(h, d_in + (size_t)b * inStride, d_out + (size_t)b * outStride, n);
    }
    if (e) return h_fail(h, SAFCONV_ERR_CUDA, "apply_device", e);
    h->count += (unsigned int)nBlocks;
    return SAFCONV_OK;
}

/* ---- the CONVOLVERS' FFT cores on their own (power-of-two N, saf_rfft conventions): test entry point, not in the public
 * header.  The public saf_rfft_* / safconv_rfft_* API (any even N) lives in safconv_rfft.c. ---- */
int safconv_debug_convolver_rfft(int N, int nBatch, const float* in, float* out, int dir)
{
    tl_err = 0; tl_msg[0] = 0;
    if (N < 64 || N > 2 * SC_MAX_M || (N & (N - 1)) || nBatch < 1 || !in || !out) {
        set_tl_error(SAFCONV_ERR_ARG, "rfft: need a power-of-two 64 <= N <= 16384, nBatch >= 1%s", "");
        return SAFCONV_ERR_ARG;
    }
    int ndev = 0;
    if (scdev_device_count(&ndev) != 0 || ndev < 1) {
        set_tl_error(SAFCONV_ERR_NO_DEVICE, "no usable CUDA device%s (libsafconv_b200 has no CPU fallback)", "");
        return SAFCONV_ERR_NO_DEVICE;
    }
    if (tl_device >= 0) scdev_set_device(tl_device);
    const int M = N / 2, logM = ilog2(M);
    const size_t inElems  = (size_t)nBatch * (dir == 0 ? (size_t)N : (size_t)2 * (M + 1));
    const size_t outElems = (size_t)nBatch * (dir == 0 ? (size_t)2 * (M + 1) : (size_t)N);
    float* tw = (float*)malloc(sizeof(float) * 2 * (size_t)M);
    void *stream = NULL, *d_tw = NULL; float *d_in = NULL, *d_out = NULL;
    int e = tw ? 0 : -1;
    if (!e) {
        for (int j = 0; j < M; j++) {                       /* evaluated in double like kiss_fft.c:358-364 */
            const double ph = -2.0 * 3.141592653589793238462643383279502884 * (double)j / (double)N;
            tw[2 * j] = (float)cos(ph); tw[2 * j + 1] = (float)sin(ph);
        }
        e = scdev_stream_create(&stream);
    }
    if (!e) e = scdev_malloc(&d_tw, sizeof(float) * 2 * (size_t)M);
    if (!e) e = scdev_malloc((void**)&d_in, sizeof(float) * inElems);
    if (!e) e = scdev_malloc((void**)&d_out, sizeof(float) * outElems);
    if (!e) e = scdev_memcpy_h2d_async(d_tw, tw, sizeof(float) * 2 * (size_t)M, stream);
    if (!e) e = scdev_memcpy_h2d_async(d_in, in, sizeof(float) * inElems, stream);
    if (!e) e = scdev_rfft(N, logM, nBatch, dir, d_in, d_out, d_tw, stream);
    if (!e) e = scdev_memcpy_d2h_async(out, d_out, sizeof(float) * outElems, stream);
    if (!e) e = scdev_stream_sync(stream); else if (stream) scdev_stream_sync(stream);
    scdev_free(d_tw); scdev_free(d_in); scdev_free(d_out);
    scdev_stream_destroy(stream);
    free(tw);
    if (e) return h_fail(NULL, e < 0 ? SAFCONV_ERR_NOMEM : SAFCONV_ERR_CUDA, "rfft", e < 0 ? 0 : e);
    return SAFCONV_OK;
}


/* ---- fftconv / fftfilt (reference saf_utility_fft.c:157-228) on the multiConv engine ---- */
static int fftconv_impl(const float* x, const float* h, int x_len, int h_len, int nCH, float* y, int keep)
{
    if (!x || !h || !y || x_len < 1 || h_len < 1 || nCH < 1) {
        set_tl_error(SAFCONV_ERR_ARG, "fftconv: invalid argument%s", "");
        return SAFCONV_ERR_ARG;
    }
    const long long y_len = (long long)x_len + h_len - 1;
    /* block size: the filter in one or a few partitions, 64 <= hop <= 1024 */
    int hop = 64;
    while (hop < h_len && hop < 1024) hop <<= 1;
    const long long nblk = (y_len + hop - 1) / hop;
    if (nblk > 0x7fffffff / 2) { set_tl_error(SAFCONV_ERR_ARG, "fftconv: signal too long%s", ""); return SAFCONV_ERR_ARG; }
    const float* chunk = h;
    safconv_handle* hd = conv_create(SC_KIND_MULTI, hop, &chunk, 1, (size_t)nCH, h_len, nCH, nCH, nCH, 0, 0);
    if (!hd) return tl_err ? tl_err : SAFCONV_ERR_CUDA;
    const size_t blkElems = (size_t)nCH * hop, total = (size_t)nblk * blkElems;
    float* hb = (float*)calloc(total, sizeof(float));           /* [nblk][nCH][hop], zero-padded */
    float *d_in = NULL, *d_out = NULL;
    int rc = SAFCONV_OK, e = 0;
    if (!hb) { rc = h_fail(NULL, SAFCONV_ERR_NOMEM, "fftconv: host staging", 0); goto done; }
    for (long long b = 0; b < nblk; b++) {
        const long long s0 = b * hop;
        if (s0 >= x_len) break;
        const size_t n = (size_t)((x_len - s0 < hop) ? x_len - s0 : hop);
        for (int c = 0; c < nCH; c++)
            memcpy(hb + (size_t)b * blkElems + (size_t)c * hop, x + (size_t)c * x_len + s0, n * sizeof(float));
    }
    e = scdev_malloc((void**)&d_in, total * sizeof(float));
    if (!e) e = scdev_malloc((void**)&d_out, total * sizeof(float));
    if (!e) e = scdev_memcpy_h2d_sync(d_in, hb, total * sizeof(float), hd->stream);
    if (e) { rc = h_fail(NULL, SAFCONV_ERR_CUDA, "fftconv: upload", e); goto done; }
    rc = safconv_apply_device_blocks(hd, d_in, d_out, (int)nblk);
    if (rc) { set_tl_error(rc, "fftconv: %.200s", hd->errmsg); goto done; }
    e = scdev_memcpy_d2h_async(hb, d_out, total * sizeof(float), hd->stream);
    if (!e) e = scdev_stream_sync(hd->stream);
    if (e) { rc = h_fail(NULL, SAFCONV_ERR_CUDA, "fftconv: download", e); goto done; }
    {
        const long long out_len = keep ? x_len : y_len;
        for (long long b = 0; b < nblk; b++) {
            const long long s0 = b * hop;
            if (s0 >= out_len) break;
            const size_t n = (size_t)((out_len - s0 < hop) ? out_len - s0 : hop);
            for (int c = 0; c < nCH; c++)
                memcpy(y + (size_t)c * out_len + s0, hb + (size_t)b * blkElems + (size_t)c * hop, n * sizeof(float));
        }
    }
done:
    scdev_free(d_in); scdev_free(d_out);
    free(hb);
    handle_free(hd);
    return rc;
}

int safconv_fftconv(const float* x, const float* h, int x_len, int h_len, int nCH, float* y)
{ return fftconv_impl(x, h, x_len, h_len, nCH, y, 0); }
int safconv_fftfilt(const float* x, const float* h, int x_len, int h_len, int nCH, float* y)
{ return fftconv_impl(x, h, x_len, h_len, nCH, y, 1); }
__attribute__((weak)) void fftconv(float* x, float* h, int x_len, int h_len, int nCH, float* y)
{ (void)fftconv_impl(x, h, x_len, h_len, nCH, y, 0); }
__attribute__((weak)) void fftfilt(float* x, float* h, int x_len, int h_len, int nCH, float* y)
{ (void)fftconv_impl(x, h, x_len, h_len, nCH, y, 1); }

/* ---- offline rendering: all frames at once, tensor-core per-bin contraction (safconv_offline.cu) ---- */
int safconv_render_offline_segment_device(void* hp, const float* d_in, float* d_out, int nFrames, int nHaloFrames)
{
    safconv_handle* h = as_handle_q(hp);
    if (!h || !d_in || !d_out || nFrames < 1 || nHaloFrames < 0 || h->pl.kind != SC_KIND_MATRIX) return SAFCONV_ERR_ARG;
    const int T = nFrames + nHaloFrames;
    int e = scdev_set_device(h->device);
    if (!e && !h->offEv[0]) for (int i = 0; i < 4 && !e; i++) e = scdev_event_create(&h->offEv[i]);
    if (!e) e = scdev_offline_prepare(&h->pl, &h->b, &h->off, T, h->stream);
    if (!e) e = scdev_offline_run(&h->pl, &h->b, &h->off, d_in, d_out, T, nHaloFrames, h->offEv, h->stream);
    if (e) return h_fail(h, SAFCONV_ERR_CUDA, "render_offline", e);
    return SAFCONV_OK;
}

int safconv_render_offline_device(void* hp, const float* d_in, float* d_out, int nFrames)
{
    return safconv_render_offline_segment_device(hp, d_in, d_out, nFrames, 0);
}

/* Host-buffer render.  Signals longer than one segment are rendered in TIME SEGMENTS (each with a P frame input
 * halo, bit-identical to the whole render up to the fp16 operand scale of the segment) through a three-stream
 * pipeline: the strided H2D copy of segment s+1 and the D2H copy of segment s-1 run beside the kernels of segment s,
 * on double-buffered device staging.  PCIe is full duplex, so the render costs about max(H2D, D2H, kernels) instead
 * of their sum -- and the device workspace is that of one segment.  (Overlap needs page-locked caller buffers;
 * pageable ones work, staged by the driver.) */
#define SC_OFF_SEG_TOTAL 512         /* frames per segment incl. halo: a multiple of the GEMM's 256-frame tiles */
int safconv_render_offline(void* hp, const float* in, float* out, int nFrames)
{
    return safconv_render_offline_segment(hp, in, out, nFrames, 0);
}

/* in [nIn][(nHaloFrames + nFrames) * hop] (the first nHaloFrames frames are history only), out [nOutLocal][nFrames * hop] */
int safconv_render_offline_segment(void* hp, const float* in, float* out, int nFrames, int nHaloFrames)
{
    safconv_handle* h = as_handle_q(hp);
    if (!h || !in || !out || nFrames < 1 || nHaloFrames < 0 || h->pl.kind != SC_KIND_MATRIX) return SAFCONV_ERR_ARG;
    int e = scdev_set_device(h->device);
    if (e) return h_fail(h, SAFCONV_ERR_CUDA, "cudaSetDevice", e);
    const size_t nIn = (size_t)h->pl.nIn, nOut = (size_t)h->pl.nOutLocal, hop = (size_t)h->pl.hop;
    const size_t len = (size_t)nFrames * hop;                    /* output row length */
    const size_t lenIn = (size_t)(nFrames + nHaloFrames) * hop;  /* input row length  */
    int rc = SAFCONV_OK;
    /* output frame t = first half of block t + second half of block t-1 (overlap-add), and block t-1 reaches back
     * to input frame t-P: P history frames per segment */
    const int halo = h->pl.P;
    const int pipelined = env_int("SAFCONV_OFF_PIPELINE", 1, 0, 1) && nFrames > SC_OFF_SEG_TOTAL && halo < SC_OFF_SEG_TOTAL / 2;
    if (!pipelined) {
        float *d_in = NULL, *d_out = NULL;
        e = scdev_malloc((void**)&d_in, sizeof(float) * nIn * lenIn);
        if (!e) e = scdev_malloc((void**)&d_out, sizeof(float) * nOut * len);
        if (!e) e = scdev_memcpy_h2d_async(d_in, in, sizeof(float) * nIn * lenIn, h->stream);
        if (!e) rc = safconv_render_offline_segment_device(hp, d_in, d_out, nFrames, nHaloFrames);
        if (!e && !rc) e = scdev_memcpy_d2h_async(out, d_out, sizeof(float) * nOut * len, h->stream);
        if (!e && !rc) e = scdev_stream_sync(h->stream); else scdev_stream_sync(h->stream);
        scdev_free(d_in); scdev_free(d_out);
        if (e) return h_fail(h, SAFCONV_ERR_CUDA, "render_offline (host buffers)", e);
        return rc;
    }
    if (!h->offIn)  e = scdev_stream_create(&h->offIn);
    if (!e && !h->offOut) e = scdev_stream_create(&h->offOut);
    for (int i = 0; i < 6 && !e; i++) if (!h->offPipeEv[i]) e = scdev_event_create_sync(&h->offPipeEv[i]);
    if (!e) e = scdev_offline_prepare(&h->pl, &h->b, &h->off, SC_OFF_SEG_TOTAL, h->stream);     /* one workspace for all segments */
    if (e) return h_fail(h, SAFCONV_ERR_CUDA, "render_offline (pipeline setup)", e);
    const int seg = SC_OFF_SEG_TOTAL - halo;                    /* new frames per segment */
    float *d_in[2] = { NULL, NULL }, *d_out[2] = { NULL, NULL };
    for (int b = 0; b < 2 && !e; b++) {                         /* sizes depend on the handle only: allocated once */
        if (!h->offStageIn[b])  e = scdev_malloc((void**)&h->offStageIn[b], sizeof(float) * nIn * SC_OFF_SEG_TOTAL * hop);
        if (!e && !h->offStageOut[b]) e = scdev_malloc((void**)&h->offStageOut[b], sizeof(float) * nOut * SC_OFF_SEG_TOTAL * hop);
        d_in[b] = h->offStageIn[b]; d_out[b] = h->offStageOut[b];
    }
    int s = 0;
    for (int t0 = 0; t0 < nFrames && !e && !rc; t0 += seg, s++) {
        const int b = s & 1;
        const int t1 = (t0 + seg < nFrames) ? t0 + seg : nFrames;
        const int hl = (t0 + nHaloFrames < halo) ? t0 + nHaloFrames : halo;     /* history frames that exist */
        const size_t inW = (size_t)(t1 - t0 + hl) * hop, outW = (size_t)(t1 - t0) * hop;
        void *evIn = h->offPipeEv[b], *evR = h->offPipeEv[2 + b], *evOut = h->offPipeEv[4 + b];
        /* input of segment s into slot b: the render of segment s-2 must be done with it */
        e = scdev_stream_wait_event(h->offIn, evR);
        if (!e) e = scdev_memcpy2d_async(d_in[b], inW * sizeof(float), in + (size_t)(nHaloFrames + t0 - hl) * hop, lenIn * sizeof(float),
                                         inW * sizeof(float), nIn, 0, h->offIn);
        if (!e) e = scdev_event_record(evIn, h->offIn);
        /* render: input landed, and the copy-back of segment s-2 has drained the output slot */
        if (!e) e = scdev_stream_wait_event(h->stream, evIn);
        if (!e) e = scdev_stream_wait_event(h->stream, evOut);
        if (!e) rc = safconv_render_offline_segment_device(hp, d_in[b], d_out[b], t1 - t0, hl);
        if (!e && !rc) e = scdev_event_record(evR, h->stream);
        /* copy back */
        if (!e && !rc) e = scdev_stream_wait_event(h->offOut, evR);
        if (!e && !rc) e = scdev_memcpy2d_async(out + (size_t)t0 * hop, len * sizeof(float), d_out[b], outW * sizeof(float),
                                                outW * sizeof(float), nOut, 1, h->offOut);
        if (!e && !rc) e = scdev_event_record(evOut, h->offOut);
    }
    {
        int e2 = scdev_stream_sync(h->offIn);
        int e3 = scdev_stream_sync(h->stream);
        int e4 = scdev_stream_sync(h->offOut);
        if (!e) e = e2 ? e2 : (e3 ? e3 : e4);
    }
    if (e) return h_fail(h, SAFCONV_ERR_CUDA, "render_offline (pipelined)", e);
    return rc;
}

int safconv_get_offline_times(void* hp, float ms[3])
{
    safconv_handle* h = as_handle(hp);
    if (!h || !ms || !h->offEv[0]) return SAFCONV_ERR_ARG;
    scdev_set_device(h->device);
    int e = scdev_stream_sync(h->stream);
    for (int i = 0; i < 3 && !e; i++) e = scdev_event_elapsed_ms(h->offEv[i], h->offEv[i + 1], &ms[i]);
    return e ? h_fail(h, SAFCONV_ERR_CUDA, "offline timing", e) : SAFCONV_OK;
}

int safconv_set_stream(void* hp, void* cudaStream)
{
    safconv_handle* h = as_handle_q(hp);
    if (!h) return SAFCONV_ERR_ARG;
    scdev_set_device(h->device);
    scdev_stream_sync(h->stream);
    h->tailReady = 0;
    if (h->graphExec) { scdev_graph_destroy(h->graphExec); h->graphExec = NULL; }
    h->stream = cudaStream ? cudaStream : h->streamOwn;
    return SAFCONV_OK;
}

void* safconv_get_stream(void* hp)
{
    safconv_handle* h = as_handle(hp);
    return h ? h->stream : NULL;
}

int safconv_synchronize(void* hp)
{
    if (scm_is_multi(hp)) return scm_synchronize(hp);
    if (scn_is_np(hp)) return SAFCONV_OK;                /* its apply is synchronous and nothing else enqueues work */
    safconv_handle* h = as_handle_q(hp);
    if (!h) return SAFCONV_ERR_ARG;
    scdev_set_device(h->device);
    int e = scdev_stream_sync(h->stream);
    return e ? h_fail(h, SAFCONV_ERR_CUDA, "synchronize", e) : SAFCONV_OK;
}

int safconv_reset_state(void* hp)
{
    if (scm_is_multi(hp)) return scm_reset_state(hp);
    if (scn_is_np(hp)) return scn_reset_state(hp);
    safconv_handle* h = as_handle_q(hp);
    if (!h) return SAFCONV_ERR_ARG;
    scdev_set_device(h->device);
    h->tailReady = 0;
    int e = scdev_memset_async(h->b.X, 0, h->bytesX, h->stream);
    if (!e) e = scdev_memset_async(h->b.tail, 0, sizeof(float) * (size_t)h->pl.nOutLocal * h->pl.hop, h->stream);
    if (!e && h->b.tail2) e = scdev_memset_async(h->b.tail2, 0, sizeof(float) * (size_t)h->pl.nOutLocal * h->pl.hop, h->stream);
    if (!e) e = scdev_memset_async(h->b.counters, 0, 4 * sizeof(unsigned int), h->stream);
    if (!e) e = scdev_stream_sync(h->stream);
    h->count = 0;
    return e ? h_fail(h, SAFCONV_ERR_CUDA, "reset_state", e) : SAFCONV_OK;
}

int safconv_get_info(void* hp, safconv_info* info)
{
    if (scm_is_multi(hp)) return scm_get_info(hp, info);
    if (scn_is_np(hp)) return scn_get_info(hp, info);
    safconv_handle* h = as_handle(hp);
    if (!h || !info) return SAFCONV_ERR_ARG;
    const scdev_plan* pl = &h->pl;
    memset(info, 0, sizeof *info);
    info->kind = pl->kind; info->hopSize = pl->hop; info->length_h = pl->len;
    info->nCHin = pl->nIn; info->nCHout = h->nCHoutTotal; info->nOutLocal = pl->nOutLocal; info->outBegin = h->outBegin;
    info->fftSize = pl->N; info->nBinsPacked = pl->M; info->numFilterBlocks = pl->P;
    info->macGrid = pl->macGrid; info->macStages = pl->macStages; info->macThreads = (SC_MAC_CWARPS + 1) * 32;
    info->maxBatch = (pl->kind == SC_KIND_TV || pl->maxBatch < 1) ? 1 : pl->maxBatch;
    info->device = h->device;
    info->bytesFilters = h->bytesH; info->bytesDelayLine = h->bytesX;
    /* algorithmic bytes per block, nBins = hop + 1 complex bins of 8 bytes */
    const double nb = (double)pl->hop + 1.0, P = pl->P, hop = pl->hop;
    if (pl->kind == SC_KIND_MATRIX) {
        const double nIn = pl->nIn, nOut = pl->nOutLocal;
        info->algBytesPerBlock    = 8.0 * nb * nIn * (nOut * P + P + 1.0) + 4.0 * hop * (nIn + nOut) + 8.0 * hop * nOut;
        info->macAlgBytesPerBlock = 8.0 * nb * nIn * (nOut * P + P);
    } else if (pl->kind == SC_KIND_MULTI) {
        const double nCH = pl->nOutLocal;
        info->algBytesPerBlock    = 8.0 * nb * nCH * (2.0 * P + 1.0) + 4.0 * hop * nCH * 4.0;
        info->macAlgBytesPerBlock = 8.0 * nb * nCH * 2.0 * P;
    }
    return SAFCONV_OK;
}

int safconv_enable_kernel_timing(void* hp, int nGroups)
{
    safconv_handle* h = as_handle_q(hp);
    if (!h || nGroups < 0) return SAFCONV_ERR_ARG;
    scdev_set_device(h->device);
    scdev_stream_sync(h->stream);
    h->tailReady = 0;
    if (h->evRing) {
        for (int i = 0; i < 4 * h->timingCap; i++) scdev_event_destroy(h->evRing[i]);
        free(h->evRing);
        h->evRing = NULL;
    }
    free(h->evBlocks); h->evBlocks = NULL;
    h->timingCap = h->timingCount = 0;
    if (nGroups == 0) return SAFCONV_OK;
    h->evRing = (void**)calloc((size_t)4 * nGroups, sizeof(void*));
    h->evBlocks = (int*)calloc((size_t)nGroups, sizeof(int));
    if (!h->evRing || !h->evBlocks) return h_fail(h, SAFCONV_ERR_NOMEM, "event ring", 0);
    for (int i = 0; i < 4 * nGroups; i++) {
        int e = scdev_event_create(&h->evRing[i]);
        if (e) return h_fail(h, SAFCONV_ERR_CUDA, "cudaEventCreate", e);
    }
    h->timingCap = nGroups;
    return SAFCONV_OK;
}

int safconv_get_kernel_totals(void* hp, float msTotal[3], int* nLaunchGroups, int* nBlocksOut)
{
    safconv_handle* h = as_handle(hp);
    if (!h || !msTotal || !h->timingCap) return SAFCONV_ERR_ARG;
    scdev_set_device(h->device);
    msTotal[0] = msTotal[1] = msTotal[2] = 0.f;
    const int n = h->timingCount;
    if (nLaunchGroups) *nLaunchGroups = n;
    if (nBlocksOut) *nBlocksOut = 0;
    if (n == 0) return SAFCONV_OK;
    const int matrix = (h->pl.kind == SC_KIND_MATRIX);
    int e = scdev_stream_sync(h->stream);
    double acc[3] = { 0, 0, 0 };
    int blocks = 0;
    for (int g = 0; g < n && !e; g++) {
        void** ev = h->evRing + 4 * (size_t)g;
        float t = 0.f;
        const int full = matrix || h->evBlocks[g] > 1;          /* multiConv: only batched groups record all four events */
        if (full) { e = scdev_event_elapsed_ms(ev[0], ev[1], &t); acc[0] += t; }
        if (!e) { e = scdev_event_elapsed_ms(ev[1], ev[2], &t); acc[1] += t; }
        if (!e && full) { e = scdev_event_elapsed_ms(ev[2], ev[3], &t); acc[2] += t; }
        blocks += h->evBlocks[g];
    }
    h->timingCount = 0;
    if (e) return h_fail(h, SAFCONV_ERR_CUDA, "kernel timing", e);
    for (int i = 0; i < 3; i++) msTotal[i] = (float)acc[i];
    if (nBlocksOut) *nBlocksOut = blocks;
    return SAFCONV_OK;
}

int safconv_get_kernel_times(void* hp, float ms[3], int* nBlocksOut)
{
    int groups = 0, blocks = 0;
    int rc = safconv_get_kernel_totals(hp, ms, &groups, &blocks);
    if (rc) return rc;
    if (nBlocksOut) *nBlocksOut = blocks;
    if (blocks > 0) for (int i = 0; i < 3; i++) ms[i] /= (float)blocks;
    return SAFCONV_OK;
}

int safconv_set_option(void* hp, const char* name, int value)
{
    if (scm_is_multi(hp)) return scm_set_option(hp, name, value);
    safconv_handle* h = as_handle_q(hp);
    if (!h || !name) return SAFCONV_ERR_ARG;
    if (!strcmp(name, "mac_hints")) { h->pl.macHints = (value < 0 || value > 2) ? 1 : value; }
    else if (!strcmp(name, "use_graph")) { h->useGraph = value ? 1 : 0; }
    else if (!strcmp(name, "batching")) { h->batching = value ? 1 : 0; }
    else if (!strcmp(name, "small_fused")) { h->smallFused = value ? 1 : 0; }
    else if (!strcmp(name, "flag_wait")) { h->flagWait = value ? 1 : 0; }
    else if (!strcmp(name, "resident_us")) { h->residentUs = value < 0 ? 0 : (value > 2000000 ? 2000000 : value); }
    else if (!strcmp(name, "detect_pinned")) { h->detectPinned = value ? 1 : 0; }
    else if (!strcmp(name, "lookahead")) { h->lookahead = (value && h->tailPass.Zp) ? 1 : 0; }
    else return SAFCONV_ERR_ARG;
    h->tailReady = 0;
    if (h->graphExec) { scdev_set_device(h->device); scdev_stream_sync(h->stream); scdev_graph_destroy(h->graphExec); h->graphExec = NULL; }
    return SAFCONV_OK;
}
