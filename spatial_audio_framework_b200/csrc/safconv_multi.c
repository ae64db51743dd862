/*
 * safconv_multi.c -- ONE convolver handle spanning several GPUs of a box (single host process, C only)
 *
 * Output channels of the matrix convolver are independent in the reference (each iteration of the `no` loop of
 * saf_matrixConv_apply reads its own filters and overlap tail and the shared delay line:
 * /root/reference/framework/modules/saf_utilities/saf_utility_matrixConv.c:218-234), multiConv channels are fully
 * independent (.c:388-413).  So device g of G owns a contiguous range of output channels (channels): its rows of the
 * filter spectra, its overlap tails and -- matrix only -- a full replica of the frequency-domain delay line.  No
 * partial sums cross GPUs.
 *
 * The handle is what the UNCHANGED drop-in calls operate on: saf_matrixConv_apply(h, in, out) / saf_multiConv_apply
 * with host pointers, synchronous, one block per call -- a SAF host (examples/src/matrixconv/matrixconv.c:142) uses N
 * GPUs without a source change (safconv_matrixConv_create_multi, or SAFCONV_DEVICES=... with the plain create).
 *
 * Threads: one persistent worker per device owns that device's shard handle (an ordinary single-device handle of
 * safconv_host.c, look-ahead apply included).  apply() posts the block to all workers (a sequence number they spin
 * on, then sleep on a condition variable) and waits for all of them -- launching N devices' kernels from one thread
 * would serialise ~5 launches x N on the caller (> the 60 us a block takes per GPU at N = 8).
 *
 * Transports (option "transport"):
 *   0 host  every device's K1 reads the page-locked input block straight over its own PCIe link and K3 writes its
 *           rows of the page-locked output block: no exchange between GPUs at all (default)
 *   1 nccl  the exchange the north star names: device 0 uploads the block, ncclBroadcast to all devices over NVLink,
 *           per-device K1 + MAC + K3 into a device buffer, ncclSend/ncclRecv gather of the nOut/G x hop shards
 *           straight into a channel-major buffer on device 0, one D2H copy -- all on a per-device side stream.
 *           NCCL is loaded with dlopen at first use (libnccl.so.2), so single-GPU hosts never need it.
 */
#define _GNU_SOURCE
#include "safconv_host_internal.h"

#include <dlfcn.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define SCM_MAX_DEV 16
enum { SCM_T_HOST = 0, SCM_T_NCCL = 1 };
enum { SCM_OP_NONE = 0, SCM_OP_CREATE, SCM_OP_APPLY, SCM_OP_RESET, SCM_OP_SYNC, SCM_OP_OPTION, SCM_OP_NCCL_SETUP, SCM_OP_QUIT };

/* the few NCCL entry points the nccl transport needs, resolved with dlsym (ABI of NCCL 2.x: ncclFloat32 == 7) */
typedef struct ncclComm* scm_ncclComm_t;
#define SCM_NCCL_FLOAT32 7
typedef struct scm_nccl {
    void* lib;
    int (*CommInitAll)(scm_ncclComm_t*, int, const int*);
    int (*CommDestroy)(scm_ncclComm_t);
    int (*Broadcast)(const void*, void*, size_t, int, int, scm_ncclComm_t, void*);
    int (*Send)(const void*, size_t, int, int, scm_ncclComm_t, void*);
    int (*Recv)(void*, size_t, int, int, scm_ncclComm_t, void*);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    const char* (*GetErrorString)(int);
} scm_nccl;

struct safconv_multi;

typedef struct scm_worker {
    struct safconv_multi* m;
    int idx, device;
    int begin, count;                 /* output channels (matrix) / channels (multi) of this device */
    safconv_handle* shard;
    pthread_t th;
    int started;
    atomic_uint req, done;            /* sequence numbers: posted / finished */
    atomic_int sleeping;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    int err;
    char errmsg[256];
    /* nccl transport */
    scm_ncclComm_t comm;
    void *commStream, *evB, *evK, *evFence;
    float* d_inAll;                   /* the whole input block on this device */
} scm_worker;

typedef struct safconv_multi {
    uint32_t magic;
    int err;
    char errmsg[256];
    int kind, nDev, hop, len, nIn, nOut;
    int transport;
    int spinUs;                       /* how long an idle worker spins before it sleeps */
    int detectPinned;
    size_t inBytes, outBytes;
    float *h_in, *h_out;              /* page-locked staging for pageable caller buffers */
    float* d_outAll;                  /* nccl transport, device 0: channel-major [nOut][hop] */
    unsigned int seq;
    /* the job the workers pick up */
    int op;
    const float* src;
    float* dst;
    const float* H;                   /* create only */
    const char* optName; int optValue;
    scm_nccl nccl;
    int ncclReady;
    scm_worker w[SCM_MAX_DEV];
} safconv_multi;

static safconv_multi* as_multi(const void* p)
{
    const safconv_multi* m = (const safconv_multi*)p;
    return (m && m->magic == SAFCONV_MAGIC_MULTI) ? (safconv_multi*)m : NULL;
}

int scm_is_multi(const void* p) { return as_multi(p) != NULL; }

static int m_fail(safconv_multi* m, int code, const char* what, const char* detail)
{
    char buf[256];
    snprintf(buf, sizeof buf, "%s%s%s", what, detail ? ": " : "", detail ? detail : "");
    if (m) { m->err = code; snprintf(m->errmsg, sizeof m->errmsg, "%s", buf); }
    sch_set_tl_error(code, "%s", buf);
    return code;
}

static void w_fail(scm_worker* w, int code, const char* what, int cudaErr)
{
    w->err = code;
    if (cudaErr) snprintf(w->errmsg, sizeof w->errmsg, "device %d: %s: %s", w->device, what, scdev_error_string(cudaErr));
    else         snprintf(w->errmsg, sizeof w->errmsg, "device %d: %s", w->device, what);
}

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e6 * (double)ts.tv_sec + 1e-3 * (double)ts.tv_nsec;
}

static inline void cpu_relax(void)
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    sched_yield();
#endif
}

/* ------------------------------------------------------------------------------------------ */
/*  NCCL (dlopen)                                                                               */
/* ------------------------------------------------------------------------------------------ */
static int nccl_load(safconv_multi* m)
{
    scm_nccl* n = &m->nccl;
    if (n->lib) return 0;
    const char* names[] = { "libnccl.so.2", "libnccl.so", NULL };
    for (int i = 0; names[i] && !n->lib; i++) n->lib = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
    if (!n->lib) return m_fail(m, SAFCONV_ERR_NO_DEVICE, "nccl transport: cannot load libnccl.so.2", dlerror());
#define SCM_SYM(field, name) do { *(void**)(&n->field) = dlsym(n->lib, name); \
        if (!n->field) { dlclose(n->lib); n->lib = NULL; return m_fail(m, SAFCONV_ERR_NO_DEVICE, "nccl transport: missing symbol", name); } } while (0)
    SCM_SYM(CommInitAll, "ncclCommInitAll");
    SCM_SYM(CommDestroy, "ncclCommDestroy");
    SCM_SYM(Broadcast, "ncclBroadcast");
    SCM_SYM(Send, "ncclSend");
    SCM_SYM(Recv, "ncclRecv");
    SCM_SYM(GroupStart, "ncclGroupStart");
    SCM_SYM(GroupEnd, "ncclGroupEnd");
    SCM_SYM(GetErrorString, "ncclGetErrorString");
#undef SCM_SYM
    return 0;
}

#define W_CUDA(w, call, what) do { if (!(w)->err) { int e__ = (call); if (e__) w_fail((w), SAFCONV_ERR_CUDA, (what), e__); } } while (0)
#define W_NCCL(w, call, what) do { if (!(w)->err) { int r__ = (call); if (r__) { (w)->err = SAFCONV_ERR_CUDA; \
        snprintf((w)->errmsg, sizeof (w)->errmsg, "device %d: %s: %s", (w)->device, (what), (w)->m->nccl.GetErrorString(r__)); } } } while (0)

/* ------------------------------------------------------------------------------------------ */
/*  worker                                                                                      */
/* ------------------------------------------------------------------------------------------ */
static void worker_take_shard_error(scm_worker* w)
{
    if (w->shard && w->shard->err && !w->err) {
        w->err = w->shard->err;
        snprintf(w->errmsg, sizeof w->errmsg, "device %d: %.200s", w->device, w->shard->errmsg);
    }
}

static void worker_create(scm_worker* w)
{
    safconv_multi* m = w->m;
    if (safconv_set_device(w->device) != SAFCONV_OK) { w_fail(w, SAFCONV_ERR_NO_DEVICE, "no such CUDA device", 0); return; }
    const float* chunk;
    size_t rows;
    int nInLocal;
    if (m->kind == SC_KIND_MATRIX) { chunk = m->H + (size_t)w->begin * m->nIn * m->len; rows = (size_t)w->count * m->nIn; nInLocal = m->nIn; }
    else                           { chunk = m->H + (size_t)w->begin * m->len;          rows = (size_t)w->count;          nInLocal = w->count; }
    w->shard = sch_conv_create(m->kind, m->hop, &chunk, 1, rows, m->len, nInLocal, w->count, m->nOut, w->begin, 0);
    if (!w->shard) {
        w->err = safconv_last_error(NULL);
        if (!w->err) w->err = SAFCONV_ERR_CUDA;
        snprintf(w->errmsg, sizeof w->errmsg, "device %d: %.200s", w->device, safconv_last_error_string(NULL));
    }
}

/* nccl transport: per-device side stream, events, the whole-input buffer (device 0 also: the gathered output) */
static void worker_nccl_setup(scm_worker* w)
{
    safconv_multi* m = w->m;
    W_CUDA(w, scdev_set_device(w->device), "cudaSetDevice");
    if (!w->commStream) W_CUDA(w, scdev_stream_create_high_priority(&w->commStream), "cudaStreamCreate");
    if (!w->evB) W_CUDA(w, scdev_event_create_sync(&w->evB), "cudaEventCreate");
    if (!w->evK) W_CUDA(w, scdev_event_create_sync(&w->evK), "cudaEventCreate");
    if (!w->evFence) W_CUDA(w, scdev_event_create_sync(&w->evFence), "cudaEventCreate");
    if (!w->d_inAll) W_CUDA(w, scdev_malloc((void**)&w->d_inAll, m->inBytes), "input block buffer");
    if (w->idx == 0 && !m->d_outAll) W_CUDA(w, scdev_malloc((void**)&m->d_outAll, m->outBytes), "gathered output buffer");
}

static void worker_apply_nccl(scm_worker* w)
{
    safconv_multi* m = w->m;
    safconv_handle* h = w->shard;
    const scm_nccl* n = &m->nccl;
    const size_t hop = (size_t)m->hop;
    h->err = SAFCONV_OK; h->errmsg[0] = 0;
    W_CUDA(w, scdev_set_device(w->device), "cudaSetDevice");
    /* input: device 0 uploads, everyone takes part in the broadcast (NVLink) */
    if (w->idx == 0) W_CUDA(w, scdev_memcpy_h2d_async(w->d_inAll, m->src, m->inBytes, w->commStream), "input upload");
    W_NCCL(w, n->Broadcast(w->d_inAll, w->d_inAll, m->inBytes / sizeof(float), SCM_NCCL_FLOAT32, 0, w->comm, w->commStream), "ncclBroadcast");
    W_CUDA(w, scdev_event_record(w->evB, w->commStream), "cudaEventRecord");
    /* this device's rows: device 0 writes them straight into the gathered buffer */
    float* d_out = (w->idx == 0) ? m->d_outAll + (size_t)w->begin * hop : h->d_out;
    const float* d_in = (m->kind == SC_KIND_MATRIX) ? w->d_inAll : w->d_inAll + (size_t)w->begin * hop;
    void* evOut;
    if (w->err) return;
    if (sch_uses_lookahead(h)) {
        sch_la_io io;
        io.k1src = d_in; io.h2dSrc = NULL; io.evSrc = w->evB; io.kout = d_out; io.d2hDst = NULL; io.sync = 0;
        W_CUDA(w, sch_apply_lookahead_io(h, &io), "apply (look-ahead)");
        evOut = h->evDone;
    } else {
        W_CUDA(w, scdev_stream_wait_event(h->stream, w->evB), "cudaStreamWaitEvent");
        if (!w->err && safconv_apply_device_blocks(h, d_in, d_out, 1)) { worker_take_shard_error(w); return; }
        W_CUDA(w, scdev_event_record(w->evK, h->stream), "cudaEventRecord");
        evOut = w->evK;
    }
    W_CUDA(w, scdev_stream_wait_event(w->commStream, evOut), "cudaStreamWaitEvent");
    /* output: shards travel to device 0 (point-to-point, grouped on the root), one D2H copy of the whole block */
    if (w->idx == 0) {
        W_NCCL(w, n->GroupStart(), "ncclGroupStart");
        for (int r = 1; r < m->nDev; r++)
            W_NCCL(w, n->Recv(m->d_outAll + (size_t)m->w[r].begin * hop, (size_t)m->w[r].count * hop, SCM_NCCL_FLOAT32, r, w->comm, w->commStream), "ncclRecv");
        W_NCCL(w, n->GroupEnd(), "ncclGroupEnd");
        W_CUDA(w, scdev_memcpy_d2h_async(m->dst, m->d_outAll, m->outBytes, w->commStream), "output download");
    } else {
        W_NCCL(w, n->Send(h->d_out, (size_t)w->count * hop, SCM_NCCL_FLOAT32, 0, w->comm, w->commStream), "ncclSend");
    }
    W_CUDA(w, scdev_stream_sync(w->commStream), "cudaStreamSynchronize");
}

static void worker_apply(scm_worker* w)
{
    safconv_multi* m = w->m;
    if (m->transport == SCM_T_NCCL) { worker_apply_nccl(w); return; }
    const size_t off = (size_t)w->begin * m->hop;
    const float* src = (m->kind == SC_KIND_MATRIX) ? m->src : m->src + off;
    w->shard->err = SAFCONV_OK; w->shard->errmsg[0] = 0;
    sch_apply_pinned(w->shard, src, m->dst + off, 0);
    worker_take_shard_error(w);
}

static void worker_free(scm_worker* w)
{
    if (w->device >= 0) scdev_set_device(w->device);
    if (w->commStream) scdev_stream_sync(w->commStream);
    if (w->shard) { sch_handle_free(w->shard); w->shard = NULL; }
    if (w->comm && w->m->nccl.CommDestroy) { w->m->nccl.CommDestroy(w->comm); w->comm = NULL; }
    scdev_event_destroy(w->evB); scdev_event_destroy(w->evK); scdev_event_destroy(w->evFence);
    scdev_stream_destroy(w->commStream);
    scdev_free(w->d_inAll);
    if (w->idx == 0) { scdev_free(w->m->d_outAll); w->m->d_outAll = NULL; }
    w->evB = w->evK = w->evFence = w->commStream = NULL; w->d_inAll = NULL;
}

static void* worker_main(void* arg)
{
    scm_worker* w = (scm_worker*)arg;
    safconv_multi* m = w->m;
    unsigned int last = 0;
    for (;;) {
        /* wait for the next job: spin for a while (back-to-back blocks arrive within microseconds), then sleep */
        unsigned int r;
        const double t0 = now_us();
        int spins = 0;
        while ((r = atomic_load_explicit(&w->req, memory_order_acquire)) == last) {
            cpu_relax();
            if ((++spins & 63) == 0 && now_us() - t0 > (double)m->spinUs) {
                pthread_mutex_lock(&w->mu);
                atomic_store(&w->sleeping, 1);
                while ((r = atomic_load(&w->req)) == last) pthread_cond_wait(&w->cv, &w->mu);
                atomic_store(&w->sleeping, 0);
                pthread_mutex_unlock(&w->mu);
                break;
            }
        }
        last = r;
        const int op = m->op;
        switch (op) {
            case SCM_OP_CREATE:     worker_create(w); break;
            case SCM_OP_APPLY:      worker_apply(w); break;
            case SCM_OP_NCCL_SETUP: worker_nccl_setup(w); break;
            case SCM_OP_RESET:      if (safconv_reset_state(w->shard)) worker_take_shard_error(w); break;
            case SCM_OP_SYNC:       if (safconv_synchronize(w->shard)) worker_take_shard_error(w); break;
            case SCM_OP_OPTION:     if (safconv_set_option(w->shard, m->optName, m->optValue)) w_fail(w, SAFCONV_ERR_ARG, "unknown option", 0); break;
            case SCM_OP_QUIT:       worker_free(w); break;
            default: break;
        }
        atomic_store_explicit(&w->done, r, memory_order_release);
        if (op == SCM_OP_QUIT) return NULL;
    }
}

/* post `op` to all workers and wait for them; returns the first worker error (0 = none) */
static int run_op(safconv_multi* m, int op)
{
    m->op = op;
    const unsigned int seq = ++m->seq;
    for (int i = 0; i < m->nDev; i++) {
        scm_worker* w = &m->w[i];
        if (!w->started) continue;
        w->err = 0; w->errmsg[0] = 0;
        atomic_store_explicit(&w->req, seq, memory_order_seq_cst);
        if (atomic_load(&w->sleeping)) {
            pthread_mutex_lock(&w->mu);
            pthread_cond_signal(&w->cv);
            pthread_mutex_unlock(&w->mu);
        }
    }
    int rc = 0;
    for (int i = 0; i < m->nDev; i++) {
        scm_worker* w = &m->w[i];
        if (!w->started) continue;
        int spins = 0;
        const double t0 = now_us();
        while (atomic_load_explicit(&w->done, memory_order_acquire) != seq) {
            cpu_relax();
            if ((++spins & 255) == 0 && now_us() - t0 > 5000.0) sched_yield();    /* long jobs (create): be polite */
        }
        if (w->err && !rc) { rc = w->err; m->err = w->err; snprintf(m->errmsg, sizeof m->errmsg, "%s", w->errmsg); }
    }
    if (rc) sch_set_tl_error(rc, "%s", m->errmsg);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/*  create / destroy                                                                            */
/* ------------------------------------------------------------------------------------------ */
void scm_destroy(void** pp)
{
    if (!pp) return;
    safconv_multi* m = as_multi(*pp);
    *pp = NULL;
    if (!m) return;
    run_op(m, SCM_OP_QUIT);
    for (int i = 0; i < m->nDev; i++) {
        scm_worker* w = &m->w[i];
        if (w->started) { pthread_join(w->th, NULL); pthread_mutex_destroy(&w->mu); pthread_cond_destroy(&w->cv); }
    }
    scdev_host_free(m->h_in); scdev_host_free(m->h_out);
    if (m->nccl.lib) dlclose(m->nccl.lib);
    m->magic = 0;
    free(m);
}

static void* multi_create(int kind, int hop, const float* H, int len, int nIn, int nOut, const int* devices, int nDev)
{
    sch_set_tl_error(SAFCONV_OK, "%s", "");
    if (!H || !devices || nDev < 1 || nDev > SCM_MAX_DEV || hop < 1 || len < 1 || nIn < 1 || nOut < 1) {
        sch_set_tl_error(SAFCONV_ERR_ARG, "create_multi: invalid argument%s (1..16 devices, H != NULL, sizes >= 1)", "");
        return NULL;
    }
    int ndevAvail = 0;
    if (scdev_device_count(&ndevAvail) != 0 || ndevAvail < 1) {
        sch_set_tl_error(SAFCONV_ERR_NO_DEVICE, "no usable CUDA device%s (libsafconv_b200 has no CPU fallback)", "");
        return NULL;
    }
    for (int i = 0; i < nDev; i++) {
        int dup = 0;
        for (int j = 0; j < i; j++) dup |= (devices[j] == devices[i]);
        if (devices[i] < 0 || devices[i] >= ndevAvail || dup) {
            sch_set_tl_error(SAFCONV_ERR_ARG, "create_multi: invalid or repeated device index%s", "");
            return NULL;
        }
    }
    if (nDev > nOut) nDev = nOut;                              /* every device needs at least one channel */
    safconv_multi* m = (safconv_multi*)calloc(1, sizeof *m);
    if (!m) { sch_set_tl_error(SAFCONV_ERR_NOMEM, "out of host memory%s", ""); return NULL; }
    m->magic = SAFCONV_MAGIC_MULTI;
    m->kind = kind; m->nDev = nDev; m->hop = hop; m->len = len; m->nIn = nIn; m->nOut = nOut;
    m->transport = sch_env_int("SAFCONV_MULTI_TRANSPORT", SCM_T_HOST, 0, 1);
    m->spinUs = sch_env_int("SAFCONV_MULTI_SPIN_US", 200, 0, 1000000);
    m->detectPinned = 1;
    m->inBytes  = sizeof(float) * (size_t)nIn * hop;
    m->outBytes = sizeof(float) * (size_t)nOut * hop;
    m->H = H;
    const int base = nOut / nDev, extra = nOut % nDev;          /* contiguous, balanced: the first `extra` devices get one more */
    int failed = 0;
    for (int i = 0; i < nDev; i++) {
        scm_worker* w = &m->w[i];
        w->m = m; w->idx = i; w->device = devices[i];
        w->begin = i * base + (i < extra ? i : extra);
        w->count = base + (i < extra ? 1 : 0);
        pthread_mutex_init(&w->mu, NULL);
        pthread_cond_init(&w->cv, NULL);
        if (pthread_create(&w->th, NULL, worker_main, w) != 0) { failed = 1; pthread_mutex_destroy(&w->mu); pthread_cond_destroy(&w->cv); break; }
        w->started = 1;
    }
    void* mp = m;
    if (failed) { m_fail(m, SAFCONV_ERR_NOMEM, "create_multi: cannot start worker threads", NULL); scm_destroy(&mp); return NULL; }
    /* every device uploads and transforms its own rows of H, all at once (one PCIe link each) */
    int rc = run_op(m, SCM_OP_CREATE);
    m->H = NULL;
    if (!rc) {
        scdev_set_device(devices[0]);
        int e = scdev_host_alloc((void**)&m->h_in, m->inBytes);
        if (!e) e = scdev_host_alloc((void**)&m->h_out, m->outBytes);
        if (e) rc = m_fail(m, SAFCONV_ERR_NOMEM, "create_multi: pinned staging", scdev_error_string(e));
    }
    if (rc) {
        char keep[256]; snprintf(keep, sizeof keep, "%s", m->errmsg);
        scm_destroy(&mp);
        sch_set_tl_error(rc, "%s", keep);
        return NULL;
    }
    if (m->transport == SCM_T_NCCL && scm_set_option(m, "transport", SCM_T_NCCL)) {
        char keep[256]; snprintf(keep, sizeof keep, "%s", m->errmsg);
        const int code = m->err;
        scm_destroy(&mp);
        sch_set_tl_error(code, "%s", keep);
        return NULL;
    }
    return m;
}

void safconv_matrixConv_create_multi(void** const phMC, int hopSize, const float* H, int length_h, int nCHin, int nCHout,
                                     const int* devices, int nDevices)
{
    if (!phMC) return;
    *phMC = multi_create(SC_KIND_MATRIX, hopSize, H, length_h, nCHin, nCHout, devices, nDevices);
}

void safconv_multiConv_create_multi(void** const phMC, int hopSize, const float* H, int length_h, int nCH,
                                    const int* devices, int nDevices)
{
    if (!phMC) return;
    *phMC = multi_create(SC_KIND_MULTI, hopSize, H, length_h, nCH, nCH, devices, nDevices);
}

/* SAFCONV_DEVICES="0,1,2,3" or "all": the plain saf_matrixConv_create / saf_multiConv_create build a multi-GPU handle.
 * Returns NULL (and leaves the thread error untouched) when the variable does not ask for more than one device. */
void* scm_create_from_env(int kind, int hop, const float* H, int len, int nIn, int nOut)
{
    const char* v = getenv("SAFCONV_DEVICES");
    if (!v || !*v) return NULL;
    int dev[SCM_MAX_DEV], n = 0, avail = 0;
    if (scdev_device_count(&avail) != 0 || avail < 2) return NULL;
    if (!strcmp(v, "all")) {
        for (int i = 0; i < avail && i < SCM_MAX_DEV; i++) dev[n++] = i;
    } else {
        const char* p = v;
        while (*p && n < SCM_MAX_DEV) {
            char* end = NULL;
            const long d = strtol(p, &end, 10);
            if (end == p) break;
            dev[n++] = (int)d;
            p = (*end == ',') ? end + 1 : end;
        }
    }
    if (n < 2) return NULL;
    return multi_create(kind, hop, H, len, nIn, nOut, dev, n);
}

/* ------------------------------------------------------------------------------------------ */
/*  apply                                                                                       */
/* ------------------------------------------------------------------------------------------ */
void scm_apply(void* p, int kind, const float* in, float* out)
{
    safconv_multi* m = as_multi(p);
    if (!m || m->kind != kind || !in || !out) return;
    m->err = SAFCONV_OK; m->errmsg[0] = 0;
    scdev_set_device(m->w[0].device);
    const int direct = m->detectPinned && scdev_is_pinned_host(in) && scdev_is_pinned_host(out)
                    && scdev_is_pinned_host((const char*)in + m->inBytes - 1) && scdev_is_pinned_host((const char*)out + m->outBytes - 1);
    if (!direct) memcpy(m->h_in, in, m->inBytes);
    m->src = direct ? in : m->h_in;
    m->dst = direct ? out : m->h_out;
    if (run_op(m, SCM_OP_APPLY)) return;
    if (!direct) memcpy(out, m->h_out, m->outBytes);
}

/* ------------------------------------------------------------------------------------------ */
/*  extension calls on a multi-GPU handle                                                       */
/* ------------------------------------------------------------------------------------------ */
int scm_last_error(void* p) { safconv_multi* m = as_multi(p); return m ? m->err : SAFCONV_ERR_ARG; }
const char* scm_last_error_string(void* p) { safconv_multi* m = as_multi(p); return m ? m->errmsg : ""; }

int scm_get_info(void* p, safconv_info* info)
{
    safconv_multi* m = as_multi(p);
    if (!m || !info) return SAFCONV_ERR_ARG;
    safconv_info s;
    memset(info, 0, sizeof *info);
    for (int i = 0; i < m->nDev; i++) {
        if (safconv_get_info(m->w[i].shard, &s)) return SAFCONV_ERR_ARG;
        if (i == 0) *info = s;
        else {
            info->bytesFilters += s.bytesFilters; info->bytesDelayLine += s.bytesDelayLine;
            info->algBytesPerBlock += s.algBytesPerBlock; info->macAlgBytesPerBlock += s.macAlgBytesPerBlock;
        }
    }
    info->nCHin = m->nIn; info->nCHout = m->nOut; info->nOutLocal = m->nOut; info->outBegin = 0;
    return SAFCONV_OK;
}

int scm_set_option(void* p, const char* name, int value)
{
    safconv_multi* m = as_multi(p);
    if (!m || !name) return SAFCONV_ERR_ARG;
    m->err = SAFCONV_OK; m->errmsg[0] = 0;
    if (!strcmp(name, "worker_spin_us")) { m->spinUs = value < 0 ? 0 : value; return SAFCONV_OK; }
    if (!strcmp(name, "detect_pinned")) { m->detectPinned = value ? 1 : 0; return SAFCONV_OK; }
    if (!strcmp(name, "transport")) {
        if (value != SCM_T_HOST && value != SCM_T_NCCL) return SAFCONV_ERR_ARG;
        if (value == SCM_T_NCCL && !m->ncclReady) {
            if (nccl_load(m)) return m->err;
            scm_ncclComm_t comms[SCM_MAX_DEV];
            int devs[SCM_MAX_DEV];
            for (int i = 0; i < m->nDev; i++) devs[i] = m->w[i].device;
            const int r = m->nccl.CommInitAll(comms, m->nDev, devs);
            if (r) return m_fail(m, SAFCONV_ERR_CUDA, "ncclCommInitAll", m->nccl.GetErrorString(r));
            for (int i = 0; i < m->nDev; i++) m->w[i].comm = comms[i];
            if (run_op(m, SCM_OP_NCCL_SETUP)) return m->err;
            m->ncclReady = 1;
        }
        /* the shards' pre-computed tails stay valid: both transports feed the same per-device look-ahead apply */
        m->transport = value;
        return SAFCONV_OK;
    }
    m->optName = name; m->optValue = value;                    /* everything else: every shard */
    return run_op(m, SCM_OP_OPTION);
}

int scm_reset_state(void* p)
{
    safconv_multi* m = as_multi(p);
    if (!m) return SAFCONV_ERR_ARG;
    m->err = SAFCONV_OK; m->errmsg[0] = 0;
    return run_op(m, SCM_OP_RESET);
}

int scm_synchronize(void* p)
{
    safconv_multi* m = as_multi(p);
    if (!m) return SAFCONV_ERR_ARG;
    m->err = SAFCONV_OK; m->errmsg[0] = 0;
    return run_op(m, SCM_OP_SYNC);
}

int safconv_multi_get_devices(void* p, int* devices, int cap)
{
    safconv_multi* m = as_multi(p);
    if (!m) return 0;
    for (int i = 0; i < m->nDev && devices && i < cap; i++) devices[i] = m->w[i].device;
    return m->nDev;
}

void* safconv_multi_get_shard(void* p, int i)
{
    safconv_multi* m = as_multi(p);
    return (m && i >= 0 && i < m->nDev) ? (void*)m->w[i].shard : NULL;
}
