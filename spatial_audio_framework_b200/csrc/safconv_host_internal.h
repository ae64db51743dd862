/*
 * safconv_host_internal.h -- private to the C host layer of libsafconv_b200.so (safconv_host.c: single-device handles,
 * safconv_multi.c: one handle spanning several GPUs of a box).  Not installed, not part of the C ABI.
 */
#ifndef SAFCONV_HOST_INTERNAL_H_INCLUDED
#define SAFCONV_HOST_INTERNAL_H_INCLUDED

#include "../../include/safconv_b200.h"
#include "safconv_dev.h"

#include <stdint.h>

#define SC_HIDDEN __attribute__((visibility("hidden")))

#define SAFCONV_MAGIC 0x5AFC0B20u
#define SAFCONV_MAGIC_MULTI 0x5AFC0B28u     /* safconv_multi.c: a handle that spans several devices */
#define SAFCONV_VERSION_STRING "safconv-b200 0.3 (sm_100a; matrixConv/multiConv/TVConv; multi-GPU handles; filter producers)"

/* mailbox of the resident latency kernel (page-locked host memory; the device side is ScMailbox in safconv_kernels.cu) */
#define SC_RES_EXIT 0xFFFFFFFFu
typedef struct sc_mailbox {
    volatile unsigned int done;          /* completion word: sequence number of the last finished block */
    unsigned int pad0[15];
    volatile unsigned long long bell;    /* doorbell, ONE 8-byte word: low half = sequence number, high half = (generation << 8) | slot */
    volatile unsigned int alive;         /* 1 while a resident kernel is polling */
    unsigned int pad1[13];
    volatile unsigned long long buf[16]; /* table of (in, out) buffer addresses, 8 slots */
} sc_mailbox;

typedef struct safconv_handle {
    uint32_t   magic;
    int        err;
    char       errmsg[256];
    scdev_plan pl;
    scdev_bufs b;
    int        device, smCount, maxSmem;
    int        nCHoutTotal, outBegin;
    void*      streamOwn;
    void*      stream;
    float     *d_in, *d_out;         /* device staging for the host-pointer API  */
    float     *h_in, *h_out;         /* pinned host staging                      */
    size_t     inBytes, outBytes;
    size_t     bytesH, bytesX, bytesZp;
    int        useGraph;
    int        smallOk;              /* the plan qualifies for the fused small-problem kernel */
    int        smallFused;           /* 1: saf_matrixConv_apply of a small problem = ONE fused kernel on mapped host buffers */
    int        batching;             /* 1: safconv_apply_device_blocks shares the FFT launches across a batch */
    int        detectPinned;         /* 1: DMA straight from/to caller buffers that are already page-locked */
    void*      graphExec;
    int        timingCap, timingCount;   /* 4 CUDA events per enqueued launch group while kernel timing is enabled */
    void**     evRing;
    int*       evBlocks;             /* blocks covered by each launch group */
    void      *graphIn, *graphOut;   /* host pointers captured in graphExec */
    scdev_offline off;               /* offline (batched frames) workspace, allocated on first use */
    void*      offEv[4];
    void      *offIn, *offOut;       /* copy streams of the pipelined host-buffer render */
    void*      offPipeEv[6];         /* per double-buffer slot: input landed, segment rendered, output copied back */
    float     *offStageIn[2], *offStageOut[2];   /* device staging of the pipelined render (kept between calls) */
    int        tvLast, tvLast2;      /* posIdx_last, posIdx_last2 (reference .c:438, 618-619) */
    /* look-ahead (matrix, P >= 2): all partitions p >= 1 of block t+1 only need spectra that are already in the
     * delay line when block t is done, so that TAIL pass is enqueued right behind block t and runs while the host
     * is away; saf_matrixConv_apply(t+1) then only pays for the newest partition (HEAD pass) */
    scdev_macpass tailPass, headPass;
    int        lookahead;            /* option: 1 = use the tail/head split in the host-pointer apply */
    int        tailReady;            /* a tail pass for the current block counter is enqueued on `stream` */
    void*      evDone;               /* recorded after the output of a block is complete: apply waits on it, not on the whole stream */
    void*      streamIn;             /* side stream: the forward FFT of the new block runs beside the tail pass */
    void       *evIn, *evFence;      /* streamIn -> stream, stream -> streamIn */
    void*      streamOut;            /* high-priority side stream: K3 (+ D2H) of block t runs beside the tail pass of block t+1 */
    void*      evMac;                /* stream -> streamOut: the partial tiles of the current block are complete */
    void*      evTail;               /* (unused since the two-deep look-ahead) */
    void*      evTailB[2];           /* end of the tail pass of block b, indexed by b & 1 */
    int        tlCap, tlN;           /* SAFCONV_TIMELINE=n: device timeline of the first n look-ahead calls (8 events each), printed once */
    void**     tlEv;                 /* [tlCap][8]: K1 start/end, head start/end, K3 start/end, tail pass start/end */
    double*    tlHost;               /* [tlCap][2]: host time at entry / return (ns) */
    char*      tlReg;
    int        oneStreamLatency;     /* latency regime: K1, head pass and K3 on one stream (no event hops between streams) */
    int        trRegime;             /* trace: regime of the current call */
    int        laDepth;              /* look-ahead depth D: tail passes cover partitions p >= D and are queued D blocks ahead */
    unsigned int tailUpTo;           /* with tailReady: tail passes of blocks count .. tailUpTo are enqueued */
    unsigned int count;              /* host mirror of the device block counter (counters[0]) */
    int        headInK3;             /* latency regime: the newest partition is added inside K3 (no head-pass launch) */
    int        trace;                /* SAFCONV_TRACE=1: per-call device timeline of the look-ahead apply on stderr (debugging) */
    void*      trEv[6];              /* head start, head end, previous tail end, K3 start, K3 end, tail end */
    volatile unsigned int* doneWord; /* page-locked word the cluster latency kernel writes its sequence number into when the block's output is complete
                                        (= &mailbox->done: the first word of the resident kernel's mailbox) */
    struct sc_mailbox* mailbox;      /* page-locked: completion word, doorbell, alive flag, buffer addresses (layout = ScMailbox in safconv_kernels.cu) */
    int        residentUs;           /* option "resident_us" / SAFCONV_RESIDENT_US: > 0 keeps the cluster latency kernel resident; it leaves after this idle time */
    int        resActive;            /* a resident kernel has been started on `stream` */
    unsigned int resGen, resNext;    /* generation of the mailbox's buffer table, next slot to replace */
    unsigned int doneSeq;
    int        flagWait;             /* option "flag_wait" / SAFCONV_FLAG_WAIT: poll doneWord instead of synchronising the stream */
    int        hostTrace;            /* SAFCONV_HOSTTRACE=1: host-side time of the zero-copy apply by segment, printed at destroy */
    double     htAcc[3];             /* ns: argument / pinned checks, launch, wait for completion */
    unsigned   htN;
} safconv_handle;


/* where one block of the look-ahead apply comes from and goes to (safconv_host.c: apply_lookahead) */
typedef struct sch_la_io {
    const float* k1src;   /* what K1 reads: a page-locked host block (zero-copy) or a device buffer                     */
    const float* h2dSrc;  /* != NULL: copy-engine upload of this host block into h->d_in first (K1 then reads h->d_in)  */
    void*        evSrc;   /* != NULL: K1 waits for this event (the input was produced on a foreign stream)              */
    float*       kout;    /* what K3 writes: a page-locked host block (zero-copy) or a device buffer                    */
    float*       d2hDst;  /* != NULL: copy-engine download of h->d_out into this host block behind K3                   */
    int          sync;    /* 1: wait until the output is complete (evDone) before returning                             */
} sch_la_io;

/* internals shared with safconv_multi.c */
SC_HIDDEN safconv_handle* sch_as_handle(void* p);
SC_HIDDEN void sch_set_tl_error(int code, const char* fmt, const char* detail);
SC_HIDDEN int  sch_fail(safconv_handle* h, int code, const char* what, int cudaErr);
SC_HIDDEN int  sch_env_int(const char* name, int dflt, int lo, int hi);
SC_HIDDEN int  sch_thread_device(void);      /* device chosen with safconv_set_device on this thread, or -1 */
SC_HIDDEN void sch_set_thread_device(int device);
SC_HIDDEN safconv_handle* sch_conv_create(int kind, int hop, const float* const* chunks, int nChunks, size_t rowsPerChunk,
                                          int len, int nIn, int nOutLocal, int nOutTotal, int outBegin, int nIRs);
SC_HIDDEN void sch_handle_free(safconv_handle* h);
/* one block, in / out page-locked host memory visible to the handle's device (synchronous) */
SC_HIDDEN void sch_apply_pinned(safconv_handle* h, const float* src, float* dst, int irIdx);
/* one block of a matrix handle with look-ahead through an explicit io description; returns a CUDA error code */
SC_HIDDEN int  sch_apply_lookahead_io(safconv_handle* h, const sch_la_io* io);
SC_HIDDEN int  sch_uses_lookahead(const safconv_handle* h);

/* safconv_rfft.c: plans of the general-size real FFT */
SC_HIDDEN int  scr_plan_init(scdev_gfft_plan* pl, int N, void* stream);
SC_HIDDEN int  scr_plan_reserve(scdev_gfft_plan* pl, int nBatch);
SC_HIDDEN void scr_plan_free(scdev_gfft_plan* pl);
int safconv_debug_fft_factors(int M, int* fac, int cap);

/* safconv_np.c: true non-partitioned convolvers (one big FFT per block) */
SC_HIDDEN int   scn_is_np(const void* p);
SC_HIDDEN int   scn_enabled(void);
SC_HIDDEN void* scn_create(int kind, int hop, const float* H, int len, int nIn, int nOut);
SC_HIDDEN void  scn_apply(void* p, int kind, const float* in, float* out);
SC_HIDDEN void  scn_destroy(void** pp);
SC_HIDDEN int   scn_last_error(void* p);
SC_HIDDEN const char* scn_last_error_string(void* p);
SC_HIDDEN int   scn_get_info(void* p, safconv_info* info);
SC_HIDDEN int   scn_reset_state(void* p);

/* safconv_multi.c */
SC_HIDDEN int   scm_is_multi(const void* p);
SC_HIDDEN void  scm_apply(void* p, int kind, const float* in, float* out);
SC_HIDDEN void  scm_destroy(void** pp);
SC_HIDDEN void* scm_create_from_env(int kind, int hop, const float* H, int len, int nIn, int nOut);
SC_HIDDEN int   scm_last_error(void* p);
SC_HIDDEN const char* scm_last_error_string(void* p);
SC_HIDDEN int   scm_get_info(void* p, safconv_info* info);
SC_HIDDEN int   scm_set_option(void* p, const char* name, int value);
SC_HIDDEN int   scm_reset_state(void* p);
SC_HIDDEN int   scm_synchronize(void* p);

#endif
