/*
 * safconv_dev.h -- internal, thin C ABI between the C host layer (safconv_host.c)
 * and the CUDA layer (safconv_kernels.cu).  Plain ints, sizes and pointers only.
 *
 * Terminology:
 *   N      FFT size (power of two >= max(64, 2*hop)); M = N/2 = complex FFT length
 *          = number of PACKED bins: bin 0 carries (DC, Nyquist) in (re, im) - both are
 *          purely real for real signals (reference: kiss_fftr.c:99-104, 137-138).
 *   kt     bin tile of SC_BK = 32 packed bins (256 bytes of float2)
 *   p      filter partition (0 = newest), P = ceil(length_h / hop)
 *   slot   ring position of a block spectrum in the frequency-domain delay line (FDL)
 *   ot     output tile (<= 64 output channels), OTsz outputs per tile
 *   unit   one (ot, kt, p): nIn x OTsz rows of 32 bins of H, contiguous in memory
 *   stage  SNI input channels of a unit = one TMA transaction of the MAC pipeline
 *   group  one (ot, kt): the P*SPU stages whose sum is one tile of the output spectrum
 */
#ifndef SAFCONV_DEV_H_INCLUDED
#define SAFCONV_DEV_H_INCLUDED

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_BK            32      /* packed bins per tile                                   */
#define SC_MAX_OT        64      /* output channels per tile                               */
#define SC_MAC_CWARPS    8       /* consumer warps of the MAC kernel (+1 producer warp)    */
#define SC_MAC_NSTAGES   4       /* default TMA pipeline depth                             */
#define SC_MAC_MAX_STAGES 16
#define SC_STAGE_H_BYTES 32768   /* default target bytes of H per stage                    */
#define SC_MAX_SNI       32      /* max input channels per stage                           */
#define SC_MAX_M         8192    /* max complex FFT length (hop <= 8192)                   */

enum { SC_KIND_MATRIX = 0, SC_KIND_MULTI = 1, SC_KIND_TV = 2 };

typedef struct scdev_plan {
    int kind;
    int hop, len, nIn, nOutLocal;
    int N, M, logM, P;
    int fftThreads;
    /* matrix MAC geometry */
    int nKT, nOT, OTsz;
    int SNI, SPU;            /* inputs per stage, stages per unit            */
    int R, WGo, WGk;         /* outputs per thread, warp groups over outputs / over rows */
    long long totalStages;
    int macGrid;
    int nGroups, nSlots;
    int RS;                  /* delay-line ring slots = P + maxBatch            */
    int maxBatch;            /* blocks per batched launch group (>= 1)          */
    int macHints;            /* L2 eviction-priority hints: 0 none, 1 filters evict_first (streamed), 2 filters evict_last (fit in L2) */
    int macSmemBytes;
    int macStages;           /* TMA pipeline depth                           */
    int macStageBytes;       /* target bytes of H per stage                  */
    /* time-varying convolver */
    int nIRs;
} scdev_plan;

/* one pass of the MAC over partitions [pLo, pLo + nP) of every group, with its own split-K bookkeeping and
 * partial-tile buffer.  The full pass (0, P) lives in scdev_plan / scdev_bufs; the host layer adds a TAIL pass
 * (1, P-1) -- everything that does not depend on the newest input block, run AHEAD of the next apply -- and a
 * HEAD pass (0, 1) for the newest block only. */
typedef struct scdev_macpass {
    int pLo, nP;
    long long totalStages;   /* nGroups * nP * SPU */
    int grid, nSlots;
    int*  ctaBase;           /* device int[grid+1]      */
    int*  grpStart;          /* device int[nGroups+1]   */
    void* Zp;                /* device float2[nSlots][OTsz][32] */
    void* ZpB;               /* second buffer (tail pass: block t+1's tail is written while K3 of block t still reads) or NULL */
    int   stages;            /* > 0: TMA pipeline depth of this launch (shallower than the plan's -> less shared memory) */
} scdev_macpass;

typedef struct scdev_bufs {
    /* all device pointers */
    void*  tw;        /* float2[M]        W_N^j                                            */
    void*  H;         /* float2: matrix [nOT][nKT][P][nIn][OTsz][32]; multi [nCH][P][M];    */
                      /*         tv     [nIRs][nOut][P][M]                                  */
    void*  X;         /* float2: matrix [nKT][RS][nIn][32];           multi [nCH][RS][M];   */
                      /*         tv     [P][M]                                              */
    void*  Zp;        /* float2[maxBatch][nSlots][OTsz][32]  split-K partial output spectra */
    float* zt;        /* float[maxBatch][nOutLocal][2*hop]   batched inverse transforms     */
    float* tail;      /* float[nOutLocal][hop]     overlap-add tails                        */
    float* tail2;     /* tv only: y_n_overlap_last                                          */
    unsigned int* counters; /* [0] block counter, [1] last-CTA ticket                       */
    int*   ctaBase;   /* int[macGrid]   first partial slot of each MAC CTA                  */
    int*   grpStart;  /* int[nGroups+1] partial slots of group g: grpStart[g]..grpStart[g+1]-1 */
    void*  wtab;      /* float2[2][M]  twiddle tables of the warp-level register FFT (multiConv batched path) or NULL */
} scdev_bufs;

/* workspace of the offline (batched frames, tensor-core) path, see safconv_offline.cu */
typedef struct scdev_offline {
    void  *XGhi, *XGlo;      /* A operand  [bin][kg][rows][16 B]  (frames x (input, re/im)), hi / lo parts; a k-group = 4 tf32 or 8 fp16 */
    void  *HGhi, *HGlo;      /* B operand  [bin][p][kg][Nn][16 B] ((output, re/im) x (input, re/im))            */
    float *Ys;               /* output spectra [bin][Tpad][Nn]                                                  */
    float *scal;             /* fp16 operands: [0] bound on the input spectra of the current render, [1] bound on the filter spectra;
                                [2] (as int) the GEMM's tile ticket, zeroed before every launch */
    int smCount;             /* CTAs of the persistent GEMM */
    int nTiles;              /* output tiles of Nn/2 (<= 64) outputs: grid.z of the GEMM                          */
    int f16, ipc;            /* 1: fp16 operands (default), 0: tf32; inputs per k-group (4 / 2)                 */
    int wfft;                /* 1: warp-register FFT transform kernels (fp16 operands, M <= 1024)                */
    void *wtab;              /* their two [M/32][32] twiddle tables                                              */
    int capFrames, capTpad, capRows, packed;
    int Nn, Kp, nKG, nKC, rowsX, tmemCols, gemmSmem, flush, fpc, opc, fftThreads;
} scdev_offline;

/* --- device / memory / stream plumbing (all return 0 on success, else a cudaError_t value) --- */
int  scdev_device_count(int* n);
int  scdev_set_device(int dev);
int  scdev_get_device(int* dev);
int  scdev_device_props(int dev, int* smCount, int* maxSmemOptin, int* ccMajor, int* ccMinor);
int  scdev_malloc(void** p, size_t bytes);
int  scdev_free(void* p);
int  scdev_host_alloc(void** p, size_t bytes);
int  scdev_host_free(void* p);
int  scdev_memset_async(void* p, int v, size_t bytes, void* stream);
int  scdev_memcpy_h2d_async(void* d, const void* h, size_t bytes, void* stream);
int  scdev_memcpy_d2h_async(void* h, const void* d, size_t bytes, void* stream);
int  scdev_memcpy_h2d_sync(void* d, const void* h, size_t bytes, void* stream);
int  scdev_memcpy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t widthBytes, size_t height,
                          int toHost, void* stream);
int  scdev_stream_create(void** s);
int  scdev_stream_create_high_priority(void** s);
int  scdev_stream_destroy(void* s);
int  scdev_stream_sync(void* s);
int  scdev_event_create(void** e);
int  scdev_event_destroy(void* e);
int  scdev_event_create_sync(void** e);          /* no timing: cheapest to record / wait on */
int  scdev_stream_wait_event(void* stream, void* e);
int  scdev_event_done(void* e);                 /* 1 complete / never recorded, 0 pending, < 0 error */
int  scdev_event_record(void* e, void* stream);
int  scdev_event_sync(void* e);
int  scdev_last_error_clear(void);             /* returns and clears the runtime's last (non-sticky) error */
int  scdev_event_elapsed_ms(void* e0, void* e1, float* ms);
int  scdev_graph_begin(void* stream);
int  scdev_graph_end(void* stream, void** graphExec);
int  scdev_graph_launch(void* graphExec, void* stream);
int  scdev_graph_destroy(void* graphExec);
const char* scdev_error_string(int err);

/* --- kernels --- */
/* one-time opt-in of >48 KB dynamic shared memory etc. for this plan */
int  scdev_prepare(const scdev_plan* pl);
/* K0: partition + forward real FFT of the time-domain filters d_h (matrix: [nOutLocal][nIn][len],
 * multi: [nCH][len], tv: [nIRs][nOut][len]) into bufs->H */
int  scdev_filter_transform(const scdev_plan* pl, const scdev_bufs* b, const float* d_h, void* stream);
/* K1: forward real FFT of nBlocks new input blocks d_in [nBlocks][nIn][hop] into ring slots (counter + b) % RS */
int  scdev_input_fft(const scdev_plan* pl, const scdev_bufs* b, const float* d_in, int nBlocks, void* stream);
/* K2: filter-streaming complex multiply-accumulate over partitions x inputs for blocks blk .. blk+nBlocks-1 of
 * the batch, streamed back to back inside ONE launch (every block streams the filter spectra once) */
int  scdev_mac(const scdev_plan* pl, const scdev_bufs* b, int blk, int nBlocks, void* stream);
/* K2 restricted to one pass (see scdev_macpass) */
/* zpSel: 0 = ps->Zp, 1 = ps->ZpB; count >= 0: block counter supplied by the host instead of read from the device */
int  scdev_mac_pass(const scdev_plan* pl, const scdev_bufs* b, const scdev_macpass* ps, int blk, int nBlocks,
                    int zpSel, long long count, void* stream);
/* K3: sum split-K partials, inverse real FFT, 1/N, overlap-add, tail save, block counter++ (one block) */
int  scdev_ifft_ola(const scdev_plan* pl, const scdev_bufs* b, float* d_out, void* stream);
/* K3 summing the partial tiles of pass p1 then pass p2 (NULL, NULL = the full pass) */
int  scdev_ifft_ola_passes(const scdev_plan* pl, const scdev_bufs* b, const scdev_macpass* p1, int zpSel1, const scdev_macpass* p2,
                           float* d_out, void* stream);
/* K3 for a batch: inverse FFTs of all nBlocks blocks in one launch, then the overlap-add chain; counter += nBlocks */
int  scdev_ifft_ola_batch(const scdev_plan* pl, const scdev_bufs* b, float* d_out, int nBlocks, void* stream);
/* offline path: allocate / grow the workspace for T frames (and build the filter operand on first use),
 * render d_in [nIn][T*hop] -> d_out [nOutLocal][(T-skip)*hop] from a zero state (the first `skip` frames are
 * history only), free.  `events` (4 CUDA events or
 * NULL) are recorded before the forward FFTs, before the GEMM, after the GEMM, and at the end. */
int  scdev_offline_prepare(const scdev_plan* pl, const scdev_bufs* b, scdev_offline* o, int T, void* stream);
int  scdev_offline_run(const scdev_plan* pl, const scdev_bufs* b, scdev_offline* o,
                       const float* d_in, float* d_out, int T, int skip, void** events, void* stream);
int  scdev_offline_free(scdev_offline* o);
/* general-size real FFT (any even N): M = N/2 point mixed-radix Stockham FFT + real split (safconv_gfft.cu) */
#define SC_GFFT_MAX_FACTORS 32
#define SC_GFFT_SMEM_M      8192     /* up to this M a transform runs inside one CTA (two M-point arrays in shared memory) */
typedef struct scdev_gfft_plan {
    int N, M;
    int nf, fac[SC_GFFT_MAX_FACTORS];   /* radices in pass order, product = M (M = 1: the single "radix" 1)          */
    void* tw;                           /* device float2[M]      W_M^e                                             */
    void* stw;                          /* device float2[M/2+1]  W_N^k                                             */
    void *w0, *w1;                      /* M > SC_GFFT_SMEM_M: device float2[maxBatch][M] ping-pong work arrays     */
    int maxBatch;
} scdev_gfft_plan;
int  scdev_gfft_smem_ok(int M);
/* dir 0: in [nBatch][N] real -> out [nBatch][N/2+1] complex (unscaled); dir 1: the inverse, scaled by 1/N, imaginary
 * parts of DC and Nyquist ignored.  Device pointers (one-CTA path: page-locked host memory works too). */
int  scdev_gfft_run(const scdev_gfft_plan* p, int dir, int nBatch, const float* in, float* out, void* stream);
/* true non-partitioned convolvers: zero-pad rows, per-bin products (matrix: summed over inputs), shifting overlap-add */
int  scdev_np_pad(const float* in, float* xpad, int rows, int len, int F, size_t inStride, void* stream);
int  scdev_np_mac(const void* H, const void* X, void* Z, int nOut, int nIn, int nBins, int multi, void* stream);
int  scdev_np_ola(const float* z, const float* ovOld, float* ovNew, float* out, int nOut, int hop, int F, void* stream);
/* batch of stand-alone real FFTs (saf_rfft conventions) on device buffers; dir 0 forward, 1 backward */
int  scdev_rfft(int N, int logM, int nBatch, int dir, const float* d_in, float* d_out, const void* d_tw, void* stream);
/* build b->wtab (no-op outside 64 <= M <= 1024) */
int  scdev_wfft_tables(const scdev_plan* pl, scdev_bufs* b, void* stream);
/* 1 if p is page-locked host memory known to CUDA (cudaHostAlloc / cudaHostRegister), else 0 */
int  scdev_is_pinned_host(const void* p);
/* small matrix problems: K1+K2+K3 in one launch, one CTA per output channel; in/out may be mapped host memory */
int  scdev_small_fits(const scdev_plan* pl, int maxSmemOptin);
int  scdev_small_resident_start(const scdev_plan* pl, const scdev_bufs* b, void* mailbox, unsigned int lastSeq, unsigned int idleUs,
                                void* stream);
int  scdev_small_resident_stamps(unsigned long long out[8]);
int  scdev_small_fused(const scdev_plan* pl, const scdev_bufs* b, const float* in, float* out, void* stream,
                       volatile unsigned int* done, unsigned int seq, int* signalled);
/* multiConv, nBlocks device-resident blocks d_in [nBlocks][nCH][hop] -> d_out [nBlocks][nCH][hop]:
 * which = 0 forward FFTs of all blocks, 1 MAC + inverse FFTs of all blocks, 2 overlap-add chain (+ counter += nBlocks) */
int  scdev_multi_batch(const scdev_plan* pl, const scdev_bufs* b, const float* d_in, float* d_out, int nBlocks, int which, void* stream);
/* multiConv: K1+K2+K3 fused, one CTA per channel */
int  scdev_multi_fused(const scdev_plan* pl, const scdev_bufs* b, const float* d_in, float* d_out, void* stream);
/* TVConv: 1-input FFT + (1..3) IR MACs + cross-fade, one CTA per output channel */
int  scdev_tv_fused(const scdev_plan* pl, const scdev_bufs* b, const float* d_in, float* d_out,
                    int irIdx, int irLast, int irLast2, void* stream);


/* --- filter producers (safconv_producers.cu): binaural Ambisonic decoder design, shoebox image-source RIRs --- */
/* Y [nSH][nD] = getRSH of the directions d_dirs [nD][2] (azimuth, elevation in degrees) */
int  scdev_prod_rsh(int order, const float* d_dirs, int nD, float* d_Y, void* stream);
/* G [nSH][nD] = (Y W Y^T)^-1 Y W; d_aug: double [nSH][2 nSH] scratch, d_flag: 1 if the Gram matrix is singular */
int  scdev_prod_lsmatrix(const float* d_Y, const float* d_w, int nD, int n, double* d_aug, float* d_G, int* d_flag, void* stream);
/* SPR decoder set-up: condition numbers of the SH transform per order 0 .. nhMax (d_Y [(nhMax+1)^2][nD]; d_aug double
 * [nS][2 nS], nS = (nhMax+1)^2), and G [n][nD] = the t-design projection folded into one matrix (d_M: double [nA][n]) */
int  scdev_prod_spr_cond(const float* d_Y, const float* d_w, int nD, int nhMax, double* d_aug, float* d_cond, void* stream);
int  scdev_prod_spr_matrix(const float* d_Ynh, const float* d_Ytd, const float* d_w, int nD, int K, int nA, int n,
                           double* d_M, float* d_G, void* stream);
/* D [nB][2][nSH] complex = H [nB][2][nD] complex times G^T; ta: bands >= bc use the HRTFs of band bc */
int  scdev_prod_ls(const void* d_H, const float* d_G, int nB, int nD, int n, int ta, int bc, void* d_D, void* stream);
int  scdev_prod_diffeq(const void* d_H, const float* d_Y, const float* d_w, int nB, int nD, int n, void* d_D, void* stream);
int  scdev_prod_magls(const void* d_H, const float* d_Y, const float* d_G, int nB, int nD, int n, int bc, void* d_D, void* d_hm, void* stream);
int  scdev_prod_scale(void* d_D, const float* d_a, int n, size_t total, void* stream);
int  scdev_prod_diffcov(const void* d_H, const float* d_Y, const float* d_w, int nB, int nD, int n, void* d_D, void* stream);
/* D [nB][rows] -> Dt [rows][nB] (complex) */
int  scdev_prod_pack(const void* d_D, int nB, int rows, void* d_Dt, void* stream);
/* image sources: d_pairs = ScpImsPair[nPairs] (safconv_prod_core.cuh); count pass -> d_stats [nPairs][2] = (images, bits of the
 * largest distance); render pass -> fp64 taps d_acc; finish: fp64 -> fp32; bank: RIRs of one receiver -> [nCh][nSrc][L] */
int  scdev_ims_count(const void* d_pairs, int nPairs, long long maxLengthVec, unsigned int* d_stats, int smCount, void* stream);
int  scdev_ims_render(const void* d_pairs, int nPairs, long long maxLengthVec, int maxOrder, const float* d_absTab, int nBands, int maxW,
                      const float* d_norms, double* d_acc, size_t totalTaps, int smCount, void* stream);
int  scdev_ims_finish(const double* d_acc, float* d_rir, size_t total, void* stream);
/* render pass, windowed version: one CTA per (pair, window of pairs[].tw taps), taps in shared memory, fp32 RIRs written
 * directly to d_rirPtrs[pair]; maxWindows = max ceil(len / tw), accDoubles = max nSH * tw over the pairs */
int  scdev_ims_render_windows(const void* d_pairs, int nPairs, int maxWindows, int accDoubles, int maxOrder, const float* d_absTab,
                              int nBands, int maxW, const float* d_norms, float* const* d_rirPtrs, void* stream);
int  scdev_ims_bank(const float* const* d_rirPtrs, const int* d_len, int nSrc, int nCh, int L, float* d_H, void* stream);
int  scdev_memcpy_d2d_async(void* dst, const void* src, size_t bytes, void* stream);
int  scdev_mem_free_bytes(size_t* freeBytes);

#ifdef __cplusplus
}
#endif
#endif
