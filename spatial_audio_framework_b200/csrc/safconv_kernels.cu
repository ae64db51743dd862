/*
 * safconv_kernels.cu -- sm_100a kernels + thin C-ABI CUDA layer of libsafconv_b200.so
 *
 * Hot path of saf_matrixConv_apply / saf_multiConv_apply / saf_TVConv_apply
 * (reference: /root/reference/framework/modules/saf_utilities/saf_utility_matrixConv.c:209-235,
 * 388-413, 546-620), re-designed for B200:
 *
 *   K0  filter_fft_kernel   create-time: partition the FIRs, forward real FFT, store the
 *                           spectra in the streaming layout of the MAC          (.c:116-125)
 *   K1  input_fft_kernel    per block: forward real FFT of the nIn input blocks into the
 *                           newest slot of the frequency-domain delay line ring (.c:211-215)
 *   K2  mac_kernel          per block: Z[no][k] = sum_p sum_ni H[no][p][ni][k] * X[t-p][ni][k]
 *                           H streamed once from HBM with TMA bulk copies through a 4-stage
 *                           mbarrier pipeline; split over (bin tile, partition range) so that
 *                           every SM streams an equal, contiguous part of H   (.c:219, 225-227)
 *   K3  ifft_ola_kernel     per block: sum the split-K partial spectra, ONE inverse real FFT per
 *                           output channel (the reference does P*nIn of them, .c:220-222),
 *                           1/N, overlap-add, tail save                        (.c:230-233)
 *   multi_fused_kernel      multiConv: K1+K2+K3 in one launch, one CTA per channel
 *   multi_fft_w_kernel /    multiConv, batches of device-resident blocks on the warp-level register FFT
 *   multi_mac_ifft_w_kernel (safconv_wfft.cuh): register sliding window over the delay line per bin
 *   small_cluster_kernel<R> small matrix problems (M = 32 R = 64 .. 1024): K1+K2+K3 in one launch on ONE thread-block
 *                           cluster of up to 8 CTAs (one warp per FFT on the register FFT, per-bin sums spread over the
 *                           CTAs, distributed-shared-memory gather), mapped host buffers; and its RESIDENT version
 *                           small_cluster_resident_kernel<R>: one block per doorbell, no launch per call
 *   small_fused_kernel      the same for the remaining small shapes: one CTA per output
 *   tv_fused_kernel         TVConv: one CTA per output channel, up to three IR sets + cross-fade; hops above 4096 as
 *                           tv_input_kernel / tv_mac_ifft_kernel / tv_xfade_kernel
 *   rfft_forward/backward   the power-of-two FFT pair on its own (saf_rfft conventions), parity-test entry points
 *   (general-size FFT: safconv_gfft.cu; offline tensor-core path: safconv_offline.cu)
 *
 * K2 runs over a PASS = partitions [pLo, pLo+nP) of every group: the full pass, and for the look-ahead
 * apply of the host layer a tail pass (D, P-D) that is enqueued ahead of the next block(s) and a head pass
 * (0, D), D = 1 or 2; K3 can also add the newest partition(s) itself (headH).
 *
 * The real FFT of size N is an M = N/2 point complex FFT (safconv_fft.cuh: radix-4/2 decimation-in-
 * frequency passes in shared memory with the last five radix-2 stages as warp-shuffle butterflies, or
 * register radix-16 passes for batches) plus a split pass; spectra are kept PACKED (M complex values,
 * bin 0 = (DC, Nyquist)).
 *
 * No CPU fallback exists: every entry point returns a CUDA error code if the device path fails.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "safconv_dev.h"
#include "safconv_fft.cuh"
#include "safconv_wfft.cuh"

/* ------------------------------------------------------------------------------------------ */
/*  K0: filter partition + forward FFT                                                         */
/*  one CTA per (row, input, partition), row = output (matrix), channel (multi) or IR x output  */
/*  (TV); the CTA index is linear over grid.x (+ grid.y beyond 2^30 CTAs): no 65535 limits     */
/* ------------------------------------------------------------------------------------------ */
struct FilterArgs {
    const float* h;        /* time-domain filters */
    float2*      H;
    const float2* tw;
    int kind, hop, len, nIn, M, logM, P, nKT, OTsz;
    int nInGrid;           /* nIn (matrix) or 1 */
    long long total;       /* rows * nInGrid * P */
};

__global__ void filter_fft_kernel(FilterArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    float2* stw = sm + SC_ALEN(a.M);
    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    const long long cta = (long long)blockIdx.y * gridDim.x + blockIdx.x;
    if (cta >= a.total) return;                         /* block-uniform */
    const int p = (int)(cta % a.P);
    const long long rowIn = cta / a.P;
    const int ni = (int)(rowIn % a.nInGrid), no = (int)(rowIn / a.nInGrid);
    const float* src;
    if (a.kind == SC_KIND_MATRIX) src = a.h + ((size_t)no * a.nIn + ni) * a.len;
    else                          src = a.h + (size_t)no * a.len;        /* multi: [ch][len]; tv: [ir*nOut+no][len] */
    /* taps [p*hop, p*hop+hop) of this FIR, zero beyond length_h (reference .c:119-121) */
    const int t0 = p * a.hop;
    for (int n = threadIdx.x; n < a.M; n += blockDim.x) {
        const int i = t0 + 2 * n;
        float2 v;
        v.x = (2 * n < a.hop && i < a.len) ? __ldg(src + i) : 0.f;
        v.y = (2 * n + 1 < a.hop && i + 1 < a.len) ? __ldg(src + i + 1) : 0.f;
        sm[padi(n, a.logM)] = v;
    }
    __syncthreads();
    cfft_dif<false>(sm, a.M, a.logM, stw, wide);

    for (int k = threadIdx.x; k <= (a.M >> 1); k += blockDim.x) {
        float2 Xk, Xmk;
        int k2 = a.M - k;
        if (k == 0) {
            const float2 z = sm[0];
            Xk = make_float2(z.x + z.y, z.x - z.y);      /* packed (DC, Nyquist) */
            k2 = 0;
            Xmk = Xk;
        } else {
            fwd_split_pair(sm, k, a.M, a.logM, spl, Xk, Xmk);
        }
        if (a.kind == SC_KIND_MATRIX) {
            const int ot = no / a.OTsz, nl = no - ot * a.OTsz;
            /* [ot][kt][p][ni][nl][32] */
            size_t o1 = (((((size_t)ot * a.nKT + (k >> 5)) * a.P + p) * a.nIn + ni) * a.OTsz + nl) * SC_BK + (k & 31);
            size_t o2 = (((((size_t)ot * a.nKT + (k2 >> 5)) * a.P + p) * a.nIn + ni) * a.OTsz + nl) * SC_BK + (k2 & 31);
            a.H[o1] = Xk;
            a.H[o2] = Xmk;
        } else {
            /* [ch][p][M] */
            float2* dst = a.H + ((size_t)no * a.P + p) * a.M;
            dst[k] = Xk;
            dst[k2] = Xmk;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  K1: forward FFT of the new input block into the newest FDL slot (matrix layout)             */
/*  grid (nIn)                                                                                  */
/* ------------------------------------------------------------------------------------------ */
struct InFftArgs {
    const float* in;       /* [B][nIn][hop] */
    float2*      X;        /* [nKT][RS][nIn][32] */
    const float2* tw;
    const unsigned int* counters;
    int hop, nIn, M, logM, RS;
};

__global__ void input_fft_kernel(InFftArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    float2* stw = sm + SC_ALEN(a.M);
    const int ni = blockIdx.x, b = blockIdx.y;
    const int slot = (int)((a.counters[0] + (unsigned)b) % (unsigned)a.RS);
    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    load_real_block(sm, a.in + ((size_t)b * a.nIn + ni) * a.hop, a.hop, a.M, a.logM);
    __syncthreads();
    cfft_dif<false>(sm, a.M, a.logM, stw, wide);
    for (int k = threadIdx.x; k <= (a.M >> 1); k += blockDim.x) {
        float2 Xk, Xmk;
        int k2 = a.M - k;
        if (k == 0) {
            const float2 z = sm[0];
            Xk = make_float2(z.x + z.y, z.x - z.y);
            k2 = 0; Xmk = Xk;
        } else {
            fwd_split_pair(sm, k, a.M, a.logM, spl, Xk, Xmk);
        }
        a.X[(((size_t)(k >> 5) * a.RS + slot) * a.nIn + ni) * SC_BK + (k & 31)] = Xk;
        a.X[(((size_t)(k2 >> 5) * a.RS + slot) * a.nIn + ni) * SC_BK + (k2 & 31)] = Xmk;
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  K2: filter-streaming complex MAC  (the HBM-roofline kernel)                                 */
/* ------------------------------------------------------------------------------------------ */

struct MacArgs {
    const float2* H;               /* [unit][nIn][OTsz][32], unit = (ot*nKT + kt)*P + p */
    const float2* X;               /* [kt][RS ring slots][nIn][32] */
    float2*       Zp;              /* [partial slot][OTsz][32] (of this block) */
    const unsigned int* counters;
    const int*    ctaBase;
    long long totalStages;
    int nIn, OTsz, P, nKT, SNI, SPU, WGo, WGk, hints;
    int pLo, nP;                   /* this pass covers partitions [pLo, pLo + nP) of every group (full pass: 0, P) */
    int useCount; unsigned int count;   /* useCount: the block counter is given by the host (`count`) instead of read from
                                         * counters[0] -- a tail pass enqueued before the previous block's K3 has bumped it */
    int NS;                        /* pipeline depth (stages)                */
    int RS, blk, nB;               /* delay-line ring size; first block of this launch inside the batch; blocks in this launch */
    size_t zpStride;               /* float2 elements of Zp per block */
    int stageHBytes, stageXBytes;  /* shared-memory bytes reserved per stage */
};

#define SC_MAC_THREADS ((SC_MAC_CWARPS + 1) * 32)

template <int R>
__global__ void __launch_bounds__(SC_MAC_THREADS, 1) mac_kernel(MacArgs a)
{
    extern __shared__ __align__(128) unsigned char smraw[];
    /* carve: [NS][stageH] [NS][stageX] [reduce 8*R*32 float2] [barriers] */
    const int NS = a.NS;
    unsigned char* smH = smraw;
    unsigned char* smX = smH + (size_t)NS * a.stageHBytes;
    float2*   red  = reinterpret_cast<float2*>(smX + (size_t)NS * a.stageXBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(red + SC_MAC_CWARPS * R * 32);
    uint64_t* empt = full + NS;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long s0 = (a.totalStages * (long long)blockIdx.x) / gridDim.x;
    const long long s1 = (a.totalStages * (long long)(blockIdx.x + 1)) / gridDim.x;
    const int nIt = (int)(s1 - s0);
    /* decompose the first stage index ONCE; afterwards (sidx, p, kt) are advanced incrementally
     * (no integer division inside the streaming loop) */
    const long long unit0 = s0 / a.SPU;                /* unit inside this pass: grp * nP + (p - pLo) */
    const int sidx0 = (int)(s0 - unit0 * a.SPU);       /* stage inside the unit          */
    const long long grp0 = unit0 / a.nP;
    const int pHi = a.pLo + a.nP;
    const int p0  = a.pLo + (int)(unit0 - grp0 * a.nP);   /* filter partition            */
    const int kt0 = (int)(grp0 % a.nKT);               /* bin tile                       */

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empt[s], SC_MAC_CWARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == SC_MAC_CWARPS) {
        /* ===================== TMA producer (one elected lane) ===================== */
        if (lane == 0) {
            /* H is read exactly once per block: evict_first -- unless the whole filter set fits in L2 (hints == 2),
             * then it should stay there for the next block */
            const uint64_t polH = (a.hints == 2) ? l2_policy_evict_last() : l2_policy_evict_first();
            const uint64_t polX = l2_policy_evict_last();               /* the FDL is re-read by every output tile */
            const float2* srcH0 = a.H + (((size_t)grp0 * a.P + p0) * a.nIn + (size_t)sidx0 * a.SNI) * a.OTsz * SC_BK;
            const size_t skipH = (size_t)(a.P - a.nP) * a.nIn * a.OTsz * SC_BK;   /* partitions of a group outside this pass */
            const uint32_t rowH = (uint32_t)a.OTsz * (SC_BK * 8);
            int s = 0; uint32_t par = 1;
            /* the blocks of a batch are streamed back to back: the pipeline never drains between blocks */
            for (int blk = 0; blk < a.nB; ++blk) {
                const unsigned int cnt0 = a.useCount ? a.count : a.counters[0];
                const int head = (int)((cnt0 + (unsigned)(a.blk + blk)) % (unsigned)a.RS);   /* newest slot */
                const float2* srcH = srcH0;
                int sidx = sidx0, p = p0, kt = kt0;
                for (int i = 0; i < nIt; ++i) {
                    mbar_wait(&empt[s], par);
                    const int ni0  = sidx * a.SNI;
                    const int cnt  = min(a.SNI, a.nIn - ni0);
                    const uint32_t bytesH = (uint32_t)cnt * rowH;
                    const uint32_t bytesX = (uint32_t)cnt * (SC_BK * 8);
                    int slot = head - p; if (slot < 0) slot += a.RS;
                    const float2* srcX = a.X + (((size_t)kt * a.RS + slot) * a.nIn + ni0) * SC_BK;
                    mbar_expect_tx(&full[s], bytesH + bytesX);
                    if (a.hints) {
                        tma_bulk_g2s_hint(smH + (size_t)s * a.stageHBytes, srcH, bytesH, &full[s], polH);
                        tma_bulk_g2s_hint(smX + (size_t)s * a.stageXBytes, srcX, bytesX, &full[s], polX);
                    } else {
                        tma_bulk_g2s(smH + (size_t)s * a.stageHBytes, srcH, bytesH, &full[s]);
                        tma_bulk_g2s(smX + (size_t)s * a.stageXBytes, srcX, bytesX, &full[s]);
                    }
                    srcH += (size_t)cnt * a.OTsz * SC_BK;               /* stages are contiguous in H */
                    if (++sidx == a.SPU) { sidx = 0; if (++p == pHi) { p = a.pLo; srcH += skipH; if (++kt == a.nKT) kt = 0; } }
                    if (++s == NS) { s = 0; par ^= 1u; }
                }
            }
        }
        return;
    }

    /* ===================== consumers: 8 warps, lane = bin inside the tile ===================== */
    const int  wo     = warp % a.WGo;                  /* which group of R outputs */
    const int  wk     = warp / a.WGo;                  /* which share of the rows  */
    const int  nvalid = min(R, a.OTsz - wo * R);       /* outputs this warp really owns */
    const bool active = (wk < a.WGk) && (nvalid > 0);
    const int  rowStride = a.OTsz * SC_BK;             /* float2 per input row of a stage */
    float2 acc[R], tot[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { acc[j] = make_float2(0.f, 0.f); tot[j] = make_float2(0.f, 0.f); }

    int s = 0; uint32_t par = 0;
    for (int blk = 0; blk < a.nB; ++blk) {
        float2* dst = a.Zp + ((size_t)blk * a.zpStride) + (size_t)a.ctaBase[blockIdx.x] * a.OTsz * SC_BK;
        int sidx = sidx0, p = p0, kt = kt0;
        for (int i = 0; i < nIt; ++i) {
            const int  cnt    = min(a.SNI, a.nIn - sidx * a.SNI);
            const bool packed = (kt == 0) && (lane == 0);  /* bin 0 holds (DC, Nyquist): two real products */

            mbar_wait(&full[s], par);
            if (active) {
                const float2* Hs = reinterpret_cast<const float2*>(smH + (size_t)s * a.stageHBytes) + wo * R * SC_BK + lane;
                const float2* Xs = reinterpret_cast<const float2*>(smX + (size_t)s * a.stageXBytes) + lane;
                for (int r = wk; r < cnt; r += a.WGk) {
                    const float2 x = Xs[r * SC_BK];
                    const float xb = packed ? 0.f : x.y;   /* re -= h.y*xb ; im += h.x*xb */
                    const float xd = packed ? x.y : x.x;   /* im += h.y*xd                */
                    const float2* hrow = Hs + (size_t)r * rowStride;
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        if (j < nvalid) {
                            const float2 h = hrow[j * SC_BK];
                            acc[j].x = fmaf(h.x, x.x, acc[j].x);
                            acc[j].x = fmaf(-h.y, xb, acc[j].x);
                            acc[j].y = fmaf(h.x, xb, acc[j].y);
                            acc[j].y = fmaf(h.y, xd, acc[j].y);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empt[s]);
            if (++s == NS) { s = 0; par ^= 1u; }

            const bool last     = (i + 1 == nIt);
            const bool endUnit  = (sidx == a.SPU - 1);
            const bool endGroup = endUnit && (p == pHi - 1);
            if (endUnit || last) {
                /* two-level accumulation: per-unit sums are folded into the running total */
#pragma unroll
                for (int j = 0; j < R; ++j) { tot[j] = caddf(tot[j], acc[j]); acc[j] = make_float2(0.f, 0.f); }
            }
            if (endGroup || last) {
                /* end of this CTA's share of group (ot,kt): emit one partial tile */
                if (a.WGk > 1) {
#pragma unroll
                    for (int j = 0; j < R; ++j) red[(warp * R + j) * 32 + lane] = tot[j];
                    asm volatile("bar.sync 1, %0;" :: "n"(SC_MAC_CWARPS * 32) : "memory");
                    if (active && wk == 0) {
#pragma unroll
                        for (int j = 0; j < R; ++j) {
                            float2 v = tot[j];
                            for (int g = 1; g < a.WGk; ++g) v = caddf(v, red[((g * a.WGo + wo) * R + j) * 32 + lane]);
                            tot[j] = v;
                        }
                    }
                    asm volatile("bar.sync 1, %0;" :: "n"(SC_MAC_CWARPS * 32) : "memory");
                }
                if (active && wk == 0) {
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        if (j < nvalid) dst[(size_t)(wo * R + j) * SC_BK + lane] = tot[j];
                }
                dst += (size_t)a.OTsz * SC_BK;
#pragma unroll
                for (int j = 0; j < R; ++j) tot[j] = make_float2(0.f, 0.f);
            }
            if (endUnit) { sidx = 0; if (++p == pHi) { p = a.pLo; if (++kt == a.nKT) kt = 0; } }
            else ++sidx;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  K3: gather split-K partials, inverse FFT, overlap-add                                        */
/*  grid (nOutLocal)                                                                             */
/* ------------------------------------------------------------------------------------------ */
struct IfftArgs {
    const float2* Zp;      /* [B][nSlots][OTsz][32] */
    const int* grpStart;   /* partial slots of group g are grpStart[g] .. grpStart[g+1]-1 (consecutive) */
    const float2* Zp2;     /* second list of partial tiles (head pass after a pre-computed tail pass), or NULL */
    const int* grpStart2;
    const float2* headH;   /* != NULL: K3 adds the newest partition itself, sum_ni H_0[no][ni][k] * X_t[ni][k], straight from */
    const float2* headX;   /*          the filter / delay-line arrays (look-ahead latency path: no separate head pass)      */
    int nHead;             /* partitions K3 adds itself: p = 0 .. nHead-1 (the look-ahead depth) */
    int P, nIn, RS;
    const float2* tw;
    float* out;            /* [B][nOutLocal][hop] */
    float* tail;           /* [nOutLocal][hop] */
    float* zt;             /* batched path: [B][nOutLocal][2*hop] scaled inverse transforms */
    unsigned int* counters;
    size_t zpStride;       /* float2 elements of Zp per block */
    int hop, M, logM, nKT, OTsz, nOutLocal, B;
    float scale;           /* 1/N */
};

/* packed-bin-aware complex multiply-accumulate */
__device__ __forceinline__ void cmac_packed(float2& acc, float2 h, float2 x, bool packed)
{
    const float xb = packed ? 0.f : x.y;
    const float xd = packed ? x.y : x.x;
    acc.x = fmaf(h.x, x.x, acc.x);
    acc.x = fmaf(-h.y, xb, acc.x);
    acc.y = fmaf(h.x, xb, acc.y);
    acc.y = fmaf(h.y, xd, acc.y);
}

/* sum the split-K partial tiles of output `no` into sm[0..M) and run the inverse real FFT (bit-reversed result) */
__device__ __forceinline__ void gather_and_ifft(const IfftArgs& a, const float2* Zp, int no, float2* sm, float2* stw)
{
    const int ot = no / a.OTsz, nl = no - ot * a.OTsz;
    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    /* Four bins of a thread at a time, in three rounds of independent loads: the slot ranges of the four groups, then up to
     * GQ partial tiles per bin and list, then whatever is left.  Beside a tail pass that saturates HBM (look-ahead,
     * throughput regime) a DRAM round trip takes 2-3 us: the first version walked the bins one after the other, ~5
     * dependent round trips each, and K3 took 50 us there instead of 15 us alone.  Summation order per bin is unchanged:
     * ascending slot of list 1, then ascending slot of list 2, then the newest partition. */
    constexpr int GK = 4, GQ = 6;
    const size_t qs = (size_t)a.OTsz * SC_BK;
    for (int kb = threadIdx.x; kb < a.M; kb += GK * blockDim.x) {
        int q0[GK], q1[GK], r0[GK], r1[GK];
#pragma unroll
        for (int j = 0; j < GK; ++j) {
            const int k = kb + j * blockDim.x;
            q0[j] = q1[j] = r0[j] = r1[j] = 0;
            if (k < a.M) {
                const int g = ot * a.nKT + (k >> 5);
                q0[j] = __ldg(a.grpStart + g); q1[j] = __ldg(a.grpStart + g + 1);
                if (a.Zp2) { r0[j] = __ldg(a.grpStart2 + g); r1[j] = __ldg(a.grpStart2 + g + 1); }
            }
        }
        float2 zz[GK];
        /* one list after the other through the same registers (K3 must stay small enough to share an SM with a MAC CTA) */
#pragma unroll 1
        for (int list = 0; list < 2; ++list) {
            const float2* base = list ? a.Zp2 : Zp;
            if (!base) break;
            float2 v[GK][GQ];
#pragma unroll
            for (int j = 0; j < GK; ++j) {
                const int k = kb + j * blockDim.x;
                const int b0 = list ? r0[j] : q0[j], b1 = list ? r1[j] : q1[j];
                const float2* src = base + ((size_t)b0 * a.OTsz + nl) * SC_BK + (k & 31);
#pragma unroll
                for (int u = 0; u < GQ; ++u) v[j][u] = (b0 + u < b1) ? src[(size_t)u * qs] : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < GK; ++j) {
                const int k = kb + j * blockDim.x;
                const int b0 = list ? r0[j] : q0[j], b1 = list ? r1[j] : q1[j];
                float2 z = list ? zz[j] : make_float2(0.f, 0.f);
#pragma unroll
                for (int u = 0; u < GQ; ++u) if (b0 + u < b1) z = caddf(z, v[j][u]);
                const float2* src = base + ((size_t)(b0 + GQ) * a.OTsz + nl) * SC_BK + (k & 31);
                for (int q = b0 + GQ; q < b1; ++q, src += qs) z = caddf(z, src[0]);
                zz[j] = z;
            }
        }
#pragma unroll
        for (int j = 0; j < GK; ++j) {
            const int k = kb + j * blockDim.x;
            if (k < a.M) sm[padi(k, a.logM)] = zz[j];
        }
    }
    if (a.headH) {
        /* look-ahead latency path: the newest partition is added here (no separate head pass).  Partition 0 of group
         * (ot, kt): rows [ni][OTsz][32] of H, row ni of the newest delay-line slot */
        const int head = (int)(a.counters[0] % (unsigned)a.RS);
        for (int k = threadIdx.x; k < a.M; k += blockDim.x) {
            const int g = ot * a.nKT + (k >> 5);
            const float2* __restrict__ Hk = a.headH + ((size_t)g * a.P * a.nIn * a.OTsz + nl) * SC_BK + (k & 31);
            const float2* __restrict__ Xk = a.headX + ((size_t)(k >> 5) * a.RS + head) * a.nIn * SC_BK + (k & 31);
            const bool packed = (k == 0);
            float2 acc = make_float2(0.f, 0.f);
            for (int p = 0; p < a.nHead; ++p) {                 /* partition p: rows p * nIn .. of the unit, ring slot head - p */
                int slot = head - p; if (slot < 0) slot += a.RS;
                const float2* __restrict__ Hp = Hk + (size_t)p * a.nIn * qs;
                const float2* __restrict__ Xp = Xk + ((ptrdiff_t)slot - head) * a.nIn * SC_BK;
#pragma unroll 16
                for (int ni = 0; ni < a.nIn; ++ni)
                    cmac_packed(acc, __ldg(Hp + (size_t)ni * qs), __ldg(Xp + (size_t)ni * SC_BK), packed);
            }
            sm[padi(k, a.logM)] = caddf(sm[padi(k, a.logM)], acc);      /* same thread wrote it above */
        }
    }
    __syncthreads();
    inv_split_all(sm, a.M, a.logM, spl);
    cfft_dif<true>(sm, a.M, a.logM, stw, wide);
}

/* one block: grid (nOutLocal); 256 threads (512 when it adds the newest partition itself), <= 128 registers so that a
 * 256-thread CTA fits beside a MAC CTA */
__global__ void __launch_bounds__(512, 1) ifft_ola_kernel(IfftArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    const int no = blockIdx.x;
    gather_and_ifft(a, a.Zp, no, sm, sm + SC_ALEN(a.M));
    ola_store(sm, a.hop, a.logM, a.scale, a.out + (size_t)no * a.hop, a.tail + (size_t)no * a.hop);
    advance_block_counter(a.counters, gridDim.x);
}

/* batch of B blocks, step 1: grid (nOutLocal, B) -> zt[b][no][0..2*hop) = z/N */
__global__ void __launch_bounds__(512, 1) ifft_batch_kernel(IfftArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    const int no = blockIdx.x, b = blockIdx.y;
    gather_and_ifft(a, a.Zp + (size_t)b * a.zpStride, no, sm, sm + SC_ALEN(a.M));
    float* z = a.zt + ((size_t)b * a.nOutLocal + no) * 2 * a.hop;
    for (int i = threadIdx.x; i < 2 * a.hop; i += blockDim.x) z[i] = time_sample(sm, i, a.logM) * a.scale;
}

/* batch step 2: overlap-add along the batch (reference .c:230-233), one thread per (no, i); same two
 * addends per sample as the one-block kernel, so both paths give identical bits */
__global__ void ola_batch_kernel(IfftArgs a)
{
    const size_t n = (size_t)a.nOutLocal * a.hop;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) {
        const int no = (int)(idx / a.hop), i = (int)(idx - (size_t)no * a.hop);
        float prev = a.tail[idx];
        for (int b = 0; b < a.B; ++b) {
            const float* z = a.zt + ((size_t)b * a.nOutLocal + no) * 2 * a.hop;
            a.out[(size_t)b * n + idx] = z[i] + prev;
            prev = z[i + a.hop];
        }
        a.tail[idx] = prev;
    }
    /* last CTA advances the block counter by B */
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&a.counters[1], 1u);
        if (t == gridDim.x - 1) {
            a.counters[1] = 0;
            __threadfence();
            atomicAdd(&a.counters[0], (unsigned)a.B);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  multiConv: everything for one channel in one CTA  (reference .c:388-413)                    */
/*  grid (nCH) ; shared memory: 3*M float2 (two work buffers + twiddles)                        */
/* ------------------------------------------------------------------------------------------ */
struct MultiArgs {
    const float* in;       /* [nCH][hop] */
    float* out;            /* [nCH][hop] */
    const float2* H;       /* [nCH][P][M] */
    float2* X;             /* [nCH][RS][M] ring, RS = P + maxBatch */
    const float2* tw;
    float* tail;
    float* zt;             /* batched path: [B][nCH][2*hop] */
    unsigned int* counters;
    int hop, M, logM, P, RS, nCH;
    int B, G;              /* batched path: blocks in the batch, consecutive blocks per CTA (MAC + inverse FFT) */
    int Q;                 /* batched path: consecutive blocks per forward-FFT CTA */
    const float2* wT1;     /* warp-FFT tables ([M/32][32] each, safconv_wfft.cuh) or NULL */
    const float2* wT2;
    float scale;
};


__global__ void multi_fused_kernel(MultiArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    float2* A = sm;                       /* FFT work array (padded) */
    float2* B = sm + SC_ALEN(a.M);        /* spectrum of the new block, natural order */
    float2* stw = B + a.M;
    const int c = blockIdx.x;
    const int head = (int)(a.counters[0] % (unsigned)a.RS);
    float2* Xc = a.X + (size_t)c * a.RS * a.M;
    const float2* Hc = a.H + (size_t)c * a.P * a.M;

    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    load_real_block(A, a.in + (size_t)c * a.hop, a.hop, a.M, a.logM);
    __syncthreads();
    cfft_dif<false>(A, a.M, a.logM, stw, wide);
    float2* Xnew = Xc + (size_t)head * a.M;
    for (int k = threadIdx.x; k <= (a.M >> 1); k += blockDim.x) {
        float2 Xk, Xmk;
        int k2 = a.M - k;
        if (k == 0) {
            const float2 z = A[0];
            Xk = make_float2(z.x + z.y, z.x - z.y);
            k2 = 0; Xmk = Xk;
        } else {
            fwd_split_pair(A, k, a.M, a.logM, spl, Xk, Xmk);
        }
        B[k] = Xk;  B[k2] = Xmk;
        Xnew[k] = Xk;  Xnew[k2] = Xmk;
    }
    __syncthreads();
    /* Z[k] = sum_p H[p][k] * X[t-p][k] ; slot(p) = head - p (mod P) */
    for (int k = threadIdx.x; k < a.M; k += blockDim.x) {
        const bool packed = (k == 0);
        float2 acc = make_float2(0.f, 0.f);
        cmac_packed(acc, __ldg(Hc + k), B[k], packed);
        int slot = head;
#pragma unroll 8
        for (int p = 1; p < a.P; ++p) {
            slot = (slot == 0) ? a.RS - 1 : slot - 1;
            cmac_packed(acc, __ldg(Hc + (size_t)p * a.M + k), Xc[(size_t)slot * a.M + k], packed);
        }
        A[padi(k, a.logM)] = acc;
    }
    __syncthreads();
    inv_split_all(A, a.M, a.logM, spl);
    cfft_dif<true>(A, a.M, a.logM, stw, wide);
    ola_store(A, a.hop, a.logM, a.scale, a.out + (size_t)c * a.hop, a.tail + (size_t)c * a.hop);
    advance_block_counter(a.counters, gridDim.x);
}

/* ---- multiConv, batch of B device-resident blocks: every (channel, block) pair is independent once the
 * spectra of all B blocks are in the ring, so the work is two fully parallel launches + the overlap-add chain ---- */

/* grid (nCH, ceil(B/Q)), Q <= SC_MULTI_FFT_Q: forward FFTs of Q consecutive blocks b of channel c into ring slots
 * (counter + b) % RS; the Q transforms advance together (one barrier per pass for all of them) */
#define SC_MULTI_FFT_Q 4
__global__ void multi_fft_batch_kernel(MultiArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    const int MP = SC_ALEN(a.M);
    float2* stw = sm + (size_t)a.Q * MP;
    const int c = blockIdx.x, b0 = blockIdx.y * a.Q;
    const int nq = min(a.Q, a.B - b0);
    const unsigned int count = a.counters[0];
    const bool wide = fft_use_wide(a.M, 1);            /* same core as the one-block kernel: identical bits */
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    for (int q = 0; q < nq; ++q)
        load_real_block(sm + (size_t)q * MP, a.in + ((size_t)(b0 + q) * a.nCH + c) * a.hop, a.hop, a.M, a.logM);
    __syncthreads();
    cfft_dif_batch<false>(sm, a.M, a.logM, stw, nq, wide);
    const int per = (a.M >> 1) + 1;
    for (int it = threadIdx.x; it < nq * per; it += blockDim.x) {
        const int q = it / per, k = it - q * per;
        const float2* s = sm + (size_t)q * MP;
        const int slot = (int)((count + (unsigned)(b0 + q)) % (unsigned)a.RS);
        float2* Xnew = a.X + ((size_t)c * a.RS + slot) * a.M;
        float2 Xk, Xmk;
        int k2 = a.M - k;
        if (k == 0) {
            const float2 z = s[0];
            Xk = make_float2(z.x + z.y, z.x - z.y);
            k2 = 0; Xmk = Xk;
        } else {
            fwd_split_pair(s, k, a.M, a.logM, spl, Xk, Xmk);
        }
        Xnew[k] = Xk;  Xnew[k2] = Xmk;
    }
}

/* grid (nCH, ceil(B/G)): for G consecutive blocks b of channel c: Z = sum_p H_p * X_{t_b - p} (same order as the
 * one-block kernel), inverse FFT -> zt[b][c][0..2*hop).  The blocks of a CTA share the channel's filter spectra and
 * all but one delay-line slot with their neighbour, so after the first block the MAC reads come out of L1. */
__global__ void multi_mac_ifft_batch_kernel(MultiArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    float2* stw = sm + SC_ALEN(a.M);
    const int c = blockIdx.x;
    const float2* __restrict__ Xc = a.X + (size_t)c * a.RS * a.M;
    const float2* __restrict__ Hc = a.H + (size_t)c * a.P * a.M;
    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    const unsigned int count = a.counters[0];
    for (int b = blockIdx.y * a.G; b < min(a.B, (int)(blockIdx.y + 1) * a.G); ++b) {
        const int head = (int)((count + (unsigned)b) % (unsigned)a.RS);
        for (int k = threadIdx.x; k < a.M; k += blockDim.x) {
            const bool packed = (k == 0);
            float2 acc = make_float2(0.f, 0.f);
            int slot = head;
#pragma unroll 8
            for (int p = 0; p < a.P; ++p) {
                cmac_packed(acc, __ldg(Hc + (size_t)p * a.M + k), __ldg(Xc + (size_t)slot * a.M + k), packed);
                slot = (slot == 0) ? a.RS - 1 : slot - 1;
            }
            sm[padi(k, a.logM)] = acc;
        }
        __syncthreads();
        inv_split_all(sm, a.M, a.logM, spl);
        cfft_dif<true>(sm, a.M, a.logM, stw, wide);
        float* z = a.zt + ((size_t)b * a.nCH + c) * 2 * a.hop;
        for (int i = threadIdx.x; i < 2 * a.hop; i += blockDim.x) z[i] = time_sample(sm, i, a.logM) * a.scale;
        __syncthreads();
    }
}

/* Same work with the operands in REGISTERS: one thread per bin (blockDim = M <= 1024), the P <= PT filter values of
 * the bin and a sliding window of its last P delay-line values stay in registers across the G consecutive blocks
 * of the CTA -- per block and bin one new 8-byte load instead of 2P.  Same summation order, identical bits.
 * Used for M <= 512. */
template <int PT>
__global__ void __launch_bounds__(512, 1) multi_mac_ifft_batch_reg_kernel(MultiArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    float2* stw = sm + SC_ALEN(a.M);
    const int c = blockIdx.x, k = threadIdx.x;
    const float2* __restrict__ Xc = a.X + (size_t)c * a.RS * a.M + k;
    const float2* __restrict__ Hc = a.H + (size_t)c * a.P * a.M + k;
    const bool wide = fft_use_wide(a.M, 1);
    const unsigned int count = a.counters[0];
    const int b0 = blockIdx.y * a.G, b1 = min(a.B, b0 + a.G);
    const bool packed = (k == 0);
    float2 h[PT], xw[PT];
    int slot = (int)((count + (unsigned)b0) % (unsigned)a.RS);          /* newest slot of block b0 */
    {
        int sl = slot;
#pragma unroll
        for (int p = 0; p < PT; ++p) {
            h[p]  = (p < a.P) ? __ldg(Hc + (size_t)p * a.M) : make_float2(0.f, 0.f);
            xw[p] = (p >= 1 && p < a.P) ? __ldg(Xc + (size_t)sl * a.M) : make_float2(0.f, 0.f);
            sl = (sl == 0) ? a.RS - 1 : sl - 1;
        }
    }
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    for (int b = b0; b < b1; ++b) {
        xw[0] = __ldg(Xc + (size_t)slot * a.M);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int p = 0; p < PT; ++p)
            if (p < a.P) cmac_packed(acc, h[p], xw[p], packed);
        sm[padi(k, a.logM)] = acc;
        __syncthreads();
        inv_split_all(sm, a.M, a.logM, spl);
        cfft_dif<true>(sm, a.M, a.logM, stw, wide);
        float* z = a.zt + ((size_t)b * a.nCH + c) * 2 * a.hop;
        for (int i = threadIdx.x; i < 2 * a.hop; i += blockDim.x) z[i] = time_sample(sm, i, a.logM) * a.scale;
        __syncthreads();
#pragma unroll
        for (int p = PT - 1; p >= 1; --p) xw[p] = xw[p - 1];
        slot = (slot + 1 == a.RS) ? 0 : slot + 1;
    }
}


/* ---- batched multiConv on the warp-level register FFT (safconv_wfft.cuh), 64 <= M <= 1024 ---- */

/* forward: grid (nCH, ceil(B/8)), 256 threads; warp q transforms block 8 blockIdx.y + q of channel c into its ring slot.
 * The spectrum leaves the registers through a warp-private padded tile so that the ring is written in 256-byte rows. */
template <int R>
__global__ void __launch_bounds__(256) multi_fft_w_kernel(MultiArgs a)
{
    constexpr int M = 32 * R, LOGR = wf_log2(R), TS = M + 32;
    extern __shared__ __align__(16) float2 smt[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x, b = blockIdx.y * 8 + warp;
    if (b >= a.B) return;
    float2* tile = smt + (size_t)warp * TS;
    const WfftLane L = wfft_lane_init<false>(a.tw, M, lane);
    const float* x = a.in + ((size_t)b * a.nCH + c) * a.hop;
    const bool vec = ((a.hop & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
    float2 v[R], X[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int n = lane + 32 * i;
        v[i] = make_float2(0.f, 0.f);
        if (vec) { if (2 * n < a.hop) v[i] = __ldg(reinterpret_cast<const float2*>(x) + n); }
        else {
            if (2 * n < a.hop)     v[i].x = __ldg(x + 2 * n);
            if (2 * n + 1 < a.hop) v[i].y = __ldg(x + 2 * n + 1);
        }
    }
    wfft<R, false>(v, a.wT1, lane, L);
    wfft_fwd_split<R>(v, X, a.wT2, lane, 0.5f);
    const int k1 = (int)(__brev((unsigned)lane) >> 27);
#pragma unroll
    for (int i = 0; i < R; ++i) tile[wf_bitrev(i, LOGR) + (R + 1) * k1] = X[i];    /* bin k = k2 + R k1 at k + (k >> LOGR) */
    __syncwarp();
    const int slot = (int)((a.counters[0] + (unsigned)b) % (unsigned)a.RS);
    float2* Xnew = a.X + ((size_t)c * a.RS + slot) * M;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int k = lane + 32 * i;
        Xnew[k] = tile[k + (k >> LOGR)];
    }
}

/* MAC + inverse: grid (nCH, ceil(B/R)), M = 32 R threads.  Phase 1, one thread per bin: the bin's P <= PT filter
 * values and a sliding window of its last P delay-line values stay in registers across the R consecutive blocks of
 * the CTA (one new 8-byte load per block), the R output spectra go to shared memory.  Phase 2, one warp per block:
 * inverse split while loading, inverse FFT in registers, 1/N, time samples back through the warp's tile to zt. */
template <int R, int PT>
__global__ void __launch_bounds__(32 * R, 1) multi_mac_ifft_w_kernel(MultiArgs a)
{
    constexpr int M = 32 * R, LOGR = wf_log2(R), TS = M + 32;
    extern __shared__ __align__(16) float2 smt[];              /* [R][TS] */
    const int c = blockIdx.x, k = threadIdx.x;
    const int b0 = blockIdx.y * R, nb = min(R, a.B - b0);
    const unsigned int count = a.counters[0];
    {
        const float2* __restrict__ Xc = a.X + (size_t)c * a.RS * M + k;
        const float2* __restrict__ Hc = a.H + (size_t)c * a.P * M + k;
        const bool packed = (k == 0);
        float2 h[PT], xw[PT];
        int slot = (int)((count + (unsigned)b0) % (unsigned)a.RS);
        {
            int sl = slot;
#pragma unroll
            for (int p = 0; p < PT; ++p) {
                h[p]  = (p < a.P) ? __ldg(Hc + (size_t)p * M) : make_float2(0.f, 0.f);
                xw[p] = (p >= 1 && p < a.P) ? __ldg(Xc + (size_t)sl * M) : make_float2(0.f, 0.f);
                sl = (sl == 0) ? a.RS - 1 : sl - 1;
            }
        }
        for (int bb = 0; bb < nb; ++bb) {
            xw[0] = __ldg(Xc + (size_t)slot * M);
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int p = 0; p < PT; ++p)
                if (p < a.P) cmac_packed(acc, h[p], xw[p], packed);
            smt[(size_t)bb * TS + k] = acc;
#pragma unroll
            for (int p = PT - 1; p >= 1; --p) xw[p] = xw[p - 1];
            slot = (slot + 1 == a.RS) ? 0 : slot + 1;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= nb) return;
    float2* tile = smt + (size_t)warp * TS;
    const WfftLane L = wfft_lane_init<true>(a.tw, M, lane);
    float2 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int kk = lane + 32 * i;
        const float2 A = tile[kk], B = tile[(M - kk) & (M - 1)];
        const float2 E = make_float2(A.x + B.x, A.y - B.y);
        const float2 D = make_float2(A.x - B.x, A.y + B.y);
        const float2 O = cmul_conjb(D, __ldg(a.tw + kk));
        v[i] = make_float2(E.x - O.y, E.y + O.x);
        if (i == 0 && lane == 0) v[i] = make_float2(A.x + A.y, A.x - A.y);      /* (DC, Nyquist) */
    }
    __syncwarp();
    wfft<R, true>(v, a.wT1, lane, L);
    const int k1 = (int)(__brev((unsigned)lane) >> 27);
#pragma unroll
    for (int i = 0; i < R; ++i)                                /* z[n], n = n2 + R k1, stored at n + (n >> LOGR) */
        tile[wf_bitrev(i, LOGR) + (R + 1) * k1] = make_float2(v[i].x * a.scale, v[i].y * a.scale);
    __syncwarp();
    float* z = a.zt + ((size_t)(b0 + warp) * a.nCH + c) * 2 * a.hop;
    const float* ts = reinterpret_cast<const float*>(tile);
    for (int s = lane; s < 2 * a.hop; s += 32) {
        const int n = s >> 1;
        z[s] = ts[2 * (n + (n >> LOGR)) + (s & 1)];
    }
}

template <int R>
static int multi_w_launch(const scdev_plan* pl, MultiArgs& a, int nBlocks, int which, cudaStream_t st)
{
    constexpr int M = 32 * R;
    const size_t tile = (size_t)(M + 32) * sizeof(float2);
    if (which == 0) {
        dim3 grid(pl->nOutLocal, (nBlocks + 7) / 8);
        if (8 * tile > 48 * 1024) SC_CHECK(sc_optin_smem(multi_fft_w_kernel<R>));
        multi_fft_w_kernel<R><<<grid, 256, 8 * tile, st>>>(a);
    } else {
        dim3 grid(pl->nOutLocal, (nBlocks + R - 1) / R);
        const size_t smem = (size_t)R * tile;
#define SC_MW_LAUNCH(PT) do {                                                                                        \
            if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(multi_mac_ifft_w_kernel<R, PT>)); \
            multi_mac_ifft_w_kernel<R, PT><<<grid, M, smem, st>>>(a); } while (0)
        if (pl->P <= 2) SC_MW_LAUNCH(2); else if (pl->P <= 4) SC_MW_LAUNCH(4); else if (pl->P <= 8) SC_MW_LAUNCH(8); else SC_MW_LAUNCH(16);
#undef SC_MW_LAUNCH
    }
    return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------ */
/*  small_fused_kernel: K1 + K2 + K3 of the MATRIX convolver in ONE launch for small real-time    */
/*  problems (C1 / C2 class: a few hundred KB of filter spectra, latency- not bandwidth-bound).   */
/*  One CTA per output channel: forward FFTs of all nIn inputs (redundant across the few CTAs,     */
/*  CTA 0 also stores them in the delay-line ring), the sum over partitions x inputs straight out  */
/*  of L2, one inverse FFT, overlap-add.  `in` / `out` may be page-locked HOST memory (zero-copy): */
/*  then a whole saf_matrixConv_apply is one kernel launch and one stream synchronisation.        */
/*  Works on the same H / delay-line layouts as the three-kernel path.                             */
/*  shared memory: nIn padded FFT arrays | nIn natural-order spectra | twiddles | reduction        */
/* ------------------------------------------------------------------------------------------ */
struct SmallArgs {
    const float* in;       /* [nIn][hop]      (device or mapped host) */
    float* out;            /* [nOutLocal][hop] (device or mapped host) */
    const float2* H;       /* [ot][kt][p][ni][OTsz][32] */
    float2* X;             /* [kt][RS][nIn][32] ring */
    const float2* tw;
    float* tail;
    unsigned int* counters;
    int hop, M, logM, P, nIn, nKT, OTsz, RS;
    float scale;
};

__global__ void small_fused_kernel(SmallArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    const int MP = SC_ALEN(a.M);
    float2* W   = sm;                                  /* nIn padded FFT work arrays (array 0 is re-used for Z) */
    float2* Xn  = sm + (size_t)a.nIn * MP;             /* nIn packed spectra of the new block, natural order    */
    float2* stw = Xn + (size_t)a.nIn * a.M;
    float2* red = stw + a.M + sc_split_len(a.M);       /* [groups][M] partial sums                              */
    const int no = blockIdx.x;
    const int ot = no / a.OTsz, nl = no - ot * a.OTsz;
    const unsigned int count = a.counters[0];          /* issued with the input loads below; graph-replay safe */
    const int head = (int)(count % (unsigned)a.RS);
    const int tid = threadIdx.x, T = blockDim.x;
    /* overlap tails are fetched now, used at the very end (hop <= 4 * blockDim) */
    float tl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = tid + u * T; tl[u] = (i < a.hop) ? a.tail[(size_t)no * a.hop + i] : 0.f; }

    /* input block: all loads of a thread are issued before the first store (the source may be mapped host
     * memory: one PCIe round trip for the whole block instead of one per channel) */
    {
        const bool vec = ((a.hop & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 7) == 0);
        const int total = a.nIn * a.M;
        for (int base = tid; base < total; base += 4 * T) {
            float2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * T;
                v[u] = make_float2(0.f, 0.f);
                if (idx < total) {
                    const int ni = idx >> a.logM, n = idx & (a.M - 1);
                    const float* x = a.in + (size_t)ni * a.hop;
                    if (vec) { if (2 * n < a.hop) v[u] = __ldg(reinterpret_cast<const float2*>(x) + n); }
                    else {
                        if (2 * n < a.hop)     v[u].x = __ldg(x + 2 * n);
                        if (2 * n + 1 < a.hop) v[u].y = __ldg(x + 2 * n + 1);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * T;
                if (idx < total) W[(size_t)(idx >> a.logM) * MP + padi(idx & (a.M - 1), a.logM)] = v[u];
            }
        }
    }
    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    __syncthreads();
    cfft_dif_batch<false>(W, a.M, a.logM, stw, a.nIn, wide);
    {
        const int per = (a.M >> 1) + 1;
        for (int it = tid; it < a.nIn * per; it += T) {
            const int ni = it / per, k = it - ni * per;
            const float2* s = W + (size_t)ni * MP;
            float2 Xk, Xmk;
            int k2 = a.M - k;
            if (k == 0) {
                const float2 z = s[0];
                Xk = make_float2(z.x + z.y, z.x - z.y);
                k2 = 0; Xmk = Xk;
            } else {
                fwd_split_pair(s, k, a.M, a.logM, spl, Xk, Xmk);
            }
            Xn[(size_t)ni * a.M + k] = Xk;  Xn[(size_t)ni * a.M + k2] = Xmk;
            if (no == 0) {
                a.X[(((size_t)(k >> 5) * a.RS + head) * a.nIn + ni) * SC_BK + (k & 31)] = Xk;
                a.X[(((size_t)(k2 >> 5) * a.RS + head) * a.nIn + ni) * SC_BK + (k2 & 31)] = Xmk;
            }
        }
    }
    __syncthreads();
    /* Z[k] = sum_p sum_ni H_p[no][ni][k] * X_{t-p}[ni][k]; the (p, ni) list is split over G = T / M thread groups */
    {
        const int G = (T >= a.M) ? T / a.M : 1;
        const int nTerms = a.P * a.nIn;
        for (int k0 = 0; k0 < a.M; k0 += T) {                      /* one round unless M > T */
            const int g = (T >= a.M) ? tid / a.M : 0;
            const int k = (T >= a.M) ? tid - g * a.M : k0 + tid;
            if (g < G && k < a.M) {
                const bool packed = (k == 0);
                const int kt = k >> 5, b = k & 31;
                float2 acc = make_float2(0.f, 0.f);
                const int t0 = (int)(((long long)nTerms * g) / G), t1 = (int)(((long long)nTerms * (g + 1)) / G);
                int p = t0 / a.nIn, ni = t0 - p * a.nIn;
                int slot = head - p; if (slot < 0) slot += a.RS;
                const float2* Hk = a.H + ((size_t)(ot * a.nKT + kt) * a.P * a.nIn * a.OTsz + nl) * SC_BK + b;
                const float2* Xk = a.X + (size_t)kt * a.RS * a.nIn * SC_BK + b;
#pragma unroll 8
                for (int t = t0; t < t1; ++t) {
                    const float2 h = __ldg(Hk + (size_t)(p * a.nIn + ni) * a.OTsz * SC_BK);
                    const float2 x = (p == 0) ? Xn[(size_t)ni * a.M + k] : Xk[((size_t)slot * a.nIn + ni) * SC_BK];
                    cmac_packed(acc, h, x, packed);
                    if (++ni == a.nIn) { ni = 0; ++p; slot = (slot == 0) ? a.RS - 1 : slot - 1; }
                }
                if (G > 1) red[(size_t)g * a.M + k] = acc;
                else W[padi(k, a.logM)] = acc;
            }
        }
        if (G > 1) {
            __syncthreads();
            for (int k = tid; k < a.M; k += T) {
                float2 z = red[k];
                for (int g = 1; g < G; ++g) z = caddf(z, red[(size_t)g * a.M + k]);
                W[padi(k, a.logM)] = z;
            }
        }
    }
    __syncthreads();
    inv_split_all(W, a.M, a.logM, spl);
    cfft_dif<true>(W, a.M, a.logM, stw, wide);
    {   /* overlap-add (reference .c:230-233) */
        float* out = a.out + (size_t)no * a.hop;
        float* tail = a.tail + (size_t)no * a.hop;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = tid + u * T;
            if (i < a.hop) {
                out[i]  = time_sample(W, i, a.logM) * a.scale + tl[u];
                tail[i] = time_sample(W, i + a.hop, a.logM) * a.scale;
            }
        }
    }
    /* block counter: the last CTA to get here bumps it (every CTA has read it long before; the kernel boundary
     * publishes the store, so no fences are needed) */
    if (tid == 0) {
        const unsigned int t = atomicAdd(&a.counters[1], 1u);
        if (t == gridDim.x - 1) { a.counters[1] = 0; a.counters[0] = count + 1u; }
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  small_cluster_kernel<R>: the same block (K1 + K2 + K3 in one launch, mapped host buffers      */
/*  allowed) for M = 32 R = 64 .. 1024 on ONE thread-block cluster of C CTAs, without redundant    */
/*  work and without block-wide FFT passes:                                                       */
/*    A  forward FFTs: input ni -> CTA ni % C, warp ni / C; one WARP per transform on the         */
/*       register FFT (safconv_wfft.cuh: compile-time M, no shared memory, no barriers), real     */
/*       split in registers, spectrum into the delay-line ring (global, L2)                       */
/*       -- cluster barrier (release / acquire: the ring slot is visible to all CTAs) --          */
/*    B  per-bin sums: unit (output, 32-bin tile) u -> CTA u % C; the P*nIn terms of a unit are    */
/*       split over the warps of the CTA (lane = bin: 256-byte coalesced rows of H and of the     */
/*       ring), fixed-order reduction through shared memory, the tile is written into the shared  */
/*       memory of the CTA that owns the output (distributed shared memory)                       */
/*       -- cluster barrier --                                                                    */
/*    C  output no -> CTA no % C, warp no / C: inverse split, register inverse FFT, overlap-add,  */
/*       coalesced stores of out / tail.                                                          */
/*  ncu (profiles/r02_C2_full_summary.csv): small_fused_kernel at C2 = 2 CTAs, 90 k warp           */
/*  instructions, 30 % issue-active, 25 us cold -- every CTA repeats all 25 input FFTs on the     */
/*  shared-memory core and walks the 100 terms of its output with 4 thread groups.                */
/* ------------------------------------------------------------------------------------------ */
#define SC_CL_THREADS 512
#define SC_CL_WARPS   (SC_CL_THREADS / 32)
#define SC_CL_MAXC    8

struct SmallCArgs {
    const float* in;       /* [nIn][hop]       (device or mapped host) */
    float* out;            /* [nOutLocal][hop] (device or mapped host) */
    const float2* H;       /* [ot][kt][p][ni][OTsz][32] */
    float2* X;             /* [kt][RS][nIn][32] ring */
    const float2* tw;      /* W_N^e, e < M */
    const float2* wT1;     /* warp-FFT tables [R][32] */
    const float2* wT2;
    float* tail;
    unsigned int* counters;
    int hop, P, nIn, nOut, nKT, OTsz, RS;
    float scale;
    volatile unsigned int* done;   /* != NULL: page-locked host word that receives `seq` once every output sample has been written */
    unsigned int seq;
    unsigned long long* stamps;    /* debugging (SAFCONV_KSTAMPS=1): %globaltimer of CTA 0 at the phase boundaries */
};
__device__ __forceinline__ void cl_stamp(unsigned long long* stamps, int i)
{
    if (stamps && threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); stamps[i] = t; }
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ uint32_t cl_ctarank()  { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cl_store_f2(const void* localSmem, uint32_t rank, float2 v)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(localSmem)), "r"(rank));
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" :: "r"(remote), "f"(v.x), "f"(v.y) : "memory");
}

/* system-scope loads: data the HOST rewrites while a resident kernel is running (input blocks, the doorbell) */
__device__ __forceinline__ unsigned int ld_sys_u32(const volatile unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const volatile unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

/* One block on the cluster.  RES = true: called from the resident kernel's loop -- nothing that changes between blocks may
 * come through a non-coherent or L1-cached path (input: system scope; ring, counter, tails: L2). */
template <int R, bool RES>
__device__ __forceinline__ void small_cluster_block(const SmallCArgs& a, float2* smc)
{
    constexpr int M = 32 * R, LOGR = wf_log2(R);
    constexpr int ZS = M + 33;                                  /* float2 per owned output: spectrum, then time-domain staging */
    float2* Zs  = smc;                                          /* [owned outputs][ZS] */
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = (int)cl_ctarank(), C = (int)cl_nctarank();
    const int nOwn = (a.nOut - c + C - 1) / C;                  /* outputs c, c + C, ... */
    const int nOwnMax = (a.nOut + C - 1) / C;
    float2* red = smc + (size_t)nOwnMax * ZS;                   /* [warps][32] partial sums */
    cl_stamp(a.stamps, 0);
    /* the input block of this warp's (first) transform: its PCIe round trip (mapped host memory) is the longest latency
     * of the kernel, so these loads are issued before anything else */
    const bool vecIn = ((a.hop & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 7) == 0);
    float2 v[R];
    int niA = c + C * warp;
    auto load_block = [&](int ni) {
        const float* x = a.in + (size_t)ni * a.hop;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int n = lane + 32 * i;
            v[i] = make_float2(0.f, 0.f);
            /* RES: ld.global.cv -- fetched again on every execution (the host rewrites the block between doorbells), coalesced like any load */
            if (vecIn) { if (2 * n < a.hop) v[i] = RES ? __ldcv(reinterpret_cast<const float2*>(x) + n) : __ldg(reinterpret_cast<const float2*>(x) + n); }
            else {
                if (2 * n < a.hop)     v[i].x = RES ? __ldcv(x + 2 * n) : __ldg(x + 2 * n);
                if (2 * n + 1 < a.hop) v[i].y = RES ? __ldcv(x + 2 * n + 1) : __ldg(x + 2 * n + 1);
            }
        }
    };
    if (R <= 16 && niA < a.nIn) load_block(niA);                /* R = 32 keeps its registers for the prologue: loaded in phase A */
    const unsigned int count = RES ? __ldcg(a.counters) : a.counters[0];
    const int head = (int)(count % (unsigned)a.RS);
    const WfftLane Lf = wfft_lane_init<false>(a.tw, M, lane);
    WfftLane Li = Lf;
    Li.w16.y = -Li.w16.y; Li.w8.y = -Li.w8.y; Li.w4.y = -Li.w4.y; Li.w2.y = -Li.w2.y;
    const int k1 = (int)(__brev((unsigned)lane) >> 27);
    /* the twiddle tables of both transforms (L2 -> L1 while the input block is still on its way over PCIe) */
    for (int l = threadIdx.x; l < (2 * M) / 16; l += SC_CL_THREADS) prefetch_l1(a.wT1 + 16 * l);     /* T1 and T2 are one array */
    for (int l = threadIdx.x; l < M / 16; l += SC_CL_THREADS) prefetch_l1(a.tw + 16 * l);
    /* overlap tail of the output this warp owns: fetched now, used at the very end (hop <= M) */
    const int noMine = c + C * warp;
    constexpr bool TLPRE = (R <= 16);                           /* R = 32: the registers are needed by the transform itself */
    float tl[TLPRE ? R : 1];
    if (TLPRE) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const int i = lane + 32 * u;
            tl[TLPRE ? u : 0] = (warp < nOwn && i < a.hop) ? __ldcg(a.tail + (size_t)noMine * a.hop + i) : 0.f;
        }
    }

    /* the rows phase B of this CTA will read -- its units' filter rows and the delay-line rows of the OLDER blocks -- are
     * pulled into L1 now: lane = row (256 bytes = two 128-byte lines) */
    {
        const int nUnits = a.nOut * a.nKT, nTerms = a.P * a.nIn;
        for (int u = c; u < nUnits; u += C) {
            const int kt = u / a.nOut, no = u - kt * a.nOut;
            const int ot = no / a.OTsz, nl = no - ot * a.OTsz;
            const float2* Hk = a.H + ((size_t)(ot * a.nKT + kt) * a.P * a.nIn * a.OTsz + nl) * SC_BK;
            const float2* Xk = a.X + (size_t)kt * a.RS * a.nIn * SC_BK;
            for (int t = threadIdx.x; t < nTerms; t += SC_CL_THREADS) {
                const int p = t / a.nIn, ni = t - p * a.nIn;
                const float2* h = Hk + (size_t)t * a.OTsz * SC_BK;
                prefetch_l1(h); prefetch_l1(h + 16);
                if (!RES && p > 0) {
                    int slot = head - p; if (slot < 0) slot += a.RS;
                    const float2* x = Xk + ((size_t)slot * a.nIn + ni) * SC_BK;
                    prefetch_l1(x); prefetch_l1(x + 16);
                }
            }
        }
    }
    /* ---- A: forward FFTs of the inputs this CTA owns ---- */
    if (R > 16 && niA < a.nIn) load_block(niA);
    while (niA < a.nIn) {                                       /* warp-uniform */
        wfft<R, false>(v, a.wT1, lane, Lf);
        float2 Xs[R];
        wfft_fwd_split<R>(v, Xs, a.wT2, lane, 0.5f);
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int k = wf_bitrev(i, LOGR) + R * k1;
            a.X[(((size_t)(k >> 5) * a.RS + head) * a.nIn + niA) * SC_BK + (k & 31)] = Xs[i];
        }
        niA += C * SC_CL_WARPS;
        if (niA < a.nIn) load_block(niA);
    }
    cl_stamp(a.stamps, 1);
    cl_sync();
    cl_stamp(a.stamps, 2);
    if (c == 0 && threadIdx.x == 0) a.counters[0] = count + 1u;     /* every CTA has read it before the barrier */

    /* ---- B: Z[no][k] = sum_p sum_ni H_p[no][ni][k] * X_{t-p}[ni][k] ---- */
    {
        const int nUnits = a.nOut * a.nKT;                       /* unit u = kt * nOut + no */
        const int nuLocal = (nUnits - c + C - 1) / C;
        const int perRound = nuLocal < SC_CL_WARPS ? (nuLocal > 0 ? nuLocal : 1) : SC_CL_WARPS;
        const int wpu = SC_CL_WARPS / perRound;                  /* warps per unit */
        const int nTerms = a.P * a.nIn;
        for (int base = 0; base < nuLocal; base += perRound) {
            const int lu = base + warp / wpu, g = warp % wpu;
            const bool valid = (warp / wpu) < perRound && lu < nuLocal;
            int no = 0, kt = 0;
            if (valid) {
                const int u = c + C * lu;
                kt = u / a.nOut; no = u - kt * a.nOut;
                const int ot = no / a.OTsz, nl = no - ot * a.OTsz;
                const bool packed = (kt == 0 && lane == 0);
                const int t0 = (int)(((long long)nTerms * g) / wpu), t1 = (int)(((long long)nTerms * (g + 1)) / wpu);
                const float2* Hk = a.H + ((size_t)(ot * a.nKT + kt) * a.P * a.nIn * a.OTsz + nl) * SC_BK + lane;
                const float2* Xk = a.X + (size_t)kt * a.RS * a.nIn * SC_BK + lane;
                float2 acc = make_float2(0.f, 0.f);
                /* eight terms per trip, all sixteen loads issued before the first use (a warp typically owns fewer than eight
                 * terms: a plain loop would walk them one L2 round trip at a time); same summation order as a plain loop */
                int p = t0 / a.nIn, ni = t0 - p * a.nIn;             /* walked incrementally: no division per term */
                int slot = head - p; if (slot < 0) slot += a.RS;
                const unsigned hStep = (unsigned)a.OTsz * SC_BK;
                for (int tb = t0; tb < t1; tb += 8) {
                    float2 hv[8], xv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int t = tb + u;
                        hv[u] = make_float2(0.f, 0.f); xv[u] = hv[u];
                        if (t < t1) {
                            hv[u] = __ldg(Hk + (unsigned)t * hStep);
                            const float2* xp = Xk + (unsigned)(slot * a.nIn + ni) * SC_BK;
                            xv[u] = (RES || p == 0) ? __ldcg(xp)  /* written by other CTAs of this launch: L2 */
                                                    : __ldg(xp);  /* older blocks: prefetched into L1 above */
                            if (++ni == a.nIn) { ni = 0; ++p; slot = (slot == 0) ? a.RS - 1 : slot - 1; }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (tb + u < t1) cmac_packed(acc, hv[u], xv[u], packed);
                }
                red[warp * 32 + lane] = acc;
            }
            __syncthreads();
            if (valid && g == 0) {
                float2 z = red[warp * 32 + lane];
                for (int gg = 1; gg < wpu; ++gg) z = caddf(z, red[(warp + gg) * 32 + lane]);
                cl_store_f2(Zs + (size_t)(no / C) * ZS + kt * 32 + lane, (uint32_t)(no % C), z);
            }
            __syncthreads();
        }
    }
    cl_stamp(a.stamps, 3);
    cl_sync();
    cl_stamp(a.stamps, 4);

    /* ---- C: inverse FFT + overlap-add of the outputs this CTA owns (reference .c:230-233) ---- */
    if (warp < nOwn) {
        float2* z = Zs + (size_t)warp * ZS;
        float2 v[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int k = lane + 32 * i;
            const float2 A = z[k], B = z[(M - k) & (M - 1)];
            const float2 E = make_float2(A.x + B.x, A.y - B.y);
            const float2 D = make_float2(A.x - B.x, A.y + B.y);
            const float2 O = cmul_conjb(D, __ldg(a.tw + k));
            v[i] = make_float2(E.x - O.y, E.y + O.x);
            if (i == 0 && lane == 0) v[i] = make_float2(A.x + A.y, A.x - A.y);      /* (DC, Nyquist) */
        }
        wfft<R, true>(v, a.wT1, lane, Li);
        __syncwarp();                                           /* the spectrum has been read by every lane */
#pragma unroll
        for (int i = 0; i < R; ++i)                             /* z[n], n = n2 + R k1, stored at n2 + (R + 1) k1 */
            z[wf_bitrev(i, LOGR) + (R + 1) * k1] = make_float2(v[i].x * a.scale, v[i].y * a.scale);
        __syncwarp();
        const float* zf = reinterpret_cast<const float*>(z);
        float* out = a.out + (size_t)noMine * a.hop;
        float* tail = a.tail + (size_t)noMine * a.hop;
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const int i = lane + 32 * u;
            if (i < a.hop) {
                const int s1 = i + a.hop;
                const int n0 = i >> 1, n1 = s1 >> 1;
                out[i]  = zf[2 * ((n0 & (R - 1)) + (R + 1) * (n0 >> LOGR)) + (i & 1)] + (TLPRE ? tl[TLPRE ? u : 0] : __ldcg(tail + i));
                tail[i] = zf[2 * ((n1 & (R - 1)) + (R + 1) * (n1 >> LOGR)) + (s1 & 1)];
            }
        }
        if (warp == 0) cl_stamp(a.stamps, 5);
        if (a.done) {
            /* host-visible completion: every owner warp fences its output stores at system scope and takes a ticket; the last
             * one writes the sequence number into the host word the caller is polling (saves the end-of-kernel
             * signalling + cudaStreamSynchronize wake-up of a few microseconds per block) */
            __threadfence_system();
            __syncwarp();
            if (lane == 0) {
                const unsigned int t = atomicAdd(&a.counters[2], 1u);
                if (t == (unsigned)a.nOut - 1u) {
                    a.counters[2] = 0;
                    __threadfence_system();
                    *a.done = a.seq;
                }
            }
        }
    }
}

template <int R>
__global__ void __launch_bounds__(SC_CL_THREADS, 1) small_cluster_kernel(SmallCArgs a)
{
    extern __shared__ __align__(16) float2 smc[];
    small_cluster_block<R, false>(a, smc);
}

/* RESIDENT version (option "resident_us"): the cluster stays on its SMs and serves one block per DOORBELL -- the host writes
 * the block's buffer addresses and a sequence number into a page-locked mailbox, CTA 0 polls it at system scope, the cluster
 * runs the block and writes the sequence number into the completion word.  A call then costs two PCIe hops and the block
 * itself; the launch + completion path of a kernel (9.2 us for an EMPTY kernel on this pool, tools/lat_floor.cu) is gone.
 * The kernel leaves on its own after `idleNs` without a doorbell (alive = 0: the host starts a new one with the next block),
 * or when the host rings SC_RES_EXIT (any other call on the handle, destroy). */
#define SC_RES_EXIT 0xFFFFFFFFu
struct ScMailbox {                       /* page-locked host memory; layout shared with safconv_host.c (sc_mailbox) */
    volatile unsigned int done;          /* completion word: sequence number of the last finished block */
    unsigned int pad0[15];
    volatile unsigned long long bell;    /* doorbell, ONE 8-byte word: low half = sequence number, high half = (generation << 8) | slot */
    volatile unsigned int alive;         /* 1 while a resident kernel is polling */
    unsigned int pad1[13];
    volatile unsigned long long buf[16]; /* table of (in, out) buffer addresses, 8 slots; `generation` changes whenever the host rewrites it */
};

template <int R>
__global__ void __launch_bounds__(SC_CL_THREADS, 1) small_cluster_resident_kernel(SmallCArgs a, ScMailbox* mb, unsigned int lastSeq,
                                                                                  unsigned long long idleNs)
{
    extern __shared__ __align__(16) float2 smc[];
    __shared__ unsigned long long s_cmd[3];                    /* CTA 0: seq, in, out of the next block */
    __shared__ unsigned long long s_tab[16];                   /* CTA 0, thread 0: its copy of the buffer table */
    const int c = (int)cl_ctarank();
    unsigned int tabGen = 0xFFFFFFFFu;                         /* generation of s_tab (thread 0 of CTA 0) */
    for (;;) {
        if (c == 0 && threadIdx.x == 0) {
            unsigned long long t0, t, bell;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            unsigned int seq;
            for (;;) {
                bell = ld_sys_u64(&mb->bell);                  /* one PCIe read: sequence number, slot and table generation together */
                seq = (unsigned int)bell;
                if (seq != lastSeq) break;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t - t0 > idleNs) { seq = SC_RES_EXIT; break; }
            }
            cl_stamp(a.stamps, 6);
            s_cmd[0] = seq;
            if (seq != SC_RES_EXIT) {
                const unsigned int hi = (unsigned int)(bell >> 32), slot = hi & 7u, gen = hi >> 8;
                if (gen != tabGen) {                           /* the host rewrote the table (new caller buffers): one more round trip, 16 loads in flight */
                    unsigned long long v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = ld_sys_u64(&mb->buf[i]);
#pragma unroll
                    for (int i = 0; i < 16; ++i) s_tab[i] = v[i];
                    tabGen = gen;
                }
                s_cmd[1] = s_tab[2 * slot]; s_cmd[2] = s_tab[2 * slot + 1];
            }
        }
        cl_sync();                                              /* CTA 0's command is visible to the whole cluster */
        unsigned long long cmd[3];
        {
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(s_cmd)), "r"(0));
#pragma unroll
            for (int i = 0; i < 3; ++i)
                asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(cmd[i]) : "r"(remote + 8u * i) : "memory");
        }
        const unsigned int seq = (unsigned int)cmd[0];
        if (seq == SC_RES_EXIT) {
            cl_sync();                                          /* nobody reads CTA 0's shared memory after it has left */
            if (c == 0 && threadIdx.x == 0) { __threadfence_system(); mb->alive = 0u; }
            return;
        }
        SmallCArgs b = a;
        b.in = reinterpret_cast<const float*>(cmd[1]); b.out = reinterpret_cast<float*>(cmd[2]);
        b.done = &mb->done; b.seq = seq;
        small_cluster_block<R, true>(b, smc);
        lastSeq = seq;
    }
}

/* the cluster version: M = 64 .. 1024, at most 16 inputs / outputs per CTA of the cluster, filters that stay in L2 */
static int small_cluster_plan_ok(const scdev_plan* pl)
{
    static int env = -1;
    if (env < 0) { const char* v = getenv("SAFCONV_SMALL_CLUSTER"); env = v ? atoi(v) : 1; }
    if (!env || pl->kind != SC_KIND_MATRIX) return 0;
    if (pl->M < 64 || pl->M > 1024 || pl->hop > pl->M) return 0;
    if (pl->nIn > SC_CL_MAXC * SC_CL_WARPS || pl->nOutLocal > SC_CL_MAXC * SC_CL_WARPS) return 0;
    if ((double)pl->P * pl->nIn * pl->nOutLocal * pl->M * 8.0 > 4.0 * 1024 * 1024) return 0;     /* one cluster = 8 SMs: more than this belongs on the whole GPU */
    return 1;
}
static int small_cluster_ok(const scdev_plan* pl, const scdev_bufs* b) { return b->wtab && small_cluster_plan_ok(pl); }

template <int R>
static int small_cluster_launch(const SmallCArgs& a, int C, cudaStream_t st)
{
    const int nOwnMax = (a.nOut + C - 1) / C;
    const size_t smem = ((size_t)nOwnMax * (32 * R + 33) + SC_CL_WARPS * 32) * sizeof(float2);
    if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(small_cluster_kernel<R>));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)C); cfg.blockDim = dim3(SC_CL_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, small_cluster_kernel<R>, a);
}

static void small_cluster_fill(SmallCArgs& a, const scdev_plan* pl, const scdev_bufs* b, int* C)
{
    a.in = NULL; a.out = NULL; a.done = NULL; a.seq = 0; a.stamps = NULL;
    a.H = (const float2*)b->H; a.X = (float2*)b->X; a.tw = (const float2*)b->tw;
    a.wT1 = (const float2*)b->wtab; a.wT2 = (const float2*)b->wtab + pl->M;
    a.tail = b->tail; a.counters = b->counters;
    a.hop = pl->hop; a.P = pl->P; a.nIn = pl->nIn; a.nOut = pl->nOutLocal; a.nKT = pl->nKT; a.OTsz = pl->OTsz; a.RS = pl->RS;
    a.scale = 1.0f / (float)pl->N;
    int work = pl->nOutLocal * pl->nKT;
    if (pl->nIn > work) work = pl->nIn;
    *C = work < SC_CL_MAXC ? work : SC_CL_MAXC;
}

static unsigned long long* g_resStamps = NULL;
template <int R>
static int small_resident_launch(const SmallCArgs& a, int C, ScMailbox* mb, unsigned int lastSeq, unsigned long long idleNs, cudaStream_t st)
{
    const int nOwnMax = (a.nOut + C - 1) / C;
    const size_t smem = ((size_t)nOwnMax * (32 * R + 33) + SC_CL_WARPS * 32) * sizeof(float2);
    if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(small_cluster_resident_kernel<R>));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)C); cfg.blockDim = dim3(SC_CL_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, small_cluster_resident_kernel<R>, a, mb, lastSeq, idleNs);
}

static int small_resident_start(const scdev_plan* pl, const scdev_bufs* b, void* mailbox, unsigned int lastSeq, unsigned int idleUs,
                                cudaStream_t st)
{
    if (!small_cluster_ok(pl, b)) return (int)cudaErrorNotSupported;
    SmallCArgs a; int C;
    small_cluster_fill(a, pl, b, &C);
    {   /* debugging (SAFCONV_KSTAMPS=1): %globaltimer of CTA 0 -- [6] doorbell seen, [0..5] phase boundaries of the block; read with scdev_small_resident_stamps */
        const char* v = getenv("SAFCONV_KSTAMPS");
        if (v && atoi(v)) { if (!g_resStamps) cudaMalloc(&g_resStamps, 64); a.stamps = g_resStamps; }
    }
    ScMailbox* mb = (ScMailbox*)mailbox;
    const unsigned long long idleNs = (unsigned long long)idleUs * 1000ull;
    switch (pl->M) {
        case 64:   return small_resident_launch<2>(a, C, mb, lastSeq, idleNs, st);
        case 128:  return small_resident_launch<4>(a, C, mb, lastSeq, idleNs, st);
        case 256:  return small_resident_launch<8>(a, C, mb, lastSeq, idleNs, st);
        case 512:  return small_resident_launch<16>(a, C, mb, lastSeq, idleNs, st);
        case 1024: return small_resident_launch<32>(a, C, mb, lastSeq, idleNs, st);
        default:   return (int)cudaErrorInvalidValue;
    }
}

static int scdev_small_cluster(const scdev_plan* pl, const scdev_bufs* b, const float* in, float* out, cudaStream_t st,
                               volatile unsigned int* done, unsigned int seq)
{
    SmallCArgs a;
    a.done = done; a.seq = seq; a.stamps = NULL;
    static int ks = -1;
    static unsigned long long* d_st = NULL;
    static double acc[5]; static int nAcc = 0;
    if (ks < 0) { const char* v = getenv("SAFCONV_KSTAMPS"); ks = v ? atoi(v) : 0; if (ks && cudaMalloc(&d_st, 64) != cudaSuccess) ks = 0; }
    if (ks) a.stamps = d_st;
    a.in = in; a.out = out; a.H = (const float2*)b->H; a.X = (float2*)b->X; a.tw = (const float2*)b->tw;
    a.wT1 = (const float2*)b->wtab; a.wT2 = (const float2*)b->wtab + pl->M;
    a.tail = b->tail; a.counters = b->counters;
    a.hop = pl->hop; a.P = pl->P; a.nIn = pl->nIn; a.nOut = pl->nOutLocal; a.nKT = pl->nKT; a.OTsz = pl->OTsz; a.RS = pl->RS;
    a.scale = 1.0f / (float)pl->N;
    int work = pl->nOutLocal * pl->nKT;
    if (pl->nIn > work) work = pl->nIn;
    const int C = work < SC_CL_MAXC ? work : SC_CL_MAXC;
    int e;
    switch (pl->M) {
        case 64:   e = small_cluster_launch<2>(a, C, st); break;
        case 128:  e = small_cluster_launch<4>(a, C, st); break;
        case 256:  e = small_cluster_launch<8>(a, C, st); break;
        case 512:  e = small_cluster_launch<16>(a, C, st); break;
        case 1024: e = small_cluster_launch<32>(a, C, st); break;
        default:   return (int)cudaErrorInvalidValue;
    }
    if (ks && !e) {                                             /* debugging only: phase durations of CTA 0, averaged */
        unsigned long long h[6];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_st, sizeof h, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 5; ++i) acc[i] += (double)(h[i + 1] - h[i]);
        if (++nAcc == 1000) {
            fprintf(stderr, "small_cluster phases (us): A %.2f, barrier %.2f, B %.2f, barrier %.2f, C %.2f\n",
                    acc[0] * 1e-6, acc[1] * 1e-6, acc[2] * 1e-6, acc[3] * 1e-6, acc[4] * 1e-6);
            nAcc = 0; for (int i = 0; i < 5; ++i) acc[i] = 0;
        }
    }
    return e;
}

/* ------------------------------------------------------------------------------------------ */
/*  TVConv: one CTA per output channel (reference .c:546-620)                                   */
/*  shared memory: 5*M float2  (X spectrum, three output frames, twiddles)                       */
/* ------------------------------------------------------------------------------------------ */
struct TvArgs {
    const float* in;       /* [hop] */
    float* out;            /* [nOut][hop] */
    const float2* H;       /* [nIRs][nOut][P][M] */
    float2* X;             /* [P][M] ring */
    const float2* tw;
    float* tail0;          /* y_n_overlap      [nOut][hop] */
    float* tail1;          /* y_n_overlap_last [nOut][hop] */
    unsigned int* counters;
    int hop, M, logM, P, nOut;
    int ir0, ir1, ir2;     /* irIdx, posIdx_last, posIdx_last2 */
    float scale;
};

__global__ void tv_fused_kernel(TvArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    const int MP = SC_ALEN(a.M);     /* padded FFT work arrays Z0, Z1, Z2 stored back to back */
    float2* Z0 = sm;
    float2* Z1 = sm + MP;
    float2* Z2 = sm + 2 * MP;
    float2* Xs = sm + 3 * MP;        /* packed spectrum of the new block, natural order */
    float2* stw = Xs + a.M;
    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    const int no = blockIdx.x;
    const int head = (int)(a.counters[0] % (unsigned)a.P);
    const bool need1 = (a.ir0 != a.ir1);
    const bool need2 = (a.ir1 != a.ir2);

    /* every CTA transforms the (single) input block; CTA 0 also stores it in the ring */
    load_real_block(Z0, a.in, a.hop, a.M, a.logM);
    __syncthreads();
    cfft_dif<false>(Z0, a.M, a.logM, stw, wide);
    for (int k = threadIdx.x; k <= (a.M >> 1); k += blockDim.x) {
        float2 Xk, Xmk;
        int k2 = a.M - k;
        if (k == 0) {
            const float2 z = Z0[0];
            Xk = make_float2(z.x + z.y, z.x - z.y);
            k2 = 0; Xmk = Xk;
        } else {
            fwd_split_pair(Z0, k, a.M, a.logM, spl, Xk, Xmk);
        }
        Xs[k] = Xk;  Xs[k2] = Xmk;
        if (no == 0) {
            a.X[(size_t)head * a.M + k] = Xk;
            a.X[(size_t)head * a.M + k2] = Xmk;
        }
    }
    __syncthreads();
    const size_t pm = (size_t)a.P * a.M;
    const float2* H0 = a.H + ((size_t)a.ir0 * a.nOut + no) * pm;
    const float2* H1 = a.H + ((size_t)a.ir1 * a.nOut + no) * pm;
    const float2* H2 = a.H + ((size_t)a.ir2 * a.nOut + no) * pm;
    for (int k = threadIdx.x; k < a.M; k += blockDim.x) {
        const bool packed = (k == 0);
        float2 z0 = make_float2(0.f, 0.f), z1 = z0, z2 = z0;
        int slot = head;
#pragma unroll 8
        for (int p = 0; p < a.P; ++p) {
            const float2 x = (p == 0) ? Xs[k] : a.X[(size_t)slot * a.M + k];
            cmac_packed(z0, __ldg(H0 + (size_t)p * a.M + k), x, packed);
            if (need1) cmac_packed(z1, __ldg(H1 + (size_t)p * a.M + k), x, packed);
            if (need2) cmac_packed(z2, __ldg(H2 + (size_t)p * a.M + k), x, packed);
            slot = (slot == 0) ? a.P - 1 : slot - 1;
        }
        if (!need1) z1 = z0;                  /* .c:587 */
        if (!need2) z2 = z1;                  /* .c:601 */
        const int ik = padi(k, a.logM);
        Z0[ik] = z0; Z1[ik] = z1; Z2[ik] = z2;
    }
    __syncthreads();
    inv_split_batch(Z0, a.M, a.logM, spl, 3);           /* Z0, Z1, Z2 are contiguous */
    cfft_dif_batch<true>(Z0, a.M, a.logM, stw, 3, wide);
    /* cross-fade (reference .c:494-497, 605-615) */
    float* t0 = a.tail0 + (size_t)no * a.hop;
    float* t1 = a.tail1 + (size_t)no * a.hop;
    float* o  = a.out + (size_t)no * a.hop;
    const float den = (float)(a.hop - 1);
    for (int i = threadIdx.x; i < a.hop; i += blockDim.x) {
        const float fin  = (float)i / den;
        const float fout = (float)(a.hop - 1 - i) / den;
        const float o1 = time_sample(Z1, i, a.logM) * a.scale + t0[i];
        const float o2 = time_sample(Z2, i, a.logM) * a.scale + t1[i];
        o[i]  = o1 * fin + o2 * fout;
        t0[i] = time_sample(Z0, i + a.hop, a.logM) * a.scale;
        t1[i] = time_sample(Z1, i + a.hop, a.logM) * a.scale;
    }
    advance_block_counter(a.counters, gridDim.x);
}


/* ---- TVConv, hop > 4096 (M = 8192): the five work arrays of tv_fused_kernel no longer fit in shared memory, so the
 * block runs as three launches through a global scratch zt[3][nOut][2*hop] (a block is >= 85 ms of audio here: launch
 * gaps are irrelevant).  Same arithmetic per bin / per sample as the fused kernel. ---- */

/* grid (1): forward FFT of the input block into ring slot `head` */
__global__ void tv_input_kernel(TvArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    float2* stw = sm + SC_ALEN(a.M);
    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    const int head = (int)(a.counters[0] % (unsigned)a.P);
    load_real_block(sm, a.in, a.hop, a.M, a.logM);
    __syncthreads();
    cfft_dif<false>(sm, a.M, a.logM, stw, wide);
    float2* Xn = a.X + (size_t)head * a.M;
    for (int k = threadIdx.x; k <= (a.M >> 1); k += blockDim.x) {
        float2 Xk, Xmk;
        int k2 = a.M - k;
        if (k == 0) {
            const float2 z = sm[0];
            Xk = make_float2(z.x + z.y, z.x - z.y);
            k2 = 0; Xmk = Xk;
        } else {
            fwd_split_pair(sm, k, a.M, a.logM, spl, Xk, Xmk);
        }
        Xn[k] = Xk;  Xn[k2] = Xmk;
    }
}

/* grid (nOut, 3): which = 0 / 1 / 2 -> IR set irIdx / posIdx_last / posIdx_last2; sets equal to the previous one are
 * skipped (reference .c:587, .c:601: the consumer re-uses the earlier result) */
__global__ void tv_mac_ifft_kernel(TvArgs a, float* zt)
{
    extern __shared__ __align__(16) float2 sm[];
    const int no = blockIdx.x, which = blockIdx.y;
    if ((which == 1 && a.ir0 == a.ir1) || (which == 2 && a.ir1 == a.ir2)) return;
    float2* stw = sm + SC_ALEN(a.M);
    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    const int ir = (which == 0) ? a.ir0 : (which == 1 ? a.ir1 : a.ir2);
    const int head = (int)(a.counters[0] % (unsigned)a.P);
    const float2* Hc = a.H + ((size_t)ir * a.nOut + no) * a.P * a.M;
    for (int k = threadIdx.x; k < a.M; k += blockDim.x) {
        const bool packed = (k == 0);
        float2 z = make_float2(0.f, 0.f);
        int slot = head;
#pragma unroll 8
        for (int p = 0; p < a.P; ++p) {
            cmac_packed(z, __ldg(Hc + (size_t)p * a.M + k), a.X[(size_t)slot * a.M + k], packed);
            slot = (slot == 0) ? a.P - 1 : slot - 1;
        }
        sm[padi(k, a.logM)] = z;
    }
    __syncthreads();
    inv_split_all(sm, a.M, a.logM, spl);
    cfft_dif<true>(sm, a.M, a.logM, stw, wide);
    float* z = zt + ((size_t)which * a.nOut + no) * 2 * a.hop;
    for (int i = threadIdx.x; i < 2 * a.hop; i += blockDim.x) z[i] = time_sample(sm, i, a.logM) * a.scale;
}

/* cross-fade + tails (reference .c:494-497, 605-615), one thread per (output, sample); last CTA bumps the counter */
__global__ void tv_xfade_kernel(TvArgs a, const float* zt)
{
    const size_t n = (size_t)a.nOut * a.hop;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) {
        const int no = (int)(idx / a.hop), i = (int)(idx - (size_t)no * a.hop);
        const int s1 = (a.ir0 != a.ir1) ? 1 : 0;
        const int s2 = (a.ir1 != a.ir2) ? 2 : s1;
        const float* z0 = zt + ((size_t)0 * a.nOut + no) * 2 * a.hop;
        const float* z1 = zt + ((size_t)s1 * a.nOut + no) * 2 * a.hop;
        const float* z2 = zt + ((size_t)s2 * a.nOut + no) * 2 * a.hop;
        const float den = (float)(a.hop - 1);
        const float fin  = (float)i / den;
        const float fout = (float)(a.hop - 1 - i) / den;
        const float o1 = z1[i] + a.tail0[idx];
        const float o2 = z2[i] + a.tail1[idx];
        a.out[idx]   = o1 * fin + o2 * fout;
        a.tail0[idx] = z0[i + a.hop];
        a.tail1[idx] = z1[i + a.hop];
    }
    advance_block_counter(a.counters, gridDim.x);
}


/* ------------------------------------------------------------------------------------------ */
/*  Stand-alone real FFT pair with the reference's saf_rfft conventions (saf_utility_fft.c:531-753: */
/*  N/2+1 interleaved complex bins, forward unscaled, backward x 1/N, Im of DC and Nyquist ignored   */
/*  by the backward transform -- kiss_fftr.c:137-138).  Same device code as the convolver kernels     */
/*  (load_real_block, cfft_dif, split passes); one CTA per transform, grid = batch.                  */
/* ------------------------------------------------------------------------------------------ */
struct RfftArgs {
    const float* in;       /* forward: [batch][N] real;            backward: [batch][N/2+1] complex */
    float* out;            /* forward: [batch][N/2+1] complex;     backward: [batch][N] real        */
    const float2* tw;
    int N, M, logM;
};

__global__ void rfft_forward_kernel(RfftArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    float2* stw = sm + SC_ALEN(a.M);
    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    load_real_block(sm, a.in + (size_t)blockIdx.x * a.N, a.N, a.M, a.logM);
    __syncthreads();
    cfft_dif<false>(sm, a.M, a.logM, stw, wide);
    float2* X = reinterpret_cast<float2*>(a.out) + (size_t)blockIdx.x * (a.M + 1);
    for (int k = threadIdx.x; k <= (a.M >> 1); k += blockDim.x) {
        if (k == 0) {
            const float2 z = sm[0];
            X[0]   = make_float2(z.x + z.y, 0.f);
            X[a.M] = make_float2(z.x - z.y, 0.f);
        } else {
            float2 Xk, Xmk;
            fwd_split_pair(sm, k, a.M, a.logM, spl, Xk, Xmk);
            X[k] = Xk;  X[a.M - k] = Xmk;
        }
    }
}

__global__ void rfft_backward_kernel(RfftArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    float2* stw = sm + SC_ALEN(a.M);
    const bool wide = fft_use_wide(a.M, 1);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = load_split_twiddles(stw, a.tw, a.M);
    const float2* X = reinterpret_cast<const float2*>(a.in) + (size_t)blockIdx.x * (a.M + 1);
    for (int k = threadIdx.x; k < a.M; k += blockDim.x) {
        float2 v = __ldg(X + k);
        if (k == 0) v = make_float2(v.x, __ldg(X + a.M).x);     /* packed (DC, Nyquist): real parts only */
        sm[padi(k, a.logM)] = v;
    }
    __syncthreads();
    inv_split_all(sm, a.M, a.logM, spl);
    cfft_dif<true>(sm, a.M, a.logM, stw, wide);
    float* x = a.out + (size_t)blockIdx.x * a.N;
    const float scale = 1.0f / (float)a.N;
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) x[i] = time_sample(sm, i, a.logM) * scale;
}

/* ------------------------------------------------------------------------------------------ */
/*  C-ABI: plumbing                                                                             */
/* ------------------------------------------------------------------------------------------ */
extern "C" {

int scdev_device_count(int* n) { return (int)cudaGetDeviceCount(n); }
int scdev_set_device(int dev) { return (int)cudaSetDevice(dev); }
int scdev_get_device(int* dev) { return (int)cudaGetDevice(dev); }
int scdev_device_props(int dev, int* smCount, int* maxSmemOptin, int* ccMajor, int* ccMinor)
{
    SC_CHECK(cudaDeviceGetAttribute(smCount, cudaDevAttrMultiProcessorCount, dev));
    SC_CHECK(cudaDeviceGetAttribute(maxSmemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SC_CHECK(cudaDeviceGetAttribute(ccMajor, cudaDevAttrComputeCapabilityMajor, dev));
    SC_CHECK(cudaDeviceGetAttribute(ccMinor, cudaDevAttrComputeCapabilityMinor, dev));
    return 0;
}
int scdev_malloc(void** p, size_t bytes) { return (int)cudaMalloc(p, bytes ? bytes : 16); }
int scdev_free(void* p) { return p ? (int)cudaFree(p) : 0; }
int scdev_host_alloc(void** p, size_t bytes) { return (int)cudaHostAlloc(p, bytes ? bytes : 16, cudaHostAllocDefault); }
int scdev_host_free(void* p) { return p ? (int)cudaFreeHost(p) : 0; }
int scdev_memset_async(void* p, int v, size_t bytes, void* stream) { return (int)cudaMemsetAsync(p, v, bytes, (cudaStream_t)stream); }
int scdev_memcpy_h2d_async(void* d, const void* h, size_t bytes, void* stream)
{ return (int)cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream); }
int scdev_memcpy_d2h_async(void* h, const void* d, size_t bytes, void* stream)
{ return (int)cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream); }
/* strided (pitched) copies: `height` rows of `widthBytes`; toHost = 0: host -> device, 1: device -> host */
int scdev_memcpy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t widthBytes, size_t height,
                         int toHost, void* stream)
{
    return (int)cudaMemcpy2DAsync(dst, dpitch, src, spitch, widthBytes, height,
                                  toHost ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, (cudaStream_t)stream);
}
/* Stream-ordered upload that has fully landed on return.  (A plain cudaMemcpy from pageable memory may
 * return while the last staged chunk is still in flight and is only ordered against the legacy default
 * stream -- not against the handle's non-blocking stream that the create-time kernels run on.) */
int scdev_memcpy_h2d_sync(void* d, const void* h, size_t bytes, void* stream)
{
    SC_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return (int)cudaStreamSynchronize((cudaStream_t)stream);
}
int scdev_stream_create(void** s)
{ cudaStream_t st; cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking); *s = (void*)st; return (int)e; }
int scdev_stream_create_high_priority(void** s)
{
    int lo = 0, hi = 0;
    SC_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));        /* hi = numerically lowest = greatest priority */
    cudaStream_t st; cudaError_t e = cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, hi); *s = (void*)st; return (int)e;
}
int scdev_stream_destroy(void* s) { return s ? (int)cudaStreamDestroy((cudaStream_t)s) : 0; }
int scdev_stream_sync(void* s) { return (int)cudaStreamSynchronize((cudaStream_t)s); }
int scdev_event_create(void** e) { cudaEvent_t ev; cudaError_t r = cudaEventCreate(&ev); *e = (void*)ev; return (int)r; }
int scdev_event_create_sync(void** e)
{ cudaEvent_t ev; cudaError_t r = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); *e = (void*)ev; return (int)r; }
int scdev_stream_wait_event(void* stream, void* e) { return (int)cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)e, 0); }
/* 1: the event's work has completed (or it was never recorded), 0: still pending, < 0: error */
int scdev_event_done(void* e)
{
    const cudaError_t r = cudaEventQuery((cudaEvent_t)e);
    if (r == cudaSuccess) return 1;
    if (r == cudaErrorNotReady) { (void)cudaGetLastError(); return 0; }
    return -1;
}
int scdev_event_destroy(void* e) { return e ? (int)cudaEventDestroy((cudaEvent_t)e) : 0; }
int scdev_event_record(void* e, void* stream) { return (int)cudaEventRecord((cudaEvent_t)e, (cudaStream_t)stream); }
int scdev_event_sync(void* e) { return (int)cudaEventSynchronize((cudaEvent_t)e); }
int scdev_last_error_clear(void) { return (int)cudaGetLastError(); }
int scdev_event_elapsed_ms(void* e0, void* e1, float* ms) { return (int)cudaEventElapsedTime(ms, (cudaEvent_t)e0, (cudaEvent_t)e1); }
int scdev_graph_begin(void* stream) { return (int)cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal); }
int scdev_graph_end(void* stream, void** graphExec)
{
    cudaGraph_t g = nullptr;
    SC_CHECK(cudaStreamEndCapture((cudaStream_t)stream, &g));
    cudaGraphExec_t ge = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    *graphExec = (void*)ge;
    return (int)e;
}
int scdev_graph_launch(void* graphExec, void* stream) { return (int)cudaGraphLaunch((cudaGraphExec_t)graphExec, (cudaStream_t)stream); }
int scdev_graph_destroy(void* graphExec) { return graphExec ? (int)cudaGraphExecDestroy((cudaGraphExec_t)graphExec) : 0; }
const char* scdev_error_string(int err) { return cudaGetErrorString((cudaError_t)err); }

/* ------------------------------------------------------------------------------------------ */
/*  C-ABI: kernel launchers                                                                     */
/* ------------------------------------------------------------------------------------------ */

static size_t fft_smem(const scdev_plan* pl, int nbuf)
{
    return ((size_t)nbuf * SC_ALEN(pl->M) + sc_split_len(pl->M)) * sizeof(float2);
}

/* forward FFTs per CTA of the batched multiConv path: 4 while the work arrays stay small enough for >= 3 CTAs per SM */
static int multi_fft_q(const scdev_plan* pl) { return fft_smem(pl, SC_MULTI_FFT_Q + 1) <= 64 * 1024 ? SC_MULTI_FFT_Q : 1; }

typedef void (*mac_fn_t)(MacArgs);
static mac_fn_t mac_fn(int R)
{
    switch (R) {
        case 1: return mac_kernel<1>;
        case 2: return mac_kernel<2>;
        case 3: return mac_kernel<3>;
        case 4: return mac_kernel<4>;
        case 5: return mac_kernel<5>;
        case 6: return mac_kernel<6>;
        case 7: return mac_kernel<7>;
        default: return mac_kernel<8>;
    }
}

int scdev_prepare(const scdev_plan* pl)
{
    /* FFT kernels hold the data and the twiddle table in shared memory: 2*M float2 (<= 128 KB).  The limit is raised to
     * the device maximum for every kernel this kind of handle can launch (see sc_optin_smem: never plan-specific). */
    SC_CHECK(sc_optin_smem(filter_fft_kernel));
    SC_CHECK(sc_optin_smem(input_fft_kernel));
    SC_CHECK(sc_optin_smem(ifft_ola_kernel));
    SC_CHECK(sc_optin_smem(ifft_batch_kernel));
    if (pl->kind == SC_KIND_MULTI) {
        SC_CHECK(sc_optin_smem(multi_fused_kernel));
        SC_CHECK(sc_optin_smem(multi_fft_batch_kernel));
        SC_CHECK(sc_optin_smem(multi_mac_ifft_batch_kernel));
        SC_CHECK(sc_optin_smem(multi_mac_ifft_batch_reg_kernel<2>));
        SC_CHECK(sc_optin_smem(multi_mac_ifft_batch_reg_kernel<4>));
        SC_CHECK(sc_optin_smem(multi_mac_ifft_batch_reg_kernel<8>));
        SC_CHECK(sc_optin_smem(multi_mac_ifft_batch_reg_kernel<16>));
    }
    if (pl->kind == SC_KIND_TV) SC_CHECK(sc_optin_smem(tv_fused_kernel));
    if (pl->kind == SC_KIND_MATRIX) {
        SC_CHECK(sc_optin_smem(mac_fn(pl->R)));
        /* The look-ahead apply runs K1 (next block) and K3 (this block) BESIDE a resident tail pass of the MAC.  CTAs
         * of kernels whose shared-memory carve-outs differ cannot share an SM (measured: with the driver's per-kernel
         * choice K1 / K3 waited for the whole 0.43 ms tail pass), so all three ask for the same, largest one. */
        SC_CHECK(cudaFuncSetAttribute(mac_fn(pl->R), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SC_CHECK(cudaFuncSetAttribute(input_fft_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        SC_CHECK(cudaFuncSetAttribute(ifft_ola_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    return 0;
}

int scdev_filter_transform(const scdev_plan* pl, const scdev_bufs* b, const float* d_h, void* stream)
{
    FilterArgs a;
    a.h = d_h; a.H = (float2*)b->H; a.tw = (const float2*)b->tw;
    a.kind = pl->kind; a.hop = pl->hop; a.len = pl->len; a.nIn = pl->nIn;
    a.M = pl->M; a.logM = pl->logM; a.P = pl->P; a.nKT = pl->nKT; a.OTsz = pl->OTsz;
    const long long rows = (pl->kind == SC_KIND_TV) ? (long long)pl->nIRs * pl->nOutLocal : pl->nOutLocal;
    a.nInGrid = (pl->kind == SC_KIND_MATRIX) ? pl->nIn : 1;
    a.total = rows * a.nInGrid * pl->P;
    const long long gx = a.total < (1ll << 30) ? a.total : (1ll << 30);
    const long long gy = (a.total + gx - 1) / gx;
    if (gy > 65535) return (int)cudaErrorInvalidValue;
    dim3 grid((unsigned)gx, (unsigned)gy, 1);
    filter_fft_kernel<<<grid, pl->fftThreads, fft_smem(pl, 2), (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

int scdev_input_fft(const scdev_plan* pl, const scdev_bufs* b, const float* d_in, int nBlocks, void* stream)
{
    InFftArgs a;
    a.in = d_in; a.X = (float2*)b->X; a.tw = (const float2*)b->tw; a.counters = b->counters;
    a.hop = pl->hop; a.nIn = pl->nIn; a.M = pl->M; a.logM = pl->logM; a.RS = pl->RS;
    dim3 grid(pl->nIn, nBlocks);
    input_fft_kernel<<<grid, pl->fftThreads, fft_smem(pl, 2), (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

int scdev_mac(const scdev_plan* pl, const scdev_bufs* b, int blk, int nBlocks, void* stream)
{
    scdev_macpass full;
    full.pLo = 0; full.nP = pl->P; full.totalStages = pl->totalStages; full.grid = pl->macGrid; full.nSlots = pl->nSlots;
    full.ctaBase = b->ctaBase; full.grpStart = b->grpStart; full.Zp = b->Zp;
    full.ZpB = NULL; full.stages = 0;
    return scdev_mac_pass(pl, b, &full, blk, nBlocks, 0, -1, stream);
}

int scdev_mac_pass(const scdev_plan* pl, const scdev_bufs* b, const scdev_macpass* ps, int blk, int nBlocks,
                   int zpSel, long long count, void* stream)
{
    MacArgs a;
    a.H = (const float2*)b->H; a.X = (const float2*)b->X;
    a.zpStride = (size_t)ps->nSlots * pl->OTsz * SC_BK;
    a.Zp = (float2*)(zpSel ? ps->ZpB : ps->Zp) + (size_t)blk * a.zpStride;
    a.useCount = count >= 0; a.count = (unsigned int)(count >= 0 ? count : 0);
    a.nB = nBlocks;
    a.counters = b->counters; a.ctaBase = ps->ctaBase;
    a.totalStages = ps->totalStages; a.pLo = ps->pLo; a.nP = ps->nP;
    a.nIn = pl->nIn; a.OTsz = pl->OTsz; a.P = pl->P; a.nKT = pl->nKT;
    a.SNI = pl->SNI; a.SPU = pl->SPU; a.WGo = pl->WGo; a.WGk = pl->WGk; a.hints = pl->macHints;
    a.NS = pl->macStages; a.RS = pl->RS; a.blk = blk;
    a.stageHBytes = pl->SNI * pl->OTsz * SC_BK * 8;
    a.stageXBytes = pl->SNI * SC_BK * 8;
    int smem = pl->macSmemBytes;
    if (ps->stages > 0 && ps->stages < pl->macStages) {        /* shallower pipeline: fits on an SM beside a full-depth pass */
        a.NS = ps->stages;
        smem = a.NS * (a.stageHBytes + a.stageXBytes) + SC_MAC_CWARPS * pl->R * 32 * 8 + 2 * a.NS * 8;
    }
    mac_fn(pl->R)<<<ps->grid, SC_MAC_THREADS, smem, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

static void fill_ifft_args(IfftArgs& a, const scdev_plan* pl, const scdev_bufs* b, float* d_out, int nBlocks)
{
    a.Zp = (const float2*)b->Zp; a.grpStart = b->grpStart; a.Zp2 = NULL; a.grpStart2 = NULL;
    a.headH = NULL; a.headX = NULL; a.nHead = 0; a.P = pl->P; a.nIn = pl->nIn; a.RS = pl->RS;
    a.tw = (const float2*)b->tw; a.out = d_out; a.tail = b->tail; a.zt = b->zt; a.counters = b->counters;
    a.zpStride = (size_t)pl->nSlots * pl->OTsz * SC_BK;
    a.hop = pl->hop; a.M = pl->M; a.logM = pl->logM; a.nKT = pl->nKT; a.OTsz = pl->OTsz;
    a.nOutLocal = pl->nOutLocal; a.B = nBlocks;
    a.scale = 1.0f / (float)pl->N;
}

int scdev_ifft_ola(const scdev_plan* pl, const scdev_bufs* b, float* d_out, void* stream)
{
    return scdev_ifft_ola_passes(pl, b, NULL, 0, NULL, d_out, stream);
}

int scdev_ifft_ola_passes(const scdev_plan* pl, const scdev_bufs* b, const scdev_macpass* p1, int zpSel1, const scdev_macpass* p2,
                          float* d_out, void* stream)
{
    IfftArgs a;
    fill_ifft_args(a, pl, b, d_out, 1);
    if (p1) { a.Zp = (const float2*)(zpSel1 ? p1->ZpB : p1->Zp); a.grpStart = p1->grpStart; }
    if (p2) { a.Zp2 = (const float2*)p2->Zp; a.grpStart2 = p2->grpStart; }
    if (p1 && !p2 && p1->pLo >= 1) {                 /* tail pass only: K3 adds the newest partition(s) itself, with twice the threads in flight */
        a.headH = (const float2*)b->H; a.headX = (const float2*)b->X; a.nHead = p1->pLo;
        int threads = 2 * pl->fftThreads; if (threads > 512) threads = 512;
        ifft_ola_kernel<<<pl->nOutLocal, threads, fft_smem(pl, 2), (cudaStream_t)stream>>>(a);
        return (int)cudaGetLastError();
    }
    ifft_ola_kernel<<<pl->nOutLocal, pl->fftThreads, fft_smem(pl, 2), (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

int scdev_ifft_ola_batch(const scdev_plan* pl, const scdev_bufs* b, float* d_out, int nBlocks, void* stream)
{
    IfftArgs a;
    fill_ifft_args(a, pl, b, d_out, nBlocks);
    dim3 grid(pl->nOutLocal, nBlocks);
    ifft_batch_kernel<<<grid, pl->fftThreads, fft_smem(pl, 2), (cudaStream_t)stream>>>(a);
    SC_CHECK(cudaGetLastError());
    const size_t n = (size_t)pl->nOutLocal * pl->hop;
    ola_batch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

static void fill_multi_args(MultiArgs& a, const scdev_plan* pl, const scdev_bufs* b, const float* d_in, float* d_out)
{
    a.in = d_in; a.out = d_out; a.H = (const float2*)b->H; a.X = (float2*)b->X;
    a.tw = (const float2*)b->tw; a.tail = b->tail; a.zt = b->zt; a.counters = b->counters;
    a.hop = pl->hop; a.M = pl->M; a.logM = pl->logM; a.P = pl->P; a.RS = pl->RS; a.nCH = pl->nOutLocal;
    a.B = 1; a.G = 1; a.Q = 1;
    a.wT1 = (const float2*)b->wtab; a.wT2 = b->wtab ? (const float2*)b->wtab + pl->M : NULL;
    a.scale = 1.0f / (float)pl->N;
}

/* multiConv, nBlocks device-resident blocks: which = 0 forward FFTs, 1 MAC + inverse FFTs, 2 overlap-add chain */
int scdev_multi_batch(const scdev_plan* pl, const scdev_bufs* b, const float* d_in, float* d_out, int nBlocks, int which, void* stream)
{
    MultiArgs a;
    fill_multi_args(a, pl, b, d_in, d_out);
    const int threads = pl->M < 64 ? 64 : (pl->M > 512 ? 512 : pl->M);
    dim3 grid(pl->nOutLocal, nBlocks);
    /* warp-FFT kernels: forward for 64 <= M <= 1024; MAC + inverse for M <= 512 and P <= 16 (register window).  Both or
     * neither, so that a handle uses one FFT algorithm throughout its batched path. */
    if (b->wtab && which <= 1 && pl->M >= 64 && pl->M <= 512 && pl->P <= 16) {
        a.B = nBlocks;
        cudaStream_t st = (cudaStream_t)stream;
        switch (pl->M) {
            case 64:  return multi_w_launch<2>(pl, a, nBlocks, which, st);
            case 128: return multi_w_launch<4>(pl, a, nBlocks, which, st);
            case 256: return multi_w_launch<8>(pl, a, nBlocks, which, st);
            default:  return multi_w_launch<16>(pl, a, nBlocks, which, st);
        }
    }
    if (which == 0) {
        a.B = nBlocks; a.Q = multi_fft_q(pl);
        dim3 gridQ(pl->nOutLocal, (nBlocks + a.Q - 1) / a.Q);
        multi_fft_batch_kernel<<<gridQ, pl->fftThreads, fft_smem(pl, a.Q + 1), (cudaStream_t)stream>>>(a);
    } else if (which == 1) {
        /* consecutive blocks per CTA: as many as keep >= ~4 CTAs per SM in flight, at most 8 */
        static int gEnv = -1;
        if (gEnv < 0) { const char* v = getenv("SAFCONV_MULTI_G"); gEnv = v ? atoi(v) : 0; }
        int G = gEnv > 0 ? gEnv : 8;
        while (G > 1 && (long long)pl->nOutLocal * ((nBlocks + G - 1) / G) < 4 * 148) G >>= 1;
        a.B = nBlocks; a.G = G;
        dim3 gridG(pl->nOutLocal, (nBlocks + G - 1) / G);
        static int regEnv = -1;
        if (regEnv < 0) { const char* v = getenv("SAFCONV_MULTI_REG"); regEnv = v ? atoi(v) : 1; }
        const int PT = pl->P <= 2 ? 2 : (pl->P <= 4 ? 4 : (pl->P <= 8 ? 8 : 16));
        if (regEnv && pl->P <= 16 && pl->M <= 512) {      /* one thread per bin, <= 128 registers per thread */
            const size_t smem = fft_smem(pl, 2);
            cudaStream_t st = (cudaStream_t)stream;
            switch (PT) {
                case 2:  multi_mac_ifft_batch_reg_kernel<2><<<gridG, pl->M, smem, st>>>(a); break;
                case 4:  multi_mac_ifft_batch_reg_kernel<4><<<gridG, pl->M, smem, st>>>(a); break;
                case 8:  multi_mac_ifft_batch_reg_kernel<8><<<gridG, pl->M, smem, st>>>(a); break;
                default: multi_mac_ifft_batch_reg_kernel<16><<<gridG, pl->M, smem, st>>>(a); break;
            }
        } else {
            multi_mac_ifft_batch_kernel<<<gridG, threads, fft_smem(pl, 2), (cudaStream_t)stream>>>(a);
        }
    } else {
        IfftArgs o;
        o.Zp = NULL; o.grpStart = NULL; o.Zp2 = NULL; o.grpStart2 = NULL; o.headH = NULL; o.headX = NULL; o.P = 0; o.nIn = 0; o.RS = 0; o.tw = NULL; o.out = d_out; o.tail = b->tail; o.zt = b->zt; o.counters = b->counters;
        o.zpStride = 0; o.hop = pl->hop; o.M = pl->M; o.logM = pl->logM; o.nKT = 0; o.OTsz = 0;
        o.nOutLocal = pl->nOutLocal; o.B = nBlocks; o.scale = 0.f;
        const size_t n = (size_t)pl->nOutLocal * pl->hop;
        ola_batch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(o);
    }
    return (int)cudaGetLastError();
}

int scdev_multi_fused(const scdev_plan* pl, const scdev_bufs* b, const float* d_in, float* d_out, void* stream)
{
    MultiArgs a;
    fill_multi_args(a, pl, b, d_in, d_out);
    /* one thread per bin (up to 512): the P filter / delay-line loads of a bin are the latency chain of this kernel */
    int threads = pl->M < 64 ? 64 : (pl->M > 512 ? 512 : pl->M);
    multi_fused_kernel<<<pl->nOutLocal, threads, fft_smem(pl, 3), (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

int scdev_tv_fused(const scdev_plan* pl, const scdev_bufs* b, const float* d_in, float* d_out,
                   int irIdx, int irLast, int irLast2, void* stream)
{
    TvArgs a;
    a.in = d_in; a.out = d_out; a.H = (const float2*)b->H; a.X = (float2*)b->X;
    a.tw = (const float2*)b->tw; a.tail0 = b->tail; a.tail1 = b->tail2; a.counters = b->counters;
    a.hop = pl->hop; a.M = pl->M; a.logM = pl->logM; a.P = pl->P; a.nOut = pl->nOutLocal;
    a.ir0 = irIdx; a.ir1 = irLast; a.ir2 = irLast2;
    a.scale = 1.0f / (float)pl->N;
    if (pl->M > 4096) {
        /* five M-point arrays do not fit in one CTA's shared memory: input FFT, (output, IR set) transforms, cross-fade */
        if (!b->zt) return (int)cudaErrorInvalidValue;
        cudaStream_t st = (cudaStream_t)stream;
        SC_CHECK(sc_optin_smem(tv_input_kernel));
        SC_CHECK(sc_optin_smem(tv_mac_ifft_kernel));
        tv_input_kernel<<<1, pl->fftThreads, fft_smem(pl, 2), st>>>(a);
        SC_CHECK(cudaGetLastError());
        tv_mac_ifft_kernel<<<dim3(pl->nOutLocal, 3), 512, fft_smem(pl, 2), st>>>(a, b->zt);
        SC_CHECK(cudaGetLastError());
        const size_t n = (size_t)pl->nOutLocal * pl->hop;
        tv_xfade_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, b->zt);
        return (int)cudaGetLastError();
    }
    tv_fused_kernel<<<pl->nOutLocal, pl->fftThreads, fft_smem(pl, 5), (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

#define SC_SMALL_THREADS 512
/* bytes of shared memory small_fused_kernel needs for this plan with `threads` threads */
static size_t small_smem(const scdev_plan* pl, int threads)
{
    const int G = threads >= pl->M ? threads / pl->M : 1;
    return ((size_t)pl->nIn * SC_ALEN(pl->M) + (size_t)pl->nIn * pl->M + pl->M + sc_split_len(pl->M)
            + (G > 1 ? (size_t)G * pl->M : 0)) * sizeof(float2);
}

int scdev_small_fits(const scdev_plan* pl, int maxSmemOptin)
{
    if (pl->kind != SC_KIND_MATRIX) return 0;
    if (small_cluster_plan_ok(pl)) return 1;
    if (small_smem(pl, SC_SMALL_THREADS) > (size_t)maxSmemOptin || small_smem(pl, SC_SMALL_THREADS) > 160 * 1024) return 0;
    if (pl->hop > 4 * SC_SMALL_THREADS) return 0;
    if ((long long)pl->nOutLocal * pl->nIn > 256) return 0;                       /* redundant forward FFTs stay cheap */
    if ((double)pl->P * pl->nIn * pl->M * 8.0 > 4.0 * 1024 * 1024) return 0;       /* per-output filter bytes: L2-resident */
    return 1;
}

/* start the resident version of the cluster latency kernel on `stream` (returns cudaErrorNotSupported if the plan is not
 * served by the cluster kernel); mailbox: page-locked, layout ScMailbox */
int scdev_small_resident_start(const scdev_plan* pl, const scdev_bufs* b, void* mailbox, unsigned int lastSeq, unsigned int idleUs,
                               void* stream)
{
    return small_resident_start(pl, b, mailbox, lastSeq, idleUs, (cudaStream_t)stream);
}

/* debugging: the last block's stamps of the resident kernel (call when the device is idle or between blocks) */
int scdev_small_resident_stamps(unsigned long long out[8])
{
    if (!g_resStamps) return -1;
    return (int)cudaMemcpy(out, g_resStamps, 64, cudaMemcpyDeviceToHost);
}

/* done / seq: optional host-visible completion word (page-locked, mapped) -- only the cluster kernel signals it; *signalled
 * tells the caller whether it may poll the word instead of synchronising the stream */
int scdev_small_fused(const scdev_plan* pl, const scdev_bufs* b, const float* in, float* out, void* stream,
                      volatile unsigned int* done, unsigned int seq, int* signalled)
{
    if (signalled) *signalled = 0;
    if (small_cluster_ok(pl, b)) {
        if (signalled) *signalled = done != NULL;
        return scdev_small_cluster(pl, b, in, out, (cudaStream_t)stream, done, seq);
    }
    SmallArgs a;
    a.in = in; a.out = out; a.H = (const float2*)b->H; a.X = (float2*)b->X; a.tw = (const float2*)b->tw;
    a.tail = b->tail; a.counters = b->counters;
    a.hop = pl->hop; a.M = pl->M; a.logM = pl->logM; a.P = pl->P; a.nIn = pl->nIn; a.nKT = pl->nKT;
    a.OTsz = pl->OTsz; a.RS = pl->RS;
    a.scale = 1.0f / (float)pl->N;
    const size_t smem = small_smem(pl, SC_SMALL_THREADS);
    if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(small_fused_kernel));
    small_fused_kernel<<<pl->nOutLocal, SC_SMALL_THREADS, smem, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

/* warp-FFT twiddle tables of a handle (b->wtab), 64 <= M <= 1024; b->tw must be uploaded already */
int scdev_wfft_tables(const scdev_plan* pl, scdev_bufs* b, void* stream)
{
    if (pl->M < 64 || pl->M > 1024) return 0;
    SC_CHECK(cudaMalloc(&b->wtab, (size_t)2 * pl->M * sizeof(float2)));
    wfft_tables_kernel<<<(pl->M + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float2*)b->tw, (float2*)b->wtab,
                                                                            (float2*)b->wtab + pl->M, pl->M, pl->logM - 5);
    return (int)cudaGetLastError();
}

/* batch of real FFTs on device buffers; d_tw = W_N^j, j < N/2 (float2); dir 0 forward, 1 backward */
int scdev_rfft(int N, int logM, int nBatch, int dir, const float* d_in, float* d_out, const void* d_tw, void* stream)
{
    RfftArgs a;
    a.in = d_in; a.out = d_out; a.tw = (const float2*)d_tw; a.N = N; a.M = N / 2; a.logM = logM;
    const int M = a.M;
    int threads = M / 4; if (threads < 32) threads = 32; if (threads > 256) threads = 256;
    const size_t smem = ((size_t)2 * SC_ALEN(M) + sc_split_len(M)) * sizeof(float2);
    if (dir == 0) {
        if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(rfft_forward_kernel));
        rfft_forward_kernel<<<nBatch, threads, smem, (cudaStream_t)stream>>>(a);
    } else {
        if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(rfft_backward_kernel));
        rfft_backward_kernel<<<nBatch, threads, smem, (cudaStream_t)stream>>>(a);
    }
    return (int)cudaGetLastError();
}

int scdev_is_pinned_host(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    /* page-locked AND visible to kernels under the same address (zero-copy paths read / write it directly) */
    return at.type == cudaMemoryTypeHost && at.devicePointer == p;
}

} /* extern "C" */
