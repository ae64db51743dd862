/*
 * safconv_offline.cu -- offline (many frames at once) rendering path of the matrix convolver.
 *
 * Same convolution as saf_matrixConv_apply called once per block from a zero state
 * (reference: /root/reference/framework/modules/saf_utilities/saf_utility_matrixConv.c:209-235), but with
 * ALL T frames of the signal available up front (BASELINE.json configs[4]: "offline batched render").
 * Then the per-bin sum over partitions x inputs
 *
 *     Y_t[no][k] = sum_p sum_ni H_p[no][ni][k] * X_{t-p}[ni][k]
 *
 * re-uses every filter value for all T frames: per bin it is a dense complex contraction, written here as
 * a REAL GEMM with fp32 accumulation on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
 * accumulators in TMEM):
 *
 *     D[frame t][(no,c)] += A[frame t][(p, ni, c')] * B[(no,c)][(p, ni, c')]
 *     A = X (frames on the UMMA M axis, 128 rows per accumulator; the partition shift t-p is a ROW OFFSET
 *         of the shared-memory operand descriptor -- the Toeplitz matrix is never materialised)
 *     B = [Hr -Hi; Hi Hr] (complex product as real 2x2 blocks), N = 2*nOut columns
 *
 * fp32 accuracy on tf32 tensor cores: both operands are split x = hi + lo (hi = rn_tf32(x), lo = rn_tf32(x - hi))
 * and every product is issued as three MMAs  lo*hi + hi*lo + hi*hi  (dropped lo*lo ~2^-24 per product); the
 * tensor core's truncating accumulation is kept short by promoting TMEM chains into fp32 registers (see the GEMM).
 *
 * Operands are kept in global memory already in the UMMA canonical K-major / no-swizzle core-matrix order
 * ([k-group of 4 floats][row][4]) so that one plain TMA bulk copy (cp.async.bulk, 1-D) per k-group lands a
 * ready-to-use shared-memory image: no tensor maps, no swizzle, and a row shift is just +16 bytes per row.
 *
 * Kernels:  offline_pack_filters_kernel  H spectra (streaming layout) -> B operand hi/lo       (once per handle)
 *           offline_fft_kernel           forward real FFT of all frames -> A operand hi/lo
 *           offline_gemm_kernel          the tcgen05 GEMM, one CTA per (256-frame tile, bin)
 *           offline_ifft_kernel          inverse real FFT per (frame, output) + overlap-add into the output signal
 *                                        (reference .c:230-233)
 */
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "safconv_dev.h"
#include "safconv_fft.cuh"
#include "safconv_wfft.cuh"

#define OFF_MT   256     /* frames per CTA tile = two M=128 accumulators                     */
#define OFF_KG   4       /* k-groups (of 4 floats) per pipeline chunk: K = 16 per chunk      */
#define OFF_NH   4       /* filter-operand pipeline stages                                   */
#define OFF_FPC  4       /* frames per forward-FFT CTA (64-byte contiguous operand writes)    */
#define OFF_OPC  8       /* outputs per inverse-FFT CTA (64-byte contiguous spectrum reads)   */
#define OFF_GEMM_THREADS 384   /* three warpgroups: warps 0-7 epilogue (accumulator promotion), warp 8 TMA producer, warp 9 MMA issuer,
                                  warps 10-11 idle (they only fill the third warpgroup, which hands its registers to the other two) */
#define OFF_EPI_REGS  232      /* setmaxnreg: 256 epilogue threads x 232 + 128 x 40 = the 384 x 168 the kernel is launched with */
#define OFF_AUX_REGS  40

/* round-to-nearest tf32 (10 explicit mantissa bits) of an fp32 value, returned in an fp32 container */
__device__ __forceinline__ float tf32_rn(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
/* x = hi + lo with hi = rn_tf32(x), lo = rn_tf32(x - hi): |x - hi - lo| <= 2^-24 |x|, both exactly representable
 * as tf32 so the tensor core's own truncation of the low 13 container bits changes nothing */
__device__ __forceinline__ float tf32_hi(float x) { return tf32_rn(x); }
__device__ __forceinline__ float tf32_lo(float x, float hi) { return tf32_rn(x - hi); }

/* ---- fp16 operands (kind::f16, twice the tf32 MMA rate, half the operand bytes) -------------------------------
 * fp16 carries the same 10 explicit mantissa bits as tf32, so the hi/lo split has the same precision -- as long as
 * the values sit inside fp16's exponent range.  Both operands are therefore multiplied by a power of two (exact)
 * chosen from a bound on their magnitude so that the largest value lands just below 2^15:
 *   A (input spectra):  |X[k]| <= sum_n |x[n]|  -> the largest l1 norm over all (channel, frame) blocks
 *   B (filter spectra): the largest |re|, |im| over the filter spectra
 * hi = rn_f16(s x) is a normal fp16 number down to 2^-29 of the bound, lo = rn_f16(s x - hi) down to 2^-18 of it;
 * below that lo is rounded on fp16's fixed subnormal grid (2^-24), i.e. with an ABSOLUTE error of 2^-40 of the
 * bound -- far below the 2^-22 relative error of the large terms that dominate every output sample.
 * The product of the two scales is divided out (exactly) when the accumulators are stored.
 * scal[0] / scal[1] hold the two bounds as float bit patterns (written with atomicMax on the device). */
__device__ __forceinline__ float pow2_scale(float bound)
{
    if (!(bound > 0.f) || bound > 3.0e38f) return 1.f;
    int e;
    (void)frexpf(bound, &e);                    /* bound = m 2^e, 0.5 <= m < 1  ->  bound 2^(15-e) < 2^15 */
    int sft = 15 - e;
    sft = sft < -60 ? -60 : (sft > 60 ? 60 : sft);
    return ldexpf(1.f, sft);
}
__device__ __forceinline__ void f16_split(float x, __half& hi, __half& lo)
{
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b)
{
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
/* the same split for a (re, im) pair with packed conversions: one cvt.rn.f16x2.f32 per part */
__device__ __forceinline__ void f16_split2(float2 x, uint32_t& hi, uint32_t& lo)
{
    const __half2 h = __floats2half2_rn(x.x, x.y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x.x - hf.x, x.y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

/* largest |v[i]| -> atomicMax on the float bit pattern (non-negative floats order like unsigned ints) */
__global__ void offline_absmax_kernel(const float* __restrict__ v, size_t n, float* out)
{
    float m = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmaxf(m, fabsf(__ldg(v + i)));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f && m <= 3.0e38f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}

/* largest l1 norm of a hop-sized input block: one warp per (frame, channel) block; grid (ceil(T/8), nIn), 256 threads */
__global__ void offline_l1max_kernel(const float* __restrict__ in, size_t inStride, int hop, int T, float* out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + warp, ni = blockIdx.y;
    float sum = 0.f;
    if (t < T) {
        const float* x = in + (size_t)ni * inStride + (size_t)t * hop;
        if ((hop & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
            const float4* x4 = reinterpret_cast<const float4*>(x);
            for (int i = lane; i < (hop >> 2); i += 32) { const float4 v = __ldg(x4 + i); sum += fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w); }
        } else {
            for (int i = lane; i < hop; i += 32) sum += fabsf(__ldg(x + i));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __shared__ float wmax[8];
    if (lane == 0) wmax[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int w = 0; w < 8; ++w) m = fmaxf(m, wmax[w]);
        if (m > 0.f && m <= 3.0e38f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  B operand: HG[hi|lo][bin][p][kg][n][16 bytes],  n = 2*no + c_out,  k = 2*ni + c_in              */
/*  (a k-group = 16 bytes = 4 tf32 or 8 fp16 values along K)                                        */
/* ------------------------------------------------------------------------------------------ */
struct PackArgs {
    const float2* H;       /* [ot][kt][p][ni][OTsz][32] */
    unsigned char* HGhi; unsigned char* HGlo;
    const float* scal;     /* F16: scal[1] = bound on the filter spectra */
    size_t total;          /* elements of H */
    int nKT, P, nIn, OTsz, nOutLocal, nKG, Nn;
    size_t tileBytes;      /* bytes of one output tile (Nn/2 outputs) of HG: output tiles are the outermost dimension */
};

template <bool F16>
__global__ void offline_pack_filters_kernel(PackArgs a)
{
    const float sc = F16 ? pow2_scale(a.scal[1]) : 1.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.total; e += stride) {
        size_t r = e;
        const int b  = (int)(r & 31); r >>= 5;
        const int nl = (int)(r % a.OTsz); r /= a.OTsz;
        const int ni = (int)(r % a.nIn);  r /= a.nIn;
        const int p  = (int)(r % a.P);    r /= a.P;
        const int kt = (int)(r % a.nKT);  r /= a.nKT;
        const int ot = (int)r;
        const int noAll = ot * a.OTsz + nl;
        if (noAll >= a.nOutLocal) continue;
        const int tile = noAll / (a.Nn >> 1), no = noAll - tile * (a.Nn >> 1);      /* output tile of the GEMM, output inside it */
        unsigned char* HGhi = a.HGhi + (size_t)tile * a.tileBytes;
        unsigned char* HGlo = a.HGlo + (size_t)tile * a.tileBytes;
        const int bin = kt * 32 + b;
        const float2 h = a.H[e];
        /* complex product as a real 2x2 block; bin 0 is the packed (DC, Nyquist) pair: two real products */
        float2 rowRe, rowIm;            /* (k = 2ni, k = 2ni+1) entries of rows n = 2no and n = 2no+1 */
        if (bin == 0) { rowRe = make_float2(h.x, 0.f);  rowIm = make_float2(0.f, h.y); }
        else          { rowRe = make_float2(h.x, -h.y); rowIm = make_float2(h.y, h.x); }
        const int k = 2 * ni;
        if (F16) {
            const size_t base = ((((size_t)bin * a.P + p) * a.nKG + (k >> 3)) * a.Nn) * 16 + (size_t)(k & 7) * 2;
            const size_t iRe = base + (size_t)(2 * no) * 16, iIm = base + (size_t)(2 * no + 1) * 16;
            __half h0, l0, h1, l1;
            f16_split(rowRe.x * sc, h0, l0); f16_split(rowRe.y * sc, h1, l1);
            *reinterpret_cast<uint32_t*>(HGhi + iRe) = pack_h2(h0, h1);
            *reinterpret_cast<uint32_t*>(HGlo + iRe) = pack_h2(l0, l1);
            f16_split(rowIm.x * sc, h0, l0); f16_split(rowIm.y * sc, h1, l1);
            *reinterpret_cast<uint32_t*>(HGhi + iIm) = pack_h2(h0, h1);
            *reinterpret_cast<uint32_t*>(HGlo + iIm) = pack_h2(l0, l1);
        } else {
            const size_t base = ((((size_t)bin * a.P + p) * a.nKG + (k >> 2)) * a.Nn) * 16 + (size_t)(k & 3) * 4;
            const size_t iRe = base + (size_t)(2 * no) * 16, iIm = base + (size_t)(2 * no + 1) * 16;
            const float2 reHi = make_float2(tf32_hi(rowRe.x), tf32_hi(rowRe.y));
            const float2 imHi = make_float2(tf32_hi(rowIm.x), tf32_hi(rowIm.y));
            *reinterpret_cast<float2*>(HGhi + iRe) = reHi;
            *reinterpret_cast<float2*>(HGhi + iIm) = imHi;
            *reinterpret_cast<float2*>(HGlo + iRe) = make_float2(tf32_lo(rowRe.x, reHi.x), tf32_lo(rowRe.y, reHi.y));
            *reinterpret_cast<float2*>(HGlo + iIm) = make_float2(tf32_lo(rowIm.x, imHi.x), tf32_lo(rowIm.y, imHi.y));
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  A operand: XG[hi|lo][bin][kg][row][16 bytes], row = (P-1) + frame (P-1 leading zero rows =     */
/*  silence before the first frame); a k-group holds (re, im) of IPC inputs: 2 (tf32) or 4 (fp16)   */
/*  grid (rows / FPC, ceil(nIn/IPC))                                                              */
/* ------------------------------------------------------------------------------------------ */
struct OffFftArgs {
    const float* in;       /* [nIn][T*hop] */
    unsigned char* XGhi; unsigned char* XGlo;
    const float2* tw;
    const float* scal;     /* fp16 operands: scal[0] = bound on the input spectra */
    size_t inStride;       /* T*hop */
    int hop, nIn, M, logM, P, T, rowsAlloc, nKG;
    int fpc;               /* frames per CTA (2 or 4): 32- or 64-byte contiguous operand stores */
    int ipc;               /* inputs per CTA = inputs per k-group: 2 (tf32) or 4 (fp16) */
    const float2* wT1;     /* warp-FFT tables (device, [R][32] each): step-2 twiddles, split twiddles */
    const float2* wT2;
};

__global__ void offline_fft_kernel(OffFftArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    const int MP = SC_ALEN(a.M);                       /* padded FFT work arrays */
    const int FPC = a.fpc, IPC = a.ipc, nTr = IPC * FPC;    /* transforms of this CTA: q = IPC*f + j (frame f, input IPC*kg+j) */
    const int logIPC = (IPC == 4) ? 2 : 1;
    float2* stw = sm + (size_t)nTr * MP;
    /* blockIdx.x walks along the frames: CTAs that run at the same time store ADJACENT 32/64-byte pieces of the
     * same operand rows, so L2 merges them into full lines and DRAM sees long bursts (the first version had the
     * input pair on x and was bound by scattered 64-byte DRAM writes, not by the FFT) */
    const int kg = blockIdx.y;
    const int row0 = blockIdx.x * FPC;
    const bool wide = fft_use_wide(a.M, nTr);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = a.tw;                      /* split-pass twiddles straight from the (L1/L2-resident) global table: keeps 3 CTAs per SM */
    /* input blocks of the transforms: for every chunk of samples the loads of (up to 8) transforms are issued
     * before the first store, so one DRAM round trip covers the whole batch */
    {
        const bool vec = ((a.hop & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 7) == 0) && ((a.inStride & 1) == 0);
        for (int q0 = 0; q0 < nTr; q0 += 8)
        for (int n0 = 0; n0 < a.M; n0 += blockDim.x) {
            const int n = n0 + threadIdx.x;
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int q = q0 + u;
                v[u] = make_float2(0.f, 0.f);
                if (q < nTr && n < a.M) {
                    const int ni = IPC * kg + (q & (IPC - 1));
                    const int t = row0 + (q >> logIPC) - (a.P - 1);
                    if (ni < a.nIn && t >= 0 && t < a.T) {
                        const float* x = a.in + (size_t)ni * a.inStride + (size_t)t * a.hop;
                        if (vec) { if (2 * n < a.hop) v[u] = __ldg(reinterpret_cast<const float2*>(x) + n); }
                        else {
                            if (2 * n < a.hop)     v[u].x = __ldg(x + 2 * n);
                            if (2 * n + 1 < a.hop) v[u].y = __ldg(x + 2 * n + 1);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (q0 + u < nTr && n < a.M) sm[(size_t)(q0 + u) * MP + padi(n, a.logM)] = v[u];
        }
    }
    __syncthreads();
    cfft_dif_batch<false>(sm, a.M, a.logM, stw, nTr, wide);

    const int half = a.M >> 1;
    if (IPC == 4) {
        /* fp16 operands: one 16-byte k-group = (re, im) of four inputs, scaled by a power of two (see pow2_scale) */
        const float sc = pow2_scale(a.scal[0]);
        for (int idx = threadIdx.x; idx < (half + 1) * FPC; idx += blockDim.x) {
            const int f = idx & (FPC - 1), k = idx / FPC;
            const float2* s0 = sm + (size_t)(4 * f) * MP;
            uint32_t hiK[4], loK[4], hiM[4], loM[4];
            int k2 = a.M - k;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2* sj = s0 + (size_t)j * MP;
                float2 x, xm;
                if (k == 0) { const float2 z = sj[0]; x = make_float2(z.x + z.y, z.x - z.y); xm = x; }
                else fwd_split_pair(sj, k, a.M, a.logM, spl, x, xm);
                __half h0, l0, h1, l1;
                f16_split(x.x * sc, h0, l0);  f16_split(x.y * sc, h1, l1);
                hiK[j] = pack_h2(h0, h1);  loK[j] = pack_h2(l0, l1);
                f16_split(xm.x * sc, h0, l0); f16_split(xm.y * sc, h1, l1);
                hiM[j] = pack_h2(h0, h1);  loM[j] = pack_h2(l0, l1);
            }
            if (k == 0) k2 = 0;
            const size_t row = (size_t)row0 + f;
            const size_t o1 = (((size_t)k * a.nKG + kg) * a.rowsAlloc + row) * 16;
            *reinterpret_cast<uint4*>(a.XGhi + o1) = make_uint4(hiK[0], hiK[1], hiK[2], hiK[3]);
            *reinterpret_cast<uint4*>(a.XGlo + o1) = make_uint4(loK[0], loK[1], loK[2], loK[3]);
            if (k2 != k) {
                const size_t o2 = (((size_t)k2 * a.nKG + kg) * a.rowsAlloc + row) * 16;
                *reinterpret_cast<uint4*>(a.XGhi + o2) = make_uint4(hiM[0], hiM[1], hiM[2], hiM[3]);
                *reinterpret_cast<uint4*>(a.XGlo + o2) = make_uint4(loM[0], loM[1], loM[2], loM[3]);
            }
        }
        return;
    }
    for (int idx = threadIdx.x; idx < (half + 1) * FPC; idx += blockDim.x) {
        const int f = idx & (FPC - 1), k = idx / FPC;
        const float2* s0 = sm + (size_t)(2 * f) * MP;
        const float2* s1 = s0 + MP;
        float2 x0, x0m, x1, x1m;
        int k2 = a.M - k;
        if (k == 0) {
            const float2 z0 = s0[0], z1 = s1[0];
            x0 = make_float2(z0.x + z0.y, z0.x - z0.y);  x0m = x0;
            x1 = make_float2(z1.x + z1.y, z1.x - z1.y);  x1m = x1;
            k2 = 0;
        } else {
            fwd_split_pair(s0, k, a.M, a.logM, spl, x0, x0m);
            fwd_split_pair(s1, k, a.M, a.logM, spl, x1, x1m);
        }
        const size_t row = (size_t)row0 + f;
        {
            const float4 v = make_float4(x0.x, x0.y, x1.x, x1.y);
            const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            const size_t o = (((size_t)k * a.nKG + kg) * a.rowsAlloc + row) * 16;
            *reinterpret_cast<float4*>(a.XGhi + o) = hi;
            *reinterpret_cast<float4*>(a.XGlo + o) = make_float4(tf32_lo(v.x, hi.x), tf32_lo(v.y, hi.y), tf32_lo(v.z, hi.z), tf32_lo(v.w, hi.w));
        }
        {
            const float4 v = make_float4(x0m.x, x0m.y, x1m.x, x1m.y);
            const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            const size_t o = (((size_t)k2 * a.nKG + kg) * a.rowsAlloc + row) * 16;
            *reinterpret_cast<float4*>(a.XGhi + o) = hi;
            *reinterpret_cast<float4*>(a.XGlo + o) = make_float4(tf32_lo(v.x, hi.x), tf32_lo(v.y, hi.y), tf32_lo(v.z, hi.z), tf32_lo(v.w, hi.w));
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  tcgen05 helpers                                                                             */
/* ------------------------------------------------------------------------------------------ */

/* shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleaved" core matrices of 8 rows x 16 bytes):
 * element (row i, 16-byte k-chunk j) lives at  start + (i/8)*SBO + (i%8)*16 + j*LBO.
 * With SBO = 128 rows are uniformly 16 bytes apart, so shifting the start address by 16*s selects rows s.. */
__device__ __forceinline__ uint64_t umma_sdesc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lboBytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sboBytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                         /* descriptor version for sm_100 */
    return d;                                       /* base offset 0, layout type 0 = no swizzle */
}

/* instruction descriptor: D = F32, A = B = TF32, both K-major, dense, M x N */
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32(int Mdim, int Ndim)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Ndim >> 3) << 17) | ((uint32_t)(Mdim >> 4) << 24);
}

/* instruction descriptor: D = F32, A = B = F16, both K-major, dense, M x N (K = 16 per instruction) */
__host__ __device__ __forceinline__ uint32_t umma_idesc_f16(int Mdim, int Ndim)
{
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(Ndim >> 3) << 17) | ((uint32_t)(Mdim >> 4) << 24);
}

template <bool F16>
__device__ __forceinline__ void umma_ss(uint32_t tmemD, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    if (F16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            :: "r"(tmemD), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            :: "r"(tmemD), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
/* arrive on an mbarrier once all previously issued MMAs of this thread have completed */
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}
/* one lane of the (fully active) warp */
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

/* two 16-column loads in flight, one wait */
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr0, uint32_t taddr1, float* v)
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr0));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr1));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

/* 64 consecutive columns: four 16-column loads in flight, one wait */
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float* v)
{
    uint32_t r[64];
#pragma unroll
    for (int g = 0; g < 4; ++g)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[16 * g + 0]), "=r"(r[16 * g + 1]), "=r"(r[16 * g + 2]), "=r"(r[16 * g + 3]), "=r"(r[16 * g + 4]),
                       "=r"(r[16 * g + 5]), "=r"(r[16 * g + 6]), "=r"(r[16 * g + 7]), "=r"(r[16 * g + 8]), "=r"(r[16 * g + 9]),
                       "=r"(r[16 * g + 10]), "=r"(r[16 * g + 11]), "=r"(r[16 * g + 12]), "=r"(r[16 * g + 13]), "=r"(r[16 * g + 14]),
                       "=r"(r[16 * g + 15])
                     : "r"(taddr + 16u * g));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

/* 32 consecutive columns, NO wait (the caller overlaps the load with the additions of the previous batch) */
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

/* ------------------------------------------------------------------------------------------ */
/*  the GEMM: grid (Tpad / 256, M bins), 320 threads                                            */
/*                                                                                              */
/*  Accumulator promotion: the tensor core adds each MMA result into its TMEM accumulator with  */
/*  truncation, which over the hundreds of chained MMAs of one output (C5: 256 k-steps x 3)       */
/*  costs ~1e-5 relative -- far outside the 1e-6 parity tolerance.  So TMEM only holds SHORT      */
/*  chains (`flush` partition steps = 6*flush MMAs per accumulator); 8 epilogue warps keep the     */
/*  running sums in registers (fp32, round-to-nearest adds) and drain one TMEM buffer set while    */
/*  the MMA thread fills the other.                                                              */
/* ------------------------------------------------------------------------------------------ */
struct OffGemmArgs {
    const unsigned char *XGhi, *XGlo, *HGhi, *HGlo;
    size_t hgTileBytes;    /* output tile z: HG + z * hgTileBytes, Ys + z * ysTileFloats */
    size_t ysTileFloats;
    const float* scal;     /* fp16 operands: the two magnitude bounds whose scales are divided out of the result */
    int* ticket;           /* tile counter of the dynamic scheduler (zeroed before the launch) */
    float* Ys;             /* [bin][Tpad][Nn] */
    int P, nKG, nKC, Nn, rowsAlloc, Tpad, rowsX, tmemCols, flush;
    int T;                 /* frames of this render: the last frame tile is cut to roundup16(T - t0) columns / one accumulator */
    int nFT, nBins, nTilesTotal;   /* frame tiles, bins, all tiles = nOutTiles * nBins * nFT (frame tile fastest) */
    uint32_t idesc;        /* instruction descriptor with the N field left empty (NT) / complete (frames on M) */
};

#define OFF_EPI_WARPS 8
#define OFF_MAX_HALF  64     /* columns per epilogue thread and frame tile: Nn/2 <= 64 */

struct OffTile { int t0, bin, z, nFr; };
__device__ __forceinline__ OffTile off_tile_decode(const OffGemmArgs& a, int tile)
{
    OffTile r;
    const int ft = tile % a.nFT; tile /= a.nFT;
    r.bin = tile % a.nBins; r.z = tile / a.nBins;
    r.t0 = ft * OFF_MT;
    const int left = a.T - r.t0;
    r.nFr = left >= OFF_MT ? OFF_MT : ((left + 15) & ~15);
    return r;
}

/* NT = true (Nn == 128): the operand roles are swapped -- filters on the UMMA M axis (A, 128 rows), the 256 frames of
 * the tile on the N axis (B) -- so one M128 x N256 x K16 instruction replaces two M128 x N128 ones.  At N = 128 an MMA
 * reads 8 KB of shared memory per 64 tensor-pipe cycles, which is the whole shared-memory read bandwidth of the SM
 * (TMA fills and everything else contend with it); at N = 256 it is 12 KB per 128 cycles.  The Toeplitz row shift
 * works on the B descriptor exactly as on the A descriptor.
 *
 * PERSISTENT: one CTA per SM walks over (frame tile, bin, output tile) work items handed out by a global ticket
 * (frame tile fastest, so the CTAs that run at the same time share a bin's filter operand in L2, exactly like the
 * hardware block scheduler did).  TMEM allocation, barrier set-up and the pipeline ramp are paid once per CTA, and the
 * final promotion + store of tile i overlaps the first chains of tile i+1 (the MMA warp only waits for the TMEM buffer
 * set to be drained, not for the global stores).  The LAST frame tile of a render is cut to roundup16(T - t0) columns
 * (NT: the UMMA N of its instructions; frames on M: one accumulator instead of two when <= 128 frames are left), so a
 * short time shard (multi-GPU: 2813 / 8 + halo = 360 frames) costs 256 + 112 columns, not 512. */
template <bool F16, bool NT>
__global__ void __launch_bounds__(OFF_GEMM_THREADS, 1) offline_gemm_kernel(OffGemmArgs a)
{
    extern __shared__ __align__(128) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t xPlane = (uint32_t)a.rowsX * 16u;        /* one k-group column of the frame tile   */
    const uint32_t xStage = 2u * OFF_KG * xPlane;           /* hi + lo                                */
    const uint32_t hPlane = (uint32_t)a.Nn * 16u;
    const uint32_t hStage = 2u * OFF_KG * hPlane;
    unsigned char* smX = smraw;                             /* [2][hi|lo][kg][rowsX][16 B]            */
    unsigned char* smH = smX + 2u * xStage;                 /* [NH][hi|lo][kg][Nn][16 B]              */
    uint64_t* xfull   = reinterpret_cast<uint64_t*>(smH + (size_t)OFF_NH * hStage);
    uint64_t* xempty  = xfull + 2;
    uint64_t* hfull   = xempty + 2;
    uint64_t* hempty  = hfull + OFF_NH;
    uint64_t* tfull   = hempty + OFF_NH;                    /* TMEM buffer set b holds a finished chain */
    uint64_t* tempty  = tfull + 2;                          /* ... has been drained by all epilogue warps */
    uint64_t* sfull   = tempty + 2;                         /* scheduler slot s holds the next tile id */
    uint64_t* sempty  = sfull + 2;                          /* ... has been read by the MMA warp and all epilogue warps */
    uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(sempty + 2);
    volatile int* sched = reinterpret_cast<volatile int*>(tmemPtr + 2);
    const int steps = a.nKC * a.P;                          /* partition steps of a tile              */
    const int nFlush = (steps + a.flush - 1) / a.flush;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1);
            mbar_init(&tfull[i], 1); mbar_init(&tempty[i], OFF_EPI_WARPS);
            mbar_init(&sfull[i], 1); mbar_init(&sempty[i], OFF_EPI_WARPS + 1);
        }
        for (int i = 0; i < OFF_NH; ++i) { mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == OFF_EPI_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmemPtr)), "r"((uint32_t)a.tmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmemPtr;
    /* register re-allocation between the warpgroups: the epilogue warps keep 128 running sums AND two 32-column batches of a
     * chain in registers (the load of batch i+1 in flight while batch i is added); producer and MMA issuer need very few */
    if (warp >= OFF_EPI_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(OFF_AUX_REGS));
    if (warp == OFF_EPI_WARPS) {
        /* ===================== scheduler + TMA producer ===================== */
        if (lane == 0) {
            int hs = 0; uint32_t hpar = 1;
            uint32_t xc = 0;                                /* frame-tile loads issued so far (all tiles) */
            int next = atomicAdd(a.ticket, 1);              /* the ticket of tile q+1 is drawn while tile q is being loaded */
            for (uint32_t q = 0; ; ++q) {
                const int sl = (int)(q & 1u);
                mbar_wait(&sempty[sl], ((q >> 1) & 1u) ^ 1u);
                const int tile = next < a.nTilesTotal ? next : -1;
                sched[sl] = tile;
                mbar_arrive(&sfull[sl]);                    /* release: the slot is visible to the waiters */
                if (tile < 0) break;
                next = atomicAdd(a.ticket, 1);
                const OffTile tl = off_tile_decode(a, tile);
                /* rows of the frame tile that are needed (frames on M: whole 128-row accumulators) */
                const uint32_t xBytes = (uint32_t)((NT ? tl.nFr : (tl.nFr > 128 ? OFF_MT : 128)) + a.P - 1) * 16u;
                for (int kc = 0; kc < a.nKC; ++kc, ++xc) {
                    const int xs = (int)(xc & 1u);
                    mbar_wait(&xempty[xs], ((xc >> 1) & 1u) ^ 1u);
                    mbar_expect_tx(&xfull[xs], 2u * OFF_KG * xBytes);
                    for (int hl = 0; hl < 2; ++hl)
                        for (int g = 0; g < OFF_KG; ++g) {
                            const unsigned char* src = (hl ? a.XGlo : a.XGhi)
                                             + (((size_t)tl.bin * a.nKG + (size_t)kc * OFF_KG + g) * a.rowsAlloc + tl.t0) * 16;
                            tma_bulk_g2s(smX + (size_t)xs * xStage + (size_t)(hl * OFF_KG + g) * xPlane, src, xBytes, &xfull[xs]);
                        }
                    for (int p = 0; p < a.P; ++p) {
                        mbar_wait(&hempty[hs], hpar);
                        mbar_expect_tx(&hfull[hs], hStage);
                        for (int hl = 0; hl < 2; ++hl) {
                            const unsigned char* src = (hl ? a.HGlo : a.HGhi) + (size_t)tl.z * a.hgTileBytes
                                             + ((((size_t)tl.bin * a.P + p) * a.nKG + (size_t)kc * OFF_KG) * a.Nn) * 16;
                            tma_bulk_g2s(smH + (size_t)hs * hStage + (size_t)hl * OFF_KG * hPlane, src, OFF_KG * hPlane, &hfull[hs]);
                        }
                        if (++hs == OFF_NH) { hs = 0; hpar ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == OFF_EPI_WARPS + 1) {
        /* ===================== MMA issuer ===================== */
        /* The whole warp runs the (uniform) control flow; one elected lane issues the tcgen05 instructions.
         * Descriptors are a per-kernel constant plus the 16-byte-unit start address, so each MMA costs a
         * couple of integer adds on the issuing thread, not a descriptor rebuild. */
        const bool leader = elect_one();
        const uint64_t aDesc0 = umma_sdesc(0, xPlane, 128);
        const uint64_t bDesc0 = umma_sdesc(0, hPlane, 128);
        const uint32_t xPlane16 = xPlane >> 4, hPlane16 = hPlane >> 4;
        int hs = 0; uint32_t hpar = 0;
        uint32_t nf = 0, xc = 0;                            /* chains / frame-tile loads consumed so far (all tiles) */
        for (uint32_t q = 0; ; ++q) {
            const int sl = (int)(q & 1u);
            mbar_wait(&sfull[sl], (q >> 1) & 1u);
            const int tile = sched[sl];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sempty[sl]);
            if (tile < 0) break;
            const OffTile tl = off_tile_decode(a, tile);
            const uint32_t idesc = NT ? (a.idesc | ((uint32_t)(tl.nFr >> 3) << 17)) : a.idesc;
            const int nAcc = NT ? 1 : (tl.nFr > 128 ? 2 : 1);
            int inChain = 0, step = 0;
            for (int kc = 0; kc < a.nKC; ++kc, ++xc) {
                const int xs = (int)(xc & 1u);
                mbar_wait(&xfull[xs], (xc >> 1) & 1u);
                const uint32_t xBase16 = smem_u32(smX + (size_t)xs * xStage) >> 4;
                for (int p = 0; p < a.P; ++p, ++step) {
                    const int tb = (int)(nf & 1u);
                    if (inChain == 0)                                       /* new chain: its TMEM buffer set must be drained */
                        mbar_wait(&tempty[tb], ((nf >> 1) & 1u) ^ 1u);
                    mbar_wait(&hfull[hs], hpar);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t hBase16 = smem_u32(smH + (size_t)hs * hStage) >> 4;
                        const uint32_t rowShift = (uint32_t)(a.P - 1 - p);  /* frame t-p = tile row +(P-1-p); 16 B per row */
                        if (NT) {
#pragma unroll
                            for (int ks = 0; ks < OFF_KG / 2; ++ks) {
                                const uint32_t x16 = xBase16 + (uint32_t)(2 * ks) * xPlane16 + rowShift;
                                const uint32_t h16 = hBase16 + (uint32_t)(2 * ks) * hPlane16;
                                const uint64_t xHi = aDesc0 + x16, xLo = aDesc0 + x16 + OFF_KG * xPlane16;
                                const uint64_t hHi = bDesc0 + h16, hLo = bDesc0 + h16 + OFF_KG * hPlane16;
                                const uint32_t d = tmem + (uint32_t)(tb * 256);
                                umma_ss<F16>(d, hLo, xHi, idesc, (inChain | ks) ? 1u : 0u);   /* lo*hi */
                                umma_ss<F16>(d, hHi, xLo, idesc, 1u);                         /* hi*lo */
                                umma_ss<F16>(d, hHi, xHi, idesc, 1u);                         /* hi*hi */
                            }
                        } else
                        for (int acc = 0; acc < nAcc; ++acc) {
#pragma unroll
                            for (int ks = 0; ks < OFF_KG / 2; ++ks) {
                                const uint32_t a16 = xBase16 + (uint32_t)(2 * ks) * xPlane16 + (uint32_t)acc * 128u + rowShift;
                                const uint32_t b16 = hBase16 + (uint32_t)(2 * ks) * hPlane16;
                                const uint64_t aHi = aDesc0 + a16, aLo = aDesc0 + a16 + OFF_KG * xPlane16;
                                const uint64_t bHi = bDesc0 + b16, bLo = bDesc0 + b16 + OFF_KG * hPlane16;
                                const uint32_t d = tmem + (uint32_t)((tb * 2 + acc) * a.Nn);
                                /* small cross terms first, the large hi*hi term last */
                                umma_ss<F16>(d, aLo, bHi, idesc, (inChain | ks) ? 1u : 0u);   /* lo*hi */
                                umma_ss<F16>(d, aHi, bLo, idesc, 1u);                         /* hi*lo */
                                umma_ss<F16>(d, aHi, bHi, idesc, 1u);                         /* hi*hi */
                            }
                        }
                        umma_commit(&hempty[hs]);                           /* filter stage free once these MMAs retire */
                    }
                    if (++hs == OFF_NH) { hs = 0; hpar ^= 1u; }
                    if (++inChain == a.flush || step + 1 == steps) {
                        if (leader) umma_commit(&tfull[tb]);                /* chain complete -> epilogue may drain it */
                        ++nf; inChain = 0;
                    }
                    __syncwarp();
                }
                if (leader) umma_commit(&xempty[xs]);
                __syncwarp();
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(OFF_EPI_REGS));
        /* ===================== epilogue warps: promote TMEM chains into fp32 register sums ===================== */
        const int q4 = warp & 3, hh = warp >> 2;           /* TMEM lane quarter (= warp % 4), column half */
        const int halfN = a.Nn >> 1;
        /* fp16 operands were scaled by exact powers of two: divide them out (exact as well) */
        const float inv = F16 ? 1.f / (pow2_scale(a.scal[0]) * pow2_scale(a.scal[1])) : 1.f;
        float sum[2][OFF_MAX_HALF];
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int c = 0; c < OFF_MAX_HALF; ++c) sum[t][c] = 0.f;
        uint32_t g = 0;                                     /* chains drained so far (all tiles) */
        for (uint32_t q = 0; ; ++q) {
            const int sl = (int)(q & 1u);
            mbar_wait(&sfull[sl], (q >> 1) & 1u);
            const int tile = sched[sl];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sempty[sl]);
            if (tile < 0) break;
            const OffTile tl = off_tile_decode(a, tile);
            for (int f = 0; f < nFlush; ++f, ++g) {
                const int tb = (int)(g & 1u);
                mbar_wait(&tfull[tb], (g >> 1) & 1u);
                tc_fence_after();
                /* full tile: the 128 columns of this thread as four 32-column batches, the load of batch i+1 in flight while
                 * batch i is added -- a chain's pace is this drain (wake-up, load latency, adds, arrive), not the MMA */
                const bool fullDrain = (halfN == 64) && (NT ? (hh * 128 + 128 <= tl.nFr) : (tl.nFr > 128));
                if (fullDrain) {
                    const uint32_t lane16 = (uint32_t)(q4 * 32) << 16;
                    const uint32_t c_t0 = NT ? (uint32_t)(tb * 256 + hh * 128) : (uint32_t)((tb * 2 + 0) * a.Nn + hh * halfN);
                    const uint32_t c_t1 = NT ? c_t0 + 64u                      : (uint32_t)((tb * 2 + 1) * a.Nn + hh * halfN);
                    uint32_t ra[32], rb[32];
                    tmem_ld32_nowait(tmem + lane16 + c_t0, ra);
                    tmem_ld32_nowait(tmem + lane16 + c_t0 + 32u, rb);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) sum[0][i] += __uint_as_float(ra[i]);
                    tmem_ld32_nowait(tmem + lane16 + c_t1, ra);
#pragma unroll
                    for (int i = 0; i < 32; ++i) sum[0][32 + i] += __uint_as_float(rb[i]);
                    tmem_ld32_nowait(tmem + lane16 + c_t1 + 32u, rb);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) sum[1][i] += __uint_as_float(ra[i]);
#pragma unroll
                    for (int i = 0; i < 32; ++i) sum[1][32 + i] += __uint_as_float(rb[i]);
                } else
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    /* NT: this thread owns frames (columns) hh*128 + t*64 .. +63; frames on M: accumulator t = frames t*128 .. */
                    if (NT ? (hh * 128 + t * 64 >= tl.nFr) : (t * 128 >= tl.nFr)) continue;     /* warp-uniform */
#pragma unroll
                    for (int c0 = 0; c0 < OFF_MAX_HALF; c0 += 32) {
                        /* NT: lane = filter row n = 32 q4 + lane, columns = frames */
                        const uint32_t col = NT ? (uint32_t)(tb * 256 + hh * 128 + t * 64 + c0)
                                                : (uint32_t)((tb * 2 + t) * a.Nn + hh * halfN + c0);
                        const uint32_t ta = tmem + ((uint32_t)(q4 * 32) << 16) + col;
                        if (halfN == 64) {                       /* full tile: all 64 columns of this half with one wait */
                            if (c0 == 0) {
                                float v[64];
                                tmem_ld64(ta, v);
#pragma unroll
                                for (int i = 0; i < 64; ++i) sum[t][i] += v[i];
                            }
                        } else if (c0 + 16 < halfN) {
                            float v[32];
                            tmem_ld16x2(ta, ta + 16, v);
#pragma unroll
                            for (int i = 0; i < 32; ++i) sum[t][c0 + i] += v[i];
                        } else if (c0 < halfN) {
                            float v[16];
                            tmem_ld16(ta, v);
#pragma unroll
                            for (int i = 0; i < 16; ++i) sum[t][c0 + i] += v[i];
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[tb]);
            }
            /* the MMA warp is already filling the buffer sets with the next tile's chains while these stores drain */
            float* Ys = a.Ys + (size_t)tl.z * a.ysTileFloats;
            if (NT) {
                /* sum[t][c] = output row n = 32 q4 + lane of frame t0 + hh*128 + t*64 + c: a warp stores 32 consecutive n.
                 * NT implies Nn == 128, so all store offsets are immediates off one base pointer per tile. */
                float* dst = Ys + ((size_t)tl.bin * a.Tpad + tl.t0 + hh * 128) * 128 + (q4 * 32 + lane);
                if (tl.nFr == OFF_MT) {
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int c = 0; c < OFF_MAX_HALF; ++c) { dst[(t * 64 + c) * 128] = sum[t][c] * inv; sum[t][c] = 0.f; }
                } else {
                    const int left = tl.nFr - hh * 128;            /* columns of this half that exist */
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int c = 0; c < OFF_MAX_HALF; ++c) {
                            if (t * 64 + c < left) dst[(t * 64 + c) * 128] = sum[t][c] * inv;
                            sum[t][c] = 0.f;
                        }
                }
            } else
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int fr = t * 128 + q4 * 32 + lane;
                float* dst = Ys + ((size_t)tl.bin * a.Tpad + tl.t0 + fr) * a.Nn + hh * halfN;
#pragma unroll
                for (int c0 = 0; c0 < OFF_MAX_HALF; c0 += 4) {
                    if (c0 < halfN && t * 128 < tl.nFr)
                        *reinterpret_cast<float4*>(dst + c0) = make_float4(sum[t][c0] * inv, sum[t][c0 + 1] * inv, sum[t][c0 + 2] * inv, sum[t][c0 + 3] * inv);
                    sum[t][c0] = sum[t][c0 + 1] = sum[t][c0 + 2] = sum[t][c0 + 3] = 0.f;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == OFF_EPI_WARPS + 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)a.tmemCols) : "memory");
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  inverse FFT per (frame, output): grid (ceil(nOut/8), T)                                      */
/* ------------------------------------------------------------------------------------------ */
struct OffIfftArgs {
    const float2* Ys;      /* [bin][Tpad][Nn/2] complex */
    float* out;            /* [nOut][(T-skip)*hop], zeroed before the launch */
    const float2* tw;
    int hop, M, logM, nOut, Nn2, Tpad, T;
    size_t ysTileC;        /* complex elements of one output tile of Ys ([tile][bin][Tpad][Nn2]) */
    int logNn2;            /* Nn2 = outputs per tile, a power of two */
    int opc;               /* outputs per inverse-FFT CTA (4 or 8): 32- or 64-byte contiguous spectrum loads */
    int skip;              /* leading halo frames that are transformed but not written to `out` */
    float scale;
    const float2* wT1;     /* warp-FFT step-2 twiddle table (device, [R][32]) */
};

__global__ void offline_ifft_kernel(OffIfftArgs a)
{
    extern __shared__ __align__(16) float2 sm[];
    const int MP = SC_ALEN(a.M);
    const int OPC = a.opc;
    float2* stw = sm + (size_t)OPC * MP;
    const int og = blockIdx.x, t = blockIdx.y;
    const bool wide = fft_use_wide(a.M, OPC);
    load_twiddles(stw, a.tw, a.M, a.logM, wide);
    const float2* spl = a.tw;                      /* split-pass twiddles straight from the (L1/L2-resident) global table: keeps 3 CTAs per SM */
    /* spectra of the OPC outputs of this frame: eight independent loads in flight per thread */
    {
        const int logOPC = (OPC == 8) ? 3 : 2;
        const int total = a.M << logOPC;
        for (int base = threadIdx.x; base < total; base += 8 * blockDim.x) {
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * blockDim.x;
                const int j = idx & (OPC - 1), k = idx >> logOPC;
                const int no = og * OPC + j;
                v[u] = (idx < total && no < a.nOut) ? __ldg(a.Ys + (size_t)(no >> a.logNn2) * a.ysTileC + (((size_t)k * a.Tpad + t) << a.logNn2) + (no & (a.Nn2 - 1))) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * blockDim.x;
                if (idx < total) sm[(size_t)(idx & (OPC - 1)) * MP + padi(idx >> logOPC, a.logM)] = v[u];
            }
        }
    }
    __syncthreads();
    {
        const int nArr = min(OPC, a.nOut - og * OPC);
        inv_split_batch(sm, a.M, a.logM, spl, nArr);
        cfft_dif_batch<true>(sm, a.M, a.logM, stw, nArr, wide);
    }
    /* overlap-add (reference .c:230-233) straight into the output signal: out[frame t] += z[0:hop],
     * out[frame t+1] += z[hop:2hop].  `out` is zeroed beforehand and every sample receives exactly two addends
     * (one from each of two CTAs), so the atomic adds give a deterministic result: 0 + a + b == 0 + b + a. */
    const int To = a.T - a.skip;
    for (int j = 0; j < OPC; ++j) {
        const int no = og * OPC + j;
        if (no >= a.nOut) break;
        const float2* s = sm + (size_t)j * MP;
        float* o = a.out + (size_t)no * To * a.hop;
        const int f0 = t - a.skip, f1 = t + 1 - a.skip;           /* output frames fed by the two halves */
        for (int i = threadIdx.x; i < a.hop; i += blockDim.x) {
            if (f0 >= 0)            atomicAdd(o + (size_t)f0 * a.hop + i, time_sample(s, i, a.logM) * a.scale);
            if (f1 >= 0 && f1 < To) atomicAdd(o + (size_t)f1 * a.hop + i, time_sample(s, i + a.hop, a.logM) * a.scale);
        }
    }
}


/* ------------------------------------------------------------------------------------------ */
/*  Warp-FFT versions of the two transform kernels (fp16 operands, M = 32 R <= 1024):            */
/*  every warp owns one transform and keeps it in registers (safconv_wfft.cuh); shared memory is  */
/*  only the staging area that turns the per-transform results into the operand layout's          */
/*  contiguous pieces (forward) / the contiguous spectrum reads into per-transform lanes (inverse) */
/* ------------------------------------------------------------------------------------------ */
#define OFFW_THREADS 256          /* 8 warps = 8 transforms per CTA */

/* forward: grid (rows / 2, ceil(nIn / 4)); warp q -> frame row0 + (q >> 2), input 4 kg + (q & 3) */
template <int R>
__global__ void __launch_bounds__(OFFW_THREADS, 2) offline_fft_w_kernel(OffFftArgs a)
{
    constexpr int M = 32 * R, LOGR = wf_log2(R);
    extern __shared__ __align__(16) unsigned char smw[];
    const float2* __restrict__ T1 = a.wT1;                     /* [R][32] W_M^(l * bitrev_R(i)), read through L1   */
    const float2* __restrict__ T2 = a.wT2;                     /* [R][32] W_N^k of the bin that (lane, slot) holds */
    uint32_t* stgHi = reinterpret_cast<uint32_t*>(smw);        /* [k2 * 32 + k1][9]: the 8 transforms (+1 pad)     */
    uint32_t* stgLo = stgHi + M * 9;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kg = blockIdx.y, row0 = blockIdx.x * 2;
    const WfftLane L = wfft_lane_init<false>(a.tw, M, lane);
    const int f = warp >> 2, ni = 4 * kg + (warp & 3);
    const int t = row0 + f - (a.P - 1);
    float2 v[R];
    {
        const bool valid = ni < a.nIn && t >= 0 && t < a.T;
        const float* x = a.in + (size_t)(valid ? ni : 0) * a.inStride + (size_t)(valid ? t : 0) * a.hop;
        const bool vec = ((a.hop & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 7) == 0) && ((a.inStride & 1) == 0);
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int n = lane + 32 * i;
            v[i] = make_float2(0.f, 0.f);
            if (valid) {
                if (vec) { if (2 * n < a.hop) v[i] = __ldg(reinterpret_cast<const float2*>(x) + n); }
                else {
                    if (2 * n < a.hop)     v[i].x = __ldg(x + 2 * n);
                    if (2 * n + 1 < a.hop) v[i].y = __ldg(x + 2 * n + 1);
                }
            }
        }
    }
    wfft<R, false>(v, T1, lane, L);

    /* real-FFT split in registers: X[k] = (E + W_N^k O) / 2, E = a + conj b, O = -i (a - conj b), b = Z[M - k] */
    const float sc = 0.5f * pow2_scale(a.scal[0]);
    const int k1 = (int)(__brev((unsigned)lane) >> 27);
    const int pl0 = (int)(__brev((unsigned)((32 - k1) & 31)) >> 27);
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int k2 = wf_bitrev(i, LOGR);
        const int ip = wf_bitrev((R - k2) % R, LOGR);
        float2 b;
        if (k2 == 0) { b.x = __shfl_sync(0xffffffffu, v[0].x, pl0);     b.y = __shfl_sync(0xffffffffu, v[0].y, pl0); }
        else         { b.x = __shfl_xor_sync(0xffffffffu, v[ip].x, 31); b.y = __shfl_xor_sync(0xffffffffu, v[ip].y, 31); }
        const float2 A = v[i];
        const float2 E = make_float2(A.x + b.x, A.y - b.y);
        const float2 O = make_float2(A.y + b.y, b.x - A.x);
        const float2 tt = cmulf(__ldg(T2 + i * 32 + lane), O);
        float2 X = make_float2((E.x + tt.x) * sc, (E.y + tt.y) * sc);
        if (k2 == 0 && lane == 0) X = make_float2((A.x + A.y) * (2.f * sc), (A.x - A.y) * (2.f * sc));   /* packed (DC, Nyquist) */
        __half h0, l0, h1, l1;
        f16_split(X.x, h0, l0);  f16_split(X.y, h1, l1);
        const int w = (k2 * 32 + k1) * 9 + warp;
        stgHi[w] = pack_h2(h0, h1);
        stgLo[w] = pack_h2(l0, l1);
    }
    __syncthreads();
    /* one 16-byte k-group (4 inputs) per row and bin; the two rows of the CTA are adjacent in the operand and are
     * written by adjacent lanes: one full 32-byte sector per lane pair and store instruction */
    for (int idx = threadIdx.x; idx < 2 * M; idx += OFFW_THREADS) {
        const int f2 = idx & 1, q = idx >> 1;
        const int kk1 = q & 31, kk2 = q >> 5;
        const int k = kk2 + R * kk1;
        const uint32_t* h = stgHi + (kk2 * 32 + kk1) * 9 + 4 * f2;
        const uint32_t* l = stgLo + (kk2 * 32 + kk1) * 9 + 4 * f2;
        const size_t o = (((size_t)k * a.nKG + kg) * a.rowsAlloc + row0 + f2) * 16;
        *reinterpret_cast<uint4*>(a.XGhi + o) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(a.XGlo + o) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

/* inverse: grid (ceil(nOut / 8), T); warp o -> output 8 og + o of frame t */
template <int R>
__global__ void __launch_bounds__(OFFW_THREADS, 2) offline_ifft_w_kernel(OffIfftArgs a)
{
    constexpr int M = 32 * R, LOGR = wf_log2(R);
    constexpr int OS = M + 33;                                 /* float2 per output in the time-domain staging */
    extern __shared__ __align__(16) unsigned char smw[];
    const float2* __restrict__ T1 = a.wT1;                     /* read through L1 */
    float2* stg = reinterpret_cast<float2*>(smw);              /* spectra [k][9] (8 outputs + pad), then time samples [8][OS] */
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int og = blockIdx.x, t = blockIdx.y;
    const WfftLane L = wfft_lane_init<true>(a.tw, M, lane);
    /* spectra of the 8 outputs of this frame: 64-byte contiguous pieces, eight independent loads in flight per thread */
    for (int base = threadIdx.x; base < M * 8; base += 8 * OFFW_THREADS) {
        float2 u[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int idx = base + q * OFFW_THREADS;
            const int j = idx & 7, k = idx >> 3;
            const int no = og * 8 + j;
            u[q] = (idx < M * 8 && no < a.nOut) ? __ldg(a.Ys + (size_t)(no >> a.logNn2) * a.ysTileC + (((size_t)k * a.Tpad + t) << a.logNn2) + (no & (a.Nn2 - 1))) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int idx = base + q * OFFW_THREADS;
            if (idx < M * 8) stg[(idx >> 3) * 9 + (idx & 7)] = u[q];
        }
    }
    __syncthreads();
    /* inverse split while loading the transform into (lane, slot) order: Zc[k] = E + i O,
     * E = A + conj B, O = (A - conj B) conj(W_N^k), B = Y[M - k]  (the 1/2 is folded into 1/N) */
    float2 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int k = lane + 32 * i;
        const float2 A = stg[k * 9 + warp], B = stg[((M - k) & (M - 1)) * 9 + warp];
        const float2 E = make_float2(A.x + B.x, A.y - B.y);
        const float2 D = make_float2(A.x - B.x, A.y + B.y);
        const float2 O = cmul_conjb(D, __ldg(a.tw + k));
        v[i] = make_float2(E.x - O.y, E.y + O.x);
        if (i == 0 && lane == 0) v[i] = make_float2(A.x + A.y, A.x - A.y);      /* (DC, Nyquist) */
    }
    wfft<R, true>(v, T1, lane, L);
    __syncthreads();                                           /* every warp is done with the spectra */
    {
        float2* ost = stg + (size_t)warp * OS;
        const int k1 = (int)(__brev((unsigned)lane) >> 27);
#pragma unroll
        for (int i = 0; i < R; ++i)                            /* z[n], n = n2 + R k1, stored at n2 + (R + 1) k1 */
            ost[wf_bitrev(i, LOGR) + (R + 1) * k1] = make_float2(v[i].x * a.scale, v[i].y * a.scale);
    }
    __syncthreads();
    /* overlap-add (reference .c:230-233) with two-addend atomic adds, exactly like offline_ifft_kernel */
    const int To = a.T - a.skip;
    const int f0 = t - a.skip, f1 = t + 1 - a.skip;
    for (int j = 0; j < 8; ++j) {
        const int no = og * 8 + j;
        if (no >= a.nOut) break;
        const float* ost = reinterpret_cast<const float*>(stg + (size_t)j * OS);
        float* o = a.out + (size_t)no * To * a.hop;
        for (int i = threadIdx.x; i < a.hop; i += OFFW_THREADS) {
            const int s1 = i + a.hop;
            const int n0 = i >> 1, n1 = s1 >> 1;
            const float z0 = ost[2 * ((n0 & (R - 1)) + (R + 1) * (n0 >> LOGR)) + (i & 1)];
            const float z1 = ost[2 * ((n1 & (R - 1)) + (R + 1) * (n1 >> LOGR)) + (s1 & 1)];
            if (f0 >= 0)            atomicAdd(o + (size_t)f0 * a.hop + i, z0);
            if (f1 >= 0 && f1 < To) atomicAdd(o + (size_t)f1 * a.hop + i, z1);
        }
    }
}


/* ---- M = 1024: the same two kernels on the shuffle-free 32 x 32 register FFT (wfft32t) ---- */

/* forward: warp q -> frame row0 + (q >> 2), input 4 kg + (q & 3); the complex spectra Z of the 8 transforms are left
 * in the warps' tiles in natural order, the whole CTA then does the real-FFT split, the fp16 hi/lo conversion and
 * the operand stores (one 16-byte k-group per thread and (bin, row)) */
__global__ void __launch_bounds__(OFFW_THREADS, 2) offline_fft_t_kernel(OffFftArgs a)
{
    constexpr int M = 1024;
    extern __shared__ __align__(16) unsigned char smw[];
    float2* tiles = reinterpret_cast<float2*>(smw);            /* [8][WFFT_TILE] */
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kg = blockIdx.y, row0 = blockIdx.x * 2;
    float2* tile = tiles + (size_t)warp * WFFT_TILE;
    {
        const int f = warp >> 2, ni = 4 * kg + (warp & 3);
        const int t = row0 + f - (a.P - 1);
        float2 v[32];
        const bool valid = ni < a.nIn && t >= 0 && t < a.T;
        const float* x = a.in + (size_t)(valid ? ni : 0) * a.inStride + (size_t)(valid ? t : 0) * a.hop;
        const bool vec = ((a.hop & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 7) == 0) && ((a.inStride & 1) == 0);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = lane + 32 * i;
            v[i] = make_float2(0.f, 0.f);
            if (valid) {
                if (vec) { if (2 * n < a.hop) v[i] = __ldg(reinterpret_cast<const float2*>(x) + n); }
                else {
                    if (2 * n < a.hop)     v[i].x = __ldg(x + 2 * n);
                    if (2 * n + 1 < a.hop) v[i].y = __ldg(x + 2 * n + 1);
                }
            }
        }
#pragma unroll
        for (int i = 16; i < 32; ++i) v[i] = make_float2(0.f, 0.f);
        /* hop <= M real samples = at most M/2 complex points: v[16..31] are zero, the first stage needs no additions */
        wfft32t<false, true>(v, a.wT1, lane, tile);
#pragma unroll
        for (int i = 0; i < 32; ++i) tile[lane + 32 * wf_bitrev(i, 5)] = v[i];      /* Z[k], natural order */
    }
    __syncthreads();
    const float sc = 0.5f * pow2_scale(a.scal[0]);
    for (int idx = threadIdx.x; idx < 2 * M; idx += OFFW_THREADS) {
        const int f = idx & 1, k = idx >> 1;
        const float2 w = __ldg(a.tw + k);
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2* z = tiles + (size_t)(4 * f + j) * WFFT_TILE;
            const float2 A = z[k], b = z[(M - k) & (M - 1)];
            /* X[k] = (E + W_N^k O) / 2, E = a + conj b, O = -i (a - conj b) */
            const float2 E = make_float2(A.x + b.x, A.y - b.y);
            const float2 O = make_float2(A.y + b.y, b.x - A.x);
            const float2 tt = cmulf(w, O);
            float2 X = make_float2((E.x + tt.x) * sc, (E.y + tt.y) * sc);
            if (k == 0) X = make_float2((A.x + A.y) * (2.f * sc), (A.x - A.y) * (2.f * sc));   /* packed (DC, Nyquist) */
            f16_split2(X, hi[j], lo[j]);
        }
        const size_t o = (((size_t)k * a.nKG + kg) * a.rowsAlloc + row0 + f) * 16;
        *reinterpret_cast<uint4*>(a.XGhi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(a.XGlo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

/* inverse: warp o -> output 8 og + o of frame t */
__global__ void __launch_bounds__(OFFW_THREADS, 2) offline_ifft_t_kernel(OffIfftArgs a)
{
    constexpr int M = 1024;
    extern __shared__ __align__(16) unsigned char smw[];
    float2* stg = reinterpret_cast<float2*>(smw);              /* spectra [k][9] (8 outputs + pad); then the warps' tiles */
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int og = blockIdx.x, t = blockIdx.y;
    {
        /* all 32 loads of a thread are in flight before the first store: ONE DRAM round trip per CTA for the 64 KB of
         * spectra (the registers are free here: the transform has not started) */
        float2 u[32];
        const int j = threadIdx.x & 7, no = og * 8 + j;
        const float2* src = a.Ys + (size_t)(no >> a.logNn2) * a.ysTileC + ((size_t)t << a.logNn2) + (no & (a.Nn2 - 1));
        const size_t kStride = (size_t)a.Tpad << a.logNn2;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const int k = (threadIdx.x >> 3) + q * (OFFW_THREADS / 8);
            u[q] = (no < a.nOut) ? __ldg(src + (size_t)k * kStride) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const int k = (threadIdx.x >> 3) + q * (OFFW_THREADS / 8);
            stg[k * 9 + j] = u[q];
        }
    }
    __syncthreads();
    float2 v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int k = lane + 32 * i;
        const float2 A = stg[k * 9 + warp], B = stg[((M - k) & (M - 1)) * 9 + warp];
        const float2 E = make_float2(A.x + B.x, A.y - B.y);
        const float2 D = make_float2(A.x - B.x, A.y + B.y);
        const float2 O = cmul_conjb(D, __ldg(a.tw + k));
        v[i] = make_float2(E.x - O.y, E.y + O.x);
        if (i == 0 && lane == 0) v[i] = make_float2(A.x + A.y, A.x - A.y);      /* (DC, Nyquist) */
    }
    __syncthreads();                                           /* every warp is done with the spectra */
    float2* tile = stg + (size_t)warp * WFFT_TILE;
    wfft32t<true>(v, a.wT1, lane, tile);
#pragma unroll
    for (int i = 0; i < 32; ++i)                               /* z[n], natural order */
        tile[lane + 32 * wf_bitrev(i, 5)] = make_float2(v[i].x * a.scale, v[i].y * a.scale);
    __syncthreads();
    const int To = a.T - a.skip;
    const int f0 = t - a.skip, f1 = t + 1 - a.skip;
    for (int j = 0; j < 8; ++j) {
        const int no = og * 8 + j;
        if (no >= a.nOut) break;
        const float* z = reinterpret_cast<const float*>(stg + (size_t)j * WFFT_TILE);
        float* o = a.out + (size_t)no * To * a.hop;
        if ((a.hop & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0)) {
            /* 16-byte vector reductions (red.global.add.v4.f32): a quarter of the atomic operations; every component
             * still receives exactly its two addends */
            for (int i = 4 * threadIdx.x; i < a.hop; i += 4 * OFFW_THREADS) {
                if (f0 >= 0)            atomicAdd(reinterpret_cast<float4*>(o + (size_t)f0 * a.hop + i), *reinterpret_cast<const float4*>(z + i));
                if (f1 >= 0 && f1 < To) atomicAdd(reinterpret_cast<float4*>(o + (size_t)f1 * a.hop + i), *reinterpret_cast<const float4*>(z + i + a.hop));
            }
        } else
        for (int i = threadIdx.x; i < a.hop; i += OFFW_THREADS) {
            if (f0 >= 0)            atomicAdd(o + (size_t)f0 * a.hop + i, z[i]);
            if (f1 >= 0 && f1 < To) atomicAdd(o + (size_t)f1 * a.hop + i, z[i + a.hop]);
        }
    }
}

template <int R> static size_t offw_fft_smem()  { return (size_t)(2 * 32 * R * 9) * 4; }
template <int R> static size_t offw_ifft_smem()
{
    const size_t M = 32 * R, a = M * 9, b = 8 * (M + 33);
    return (a > b ? a : b) * 8;
}

template <int R>
static int offw_launch(const OffFftArgs* f, const OffIfftArgs* i, dim3 grid, cudaStream_t st)
{
    if (f) {
        SC_CHECK(sc_optin_smem(offline_fft_w_kernel<R>));
        offline_fft_w_kernel<R><<<grid, OFFW_THREADS, offw_fft_smem<R>(), st>>>(*f);
    } else {
        SC_CHECK(sc_optin_smem(offline_ifft_w_kernel<R>));
        offline_ifft_w_kernel<R><<<grid, OFFW_THREADS, offw_ifft_smem<R>(), st>>>(*i);
    }
    return (int)cudaGetLastError();
}

static int offw_dispatch(int M, int variant, const OffFftArgs* f, const OffIfftArgs* i, dim3 grid, cudaStream_t st)
{
    if (M == 1024 && variant == 2) {                           /* shuffle-free 32 x 32 version */
        if (f) {
            const size_t smem = (size_t)8 * WFFT_TILE * 8;
            SC_CHECK(sc_optin_smem(offline_fft_t_kernel));
            offline_fft_t_kernel<<<grid, OFFW_THREADS, smem, st>>>(*f);
        } else {
            const size_t smem = (size_t)1024 * 9 * 8;           /* >= 8 tiles */
            SC_CHECK(sc_optin_smem(offline_ifft_t_kernel));
            offline_ifft_t_kernel<<<grid, OFFW_THREADS, smem, st>>>(*i);
        }
        return (int)cudaGetLastError();
    }
    switch (M) {
        case 64:   return offw_launch<2>(f, i, grid, st);
        case 128:  return offw_launch<4>(f, i, grid, st);
        case 256:  return offw_launch<8>(f, i, grid, st);
        case 512:  return offw_launch<16>(f, i, grid, st);
        case 1024: return offw_launch<32>(f, i, grid, st);
        default:   return (int)cudaErrorInvalidValue;
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  C-ABI                                                                                        */
/* ------------------------------------------------------------------------------------------ */
extern "C" {

static size_t roundup(size_t v, size_t m) { return (v + m - 1) / m * m; }

int scdev_offline_free(scdev_offline* o)
{
    if (!o) return 0;
    cudaFree(o->XGhi); cudaFree(o->XGlo); cudaFree(o->HGhi); cudaFree(o->HGlo); cudaFree(o->Ys); cudaFree(o->scal); cudaFree(o->wtab);
    o->XGhi = o->XGlo = o->HGhi = o->HGlo = NULL; o->Ys = NULL; o->scal = NULL; o->wtab = NULL;
    o->capFrames = 0; o->packed = 0;
    return 0;
}

/* (re)allocate the workspace for T frames and build the filter operand once */
int scdev_offline_prepare(const scdev_plan* pl, const scdev_bufs* b, scdev_offline* o, int T, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->kind != SC_KIND_MATRIX) return (int)cudaErrorInvalidValue;
    int Nn = 32;                                                     /* UMMA N: 2*nOut padded to 32 / 64 / 128 */
    while (Nn < 2 * pl->nOutLocal) Nn <<= 1;
    if (Nn > 2 * OFF_MAX_HALF) Nn = 2 * OFF_MAX_HALF;                /* more than 64 outputs: tiles of 64 (grid.z of the GEMM) */
    o->nTiles = (2 * pl->nOutLocal + Nn - 1) / Nn;
    if (pl->P - 1 > 1024) return (int)cudaErrorInvalidValue;
    o->Nn = Nn;
    if (!o->packed) {                                                /* operand type is fixed per handle */
        const char* v = getenv("SAFCONV_OFF_KIND");
        o->f16 = !(v && !strcmp(v, "tf32"));
    }
    const int epk = o->f16 ? 8 : 4;                                  /* values per 16-byte k-group */
    o->ipc = epk / 2;
    o->Kp = (int)roundup((size_t)2 * pl->nIn, (size_t)epk * OFF_KG);
    o->nKG = o->Kp / epk;
    o->nKC = o->nKG / OFF_KG;
    o->rowsX = OFF_MT + pl->P - 1;
    o->tmemCols = 4 * Nn;                                            /* two buffer sets x two frame tiles */
    o->gemmSmem = 2 * (2 * OFF_KG * o->rowsX * 16) + OFF_NH * (2 * OFF_KG * Nn * 16) + (12 + 2 * OFF_NH) * 8 + 16;
    if (!o->smCount) {
        int dev = 0;
        SC_CHECK(cudaGetDevice(&dev));
        SC_CHECK(cudaDeviceGetAttribute(&o->smCount, cudaDevAttrMultiProcessorCount, dev));
    }
    {
        const char* v = getenv("SAFCONV_OFF_FLUSH");
        int fl = v ? atoi(v) : 1;
        o->flush = fl < 1 ? 1 : fl;
        v = getenv("SAFCONV_OFF_FPC");      o->fpc = v ? ((atoi(v) == 2) ? 2 : 4) : (o->f16 ? 2 : 4);
        v = getenv("SAFCONV_OFF_OPC");      o->opc = (v && atoi(v) == 4) ? 4 : 8;
        v = getenv("SAFCONV_OFF_THREADS");  o->fftThreads = (v && atoi(v) == 128) ? 128 : 256;
        /* warp-FFT transform kernels: fp16 operands, M = 64 .. 1024 (SAFCONV_OFF_WFFT=0 selects the shared-memory ones) */
        v = getenv("SAFCONV_OFF_WFFT");
        o->wfft = (o->f16 && pl->M >= 64 && pl->M <= 1024) ? (v ? atoi(v) : 2) : 0;     /* 2: shuffle-free version at M = 1024 */
        if (o->wfft && !o->wtab) {
            SC_CHECK(cudaMalloc((void**)&o->wtab, (size_t)2 * pl->M * sizeof(float2)));
            wfft_tables_kernel<<<(pl->M + 255) / 256, 256, 0, st>>>((const float2*)b->tw, (float2*)o->wtab, (float2*)o->wtab + pl->M,
                                                                     pl->M, pl->logM - 5);
            SC_CHECK(cudaGetLastError());
        }
        /* fewer transforms per CTA when the FFT work arrays of the default batch do not fit in shared memory */
        const size_t arr = (size_t)SC_ALEN(pl->M) * 8, cap = 227 * 1024;
        while (o->fpc > 1 && (size_t)(o->ipc * o->fpc + 1) * arr > cap) o->fpc >>= 1;
        if (o->opc == 8 && 9 * arr > cap) o->opc = 4;
        if ((size_t)(o->ipc * o->fpc + 1) * arr > cap || (size_t)(o->opc + 1) * arr > cap) return (int)cudaErrorInvalidValue;
    }
    if (!o->packed) {
        const size_t hgTile = (size_t)pl->M * pl->P * o->nKG * Nn * 16;
        const size_t hgBytes = hgTile * o->nTiles;
        SC_CHECK(cudaMalloc((void**)&o->HGhi, hgBytes));
        SC_CHECK(cudaMalloc((void**)&o->HGlo, hgBytes));
        SC_CHECK(cudaMalloc((void**)&o->scal, 4 * sizeof(float)));
        SC_CHECK(cudaMemsetAsync(o->HGhi, 0, hgBytes, st));
        SC_CHECK(cudaMemsetAsync(o->HGlo, 0, hgBytes, st));
        SC_CHECK(cudaMemsetAsync(o->scal, 0, 4 * sizeof(float), st));
        PackArgs a;
        a.H = (const float2*)b->H; a.HGhi = (unsigned char*)o->HGhi; a.HGlo = (unsigned char*)o->HGlo; a.scal = o->scal;
        a.total = (size_t)pl->nOT * pl->nKT * pl->P * pl->nIn * pl->OTsz * SC_BK;
        a.nKT = pl->nKT; a.P = pl->P; a.nIn = pl->nIn; a.OTsz = pl->OTsz; a.nOutLocal = pl->nOutLocal;
        a.nKG = o->nKG; a.Nn = Nn; a.tileBytes = hgTile;
        if (o->f16) {
            offline_absmax_kernel<<<148 * 8, 256, 0, st>>>((const float*)b->H, 2 * a.total, o->scal + 1);
            SC_CHECK(cudaGetLastError());
            offline_pack_filters_kernel<true><<<148 * 8, 256, 0, st>>>(a);
        } else {
            offline_pack_filters_kernel<false><<<148 * 8, 256, 0, st>>>(a);
        }
        SC_CHECK(cudaGetLastError());
        SC_CHECK(sc_optin_smem(offline_gemm_kernel<true, false>));
        SC_CHECK(sc_optin_smem(offline_gemm_kernel<false, false>));
        SC_CHECK(sc_optin_smem(offline_gemm_kernel<true, true>));
        SC_CHECK(sc_optin_smem(offline_gemm_kernel<false, true>));
        SC_CHECK(sc_optin_smem(offline_fft_kernel));
        SC_CHECK(sc_optin_smem(offline_ifft_kernel));
        o->packed = 1;
    }
    if (T > o->capFrames) {
        SC_CHECK(cudaStreamSynchronize(st));
        cudaFree(o->XGhi); cudaFree(o->XGlo); cudaFree(o->Ys);
        o->XGhi = o->XGlo = o->Ys = NULL; o->capFrames = 0;
        const size_t Tpad = roundup((size_t)T, OFF_MT);
        const size_t rowsAlloc = roundup(Tpad + pl->P - 1, OFF_FPC);
        const size_t xgBytes = (size_t)pl->M * o->nKG * rowsAlloc * 16;
        SC_CHECK(cudaMalloc((void**)&o->XGhi, xgBytes));
        SC_CHECK(cudaMalloc((void**)&o->XGlo, xgBytes));
        SC_CHECK(cudaMalloc((void**)&o->Ys, (size_t)o->nTiles * pl->M * Tpad * Nn * sizeof(float)));
        /* padding k-groups (odd nIn / K padding) are never written by the FFT kernel: zero them once */
        SC_CHECK(cudaMemsetAsync(o->XGhi, 0, xgBytes, st));
        SC_CHECK(cudaMemsetAsync(o->XGlo, 0, xgBytes, st));
        o->capFrames = T; o->capTpad = (int)Tpad; o->capRows = (int)rowsAlloc;
    }
    return 0;
}

/* d_in [nIn][T*hop] -> d_out [nOutLocal][(T-skip)*hop]; zero state before the first frame; the first `skip`
 * frames are a halo (history for the frames that follow) whose output is not written */
int scdev_offline_run(const scdev_plan* pl, const scdev_bufs* b, scdev_offline* o,
                      const float* d_in, float* d_out, int T, int skip, void** events, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (T < 1 || T > o->capFrames || skip < 0 || skip >= T) return (int)cudaErrorInvalidValue;
    const int Tpad = (int)roundup((size_t)T, OFF_MT);
    const int rowsAlloc = o->capRows;           /* row stride of the operand as allocated */
    /* operand rows that hold this render's frames (row = P-1 + frame); rows up to the end of the last 256-frame tile
     * keep whatever an earlier render left there: they only reach GEMM columns >= T, which are never read */
    const int rowsUsed = (int)roundup((size_t)T + pl->P - 1, OFF_FPC);

    if (events) SC_CHECK(cudaEventRecord((cudaEvent_t)events[0], st));
    if (o->f16) {
        /* bound on the input spectra of this render: the largest l1 norm of a block (see pow2_scale) */
        SC_CHECK(cudaMemsetAsync(o->scal, 0, sizeof(float), st));
        dim3 grid((T + 7) / 8, pl->nIn);
        offline_l1max_kernel<<<grid, 256, 0, st>>>(d_in, (size_t)T * pl->hop, pl->hop, T, o->scal);
        SC_CHECK(cudaGetLastError());
    }
    OffFftArgs f;
    f.in = d_in; f.XGhi = (unsigned char*)o->XGhi; f.XGlo = (unsigned char*)o->XGlo; f.tw = (const float2*)b->tw; f.scal = o->scal;
    f.inStride = (size_t)T * pl->hop;
    f.hop = pl->hop; f.nIn = pl->nIn; f.M = pl->M; f.logM = pl->logM; f.P = pl->P; f.T = T;
    f.rowsAlloc = rowsAlloc; f.nKG = o->nKG;
    f.wT1 = (const float2*)o->wtab; f.wT2 = o->wtab ? (const float2*)o->wtab + pl->M : NULL;
    {
        f.fpc = o->fpc; f.ipc = o->ipc;
        if (o->wfft) {
            dim3 grid(rowsUsed / 2, (pl->nIn + 3) / 4);
            { const int e_ = offw_dispatch(pl->M, o->wfft, &f, NULL, grid, st); if (e_) return e_; }
        } else {
            dim3 grid(rowsUsed / o->fpc, (pl->nIn + o->ipc - 1) / o->ipc);
            offline_fft_kernel<<<grid, o->fftThreads, (size_t)(o->ipc * o->fpc + 1) * SC_ALEN(pl->M) * 8, st>>>(f);
            SC_CHECK(cudaGetLastError());
        }
    }
    if (events) SC_CHECK(cudaEventRecord((cudaEvent_t)events[1], st));
    OffGemmArgs g;
    g.XGhi = (const unsigned char*)o->XGhi; g.XGlo = (const unsigned char*)o->XGlo;
    g.HGhi = (const unsigned char*)o->HGhi; g.HGlo = (const unsigned char*)o->HGlo; g.Ys = o->Ys; g.scal = o->scal;
    g.P = pl->P; g.nKG = o->nKG; g.nKC = o->nKC; g.Nn = o->Nn; g.rowsAlloc = rowsAlloc; g.Tpad = o->capTpad;
    g.rowsX = o->rowsX; g.tmemCols = o->tmemCols; g.flush = o->flush;
    g.T = T; g.nFT = (T + OFF_MT - 1) / OFF_MT; g.nBins = pl->M; g.nTilesTotal = g.nFT * pl->M * o->nTiles;
    g.ticket = reinterpret_cast<int*>(o->scal + 2);
    SC_CHECK(cudaMemsetAsync(g.ticket, 0, sizeof(int), st));
    g.hgTileBytes = (size_t)pl->M * pl->P * o->nKG * o->Nn * 16;
    g.ysTileFloats = (size_t)pl->M * o->capTpad * o->Nn;
    /* frames on the UMMA N axis (one M128 x N256 instruction per product term) when the filter rows fill M = 128 */
    static int ntEnv = -1;
    if (ntEnv < 0) { const char* v = getenv("SAFCONV_OFF_NT"); ntEnv = v ? atoi(v) : 1; }
    const int nt = ntEnv && o->Nn == 128;
    const int mm = 128, nn = nt ? 0 : o->Nn;
    g.idesc = o->f16 ? umma_idesc_f16(mm, nn) : umma_idesc_tf32(mm, nn);      /* NT: the N field is filled in per tile */
    {
        static int ctasEnv = -1;
        if (ctasEnv < 0) { const char* v = getenv("SAFCONV_OFF_CTAS"); ctasEnv = v ? atoi(v) : 0; }
        int ctas = ctasEnv > 0 ? ctasEnv : o->smCount;
        if (ctas > g.nTilesTotal) ctas = g.nTilesTotal;
        dim3 grid(ctas);
        if (o->f16) { if (nt) offline_gemm_kernel<true, true><<<grid, OFF_GEMM_THREADS, o->gemmSmem, st>>>(g);
                      else    offline_gemm_kernel<true, false><<<grid, OFF_GEMM_THREADS, o->gemmSmem, st>>>(g); }
        else        { if (nt) offline_gemm_kernel<false, true><<<grid, OFF_GEMM_THREADS, o->gemmSmem, st>>>(g);
                      else    offline_gemm_kernel<false, false><<<grid, OFF_GEMM_THREADS, o->gemmSmem, st>>>(g); }
        SC_CHECK(cudaGetLastError());
    }
    if (events) SC_CHECK(cudaEventRecord((cudaEvent_t)events[2], st));
    OffIfftArgs i;
    i.Ys = (const float2*)o->Ys; i.out = d_out; i.tw = (const float2*)b->tw;
    i.hop = pl->hop; i.M = pl->M; i.logM = pl->logM; i.nOut = pl->nOutLocal; i.Nn2 = o->Nn / 2;
    i.Tpad = o->capTpad; i.T = T; i.skip = skip; i.scale = 1.0f / (float)pl->N;
    i.ysTileC = (size_t)pl->M * o->capTpad * (o->Nn / 2);
    i.logNn2 = 0; while ((1 << i.logNn2) < o->Nn / 2) i.logNn2++;
    i.wT1 = (const float2*)o->wtab;
    {
        i.opc = o->opc;
        SC_CHECK(cudaMemsetAsync(d_out, 0, sizeof(float) * (size_t)pl->nOutLocal * (T - skip) * pl->hop, st));
        if (o->wfft) {
            dim3 grid((pl->nOutLocal + 7) / 8, T);
            { const int e_ = offw_dispatch(pl->M, o->wfft, NULL, &i, grid, st); if (e_) return e_; }
        } else {
            dim3 grid((pl->nOutLocal + o->opc - 1) / o->opc, T);
            offline_ifft_kernel<<<grid, o->fftThreads, (size_t)(o->opc + 1) * SC_ALEN(pl->M) * 8, st>>>(i);
            SC_CHECK(cudaGetLastError());
        }
    }
    if (events) SC_CHECK(cudaEventRecord((cudaEvent_t)events[3], st));
    return 0;
}

} /* extern "C" */
