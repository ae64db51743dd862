/*
 * safconv_fft.cuh -- device-side building blocks shared by the CUDA translation units of libsafconv_b200:
 * complex helpers, the shared-memory / warp-shuffle FFT core, real-FFT split passes, overlap-add epilogue,
 * and the mbarrier / TMA-bulk-copy PTX wrappers.
 */
#ifndef SAFCONV_FFT_CUH_INCLUDED
#define SAFCONV_FFT_CUH_INCLUDED

#include <cuda_runtime.h>
#include <stdint.h>

#define SC_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

/* Opt-in to more than 48 KB of dynamic shared memory.  cudaFuncAttributeMaxDynamicSharedMemorySize is a LIMIT that is
 * global per (function, device): a plan-specific value set by one handle would be lowered by the next handle that
 * shares the kernel instantiation, and the first handle's launches would then fail.  So the limit is only ever set
 * to the device's opt-in maximum (227 KB on sm_100) -- monotone by construction, whatever handles are alive. */
template <class F>
static inline cudaError_t sc_optin_smem(F fn)
{
    static int optin[64];                           /* per device, 0 = not queried yet */
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int mx = (dev >= 0 && dev < 64) ? optin[dev] : 0;
    if (!mx) {
        e = cudaDeviceGetAttribute(&mx, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) optin[dev] = mx;
    }
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
}

/* ------------------------------------------------------------------------------------------ */
/*  small device helpers                                                                      */
/* ------------------------------------------------------------------------------------------ */

__device__ __forceinline__ float2 cmulf(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conjb(float2 a, float2 b)   /* a * conj(b) */
{
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ int bitrev(int v, int logM) { return (int)(__brev((unsigned)v) >> (32 - logM)); }

/* FFT work arrays in shared memory are PADDED: element i of an M-point array lives at
 *     padi(i) = i + (i >> 4) + (i >> (logM-4))
 * (one extra float2 after every 16 elements and one after every M/16; array length SC_ALEN(M)).
 *  - the register passes let a thread own 2..16 CONSECUTIVE elements (stride-1 final pass) or elements at a
 *    small stride: the per-16 skew puts the 16 lanes of a half-warp on 16 different banks in every case;
 *  - the FFT leaves its result in bit-reversed order and the epilogues read s[bitrev(k)] for consecutive k,
 *    i.e. lane stride M/32 elements: the two skews together make those reads conflict-free as well;
 *  - SC_ALEN(M) = 2 (mod 16) elements, so consecutive arrays of a batch start on different banks. */
#define SC_ALEN(M) ((M) + ((M) >> 4) + 18)
__device__ __forceinline__ int padi(int i, int logM) { return i + (i >> 4) + (i >> (logM - 4)); }

/* ------------------------------------------------------------------------------------------ */
/*  M-point complex FFT in shared memory, decimation in frequency, in place                     */
/*  input: natural order (padded indexing) ; output: element padi(bitrev(k)) holds bin k          */
/*                                                                                              */
/*  Register passes: a pass of radix R = 16 on sub-transforms of length L lets one thread own the  */
/*  16 elements j + i*(L/16) of a sub-transform, run the complete 16-point DIF butterfly (four      */
/*  radix-2 stages, internal twiddles W_16^k are compile-time constants) in registers, multiply     */
/*  output q by the external twiddle W_L^(j*q) and store in place: four FFT stages per trip         */
/*  through shared memory.  log2(M) mod 4 left-over stages run as a final radix-2/4/8 register      */
/*  pass over consecutive elements.  All transforms of a CTA advance together (one barrier per pass).*/
/*  tw = the per-pass tables load_twiddles() builds in shared memory: for every radix-16 pass with   */
/*  L > 16, W_L^(j*q) at [(q-1)*(L/16) + j], q = 1..15 (unit stride in j).                           */
/*  INV conjugates every twiddle (unnormalised inverse transform).                                */
/*  Requires M >= 32; ends with __syncthreads().                                                  */
/* ------------------------------------------------------------------------------------------ */

/* d * W_16^k  (forward: exp(-2 pi i k / 16); INV: the conjugate), k a compile-time constant after unrolling */
template <bool INV>
__device__ __forceinline__ float2 mul_w16(float2 d, int k)
{
    const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
    if (k == 0) return d;
    if (k == 4) return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    float wx, wy;                                   /* W_16^k = (wx, -wy) forward */
    switch (k) {
        case 1: wx = C1;  wy = S1; break;
        case 2: wx = R2;  wy = R2; break;
        case 3: wx = S1;  wy = C1; break;
        case 5: wx = -S1; wy = C1; break;
        case 6: wx = -R2; wy = R2; break;
        default: wx = -C1; wy = S1; break;          /* 7 */
    }
    if (INV) wy = -wy;
    /* d * (wx - i wy) */
    return make_float2(d.x * wx + d.y * wy, d.y * wx - d.x * wy);
}

/* complete R-point DIF butterfly on registers; v[i] ends up holding output bitrev_R(i), exactly like the
 * in-place radix-2 DIF would leave it at position i */
template <int R, bool INV>
__device__ __forceinline__ void dif_regs(float2 (&v)[R])
{
#pragma unroll
    for (int h = R / 2; h >= 1; h >>= 1) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if ((i & h) == 0) {
                const float2 a = v[i], b = v[i + h];
                v[i] = caddf(a, b);
                v[i + h] = mul_w16<INV>(csubf(a, b), (i & (h - 1)) * (8 / h));
            }
        }
    }
}

template <int R, bool INV>
__device__ __forceinline__ void final_pass(float2* s, const int M, const int logM, const int nArr)
{
    const int per = M / R;                          /* threads' worth of work per transform */
    for (int it = threadIdx.x; it < nArr * per; it += blockDim.x) {
        const int arr = it / per, t = it - arr * per;
        float2* sa = s + (size_t)arr * SC_ALEN(M);
        float2 v[R];
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = sa[padi(t * R + i, logM)];
        dif_regs<R, INV>(v);
#pragma unroll
        for (int i = 0; i < R; ++i) sa[padi(t * R + i, logM)] = v[i];
    }
    __syncthreads();
}

template <bool INV>
__device__ void cfft_dif_batch_wide(float2* s, const int M, const int logM, const float2* __restrict__ tw, const int nArr)
{
    const int tid = threadIdx.x, T = blockDim.x;
    int logL = logM, off = 0;
    while (logL >= 4) {
        const int logSt = logL - 4, st = 1 << logSt;                /* element stride inside a butterfly */
        const int per = M >> 4;                                     /* butterflies per transform */
        for (int it = tid; it < nArr * per; it += T) {
            const int arr = it >> (logM - 4), t = it & (per - 1);
            float2* sa = s + (size_t)arr * SC_ALEN(M);
            const int j = t & (st - 1);
            const int base = ((t >> logSt) << logL) + j;            /* sub-transform start + j */
            float2 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = sa[padi(base + (i << logSt), logM)];
            dif_regs<16, INV>(v);
            if (logSt > 0) {
#pragma unroll
                for (int i = 1; i < 16; ++i) {
                    const int q = ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3);   /* bitrev4(i) */
                    float2 w = tw[off + (q - 1) * st + j];
                    if (INV) w.y = -w.y;
                    v[i] = cmulf(v[i], w);
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) sa[padi(base + (i << logSt), logM)] = v[i];
        }
        __syncthreads();
        if (logSt > 0) off += 15 * st;
        logL -= 4;
    }
    if (logL == 1) final_pass<2, INV>(s, M, logM, nArr);
    else if (logL == 2) final_pass<4, INV>(s, M, logM, nArr);
    else if (logL == 3) final_pass<8, INV>(s, M, logM, nArr);
}

/* ---- "narrow" core: radix-4 / radix-2 passes through shared memory (4 points per thread per pass) and the last
 * five stages as warp-shuffle butterflies.  More barriers and more instructions per point than the register
 * radix-16 passes, but M/4 threads work on one transform -- the better choice when a CTA owns a single transform
 * (the block-latency kernels K1 / K3 / multiConv / TVConv).  Its per-pass tables: W_L^j then W_L^2j for every
 * radix-4 pass, W_L^j for the radix-2 pass, W_32^j (j < 16) for the shuffle stages. ---- */
template <bool INV>
__device__ __forceinline__ float2 twd(const float2* __restrict__ tw, int idx)
{
    float2 w = tw[idx];
    if (INV) w.y = -w.y;
    return w;
}

template <bool INV>
__device__ void cfft_dif_batch_narrow(float2* s, const int M, const int logM, const float2* __restrict__ tw, const int nArr)
{
    const int tid = threadIdx.x, T = blockDim.x;
    int L = M;                 /* current sub-transform length */
    int nsm = logM - 5;        /* radix-2 stages done through shared memory (spans M/2 .. 32) */
    int off = 0;               /* start of this pass's twiddle table inside tw (see load_twiddles) */

    /* two radix-2 stages fused per pass */
    while (nsm >= 2) {
        const int q = L >> 2;
        const int per = M >> 2;                    /* butterflies per transform */
        for (int it = tid; it < nArr * per; it += T) {
            const int arr = it >> (logM - 2), i = it & (per - 1);
            float2* sa = s + (size_t)arr * SC_ALEN(M);
            const int j = i & (q - 1);
            const int base = ((i - j) << 2) + j;
            const int i0 = padi(base, logM), i1 = padi(base + q, logM), i2 = padi(base + 2 * q, logM), i3 = padi(base + 3 * q, logM);
            const float2 a0 = sa[i0], a1 = sa[i1], a2 = sa[i2], a3 = sa[i3];
            const float2 w1 = twd<INV>(tw, off + j);         /* W_L^j  */
            const float2 w2 = twd<INV>(tw, off + q + j);     /* W_L^2j */
            const float2 u0 = caddf(a0, a2);
            const float2 u1 = caddf(a1, a3);
            const float2 v0 = cmulf(csubf(a0, a2), w1);
            float2 d1 = csubf(a1, a3);
            /* W_L^(j+L/4) = W_L^j * (-i) forward, * (+i) inverse */
            d1 = INV ? make_float2(-d1.y, d1.x) : make_float2(d1.y, -d1.x);
            const float2 v1 = cmulf(d1, w1);
            sa[i0] = caddf(u0, u1);
            sa[i1] = cmulf(csubf(u0, u1), w2);
            sa[i2] = caddf(v0, v1);
            sa[i3] = cmulf(csubf(v0, v1), w2);
        }
        __syncthreads();
        off += 2 * q;
        L >>= 2;
        nsm -= 2;
    }
    if (nsm == 1) {
        const int half = L >> 1;
        const int per = M >> 1;
        for (int it = tid; it < nArr * per; it += T) {
            const int arr = it >> (logM - 1), i = it & (per - 1);
            float2* sa = s + (size_t)arr * SC_ALEN(M);
            const int j = i & (half - 1);
            const int base = ((i - j) << 1) + j;
            const int i0 = padi(base, logM), i1 = padi(base + half, logM);
            const float2 a = sa[i0], b = sa[i1];
            const float2 w = twd<INV>(tw, off + j);
            sa[i0] = caddf(a, b);
            sa[i1] = cmulf(csubf(a, b), w);
        }
        __syncthreads();
        off += half;
        L >>= 1;
    }
    /* L == 32: the last five stages (spans 16,8,4,2,1) stay inside one warp; rows of 32 points of all
     * transforms are contiguous, so the batch is just more rows.
     * Branch-free butterflies: lane l (partner l^h) computes  t = o + sg*v  (lower half: v+o, upper half: o-v)
     * and multiplies by its own per-lane factor (1 in the lower half, the twiddle in the upper half). */
    {
        const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
        const float2 one = make_float2(1.f, 0.f);
        /* tw + off : the 16-entry table W_32^j, j < 16 */
        const float2 w16 = (lane & 16) ? twd<INV>(tw, off + (lane & 15))     : one;
        const float2 w8  = (lane & 8)  ? twd<INV>(tw, off + 2 * (lane & 7)) : one;
        const float2 w4  = (lane & 4)  ? twd<INV>(tw, off + 4 * (lane & 3)) : one;
        const float2 w2  = (lane & 2)  ? twd<INV>(tw, off + 8 * (lane & 1)) : one;
        const float s16 = (lane & 16) ? -1.f : 1.f, s8 = (lane & 8) ? -1.f : 1.f, s4 = (lane & 4) ? -1.f : 1.f,
                    s2 = (lane & 2) ? -1.f : 1.f, s1 = (lane & 1) ? -1.f : 1.f;
        for (int row = warp; row < nArr * (M >> 5); row += nwarps) {
            const int arr = row >> (logM - 5), rr = row & ((M >> 5) - 1);
            float2* sp = s + (size_t)arr * SC_ALEN(M) + padi(rr * 32 + lane, logM);
            float2 v = *sp;
#define SC_SHFL_STAGE(HALF, SG, W)                                                   \
            {                                                                        \
                const float ox = __shfl_xor_sync(0xffffffffu, v.x, HALF);            \
                const float oy = __shfl_xor_sync(0xffffffffu, v.y, HALF);            \
                const float tx = fmaf(SG, v.x, ox), ty = fmaf(SG, v.y, oy);          \
                v.x = tx * W.x - ty * W.y;                                           \
                v.y = tx * W.y + ty * W.x;                                           \
            }
            SC_SHFL_STAGE(16, s16, w16)
            SC_SHFL_STAGE(8,  s8,  w8)
            SC_SHFL_STAGE(4,  s4,  w4)
            SC_SHFL_STAGE(2,  s2,  w2)
#undef SC_SHFL_STAGE
            {   /* span 1: twiddle is 1 */
                const float ox = __shfl_xor_sync(0xffffffffu, v.x, 1);
                const float oy = __shfl_xor_sync(0xffffffffu, v.y, 1);
                v.x = fmaf(s1, v.x, ox);
                v.y = fmaf(s1, v.y, oy);
            }
            *sp = v;
        }
        __syncthreads();
    }
}


/* which core a CTA uses: register radix-16 passes when the batch gives every thread at least one 16-point
 * butterfly per pass, else the narrow core.  A kernel decides ONCE (nArr = the smallest batch it transforms) and
 * hands the same `wide` to load_twiddles() and to every cfft_dif*() call: the two cores use different tables. */
__device__ __forceinline__ bool fft_use_wide(int M, int nArr) { return nArr * (M >> 4) >= (int)blockDim.x; }

template <bool INV>
__device__ __forceinline__ void cfft_dif_batch(float2* s, const int M, const int logM, const float2* __restrict__ tw,
                                               const int nArr, const bool wide)
{
    if (wide) cfft_dif_batch_wide<INV>(s, M, logM, tw, nArr);
    else      cfft_dif_batch_narrow<INV>(s, M, logM, tw, nArr);
}

template <bool INV>
__device__ __forceinline__ void cfft_dif(float2* s, const int M, const int logM, const float2* __restrict__ tw, const bool wide)
{
    cfft_dif_batch<INV>(s, M, logM, tw, 1, wide);
}

/* Build the per-pass twiddle tables of cfft_dif_batch in shared memory (fewer than M float2 in total): for every
 * radix-16 pass of sub-length L > 16 the table W_L^(j*q) at [(q-1)*(L/16) + j], q = 1..15, j < L/16, taken from the
 * global table gtw[e] = exp(-2 pi i e / 2M), e < M  (W_L^x = W_2M^(x * 2M/L); the upper half by W^(M+e) = -W^e). */
__device__ __forceinline__ void load_twiddles_wide(float2* stw, const float2* __restrict__ gtw, int M, int logM)
{
    int logL = logM, off = 0;
    while (logL >= 4) {
        const int logSt = logL - 4, st = 1 << logSt;
        if (logSt > 0) {
            const int sh = logM + 1 - logL;                              /* 2M / L = 2^sh */
            for (int idx = threadIdx.x; idx < 15 * st; idx += blockDim.x) {
                const int q = (idx >> logSt) + 1, j = idx & (st - 1);
                const int e = (j * q) << sh;                             /* < 2M */
                float2 w = __ldg(gtw + (e & (M - 1)));
                if (e >= M) { w.x = -w.x; w.y = -w.y; }
                stw[off + idx] = w;
            }
            off += 15 * st;
        }
        logL -= 4;
    }
}

__device__ __forceinline__ void load_twiddles_narrow(float2* stw, const float2* __restrict__ gtw, int M, int logM)
{
    const int tid = threadIdx.x, T = blockDim.x;
    int L = M, nsm = logM - 5, off = 0;
    while (nsm >= 2) {
        const int q = L >> 2, tstr = (2 * M) / L;
        for (int j = tid; j < q; j += T) {
            stw[off + j]     = __ldg(gtw + j * tstr);
            stw[off + q + j] = __ldg(gtw + 2 * j * tstr);
        }
        off += 2 * q; L >>= 2; nsm -= 2;
    }
    if (nsm == 1) {
        const int half = L >> 1, tstr = (2 * M) / L;
        for (int j = tid; j < half; j += T) stw[off + j] = __ldg(gtw + j * tstr);
        off += half;
    }
    if (tid < 16) stw[off + tid] = __ldg(gtw + tid * (M >> 4));
}


__device__ __forceinline__ void load_twiddles(float2* stw, const float2* __restrict__ gtw, int M, int logM, bool wide)
{
    if (wide) load_twiddles_wide(stw, gtw, M, logM);
    else      load_twiddles_narrow(stw, gtw, M, logM);
}

/* The real-FFT split passes need W_N^k, k <= M/2 (unit stride).  For M <= 4096 the kernels keep a copy in shared
 * memory right after the M-entry pass-table region (one L2 round trip less in the middle of the kernel); for
 * larger M they read the global table. */
__host__ __device__ __forceinline__ int sc_split_len(int M) { return M <= 4096 ? (M / 2 + 2) : 0; }
__device__ __forceinline__ const float2* load_split_twiddles(float2* stw, const float2* __restrict__ gtw, int M)
{
    if (!sc_split_len(M)) return gtw;
    float2* spl = stw + M;
    for (int k = threadIdx.x; k <= (M >> 1); k += blockDim.x) spl[k] = __ldg(gtw + k);
    return spl;
}

/* load one real block of `hop` samples (zero-padded to N = 2M) as M complex values z[n] = x[2n] + i x[2n+1] */
__device__ __forceinline__ void load_real_block(float2* s, const float* __restrict__ x, int hop, int M, int logM)
{
    const int tid = threadIdx.x, T = blockDim.x;
    if ((hop & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
        /* 16-byte loads: two complex points (four samples) per thread and load */
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const int h4 = hop >> 2;
        for (int n2 = tid; n2 < (M >> 1); n2 += T) {
            const float4 v = (n2 < h4) ? __ldg(x4 + n2) : make_float4(0.f, 0.f, 0.f, 0.f);
            s[padi(2 * n2, logM)]     = make_float2(v.x, v.y);
            s[padi(2 * n2 + 1, logM)] = make_float2(v.z, v.w);
        }
    } else if ((hop & 1) == 0 && ((reinterpret_cast<uintptr_t>(x) & 7) == 0)) {
        const float2* x2 = reinterpret_cast<const float2*>(x);
        const int h2 = hop >> 1;
        for (int n = tid; n < M; n += T) s[padi(n, logM)] = (n < h2) ? __ldg(x2 + n) : make_float2(0.f, 0.f);
    } else {
        for (int n = tid; n < M; n += T) {
            const int i = 2 * n;
            float2 v;
            v.x = (i < hop) ? __ldg(x + i) : 0.f;
            v.y = (i + 1 < hop) ? __ldg(x + i + 1) : 0.f;
            s[padi(n, logM)] = v;
        }
    }
}

/* forward split pass for the bin pair (k, M-k), 1 <= k <= M/2, from the bit-reversed complex FFT in s.
 * X[k] = E + W_N^k O,  X[M-k] = conj(E - W_N^k O),  E = (a + conj b)/2, O = -i (a - conj b)/2 */
__device__ __forceinline__ void fwd_split_pair(const float2* s, int k, int M, int logM,
                                               const float2* __restrict__ tw, float2& Xk, float2& Xmk)
{
    const float2 a = s[padi(bitrev(k, logM), logM)];
    const float2 b = s[padi(bitrev(M - k, logM), logM)];
    const float2 E = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
    const float2 O = make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x));
    const float2 t = cmulf(tw[k], O);                /* tw: split table W_N^k (shared copy or the global table) */
    Xk  = make_float2(E.x + t.x, E.y + t.y);
    Xmk = make_float2(E.x - t.x, t.y - E.y);
}

/* inverse split pass, in place on the packed natural-order spectrum Z (pair k, M-k; 1 <= k <= M/2):
 * Zc[k] = E + iO, Zc[M-k] = conj(E) + i conj(O), E = A + conj B, O = (A - conj B) W_N^-k  (the 1/2 is folded into 1/N) */
__device__ __forceinline__ void inv_split_pair(float2* Z, int k, int M, int logM, const float2* __restrict__ tw)
{
    const int ik = padi(k, logM), im = padi(M - k, logM);
    const float2 A = Z[ik], B = Z[im];
    const float2 E = make_float2(A.x + B.x, A.y - B.y);
    const float2 D = make_float2(A.x - B.x, A.y + B.y);
    const float2 O = cmul_conjb(D, tw[k]);           /* tw: split table W_N^k (shared copy or the global table) */
    Z[ik] = make_float2(E.x - O.y, E.y + O.x);
    Z[im] = make_float2(E.x + O.y, O.x - E.y);
}

/* inverse split pass of nArr packed spectra stored back to back */
__device__ __forceinline__ void inv_split_batch(float2* Z, int M, int logM, const float2* __restrict__ tw, int nArr)
{
    const int per = (M >> 1) + 1;
    for (int it = threadIdx.x; it < nArr * per; it += blockDim.x) {
        const int arr = it / per, k = it - arr * per;
        float2* Za = Z + (size_t)arr * SC_ALEN(M);
        if (k == 0) {
            const float2 A = Za[0];                      /* (DC, Nyquist) */
            Za[0] = make_float2(A.x + A.y, A.x - A.y);
        } else {
            inv_split_pair(Za, k, M, logM, tw);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void inv_split_all(float2* Z, int M, int logM, const float2* __restrict__ tw)
{
    inv_split_batch(Z, M, logM, tw, 1);
}

/* time sample j of the (bit-reversed) inverse transform result */
__device__ __forceinline__ float time_sample(const float2* s, int j, int logM)
{
    const float2 v = s[padi(bitrev(j >> 1, logM), logM)];
    return (j & 1) ? v.y : v.x;
}

/* overlap-add epilogue (reference .c:230-233): out[i] = z[i]/N + tail[i]; tail[i] = z[i+hop]/N */
__device__ __forceinline__ void ola_store(const float2* s, int hop, int logM, float scale,
                                          float* __restrict__ out, float* __restrict__ tail)
{
    for (int i = threadIdx.x; i < hop; i += blockDim.x) {
        const float z0 = time_sample(s, i, logM) * scale;
        const float z1 = time_sample(s, i + hop, logM) * scale;
        out[i]  = z0 + tail[i];
        tail[i] = z1;
    }
}

/* "last CTA increments the block counter" : counters[0] = block counter, counters[1] = ticket */
__device__ __forceinline__ void advance_block_counter(unsigned int* counters, unsigned int nCtas)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&counters[1], 1u);
        if (t == nCtas - 1) {
            counters[1] = 0;
            __threadfence();
            atomicAdd(&counters[0], 1u);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/*  mbarrier / TMA bulk copy                                                                    */
/* ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
/* TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}


#endif /* SAFCONV_FFT_CUH_INCLUDED */
