/*
 * safconv_wfft.cuh -- warp-level register FFT: one warp transforms M = 32*R complex points (R = 2..32, a
 * compile-time constant) without shared memory and without block barriers.
 *
 * Lane j holds v[i] = x[j + 32 i], i < R.  Decimation in frequency in three steps:
 *   1. an R-point DIF FFT over i in registers (all twiddles are compile-time constants W_32^e),
 *   2. the twiddle W_M^(j k2) (k2 = bitrev_R(i)) from a table laid out [slot][lane] (unit stride over lanes),
 *   3. a 32-point DIF FFT across the lanes of every register slot: five branch-free __shfl_xor butterflies.
 * Result: lane l, slot i holds X[k],  k = bitrev_R(i) + R * bitrev_5(l).
 * With everything indexed by compile-time R the address arithmetic of the shared-memory cores (safconv_fft.cuh)
 * disappears: ~80 instructions per point instead of 145-200 (ncu, profiles/r01_offline_c5_full_summary.csv).
 *
 * The real-FFT split passes pair bin k with M-k: in this layout the partner of (lane l, slot i) sits in lane
 * l ^ 31 (k2 != 0; slot bitrev_R(R - k2), a compile-time constant) or in lane bitrev_5(32 - bitrev_5(l))
 * (k2 = 0, slot 0), so the forward split is two more shuffles per point.
 */
#ifndef SAFCONV_WFFT_CUH_INCLUDED
#define SAFCONV_WFFT_CUH_INCLUDED

#include "safconv_fft.cuh"

__host__ __device__ constexpr int wf_log2(int r) { return r <= 1 ? 0 : 1 + wf_log2(r >> 1); }
__host__ __device__ constexpr int wf_bitrev(int x, int bits)
{
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((x >> b) & 1) << (bits - 1 - b);
    return r;
}

/* d * W_32^e (forward: exp(-2 pi i e / 32); INV: the conjugate), e = 0..15 a compile-time constant after unrolling */
template <bool INV>
__device__ __forceinline__ float2 mul_w32(float2 d, int e)
{
    if (e == 0) return d;
    if (e == 8) return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    float c, s;
    switch (e) {
        case 1:  c =  0.98078528040323044913f; s = 0.19509032201612826785f; break;
        case 2:  c =  0.92387953251128675613f; s = 0.38268343236508977173f; break;
        case 3:  c =  0.83146961230254523708f; s = 0.55557023301960222474f; break;
        case 4:  c =  0.70710678118654752440f; s = 0.70710678118654752440f; break;
        case 5:  c =  0.55557023301960222474f; s = 0.83146961230254523708f; break;
        case 6:  c =  0.38268343236508977173f; s = 0.92387953251128675613f; break;
        case 7:  c =  0.19509032201612826785f; s = 0.98078528040323044913f; break;
        case 9:  c = -0.19509032201612826785f; s = 0.98078528040323044913f; break;
        case 10: c = -0.38268343236508977173f; s = 0.92387953251128675613f; break;
        case 11: c = -0.55557023301960222474f; s = 0.83146961230254523708f; break;
        case 12: c = -0.70710678118654752440f; s = 0.70710678118654752440f; break;
        case 13: c = -0.83146961230254523708f; s = 0.55557023301960222474f; break;
        case 14: c = -0.92387953251128675613f; s = 0.38268343236508977173f; break;
        default: c = -0.98078528040323044913f; s = 0.19509032201612826785f; break;   /* 15 */
    }
    if (INV) s = -s;
    return make_float2(d.x * c + d.y * s, d.y * c - d.x * s);      /* d * (c - i s) */
}

/* R-point DIF FFT on registers (R <= 32); v[i] ends up holding output bitrev_R(i) */
template <int R, bool INV>
__device__ __forceinline__ void dif_regs32(float2 (&v)[R])
{
#pragma unroll
    for (int h = R / 2; h >= 1; h >>= 1) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if ((i & h) == 0) {
                const float2 a = v[i], b = v[i + h];
                v[i] = caddf(a, b);
                v[i + h] = mul_w32<INV>(csubf(a, b), (i & (h - 1)) * (16 / h));
            }
        }
    }
}

/* the same for an input whose upper half v[16..31] is zero (a zero-padded block: hop <= M real samples = M/2 complex
 * points): the first stage degenerates to v[i+16] = v[i] * W_32^i -- no additions */
template <bool INV>
__device__ __forceinline__ void dif_regs32_zpad(float2 (&v)[32])
{
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i + 16] = mul_w32<INV>(v[i], i);
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if ((i & h) == 0) {
                const float2 a = v[i], b = v[i + h];
                v[i] = caddf(a, b);
                v[i + h] = mul_w32<INV>(csubf(a, b), (i & (h - 1)) * (16 / h));
            }
        }
    }
}

/* per-lane constants of the five shuffle stages */
struct WfftLane {
    float2 w16, w8, w4, w2;
    float  s16, s8, s4, s2, s1;
};

/* gtw[e] = exp(-2 pi i e / 2M), e < M  ->  W_32^e = gtw[e * M / 16]  (M >= 16) */
template <bool INV>
__device__ __forceinline__ WfftLane wfft_lane_init(const float2* __restrict__ gtw, int M, int lane)
{
    WfftLane L;
    const float2 one = make_float2(1.f, 0.f);
    const int u = M >> 4;
    L.w16 = (lane & 16) ? __ldg(gtw + (lane & 15) * u)       : one;
    L.w8  = (lane & 8)  ? __ldg(gtw + 2 * (lane & 7) * u)    : one;
    L.w4  = (lane & 4)  ? __ldg(gtw + 4 * (lane & 3) * u)    : one;
    L.w2  = (lane & 2)  ? __ldg(gtw + 8 * (lane & 1) * u)    : one;
    if (INV) { L.w16.y = -L.w16.y; L.w8.y = -L.w8.y; L.w4.y = -L.w4.y; L.w2.y = -L.w2.y; }
    L.s16 = (lane & 16) ? -1.f : 1.f;  L.s8 = (lane & 8) ? -1.f : 1.f;  L.s4 = (lane & 4) ? -1.f : 1.f;
    L.s2  = (lane & 2)  ? -1.f : 1.f;  L.s1 = (lane & 1) ? -1.f : 1.f;
    return L;
}

/* step-2 table T1[i * 32 + l] = W_M^(l * bitrev_R(i)), M = 32 R; built by the whole CTA (caller synchronises) */
template <int R>
__device__ __forceinline__ void wfft_build_table(float2* T1, const float2* __restrict__ gtw)
{
    constexpr int M = 32 * R, LOGR = wf_log2(R);
    for (int idx = threadIdx.x; idx < M; idx += blockDim.x) {
        const int i = idx >> 5, l = idx & 31;
        const int k2 = (LOGR == 0) ? 0 : (int)(__brev((unsigned)i) >> (32 - LOGR));
        const int e = 2 * l * k2;                               /* W_M^x = W_2M^(2x), e < 2M */
        float2 w = __ldg(gtw + (e & (M - 1)));
        if (e >= M) { w.x = -w.x; w.y = -w.y; }
        T1[idx] = w;
    }
}

/* one butterfly stage across lanes: lane l (partner l ^ HALF) gets  (o + sg v) * w  -- v + o in the lower half
 * (w = 1), (o - v) * twiddle in the upper half */
#define WFFT_STAGE(V, HALF, SG, W)                                              \
    {                                                                           \
        const float ox = __shfl_xor_sync(0xffffffffu, (V).x, HALF);             \
        const float oy = __shfl_xor_sync(0xffffffffu, (V).y, HALF);             \
        const float tx = fmaf(SG, (V).x, ox), ty = fmaf(SG, (V).y, oy);         \
        (V).x = tx * (W).x - ty * (W).y;                                        \
        (V).y = tx * (W).y + ty * (W).x;                                        \
    }

template <int R, bool INV>
__device__ __forceinline__ void wfft(float2 (&v)[R], const float2* __restrict__ T1, int lane, const WfftLane& L)
{
    dif_regs32<R, INV>(v);
#pragma unroll
    for (int i = 1; i < R; ++i) {
        float2 w = __ldg(T1 + i * 32 + lane);
        if (INV) w.y = -w.y;
        v[i] = cmulf(v[i], w);
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
        WFFT_STAGE(v[i], 16, L.s16, L.w16)
        WFFT_STAGE(v[i], 8,  L.s8,  L.w8)
        WFFT_STAGE(v[i], 4,  L.s4,  L.w4)
        WFFT_STAGE(v[i], 2,  L.s2,  L.w2)
        {   /* span 1: twiddle is 1 */
            const float ox = __shfl_xor_sync(0xffffffffu, v[i].x, 1);
            const float oy = __shfl_xor_sync(0xffffffffu, v[i].y, 1);
            v[i].x = fmaf(L.s1, v[i].x, ox);
            v[i].y = fmaf(L.s1, v[i].y, oy);
        }
    }
}

/* the two [R][32] tables of the warp-FFT kernels in global memory, built once per handle (read through L1):
 * T1[i][l] = W_M^(l * bitrev_R(i)),  T2[i][l] = W_N^k for the bin k = bitrev_R(i) + R * bitrev_5(l) that
 * (lane l, slot i) holds after the transform */
static __global__ void wfft_tables_kernel(const float2* __restrict__ gtw, float2* T1, float2* T2, int M, int logR)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M) return;
    const int R = M >> 5, i = idx >> 5, l = idx & 31;
    const int k2 = logR ? (int)(__brev((unsigned)i) >> (32 - logR)) : 0;
    const int e = 2 * l * k2;
    float2 w = gtw[e & (M - 1)];
    if (e >= M) { w.x = -w.x; w.y = -w.y; }
    T1[idx] = w;
    T2[idx] = gtw[k2 + R * (int)(__brev((unsigned)l) >> 27)];
}

/* forward real-FFT split of a transform held in the wfft<R> layout: returns in X[i] the packed real spectrum value
 * of the bin k = bitrev_R(i) + R * bitrev_5(lane) (bin 0 = (DC, Nyquist)), scaled by 2 * half:
 * X[k] = (E + W_N^k O) / 2, E = a + conj b, O = -i (a - conj b), b = Z[M - k] fetched by shuffle */
template <int R>
__device__ __forceinline__ void wfft_fwd_split(const float2 (&v)[R], float2 (&X)[R], const float2* __restrict__ T2, int lane, float half)
{
    constexpr int LOGR = wf_log2(R);
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
        X[i] = make_float2((E.x + tt.x) * half, (E.y + tt.y) * half);
        if (k2 == 0 && lane == 0) X[i] = make_float2((A.x + A.y) * (2.f * half), (A.x - A.y) * (2.f * half));
    }
}

/* M = 1024 = 32 x 32 without shuffles: 32-point DIF in registers, twiddle, transpose through a warp-private
 * shared-memory tile (32 x 33 float2: unit-stride writes, stride-33 reads, both conflict-free), second 32-point DIF in
 * registers.  Lane j holds v[i] = x[j + 32 i] on entry; on exit lane l, slot i holds X[l + 32 * bitrev_5(i)].
 * The shuffle version issues 10 SHFL per point through the same LSU data pipe that serves shared memory (ncu: the
 * transform kernels were MIO-bound on it); this one needs 2 shared-memory accesses per point. */
#define WFFT_TILE 1056     /* float2 per warp: 32 rows x 33 */
template <bool INV, bool ZPAD = false>
__device__ __forceinline__ void wfft32t(float2 (&v)[32], const float2* __restrict__ T1, int lane, float2* tile)
{
    if (ZPAD) dif_regs32_zpad<INV>(v); else dif_regs32<32, INV>(v);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float2 w = __ldg(T1 + i * 32 + lane);
        if (INV) w.y = -w.y;
        if (i) v[i] = cmulf(v[i], w);
        tile[wf_bitrev(i, 5) * 33 + lane] = v[i];               /* row k2, column j */
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tile[lane * 33 + j];    /* lane = k2 */
    dif_regs32<32, INV>(v);
    __syncwarp();                                               /* the tile may be overwritten by the caller */
}

#endif /* SAFCONV_WFFT_CUH_INCLUDED */
