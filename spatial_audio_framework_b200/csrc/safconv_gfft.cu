/*
 * safconv_gfft.cu -- general-size real FFT kernels (any even N) + their thin C-ABI launchers
 *
 * Device side of saf_rfft_create / forward / backward / destroy (reference saf_utility_fft.c:531-753, KissFFT
 * resources/kissFFT/kiss_fftr.c:69-161, kiss_fft.c:93-331) and of the true non-partitioned convolver modes
 * (saf_utility_matrixConv.c:71-96, 174-207): an M = N/2 point complex mixed-radix Stockham FFT (safconv_gfft.cuh)
 * plus the real-FFT split pass.
 *
 *   gfft_smem_kernel     M <= SC_GFFT_SMEM_M: one CTA per transform, both ping-pong arrays in shared memory, every pass
 *                        + the split inside one launch; input / output may be page-locked host memory (zero-copy)
 *   gfft_pass_kernel     larger M: one launch per pass over the whole batch, ping-pong between two device arrays
 *   gfft_split_kernel    forward / inverse split for the multi-launch path
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "safconv_dev.h"
#include "safconv_fft.cuh"
#include "safconv_gfft.cuh"

struct GfftArgs {
    const float* in;        /* forward: [batch][N] real;          backward: [batch][M+1] complex */
    float*       out;       /* forward: [batch][M+1] complex;     backward: [batch][N] real      */
    const float2* tw;       /* W_M^e, e < M      */
    const float2* stw;      /* W_N^k, k <= M/2   */
    int N, M, nf;
    int fac[SC_GFFT_MAX_FACTORS];
    float scale;            /* 1/N */
};

/* all passes of one transform between the shared-memory arrays a and b; returns the array that holds the result */
template <bool INV>
__device__ float2* gfft_passes_smem(float2* a, float2* b, const GfftArgs& g, float lastScale)
{
    int Ns = 1;
    for (int f = 0; f < g.nf; ++f) {
        const int R = g.fac[f];
        const float sc = (f == g.nf - 1) ? lastScale : 1.0f;
        const int nb = g.M / R;
        switch (R) {
            case 2: for (int j = threadIdx.x; j < nb; j += blockDim.x) gfft_bfly<2, INV>(a, b, g.M, Ns, g.tw, j, sc); break;
            case 3: for (int j = threadIdx.x; j < nb; j += blockDim.x) gfft_bfly<3, INV>(a, b, g.M, Ns, g.tw, j, sc); break;
            case 4: for (int j = threadIdx.x; j < nb; j += blockDim.x) gfft_bfly<4, INV>(a, b, g.M, Ns, g.tw, j, sc); break;
            case 5: for (int j = threadIdx.x; j < nb; j += blockDim.x) gfft_bfly<5, INV>(a, b, g.M, Ns, g.tw, j, sc); break;
            default: for (int i = threadIdx.x; i < g.M; i += blockDim.x) gfft_generic_elem<INV>(a, b, g.M, Ns, R, g.tw, i, sc); break;
        }
        __syncthreads();
        float2* t = a; a = b; b = t;
        Ns *= R;
    }
    return a;
}

/* grid (batch); shared memory: 2 * M float2 */
__global__ void gfft_smem_forward_kernel(GfftArgs g)
{
    extern __shared__ __align__(16) float2 gsm[];
    float2 *a = gsm, *b = gsm + g.M;
    const float* x = g.in + (size_t)blockIdx.x * g.N;
    if ((reinterpret_cast<uintptr_t>(x) & 7) == 0) {
        const float2* x2 = reinterpret_cast<const float2*>(x);
        for (int n = threadIdx.x; n < g.M; n += blockDim.x) a[n] = x2[n];
    } else {
        for (int n = threadIdx.x; n < g.M; n += blockDim.x) a[n] = make_float2(x[2 * n], x[2 * n + 1]);
    }
    __syncthreads();
    const float2* Z = gfft_passes_smem<false>(a, b, g, 1.0f);
    float2* X = reinterpret_cast<float2*>(g.out) + (size_t)blockIdx.x * (g.M + 1);
    for (int k = threadIdx.x; k <= (g.M >> 1); k += blockDim.x) gfft_fwd_split(Z, X, g.M, g.stw, k);
}

__global__ void gfft_smem_backward_kernel(GfftArgs g)
{
    extern __shared__ __align__(16) float2 gsm[];
    float2 *a = gsm, *b = gsm + g.M;
    const float2* X = reinterpret_cast<const float2*>(g.in) + (size_t)blockIdx.x * (g.M + 1);
    for (int k = threadIdx.x; k <= (g.M >> 1); k += blockDim.x) gfft_inv_split(X, a, g.M, g.stw, k);
    __syncthreads();
    const float2* z = gfft_passes_smem<true>(a, b, g, g.scale);
    float* x = g.out + (size_t)blockIdx.x * g.N;
    if ((reinterpret_cast<uintptr_t>(x) & 7) == 0) {
        float2* x2 = reinterpret_cast<float2*>(x);
        for (int n = threadIdx.x; n < g.M; n += blockDim.x) x2[n] = z[n];
    } else {
        for (int n = threadIdx.x; n < g.M; n += blockDim.x) { x[2 * n] = z[n].x; x[2 * n + 1] = z[n].y; }
    }
}

/* ---- multi-launch path: one pass over the whole batch.  in / out: [batch][stride] float2 ---- */
struct GfftPassArgs {
    const float2* in; float2* out;
    size_t inStride, outStride;      /* float2 elements between consecutive transforms */
    const float2* tw;
    int M, Ns, R;
    float scale;
};

template <int R, bool INV>
__global__ void gfft_pass_kernel(GfftPassArgs p)
{
    const int nb = p.M / R;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    gfft_bfly<R, INV>(p.in + (size_t)blockIdx.y * p.inStride, p.out + (size_t)blockIdx.y * p.outStride, p.M, p.Ns, p.tw, j, p.scale);
}

template <bool INV>
__global__ void gfft_pass_generic_kernel(GfftPassArgs p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.M) return;
    gfft_generic_elem<INV>(p.in + (size_t)blockIdx.y * p.inStride, p.out + (size_t)blockIdx.y * p.outStride, p.M, p.Ns, p.R, p.tw, i, p.scale);
}

/* forward split: Z [batch][M] -> X [batch][M+1]; inverse split: X [batch][M+1] -> Zc [batch][M] */
__global__ void gfft_split_kernel(const float2* src, float2* dst, const float2* stw, int M, int inverse)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > (M >> 1)) return;
    if (inverse) gfft_inv_split(src + (size_t)blockIdx.y * (M + 1), dst + (size_t)blockIdx.y * M, M, stw, k);
    else         gfft_fwd_split(src + (size_t)blockIdx.y * M, dst + (size_t)blockIdx.y * (M + 1), M, stw, k);
}

/* real [batch][N] (any alignment) -> complex [batch][M] */
__global__ void gfft_pack_kernel(const float* x, float2* z, int M)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= M) return;
    const float* xb = x + (size_t)blockIdx.y * 2 * M;
    z[(size_t)blockIdx.y * M + n] = make_float2(xb[2 * n], xb[2 * n + 1]);
}

template <bool INV>
static int gfft_launch_pass(const GfftPassArgs& p, int nBatch, cudaStream_t st)
{
    const int T = 256;
    if (nBatch > 65535) return (int)cudaErrorInvalidValue;
    switch (p.R) {
        case 2: gfft_pass_kernel<2, INV><<<dim3((p.M / 2 + T - 1) / T, nBatch), T, 0, st>>>(p); break;
        case 3: gfft_pass_kernel<3, INV><<<dim3((p.M / 3 + T - 1) / T, nBatch), T, 0, st>>>(p); break;
        case 4: gfft_pass_kernel<4, INV><<<dim3((p.M / 4 + T - 1) / T, nBatch), T, 0, st>>>(p); break;
        case 5: gfft_pass_kernel<5, INV><<<dim3((p.M / 5 + T - 1) / T, nBatch), T, 0, st>>>(p); break;
        default: gfft_pass_generic_kernel<INV><<<dim3((p.M + T - 1) / T, nBatch), T, 0, st>>>(p); break;
    }
    return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------ */
/*  true non-partitioned convolvers (reference saf_utility_matrixConv.c:174-207, 368-386): ONE FFT of      */
/*  fftSize = numOvrlpAddBlocks * hop per block and channel, per-bin products, fftSize-long overlap-add    */
/* ------------------------------------------------------------------------------------------ */

/* rows of `len` samples (row r at in + r*inStride) -> zero-padded rows of F samples (.c:177, .c:91, .c:371) */
__global__ void np_pad_kernel(const float* __restrict__ in, float* __restrict__ xpad, int len, int F, size_t inStride)
{
    const float* src = in + (size_t)blockIdx.y * inStride;
    float* dst = xpad + (size_t)blockIdx.y * F;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F; i += gridDim.x * blockDim.x) dst[i] = (i < len) ? src[i] : 0.f;
}

/* Z[no][k] = sum_ni H[no][ni][k] * X[ni][k]  (matrix; the reference multiplies (.c:186) and sums AFTER nIn inverse
 * FFTs (.c:192-195) -- the inverse FFT is linear, so summing per bin first gives the same result with one inverse FFT
 * per output);  multi: Z[c][k] = H[c][k] * X[c][k] (.c:378).  grid (ceil(nBins/256), nOut) */
__global__ void np_mac_kernel(const float2* __restrict__ H, const float2* __restrict__ X, float2* __restrict__ Z,
                              int nIn, int nBins, int multi)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nBins) return;
    const int no = blockIdx.y;
    if (multi) {
        Z[(size_t)no * nBins + k] = gf_mul(H[(size_t)no * nBins + k], X[(size_t)no * nBins + k]);
        return;
    }
    const float2* h = H + (size_t)no * nIn * nBins + k;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int ni = 0; ni < nIn; ++ni) {
        const float2 a = h[(size_t)ni * nBins], x = X[(size_t)ni * nBins + k];
        acc.x = fmaf(a.x, x.x, acc.x); acc.x = fmaf(-a.y, x.y, acc.x);
        acc.y = fmaf(a.x, x.y, acc.y); acc.y = fmaf(a.y, x.x, acc.y);
    }
    Z[(size_t)no * nBins + k] = acc;
}

/* shift the overlap-add buffer by one hop, add the new fftSize-long block, emit the first hop samples (.c:198-205):
 * ovNew[i] = (i + hop < F ? ovOld[i + hop] : 0) + z[i];  out[i < hop] = ovNew[i].  grid (ceil(F/256), nOut) */
__global__ void np_ola_kernel(const float* __restrict__ z, const float* __restrict__ ovOld, float* __restrict__ ovNew,
                              float* __restrict__ out, int hop, int F)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F) return;
    const size_t row = (size_t)blockIdx.y * F;
    const float v = ((i + hop < F) ? ovOld[row + i + hop] : 0.f) + z[row + i];
    ovNew[row + i] = v;
    if (i < hop) out[(size_t)blockIdx.y * hop + i] = v;
}

extern "C" {

int scdev_np_pad(const float* in, float* xpad, int rows, int len, int F, size_t inStride, void* stream)
{
    if (rows < 1) return 0;
    if (rows > 65535) return (int)cudaErrorInvalidValue;
    int gx = (F + 255) / 256; if (gx > 64) gx = 64;
    np_pad_kernel<<<dim3(gx, rows), 256, 0, (cudaStream_t)stream>>>(in, xpad, len, F, inStride);
    return (int)cudaGetLastError();
}

int scdev_np_mac(const void* H, const void* X, void* Z, int nOut, int nIn, int nBins, int multi, void* stream)
{
    if (nOut > 65535) return (int)cudaErrorInvalidValue;
    np_mac_kernel<<<dim3((nBins + 255) / 256, nOut), 256, 0, (cudaStream_t)stream>>>((const float2*)H, (const float2*)X, (float2*)Z, nIn, nBins, multi);
    return (int)cudaGetLastError();
}

int scdev_np_ola(const float* z, const float* ovOld, float* ovNew, float* out, int nOut, int hop, int F, void* stream)
{
    if (nOut > 65535) return (int)cudaErrorInvalidValue;
    np_ola_kernel<<<dim3((F + 255) / 256, nOut), 256, 0, (cudaStream_t)stream>>>(z, ovOld, ovNew, out, hop, F);
    return (int)cudaGetLastError();
}

int scdev_gfft_smem_ok(int M) { return M >= 1 && M <= SC_GFFT_SMEM_M; }

static void gfft_fill(GfftArgs& g, const scdev_gfft_plan* p, const float* in, float* out)
{
    g.in = in; g.out = out; g.tw = (const float2*)p->tw; g.stw = (const float2*)p->stw;
    g.N = p->N; g.M = p->M; g.nf = p->nf;
    for (int i = 0; i < SC_GFFT_MAX_FACTORS; ++i) g.fac[i] = p->fac[i];
    g.scale = 1.0f / (float)p->N;
}

static int gfft_threads(int M)
{
    int t = (M / 4 + 31) / 32 * 32;
    if (t < 64) t = 64;
    if (t > 512) t = 512;
    return t;
}

/* dir 0: in [nBatch][N] real -> out [nBatch][M+1] complex; dir 1: the inverse (x 1/N).  Pointers: device memory, or --
 * smem path only -- page-locked host memory.  nBatch <= p->maxBatch on the multi-launch path (work buffers). */
int scdev_gfft_run(const scdev_gfft_plan* p, int dir, int nBatch, const float* in, float* out, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const int M = p->M;
    if (nBatch < 1) return 0;
    if (scdev_gfft_smem_ok(M)) {
        GfftArgs g;
        gfft_fill(g, p, in, out);
        const size_t smem = (size_t)2 * M * sizeof(float2);
        if (dir == 0) {
            if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(gfft_smem_forward_kernel));
            gfft_smem_forward_kernel<<<nBatch, gfft_threads(M), smem, st>>>(g);
        } else {
            if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(gfft_smem_backward_kernel));
            gfft_smem_backward_kernel<<<nBatch, gfft_threads(M), smem, st>>>(g);
        }
        return (int)cudaGetLastError();
    }
    if (nBatch > p->maxBatch || !p->w0 || !p->w1) return (int)cudaErrorInvalidValue;
    const int T = 256;
    float2 *w0 = (float2*)p->w0, *w1 = (float2*)p->w1;
    GfftPassArgs a;
    a.tw = (const float2*)p->tw; a.M = M; a.scale = 1.0f;
    int Ns = 1;
    if (dir == 0) {
        /* x -> (pack if misaligned) -> passes -> split -> X */
        const float2* src;
        if ((reinterpret_cast<uintptr_t>(in) & 7) == 0) src = reinterpret_cast<const float2*>(in);
        else {
            gfft_pack_kernel<<<dim3((M + T - 1) / T, nBatch), T, 0, st>>>(in, w1, M);
            SC_CHECK(cudaGetLastError());
            src = w1;
        }
        float2* dst = w0;
        for (int f = 0; f < p->nf; ++f) {
            a.in = src; a.out = dst; a.inStride = M; a.outStride = M; a.Ns = Ns; a.R = p->fac[f];
            int e = gfft_launch_pass<false>(a, nBatch, st);
            if (e) return e;
            Ns *= p->fac[f];
            src = dst; dst = (dst == w0) ? w1 : w0;
        }
        gfft_split_kernel<<<dim3((M / 2 + 1 + T - 1) / T, nBatch), T, 0, st>>>(src, reinterpret_cast<float2*>(out), (const float2*)p->stw, M, 0);
        return (int)cudaGetLastError();
    }
    /* X -> inverse split -> passes (last one scaled, straight into x when it is 8-byte aligned) */
    gfft_split_kernel<<<dim3((M / 2 + 1 + T - 1) / T, nBatch), T, 0, st>>>(reinterpret_cast<const float2*>(in), w0, (const float2*)p->stw, M, 1);
    SC_CHECK(cudaGetLastError());
    const bool direct = (reinterpret_cast<uintptr_t>(out) & 7) == 0;
    const float2* src = w0;
    float2* dst = w1;
    for (int f = 0; f < p->nf; ++f) {
        const bool last = (f == p->nf - 1);
        a.in = src; a.Ns = Ns; a.R = p->fac[f]; a.inStride = M; a.outStride = M;
        a.out = (last && direct) ? reinterpret_cast<float2*>(out) : dst;
        a.scale = last ? 1.0f / (float)p->N : 1.0f;
        int e = gfft_launch_pass<true>(a, nBatch, st);
        if (e) return e;
        Ns *= p->fac[f];
        src = a.out; dst = (dst == w0) ? w1 : w0;
    }
    if (!direct) SC_CHECK(cudaMemcpyAsync(out, src, (size_t)nBatch * M * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    return 0;
}

} /* extern "C" */
