/*
 * safconv_producers.cu -- device side of the filter PRODUCERS that feed the convolvers + their
 * thin C-ABI launchers (host sequencing: safconv_producers.c)
 *
 *   binaural Ambisonic decoder filters   /root/reference/framework/modules/saf_hoa/saf_hoa.c:393-497
 *                                        (getBinauralAmbiDecoderMtx / getBinauralAmbiDecoderFilters) with the designs of
 *                                        saf_hoa_internal.c:162-228 (LS), :230-330 (LSDIFFEQ), :332-430 (SPR),
 *                                        :432-523 (TA), :525-623 (MagLS), max-rE weighting saf_hoa.c:427-445 and diffuse-field
 *                                        covariance matching saf_hoa.c:497-604
 *   shoebox image-source RIRs            /root/reference/framework/modules/saf_reverb/saf_reverb.c:184-295
 *                                        (ims_shoebox_computeEchograms / ims_shoebox_renderRIRs) with
 *                                        saf_reverb_internal.c:269-710
 *
 * Re-designed, not translated:
 * - The reference solves the SAME (nSH x nSH) normal equations once per band with cgesv (saf_hoa_internal.c:215-224).
 *   Here G = (Y W Y^T)^-1 Y W is formed once (fp64) and every least-squares decoder is ONE product D = H G^T over all
 *   bands (prod_ls_kernel); the per-band designs (diffuse EQ, covariance matching) are one CTA per band on top of it,
 *   MagLS -- a recurrence over the bands -- walks the bands on one thread-block cluster with the direction slices of Y
 *   and G resident in shared memory; SPR folds its t-design projection into the same kind of matrix G.
 * - The reference materialises every image source of every band in echogram containers, sorts them by time, and adds
 *   them into per-band RIRs that are summed afterwards (saf_reverb_internal.c:269-710: ~13 arrays of nImages floats per
 *   band).  Here one thread per lattice point goes from reflection orders to RIR taps in registers: geometry (bit-exact
 *   fp32, safconv_prod_core.cuh), SH encoding, wall absorption from per-axis tables, and fp64 atomic adds into the taps;
 *   nothing but the RIR ever touches HBM.  The fp64 accumulator makes the result independent of the order in which the
 *   hardware retires the atomics to far below fp32 resolution.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "safconv_dev.h"
#include "safconv_fft.cuh"
#include "safconv_sh.cuh"
#include "safconv_prod_core.cuh"

/* ================================================================================================================ */
/*  small reductions                                                                                                 */
/* ================================================================================================================ */
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* sum of K doubles per thread over the whole CTA; result valid in every thread.  red: shared double[K * 32] */
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* red)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < K; k++) red[k * 32 + w] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double s = 0.0;
        for (int i = 0; i < nw; i++) s += red[k * 32 + i];
        v[k] = s;
    }
}

/* ================================================================================================================ */
/*  decoder design                                                                                                   */
/* ================================================================================================================ */

/* Y[q][d]: real SH of every measurement direction (getRSH).  One thread per direction. */
__global__ void prod_rsh_kernel(int order, const float* __restrict__ dirs_deg, int nD, float* __restrict__ Y)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < nD) scsh_rsh_dir(order, dirs_deg[2 * d], dirs_deg[2 * d + 1], Y + d, nD);
}

/* A[i][j] = sum_d Y[i][d] w[d] Y[j][d] (Yna_W_Yna, saf_hoa_internal.c:202-214), written into the left half of the
 * augmented matrix [A | I] (row stride 2n).  grid (nSH, nSH) */
__global__ void prod_gram_kernel(const float* __restrict__ Y, const float* __restrict__ w, int nD, int n, double* __restrict__ aug)
{
    __shared__ double red[32];
    const int i = blockIdx.y, j = blockIdx.x;
    double v[1] = {0.0};
    for (int d = threadIdx.x; d < nD; d += blockDim.x) v[0] += (double)Y[(size_t)i * nD + d] * (double)w[d] * (double)Y[(size_t)j * nD + d];
    block_sum<1>(v, red);
    if (threadIdx.x == 0) {
        aug[(size_t)i * 2 * n + j] = v[0];
        aug[(size_t)i * 2 * n + n + j] = (i == j) ? 1.0 : 0.0;
    }
}

/* Gauss-Jordan on [A | I] -> [I | A^-1], A symmetric positive definite (no pivoting needed), one CTA.
 * flag[0] = 1 if a pivot is not positive (the grid cannot resolve this SH order). */
__global__ void prod_spd_inverse_kernel(double* __restrict__ aug, int n, int* __restrict__ flag)
{
    extern __shared__ double fac[];            /* n */
    __shared__ double piv;
    const int W = 2 * n;
    for (int k = 0; k < n; k++) {
        if (threadIdx.x == 0) piv = aug[(size_t)k * W + k];
        __syncthreads();
        const double p = piv;
        if (!(p > 0.0)) { if (threadIdx.x == 0) flag[0] = 1; return; }
        for (int i = threadIdx.x; i < n; i += blockDim.x) fac[i] = (i == k) ? 0.0 : aug[(size_t)i * W + k];
        __syncthreads();
        const double ip = 1.0 / p;
        for (int j = threadIdx.x; j < W; j += blockDim.x) aug[(size_t)k * W + j] *= ip;
        __syncthreads();
        for (int e = threadIdx.x; e < n * W; e += blockDim.x) {
            const int i = e / W, j = e - i * W;
            const double f = fac[i];
            if (f != 0.0) aug[(size_t)i * W + j] -= f * aug[(size_t)k * W + j];
        }
        __syncthreads();
    }
}

/* G[i][d] = sum_j Ainv[i][j] Y[j][d] w[d]  (fp64 sum, stored as fp32).  grid (ceil(nD / 128), nSH) */
__global__ void prod_g_kernel(const double* __restrict__ aug, const float* __restrict__ Y, const float* __restrict__ w,
                              int nD, int n, float* __restrict__ G)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (d >= nD) return;
    double s = 0.0;
    for (int j = 0; j < n; j++) s += aug[(size_t)i * 2 * n + n + j] * (double)Y[(size_t)j * nD + d];
    G[(size_t)i * nD + d] = (float)(s * (double)w[d]);
}

/* SPR decoder (saf_hoa_internal.c:332-430), set-up 1: condition number of the SH transform of the measurement grid per order
 * (checkCondNumberSHTReal, saf_sh.c:884-953: largest / smallest singular value of Y_n W Y_n^T, which are its eigenvalues).
 * The Gram matrix of the highest order holds those of all lower orders as leading blocks (ACN order nests).  One CTA per
 * order: power iteration for the largest eigenvalue, power iteration on (sigma I - A) for the smallest -- the Rayleigh
 * quotient converges to the VALUE long before the vector does, and only "below or above 100" is asked.
 * aug: [nMax][2 nMax] with the Gram matrix in the left half (prod_gram_kernel); scale: undoes the 4 pi of the N3D scaling.
 * grid (orders); dynamic shared memory: 2 nMax doubles. */
__global__ void __launch_bounds__(512) prod_cond_kernel(const double* __restrict__ aug, int nMax, int itMax, int itMin, double scale,
                                                        float* __restrict__ cond)
{
    extern __shared__ double csm[];
    __shared__ double red[2 * 32];
    const int m = (blockIdx.x + 1) * (blockIdx.x + 1), W = 2 * nMax;
    double *v = csm, *u = csm + nMax;
    double lmax = 0.0, lmin = 0.0;
    for (int phase = 0; phase < 2; phase++) {
        const double sigma = phase ? 1.01 * lmax : 0.0;
        double s[2] = {0.0, 0.0};
        for (int i = threadIdx.x; i < m; i += blockDim.x) { const double x = 1.0 + 0.37 * sin(1.3 * i + 0.5 + phase); v[i] = x; s[0] += x * x; }
        block_sum<2>(s, red);
        const double i0 = rsqrt(s[0]);
        for (int i = threadIdx.x; i < m; i += blockDim.x) v[i] *= i0;
        __syncthreads();
        double lambda = 0.0;
        const int iters = phase ? itMin : itMax;
        for (int it = 0; it < iters; it++) {
            double d[2] = {0.0, 0.0};
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                double a0 = 0.0, a1 = 0.0;
                int j = 0;
                for (; j + 1 < m; j += 2) { a0 += aug[(size_t)j * W + i] * v[j]; a1 += aug[(size_t)(j + 1) * W + i] * v[j + 1]; }   /* A symmetric: column i, coalesced */
                if (j < m) a0 += aug[(size_t)j * W + i] * v[j];
                const double r = phase ? sigma * v[i] - (a0 + a1) : (a0 + a1);
                u[i] = r; d[0] += v[i] * r; d[1] += r * r;
            }
            block_sum<2>(d, red);
            lambda = d[0];
            const double inv = d[1] > 0.0 ? rsqrt(d[1]) : 0.0;
            for (int i = threadIdx.x; i < m; i += blockDim.x) v[i] = u[i] * inv;
            __syncthreads();
        }
        if (phase == 0) lmax = lambda; else lmin = sigma - lambda;
    }
    if (lmin < 0.0) lmin = 0.0;
    if (threadIdx.x == 0) cond[blockIdx.x] = (float)((lmax * scale) / (lmin * scale + 2.23e-7));
}

/* SPR set-up 2: M[a][i] = (1 / K) sum_k Ytd[a][k] Ytd[i][k], a < nA (interpolation order), i < n (decoding order) --
 * the t-design quadrature of the SH products (:413-420 folded into :399-412).  grid (n, nA) */
__global__ void prod_spr_m_kernel(const float* __restrict__ Ytd, int K, int n, double* __restrict__ M)
{
    __shared__ double red[32];
    const int i = blockIdx.x, a = blockIdx.y;
    double v[1] = {0.0};
    for (int k = threadIdx.x; k < K; k += blockDim.x) v[0] += (double)Ytd[(size_t)a * K + k] * (double)Ytd[(size_t)i * K + k];
    block_sum<1>(v, red);
    if (threadIdx.x == 0) M[(size_t)a * n + i] = v[0] / (double)K;
}

/* SPR set-up 3: G[i][d] = w[d] sum_a Ynh[a][d] M[a][i]: with it the SPR decoder of every band is the same product
 * D = H G^T as the least-squares one (prod_ls_kernel).  grid (ceil(nD / 128), n) */
__global__ void prod_spr_g_kernel(const float* __restrict__ Ynh, const double* __restrict__ M, const float* __restrict__ w,
                                  int nD, int nA, int n, float* __restrict__ G)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (d >= nD) return;
    double s = 0.0;
    for (int a = 0; a < nA; a++) s += (double)Ynh[(size_t)a * nD + d] * M[(size_t)a * n + i];
    G[(size_t)i * nD + d] = (float)(s * (double)w[d]);
}

/* least-squares decoder of every band: D[band][ear][i] = sum_d H[src][ear][d] G[i][d], src = band, or the cut-off band
 * for band >= bc when `ta` (TA decoder as written: saf_hoa_internal.c:492-505 re-uses the cut-off band's HRTFs because
 * the phase term multiplies by exp(0)).  grid (nB, 2), one warp per output i in turn. */
__global__ void prod_ls_kernel(const float2* __restrict__ H, const float* __restrict__ G, int nB, int nD, int n,
                               int ta, int bc, float2* __restrict__ D)
{
    const int band = blockIdx.x, ear = blockIdx.y;      /* bands on grid.x: no 65535 limit */
    const int src = (ta && band >= bc) ? bc : band;
    const float2* h = H + ((size_t)src * 2 + ear) * nD;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = w; i < n; i += nw) {
        const float* g = G + (size_t)i * nD;
        double re = 0.0, im = 0.0;
        for (int d = lane; d < nD; d += 32) { const float2 hv = h[d]; const double gv = (double)g[d]; re += gv * (double)hv.x; im += gv * (double)hv.y; }
        re = warp_sum(re); im = warp_sum(im);
        if (lane == 0) D[((size_t)band * 2 + ear) * n + i] = make_float2((float)re, (float)im);
    }
}

/* LSDIFFEQ (saf_hoa_internal.c:284-319): scale the band's decoder by the mean over the ears of
 * sqrt(E_ref / (E_ls + 2.23e-7)), E = diffuse-field energy of the HRTFs / of the decoded HRTFs D Y.  grid (nB) */
__global__ void prod_diffeq_kernel(const float2* __restrict__ H, const float* __restrict__ Y, const float* __restrict__ w,
                                   int nD, int n, float2* __restrict__ D)
{
    extern __shared__ float2 sD[];             /* 2 n */
    __shared__ double red[4 * 32];
    const int band = blockIdx.x;
    float2* Db = D + (size_t)band * 2 * n;
    for (int e = threadIdx.x; e < 2 * n; e += blockDim.x) sD[e] = Db[e];
    __syncthreads();
    double v[4] = {0.0, 0.0, 0.0, 0.0};        /* ref ear 0, ref ear 1, ls ear 0, ls ear 1 */
    for (int d = threadIdx.x; d < nD; d += blockDim.x) {
        double l0r = 0, l0i = 0, l1r = 0, l1i = 0;
        for (int i = 0; i < n; i++) {
            const double y = (double)Y[(size_t)i * nD + d];
            l0r += y * sD[i].x; l0i += y * sD[i].y; l1r += y * sD[n + i].x; l1i += y * sD[n + i].y;
        }
        const float2 h0 = H[((size_t)band * 2) * nD + d], h1 = H[((size_t)band * 2 + 1) * nD + d];
        const double wd = (double)w[d];
        v[0] += wd * ((double)h0.x * h0.x + (double)h0.y * h0.y);
        v[1] += wd * ((double)h1.x * h1.x + (double)h1.y * h1.y);
        v[2] += wd * (l0r * l0r + l0i * l0i);
        v[3] += wd * (l1r * l1r + l1i * l1i);
    }
    block_sum<4>(v, red);
    const float Gh = (sqrtf((float)v[0] / ((float)v[2] + 2.23e-7f)) + sqrtf((float)v[1] / ((float)v[3] + 2.23e-7f))) / 2.0f;
    for (int e = threadIdx.x; e < 2 * n; e += blockDim.x) Db[e] = make_float2(sD[e].x * Gh, sD[e].y * Gh);
}

/* MagLS above the cut-off (saf_hoa_internal.c:596-612): band b takes the MAGNITUDES of its HRTFs and the PHASES of the
 * HRTFs that the decoder of band b-1 reproduces, and fits those least-squares.  A recurrence over the bands: one
 * resident CTA, per band  (A) hm[e][d] = |H_b[e][d]| * unit(sum_i D_{b-1}[e][i] Y[i][d])  (B) D_b[e][i] = sum_d hm[e][d] G[i][d].
 * hm: global scratch float2[2 nD].  Dynamic shared memory: float2[2 n] (the previous band's decoder). */
__global__ void prod_magls_kernel(const float2* __restrict__ H, const float* __restrict__ Y, const float* __restrict__ G,
                                  int nB, int nD, int n, int bc, float2* __restrict__ D, float2* __restrict__ hm)
{
    extern __shared__ float2 sD[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int e = threadIdx.x; e < 2 * n; e += blockDim.x) sD[e] = D[(size_t)bc * 2 * n + e];
    __syncthreads();
    for (int band = bc + 1; band < nB; band++) {
        for (int t = threadIdx.x; t < 2 * nD; t += blockDim.x) {
            const int ear = t / nD, d = t - ear * nD;
            double re = 0.0, im = 0.0;
            const float2* dd = sD + ear * n;
            for (int i = 0; i < n; i++) { const double y = (double)Y[(size_t)i * nD + d]; re += y * dd[i].x; im += y * dd[i].y; }
            const float2 h = H[((size_t)band * 2 + ear) * nD + d];
            const double mag = sqrt((double)h.x * h.x + (double)h.y * h.y);
            const double nr = sqrt(re * re + im * im);
            /* |H| exp(i atan2(im, re)); atan2(0, 0) = 0 */
            hm[t] = (nr > 0.0) ? make_float2((float)(mag * re / nr), (float)(mag * im / nr)) : make_float2((float)mag, 0.0f);
        }
        __syncthreads();
        for (int o = w; o < 2 * n; o += nw) {
            const int ear = o / n, i = o - ear * n;
            const float* g = G + (size_t)i * nD;
            const float2* hh = hm + (size_t)ear * nD;
            double re = 0.0, im = 0.0;
            for (int d = lane; d < nD; d += 32) { const float2 hv = hh[d]; const double gv = (double)g[d]; re += gv * (double)hv.x; im += gv * (double)hv.y; }
            re = warp_sum(re); im = warp_sum(im);
            if (lane == 0) { const float2 r = make_float2((float)re, (float)im); sD[o] = r; D[(size_t)band * 2 * n + o] = r; }
        }
        __syncthreads();
    }
}

/* MagLS recurrence on a thread-block CLUSTER: the directions are cut into one slice per CTA; every CTA keeps its slice of
 * Y and G in SHARED memory for the whole walk over the bands (the single-CTA kernel re-reads both from L2 for every band:
 * 27 us per band), computes per band (A) the target responses of its directions from the previous band's decoder and
 * (B) its slice's contribution to the new decoder, and hands the 2 nSH partial sums to ALL CTAs through distributed
 * shared memory (st.shared::cluster); after one cluster barrier per band every CTA adds the C partials in rank order --
 * the same fp64 sum in every CTA, so all of them hold the same new decoder and go on.  The partial buffers are
 * double-buffered by band parity (a CTA can be at most one barrier ahead).  grid = one cluster of C CTAs. */
struct MaglsArgs { const float2* H; const float* Y; const float* G; float2* D; int nB, nD, n, bc, S; };

__device__ __forceinline__ uint32_t pcl_ctarank()  { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t pcl_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void pcl_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void pcl_store_d2(const void* localSmem, uint32_t rank, double x, double y)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(localSmem)), "r"(rank));
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" :: "r"(remote), "d"(x), "d"(y) : "memory");
}

__global__ void __launch_bounds__(1024, 1) prod_magls_cluster_kernel(MaglsArgs a)
{
    extern __shared__ __align__(16) unsigned char msm[];
    const int C = (int)pcl_nctarank(), rank = (int)pcl_ctarank();
    const int n = a.n, S = a.S, nD = a.nD, d0 = rank * S;
    const int ns = max(0, min(nD, d0 + S) - d0);
    double2* part = reinterpret_cast<double2*>(msm);                 /* [2][C][2 n] */
    float2*  sD   = reinterpret_cast<float2*>(part + 2 * C * 2 * n); /* [2 n]       */
    float2*  hm   = sD + 2 * n;                                      /* [2][S]      */
    float*   Ys   = reinterpret_cast<float*>(hm + 2 * S);            /* [n][S]      */
    float*   Gs   = Ys + (size_t)n * S;                              /* [n][S]      */
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int e = threadIdx.x; e < n * S; e += blockDim.x) {
        const int i = e / S, dl = e - i * S;
        const bool in = dl < ns;
        Ys[e] = in ? a.Y[(size_t)i * nD + d0 + dl] : 0.0f;
        Gs[e] = in ? a.G[(size_t)i * nD + d0 + dl] : 0.0f;
    }
    for (int e = threadIdx.x; e < 2 * n; e += blockDim.x) sD[e] = a.D[(size_t)a.bc * 2 * n + e];
    __syncthreads();
    pcl_sync();                                  /* every CTA of the cluster is running: its shared memory may be written */
    const int t = threadIdx.x;
    const bool act = t < 2 * ns;
    const int ear = act ? t / ns : 0, dl = act ? t - ear * ns : 0;
    float2 hcur = (act && a.bc + 1 < a.nB) ? a.H[((size_t)(a.bc + 1) * 2 + ear) * nD + d0 + dl] : make_float2(0.f, 0.f);
    for (int band = a.bc + 1; band < a.nB; band++) {
        /* the next band's HRTF magnitudes are requested now and used one band later */
        const float2 hnext = (act && band + 1 < a.nB) ? a.H[((size_t)(band + 1) * 2 + ear) * nD + d0 + dl] : make_float2(0.f, 0.f);
        if (act) {
            /* four independent partial sums: the fp64 FMA chain is what a band waits for */
            double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0, i0 = 0.0, i1 = 0.0, i2 = 0.0, i3 = 0.0;
            const float2* dd = sD + ear * n;
            int i = 0;
            for (; i + 3 < n; i += 4) {
                const double y0 = (double)Ys[i * S + dl], y1 = (double)Ys[(i + 1) * S + dl], y2 = (double)Ys[(i + 2) * S + dl], y3 = (double)Ys[(i + 3) * S + dl];
                r0 += y0 * dd[i].x; i0 += y0 * dd[i].y; r1 += y1 * dd[i + 1].x; i1 += y1 * dd[i + 1].y;
                r2 += y2 * dd[i + 2].x; i2 += y2 * dd[i + 2].y; r3 += y3 * dd[i + 3].x; i3 += y3 * dd[i + 3].y;
            }
            for (; i < n; i++) { const double y = (double)Ys[i * S + dl]; r0 += y * dd[i].x; i0 += y * dd[i].y; }
            const double re = (r0 + r1) + (r2 + r3), im = (i0 + i1) + (i2 + i3);
            const double mag = sqrt((double)hcur.x * hcur.x + (double)hcur.y * hcur.y);
            const double nr = sqrt(re * re + im * im);
            hm[ear * S + dl] = (nr > 0.0) ? make_float2((float)(mag * re / nr), (float)(mag * im / nr)) : make_float2((float)mag, 0.0f);
        }
        __syncthreads();
        const int buf = band & 1;
        for (int o0 = 4 * w; o0 < 2 * n; o0 += 4 * nw) {           /* four outputs of a warp at a time: their reductions overlap */
            double re[4] = {0.0, 0.0, 0.0, 0.0}, im[4] = {0.0, 0.0, 0.0, 0.0};
            for (int d = lane; d < ns; d += 32) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int o = min(o0 + k, 2 * n - 1), eo = o / n, i = o - eo * n;
                    const float2 hv = hm[eo * S + d]; const double gv = (double)Gs[(size_t)i * S + d];
                    re[k] += gv * (double)hv.x; im[k] += gv * (double)hv.y;
                }
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1)
#pragma unroll
                for (int k = 0; k < 4; k++) { re[k] += __shfl_xor_sync(0xffffffffu, re[k], sft); im[k] += __shfl_xor_sync(0xffffffffu, im[k], sft); }
            if (lane < C)
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (o0 + k < 2 * n) pcl_store_d2(&part[((size_t)buf * C + rank) * 2 * n + o0 + k], (uint32_t)lane, re[k], im[k]);
        }
        pcl_sync();
        for (int o = threadIdx.x; o < 2 * n; o += blockDim.x) {
            double re = 0.0, im = 0.0;
            for (int r = 0; r < C; r++) { const double2 v = part[((size_t)buf * C + r) * 2 * n + o]; re += v.x; im += v.y; }
            const float2 res = make_float2((float)re, (float)im);
            sD[o] = res;
            if (rank == 0) a.D[(size_t)band * 2 * n + o] = res;
        }
        __syncthreads();
        hcur = hnext;
    }
    pcl_sync();                                  /* no CTA leaves while a neighbour could still write into it */
}

/* max-rE weighting (saf_hoa.c:427-445): D[band][ear][i] *= a[i] */
__global__ void prod_scale_kernel(float2* __restrict__ D, const float* __restrict__ a, int n, size_t total)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < total) { const float s = a[e % n]; D[e].x *= s; D[e].y *= s; }
}

/* diffuse-field covariance matching (saf_hoa.c:540-596), every band except the last (Nyquist).  grid (nB - 1) */
__global__ void prod_diffcov_kernel(const float2* __restrict__ H, const float* __restrict__ Y, const float* __restrict__ w,
                                    int nD, int n, float2* __restrict__ D)
{
    extern __shared__ float2 sD[];             /* 2 n */
    __shared__ double red[8 * 32];
    __shared__ scp_cd sM[2][2];
    const int band = blockIdx.x;
    float2* Db = D + (size_t)band * 2 * n;
    for (int e = threadIdx.x; e < 2 * n; e += blockDim.x) sD[e] = Db[e];
    __syncthreads();
    /* C_ref = H W H^H, C_ambi = (D Y) W (D Y)^H: [0][0], Re [0][1], Im [0][1], [1][1] each */
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int d = threadIdx.x; d < nD; d += blockDim.x) {
        double a0r = 0, a0i = 0, a1r = 0, a1i = 0;
        for (int i = 0; i < n; i++) {
            const double y = (double)Y[(size_t)i * nD + d];
            a0r += y * sD[i].x; a0i += y * sD[i].y; a1r += y * sD[n + i].x; a1i += y * sD[n + i].y;
        }
        const float2 h0 = H[((size_t)band * 2) * nD + d], h1 = H[((size_t)band * 2 + 1) * nD + d];
        const double wd = (double)w[d];
        v[0] += wd * ((double)h0.x * h0.x + (double)h0.y * h0.y);
        v[1] += wd * ((double)h0.x * h1.x + (double)h0.y * h1.y);      /* h0 conj(h1) */
        v[2] += wd * ((double)h0.y * h1.x - (double)h0.x * h1.y);
        v[3] += wd * ((double)h1.x * h1.x + (double)h1.y * h1.y);
        v[4] += wd * (a0r * a0r + a0i * a0i);
        v[5] += wd * (a0r * a1r + a0i * a1i);
        v[6] += wd * (a0i * a1r - a0r * a1i);
        v[7] += wd * (a1r * a1r + a1i * a1i);
    }
    block_sum<8>(v, red);
    if (threadIdx.x == 0) {
        scp_cd M[2][2];
        scp_diffcov_M(v[0], scp_c(v[1], v[2]), v[3], v[4], scp_c(v[5], v[6]), v[7], M);
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) sM[i][j] = M[i][j];
    }
    __syncthreads();
    /* D' = M^H D: D'[e][i] = conj(M[0][e]) D[0][i] + conj(M[1][e]) D[1][i] */
    for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) {
        const int e = t / n, i = t - e * n;
        const scp_cd d0 = scp_c(sD[i].x, sD[i].y), d1 = scp_c(sD[n + i].x, sD[n + i].y);
        const scp_cd r = scp_cadd(scp_cmul(scp_conj(sM[0][e]), d0), scp_cmul(scp_conj(sM[1][e]), d1));
        Db[t] = make_float2((float)r.re, (float)r.im);
    }
}

/* D[band][row] -> Dt[row][band], row = ear * nSH + i: the batch layout of the inverse real FFT (saf_hoa.c:482-489) */
__global__ void prod_pack_kernel(const float2* __restrict__ D, int nB, int rows, float2* __restrict__ Dt)
{
    __shared__ float2 tile[32][33];
    const int b0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int b = b0 + y, r = r0 + threadIdx.x;
        if (b < nB && r < rows) tile[y][threadIdx.x] = D[(size_t)b * rows + r];
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int r = r0 + y, b = b0 + threadIdx.x;
        if (b < nB && r < rows) Dt[(size_t)r * nB + b] = tile[threadIdx.x][y];
    }
}

/* ================================================================================================================ */
/*  image-source simulator                                                                                           */
/* ================================================================================================================ */
struct ImsArgs {
    const ScpImsPair* pairs;
    const float* absTab;      /* [3][nBands][maxW] wall-reflection products per axis and reflection order             */
    int nBands, maxW;         /* maxW = 2 * max(Nx, Ny, Nz) + 1; entry of order o along an axis with half-width N: o + N */
    const float* norms;       /* scsh_recur_norms table                                                               */
    double* acc;              /* fp64 taps, pair p at pairs[p].accOff, [nSH][len]                                     */
    unsigned int* stats;      /* per pair: [0] image count, [1] bits of the largest distance                          */
};

/* pass 1: how many images, and the latest arrival (-> RIR length).  grid (x, nPairs) */
__global__ void ims_count_kernel(ImsArgs a)
{
    const ScpImsPair p = a.pairs[blockIdx.y];
    unsigned int cnt = 0, dmaxBits = 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < p.lengthVec; q += (long long)gridDim.x * blockDim.x) {
        int ii, jj, kk; float sx, sy, sz, d;
        scp_ims_lattice(&p, q, &ii, &jj, &kk);
        if (scp_ims_image(&p, ii, jj, kk, &sx, &sy, &sz, &d)) { cnt++; dmaxBits = max(dmaxBits, __float_as_uint(d)); }   /* d >= 0: bit order = value order */
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, o); dmaxBits = max(dmaxBits, __shfl_xor_sync(0xffffffffu, dmaxBits, o)); }
    if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd(&a.stats[2 * blockIdx.y], cnt); atomicMax(&a.stats[2 * blockIdx.y + 1], dmaxBits); }
}

/* one image source -> its nSH tap contributions: wall absorption (saf_reverb_internal.c:589-633: product over the three
 * axes per band, summed over the bands -- renderRIR adds the band RIRs without filtering them, :697-702), direction, SH
 * encoding (:556-569); add(ch, value) receives them */
template <int NMAX, class Add>
__device__ __forceinline__ void ims_image_to_taps(const ScpImsPair& p, const ImsArgs& a, int ii, int jj, int kk,
                                                  float sx, float sy, float sz, float att, Add add)
{
    const float* tx = a.absTab, *ty = a.absTab + (size_t)a.nBands * a.maxW, *tz = a.absTab + 2 * (size_t)a.nBands * a.maxW;
    double tot = 0.0;
    for (int b = 0; b < a.nBands; b++)
        tot += (double)(tx[b * a.maxW + ii + p.Nx] * ty[b * a.maxW + jj + p.Ny] * tz[b * a.maxW + kk + p.Nz]);
    if (p.order == 0) { add(0, (double)att * tot); return; }
    float azi, incl;
    scp_ims_direction(sx, sy, sz, &azi, &incl);
    float Yv[(NMAX + 1) * (NMAX + 1)];
    scsh_shreal_recur_dir<NMAX>(p.order, azi, incl, a.norms, Yv);
#pragma unroll
    for (int ch = 0; ch < (NMAX + 1) * (NMAX + 1); ch++)
        if (ch < p.nSH) add(ch, (double)(Yv[ch] * att) * tot);
}

/* pass 2, global-atomics version: every lattice point straight into the fp64 taps in HBM.  grid (x, nPairs) */
template <int NMAX>
__global__ void __launch_bounds__(128) ims_render_kernel(ImsArgs a)
{
    const ScpImsPair p = a.pairs[blockIdx.y];
    double* acc = a.acc + p.accOff;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < p.lengthVec; q += (long long)gridDim.x * blockDim.x) {
        int ii, jj, kk; float sx, sy, sz, d;
        scp_ims_lattice(&p, q, &ii, &jj, &kk);
        if (!scp_ims_image(&p, ii, jj, kk, &sx, &sy, &sz, &d)) continue;
        float time, att;
        const int tap = scp_ims_tap(&p, d, &time, &att);
        if (tap < 0 || tap >= p.len) continue;
        ims_image_to_taps<NMAX>(p, a, ii, jj, kk, sx, sy, sz, att,
                                [&](int ch, double v) { atomicAdd(&acc[(size_t)ch * p.len + tap], v); });
    }
}

/* pass 2, windowed version: one CTA owns the taps [w0, w0 + tw) of one pair.  It intersects the lattice rows with the
 * window's spherical shell (scp_ims_row_ranges: a handful of candidates per row instead of a scan of the whole lattice),
 * decides every candidate with the exact fp32 geometry, queues the hits in shared memory so that the expensive part --
 * SH encoding, one thread per image -- runs with full warps, accumulates the taps in SHARED memory (fp64 atomics that
 * never leave the SM) and writes its window of the fp32 RIR once: no atomics in HBM, no fp64 accumulator array, no
 * conversion pass, no memset.  grid (windows, nPairs); dynamic shared memory: double [nSH][tw] + long long [qcap] + the queue counter. */
#define IMS_ROWS_PER_ROUND 8192
template <int NMAX>
__global__ void __launch_bounds__(128) ims_window_kernel(ImsArgs a, float* const* __restrict__ rirPtrs, int accDoubles, int qcap)
{
    extern __shared__ __align__(16) unsigned char ims_smem[];     /* no static shared memory: the opt-in limit is for dynamic + static */
    const ScpImsPair p = a.pairs[blockIdx.y];
    const int tw = p.tw, w0 = blockIdx.x * tw;
    if (w0 >= p.len) return;
    double* sacc = reinterpret_cast<double*>(ims_smem);
    unsigned long long* queue = reinterpret_cast<unsigned long long*>(sacc + accDoubles);
    int& qn = *reinterpret_cast<int*>(queue + qcap);
    const int nAcc = p.nSH * tw;
    for (int e = threadIdx.x; e < nAcc; e += blockDim.x) sacc[e] = 0.0;
    if (threadIdx.x == 0) qn = 0;
    __syncthreads();
    double dlo, dhi; int jr, kr;
    scp_ims_window_range(&p, w0, tw, &dlo, &dhi);
    scp_ims_window_rows(&p, dhi, &jr, &kr);
    const int wj = 2 * jr + 1, nRows = wj * (2 * kr + 1);
    /* rounds of `rows` lattice rows: enumerate into the queue, then drain it with full warps.  A round whose hits do not
     * fit the queue is enumerated again with half the rows (uniform decision, nothing was processed yet); one row never
     * has more than 2 Nx + 1 <= qcap hits (the host layer checks that before choosing this kernel). */
    int rows = IMS_ROWS_PER_ROUND;
    for (int base = 0; base < nRows;) {
        const int end = min(nRows, base + rows);
        for (int r = base + threadIdx.x; r < end; r += blockDim.x) {
            const int jj = r % wj - jr, kk = r / wj - kr;
            int lo[4], hi[4];
            const int nr = scp_ims_row_ranges(&p, jj, kk, dlo, dhi, lo, hi);
#pragma unroll
            for (int s = 0; s < 4; s++)
                for (int ii = (s < nr) ? lo[s] : 1; ii <= ((s < nr) ? hi[s] : 0); ii += 2) {
                    float sx, sy, sz, d, time, att;
                    if (!scp_ims_image(&p, ii, jj, kk, &sx, &sy, &sz, &d)) continue;
                    const int tap = scp_ims_tap(&p, d, &time, &att);
                    if (tap < w0 || tap >= w0 + tw || tap >= p.len) continue;
                    const int slot = atomicAdd(&qn, 1);
                    if (slot < qcap) queue[slot] = ((unsigned long long)(kk + p.Nz) << 42) | ((unsigned long long)(jj + p.Ny) << 21) | (unsigned long long)(ii + p.Nx);
                }
        }
        __syncthreads();
        const int n = qn;
        __syncthreads();
        if (threadIdx.x == 0) qn = 0;
        if (n > qcap && rows > 1) { rows >>= 1; __syncthreads(); continue; }       /* too many hits: same rows again, half as many */
        for (int e = threadIdx.x; e < min(n, qcap); e += blockDim.x) {
            const unsigned long long v = queue[e];
            const int ii = (int)(v & 0x1fffffu) - p.Nx, jj = (int)((v >> 21) & 0x1fffffu) - p.Ny, kk = (int)(v >> 42) - p.Nz;
            float sx, sy, sz, d, time, att;
            scp_ims_image(&p, ii, jj, kk, &sx, &sy, &sz, &d);
            const int t = scp_ims_tap(&p, d, &time, &att) - w0;
            ims_image_to_taps<NMAX>(p, a, ii, jj, kk, sx, sy, sz, att, [&](int ch, double v2) { atomicAdd(&sacc[ch * tw + t], v2); });
        }
        __syncthreads();
        base = end;
        if (rows < IMS_ROWS_PER_ROUND && 4 * n < qcap) rows <<= 1;
    }
    float* rir = rirPtrs[blockIdx.y];
    for (int e = threadIdx.x; e < nAcc; e += blockDim.x) {
        const int ch = e / tw, t = e - ch * tw;
        if (w0 + t < p.len) rir[(size_t)ch * p.len + w0 + t] = (float)sacc[e];
    }
}

/* fp64 taps -> fp32 RIR (same layout) */
__global__ void ims_finish_kernel(const double* __restrict__ acc, float* __restrict__ rir, size_t total)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < total) rir[e] = (float)acc[e];
}

/* RIRs of one receiver -> convolver filter bank Hflat[ch][src][L] (zero-padded to L): the layout saf_matrixConv_create
 * takes (nCHout x nCHin x length_h, saf_utility_matrixConv.h:48).  grid (ceil(L / 256), nSrc, nCh) */
struct BankArgs { const float* const* rir; const int* len; float* H; int nSrc, L; };
__global__ void ims_bank_kernel(BankArgs b)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y, ch = blockIdx.z;
    if (t >= b.L) return;
    const int len = b.len[s];
    b.H[((size_t)ch * b.nSrc + s) * b.L + t] = (t < len) ? b.rir[s][(size_t)ch * len + t] : 0.0f;
}

/* ================================================================================================================ */
/*  launchers                                                                                                        */
/* ================================================================================================================ */
extern "C" {

int scdev_prod_rsh(int order, const float* d_dirs, int nD, float* d_Y, void* stream)
{
    prod_rsh_kernel<<<(nD + 127) / 128, 128, 0, (cudaStream_t)stream>>>(order, d_dirs, nD, d_Y);
    return (int)cudaGetLastError();
}

/* G = (Y W Y^T)^-1 Y W.  d_aug: double[n][2n] scratch, d_flag: int (1 = singular). */
int scdev_prod_lsmatrix(const float* d_Y, const float* d_w, int nD, int n, double* d_aug, float* d_G, int* d_flag, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    SC_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    prod_gram_kernel<<<dim3(n, n), 128, 0, st>>>(d_Y, d_w, nD, n, d_aug);
    prod_spd_inverse_kernel<<<1, 1024, n * sizeof(double), st>>>(d_aug, n, d_flag);
    prod_g_kernel<<<dim3((nD + 127) / 128, n), 128, 0, st>>>(d_aug, d_Y, d_w, nD, n, d_G);
    return (int)cudaGetLastError();
}

/* SPR: condition numbers of orders 0 .. nhMax from the Gram matrix of d_Y [(nhMax+1)^2][nD] with weights d_w (ones if the
 * caller gave none); d_aug: double [nS][2 nS] scratch, nS = (nhMax+1)^2; d_cond: float [nhMax + 1] */
int scdev_prod_spr_cond(const float* d_Y, const float* d_w, int nD, int nhMax, double* d_aug, float* d_cond, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const int nS = (nhMax + 1) * (nhMax + 1);
    prod_gram_kernel<<<dim3(nS, nS), 128, 0, st>>>(d_Y, d_w, nD, nS, d_aug);
    prod_cond_kernel<<<nhMax + 1, 512, 2 * nS * sizeof(double), st>>>(d_aug, nS, 100, 800, 1.0 / (4.0 * 3.14159265358979323846), d_cond);
    return (int)cudaGetLastError();
}

/* SPR: G [n][nD] from Ynh [nA][nD] (measurement grid, interpolation order), Ytd [nA][K] (t-design, same order), weights */
int scdev_prod_spr_matrix(const float* d_Ynh, const float* d_Ytd, const float* d_w, int nD, int K, int nA, int n,
                          double* d_M, float* d_G, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    prod_spr_m_kernel<<<dim3(n, nA), 64, 0, st>>>(d_Ytd, K, n, d_M);
    prod_spr_g_kernel<<<dim3((nD + 127) / 128, n), 128, 0, st>>>(d_Ynh, d_M, d_w, nD, nA, n, d_G);
    return (int)cudaGetLastError();
}

int scdev_prod_ls(const void* d_H, const float* d_G, int nB, int nD, int n, int ta, int bc, void* d_D, void* stream)
{
    prod_ls_kernel<<<dim3(nB, 2), 256, 0, (cudaStream_t)stream>>>((const float2*)d_H, d_G, nB, nD, n, ta, bc, (float2*)d_D);
    return (int)cudaGetLastError();
}

int scdev_prod_diffeq(const void* d_H, const float* d_Y, const float* d_w, int nB, int nD, int n, void* d_D, void* stream)
{
    prod_diffeq_kernel<<<nB, 256, 2 * n * sizeof(float2), (cudaStream_t)stream>>>((const float2*)d_H, d_Y, d_w, nD, n, (float2*)d_D);
    return (int)cudaGetLastError();
}

int scdev_prod_magls(const void* d_H, const float* d_Y, const float* d_G, int nB, int nD, int n, int bc, void* d_D, void* d_hm, void* stream)
{
    if (bc + 1 >= nB) return 0;
    /* cluster version: 8 CTAs, each with its direction slice of Y and G in shared memory (needs 2 S <= 1024 threads' worth of
     * directions per CTA and the slices + partial buffers within the opt-in shared memory); else one CTA streaming from L2 */
    const int C = 8, S = (nD + C - 1) / C;
    const size_t smem = sizeof(double) * 2 * ((size_t)2 * C * 2 * n) + sizeof(float) * 2 * ((size_t)2 * n + 2 * S) + sizeof(float) * 2 * (size_t)n * S;
    const char* env = getenv("SAFCONV_MAGLS_CLUSTER");
    if (!(env && env[0] == '0') && 2 * S <= 1024 && smem <= 222 * 1024) {
        MaglsArgs a = { (const float2*)d_H, d_Y, d_G, (float2*)d_D, nB, nD, n, bc, S };
        if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(prod_magls_cluster_kernel));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)C); cfg.blockDim = dim3(2 * S <= 512 ? 512 : 1024); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return (int)cudaLaunchKernelEx(&cfg, prod_magls_cluster_kernel, a);
    }
    prod_magls_kernel<<<1, 1024, 2 * n * sizeof(float2), (cudaStream_t)stream>>>((const float2*)d_H, d_Y, d_G, nB, nD, n, bc, (float2*)d_D, (float2*)d_hm);
    return (int)cudaGetLastError();
}

int scdev_prod_scale(void* d_D, const float* d_a, int n, size_t total, void* stream)
{
    prod_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((float2*)d_D, d_a, n, total);
    return (int)cudaGetLastError();
}

int scdev_prod_diffcov(const void* d_H, const float* d_Y, const float* d_w, int nB, int nD, int n, void* d_D, void* stream)
{
    if (nB < 2) return 0;
    prod_diffcov_kernel<<<nB - 1, 256, 2 * n * sizeof(float2), (cudaStream_t)stream>>>((const float2*)d_H, d_Y, d_w, nD, n, (float2*)d_D);
    return (int)cudaGetLastError();
}

int scdev_prod_pack(const void* d_D, int nB, int rows, void* d_Dt, void* stream)
{
    prod_pack_kernel<<<dim3((nB + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>((const float2*)d_D, nB, rows, (float2*)d_Dt);
    return (int)cudaGetLastError();
}

static int ims_grid_x(long long maxLen, int smCount)
{
    long long g = (maxLen + 127) / 128;
    const long long cap = (long long)smCount * 16;          /* 16 CTAs of 128 threads per SM, grid-stride beyond that */
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int scdev_ims_count(const void* d_pairs, int nPairs, long long maxLengthVec, unsigned int* d_stats, int smCount, void* stream)
{
    ImsArgs a = {};
    a.pairs = (const ScpImsPair*)d_pairs; a.stats = d_stats;
    SC_CHECK(cudaMemsetAsync(d_stats, 0, sizeof(unsigned int) * 2 * (size_t)nPairs, (cudaStream_t)stream));
    ims_count_kernel<<<dim3(ims_grid_x(maxLengthVec, smCount), nPairs), 128, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

int scdev_ims_render(const void* d_pairs, int nPairs, long long maxLengthVec, int maxOrder, const float* d_absTab, int nBands, int maxW,
                     const float* d_norms, double* d_acc, size_t totalTaps, int smCount, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    ImsArgs a = {};
    a.pairs = (const ScpImsPair*)d_pairs; a.absTab = d_absTab; a.nBands = nBands; a.maxW = maxW; a.norms = d_norms; a.acc = d_acc;
    SC_CHECK(cudaMemsetAsync(d_acc, 0, sizeof(double) * totalTaps, st));
    const dim3 grid(ims_grid_x(maxLengthVec, smCount), nPairs);
    if (maxOrder <= 3)      ims_render_kernel<3><<<grid, 128, 0, st>>>(a);
    else if (maxOrder <= 7) ims_render_kernel<7><<<grid, 128, 0, st>>>(a);
    else                    ims_render_kernel<SCSH_MAX_ORDER><<<grid, 128, 0, st>>>(a);
    return (int)cudaGetLastError();
}

/* windowed render: d_rirPtrs[pair] = that pair's fp32 RIR [nSH][len]; maxWindows = max over pairs of ceil(len / tw);
 * accDoubles = max over pairs of nSH * tw */
int scdev_ims_render_windows(const void* d_pairs, int nPairs, int maxWindows, int accDoubles, int maxOrder, const float* d_absTab,
                             int nBands, int maxW, const float* d_norms, float* const* d_rirPtrs, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    ImsArgs a = {};
    a.pairs = (const ScpImsPair*)d_pairs; a.absTab = d_absTab; a.nBands = nBands; a.maxW = maxW; a.norms = d_norms;
    const int qcap = 2048;
    const size_t smem = sizeof(double) * (size_t)accDoubles + sizeof(unsigned long long) * qcap + 16;     /* + the queue counter */
    const dim3 grid(maxWindows, nPairs);
    if (maxOrder <= 3) {
        if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(ims_window_kernel<3>));
        ims_window_kernel<3><<<grid, 128, smem, st>>>(a, d_rirPtrs, accDoubles, qcap);
    } else if (maxOrder <= 7) {
        if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(ims_window_kernel<7>));
        ims_window_kernel<7><<<grid, 128, smem, st>>>(a, d_rirPtrs, accDoubles, qcap);
    } else {
        if (smem > 48 * 1024) SC_CHECK(sc_optin_smem(ims_window_kernel<SCSH_MAX_ORDER>));
        ims_window_kernel<SCSH_MAX_ORDER><<<grid, 128, smem, st>>>(a, d_rirPtrs, accDoubles, qcap);
    }
    return (int)cudaGetLastError();
}

int scdev_ims_finish(const double* d_acc, float* d_rir, size_t total, void* stream)
{
    ims_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_acc, d_rir, total);
    return (int)cudaGetLastError();
}

int scdev_ims_bank(const float* const* d_rirPtrs, const int* d_len, int nSrc, int nCh, int L, float* d_H, void* stream)
{
    BankArgs b = { d_rirPtrs, d_len, d_H, nSrc, L };
    ims_bank_kernel<<<dim3((L + 255) / 256, nSrc, nCh), 256, 0, (cudaStream_t)stream>>>(b);
    return (int)cudaGetLastError();
}

int scdev_mem_free_bytes(size_t* freeBytes)
{
    size_t total = 0;
    return (int)cudaMemGetInfo(freeBytes, &total);
}

int scdev_memcpy_d2d_async(void* dst, const void* src, size_t bytes, void* stream)
{ return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream); }

}  /* extern "C" */
