/*
 * safconv_gfft.cuh -- general-size FFT building blocks (any M = product of small primes and/or arbitrary primes)
 *
 * The reference's real FFT accepts ANY even size N (saf_rfft_create, saf_utility_fft.c:531-613: "Only even (non zero)
 * FFT sizes are supported") and its default backend KissFFT factors N/2 into 4s, 2s, 3s, 5s and whatever primes remain
 * (resources/kissFFT/kiss_fft.c:310-331, butterflies :93-235).  The non-partitioned convolver modes use such sizes
 * (fftSize = numOvrlpAddBlocks * hopSize, saf_utility_matrixConv.c:71-96: e.g. 1280).
 *
 * Here: a mixed-radix STOCKHAM autosort FFT (natural order in, natural order out, no bit reversal, ping-pong between two
 * arrays).  One pass of radix R over an M-point array with Ns = product of the radices already processed:
 *
 *     butterfly j < M/R:   k = j mod Ns
 *         v[r]  = in[j + r*M/R] * W_(Ns*R)^(r*k)          r = 0..R-1
 *         V     = DFT_R(v)
 *         out[(j div Ns)*Ns*R + k + q*Ns] = V[q]          q = 0..R-1
 *
 * Radices 2, 3, 4, 5 are register butterflies; any other prime runs the O(R) sum per output element (one thread per
 * element), like kf_bfly_generic.  Twiddles come from ONE table W_M^e, e < M, evaluated in double and rounded
 * (kiss_fft.c:358-364): W_(Ns*R)^x = W_M^(x * M/(Ns*R)).
 *
 * Every function is __host__ __device__: tests/test_gfft_host.py compiles this header into a host program (nvcc, no GPU
 * needed) and checks the passes and the real-FFT split against numpy; the kernels in safconv_gfft.cu call the same code.
 */
#ifndef SAFCONV_GFFT_CUH_INCLUDED
#define SAFCONV_GFFT_CUH_INCLUDED

#include <cuda_runtime.h>

#define SC_GFFT_HD __host__ __device__ __forceinline__

SC_GFFT_HD float2 gf_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SC_GFFT_HD float2 gf_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SC_GFFT_HD float2 gf_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
/* a * (-i) forward, a * (+i) inverse */
template <bool INV> SC_GFFT_HD float2 gf_rot(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
template <bool INV> SC_GFFT_HD float2 gf_tw(const float2* __restrict__ tw, int e)
{
    float2 w = tw[e];
    if (INV) w.y = -w.y;
    return w;
}

/* R-point DFT in registers (forward: exp(-2 pi i r q / R); INV: the conjugate), natural order in and out */
template <int R, bool INV> struct GfDft;

template <bool INV> struct GfDft<2, INV> {
    static SC_GFFT_HD void run(float2 (&v)[2])
    {
        const float2 a = v[0], b = v[1];
        v[0] = gf_add(a, b); v[1] = gf_sub(a, b);
    }
};

template <bool INV> struct GfDft<3, INV> {
    static SC_GFFT_HD void run(float2 (&v)[3])
    {
        const float S3 = 0.86602540378443864676f;                   /* sin(pi/3) */
        const float2 t1 = gf_add(v[1], v[2]), t2 = gf_sub(v[1], v[2]);
        const float2 m = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
        const float2 s = gf_rot<INV>(make_float2(S3 * t2.x, S3 * t2.y));      /* -/+ i * S3 * t2 */
        v[0] = gf_add(v[0], t1);
        v[1] = gf_add(m, s);
        v[2] = gf_sub(m, s);
    }
};

template <bool INV> struct GfDft<4, INV> {
    static SC_GFFT_HD void run(float2 (&v)[4])
    {
        const float2 a = gf_add(v[0], v[2]), b = gf_sub(v[0], v[2]);
        const float2 c = gf_add(v[1], v[3]), d = gf_rot<INV>(gf_sub(v[1], v[3]));   /* -/+ i (v1 - v3) */
        v[0] = gf_add(a, c); v[2] = gf_sub(a, c);
        v[1] = gf_add(b, d); v[3] = gf_sub(b, d);
    }
};

template <bool INV> struct GfDft<5, INV> {
    static SC_GFFT_HD void run(float2 (&v)[5])
    {
        const float C1 = 0.30901699437494742410f, C2 = -0.80901699437494742410f;   /* cos(2 pi/5), cos(4 pi/5) */
        const float S1 = 0.95105651629515357212f, S2 = 0.58778525229247312917f;    /* sin(2 pi/5), sin(4 pi/5) */
        const float2 t1 = gf_add(v[1], v[4]), t2 = gf_add(v[2], v[3]);
        const float2 t3 = gf_sub(v[1], v[4]), t4 = gf_sub(v[2], v[3]);
        const float2 a1 = make_float2(v[0].x + C1 * t1.x + C2 * t2.x, v[0].y + C1 * t1.y + C2 * t2.y);
        const float2 a2 = make_float2(v[0].x + C2 * t1.x + C1 * t2.x, v[0].y + C2 * t1.y + C1 * t2.y);
        const float2 b1 = gf_rot<INV>(make_float2(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y));   /* -/+ i (S1 t3 + S2 t4) */
        const float2 b2 = gf_rot<INV>(make_float2(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y));   /* -/+ i (S2 t3 - S1 t4) */
        v[0] = gf_add(v[0], gf_add(t1, t2));
        v[1] = gf_add(a1, b1); v[4] = gf_sub(a1, b1);
        v[2] = gf_add(a2, b2); v[3] = gf_sub(a2, b2);
    }
};

/* one radix-R butterfly j < M/R of a Stockham pass; `scale` multiplies the stored values (1/N on the last inverse pass) */
template <int R, bool INV>
SC_GFFT_HD void gfft_bfly(const float2* __restrict__ in, float2* __restrict__ out, int M, int Ns,
                          const float2* __restrict__ tw, int j, float scale)
{
    const int nb = M / R;
    const int k = j % Ns;
    const int tstep = M / (Ns * R);
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        v[r] = in[j + r * nb];
        if (r > 0 && k > 0) v[r] = gf_mul(v[r], gf_tw<INV>(tw, r * k * tstep));      /* r*k*tstep < M */
    }
    GfDft<R, INV>::run(v);
    const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) out[j0 + q * Ns] = make_float2(v[q].x * scale, v[q].y * scale);
}

/* any other radix (primes >= 7): output element idx < M of the pass, an O(R) sum (kf_bfly_generic, kiss_fft.c:187-235) */
template <bool INV>
SC_GFFT_HD void gfft_generic_elem(const float2* __restrict__ in, float2* __restrict__ out, int M, int Ns, int R,
                                  const float2* __restrict__ tw, int idx, float scale)
{
    const int nb = M / R;
    const int j = idx % nb, q = idx / nb;
    const int k = j % Ns;
    const int tstep = M / (Ns * R);
    const long long span = (long long)Ns * R;
    const long long kq = (long long)k + (long long)q * Ns;
    float2 acc = in[j];
    for (int r = 1; r < R; ++r) {
        const int e = (int)(((long long)r * kq) % span) * tstep;
        acc = gf_add(acc, gf_mul(in[j + r * nb], gf_tw<INV>(tw, e)));
    }
    out[(j / Ns) * Ns * R + k + q * Ns] = make_float2(acc.x * scale, acc.y * scale);
}

/* real-FFT split passes on NATURAL-order arrays (kiss_fftr.c:69-123, 125-161), stw[k] = exp(-2 pi i k / N), k <= M/2.
 * forward, pair (k, M-k), 1 <= k <= M/2: X[k] = E + W_N^k O, X[M-k] = conj(E - W_N^k O),
 * E = (a + conj b)/2, O = -i (a - conj b)/2, a = Z[k], b = Z[M-k] */
SC_GFFT_HD void gfft_fwd_split(const float2* __restrict__ Z, float2* __restrict__ X, int M, const float2* __restrict__ stw, int k)
{
    if (k == 0) {
        const float2 z = Z[0];
        X[0] = make_float2(z.x + z.y, 0.f);
        X[M] = make_float2(z.x - z.y, 0.f);
        return;
    }
    const float2 a = Z[k], b = Z[M - k];
    const float2 E = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
    const float2 O = make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x));
    const float2 t = gf_mul(stw[k], O);
    X[k]     = make_float2(E.x + t.x, E.y + t.y);
    X[M - k] = make_float2(E.x - t.x, t.y - E.y);
}

/* inverse, pair (k, M-k): Zc[k] = E + iO, Zc[M-k] = conj(E) + i conj(O), E = A + conj B, O = (A - conj B) W_N^-k
 * (the factor 1/2 is folded into the final 1/N); k = 0 uses only the real parts of X[0] and X[M] (kiss_fftr.c:137-138) */
SC_GFFT_HD void gfft_inv_split(const float2* __restrict__ X, float2* __restrict__ Zc, int M, const float2* __restrict__ stw, int k)
{
    if (k == 0) {
        const float a = X[0].x, b = X[M].x;
        Zc[0] = make_float2(a + b, a - b);
        return;
    }
    const float2 A = X[k], B = X[M - k];
    const float2 E = make_float2(A.x + B.x, A.y - B.y);
    const float2 D = make_float2(A.x - B.x, A.y + B.y);
    const float2 w = stw[k];
    const float2 O = make_float2(D.x * w.x + D.y * w.y, D.y * w.x - D.x * w.y);   /* D * conj(w) */
    Zc[k]     = make_float2(E.x - O.y, E.y + O.x);
    Zc[M - k] = make_float2(E.x + O.y, O.x - E.y);
}

#endif /* SAFCONV_GFFT_CUH_INCLUDED */
