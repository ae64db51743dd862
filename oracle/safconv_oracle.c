/*
 * safconv_oracle.c  --  TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * A plain-C, single-threaded CPU restatement of the reference's multichannel
 * convolution path (SAF v1.3.0 fork under /root/reference), written so that the
 * CUDA implementation in spatial_audio_framework_b200/csrc can be checked on
 * machines where /root/reference does not exist (the GPU box).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this file's shared object, and only as the
 * checker.  The product library (libsafconv_b200.so) never links or calls it.
 *
 * Parity status: PINNED AGAINST THE COMPILED REFERENCE.
 *   The reference's own tests hold no golden vectors for this path
 *   (test__saf_matrixConv has zero assertions), so the pin is
 *   the reference itself: oracle/Makefile compiles the unmodified reference
 *   sources in place into oracle/_ref/libsaf_ref_conv_*.so, and
 *   tests/test_oracle_vs_reference.py checks that every function below
 *   reproduces that library's output BIT-FOR-BIT (float32) on seeded inputs,
 *   in both modes, including non-power-of-two FFT sizes (radix 3/5/generic).
 *   Golden outputs produced by the compiled reference are committed under
 *   tests/golden/ (made by tests/golden/make_golden.py) so the same pin holds
 *   where the reference tree is absent.
 *
 * Every function states the reference file:line it follows.  Paths are
 * relative to /root/reference/framework/.
 *   MC  = modules/saf_utilities/saf_utility_matrixConv.c
 *   FFT = modules/saf_utilities/saf_utility_fft.c
 *   KF  = resources/kissFFT/kiss_fft.c
 *   KFR = resources/kissFFT/kiss_fftr.c
 *
 * The arithmetic is arranged so that each float operation has the same
 * operands and order as in the reference (compile with -ffp-contract=off);
 * the control structure, data structures and naming are our own.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float re, im; } cpx;

static inline cpx c_mul(cpx a, cpx b)   /* product order of KF guts C_MUL (_kiss_fft_guts.h) */
{
    cpx m;
    m.re = a.re * b.re - a.im * b.im;
    m.im = a.re * b.im + a.im * b.re;
    return m;
}
static inline cpx c_add(cpx a, cpx b) { cpx r = { a.re + b.re, a.im + b.im }; return r; }
static inline cpx c_sub(cpx a, cpx b) { cpx r = { a.re - b.re, a.im - b.im }; return r; }

/* ------------------------------------------------------------------------- */
/*  Mixed-radix complex FFT (restates KF:22-303 kf_bfly* / kf_work,          */
/*  KF:310-331 kf_factor, KF:340-370 kiss_fft_alloc)                         */
/* ------------------------------------------------------------------------- */

#define ORC_MAX_STAGES 32

typedef struct {
    int  n;                       /* complex length */
    int  inverse;
    int  nstages;
    int  radix[ORC_MAX_STAGES];   /* p_i  */
    int  rest[ORC_MAX_STAGES];    /* m_i = n / (p_0*...*p_i) */
    cpx* tw;                      /* tw[i] = exp(-/+ 2 pi i / n), evaluated in double (KF:358-364) */
} cfft_plan;

/* KF:310-331 : peel 4s, then 2s, then odd primes 3,5,7,...; stop trying once p > floor(sqrt(n)) */
static void plan_factor(cfft_plan* pl)
{
    int n = pl->n, p = 4, s = 0;
    double root = floor(sqrt((double)n));
    do {
        while (n % p) {
            if (p == 4) p = 2;
            else if (p == 2) p = 3;
            else p += 2;
            if (p > root) p = n;
        }
        n /= p;
        pl->radix[s] = p;
        pl->rest[s]  = n;
        s++;
    } while (n > 1);
    pl->nstages = s;
}

static cfft_plan* cfft_plan_new(int n, int inverse)
{
    cfft_plan* pl = (cfft_plan*)calloc(1, sizeof(cfft_plan));
    pl->n = n; pl->inverse = inverse;
    pl->tw = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    for (int i = 0; i < n; i++) {
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        double ph = -2 * pi * i / n;              /* KF:360-363 (same expression, same rounding) */
        if (inverse) ph *= -1;
        pl->tw[i].re = (float)cos(ph);
        pl->tw[i].im = (float)sin(ph);
    }
    plan_factor(pl);
    return pl;
}

static void cfft_plan_free(cfft_plan* pl) { if (pl) { free(pl->tw); free(pl); } }

/* radix-2 combine, KF:22-44 */
static void comb2(cpx* v, int m, int ts, const cfft_plan* pl)
{
    for (int k = 0; k < m; k++) {
        cpx t = c_mul(v[k + m], pl->tw[(size_t)k * ts]);
        v[k + m] = c_sub(v[k], t);
        v[k]     = c_add(v[k], t);
    }
}

/* radix-4 combine, KF:46-91 */
static void comb4(cpx* v, int m, int ts, const cfft_plan* pl)
{
    for (int k = 0; k < m; k++) {
        cpx a = c_mul(v[k + m],     pl->tw[(size_t)k * ts]);
        cpx b = c_mul(v[k + 2 * m], pl->tw[(size_t)k * ts * 2]);
        cpx c = c_mul(v[k + 3 * m], pl->tw[(size_t)k * ts * 3]);
        cpx d   = c_sub(v[k], b);        /* scratch[5] */
        v[k]    = c_add(v[k], b);
        cpx apc = c_add(a, c);           /* scratch[3] */
        cpx amc = c_sub(a, c);           /* scratch[4] */
        v[k + 2 * m] = c_sub(v[k], apc);
        v[k]         = c_add(v[k], apc);
        if (pl->inverse) {
            v[k + m].re     = d.re - amc.im;  v[k + m].im     = d.im + amc.re;
            v[k + 3 * m].re = d.re + amc.im;  v[k + 3 * m].im = d.im - amc.re;
        } else {
            v[k + m].re     = d.re + amc.im;  v[k + m].im     = d.im - amc.re;
            v[k + 3 * m].re = d.re - amc.im;  v[k + 3 * m].im = d.im + amc.re;
        }
    }
}

/* radix-3 combine, KF:93-135 */
static void comb3(cpx* v, int m, int ts, const cfft_plan* pl)
{
    const float e3i = pl->tw[(size_t)ts * m].im;
    for (int k = 0; k < m; k++) {
        cpx a = c_mul(v[k + m],     pl->tw[(size_t)k * ts]);
        cpx b = c_mul(v[k + 2 * m], pl->tw[(size_t)k * ts * 2]);
        cpx s = c_add(a, b);
        cpx d = c_sub(a, b);
        v[k + m].re = v[k].re - s.re * .5f;
        v[k + m].im = v[k].im - s.im * .5f;
        d.re *= e3i; d.im *= e3i;
        v[k] = c_add(v[k], s);
        v[k + 2 * m].re = v[k + m].re + d.im;
        v[k + 2 * m].im = v[k + m].im - d.re;
        v[k + m].re -= d.im;
        v[k + m].im += d.re;
    }
}

/* radix-5 combine, KF:137-196 */
static void comb5(cpx* v, int m, int ts, const cfft_plan* pl)
{
    const cpx ya = pl->tw[(size_t)ts * m], yb = pl->tw[(size_t)ts * 2 * m];
    for (int u = 0; u < m; u++) {
        cpx s0 = v[u];
        cpx s1 = c_mul(v[u + m],     pl->tw[(size_t)u * ts]);
        cpx s2 = c_mul(v[u + 2 * m], pl->tw[(size_t)2 * u * ts]);
        cpx s3 = c_mul(v[u + 3 * m], pl->tw[(size_t)3 * u * ts]);
        cpx s4 = c_mul(v[u + 4 * m], pl->tw[(size_t)4 * u * ts]);
        cpx s7 = c_add(s1, s4), s10 = c_sub(s1, s4);
        cpx s8 = c_add(s2, s3), s9  = c_sub(s2, s3);
        cpx s5, s6, s11, s12;

        v[u].re += s7.re + s8.re;
        v[u].im += s7.im + s8.im;

        s5.re = s0.re + s7.re * ya.re + s8.re * yb.re;
        s5.im = s0.im + s7.im * ya.re + s8.im * yb.re;
        s6.re =  s10.im * ya.im + s9.im * yb.im;
        s6.im = -(s10.re * ya.im) - s9.re * yb.im;
        v[u + m]     = c_sub(s5, s6);
        v[u + 4 * m] = c_add(s5, s6);

        s11.re = s0.re + s7.re * yb.re + s8.re * ya.re;
        s11.im = s0.im + s7.im * yb.re + s8.im * ya.re;
        s12.re = -(s10.im * yb.im) + s9.im * ya.im;
        s12.im =  s10.re * yb.im - s9.re * ya.im;
        v[u + 2 * m] = c_add(s11, s12);
        v[u + 3 * m] = c_sub(s11, s12);
    }
}

/* any other prime radix, KF:198-237 */
static void comb_any(cpx* v, int m, int p, int ts, const cfft_plan* pl)
{
    cpx* tmp = (cpx*)malloc(sizeof(cpx) * (size_t)p);
    for (int u = 0; u < m; u++) {
        for (int q = 0, k = u; q < p; q++, k += m) tmp[q] = v[k];
        for (int q1 = 0, k = u; q1 < p; q1++, k += m) {
            int idx = 0;
            v[k] = tmp[0];
            for (int q = 1; q < p; q++) {
                idx += ts * k;
                if (idx >= pl->n) idx -= pl->n;
                v[k] = c_add(v[k], c_mul(tmp[q], pl->tw[idx]));
            }
        }
    }
    free(tmp);
}

/* decimation-in-time recursion, KF:239-303 (kf_work): `ts` is the twiddle/input stride of this level */
static void cfft_level(cpx* out, const cpx* in, int ts, int stage, const cfft_plan* pl)
{
    const int p = pl->radix[stage], m = pl->rest[stage];
    if (m == 1) {
        for (int q = 0; q < p; q++) out[q] = in[(size_t)q * ts];
    } else {
        for (int q = 0; q < p; q++)
            cfft_level(out + (size_t)q * m, in + (size_t)q * ts, ts * p, stage + 1, pl);
    }
    switch (p) {
        case 2:  comb2(out, m, ts, pl); break;
        case 3:  comb3(out, m, ts, pl); break;
        case 4:  comb4(out, m, ts, pl); break;
        case 5:  comb5(out, m, ts, pl); break;
        default: comb_any(out, m, p, ts, pl); break;
    }
}

static void cfft_exec(const cfft_plan* pl, const cpx* in, cpx* out) { cfft_level(out, in, 1, 0, pl); }

/* ------------------------------------------------------------------------- */
/*  Real FFT of even length N via an N/2-point complex FFT                    */
/*  (restates KFR:29-67 kiss_fftr_alloc, KFR:69-123 kiss_fftr,                */
/*   KFR:125-161 kiss_fftri; wrapped as FFT:531-753 saf_rfft_*)               */
/* ------------------------------------------------------------------------- */

typedef struct {
    int N;                         /* real length (even) */
    cfft_plan *fwd, *bwd;
    cpx *work;                     /* N/2 */
    cpx *st_fwd, *st_bwd;          /* "super twiddles", N/4 each (KFR:59-65) */
} rfft_plan;

static rfft_plan* rfft_plan_new(int N)
{
    rfft_plan* r = (rfft_plan*)calloc(1, sizeof(rfft_plan));
    const int nc = N / 2;
    r->N = N;
    r->fwd = cfft_plan_new(nc, 0);
    r->bwd = cfft_plan_new(nc, 1);
    r->work = (cpx*)malloc(sizeof(cpx) * (size_t)nc);
    r->st_fwd = (cpx*)malloc(sizeof(cpx) * (size_t)(nc / 2 + 1));
    r->st_bwd = (cpx*)malloc(sizeof(cpx) * (size_t)(nc / 2 + 1));
    for (int i = 0; i < nc / 2; i++) {
        double ph = -3.14159265358979323846264338327 * ((double)(i + 1) / nc + .5);   /* KFR:60-61 */
        r->st_fwd[i].re = (float)cos(ph);  r->st_fwd[i].im = (float)sin(ph);
        ph *= -1;                                                                      /* KFR:62-63 */
        r->st_bwd[i].re = (float)cos(ph);  r->st_bwd[i].im = (float)sin(ph);
    }
    return r;
}

static void rfft_plan_free(rfft_plan* r)
{
    if (!r) return;
    cfft_plan_free(r->fwd); cfft_plan_free(r->bwd);
    free(r->work); free(r->st_fwd); free(r->st_bwd); free(r);
}

/* FFT:709-710 -> KFR:69-123 ; X has N/2+1 bins, unscaled */
static void rfft_forward(rfft_plan* r, const float* x, cpx* X)
{
    const int nc = r->N / 2;
    cfft_exec(r->fwd, (const cpx*)x, r->work);
    const float dr = r->work[0].re, di = r->work[0].im;
    X[0].re  = dr + di;  X[0].im  = 0.f;
    X[nc].re = dr - di;  X[nc].im = 0.f;
    for (int k = 1; k <= nc / 2; k++) {
        cpx a = r->work[k];
        cpx b = { r->work[nc - k].re, -r->work[nc - k].im };
        cpx s = c_add(a, b), d = c_sub(a, b);
        cpx t = c_mul(d, r->st_fwd[k - 1]);
        X[k].re      = (s.re + t.re) * .5f;
        X[k].im      = (s.im + t.im) * .5f;
        X[nc - k].re = (s.re - t.re) * .5f;
        X[nc - k].im = (t.im - s.im) * .5f;
    }
}

/* FFT:749-752 -> KFR:125-161 followed by scaling with 1/N ; imaginary parts of bins 0 and N/2 are ignored */
static void rfft_backward(rfft_plan* r, const cpx* X, float* x)
{
    const int nc = r->N / 2;
    r->work[0].re = X[0].re + X[nc].re;
    r->work[0].im = X[0].re - X[nc].re;
    for (int k = 1; k <= nc / 2; k++) {
        cpx a = X[k];
        cpx b = { X[nc - k].re, -X[nc - k].im };
        cpx e = c_add(a, b), d = c_sub(a, b);
        cpx o = c_mul(d, r->st_bwd[k - 1]);
        r->work[k]      = c_add(e, o);
        r->work[nc - k] = c_sub(e, o);
        r->work[nc - k].im *= -1;
    }
    cfft_exec(r->bwd, r->work, (cpx*)x);
    const float sc = 1.0f / (float)r->N;          /* FFT:751 cblas_sscal(N, 1/N) */
    for (int i = 0; i < r->N; i++) x[i] *= sc;
}

/* exported wrappers (used by the oracle-vs-reference FFT tests) */
void orc_rfft_forward(int N, const float* x, float* X)
{
    rfft_plan* r = rfft_plan_new(N);
    rfft_forward(r, x, (cpx*)X);
    rfft_plan_free(r);
}
void orc_rfft_backward(int N, const float* X, float* x)
{
    rfft_plan* r = rfft_plan_new(N);
    rfft_backward(r, (const cpx*)X, x);
    rfft_plan_free(r);
}

/* ------------------------------------------------------------------------- */
/*  Shared convolver engine                                                   */
/* ------------------------------------------------------------------------- */

/* utility_cvvmul, C99 branch (modules/saf_utilities/saf_utility_veclib.c:1196-1206) */
static void spec_mul(const cpx* a, const cpx* b, size_t n, cpx* c)
{
    for (size_t i = 0; i < n; i++) c[i] = c_mul(a[i], b[i]);
}

/* fftconv / fftfilt (modules/saf_utilities/saf_utility_fft.c:157-228): per-channel linear convolution through ONE
 * real FFT of size nextpow2(x_len + h_len - 1) (saf_utility_misc.c:64-80), spectra multiplied with utility_cvvmul */
void orc_fftconv(const float* x, const float* h, int x_len, int h_len, int nCH, float* y)
{
    const int y_len = x_len + h_len - 1;
    int fftSize = 1;
    do { fftSize *= 2; } while (fftSize < y_len);                 /* nextpow2: smallest 2^k >= y_len, k >= 1 */
    const int nBins = fftSize / 2 + 1;
    float* h0 = (float*)calloc((size_t)fftSize, sizeof(float));
    float* x0 = (float*)calloc((size_t)fftSize, sizeof(float));
    float* y0 = (float*)malloc((size_t)fftSize * sizeof(float));
    cpx* H = (cpx*)malloc((size_t)nBins * sizeof(cpx));
    cpx* X = (cpx*)malloc((size_t)nBins * sizeof(cpx));
    cpx* Y = (cpx*)malloc((size_t)nBins * sizeof(cpx));
    rfft_plan* r = rfft_plan_new(fftSize);
    for (int i = 0; i < nCH; i++) {
        memcpy(h0, h + (size_t)i * h_len, (size_t)h_len * sizeof(float));
        memcpy(x0, x + (size_t)i * x_len, (size_t)x_len * sizeof(float));
        rfft_forward(r, x0, X);
        rfft_forward(r, h0, H);
        spec_mul(X, H, (size_t)nBins, Y);
        rfft_backward(r, Y, y0);
        memcpy(y + (size_t)i * y_len, y0, (size_t)y_len * sizeof(float));
    }
    rfft_plan_free(r);
    free(h0); free(x0); free(y0); free(H); free(X); free(Y);
}

void orc_fftfilt(const float* x, const float* h, int x_len, int h_len, int nCH, float* y)
{
    const int y_len = x_len + h_len - 1;
    float* t = (float*)malloc((size_t)nCH * y_len * sizeof(float));
    orc_fftconv(x, h, x_len, h_len, nCH, t);
    for (int i = 0; i < nCH; i++) memcpy(y + (size_t)i * x_len, t + (size_t)i * y_len, (size_t)x_len * sizeof(float));
    free(t);
}

static int ceil_div_as_ref(int num, int den)      /* (int)ceilf((float)num/(float)den), MC:102 */
{
    return (int)ceilf((float)num / (float)den);
}

/* layout kinds */
enum { ORC_MATRIX = 0, ORC_MULTI = 1 };

typedef struct {
    int kind, part;
    int hop, N, nBins, len, nIn, nOut, P, nOLA;
    rfft_plan* fft;
    /* partitioned mode */
    cpx*  Hf;        /* matrix: [nOut][P][nIn][nBins] ; multi: [P][nCH][nBins] */
    cpx*  fdl;       /* [P][nIn][nBins], slot 0 = newest (MC:211) */
    cpx*  prod;      /* [P][nIn][nBins] */
    float* seg;      /* [P*nIn][N] */
    float* tail;     /* [nOut][hop] */
    /* non-partitioned mode */
    cpx*  Xf;        /* [nIn][nBins] */
    float* ola;      /* [nOut][N] */
    /* scratch */
    float* xpad;     /* N (partitioned) or nIn*N (non-partitioned matrix) */
    float* acc;      /* N */
    float* tmp;      /* N */
} orc_conv;

static void conv_free(orc_conv* h)
{
    if (!h) return;
    rfft_plan_free(h->fft);
    free(h->Hf); free(h->fdl); free(h->prod); free(h->seg); free(h->tail);
    free(h->Xf); free(h->ola); free(h->xpad); free(h->acc); free(h->tmp);
    free(h);
}

/* MC:49-130 (matrix) and MC:257-328 (multi).  For ORC_MULTI nIn == nOut == nCH and filters are [nCH][len]. */
static orc_conv* conv_new(int kind, int hop, const float* H, int len, int nIn, int nOut, int part)
{
    orc_conv* h = (orc_conv*)calloc(1, sizeof(orc_conv));
    h->kind = kind; h->part = part ? 1 : 0;
    h->hop = hop; h->len = len; h->nIn = nIn; h->nOut = nOut;
    const size_t nPairs = (kind == ORC_MATRIX) ? (size_t)nOut * nIn : (size_t)nOut;

    if (!h->part) {
        /* MC:71-96 / MC:277-298 : one FFT long enough for hop+len-1 samples, multiple of hop */
        h->nOLA  = (int)(ceilf((float)(hop + len - 1) / (float)hop) + 0.1f);
        h->N     = h->nOLA * hop;
        h->nBins = h->N / 2 + 1;
        h->fft   = rfft_plan_new(h->N);
        h->Hf    = (cpx*)malloc(sizeof(cpx) * nPairs * h->nBins);
        h->Xf    = (cpx*)calloc((size_t)nIn * h->nBins, sizeof(cpx));
        h->ola   = (float*)calloc((size_t)nOut * h->N, sizeof(float));
        h->xpad  = (float*)calloc((size_t)h->N, sizeof(float));
        h->acc   = (float*)malloc(sizeof(float) * h->N);
        h->tmp   = (float*)malloc(sizeof(float) * h->N);
        h->prod  = (cpx*)malloc(sizeof(cpx) * h->nBins);
        float* hp = (float*)calloc((size_t)h->N, sizeof(float));
        for (size_t q = 0; q < nPairs; q++) {
            memcpy(hp, H + q * len, sizeof(float) * (size_t)len);                   /* MC:91 */
            rfft_forward(h->fft, hp, h->Hf + q * h->nBins);                       /* MC:92 */
        }
        free(hp);
        return h;
    }

    /* MC:97-129 / MC:299-327 */
    h->N     = 2 * hop;
    h->nBins = hop + 1;
    h->P     = ceil_div_as_ref(len, hop);
    h->fft   = rfft_plan_new(h->N);
    const size_t slot = (size_t)nIn * h->nBins;              /* one FDL slot */
    h->Hf   = (cpx*)malloc(sizeof(cpx) * (kind == ORC_MATRIX ? (size_t)nOut : 1) * h->P * slot);
    h->fdl  = (cpx*)calloc((size_t)h->P * slot, sizeof(cpx));
    h->prod = (cpx*)malloc(sizeof(cpx) * (size_t)h->P * slot);
    h->seg  = (float*)malloc(sizeof(float) * (size_t)h->P * nIn * h->N);
    h->tail = (float*)calloc((size_t)nOut * hop, sizeof(float));
    h->xpad = (float*)calloc((size_t)h->N, sizeof(float));
    h->acc  = (float*)malloc(sizeof(float) * h->N);
    float* hp  = (float*)calloc((size_t)h->P * hop, sizeof(float));
    float* hp2 = (float*)calloc((size_t)h->N, sizeof(float));
    for (int no = 0; no < nOut; no++) {
        const int niEnd = (kind == ORC_MATRIX) ? nIn : 1;
        for (int j = 0; j < niEnd; j++) {
            const int ni = (kind == ORC_MATRIX) ? j : no;
            const float* src = (kind == ORC_MATRIX) ? H + ((size_t)no * nIn + ni) * len : H + (size_t)no * len;
            memcpy(hp, src, sizeof(float) * (size_t)len);                          /* MC:119 (tail of hp stays 0) */
            for (int p = 0; p < h->P; p++) {
                memcpy(hp2, hp + (size_t)p * hop, sizeof(float) * (size_t)hop);    /* MC:121 (upper half stays 0) */
                cpx* dst = (kind == ORC_MATRIX)
                         ? h->Hf + (size_t)no * h->P * slot + (size_t)p * slot + (size_t)ni * h->nBins   /* MC:122 */
                         : h->Hf + (size_t)p * slot + (size_t)ni * h->nBins;                             /* MC:321 */
                rfft_forward(h->fft, hp2, dst);
            }
        }
    }
    free(hp); free(hp2);
    return h;
}

/* MC:163-236 (matrix) and MC:357-414 (multi) */
static void conv_apply(orc_conv* h, const float* in, float* out)
{
    const int hop = h->hop, N = h->N, nb = h->nBins;

    if (!h->part) {
        /* forward transforms of the zero-padded inputs, MC:176-179 / MC:370-373 */
        for (int ni = 0; ni < h->nIn; ni++) {
            memcpy(h->xpad, in + (size_t)ni * hop, sizeof(float) * (size_t)hop);
            rfft_forward(h->fft, h->xpad, h->Xf + (size_t)ni * nb);
        }
        for (int no = 0; no < h->nOut; no++) {
            float* ola = h->ola + (size_t)no * N;
            if (h->kind == ORC_MATRIX) {
                /* MC:186-195 : per-input product, inverse FFT, time-domain sum in input order */
                memset(h->acc, 0, sizeof(float) * (size_t)N);
                for (int ni = 0; ni < h->nIn; ni++) {
                    spec_mul(h->Hf + ((size_t)no * h->nIn + ni) * nb, h->Xf + (size_t)ni * nb, (size_t)nb, h->prod);
                    rfft_backward(h->fft, h->prod, h->tmp);
                    for (int i = 0; i < N; i++) h->acc[i] += h->tmp[i];
                }
            } else {
                /* MC:376-378 */
                spec_mul(h->Hf + (size_t)no * nb, h->Xf + (size_t)no * nb, (size_t)nb, h->prod);
                rfft_backward(h->fft, h->prod, h->acc);
            }
            /* shift the overlap-add buffer by one hop, clear its last hop, add, emit first hop:
             * MC:198-205 / MC:381-384 */
            memmove(ola, ola + hop, sizeof(float) * (size_t)(h->nOLA - 1) * hop);
            memset(ola + (size_t)(h->nOLA - 1) * hop, 0, sizeof(float) * (size_t)hop);
            for (int i = 0; i < N; i++) ola[i] += h->acc[i];
            memcpy(out + (size_t)no * hop, ola, sizeof(float) * (size_t)hop);
        }
        return;
    }

    const size_t slot = (size_t)h->nIn * nb;
    /* age the delay line by one slot and transform the new block into slot 0, MC:211-215 / MC:390-394 */
    memmove(h->fdl + slot, h->fdl, sizeof(cpx) * (size_t)(h->P - 1) * slot);
    for (int ni = 0; ni < h->nIn; ni++) {
        memcpy(h->xpad, in + (size_t)ni * hop, sizeof(float) * (size_t)hop);
        rfft_forward(h->fft, h->xpad, h->fdl + (size_t)ni * nb);
    }

    if (h->kind == ORC_MULTI)
        spec_mul(h->Hf, h->fdl, (size_t)h->P * slot, h->prod);                            /* MC:397 */

    for (int no = 0; no < h->nOut; no++) {
        memset(h->acc, 0, sizeof(float) * (size_t)N);
        if (h->kind == ORC_MATRIX) {
            spec_mul(h->Hf + (size_t)no * h->P * slot, h->fdl, (size_t)h->P * slot, h->prod);   /* MC:219 */
            const int nseg = h->P * h->nIn;
            for (int s = 0; s < nseg; s++)                                                /* MC:220-222 */
                rfft_backward(h->fft, h->prod + (size_t)s * nb, h->seg + (size_t)s * N);
            for (int s = 0; s < nseg; s++)                                                /* MC:225-227 */
                for (int i = 0; i < N; i++) h->acc[i] += h->seg[(size_t)s * N + i];
        } else {
            for (int p = 0; p < h->P; p++)                                                /* MC:399-400 */
                rfft_backward(h->fft, h->prod + (size_t)p * slot + (size_t)no * nb, h->seg + (size_t)p * N);
            for (int p = 0; p < h->P; p++)                                                /* MC:403-405 */
                for (int i = 0; i < N; i++) h->acc[i] += h->seg[(size_t)p * N + i];
        }
        float* tail = h->tail + (size_t)no * hop;
        for (int i = 0; i < hop; i++) out[(size_t)no * hop + i] = h->acc[i] + tail[i];    /* MC:230 / MC:408 */
        memcpy(tail, h->acc + hop, sizeof(float) * (size_t)hop);                          /* MC:233 / MC:411 */
    }
}

/* exported C API of the oracle: same argument order and meaning as the reference API
 * (modules/saf_utilities/saf_utility_matrixConv.h:55-86,109-136), returning the handle */
void* orc_matrixConv_create(int hop, const float* H, int len, int nIn, int nOut, int usePart)
{ return conv_new(ORC_MATRIX, hop, H, len, nIn, nOut, usePart); }
void  orc_matrixConv_apply(void* h, const float* in, float* out) { conv_apply((orc_conv*)h, in, out); }
void  orc_matrixConv_destroy(void* h) { conv_free((orc_conv*)h); }

void* orc_multiConv_create(int hop, const float* H, int len, int nCH, int usePart)
{ return conv_new(ORC_MULTI, hop, H, len, nCH, nCH, usePart); }
void  orc_multiConv_apply(void* h, const float* in, float* out) { conv_apply((orc_conv*)h, in, out); }
void  orc_multiConv_destroy(void* h) { conv_free((orc_conv*)h); }

/* ------------------------------------------------------------------------- */
/*  Time-varying convolver (MC:423-620), 1 input -> nOut, nIRs selectable     */
/* ------------------------------------------------------------------------- */

typedef struct {
    int hop, N, nBins, len, nIRs, nOut, P;
    int idx1, idx2;                       /* posIdx_last, posIdx_last2 */
    rfft_plan* fft;
    cpx*  Hf;                             /* [nIRs][nOut][P][nBins] */
    cpx  *fdl, *prod;                     /* [P][nBins] */
    float *seg;                           /* [P][N] */
    float *z0, *z1, *z2;                  /* N each */
    float *tail0, *tail1;                 /* [nOut][hop] : y_n_overlap, y_n_overlap_last */
    float *fin, *fout, *xpad;
} orc_tv;

/* MC:441-512 ; H is given FLAT here as [nIRs][nOut][len] (the reference takes float** with one FLAT nOut x len row per IR) */
void* orc_TVConv_create(int hop, const float* H, int len, int nIRs, int nOut, int initIdx)
{
    orc_tv* h = (orc_tv*)calloc(1, sizeof(orc_tv));
    h->hop = hop; h->len = len; h->nIRs = nIRs; h->nOut = nOut;
    h->idx1 = h->idx2 = (initIdx < nIRs) ? initIdx : 0;                /* MC:459-465 */
    h->N = 2 * hop; h->nBins = hop + 1; h->P = ceil_div_as_ref(len, hop);
    h->fft  = rfft_plan_new(h->N);
    const size_t pb = (size_t)h->P * h->nBins;
    h->Hf   = (cpx*)malloc(sizeof(cpx) * (size_t)nIRs * nOut * pb);
    h->fdl  = (cpx*)calloc(pb, sizeof(cpx));
    h->prod = (cpx*)malloc(sizeof(cpx) * pb);
    h->seg  = (float*)malloc(sizeof(float) * (size_t)h->P * h->N);
    h->z0 = (float*)malloc(sizeof(float) * h->N);
    h->z1 = (float*)malloc(sizeof(float) * h->N);
    h->z2 = (float*)malloc(sizeof(float) * h->N);
    h->tail0 = (float*)calloc((size_t)nOut * hop, sizeof(float));
    h->tail1 = (float*)calloc((size_t)nOut * hop, sizeof(float));
    h->fin  = (float*)malloc(sizeof(float) * hop);
    h->fout = (float*)malloc(sizeof(float) * hop);
    h->xpad = (float*)calloc((size_t)h->N, sizeof(float));
    for (int n = 0; n < hop; n++) {                                     /* MC:494-497 */
        h->fin[n]  = (float)n / (float)(hop - 1);
        h->fout[n] = (float)(hop - 1 - n) / (float)(hop - 1);
    }
    float* hp  = (float*)calloc((size_t)h->P * hop, sizeof(float));
    float* hp2 = (float*)calloc((size_t)h->N, sizeof(float));
    for (int ir = 0; ir < nIRs; ir++)
        for (int no = 0; no < nOut; no++) {
            memcpy(hp, H + ((size_t)ir * nOut + no) * len, sizeof(float) * (size_t)len);   /* MC:502 */
            for (int p = 0; p < h->P; p++) {
                memcpy(hp2, hp + (size_t)p * hop, sizeof(float) * (size_t)hop);            /* MC:504 */
                rfft_forward(h->fft, hp2, h->Hf + ((size_t)ir * nOut + no) * pb + (size_t)p * h->nBins);
            }
        }
    free(hp); free(hp2);
    return h;
}

void orc_TVConv_destroy(void* hv)
{
    orc_tv* h = (orc_tv*)hv;
    if (!h) return;
    rfft_plan_free(h->fft);
    free(h->Hf); free(h->fdl); free(h->prod); free(h->seg);
    free(h->z0); free(h->z1); free(h->z2); free(h->tail0); free(h->tail1);
    free(h->fin); free(h->fout); free(h->xpad); free(h);
}

/* one filtered 2*hop frame for IR `ir`, output `no`: MC:561-569 */
static void tv_frame(orc_tv* h, int ir, int no, float* z)
{
    const size_t pb = (size_t)h->P * h->nBins;
    spec_mul(h->Hf + ((size_t)ir * h->nOut + no) * pb, h->fdl, pb, h->prod);
    for (int p = 0; p < h->P; p++)
        rfft_backward(h->fft, h->prod + (size_t)p * h->nBins, h->seg + (size_t)p * h->N);
    memset(z, 0, sizeof(float) * (size_t)h->N);
    for (int p = 0; p < h->P; p++)
        for (int i = 0; i < h->N; i++) z[i] += h->seg[(size_t)p * h->N + i];
}

/* MC:546-620 */
void orc_TVConv_apply(void* hv, const float* in, float* out, int irIdx)
{
    orc_tv* h = (orc_tv*)hv;
    const int hop = h->hop;
    memmove(h->fdl + h->nBins, h->fdl, sizeof(cpx) * (size_t)(h->P - 1) * h->nBins);   /* MC:557 */
    memcpy(h->xpad, in, sizeof(float) * (size_t)hop);                                  /* MC:559 */
    rfft_forward(h->fft, h->xpad, h->fdl);                                             /* MC:560 */
    for (int no = 0; no < h->nOut; no++) {
        tv_frame(h, irIdx, no, h->z0);
        if (irIdx != h->idx1) tv_frame(h, h->idx1, no, h->z1);                         /* MC:575-585 */
        else memcpy(h->z1, h->z0, sizeof(float) * (size_t)h->N);                       /* MC:587 */
        if (h->idx1 != h->idx2) tv_frame(h, h->idx2, no, h->z2);                       /* MC:589-599 */
        else memcpy(h->z2, h->z1, sizeof(float) * (size_t)h->N);                       /* MC:601 */
        float* t0 = h->tail0 + (size_t)no * hop;
        float* t1 = h->tail1 + (size_t)no * hop;
        for (int i = 0; i < hop; i++) {                                                /* MC:605-611 */
            const float o1 = h->z1[i] + t0[i];
            const float o2 = h->z2[i] + t1[i];
            const float a = o1 * h->fin[i];
            const float b = o2 * h->fout[i];
            out[(size_t)no * hop + i] = a + b;
        }
        memcpy(t0, h->z0 + hop, sizeof(float) * (size_t)hop);                          /* MC:614 */
        memcpy(t1, h->z1 + hop, sizeof(float) * (size_t)hop);                          /* MC:615 */
    }
    h->idx2 = h->idx1;                                                                 /* MC:618-619 */
    h->idx1 = irIdx;
}

/* ------------------------------------------------------------------------- */
/*  fp64 ground truth: direct time-domain convolution                          */
/*  y[no][n] = sum_ni sum_k h[no][ni][k] * x[ni][n-k]          */
/* ------------------------------------------------------------------------- */

/* matrix form; x is [nIn][T]; evaluates samples n in [n0, n1) for the listed output channels */
void orc_truth_matrix(const float* H, int len, int nIn, int nOut, const float* x, long T,
                      const int* outs, int nOuts, long n0, long n1, double* y /* [nOuts][n1-n0] */)
{
    (void)nOut;
    for (int j = 0; j < nOuts; j++) {
        const int no = outs[j];
        for (long n = n0; n < n1; n++) {
            double acc = 0.0;
            for (int ni = 0; ni < nIn; ni++) {
                const float* hh = H + ((size_t)no * nIn + ni) * len;
                const float* xx = x + (size_t)ni * T;
                const long kmax = (n < len - 1) ? n : (len - 1);
                double a = 0.0;
                for (long k = 0; k <= kmax; k++) a += (double)hh[k] * (double)xx[n - k];
                acc += a;
            }
            y[(size_t)j * (n1 - n0) + (n - n0)] = acc;
        }
    }
}

/* diagonal (multiConv) form; H is [nCH][len], x is [nCH][T] */
void orc_truth_multi(const float* H, int len, int nCH, const float* x, long T,
                     const int* chans, int nChans, long n0, long n1, double* y /* [nChans][n1-n0] */)
{
    (void)nCH;
    for (int j = 0; j < nChans; j++) {
        const int c = chans[j];
        const float* hh = H + (size_t)c * len;
        const float* xx = x + (size_t)c * T;
        for (long n = n0; n < n1; n++) {
            const long kmax = (n < len - 1) ? n : (len - 1);
            double a = 0.0;
            for (long k = 0; k <= kmax; k++) a += (double)hh[k] * (double)xx[n - k];
            y[(size_t)j * (n1 - n0) + (n - n0)] = a;
        }
    }
}
