/*
 * safconv_prod_core.cuh -- per-element arithmetic of the filter producers, __host__ __device__ (the kernels of
 * safconv_producers.cu call these; tests/test_producers_host.py compiles the same header into a host program).
 *
 *   image-source geometry   /root/reference/framework/modules/saf_reverb/saf_reverb_internal.c:269-397 (coreInitT),
 *                           :399-521 (coreInitN), :640-685 (tap index of renderRIR)
 *   2 x 2 covariance match  /root/reference/framework/modules/saf_hoa/saf_hoa.c:497-604 (applyDiffCovMatching)
 *
 * The geometry must reproduce the reference's fp32 results BIT FOR BIT: an image source lands on tap
 * (int)(time * fs + 0.5f), so one ulp in the distance can move a reflection by a sample.  The reference is compiled for
 * baseline x86-64 (no fused multiply-add, IEEE sqrt / divide), so every step is an explicitly rounded fp32 operation
 * here (__fmul_rn / __fadd_rn never contract into an FMA); host builds of this header use -ffp-contract=off.
 */
#ifndef SAFCONV_PROD_CORE_CUH_INCLUDED
#define SAFCONV_PROD_CORE_CUH_INCLUDED

#include <math.h>

#if defined(__CUDACC__)
#define SCP_HD __host__ __device__ __forceinline__
#else
#define SCP_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define SCP_MUL(a, b) __fmul_rn((a), (b))
#define SCP_ADD(a, b) __fadd_rn((a), (b))
#define SCP_SUB(a, b) __fsub_rn((a), (b))
#define SCP_DIV(a, b) __fdiv_rn((a), (b))
#define SCP_SQRT(a)   __fsqrt_rn((a))
#else
#define SCP_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define SCP_ADD(a, b) ((float)((float)(a) + (float)(b)))
#define SCP_SUB(a, b) ((float)((float)(a) - (float)(b)))
#define SCP_DIV(a, b) ((float)((float)(a) / (float)(b)))
#define SCP_SQRT(a)   sqrtf((a))
#endif

/* one source / receiver pair of an image-source scene (device-visible, plain data) */
typedef struct ScpImsPair {
    float room[3];          /* shoebox dimensions                                                                   */
    float so[3], ro[3];     /* source / receiver relative to the room centre (coreInit :288-294)                    */
    float c_ms, fs, dmax;   /* speed of sound, sample rate, maximum distance (T mode)                               */
    int   mode;             /* 0: all images closer than dmax (coreInitT); 1: all images up to order maxN (coreInitN) */
    int   Nx, Ny, Nz;       /* lattice half-widths (N mode: all three = maxN)                                       */
    long long lengthVec;    /* (2Nx+1)(2Ny+1)(2Nz+1)                                                                */
    int   order, nSH;       /* receiver SH order / channels                                                          */
    int   len;              /* RIR length in samples (known after the count pass)                                   */
    long long accOff;       /* offset of this pair's [nSH][len] block in the fp64 accumulator (global-atomics render) */
    int   tw;               /* taps per window of the windowed render (a multiple of 32)                            */
    int   pad_;
} ScpImsPair;

/* lattice point q -> reflection orders (i fastest, then j, then k: the order of :304-321 / :431-449) */
SCP_HD void scp_ims_lattice(const ScpImsPair* p, long long q, int* ii, int* jj, int* kk)
{
    const long long wx = 2 * p->Nx + 1, wy = 2 * p->Ny + 1;
    *ii = (int)(q % wx) - p->Nx;
    *jj = (int)((q / wx) % wy) - p->Ny;
    *kk = (int)(q / (wx * wy)) - p->Nz;
}

/* image source of lattice point (ii, jj, kk): position relative to the receiver and distance; returns 1 if the image
 * belongs to the echogram (:365-372 / :453-459). */
SCP_HD int scp_ims_image(const ScpImsPair* p, int ii, int jj, int kk, float* sx, float* sy, float* sz, float* d)
{
    if (p->mode == 1) {
        const int ord = (ii < 0 ? -ii : ii) + (jj < 0 ? -jj : jj) + (kk < 0 ? -kk : kk);
        if (ord > p->Nx) return 0;
    }
    const float sgx = (ii & 1) ? -1.0f : 1.0f, sgy = (jj & 1) ? -1.0f : 1.0f, sgz = (kk & 1) ? -1.0f : 1.0f;   /* powf(-1, n) */
    const float x = SCP_SUB(SCP_ADD(SCP_MUL((float)ii, p->room[0]), SCP_MUL(sgx, p->so[0])), p->ro[0]);
    const float y = SCP_SUB(SCP_ADD(SCP_MUL((float)jj, p->room[1]), SCP_MUL(sgy, p->so[1])), p->ro[1]);
    const float z = SCP_SUB(SCP_ADD(SCP_MUL((float)kk, p->room[2]), SCP_MUL(sgz, p->so[2])), p->ro[2]);
    const float dd = SCP_SQRT(SCP_ADD(SCP_ADD(SCP_MUL(x, x), SCP_MUL(y, y)), SCP_MUL(z, z)));
    *sx = x; *sy = y; *sz = z; *d = dd;
    if (p->mode == 0 && !(dd < p->dmax)) return 0;
    return 1;
}

/* propagation time, tap index (:676) and distance attenuation (:381-383) of an image at distance d */
SCP_HD int scp_ims_tap(const ScpImsPair* p, float d, float* time, float* att)
{
    const float t = SCP_DIV(d, p->c_ms);
    *time = t;
    *att = (d <= 1.0f) ? 1.0f : SCP_DIV(1.0f, d);
    return (int)SCP_ADD(SCP_MUL(t, p->fs), 0.5f);
}

/* RIR length from the latest arrival (:655-657): (int)(endtime * fs + 1.0f) + 1 */
SCP_HD int scp_ims_length(const ScpImsPair* p, float dLast)
{
    const float endtime = SCP_DIV(dLast, p->c_ms);
    return (int)SCP_ADD(SCP_MUL(endtime, p->fs), 1.0f) + 1;
}

/* direction of an image as the receiver module sees it (:556-559): unitCart2sph (saf_utility_geometry.c:351-364), then
 * elevation -> inclination */
SCP_HD void scp_ims_direction(float sx, float sy, float sz, float* azi, float* incl)
{
    *azi = atan2f(sy, sx);
    const float elev = atan2f(sz, SCP_SQRT(SCP_ADD(SCP_MUL(sx, sx), SCP_MUL(sy, sy))));
    *incl = SCP_SUB(3.14159265358979323846264338327950288f / 2.0f, elev);
}

/* ---- windowed render: which lattice points can land in the taps [w0, w0 + tw) --------------------------------------
 * tap = (int)(d / c * fs + 0.5f) in fp32, so an image of the window has (w0 - 0.5) c / fs <= d < (w0 + tw - 0.5) c / fs up
 * to a few fp32 roundings (relative 1e-7; the + 0.5f at 96 000 taps rounds by up to 0.004 taps = 4e-8 of d).  The range
 * returned here is widened by 1e-5 relative + 1e-4 m: CONSERVATIVE -- every image of the window is inside, a few outside
 * are too, and the exact tap test of the caller decides. */
SCP_HD void scp_ims_window_range(const ScpImsPair* p, int w0, int tw, double* dlo, double* dhi)
{
    const double k = (double)p->c_ms / (double)p->fs;
    double lo = ((double)w0 - 0.5) * k, hi = ((double)(w0 + tw) - 0.5) * k;
    lo = lo * (1.0 - 1e-5) - 1e-4; hi = hi * (1.0 + 1e-5) + 1e-4;
    *dlo = lo > 0.0 ? lo : 0.0; *dhi = hi;
}

/* rows (jj, kk) that can hold an image closer than dhi: |jj| <= *jr, |kk| <= *kr (clamped to the lattice) */
SCP_HD void scp_ims_window_rows(const ScpImsPair* p, double dhi, int* jr, int* kr)
{
    const double ey = fabs((double)p->so[1]) + fabs((double)p->ro[1]), ez = fabs((double)p->so[2]) + fabs((double)p->ro[2]);
    int j = (int)ceil((dhi + ey) / (double)p->room[1]) + 1, k = (int)ceil((dhi + ez) / (double)p->room[2]) + 1;
    *jr = j < p->Ny ? j : p->Ny; *kr = k < p->Nz ? k : p->Nz;
}

/* candidates of lattice row (jj, kk) for the distance range [dlo, dhi): up to FOUR inclusive ii ranges of stride 2
 * (lo[s] > hi[s]: empty), pairwise disjoint, clamped to the lattice -- the even and the odd reflection orders are two
 * arithmetic progressions in x (x = ii Lx + so.x - ro.x for even ii, ii Lx - so.x - ro.x for odd ii), and the shell cuts
 * the row in one interval (|x| < xb) or two (xa <= |x| < xb).  Returns the number of ranges written (0, 2 or 4). */
SCP_HD int scp_ims_row_ranges(const ScpImsPair* p, int jj, int kk, double dlo, double dhi, int lo[4], int hi[4])
{
    const float sgy = (jj & 1) ? -1.0f : 1.0f, sgz = (kk & 1) ? -1.0f : 1.0f;
    const double y = (double)SCP_SUB(SCP_ADD(SCP_MUL((float)jj, p->room[1]), SCP_MUL(sgy, p->so[1])), p->ro[1]);
    const double z = (double)SCP_SUB(SCP_ADD(SCP_MUL((float)kk, p->room[2]), SCP_MUL(sgz, p->so[2])), p->ro[2]);
    const double r2 = y * y + z * z;
    const double hi2 = dhi * dhi - r2;
    if (!(hi2 > 0.0)) return 0;
    const double xb = sqrt(hi2) * (1.0 + 1e-7) + 1e-7;
    const double lo2 = dlo * dlo - r2;
    double xa = lo2 > 0.0 ? sqrt(lo2) * (1.0 - 1e-7) - 1e-7 : 0.0;
    if (xa < 1e-3) xa = 0.0;                        /* the two intervals would (nearly) touch: one interval */
    const double Lx = (double)p->room[0];
    const int N = p->Nx;
    int n = 0;
    for (int par = 0; par < 2; par++) {
        const double off = (par ? -(double)p->so[0] : (double)p->so[0]) - (double)p->ro[0];
        for (int side = 0; side < (xa > 0.0 ? 2 : 1); side++) {
            const double xl = (xa > 0.0) ? (side ? xa : -xb) : -xb, xh = (xa > 0.0) ? (side ? xb : -xa) : xb;
            int a = (int)ceil((xl - off) / Lx), b = (int)floor((xh - off) / Lx);
            if (a < -N) a = -N;
            if (b > N) b = N;
            if ((a & 1) != par) a++;
            lo[n] = a; hi[n] = b; n++;
        }
    }
    return n;
}

/* ---- 2 x 2 diffuse-field covariance matching (fp64) ------------------------------------------------------------ */
typedef struct scp_cd { double re, im; } scp_cd;
SCP_HD scp_cd scp_c(double re, double im) { scp_cd r; r.re = re; r.im = im; return r; }
SCP_HD scp_cd scp_cmul(scp_cd a, scp_cd b) { return scp_c(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
SCP_HD scp_cd scp_cmulc(scp_cd a, scp_cd b) { return scp_c(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im); }   /* a * conj(b) */
SCP_HD scp_cd scp_cadd(scp_cd a, scp_cd b) { return scp_c(a.re + b.re, a.im + b.im); }
SCP_HD scp_cd scp_conj(scp_cd a) { return scp_c(a.re, -a.im); }
SCP_HD scp_cd scp_cscale(scp_cd a, double s) { return scp_c(a.re * s, a.im * s); }

/* C = [[c00, c01], [conj(c01), c11]] Hermitian -> upper factor X with X^H X = C (utility_cchol, veclib.c:4103-4160).
 * Returns 0 if C is not positive definite. */
SCP_HD int scp_chol2(double c00, scp_cd c01, double c11, scp_cd X[2][2])
{
    if (!(c00 > 0.0)) return 0;
    const double x00 = sqrt(c00);
    const scp_cd x01 = scp_cscale(c01, 1.0 / x00);
    const double r = c11 - (x01.re * x01.re + x01.im * x01.im);
    if (!(r > 0.0)) return 0;
    X[0][0] = scp_c(x00, 0.0); X[0][1] = x01; X[1][0] = scp_c(0.0, 0.0); X[1][1] = scp_c(sqrt(r), 0.0);
    return 1;
}

SCP_HD void scp_mm2(const scp_cd A[2][2], const scp_cd B[2][2], scp_cd C[2][2])
{
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) C[i][j] = scp_cadd(scp_cmul(A[i][0], B[0][j]), scp_cmul(A[i][1], B[1][j]));
}

/* M of applyDiffCovMatching (:556-593): with X = chol(C_ref), Xa = chol(C_ambi), U S V^H = svd(Xa^H X):
 * M = Xa^-1 (V U^H X).  V U^H is the unitary polar factor of (Xa^H X)^H, which is unique for a non-singular matrix, so
 * it is formed here without an SVD: Q = A^H (A A^H)^(-1/2), the inverse square root of the 2 x 2 Hermitian matrix in
 * closed form.  The decoder of the band then becomes M^H D.  Returns 0 (M = identity) if a factor does not exist. */
SCP_HD int scp_diffcov_M(double r00, scp_cd r01, double r11, double a00, scp_cd a01, double a11, scp_cd M[2][2])
{
    scp_cd X[2][2], Xa[2][2], A[2][2], P[2][2];
    M[0][0] = scp_c(1.0, 0.0); M[0][1] = scp_c(0.0, 0.0); M[1][0] = scp_c(0.0, 0.0); M[1][1] = scp_c(1.0, 0.0);
    if (!scp_chol2(r00, r01, r11, X) || !scp_chol2(a00, a01, a11, Xa)) return 0;
    /* A = Xa^H X */
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            A[i][j] = scp_cadd(scp_cmul(scp_conj(Xa[0][i]), X[0][j]), scp_cmul(scp_conj(Xa[1][i]), X[1][j]));
    /* P = A A^H (Hermitian positive definite) */
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) P[i][j] = scp_cadd(scp_cmulc(A[i][0], A[j][0]), scp_cmulc(A[i][1], A[j][1]));
    const double p00 = P[0][0].re, p11 = P[1][1].re;
    const double det = p00 * p11 - (P[0][1].re * P[0][1].re + P[0][1].im * P[0][1].im);
    if (!(det > 0.0)) return 0;
    const double s = sqrt(det), t = sqrt(p00 + p11 + 2.0 * s);
    if (!(t > 0.0)) return 0;
    /* sqrt(P) = (P + s I) / t;  its inverse = adj / det_sqrt, det_sqrt = s */
    const double q00 = (p00 + s) / t, q11 = (p11 + s) / t;
    const scp_cd q01 = scp_cscale(P[0][1], 1.0 / t);
    const double dq = q00 * q11 - (q01.re * q01.re + q01.im * q01.im);
    if (!(dq > 0.0)) return 0;
    scp_cd R[2][2];                                   /* (A A^H)^(-1/2) */
    R[0][0] = scp_c(q11 / dq, 0.0); R[1][1] = scp_c(q00 / dq, 0.0);
    R[0][1] = scp_cscale(q01, -1.0 / dq); R[1][0] = scp_conj(R[0][1]);
    scp_cd AH[2][2], Q[2][2], QX[2][2];
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) AH[i][j] = scp_conj(A[j][i]);
    scp_mm2(AH, R, Q);
    scp_mm2(Q, X, QX);
    /* M = Xa^-1 QX, Xa upper triangular */
    const double i11 = 1.0 / Xa[1][1].re, i00 = 1.0 / Xa[0][0].re;
    for (int j = 0; j < 2; j++) {
        M[1][j] = scp_cscale(QX[1][j], i11);
        const scp_cd tmp = scp_cmul(Xa[0][1], M[1][j]);
        M[0][j] = scp_cscale(scp_c(QX[0][j].re - tmp.re, QX[0][j].im - tmp.im), i00);
    }
    return 1;
}

#endif
