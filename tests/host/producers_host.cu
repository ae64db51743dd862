/* Host build of the filter producers' arithmetic (tests/test_producers_host.py): the __host__ __device__ functions of
 * csrc/safconv_sh.cuh and csrc/safconv_prod_core.cuh -- the very code the kernels of safconv_producers.cu call -- driven by
 * plain loops that walk the lattice / the directions in the order a grid-stride kernel would.  No GPU needed. */
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../spatial_audio_framework_b200/csrc/safconv_sh.cuh"
#include "../../spatial_audio_framework_b200/csrc/safconv_prod_core.cuh"

extern "C" {

void ph_rsh(int order, const float* dirs_deg, int nD, float* Y)
{
    for (int d = 0; d < nD; d++) scsh_rsh_dir(order, dirs_deg[2 * d], dirs_deg[2 * d + 1], Y + d, nD);
}

/* dirs_rad [nD][2] = (azimuth, inclination); Y [(order+1)^2][nD] */
void ph_shreal_recur(int order, const float* dirs_rad, int nD, float* Y)
{
    float c[SCSH_MAX_ORDER + 1][SCSH_MAX_ORDER + 1];
    scsh_recur_norms(SCSH_MAX_ORDER, c);
    const int n = (order + 1) * (order + 1);
    float Yv[(SCSH_MAX_ORDER + 1) * (SCSH_MAX_ORDER + 1)];
    for (int d = 0; d < nD; d++) {
        if (order <= 3)      scsh_shreal_recur_dir<3>(order, dirs_rad[2 * d], dirs_rad[2 * d + 1], &c[0][0], Yv);
        else if (order <= 7) scsh_shreal_recur_dir<7>(order, dirs_rad[2 * d], dirs_rad[2 * d + 1], &c[0][0], Yv);
        else                 scsh_shreal_recur_dir<SCSH_MAX_ORDER>(order, dirs_rad[2 * d], dirs_rad[2 * d + 1], &c[0][0], Yv);
        for (int q = 0; q < n; q++) Y[(size_t)q * nD + d] = Yv[q];
    }
}

double ph_legendre(int n, double x) { return scsh_legendre(n, x); }

/* M (2 x 2 complex, row-major re/im pairs) of the diffuse-field covariance matching; returns scp_diffcov_M's flag */
int ph_diffcov_M(const double* cref /* c00, re c01, im c01, c11 */, const double* camb, double* M)
{
    scp_cd m[2][2];
    const int ok = scp_diffcov_M(cref[0], scp_c(cref[1], cref[2]), cref[3], camb[0], scp_c(camb[1], camb[2]), camb[3], m);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) { M[(i * 2 + j) * 2] = m[i][j].re; M[(i * 2 + j) * 2 + 1] = m[i][j].im; }
    return ok;
}

/* One source / receiver pair through the two passes of ims_shoebox_renderRIRs exactly as safconv_producers.c /
 * safconv_producers.cu sequence them.  pair: room, so, ro, c_ms, fs, dmax, mode, Nx..Nz, lengthVec, order, nSH filled by
 * the caller (the Python test mirrors the host layer's fp32 set-up); absTab [3][nBands][maxW].  First call with rir == NULL
 * returns the length and image count; second call fills rir [nSH][len] (fp64 taps rounded to fp32) and taps[nImages]. */
int ph_ims_pair(ScpImsPair* pair, const float* absTab, int nBands, int maxW, float* rir, int* nImages, int* taps)
{
    ScpImsPair p = *pair;
    if (!rir) {
        unsigned int cnt = 0; float dLast = 0.0f;
        for (long long q = 0; q < p.lengthVec; q++) {
            int ii, jj, kk; float sx, sy, sz, d;
            scp_ims_lattice(&p, q, &ii, &jj, &kk);
            if (scp_ims_image(&p, ii, jj, kk, &sx, &sy, &sz, &d)) { cnt++; if (d > dLast) dLast = d; }
        }
        *nImages = (int)cnt;
        if (!cnt) return 0;
        pair->len = scp_ims_length(&p, dLast);
        return pair->len;
    }
    float c[SCSH_MAX_ORDER + 1][SCSH_MAX_ORDER + 1];
    scsh_recur_norms(SCSH_MAX_ORDER, c);
    std::vector<double> acc((size_t)p.nSH * p.len, 0.0);
    const float *tx = absTab, *ty = absTab + (size_t)nBands * maxW, *tz = absTab + 2 * (size_t)nBands * maxW;
    int n = 0;
    for (long long q = 0; q < p.lengthVec; q++) {
        int ii, jj, kk; float sx, sy, sz, d;
        scp_ims_lattice(&p, q, &ii, &jj, &kk);
        if (!scp_ims_image(&p, ii, jj, kk, &sx, &sy, &sz, &d)) continue;
        float time, att;
        const int tap = scp_ims_tap(&p, d, &time, &att);
        if (taps) taps[n] = tap;
        n++;
        if (tap < 0 || tap >= p.len) continue;
        double tot = 0.0;
        for (int b = 0; b < nBands; b++) tot += (double)(tx[b * maxW + ii + p.Nx] * ty[b * maxW + jj + p.Ny] * tz[b * maxW + kk + p.Nz]);
        if (p.order == 0) { acc[tap] += (double)att * tot; continue; }
        float azi, incl, Yv[(SCSH_MAX_ORDER + 1) * (SCSH_MAX_ORDER + 1)];
        scp_ims_direction(sx, sy, sz, &azi, &incl);
        if (p.order <= 3)      scsh_shreal_recur_dir<3>(p.order, azi, incl, &c[0][0], Yv);
        else if (p.order <= 7) scsh_shreal_recur_dir<7>(p.order, azi, incl, &c[0][0], Yv);
        else                   scsh_shreal_recur_dir<SCSH_MAX_ORDER>(p.order, azi, incl, &c[0][0], Yv);
        for (int ch = 0; ch < p.nSH; ch++) acc[(size_t)ch * p.len + tap] += (double)(Yv[ch] * att) * tot;
    }
    for (size_t e = 0; e < acc.size(); e++) rir[e] = (float)acc[e];
    *nImages = n;
    return p.len;
}

/* The windowed render (ims_window_kernel), one "CTA" after the other: every window of pair->tw taps enumerates its
 * candidates with scp_ims_window_range / _rows / scp_ims_row_ranges, decides them with the exact geometry and adds them to
 * its own taps.  Returns the number of images processed over all windows (must equal the count pass: every image is
 * found by exactly one window); candidates = lattice points the windows looked at (efficiency of the enumeration). */
long long ph_ims_pair_windows(const ScpImsPair* pair, const float* absTab, int nBands, int maxW, float* rir, long long* candidates)
{
    const ScpImsPair p = *pair;
    float c[SCSH_MAX_ORDER + 1][SCSH_MAX_ORDER + 1];
    scsh_recur_norms(SCSH_MAX_ORDER, c);
    const float *tx = absTab, *ty = absTab + (size_t)nBands * maxW, *tz = absTab + 2 * (size_t)nBands * maxW;
    long long nImg = 0, nCand = 0;
    const int tw = p.tw;
    std::vector<double> sacc((size_t)p.nSH * tw);
    for (int w0 = 0; w0 < p.len; w0 += tw) {
        std::fill(sacc.begin(), sacc.end(), 0.0);
        double dlo, dhi; int jr, kr;
        scp_ims_window_range(&p, w0, tw, &dlo, &dhi);
        scp_ims_window_rows(&p, dhi, &jr, &kr);
        const int wj = 2 * jr + 1, nRows = wj * (2 * kr + 1);
        for (int r = 0; r < nRows; r++) {
            const int jj = r % wj - jr, kk = r / wj - kr;
            int lo[4], hi[4];
            const int nr = scp_ims_row_ranges(&p, jj, kk, dlo, dhi, lo, hi);
            for (int s = 0; s < nr; s++)
                for (int ii = lo[s]; ii <= hi[s]; ii += 2) {
                    nCand++;
                    float sx, sy, sz, d, time, att;
                    if (!scp_ims_image(&p, ii, jj, kk, &sx, &sy, &sz, &d)) continue;
                    const int tap = scp_ims_tap(&p, d, &time, &att);
                    if (tap < w0 || tap >= w0 + tw || tap >= p.len) continue;
                    nImg++;
                    const int t = tap - w0;
                    double tot = 0.0;
                    for (int b = 0; b < nBands; b++) tot += (double)(tx[b * maxW + ii + p.Nx] * ty[b * maxW + jj + p.Ny] * tz[b * maxW + kk + p.Nz]);
                    if (p.order == 0) { sacc[t] += (double)att * tot; continue; }
                    float azi, incl, Yv[(SCSH_MAX_ORDER + 1) * (SCSH_MAX_ORDER + 1)];
                    scp_ims_direction(sx, sy, sz, &azi, &incl);
                    if (p.order <= 3)      scsh_shreal_recur_dir<3>(p.order, azi, incl, &c[0][0], Yv);
                    else if (p.order <= 7) scsh_shreal_recur_dir<7>(p.order, azi, incl, &c[0][0], Yv);
                    else                   scsh_shreal_recur_dir<SCSH_MAX_ORDER>(p.order, azi, incl, &c[0][0], Yv);
                    for (int ch = 0; ch < p.nSH; ch++) sacc[(size_t)ch * tw + t] += (double)(Yv[ch] * att) * tot;
                }
        }
        for (int ch = 0; ch < p.nSH; ch++)
            for (int t = 0; t < tw && w0 + t < p.len; t++) rir[(size_t)ch * p.len + w0 + t] = (float)sacc[(size_t)ch * tw + t];
    }
    if (candidates) *candidates = nCand;
    return nImg;
}

int ph_sizeof_pair(void) { return (int)sizeof(ScpImsPair); }

}
