/* Host build of the general-size FFT building blocks (spatial_audio_framework_b200/csrc/safconv_gfft.cuh): the same
 * __host__ __device__ pass / split functions the kernels of safconv_gfft.cu call, driven by plain loops on the CPU.
 * Test infrastructure for tests/test_gfft_host.py (index algebra, butterflies, split passes, factor order) -- compiled
 * with nvcc, needs no GPU.  Not part of the product. */
#include <vector>
#include <cmath>
#include "../../spatial_audio_framework_b200/csrc/safconv_gfft.cuh"

template <bool INV>
static std::vector<float2> run_passes(std::vector<float2> a, int M, const int* fac, int nf, const std::vector<float2>& tw, float lastScale)
{
    std::vector<float2> b(M);
    int Ns = 1;
    for (int f = 0; f < nf; ++f) {
        const int R = fac[f];
        const float sc = (f == nf - 1) ? lastScale : 1.0f;
        const int nb = M / R;
        for (int j = 0; j < nb; ++j) {
            switch (R) {
                case 2: gfft_bfly<2, INV>(a.data(), b.data(), M, Ns, tw.data(), j, sc); break;
                case 3: gfft_bfly<3, INV>(a.data(), b.data(), M, Ns, tw.data(), j, sc); break;
                case 4: gfft_bfly<4, INV>(a.data(), b.data(), M, Ns, tw.data(), j, sc); break;
                case 5: gfft_bfly<5, INV>(a.data(), b.data(), M, Ns, tw.data(), j, sc); break;
                default: for (int q = 0; q < R; ++q) gfft_generic_elem<INV>(a.data(), b.data(), M, Ns, R, tw.data(), j + q * nb, sc); break;
            }
        }
        a.swap(b);
        Ns *= R;
    }
    return a;
}

extern "C" int gfft_host_rfft(int N, int dir, const int* fac, int nf, const float* in, float* out)
{
    const int M = N / 2;
    const double pi = 3.141592653589793238462643383279502884;
    std::vector<float2> tw(M), stw(M / 2 + 1);
    for (int i = 0; i < M; ++i) tw[i] = make_float2((float)cos(-2.0 * pi * i / M), (float)sin(-2.0 * pi * i / M));
    for (int k = 0; k <= M / 2; ++k) stw[k] = make_float2((float)cos(-2.0 * pi * k / N), (float)sin(-2.0 * pi * k / N));
    if (dir == 0) {
        std::vector<float2> z(M);
        for (int n = 0; n < M; ++n) z[n] = make_float2(in[2 * n], in[2 * n + 1]);
        std::vector<float2> Z = run_passes<false>(z, M, fac, nf, tw, 1.0f);
        float2* X = reinterpret_cast<float2*>(out);
        for (int k = 0; k <= M / 2; ++k) gfft_fwd_split(Z.data(), X, M, stw.data(), k);
    } else {
        std::vector<float2> zc(M);
        const float2* X = reinterpret_cast<const float2*>(in);
        for (int k = 0; k <= M / 2; ++k) gfft_inv_split(X, zc.data(), M, stw.data(), k);
        std::vector<float2> z = run_passes<true>(zc, M, fac, nf, tw, 1.0f / (float)N);
        for (int n = 0; n < M; ++n) { out[2 * n] = z[n].x; out[2 * n + 1] = z[n].y; }
    }
    return 0;
}
