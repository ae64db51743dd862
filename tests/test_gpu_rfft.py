"""GPU tests of the saf_rfft drop-in (saf_rfft_create / forward / backward / destroy, reference
saf_utility_fft.h:240-276, .c:531-753) on the general-size device FFT (csrc/safconv_gfft.cu, csrc/safconv_rfft.c):
every size the reference's own test walks (test__utilities_module.c:381-404, 16 ... 1048576 incl. the non-power-of-two
ones), the golden vectors produced by the compiled reference (tests/golden/rfft_*.npz incl. rfft_80 and rfft_1280), the
oracle's KissFFT restatement, odd halves and large prime factors, batches, page-locked buffers, error paths."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden_files

pytestmark = pytest.mark.gpu

# test__saf_rfft: fftSizesToTest[24]
REFERENCE_SIZES = [16, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 1048576,
                   80, 160, 320, 640, 1280, 240, 480, 960, 1920, 3840, 7680, 15360, 30720]


def oracle_forward(orc, x):
    Xo, _ = orc.oracle_rfft(len(x), x)
    Xo = Xo.reshape(-1, 2)
    return Xo[:, 0] + 1j * Xo[:, 1]


@pytest.mark.parametrize("N", REFERENCE_SIZES)
def test_saf_rfft_reference_sizes(saf, orc, N):
    """forward == the reference's KissFFT (rel L2 <= 1e-6), forward -> backward == identity within the reference's own
    tolerance (1e-5, test__saf_rfft), through the drop-in calls on a resident handle."""
    rng = np.random.default_rng(N)
    x = rng.uniform(-1, 1, N).astype(np.float32)
    f = saf.RFFT(N)
    assert int(np.prod(f.factors())) == N // 2
    X = f.forward(x)
    Xo = oracle_forward(orc, x)
    l2 = np.linalg.norm(X - Xo) / np.linalg.norm(Xo)
    assert l2 <= 1e-6, (N, l2)
    assert X[0].imag == 0 and X[-1].imag == 0
    xb = f.backward(X)
    assert np.abs(xb - x).max() <= 1e-5, N
    # a second call on the same handle (nothing is re-allocated), different data
    x2 = rng.uniform(-1, 1, N).astype(np.float32)
    assert np.abs(f.backward(f.forward(x2)) - x2).max() <= 1e-5
    f.destroy()


@pytest.mark.parametrize("path", golden_files("rfft"), ids=lambda p: p.stem)
def test_saf_rfft_against_reference_golden(saf, path):
    """Fixtures made by the compiled reference's saf_rfft_forward / backward (tests/golden/make_golden.py)."""
    g = np.load(path)
    N = int(g["N"])
    Xref = g["X"].astype(np.float32).reshape(-1, 2)
    Xref = Xref[:, 0] + 1j * Xref[:, 1]
    f = saf.RFFT(N)
    X = f.forward(g["x"])
    assert np.abs(X - Xref).max() / np.abs(Xref).max() <= 1e-6, path.stem
    xb = f.backward(Xref)
    assert np.abs(xb - g["xb"]).max() <= 1e-6 * max(np.abs(g["xb"]).max(), 1.0), path.stem
    f.destroy()


@pytest.mark.parametrize("N", [2, 4, 6, 10, 14, 22, 2 * 49, 2 * 77, 2 * 243, 2 * 625, 1000, 2 * 3 * 7 * 11, 2 * 1009, 2 * 8191,
                               2 * 9973, 18 * 4096, 2 * 3 * 5 * 7 * 11 * 13])
def test_saf_rfft_odd_halves_and_large_primes(saf, orc, N):
    """Anything even is legal in the reference (saf_utility_fft.c:542): odd N/2, prime factors beyond 5 (generic-radix
    pass like kf_bfly_generic), sizes on both sides of the one-CTA / multi-launch switch."""
    rng = np.random.default_rng(N + 1)
    x = rng.uniform(-1, 1, N).astype(np.float32)
    f = saf.RFFT(N)
    X = f.forward(x)
    ref = np.fft.rfft(x.astype(np.float64))
    big = max(f.factors())
    tol = 1e-6 if big <= 5 else 1e-6 * max(1.0, np.sqrt(big) / 4)       # O(R) sums of the generic radix round more
    assert np.linalg.norm(X - ref) / np.linalg.norm(ref) <= tol, (N, f.factors())
    Xo = oracle_forward(orc, x)
    assert np.linalg.norm(X - Xo) / np.linalg.norm(Xo) <= 2 * tol
    Xin = X.copy()
    Xin[0] += 5j                                                      # imaginary parts of DC / Nyquist are ignored
    Xin[-1] -= 3j
    assert np.abs(f.backward(Xin) - x).max() <= 1e-5 * max(1.0, np.sqrt(big) / 8)
    f.destroy()


@pytest.mark.parametrize("N", [64, 1280, 2 * 1009, 30720, 131072])
def test_saf_rfft_batches_and_pinned_buffers(saf, N):
    import torch
    rng = np.random.default_rng(N + 2)
    nb = 7
    x = rng.uniform(-1, 1, (nb, N)).astype(np.float32)
    f = saf.RFFT(N)
    X = f.forward(x)                                                  # batch grows the handle's buffers
    ref = np.fft.rfft(x.astype(np.float64), axis=1)
    assert np.linalg.norm(X - ref) / np.linalg.norm(ref) <= 2e-6
    assert np.abs(f.backward(X) - x).max() <= 2e-5
    one = f.forward(x[3])                                             # back to single transforms on the grown handle
    assert np.array_equal(one, X[3])
    # page-locked caller buffers, single transform through the drop-in call
    xin = torch.from_numpy(x[0].copy()).pin_memory()
    Xout = torch.empty((N // 2 + 1, 2), dtype=torch.float32).pin_memory()
    lib = saf.lib()
    lib.saf_rfft_forward(f._h, C.cast(xin.data_ptr(), C.POINTER(C.c_float)), C.c_void_p(Xout.data_ptr()))
    assert lib.safconv_rfft_last_error(f._h) == 0
    Xp = Xout.numpy()[:, 0] + 1j * Xout.numpy()[:, 1]
    assert np.array_equal(Xp, X[0])
    # an unaligned (4-byte) time-domain pointer
    buf = np.zeros(N + 1, np.float32)
    buf[1:] = x[1]
    Xu = f.forward(buf[1:])
    assert np.linalg.norm(Xu - ref[1]) / np.linalg.norm(ref[1]) <= 2e-6
    f.destroy()


def test_saf_rfft_stateless_helpers_any_even_size(saf):
    import spatial_audio_framework_b200 as pkg
    rng = np.random.default_rng(3)
    for N in (96, 500, 4096):
        x = rng.uniform(-1, 1, (3, N)).astype(np.float32)
        X = pkg.rfft_forward(x)
        ref = np.fft.rfft(x.astype(np.float64), axis=1)
        assert np.linalg.norm(X - ref) / np.linalg.norm(ref) <= 1e-6
        assert np.abs(pkg.rfft_backward(X) - x).max() <= 1e-5


def test_saf_rfft_errors(saf):
    lib = saf.lib()
    for N in (0, 1, 7, -4):
        h = C.c_void_p(5)
        lib.saf_rfft_create(C.byref(h), N)
        assert not h.value and b"even" in lib.safconv_last_error_string(None)
    h = C.c_void_p()
    lib.saf_rfft_destroy(C.byref(h))                                  # destroy(NULL) is a no-op
    x = np.zeros(8, np.float32)
    lib.saf_rfft_forward(None, x.ctypes.data_as(C.POINTER(C.c_float)), x.ctypes.data_as(C.c_void_p))   # NULL handle: no-op
    f = saf.RFFT(8)
    assert lib.safconv_rfft_batch(f._h, 2, 1, x.ctypes.data_as(C.POINTER(C.c_float)), x.ctypes.data_as(C.POINTER(C.c_float))) == 1
    f.destroy()
