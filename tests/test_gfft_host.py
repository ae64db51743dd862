"""CPU test of the general-size FFT building blocks: csrc/safconv_gfft.cuh is __host__ __device__, so the very functions
the kernels call (radix-2/3/4/5 butterflies, the generic-radix element, Stockham index algebra, real-FFT split passes) are
compiled into a host library with nvcc and checked against numpy and the oracle's KissFFT restatement -- no GPU needed.
The factor order comes from the product's own planner (safconv_debug_fft_factors = kf_factor, kiss_fft.c:310-331)."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "host" / "gfft_host.cu"
OUT = ROOT / "tests" / "host" / "_build" / "libgfft_host.so"

# the reference's own test sizes (test__utilities_module.c:381-384) up to 2^16, plus odd halves and big primes
SIZES = [2, 4, 6, 10, 14, 16, 22, 80, 160, 240, 256, 320, 480, 500, 640, 960, 1024, 1280, 1920, 2048, 3840, 4096, 7680, 8192,
         15360, 16384, 30720, 65536, 2 * 49, 2 * 77, 2 * 1009, 2 * 3 * 7 * 11, 1000, 2 * 625, 2 * 243]


@pytest.fixture(scope="module")
def host_lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    OUT.parent.mkdir(exist_ok=True)
    if not OUT.exists() or OUT.stat().st_mtime < max(SRC.stat().st_mtime, (ROOT / "spatial_audio_framework_b200" / "csrc" / "safconv_gfft.cuh").stat().st_mtime):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", str(OUT), str(SRC)], check=True)
    lib = C.CDLL(str(OUT))
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.gfft_host_rfft.argtypes = [C.c_int, C.c_int, ip, C.c_int, fp, fp]
    return lib


def factors(saf, M):
    buf = (C.c_int * 32)()
    n = saf.lib().safconv_debug_fft_factors(M, buf, 32)
    return [buf[i] for i in range(n)]


def test_factor_order_is_kissffts(saf):
    assert factors(saf, 1) == [1]
    assert factors(saf, 640) == [4, 4, 4, 2, 5]           # 4s first, then 2, then odd primes ascending
    assert factors(saf, 512) == [4, 4, 4, 4, 2]
    assert factors(saf, 15360) == [4, 4, 4, 4, 4, 3, 5]
    assert factors(saf, 1009) == [1009]
    assert factors(saf, 3 * 7 * 11) == [3, 7, 11]
    for M in (2, 3, 40, 77, 120, 625, 2 ** 19, 30030):
        f = factors(saf, M)
        assert int(np.prod(f)) == M and all(x >= 2 for x in f)


@pytest.mark.parametrize("N", SIZES)
def test_host_build_of_the_fft_core_vs_numpy_and_oracle(saf, orc, host_lib, N):
    rng = np.random.default_rng(N)
    x = rng.uniform(-1, 1, N).astype(np.float32)
    fac = factors(saf, N // 2)
    facbuf = (C.c_int * len(fac))(*fac)
    fp = C.POINTER(C.c_float)
    X = np.empty((N // 2 + 1, 2), np.float32)
    host_lib.gfft_host_rfft(N, 0, facbuf, len(fac), x.ctypes.data_as(fp), X.ctypes.data_as(fp))
    Xc = X[:, 0] + 1j * X[:, 1]
    ref = np.fft.rfft(x.astype(np.float64))
    tol = 2e-7 * (np.log2(N) + 4 + max(fac))             # fp32 rounding grows with the number of passes / the generic radix
    assert np.linalg.norm(Xc - ref) / np.linalg.norm(ref) < tol
    assert X[0, 1] == 0.0 and X[-1, 1] == 0.0            # DC and Nyquist are purely real (kiss_fftr.c:99-104)
    # the reference's own KissFFT (oracle restatement), same conventions
    Xo, _ = orc.oracle_rfft(N, x)
    Xo = Xo.reshape(-1, 2)
    Xo = Xo[:, 0] + 1j * Xo[:, 1]
    assert np.linalg.norm(Xc - Xo) / np.linalg.norm(Xo) < 2 * tol
    # backward: 1/N scaling, imaginary parts of DC / Nyquist ignored (kiss_fftr.c:137-138)
    Xin = X.copy()
    Xin[0, 1] = 123.0
    Xin[-1, 1] = -77.0
    y = np.empty(N, np.float32)
    host_lib.gfft_host_rfft(N, 1, facbuf, len(fac), Xin.ctypes.data_as(fp), y.ctypes.data_as(fp))
    assert np.abs(y - x).max() < 1e-5                     # the reference's own round-trip tolerance (test__saf_rfft)
