"""Index algebra of the warp-level register FFT (csrc/safconv_wfft.cuh), emulated in numpy on the CPU:
which bin ends up in which (lane, register slot), where the real-FFT split finds its partner, and the
shuffle-free 32 x 32 variant.  Guards the mapping the CUDA kernels and their staging code rely on."""
import numpy as np
import pytest


def bitrev(x, bits):
    r = 0
    for b in range(bits):
        r |= ((x >> b) & 1) << (bits - 1 - b)
    return r


def dif_regs(v, R, sgn):
    """R-point DIF over the slot axis with the W_32^e constants of mul_w32 (dif_regs32)."""
    v = v.copy()
    h = R // 2
    while h >= 1:
        for i in range(R):
            if not (i & h):
                a, b = v[:, i].copy(), v[:, i + h].copy()
                v[:, i] = a + b
                v[:, i + h] = (a - b) * np.exp(sgn * 2j * np.pi * ((i & (h - 1)) * (16 // h)) / 32)
        h //= 2
    return v


def wfft(v, R, inv=False):
    """wfft<R>: lane j holds x[j + 32 i]; returns X[bitrev_R(i) + R * bitrev_5(l)] at (lane l, slot i)."""
    sgn = 1 if inv else -1
    logR, M = int(np.log2(R)), 32 * R
    v = dif_regs(v, R, sgn)
    for i in range(R):                                   # table T1[i][l] = W_M^(l * bitrev_R(i))
        v[:, i] *= np.exp(sgn * 2j * np.pi * (np.arange(32) * bitrev(i, logR)) / M)
    for half in (16, 8, 4, 2, 1):                        # WFFT_STAGE: partner l ^ half
        nv = v.copy()
        for l in range(32):
            o = v[l ^ half]
            if l & half:
                nv[l] = (o - v[l]) * np.exp(sgn * 2j * np.pi * ((l & (half - 1)) * (16 // half)) / 32)
            else:
                nv[l] = v[l] + o
        v = nv
    return v


@pytest.mark.parametrize("R", [2, 4, 8, 16, 32])
def test_wfft_layout_and_split_partner(R):
    rng = np.random.default_rng(R)
    M, logR = 32 * R, int(np.log2(R))
    x = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    v = np.array([[x[j + 32 * i] for i in range(R)] for j in range(32)])
    X, Xi = np.fft.fft(x), np.fft.ifft(x) * M
    out, outi = wfft(v, R), wfft(v, R, inv=True)
    for l in range(32):
        for i in range(R):
            k = bitrev(i, logR) + R * bitrev(l, 5)
            assert abs(out[l, i] - X[k]) < 1e-9 * M and abs(outi[l, i] - Xi[k]) < 1e-9 * M
    # wfft_fwd_split: the partner bin M - k sits in lane l ^ 31, slot bitrev_R(R - k2)  (k2 != 0)
    # or in lane bitrev_5(32 - bitrev_5(l)), slot 0  (k2 == 0)
    for l in range(32):
        k1 = bitrev(l, 5)
        pl0 = bitrev((32 - k1) & 31, 5)
        for i in range(R):
            k2 = bitrev(i, logR)
            k = k2 + R * k1
            if k == 0:
                continue
            pl, pi = (pl0, 0) if k2 == 0 else (l ^ 31, bitrev((R - k2) % R, logR))
            assert bitrev(pi, logR) + R * bitrev(pl, 5) == M - k


def test_wfft32t_transpose_variant():
    """wfft32t: 32-point DIF, twiddle, transpose through the 32 x 33 tile, second 32-point DIF:
    lane l, slot i ends with X[l + 32 * bitrev_5(i)]."""
    rng = np.random.default_rng(7)
    x = rng.standard_normal(1024) + 1j * rng.standard_normal(1024)
    v = np.array([[x[j + 32 * i] for i in range(32)] for j in range(32)])
    for inv, ref in ((False, np.fft.fft(x)), (True, np.fft.ifft(x) * 1024)):
        sgn = 1 if inv else -1
        a = dif_regs(v, 32, sgn)
        tile = np.zeros((32, 33), complex)
        for i in range(32):
            k2 = bitrev(i, 5)
            tile[k2, :32] = a[:, i] * np.exp(sgn * 2j * np.pi * (np.arange(32) * k2) / 1024)    # row k2, column j
        u = dif_regs(tile[:, :32].copy(), 32, sgn)                                              # lane = k2, slot = j
        for l in range(32):
            for i in range(32):
                assert abs(u[l, i] - ref[l + 32 * bitrev(i, 5)]) < 1e-8


def test_tile_index_skews_are_conflict_free():
    """Shared-memory index skews used around the warp FFT: row stride 33 (transpose tile), stride R + 1 for the
    time-sample / spectrum staging, stride 9 words for the forward operand staging -- all distinct banks per warp."""
    for stride_words in (2 * 33,):                       # float2 tile: stride 33 float2 = 66 words, 64-bit accesses
        banks = {(stride_words * l) % 32 for l in range(16)}                  # a half-warp per 64-bit wavefront
        assert len(banks) == 16 and all(b % 2 == 0 for b in banks)
    for R in (2, 4, 8, 16, 32):
        banks = {(2 * (R + 1) * l) % 32 for l in range(16)}
        assert len(banks) == 16
    assert len({(9 * l) % 32 for l in range(32)}) == 32
