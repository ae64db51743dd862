"""The oracle restatement (oracle/safconv_oracle.c) must reproduce the compiled reference
(oracle/_ref, built from /root/reference by oracle/Makefile) BIT FOR BIT, and both must match the
committed golden fixtures.  This is the pin that makes the oracle trustworthy on the GPU box."""
import numpy as np
import pytest

from conftest import golden_files

import oracle as O

needs_ref = pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref not built (no /root/reference)")


@needs_ref
@pytest.mark.parametrize("N", [16, 64, 256, 2048, 4096, 80, 240, 1280, 3840, 2 * 7 * 11, 34])
def test_rfft_bit_exact(N):
    x = np.random.default_rng(N).uniform(-1, 1, N).astype(np.float32)
    X1, x1 = O.ref_rfft(N, x)
    X2, x2 = O.oracle_rfft(N, x)
    assert np.array_equal(X1, X2)
    assert np.array_equal(x1, x2)
    # reference test__saf_rfft (test__utilities_module.c:381-404): round trip within 1e-5
    assert np.abs(x2 - x).max() < 1e-5


@needs_ref
@pytest.mark.parametrize("hop,L,nIn,nOut,part,nblk", [
    (256, 1024, 4, 2, 1, 10), (256, 1024, 4, 2, 0, 10),
    (128, 512, 25, 2, 1, 8),
    (2048, 512, 3, 5, 1, 3),
    (96, 200, 2, 3, 1, 8), (100, 333, 2, 3, 0, 8),
    (64, 64, 1, 1, 1, 5), (64, 1, 2, 2, 1, 4), (64, 1, 2, 2, 0, 4),
    (32, 1000, 2, 2, 1, 40),
])
def test_matrixconv_bit_exact(hop, L, nIn, nOut, part, nblk):
    rng = np.random.default_rng(hop * 7 + L)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    a = O.RefMatrixConv(hop, H, part).run(x)
    b = O.OracleMatrixConv(hop, H, part).run(x)
    assert np.array_equal(a, b)


@needs_ref
@pytest.mark.parametrize("hop,L,nCH,part,nblk", [
    (512, 4096, 4, 1, 12), (256, 1024, 5, 0, 8), (128, 500, 3, 1, 10), (128, 500, 3, 0, 10), (50, 77, 2, 1, 9),
])
def test_multiconv_bit_exact(hop, L, nCH, part, nblk):
    rng = np.random.default_rng(hop * 3 + L)
    H = rng.uniform(-1, 1, (nCH, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nCH, hop * nblk)).astype(np.float32)
    a = O.RefMultiConv(hop, H, part).run(x)
    b = O.OracleMultiConv(hop, H, part).run(x)
    assert np.array_equal(a, b)


@needs_ref
def test_tvconv_bit_exact():
    rng = np.random.default_rng(5)
    hop, L, nIRs, nOut = 64, 300, 5, 2
    H = rng.uniform(-1, 1, (nIRs, nOut, L)).astype(np.float32)
    r, o = O.RefTVConv(hop, H, 7), O.OracleTVConv(hop, H, 7)   # initIdx out of range -> 0 (.c:459-465)
    for ir in [0, 0, 1, 4, 4, 4, 2, 3, 3, 0]:
        x = rng.uniform(-1, 1, hop).astype(np.float32)
        assert np.array_equal(r.apply(x, ir), o.apply(x, ir))


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.stem)
def test_oracle_matches_golden(path):
    """Golden fixtures were produced by the compiled reference (tests/golden/make_golden.py)."""
    g = np.load(path)
    kind = str(g["kind"])
    if kind == "matrix":
        y = O.OracleMatrixConv(int(g["hop"]), g["H"], int(g["part"])).run(g["x"])
    elif kind == "multi":
        y = O.OracleMultiConv(int(g["hop"]), g["H"], int(g["part"])).run(g["x"])
    elif kind == "tv":
        hop = int(g["hop"])
        tv = O.OracleTVConv(hop, g["H"], int(g["initIdx"]))
        y = np.concatenate([tv.apply(g["x"][0, i * hop:(i + 1) * hop], int(ir)) for i, ir in enumerate(g["seq"])], axis=1)
    elif kind == "fftconv":
        assert np.array_equal(O.oracle_fftconv(g["x"], g["h"]), g["y"])
        assert np.array_equal(O.oracle_fftconv(g["x"], g["h"], True), g["yfilt"])
        return
    else:
        X, xb = O.oracle_rfft(int(g["N"]), g["x"])
        assert np.array_equal(X, g["X"]) and np.array_equal(xb, g["xb"])
        return
    assert np.array_equal(y, g["y"])


@pytest.mark.parametrize("path", golden_files("matrix") + golden_files("multi"), ids=lambda p: p.stem)
def test_golden_close_to_fp64_truth(path):
    """Sanity of the fixtures themselves: the reference output is the causal linear convolution."""
    g = np.load(path)
    H, x, y = g["H"], g["x"], g["y"]
    if str(g["kind"]) == "matrix":
        t = O.truth_matrix(H, x, np.arange(H.shape[0]), 0, x.shape[1])
    else:
        t = O.truth_multi(H, x, np.arange(H.shape[0]), 0, x.shape[1])
    assert np.linalg.norm(y - t) / np.linalg.norm(t) < 5e-7


@pytest.mark.parametrize("nCH,xl,hl", [(3, 1000, 129), (1, 5, 7), (2, 4096, 4096), (4, 300, 1), (2, 1, 9)])
def test_fftconv_fftfilt_oracle_equals_reference(nCH, xl, hl):
    """fftconv / fftfilt (saf_utility_fft.c:157-228): restatement == compiled reference bit for bit, and both
    close to the fp64 direct convolution."""
    rng = np.random.default_rng(nCH + xl + hl)
    x = rng.uniform(-1, 1, (nCH, xl)).astype(np.float32)
    h = rng.uniform(-1, 1, (nCH, hl)).astype(np.float32)
    for filt in (False, True):
        a = O.oracle_fftconv(x, h, filt)
        b = O.ref_fftconv(x, h, filt)
        assert np.array_equal(a, b)
        t = np.stack([np.convolve(x[i].astype(np.float64), h[i].astype(np.float64)) for i in range(nCH)])[:, :a.shape[1]]
        assert np.abs(a - t).max() <= 2e-6 * max(np.abs(t).max(), 1e-30)
