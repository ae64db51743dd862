"""Generate tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref).

Run in the build container (where /root/reference exists and `make -C oracle ref` works):

    python tests/golden/make_golden.py

Each fixture stores the exact inputs (filters, signal, IR index sequence) and the output of the
reference's own saf_matrixConv / saf_multiConv / saf_TVConv
(/root/reference/framework/modules/saf_utilities/saf_utility_matrixConv.c) on them, so that the
pin "oracle == reference" and the GPU parity tests also hold on machines without the reference
tree.  The reference's test-suite itself has no golden vectors for this path.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent

MATRIX = [
    # name,            hop, L,    nIn, nOut, part, blocks
    ("matrix_c1_part",  256, 1024, 4,   2,    1,    8),    # BASELINE.json configs[0]
    ("matrix_c1_nopart", 256, 1024, 4,  2,    0,    8),
    ("matrix_c2_part",  128, 512,  25,  2,    1,    8),    # BASELINE.json configs[1]
    ("matrix_ragged_part", 96, 250, 3,  5,    1,    9),    # non-pow2 hop, L not a multiple of hop
    ("matrix_ragged_nopart", 100, 333, 2, 3,  0,    7),    # non-pow2 FFT size in the reference (500)
    ("matrix_ut_small", 2048, 512, 3,   4,    1,    3),    # unit-test hop/taps (P = 1), fewer channels
]
MULTI = [
    ("multi_part",   128, 1000, 4, 1, 12),
    ("multi_nopart", 256, 1024, 3, 0, 6),                  # test__examples.c:323 shape, fewer channels
    ("multi_c3_small", 512, 4096, 2, 1, 10),               # BASELINE.json configs[2], 2 of 256 channels
]


def main():
    lib, kind = O.load_reference()
    print("reference variant:", kind)
    rng = np.random.default_rng(20261018)
    for name, hop, L, nIn, nOut, part, nblk in MATRIX:
        H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
        x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
        y = O.RefMatrixConv(hop, H, part).run(x)
        np.savez(OUT / f"{name}.npz", kind="matrix", hop=hop, part=part, H=H, x=x, y=y)
        print(name, y.shape, float(np.abs(y).max()))
    for name, hop, L, nCH, part, nblk in MULTI:
        H = rng.uniform(-1, 1, (nCH, L)).astype(np.float32)
        x = rng.uniform(-1, 1, (nCH, hop * nblk)).astype(np.float32)
        y = O.RefMultiConv(hop, H, part).run(x)
        np.savez(OUT / f"{name}.npz", kind="multi", hop=hop, part=part, H=H, x=x, y=y)
        print(name, y.shape, float(np.abs(y).max()))
    # time-varying convolver with IR switches (exercises the 3-way cross-fade)
    hop, L, nIRs, nOut = 128, 700, 4, 3
    H = rng.uniform(-1, 1, (nIRs, nOut, L)).astype(np.float32)
    seq = np.array([1, 1, 2, 2, 2, 0, 3, 3, 1, 1, 1, 2], np.int32)
    x = rng.uniform(-1, 1, (1, hop * len(seq))).astype(np.float32)
    tv = O.RefTVConv(hop, H, 1)
    y = np.concatenate([tv.apply(x[0, i * hop:(i + 1) * hop], int(ir)) for i, ir in enumerate(seq)], axis=1)
    np.savez(OUT / "tvconv_switch.npz", kind="tv", hop=hop, initIdx=1, H=H, x=x, y=y, seq=seq)
    print("tvconv_switch", y.shape)
    # saf_rfft forward values for a few sizes (incl. non-pow2) -- pins the oracle's FFT restatement
    for N in (64, 256, 2048, 80, 1280):
        xx = rng.uniform(-1, 1, N).astype(np.float32)
        X, xb = O.ref_rfft(N, xx)
        np.savez(OUT / f"rfft_{N}.npz", kind="rfft", N=N, x=xx, X=X, xb=xb)


def fftconv_fixtures():
    """fftconv / fftfilt (saf_utility_fft.c:157-228) of the compiled reference; own seed so that adding them leaves
    the fixtures above untouched."""
    rng = np.random.default_rng(20261019)
    for name, nCH, xl, hl in (("fftconv_a", 3, 700, 129), ("fftconv_b", 2, 2048, 2048), ("fftconv_c", 1, 5, 7)):
        x = rng.uniform(-1, 1, (nCH, xl)).astype(np.float32)
        h = rng.uniform(-1, 1, (nCH, hl)).astype(np.float32)
        np.savez(OUT / f"{name}.npz", kind="fftconv", x=x, h=h, y=O.ref_fftconv(x, h), yfilt=O.ref_fftconv(x, h, True))
        print(name, nCH, xl, hl)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fftconv":
        fftconv_fixtures()
    else:
        main()
        fftconv_fixtures()
