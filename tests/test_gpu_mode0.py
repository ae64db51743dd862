"""GPU tests of the TRUE non-partitioned convolver modes (csrc/safconv_np.c; reference
saf_utility_matrixConv.c:71-96, 174-207 and :277-298, 368-386): one FFT of numOvrlpAddBlocks * hopSize points per block on
the general-size device FFT, fftSize-long shifting overlap-add.  Checked against the oracle's / the compiled reference's
mode 0 (usePartFLAG = 0) and the golden fixtures; the plan must really use the reference's FFT size."""
import numpy as np
import pytest

from conftest import TOL_MAXABS_FS, TOL_REL_L2, err_metrics, golden_files

pytestmark = pytest.mark.gpu


@pytest.fixture()
def true_mode0(saf):
    saf.lib().safconv_set_true_mode0(1)
    yield
    saf.lib().safconv_set_true_mode0(0)


def check(y, ref, what=""):
    ma, l2 = err_metrics(y, ref)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, f"{what}: max-abs/fs {ma:.3g}, rel-L2 {l2:.3g}"


def ref_fft_size(hop, L):
    return int(np.ceil(np.float32(hop + L - 1) / np.float32(hop)) + 0.1) * hop        # .c:73-75


# hop, L, nIn, nOut, blocks
CASES = [
    (256, 1024, 4, 2, 12),        # C1 -> fftSize 1280 = 2^8 * 5
    (100, 333, 2, 3, 9),          # 500 = 2^2 * 5^3
    (17, 40, 3, 3, 11),           # 68 = 4 * 17: generic radix 17
    (128, 512, 25, 2, 10),        # C2 -> 640
    (2048, 512, 32, 40, 3),       # unit-test shape -> 4096
    (1024, 20000, 3, 2, 25),      # 21504 = 2^10 * 3 * 7: multi-launch FFT, generic radix 7; 21 blocks of overlap
    (64, 1, 2, 2, 5),             # one-tap filters: fftSize == hop
    (300, 7000, 2, 1, 30),        # 7500 = 2^2 * 3 * 5^4
]


@pytest.mark.parametrize("hop,L,nIn,nOut,nblk", CASES)
def test_true_mode0_matrixconv_vs_oracle(saf, orc, true_mode0, hop, L, nIn, nOut, nblk):
    rng = np.random.default_rng(hop + L + nIn)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 0).run(x)
    mc = saf.MatrixConv(hop, H, 0)
    info = mc.info()
    assert info.fftSize == ref_fft_size(hop, L) and info.numFilterBlocks == 1      # the reference's size, not a power of two
    y = mc.run(x)
    check(y, ref, "matrixConv, true mode 0")
    mc.reset_state()
    check(mc.run(x), ref, "matrixConv, true mode 0, after reset")
    mc.destroy()
    # usePartFLAG = 1 is untouched by the switch
    mp = saf.MatrixConv(hop, H, 1)
    assert mp.info().fftSize & (mp.info().fftSize - 1) == 0
    mp.destroy()


@pytest.mark.parametrize("hop,L,nCH,nblk", [(256, 1024, 32, 8), (50, 77, 3, 9), (512, 4096, 16, 12), (1000, 30000, 2, 40)])
def test_true_mode0_multiconv_vs_oracle(saf, orc, true_mode0, hop, L, nCH, nblk):
    rng = np.random.default_rng(hop + L + nCH)
    H = rng.uniform(-1, 1, (nCH, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nCH, hop * nblk)).astype(np.float32)
    ref = orc.OracleMultiConv(hop, H, 0).run(x)
    mc = saf.MultiConv(hop, H, 0)
    assert mc.info().fftSize == ref_fft_size(hop, L)
    check(mc.run(x), ref, "multiConv, true mode 0")
    mc.destroy()


def test_true_mode0_goldens_from_the_compiled_reference(saf, true_mode0):
    seen = 0
    for path in golden_files("matrix") + golden_files("multi"):
        g = np.load(path)
        if int(g["part"]) != 0:
            continue
        seen += 1
        cls = saf.MatrixConv if str(g["kind"]) == "matrix" else saf.MultiConv
        mc = cls(int(g["hop"]), g["H"], 0)
        assert mc.info().numFilterBlocks == 1
        check(mc.run(g["x"]), g["y"], path.stem)
        mc.destroy()
    assert seen >= 3


def test_true_mode0_odd_fft_size_falls_back_to_the_partitioned_engine(saf, orc, true_mode0):
    """hop 17, L 35 -> 3 blocks * 17 = 51: the reference's own saf_rfft_create asserts an even size; served partitioned."""
    hop, L = 17, 35
    rng = np.random.default_rng(1)
    H = rng.uniform(-1, 1, (2, 2, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (2, hop * 9)).astype(np.float32)
    mc = saf.MatrixConv(hop, H, 0)
    assert mc.info().fftSize == 64
    check(mc.run(x), orc.OracleMatrixConv(hop, H, 1).run(x), "odd fftSize")
    mc.destroy()


@pytest.mark.parametrize("part", [1, 0])
@pytest.mark.parametrize("hop,L,nIn,nOut,nblk", [(16384, 20000, 2, 3, 4), (12000, 5000, 3, 2, 5), (10000, 30000, 1, 2, 6)])
def test_block_sizes_above_8192(saf, orc, hop, L, nIn, nOut, nblk, part):
    """The reference takes any hopSize (saf_utility_matrixConv.c:100); the partitioned engine stops at 8192 (one block's FFT
    in one CTA's shared memory), larger blocks are served by the big-FFT engine for EITHER usePartFLAG -- same causal linear
    convolution as the reference's partitioned mode."""
    rng = np.random.default_rng(hop + L)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H, part)
    assert mc.info().fftSize == ref_fft_size(hop, L)
    check(mc.run(x), ref, f"matrixConv, hop {hop}")
    mc.destroy()
    Hm = rng.uniform(-1, 1, (nIn + 1, L)).astype(np.float32)
    xm = rng.uniform(-1, 1, (nIn + 1, hop * nblk)).astype(np.float32)
    mm = saf.MultiConv(hop, Hm, part)
    check(mm.run(xm), orc.OracleMultiConv(hop, Hm, 1).run(xm), f"multiConv, hop {hop}")
    mm.destroy()
