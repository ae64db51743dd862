"""Drop-in integration test: the reference's OWN plug-in cores
examples/src/matrixconv/matrixconv.c, examples/src/multiconv/multiconv.c and examples/src/tvconv/tvconv.c, compiled UNMODIFIED by
oracle/Makefile, once linked with the reference's saf_utility_matrixConv.c (libsaf_ref_examples_ref.so) and
once with saf_matrixConv_* / saf_multiConv_* resolved from libsafconv_b200.so (libsaf_ref_examples_b200.so).
Both are driven like a plug-in host would (per-host-block float** audio, FIFO re-blocking to the clamped
frame size) and must produce the same audio within the north_star tolerance."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from conftest import TOL_MAXABS_FS, TOL_REL_L2, err_metrics

ROOT = Path(__file__).resolve().parents[1]
REF_SO = ROOT / "oracle" / "_ref" / "libsaf_ref_examples_ref.so"
B200_SO = ROOT / "oracle" / "_ref" / "libsaf_ref_examples_b200.so"
fpp = C.POINTER(C.POINTER(C.c_float))
fp = C.POINTER(C.c_float)


def _rows(a):
    """float** view of a 2-D float32 array (keeps the array alive through the returned tuple)."""
    a = np.ascontiguousarray(a, np.float32)
    ptrs = (fp * a.shape[0])(*[a[i].ctypes.data_as(fp) for i in range(a.shape[0])])
    return ptrs, a


def _drive(lib, name, H2d, nIn, part, host_block, x, nOutAsk):
    """create -> set filters -> init -> process host blocks; returns y[nOutAsk, T]."""
    g = lambda f: getattr(lib, f"{name}_{f}")
    h = C.c_void_p()
    g("create").argtypes = [C.POINTER(C.c_void_p)]
    g("create")(C.byref(h))
    if name == "matrixconv":
        lib.matrixconv_setNumInputChannels.argtypes = [C.c_void_p, C.c_int]
        lib.matrixconv_setNumInputChannels(h, nIn)
    g("setFilters").argtypes = [C.c_void_p, fpp, C.c_int, C.c_int, C.c_int]
    Hp, Hk = _rows(H2d)
    g("setFilters")(h, Hp, H2d.shape[0], H2d.shape[1], 48000)
    g("setEnablePart").argtypes = [C.c_void_p, C.c_int]
    g("setEnablePart")(h, part)
    g("init").argtypes = [C.c_void_p, C.c_int, C.c_int]
    g("init")(h, 48000, host_block)
    g("process").argtypes = [C.c_void_p, fpp, fpp, C.c_int, C.c_int, C.c_int]
    g("getProcessingDelay").argtypes = [C.c_void_p]
    delay = g("getProcessingDelay")(h)
    T = x.shape[1]
    y = np.zeros((nOutAsk, T), np.float32)
    for b in range(T // host_block):
        ip, ik = _rows(x[:, b * host_block:(b + 1) * host_block])
        ob = np.zeros((nOutAsk, host_block), np.float32)
        op, ok = _rows(ob)
        g("process")(h, ip, op, x.shape[0], nOutAsk, host_block)
        y[:, b * host_block:(b + 1) * host_block] = ok
    g("destroy").argtypes = [C.POINTER(C.c_void_p)]
    g("destroy")(C.byref(h))
    return y, delay


@pytest.mark.gpu
@pytest.mark.parametrize("host_block,part", [(128, 1), (1024, 1), (512, 0)])
def test_reference_matrixconv_core_runs_on_the_b200_library(saf, host_block, part):
    if not (REF_SO.exists() and B200_SO.exists()):
        pytest.skip("oracle/_ref example cores not built (need /root/reference at build time)")
    rng = np.random.default_rng(host_block)
    nIn, nOut, L, T = 5, 3, 1500, 8192
    H = rng.uniform(-1, 1, (nOut, nIn * L)).astype(np.float32)     # one row per output: nIn filters back to back
    x = rng.uniform(-1, 1, (nIn, T)).astype(np.float32)
    y_ref, d_ref = _drive(C.CDLL(str(REF_SO)), "matrixconv", H, nIn, part, host_block, x, nOut)
    y_gpu, d_gpu = _drive(C.CDLL(str(B200_SO)), "matrixconv", H, nIn, part, host_block, x, nOut)
    assert d_ref == d_gpu == max(512, host_block)                   # one frame of latency (matrixconv.c:310-314)
    assert np.abs(y_ref).max() > 1.0                                 # the convolver really ran
    ma, l2 = err_metrics(y_gpu, y_ref)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, (ma, l2)
    # and it is the expected convolution, delayed by one frame
    frame = max(512, host_block)
    exp = np.zeros((nOut, T))
    Hm = H.reshape(nOut, nIn, L).astype(np.float64)
    for no in range(nOut):
        for ni in range(nIn):
            exp[no] += np.convolve(x[ni].astype(np.float64), Hm[no, ni])[:T]
    nvalid = (T // frame - 1) * frame
    ma, l2 = err_metrics(y_gpu[:, frame:frame + nvalid], exp[:, :nvalid])
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, (ma, l2)


@pytest.mark.gpu
@pytest.mark.parametrize("host_block,part", [(256, 1), (2048, 0)])
def test_reference_multiconv_core_runs_on_the_b200_library(saf, host_block, part):
    if not (REF_SO.exists() and B200_SO.exists()):
        pytest.skip("oracle/_ref example cores not built (need /root/reference at build time)")
    rng = np.random.default_rng(host_block + 1)
    nCH, L, T = 6, 2000, 8192
    H = rng.uniform(-1, 1, (nCH, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nCH, T)).astype(np.float32)
    y_ref, _ = _drive(C.CDLL(str(REF_SO)), "multiconv", H, nCH, part, host_block, x, nCH)
    y_gpu, _ = _drive(C.CDLL(str(B200_SO)), "multiconv", H, nCH, part, host_block, x, nCH)
    assert np.abs(y_ref).max() > 1.0
    ma, l2 = err_metrics(y_gpu, y_ref)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, (ma, l2)


def _drive_tvconv(lib, irs, positions, host_block, x, walk):
    """tvconv_create -> (test shim: load IRs + listener positions) -> init -> process host blocks while the target
    position walks along `walk` (one entry per host block); returns y[nCh, T] and the position index per block."""
    nPos, nCh, L = irs.shape
    h = C.c_void_p()
    lib.tvconv_create.argtypes = [C.POINTER(C.c_void_p)]
    lib.tvconv_create(C.byref(h))
    lib.tvconv_testload.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp]
    irs = np.ascontiguousarray(irs, np.float32)
    positions = np.ascontiguousarray(positions, np.float32)
    lib.tvconv_testload(h, irs.ctypes.data_as(fp), nPos, nCh, L, 48000, positions.ctypes.data_as(fp))
    lib.tvconv_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.tvconv_init(h, 48000, host_block)
    lib.tvconv_setTargetPosition.argtypes = [C.c_void_p, C.c_float, C.c_int]
    lib.tvconv_getListenerPositionIdx.argtypes = [C.c_void_p]
    lib.tvconv_process.argtypes = [C.c_void_p, fpp, fpp, C.c_int, C.c_int, C.c_int]
    T = x.shape[1]
    y = np.zeros((nCh, T), np.float32)
    idx = []
    for b in range(T // host_block):
        lib.tvconv_setTargetPosition(h, float(walk[b]), 0)
        idx.append(lib.tvconv_getListenerPositionIdx(h))
        ip, ik = _rows(x[:, b * host_block:(b + 1) * host_block])
        ob = np.zeros((nCh, host_block), np.float32)
        op, ok = _rows(ob)
        lib.tvconv_process(h, ip, op, 1, nCh, host_block)
        y[:, b * host_block:(b + 1) * host_block] = ok
    lib.tvconv_destroy.argtypes = [C.POINTER(C.c_void_p)]
    lib.tvconv_destroy(C.byref(h))
    return y, idx


@pytest.mark.gpu
@pytest.mark.parametrize("host_block,L,nblk", [(512, 3000, 24), (2048, 9000, 10), (8192, 20000, 6), (16384, 5000, 3)])
def test_reference_tvconv_core_runs_on_the_b200_library(saf, host_block, L, nblk):
    """examples/src/tvconv/tvconv.c UNMODIFIED on saf_TVConv_* from libsafconv_b200.so vs the same core on the
    reference's own convolver: frames clamp to [512, 8192] (tvconv_internal.h:43-44) -- 8192-sample frames included --,
    the listener position (= impulse-response set) changes while audio runs (3-way cross-fade, .c:577-611)."""
    if not (REF_SO.exists() and B200_SO.exists()):
        pytest.skip("oracle/_ref example cores not built (need /root/reference at build time)")
    rng = np.random.default_rng(host_block)
    nPos, nCh = 6, 4
    irs = (rng.uniform(-1, 1, (nPos, nCh, L)) * np.exp(-4.0 * np.arange(L) / L)).astype(np.float32)
    positions = np.zeros((nPos, 3), np.float32)
    positions[:, 0] = np.arange(nPos)                      # listeners on a line along x
    T = host_block * nblk
    x = rng.uniform(-1, 1, (1, T)).astype(np.float32)
    walk = rng.uniform(-0.4, nPos - 0.6, nblk)
    walk[1] = walk[0]                                      # a repeated position: no cross-fade for that block
    y_ref, i_ref = _drive_tvconv(C.CDLL(str(REF_SO)), irs, positions, host_block, x, walk)
    y_gpu, i_gpu = _drive_tvconv(C.CDLL(str(B200_SO)), irs, positions, host_block, x, walk)
    assert i_ref == i_gpu and len(set(i_ref)) > 1          # the position really moved
    assert np.abs(y_ref).max() > 0.5                        # the convolver really ran
    ma, l2 = err_metrics(y_gpu, y_ref)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, (ma, l2)


def test_example_core_libraries_export_the_reference_api():
    if not (REF_SO.exists() and B200_SO.exists()):
        pytest.skip("oracle/_ref example cores not built")
    import subprocess
    for so in (REF_SO, B200_SO):
        syms = subprocess.run(["nm", "-D", str(so)], capture_output=True, text=True).stdout
        for s in ("matrixconv_create", "matrixconv_process", "matrixconv_setFilters", "multiconv_create", "multiconv_process",
                  "tvconv_create", "tvconv_process", "tvconv_setTargetPosition"):
            assert s in syms
    # the b200 flavour takes the convolver from the product library, the ref flavour carries its own
    u = subprocess.run(["nm", "-D", "--undefined-only", str(B200_SO)], capture_output=True, text=True).stdout
    assert "saf_matrixConv_apply" in u and "saf_multiConv_apply" in u and "saf_TVConv_apply" in u
    r = subprocess.run(["nm", "-D", "--undefined-only", str(REF_SO)], capture_output=True, text=True).stdout
    assert "saf_matrixConv_apply" not in r


REVERB_SO = ROOT / "oracle" / "_ref" / "libsaf_ref_reverbtest_b200.so"


@pytest.mark.gpu
@pytest.mark.skipif(not REVERB_SO.exists(), reason="oracle/_ref/libsaf_ref_reverbtest_b200.so not built (needs /root/reference)")
def test_reference_ims_unit_test_runs_on_the_library(saf):
    """The reference's own unit test of the image-source simulator, test__ims_shoebox_RIR
    (test/src/test__reverb_module.c:27-96: add / remove / re-add sources, move a source and the receiver ten times,
    computeEchograms + renderRIRs after every move), compiled unmodified with ims_shoebox_* resolved from
    libsafconv_b200.so.  It asserts nothing itself; a hook in front of its final ims_shoebox_destroy
    (oracle/shim/reverbtest_hooks.c) copies the rendered RIRs out, and they must equal the compiled reference's for the same
    sequence (tests/golden/producers_ims.npz, ut_*)."""
    from conftest import GOLDEN
    g = np.load(GOLDEN / "producers_ims.npz")
    L = C.CDLL(str(REVERB_SO))
    L.test__ims_shoebox_RIR()
    n = L.reverbtest_num_captured()          # (the hook probes every receiver / source ID, so the thread's last error is its own)
    assert n == 3
    L.reverbtest_get.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(fp)]
    seen = set()
    for i in range(n):
        rid, sid, ln, nch, data = C.c_int(), C.c_int(), C.c_int(), C.c_int(), fp()
        assert L.reverbtest_get(i, C.byref(rid), C.byref(sid), C.byref(ln), C.byref(nch), C.byref(data)) == 0
        rir = np.ctypeslib.as_array(data, shape=(nch.value, ln.value)).copy()
        ref = g[f"ut_rir_r{rid.value}_s{sid.value}"]
        assert rir.shape == ref.shape
        ma, l2 = err_metrics(rir, ref)
        assert l2 <= TOL_REL_L2 and ma <= TOL_MAXABS_FS, (rid.value, sid.value, ma, l2)
        seen.add(sid.value)
    assert len(seen) == 3
