"""CPU tests of the filter producers (no GPU needed):

1. the numpy fp64 restatement (oracle/producers.py) against the golden outputs of the compiled reference
   (tests/golden/producers_*.npz) and, when oracle/_ref/libsaf_ref_producers.so is here, against the reference itself on
   fresh inputs -- this pins the checker;
2. the device arithmetic -- csrc/safconv_sh.cuh and csrc/safconv_prod_core.cuh are __host__ __device__, the SAME functions
   the kernels of csrc/safconv_producers.cu call -- compiled into a host library (tests/host/producers_host.cu) and checked
   against the same goldens: SH values, image-source geometry / tap positions (bit-exact), whole RIRs, the 2 x 2 covariance
   matching;
3. host logic of the C layer that needs no device (argument errors, exported symbols are covered by tests/test_abi.py).
"""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, err_metrics

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "spatial_audio_framework_b200" / "csrc"
SRC = ROOT / "tests" / "host" / "producers_host.cu"
OUT = ROOT / "tests" / "host" / "_build" / "libproducers_host.so"

from oracle import producers as PR  # noqa: E402
from spatial_audio_framework_b200 import synth  # noqa: E402


def gold(name):
    return np.load(GOLDEN / f"producers_{name}.npz")


def ulps(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    return np.abs(a - b) / np.maximum(np.spacing(np.abs(b)), np.float32(1e-30))


# ------------------------------------------------------------------------------------------------------------------
#  1. the restatement is pinned on the reference
# ------------------------------------------------------------------------------------------------------------------
def test_np_sh_vs_reference_golden():
    g = gold("sh")
    Y = PR.np_rsh(10, g["dirs_deg"]).astype(np.float32)
    # fp64 evaluation rounded to fp32 and scaled in fp32 like getRSH: equal to the last bit almost everywhere
    assert np.abs(Y - g["rsh10"]).max() < 2e-6 * np.abs(g["rsh10"]).max()
    Yr = PR.np_shreal_recur_f32(7, g["dirs_rad"][:, 0], g["dirs_rad"][:, 1])
    assert np.abs(Yr - g["recur7"]).max() < 2e-6
    for o in range(11):
        assert np.allclose(PR.np_maxre(o), g["maxre"][o][:(o + 1) ** 2], rtol=0, atol=1e-7)


DEC_CASES = [(m, dc, mr) for m in (PR.DEFAULT, PR.LS, PR.LSDIFFEQ, PR.TA, PR.MAGLS) for dc in (0, 1) for mr in (0, 1)]


@pytest.mark.parametrize("m,dc,mr", DEC_CASES)
def test_np_decoder_vs_reference_golden(m, dc, mr):
    g = gold("decoder")
    ref = g[f"f_m{m}_dc{dc}_mr{mr}"]
    mine = PR.np_decoder_filters(g["hrtfs"], g["dirs_deg"], int(g["fftSize"]), float(g["fs"]), m, int(g["order"]),
                                 g["itd_s"], None, dc, mr)
    ma, l2 = err_metrics(ref, mine)
    # the reference solves in fp32 (cgesv per band); MagLS chains ~60 bands of fp32 phase continuation
    tol = 5e-5 if m == PR.MAGLS else 3e-6
    assert l2 < tol and ma < tol, (ma, l2)


def test_np_decoder_weights_and_orders_vs_golden():
    g = gold("decoder")
    H, d, itd, fftSize, fs = g["hrtfs"], g["dirs_deg"], g["itd_s"], int(g["fftSize"]), float(g["fs"])
    for key, kw in (("f_m1_weights", dict(method=PR.LS, order=3, weights=g["weights"], diffCM=1, maxRE=1)),
                    ("f_m5_o1", dict(method=PR.MAGLS, order=1)),
                    ("f_m2_o5", dict(method=PR.LSDIFFEQ, order=5, maxRE=1))):
        mine = PR.np_decoder_filters(H, d, fftSize, fs, kw["method"], kw["order"], itd, kw.get("weights"), kw.get("diffCM", 0), kw.get("maxRE", 0))
        ma, l2 = err_metrics(g[key], mine)
        assert l2 < (5e-5 if kw["method"] == PR.MAGLS else 3e-6), (key, ma, l2)


@pytest.mark.skipif(not PR.producers_reference_available(), reason="the t-design comes from oracle/_ref/libsaf_ref_producers.so")
def test_np_spr_decoder_vs_reference_golden():
    """SPR (saf_hoa_internal.c:332-430): interpolation order from the condition numbers, projection on the reference's t-design"""
    g = gold("decoder")
    R = PR.load_producers_reference()
    H, d, itd, fftSize, fs, order = g["hrtfs"], g["dirs_deg"], g["itd_s"], int(g["fftSize"]), float(g["fs"]), int(g["order"])
    nh, conds = PR.np_spr_order(d, d.shape[0])
    drad = np.stack([np.radians(d[:, 0]), np.pi / 2 - np.radians(d[:, 1])], 1).astype(np.float32)
    cref = R.cond_numbers(len(conds) - 1, drad)
    ok = cref < 1000
    assert np.allclose(conds[ok], cref[ok], rtol=2e-3) and nh >= order
    for dc in (0, 1):
        for mr in (0, 1):
            mine = PR.np_decoder_filters(H, d, fftSize, fs, PR.SPR, order, itd, None, dc, mr, tdesign_deg=R.tdesign(2 * order))
            ma, l2 = err_metrics(g[f"f_m3_dc{dc}_mr{mr}"], mine)
            assert l2 < 3e-6 and ma < 3e-6, (dc, mr, ma, l2)
    w = g["weights"] * np.float32(4 * np.pi)
    mine = PR.np_decoder_filters(H, d, fftSize, fs, PR.SPR, 1, itd, w, 0, 0, tdesign_deg=R.tdesign(2))
    ma, l2 = err_metrics(g["f_m3_o1_weights"], mine)
    assert l2 < 3e-6, (ma, l2)


IMS_CASES = {   # name: (order, maxN, maxTime_s, nBands, src, rec) -- the cases of tests/golden/make_golden_producers.py
    "t_o0": (0, -1, 0.10, 7, [5.1, 6.0, 1.1], [8.8, 5.5, 1.0]),
    "t_o3": (3, -1, 0.08, 7, [2.1, 1.0, 1.3], [8.8, 5.5, 0.9]),
    "t_o7": (7, -1, 0.04, 2, [4.4, 3.0, 1.4], [3.3, 2.5, 1.7]),
    "n_o2": (2, 4, -1.0, 7, [6.4, 4.0, 1.3], [1.0, 6.5, 2.0]),
    "n_o5_direct": (5, 0, -1.0, 7, [8.5, 5.0, 1.8], [8.8, 5.5, 0.9]),
}


@pytest.mark.parametrize("name", sorted(IMS_CASES))
def test_np_ims_vs_reference_golden(name):
    g = gold("ims")
    order, maxN, maxT, nB, src, rec = IMS_CASES[name]
    rir, idx = PR.np_ims_rir(synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL[:nB], nB, 343.0, 48e3, src, rec, order, maxN, maxT)
    ref = g[f"{name}_rir"]
    assert rir.shape == ref.shape
    # same image set, same taps (the reference's sorted propagation times -> (int)(t * fs + 0.5f))
    t = g[f"{name}_times"]
    ref_idx = (t * np.float32(48e3) + np.float32(0.5)).astype(np.int64)
    assert np.array_equal(np.sort(ref_idx), np.sort(idx))
    ma, l2 = err_metrics(rir, ref)
    assert l2 < 1e-6 and ma < 1e-6, (ma, l2)


@pytest.mark.skipif(not PR.producers_reference_available(), reason="oracle/_ref/libsaf_ref_producers.so not built")
def test_np_restatement_vs_live_reference_fresh_inputs():
    R = PR.load_producers_reference()
    H, d, itd = synth.synthetic_hrtfs(200, 64, 44100.0, seed=99)
    for m, order, dc, mr in ((PR.LS, 4, 1, 1), (PR.LSDIFFEQ, 2, 0, 0), (PR.TA, 3, 1, 0), (PR.MAGLS, 2, 0, 1)):
        a = R.decoder_filters(H, d, 64, 44100.0, m, order, itd, None, dc, mr)
        b = PR.np_decoder_filters(H, d, 64, 44100.0, m, order, itd, None, dc, mr)
        ma, l2 = err_metrics(a, b)
        assert l2 < (5e-5 if m == PR.MAGLS else 3e-6), (m, ma, l2)
    # a bigger echogram than the goldens hold: ~20 000 image sources, every tap position must agree
    s = R.ims([6.0, 5.0, 2.8], synth.IMS_TEST_ABS_WALL, 125.0, 7, 343.0, 48e3)
    sid = s.add_source([1.2, 3.3, 1.1]); rid = s.add_receiver_sh(2, [4.4, 1.7, 1.6])
    s.compute_echograms(-1, 0.22); s.render_rirs(0)
    ref = s.rir(rid, sid); t = s.echogram_times(rid, sid)
    s.destroy()
    rir, idx = PR.np_ims_rir([6.0, 5.0, 2.8], synth.IMS_TEST_ABS_WALL, 7, 343.0, 48e3, [1.2, 3.3, 1.1], [4.4, 1.7, 1.6], 2, -1, 0.22)
    assert t.size > 15000 and t.size == idx.size
    assert np.array_equal(np.sort((t * np.float32(48e3) + np.float32(0.5)).astype(np.int64)), np.sort(idx))
    ma, l2 = err_metrics(rir, ref)
    assert l2 < 1e-6, (ma, l2)


# ------------------------------------------------------------------------------------------------------------------
#  2. the device arithmetic, compiled for the host
# ------------------------------------------------------------------------------------------------------------------
class ScpImsPair(C.Structure):      # csrc/safconv_prod_core.cuh
    _fields_ = [("room", C.c_float * 3), ("so", C.c_float * 3), ("ro", C.c_float * 3),
                ("c_ms", C.c_float), ("fs", C.c_float), ("dmax", C.c_float), ("mode", C.c_int),
                ("Nx", C.c_int), ("Ny", C.c_int), ("Nz", C.c_int), ("lengthVec", C.c_longlong),
                ("order", C.c_int), ("nSH", C.c_int), ("len", C.c_int), ("accOff", C.c_longlong),
                ("tw", C.c_int), ("pad_", C.c_int)]


@pytest.fixture(scope="module")
def host_lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    OUT.parent.mkdir(exist_ok=True)
    newest = max(p.stat().st_mtime for p in (SRC, CSRC / "safconv_sh.cuh", CSRC / "safconv_prod_core.cuh"))
    if not OUT.exists() or OUT.stat().st_mtime < newest:
        subprocess.run([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
                        "-Wno-deprecated-gpu-targets", "-o", str(OUT), str(SRC)], check=True)
    L = C.CDLL(str(OUT))
    fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.ph_rsh.argtypes = [C.c_int, fp, C.c_int, fp]
    L.ph_shreal_recur.argtypes = [C.c_int, fp, C.c_int, fp]
    L.ph_legendre.argtypes = [C.c_int, C.c_double]; L.ph_legendre.restype = C.c_double
    L.ph_diffcov_M.argtypes = [dp, dp, dp]
    L.ph_ims_pair.argtypes = [C.POINTER(ScpImsPair), fp, C.c_int, C.c_int, fp, ip, ip]
    L.ph_ims_pair_windows.argtypes = [C.POINTER(ScpImsPair), fp, C.c_int, C.c_int, fp, C.POINTER(C.c_longlong)]
    L.ph_ims_pair_windows.restype = C.c_longlong
    assert L.ph_sizeof_pair() == C.sizeof(ScpImsPair)
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def test_device_rsh_vs_reference_golden(host_lib):
    g = gold("sh")
    d = np.ascontiguousarray(g["dirs_deg"])
    for order, key in ((10, "rsh10"), (3, "rsh3")):
        Y = np.zeros(((order + 1) ** 2, d.shape[0]), np.float32)
        host_lib.ph_rsh(order, _fp(d), d.shape[0], _fp(Y))
        u = ulps(Y, g[key])
        # a different (equally fp64-accurate) Legendre recurrence: the fp32 results differ by last-bit rounding at most --
        # except where the value itself is a rounding residue (|Y| < 1e-6 next to values of order 1, e.g. at the poles)
        big = np.abs(g[key]) > 1e-5
        assert (u[big] <= 2).all(), u[big].max()
        assert (u[big] == 0).mean() > 0.9
        assert np.abs(Y - g[key]).max() < 5e-7 * np.abs(g[key]).max()


def test_device_shreal_recur_vs_reference_golden(host_lib):
    g = gold("sh")
    d = np.ascontiguousarray(g["dirs_rad"])
    for order, key in ((2, "recur2"), (7, "recur7"), (10, "recur10")):
        Y = np.zeros(((order + 1) ** 2, d.shape[0]), np.float32)
        host_lib.ph_shreal_recur(order, _fp(d), d.shape[0], _fp(Y))
        # the same fp32 operations in the same order as saf_sh.c:255-330 (host build: no FMA contraction): bit-equal
        assert np.array_equal(Y, g[key]), (order, np.abs(Y - g[key]).max())


def test_device_maxre_legendre(host_lib):
    g = gold("sh")
    for o in range(11):
        x = float(np.cos(np.float32(137.9) * (np.float32(np.pi) / np.float32(180.0)) / (np.float32(o) + np.float32(1.51)), dtype=np.float32))
        for n in range(o + 1):
            assert abs(np.float32(host_lib.ph_legendre(n, x)) - g["maxre"][o][n * n]) <= 1.2e-7


def test_device_diffcov_2x2_vs_svd(host_lib):
    rng = np.random.default_rng(3)
    dp = C.POINTER(C.c_double)
    for _ in range(200):
        def herm():
            A = rng.standard_normal((2, 5)) + 1j * rng.standard_normal((2, 5))
            return A @ A.conj().T
        Cr, Ca = herm(), herm() * rng.uniform(0.01, 100)
        cref = np.array([Cr[0, 0].real, Cr[0, 1].real, Cr[0, 1].imag, Cr[1, 1].real])
        camb = np.array([Ca[0, 0].real, Ca[0, 1].real, Ca[0, 1].imag, Ca[1, 1].real])
        M = np.zeros(8)
        assert host_lib.ph_diffcov_M(cref.ctypes.data_as(dp), camb.ctypes.data_as(dp), M.ctypes.data_as(dp)) == 1
        M = (M[0::2] + 1j * M[1::2]).reshape(2, 2)
        # applyDiffCovMatching as written (saf_hoa.c:556-593): Cholesky factors, SVD, M = Xa^-1 V U^H X
        X = np.linalg.cholesky(Cr).conj().T; Xa = np.linalg.cholesky(Ca).conj().T
        U, _, Vh = np.linalg.svd(Xa.conj().T @ X)
        Mref = np.linalg.solve(Xa, Vh.conj().T @ (U.conj().T @ X))
        assert np.abs(M - Mref).max() < 1e-9 * max(1.0, np.abs(Mref).max())
        # and it does what it is for: M^H C_ambi M = C_ref
        assert np.abs(M.conj().T @ Ca @ M - Cr).max() < 1e-9 * np.abs(Cr).max()
    # not positive definite -> identity, flag 0 (utility_cchol zeroes its output there, veclib.c:4137-4147)
    bad = np.array([1.0, 2.0, 0.0, 1.0]); M = np.zeros(8)
    assert host_lib.ph_diffcov_M(bad.ctypes.data_as(dp), bad.ctypes.data_as(dp), M.ctypes.data_as(dp)) == 0
    assert np.array_equal(M, [1, 0, 0, 0, 0, 0, 1, 0])


def host_pair(room, c_ms, fs, src, rec, order, maxN, maxT):
    """the fp32 set-up of safconv_producers.c:safconv_ims_shoebox_renderRIRs, mirrored"""
    f = np.float32
    room = np.asarray(room, f); src = np.asarray(src, f); rec = np.asarray(rec, f)
    p = ScpImsPair()
    ry, sy = room[1] - rec[1], room[1] - src[1]
    so = [src[0] - room[0] / f(2), room[1] / f(2) - sy, src[2] - room[2] / f(2)]
    ro = [rec[0] - room[0] / f(2), room[1] / f(2) - ry, rec[2] - room[2] / f(2)]
    for i in range(3):
        p.room[i] = room[i]; p.so[i] = so[i]; p.ro[i] = ro[i]
    p.c_ms = c_ms; p.fs = fs
    if maxT > 0:
        dmax = f(maxT) * f(c_ms)
        p.mode = 0; p.dmax = dmax
        p.Nx, p.Ny, p.Nz = (int(dmax / room[i] + f(1.0)) for i in range(3))
    else:
        p.mode = 1; p.dmax = 0.0
        p.Nx = p.Ny = p.Nz = int(maxN)
    p.lengthVec = (2 * p.Nx + 1) * (2 * p.Ny + 1) * (2 * p.Nz + 1)
    p.order = order; p.nSH = (order + 1) ** 2
    return p


def wall_tables(abs_wall, N3):
    """wall_product of safconv_producers.c (the fp32 powf expressions of saf_reverb_internal.c:601-627)"""
    f = np.float32
    aw = np.asarray(abs_wall, f)
    nB = aw.shape[0]; maxW = 2 * max(N3) + 1
    tab = np.zeros((3, nB, maxW), f)
    for ax in range(3):
        N = N3[ax]
        for b in range(nB):
            r0, r1 = np.sqrt(f(1) - aw[b, 2 * ax]), np.sqrt(f(1) - aw[b, 2 * ax + 1])
            for o in range(-N, N + 1):
                a = f(abs(o))
                if o % 2 == 0:
                    v = np.power(r0, a / f(2), dtype=f) * np.power(r1, a / f(2), dtype=f)
                elif o > 0:
                    v = np.power(r0, np.ceil(f(o) / f(2)), dtype=f) * np.power(r1, np.floor(f(o) / f(2)), dtype=f)
                else:
                    v = np.power(r0, np.floor(a / f(2)), dtype=f) * np.power(r1, np.ceil(a / f(2)), dtype=f)
                tab[ax, b, o + N] = v
    return tab, maxW


@pytest.mark.parametrize("name", sorted(IMS_CASES))
def test_device_ims_pair_vs_reference_golden(host_lib, name):
    g = gold("ims")
    order, maxN, maxT, nB, src, rec = IMS_CASES[name]
    p = host_pair(synth.IMS_TEST_ROOM, 343.0, 48e3, src, rec, order, maxN, maxT)
    tab, maxW = wall_tables(synth.IMS_TEST_ABS_WALL[:nB], (p.Nx, p.Ny, p.Nz))
    n = C.c_int()
    length = host_lib.ph_ims_pair(C.byref(p), _fp(tab), nB, maxW, None, C.byref(n), None)
    ref = g[f"{name}_rir"]
    t = g[f"{name}_times"]
    assert n.value == t.size and length == ref.shape[1]            # same image set, same RIR length
    rir = np.zeros((p.nSH, length), np.float32); taps = np.zeros(n.value, np.int32)
    host_lib.ph_ims_pair(C.byref(p), _fp(tab), nB, maxW, _fp(rir), C.byref(n), taps.ctypes.data_as(C.POINTER(C.c_int)))
    ref_taps = (t * np.float32(48e3) + np.float32(0.5)).astype(np.int64)
    assert np.array_equal(np.sort(taps), np.sort(ref_taps))          # tap positions bit-exact
    ma, l2 = err_metrics(rir, ref)
    assert l2 < 1e-6 and ma < 1e-6, (ma, l2)
    assert np.array_equal(rir[0] != 0, ref[0] != 0)              # omni channel: all terms positive


def window_taps(nSH):
    """the host layer's choice of taps per window (safconv_producers.c)"""
    return max(32, min(1024, (6144 // nSH) & ~31))


WINDOW_CASES = dict(IMS_CASES)
WINDOW_CASES.update({
    "t_o2_long": (2, -1, 0.45, 7, [5.1, 6.0, 1.1], [8.8, 5.5, 0.9]),          # 73 000 images, 21 601 taps
    "t_o10": (10, -1, 0.05, 3, [1.0, 1.5, 1.2], [5.5, 2.5, 1.8]),             # 32-tap windows
    "t_close": (1, -1, 0.03, 2, [5.0, 3.5, 1.5], [5.0, 3.5, 1.5]),            # source ON the receiver: d = 0 image, symmetric rows
    "t_wall": (3, -1, 0.07, 7, [0.0, 0.0, 0.0], [10.0, 7.0, 3.0]),            # source and receiver in opposite corners (coincident images)
    "n_o1_big": (1, 9, -1.0, 4, [2.2, 6.1, 0.3], [7.7, 0.9, 2.6]),            # order-limited lattice, 1 500 images
})


@pytest.mark.parametrize("name", sorted(WINDOW_CASES))
def test_device_ims_windowed_render_equals_lattice_scan(host_lib, name):
    """the windowed render (one CTA per window of taps, candidates from the window's spherical shell) finds exactly the
    images of the full lattice scan, each once, and produces the same RIR"""
    order, maxN, maxT, nB, src, rec = WINDOW_CASES[name]
    p = host_pair(synth.IMS_TEST_ROOM, 343.0, 48e3, src, rec, order, maxN, maxT)
    tab, maxW = wall_tables(synth.IMS_TEST_ABS_WALL[:nB], (p.Nx, p.Ny, p.Nz))
    n = C.c_int()
    length = host_lib.ph_ims_pair(C.byref(p), _fp(tab), nB, maxW, None, C.byref(n), None)
    scan = np.zeros((p.nSH, length), np.float32)
    host_lib.ph_ims_pair(C.byref(p), _fp(tab), nB, maxW, _fp(scan), C.byref(n), None)
    p.tw = window_taps(p.nSH)
    win = np.full((p.nSH, length), np.nan, np.float32)
    cand = C.c_longlong()
    found = host_lib.ph_ims_pair_windows(C.byref(p), _fp(tab), nB, maxW, _fp(win), C.byref(cand))
    assert found == n.value, (found, n.value)                                  # every image, exactly once
    assert not np.isnan(win).any()                                             # every tap of every window written
    # the same fp64 sums in a different order: equal to fp32 rounding of the fp64 result (almost always bit-equal)
    assert np.allclose(win, scan, rtol=3e-7, atol=0) and (win == scan).mean() > 0.999
    assert p.lengthVec >= found
    print(name, "images", found, "candidates looked at", cand.value, "lattice", p.lengthVec)


def test_device_ims_broadband_vs_restatement(host_lib):
    """nBands = 1 (the reference's own renderRIRs crashes there: FIRFilterbank with zero cut-offs) vs the restatement"""
    aw = synth.IMS_TEST_ABS_WALL[3:4]
    p = host_pair([7.0, 4.0, 3.5], 343.0, 44100.0, [1.0, 1.5, 1.2], [5.5, 2.5, 1.8], 4, -1, 0.06)
    tab, maxW = wall_tables(aw, (p.Nx, p.Ny, p.Nz))
    n = C.c_int()
    length = host_lib.ph_ims_pair(C.byref(p), _fp(tab), 1, maxW, None, C.byref(n), None)
    rir = np.zeros((p.nSH, length), np.float32)
    host_lib.ph_ims_pair(C.byref(p), _fp(tab), 1, maxW, _fp(rir), C.byref(n), None)
    ref, idx = PR.np_ims_rir([7.0, 4.0, 3.5], aw, 1, 343.0, 44100.0, [1.0, 1.5, 1.2], [5.5, 2.5, 1.8], 4, -1, 0.06)
    assert ref.shape == rir.shape and idx.size == n.value
    ma, l2 = err_metrics(rir, ref)
    assert l2 < 1e-6, (ma, l2)


# ------------------------------------------------------------------------------------------------------------------
#  3. host layer without a device
# ------------------------------------------------------------------------------------------------------------------
def test_producer_argument_errors_need_no_device(saf):
    P = saf.producers
    H, d, itd = synth.synthetic_hrtfs(50, 16)
    with pytest.raises(saf.SafConvError, match="fftSize must be even"):
        P.decoder_filters(np.zeros((8, 2, 50), np.complex64), d, 15, 48000.0, P.DECODER_LS, 1)
    with pytest.raises(saf.SafConvError, match="needs the t-design of degree"):
        P.decoder_filters(H, d, 16, 48000.0, P.DECODER_SPR, 5)                 # nothing registered for degree 10
    with pytest.raises(saf.SafConvError, match="SPR needs order >= 1"):
        P.decoder_filters(H, d, 16, 48000.0, P.DECODER_SPR, 0)
    with pytest.raises(saf.SafConvError, match="safconv_register_tdesign"):
        P.register_tdesign(0, np.zeros((4, 2), np.float32))
    with pytest.raises(saf.SafConvError, match="order > 10"):
        P.decoder_filters(H, d, 16, 48000.0, P.DECODER_LS, 11)
    L = saf.lib()
    out = np.full((2, 4, 16), 7.0, np.float32)
    # the reference-named void function leaves the output untouched on error and records the reason per thread
    P._L().getBinauralAmbiDecoderFilters(H.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.POINTER(C.c_float)), 50, 16, 48000.0,
                                         P.DECODER_SPR, 5, None, None, 0, 0, out.ctypes.data_as(C.POINTER(C.c_float)))
    assert (out == 7.0).all() and L.safconv_last_error(None) == 1
    assert b"SPR" in L.safconv_last_error_string(None)


def test_producer_errors_without_a_device(saf):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    P = saf.producers
    H, d, itd = synth.synthetic_hrtfs(50, 16)
    with pytest.raises(saf.SafConvError, match="no usable CUDA device"):
        P.decoder_filters(H, d, 16, 48000.0, P.DECODER_LS, 1)
    with pytest.raises(saf.SafConvError, match="no usable CUDA device"):
        P.ImsShoebox(synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL, 125.0, 7, 343.0, 48e3)
    h = C.c_void_p()
    P._L().safconv_matrixConv_create_device(C.byref(h), 128, None, 100, 2, 2)
    assert not h and saf.lib().safconv_last_error(None) != 0


def test_spr_finds_saf_tdesign_tables_in_the_host_process(tmp_path):
    """inside a SAF host the SPR decoder needs no registration: the library's weak references to SAF's
    __HANDLES_Tdesign_dirs_deg / __Tdesign_nPoints_per_degree resolve against the process.  A stand-in "host" library
    defines the two symbols (a made-up degree-2 design); loaded first, it takes the call past the t-design check."""
    import subprocess
    import sys
    src = tmp_path / "fakesaf.c"
    src.write_text("static const float d2[4][2] = {{0,35.26f},{180,35.26f},{90,-35.26f},{-90,-35.26f}};\n"
                   "const float* __HANDLES_Tdesign_dirs_deg[21] = {0, &d2[0][0]};\n"
                   "const int __Tdesign_nPoints_per_degree[21] = {2, 4};\n")
    so = tmp_path / "libfakesaf.so"
    subprocess.run(["gcc", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    code = f"""
import ctypes as C, sys
sys.path.insert(0, {str(ROOT)!r})
if {{host}}: C.CDLL({str(so)!r}, mode=C.RTLD_GLOBAL)
import numpy as np, spatial_audio_framework_b200 as saf
from spatial_audio_framework_b200 import synth
H, d, itd = synth.synthetic_hrtfs(50, 16)
try:
    saf.producers.decoder_filters(H, d, 16, 48000.0, saf.producers.DECODER_SPR, 1)
    print("OK")
except saf.SafConvError as e:
    print("ERR", e)
"""
    with_host = subprocess.run([sys.executable, "-c", code.format(host=True)], capture_output=True, text=True).stdout
    without = subprocess.run([sys.executable, "-c", code.format(host=False)], capture_output=True, text=True).stdout
    assert "needs the t-design" in without
    assert "needs the t-design" not in with_host and ("OK" in with_host or "no usable CUDA device" in with_host), with_host


def test_device_rsh_up_to_order_20(host_lib):
    """the SPR decoder evaluates the measurement grid up to SH order 20 (441 channels): the device evaluator against the
    fp64 restatement there (no golden: the reference's own getSHreal needs factorials of 40 in double)"""
    d = synth.fibonacci_grid_deg(300)
    d[:2] = [[0, 90], [0, -90]]
    Y = np.zeros((441, 300), np.float32)
    host_lib.ph_rsh(20, _fp(np.ascontiguousarray(d)), 300, _fp(Y))
    ref = PR.np_rsh(20, d)
    assert np.isfinite(Y).all()
    assert np.abs(Y - ref).max() < 2e-6 * np.abs(ref).max()
    # orthonormality on a dense grid: (1 / N) Y Y^T = I up to the quadrature error of a Fibonacci grid at low orders
    G = (Y[:49].astype(np.float64) @ Y[:49].T.astype(np.float64)) / 300
    assert np.abs(G - np.eye(49)).max() < 0.1


def test_device_ims_windowed_render_random_scenes(host_lib):
    """randomised rooms / positions / window sizes / modes: the windowed enumeration must find every image of the lattice
    scan exactly once (scp_ims_window_range / _rows / scp_ims_row_ranges are conservative, the exact tap test decides)"""
    rng = np.random.default_rng(2024)
    for trial in range(40):
        room = rng.uniform(1.5, 12.0, 3).astype(np.float32)
        src = (rng.uniform(0.02, 0.98, 3) * room).astype(np.float32)
        rec = (rng.uniform(0.02, 0.98, 3) * room).astype(np.float32)
        if trial % 5 == 0:
            rec = src.copy()                                   # coincident source and receiver
        if trial % 7 == 0:
            src = np.array([0, 0, 0], np.float32)              # on a corner: images coincide pairwise
        order = int(rng.integers(0, 4))
        fs = float(rng.choice([8000.0, 44100.0, 48000.0, 96000.0]))
        c = float(rng.uniform(300.0, 360.0))
        if trial % 3 == 0:
            p = host_pair(room, c, fs, src, rec, order, int(rng.integers(0, 9)), -1.0)
        else:
            dist = float(np.linalg.norm(src - rec))
            p = host_pair(room, c, fs, src, rec, order, -1, (dist + float(rng.uniform(1.0, 25.0))) / c)
        nB = int(rng.integers(1, 4))
        tab, maxW = wall_tables(synth.IMS_TEST_ABS_WALL[:nB], (p.Nx, p.Ny, p.Nz))
        n = C.c_int()
        length = host_lib.ph_ims_pair(C.byref(p), _fp(tab), nB, maxW, None, C.byref(n), None)
        if n.value == 0:
            continue
        scan = np.zeros((p.nSH, length), np.float32)
        host_lib.ph_ims_pair(C.byref(p), _fp(tab), nB, maxW, _fp(scan), C.byref(n), None)
        p.tw = int(rng.choice([32, 64, 96, 384, 1024]))
        win = np.full((p.nSH, length), np.nan, np.float32)
        found = host_lib.ph_ims_pair_windows(C.byref(p), _fp(tab), nB, maxW, _fp(win), None)
        assert found == n.value, (trial, found, n.value)
        assert not np.isnan(win).any()
        assert np.allclose(win, scan, rtol=1e-6, atol=1e-12), trial
