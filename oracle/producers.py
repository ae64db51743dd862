"""TEST INFRASTRUCTURE ONLY -- CPU checkers for the filter PRODUCERS.

Two things live here:

* ``RefProducers``: ctypes access to the UNMODIFIED reference functions compiled in place into
  ``oracle/_ref/libsaf_ref_producers.so`` (``oracle/Makefile``): ``getBinauralAmbiDecoderFilters``
  (``/root/reference/framework/modules/saf_hoa/saf_hoa.c:452-497``) and the image-source simulator
  ``ims_shoebox_*`` (``/root/reference/framework/modules/saf_reverb/saf_reverb.c:36-856``).
* a numpy **fp64** restatement of the same algorithms (``np_*`` below), each function citing the reference lines it
  follows.  It is the "truth" the tests measure both the reference (fp32 LAPACK) and the CUDA path against, and the
  checker that travels to the GPU box when ``oracle/_ref`` is not there.

Pinned: ``tests/test_producers_oracle.py`` checks the restatement against the compiled reference on every decoder
method / flag and on the IMS scenes of the reference's own unit test (``test/src/test__reverb_module.c:27-96``);
``tests/golden/producers_*.npz`` hold outputs of the compiled reference (``tests/golden/make_golden_producers.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import math
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_REF_SO = HERE / "_ref" / "libsaf_ref_producers.so"

_f32p = C.POINTER(C.c_float)

# decoder methods, saf_hoa.h:131-171
DEFAULT, LS, LSDIFFEQ, SPR, TA, MAGLS = range(6)


def producers_reference_available() -> bool:
    return _REF_SO.exists()


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


class RefProducers:
    """The compiled reference.  Fortran LAPACK prints to stdout when the uplo shim is absent; it is linked in."""

    def __init__(self):
        if not _REF_SO.exists():
            raise FileNotFoundError(str(_REF_SO))
        L = C.CDLL(str(_REF_SO))
        L.getBinauralAmbiDecoderFilters.restype = None
        L.getBinauralAmbiDecoderFilters.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                                    _f32p, _f32p, C.c_int, C.c_int, _f32p]
        L.getRSH.argtypes = [C.c_int, _f32p, C.c_int, _f32p]
        L.getSHreal_recur.argtypes = [C.c_int, _f32p, C.c_int, _f32p]
        L.getMaxREweights.argtypes = [C.c_int, C.c_int, _f32p]
        L.ims_shoebox_create.argtypes = [C.POINTER(C.c_void_p), _f32p, _f32p, C.c_float, C.c_int, C.c_float, C.c_float]
        L.ims_shoebox_destroy.argtypes = [C.POINTER(C.c_void_p)]
        L.ims_shoebox_computeEchograms.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.ims_shoebox_renderRIRs.argtypes = [C.c_void_p, C.c_int]
        L.ims_shoebox_addSource.argtypes = [C.c_void_p, _f32p, C.c_void_p]
        L.ims_shoebox_addReceiverSH.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_void_p]
        L.ims_shoebox_updateSource.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.ims_shoebox_updateReceiver.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.ims_shoebox_removeSource.argtypes = [C.c_void_p, C.c_int]
        L.ims_shoebox_removeReceiver.argtypes = [C.c_void_p, C.c_int]
        L.ims_shoebox_setRoomDimensions.argtypes = [C.c_void_p, _f32p]
        L.ims_shoebox_setWallAbsCoeffs.argtypes = [C.c_void_p, _f32p]
        L.oracle_ims_get_rir.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(_f32p), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_ims_get_echogram.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p, C.c_int]
        L.oracle_tdesign.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.oracle_tdesign.restype = _f32p
        L.checkCondNumberSHTReal.argtypes = [C.c_int, _f32p, C.c_int, _f32p, _f32p]
        self.L = L

    def tdesign(self, degree):
        """the reference's minimum t-design of `degree` (1..21): [nPoints, 2] = (azimuth, elevation) in degrees"""
        n = C.c_int()
        p = self.L.oracle_tdesign(int(degree), C.byref(n))
        if not p:
            raise ValueError("degree must be 1..21")
        return np.ctypeslib.as_array(p, shape=(n.value, 2)).copy()

    def cond_numbers(self, order, dirs_rad, weights=None):
        """checkCondNumberSHTReal (saf_sh.c:884-953): condition number of the SH transform per order 0..order"""
        d = np.ascontiguousarray(dirs_rad, np.float32)
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        out = np.zeros(order + 1, np.float32)
        self.L.checkCondNumberSHTReal(int(order), _fp(d), d.shape[0], None if w is None else _fp(w), _fp(out))
        return out

    # --- saf_hoa -------------------------------------------------------------------------------------------------
    def decoder_filters(self, hrtfs, dirs_deg, fftSize, fs, method, order, itd_s=None, weights=None,
                        diffCM=0, maxRE=0):
        hrtfs = np.ascontiguousarray(hrtfs, np.complex64)
        dirs = np.ascontiguousarray(dirs_deg, np.float32)
        nB, nE, nD = hrtfs.shape
        assert nE == 2 and nB == fftSize // 2 + 1 and dirs.shape == (nD, 2)
        nSH = (order + 1) ** 2
        out = np.zeros((2, nSH, fftSize), np.float32)
        itd = None if itd_s is None else np.ascontiguousarray(itd_s, np.float32)
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        self.L.getBinauralAmbiDecoderFilters(hrtfs.ctypes.data_as(C.c_void_p), _fp(dirs), nD, fftSize, float(fs),
                                             int(method), int(order), None if itd is None else _fp(itd),
                                             None if w is None else _fp(w), int(diffCM), int(maxRE), _fp(out))
        return out

    def rsh(self, order, dirs_deg):
        dirs = np.ascontiguousarray(dirs_deg, np.float32)
        Y = np.zeros(((order + 1) ** 2, dirs.shape[0]), np.float32)
        self.L.getRSH(order, _fp(dirs), dirs.shape[0], _fp(Y))
        return Y

    def shreal_recur(self, order, dirs_rad):
        dirs = np.ascontiguousarray(dirs_rad, np.float32)
        Y = np.zeros(((order + 1) ** 2, dirs.shape[0]), np.float32)
        self.L.getSHreal_recur(order, _fp(dirs), dirs.shape[0], _fp(Y))
        return Y

    def maxre(self, order):
        a = np.zeros((order + 1) ** 2, np.float32)
        self.L.getMaxREweights(order, 0, _fp(a))
        return a

    # --- saf_reverb ----------------------------------------------------------------------------------------------
    def ims(self, room, abs_wall, lowest_band, n_bands, c_ms, fs):
        return _RefIms(self.L, room, abs_wall, lowest_band, n_bands, c_ms, fs)


class _RefIms:
    def __init__(self, L, room, abs_wall, lowest_band, n_bands, c_ms, fs):
        self.L = L
        self.h = C.c_void_p()
        room = np.ascontiguousarray(room, np.float32)
        aw = np.ascontiguousarray(abs_wall, np.float32)
        L.ims_shoebox_create(C.byref(self.h), _fp(room), _fp(aw), float(lowest_band), int(n_bands), float(c_ms), float(fs))

    def add_source(self, xyz):
        return self.L.ims_shoebox_addSource(self.h, _fp(np.ascontiguousarray(xyz, np.float32)), None)

    def add_receiver_sh(self, order, xyz):
        return self.L.ims_shoebox_addReceiverSH(self.h, int(order), _fp(np.ascontiguousarray(xyz, np.float32)), None)

    def update_source(self, sid, xyz):
        self.L.ims_shoebox_updateSource(self.h, sid, _fp(np.ascontiguousarray(xyz, np.float32)))

    def update_receiver(self, rid, xyz):
        self.L.ims_shoebox_updateReceiver(self.h, rid, _fp(np.ascontiguousarray(xyz, np.float32)))

    def remove_source(self, sid):
        self.L.ims_shoebox_removeSource(self.h, sid)

    def remove_receiver(self, rid):
        self.L.ims_shoebox_removeReceiver(self.h, rid)

    def set_room(self, room):
        self.L.ims_shoebox_setRoomDimensions(self.h, _fp(np.ascontiguousarray(room, np.float32)))

    def set_abs(self, abs_wall):
        self.L.ims_shoebox_setWallAbsCoeffs(self.h, _fp(np.ascontiguousarray(abs_wall, np.float32)))

    def compute_echograms(self, maxN, maxTime_s):
        self.L.ims_shoebox_computeEchograms(self.h, int(maxN), float(maxTime_s))

    def render_rirs(self, frac=0):
        self.L.ims_shoebox_renderRIRs(self.h, int(frac))

    def rir(self, rid, sid):
        p = _f32p(); n = C.c_int(); ch = C.c_int()
        if self.L.oracle_ims_get_rir(self.h, rid, sid, C.byref(p), C.byref(n), C.byref(ch)) != 0 or not p:
            return None
        return np.ctypeslib.as_array(p, shape=(ch.value, n.value)).copy()

    def echogram_times(self, rid, sid):
        n = self.L.oracle_ims_get_echogram(self.h, rid, sid, None, 0)
        t = np.zeros(max(n, 1), np.float32)
        self.L.oracle_ims_get_echogram(self.h, rid, sid, _fp(t), n)
        return t[:n]

    def destroy(self):
        if self.h:
            self.L.ims_shoebox_destroy(C.byref(self.h))
            self.h = C.c_void_p()


_ref = None


def load_producers_reference() -> RefProducers:
    global _ref
    if _ref is None:
        _ref = RefProducers()
    return _ref


# =====================================================================================================================
#  fp64 restatement: spherical harmonics
# =====================================================================================================================

def np_legendre_all(order: int, x: np.ndarray) -> np.ndarray:
    """Unnormalised associated Legendre functions WITHOUT the Condon-Shortley phase, P[n][m][dir], fp64.
    (saf_sh.c:53-126 computes the same values with the phase and getSHreal :217-226 cancels it again.)"""
    x = np.asarray(x, np.float64)
    s = np.sqrt(np.maximum(0.0, 1.0 - x * x))
    P = np.zeros((order + 1, order + 1, x.size))
    P[0, 0] = 1.0
    for m in range(1, order + 1):
        P[m, m] = P[m - 1, m - 1] * (2 * m - 1) * s
    for m in range(0, order):
        P[m + 1, m] = (2 * m + 1) * x * P[m, m]
    for m in range(0, order + 1):
        for n in range(m + 2, order + 1):
            P[n, m] = ((2 * n - 1) * x * P[n - 1, m] - (n + m - 1) * P[n - 2, m]) / (n - m)
    return P


def np_shreal(order: int, azi_rad: np.ndarray, incl_rad: np.ndarray) -> np.ndarray:
    """Real orthonormal SH (ACN order), (order+1)^2 x nDirs, fp64 -- getSHreal, saf_sh.c:190-253."""
    azi = np.asarray(azi_rad, np.float64)
    P = np_legendre_all(order, np.cos(np.asarray(incl_rad, np.float64)))
    Y = np.zeros(((order + 1) ** 2, azi.size))
    for n in range(order + 1):
        for m in range(-n, n + 1):
            am = abs(m)
            norm = math.sqrt((2 * n + 1) * math.factorial(n - am) / (4 * math.pi * math.factorial(n + am)))
            if m < 0:
                Y[n * n + n + m] = norm * P[n, am] * math.sqrt(2.0) * np.sin(am * azi)
            elif m == 0:
                Y[n * n + n] = norm * P[n, 0]
            else:
                Y[n * n + n + m] = norm * P[n, am] * math.sqrt(2.0) * np.cos(am * azi)
    return Y


def np_rsh(order: int, dirs_deg: np.ndarray) -> np.ndarray:
    """getRSH, saf_hoa.c:118-150: N3D real SH without the 1/sqrt(4pi) term; directions [azi, elev] in degrees.
    The degree->radian conversion is done in fp32 like the reference (:138-141), the rest in fp64."""
    d = np.asarray(dirs_deg, np.float32)
    pi = np.float32(math.pi)
    azi = d[:, 0] * pi / np.float32(180.0)
    incl = pi / np.float32(2.0) - (d[:, 1] * pi / np.float32(180.0))
    return np_shreal(order, azi.astype(np.float64), incl.astype(np.float64)) * math.sqrt(4.0 * math.pi)


def np_maxre(order: int) -> np.ndarray:
    """getMaxREweights (diagMtxFlag = 0), saf_hoa.c:235-266: P_n(cos(137.9 deg / (order + 1.51))) per SH channel."""
    x = float(np.cos(np.float32(137.9) * (np.float32(math.pi) / np.float32(180.0)) / (np.float32(order) + np.float32(1.51)),
                     dtype=np.float32))
    a = np.zeros((order + 1) ** 2)
    p0, p1 = 1.0, x
    for n in range(order + 1):
        pn = p0 if n == 0 else p1
        if n >= 2:
            pn = ((2 * n - 1) * x * p1 - (n - 1) * p0) / n
            p0, p1 = p1, pn
        a[n * n:(n + 1) * (n + 1)] = pn
    return a


# =====================================================================================================================
#  fp64 restatement: binaural Ambisonic decoder filters
# =====================================================================================================================

def _weights(nD, weights):
    return np.full(nD, 1.0 / nD) if weights is None else np.asarray(weights, np.float64)


def _band_cutoff(freqs):
    """first index of the band closest to 1.5 kHz (saf_hoa_internal.c:465-473), fp32 comparison like the reference"""
    f = np.asarray(freqs, np.float32)
    return int(np.argmin(np.abs(f - np.float32(1.5e3))))


def np_spr_order(dirs_deg, n_dirs, weights=None):
    """the SH order the SPR decoder interpolates the HRTF set with (saf_hoa_internal.c:357-371): the LAST order up to
    min(int(sqrt(N_dirs) - 1), 20) whose SH transform has a condition number below 100 (checkCondNumberSHTReal)"""
    nh_max = min(int(np.float32(np.sqrt(np.float32(n_dirs))) - np.float32(1.0)), 20)
    Y = np_rsh(nh_max, dirs_deg) / math.sqrt(4.0 * math.pi)
    w = np.ones(n_dirs) if weights is None else np.asarray(weights, np.float64)
    conds = []
    for n in range(nh_max + 1):
        Yn = Y[:(n + 1) ** 2]
        s = np.linalg.svd((Yn * w) @ Yn.T, compute_uv=False)
        conds.append(s.max() / (s.min() + 2.23e-7))
    nh = 0
    for i, c in enumerate(conds):
        nh = i if c < 100.0 else nh
    return nh, np.array(conds)


def np_decoder_mtx(hrtfs, dirs_deg, method, order, freqs=None, itd_s=None, weights=None, diffCM=0, maxRE=0, tdesign_deg=None):
    """getBinauralAmbiDecoderMtx, saf_hoa.c:393-450 (+ saf_hoa_internal.c:162-623), fp64.
    hrtfs: nBands x 2 x nDirs complex.  Returns nBands x 2 x nSH complex128."""
    H = np.asarray(hrtfs, np.complex128)
    nB, _, nD = H.shape
    Y = np_rsh(order, dirs_deg).astype(np.float32).astype(np.float64)     # the reference keeps Y in fp32 (:185-189)
    w = _weights(nD, weights)
    YW = Y * w[None, :]
    A = YW @ Y.T                                                          # Yna_W_Yna, :207-214
    G = np.linalg.solve(A, YW)                                            # nSH x nDirs: B = G H^H, decMtx = conj(B)^T = H G^T
    if method in (DEFAULT, LS):                                           # getBinDecoder_LS :162-228
        D = H @ G.T
    elif method == LSDIFFEQ:                                              # getBinDecoder_LSDIFFEQ :230-330
        D = H @ G.T
        Hls = D @ Y
        c_ref = np.einsum("bed,d->be", np.abs(H) ** 2, w)
        c_ls = np.einsum("bed,d->be", np.abs(Hls) ** 2, w)
        Gh = np.mean(np.sqrt(c_ref / (c_ls + 2.23e-7)), axis=1)
        D = D * Gh[:, None, None]
    elif method == TA:                                                    # getBinDecoder_TA :432-523
        # the phase term of :494-497 is exp(0 * itd/2) = 1 as written: bands above the cut-off re-use the HRTFs OF THE
        # CUT-OFF BAND unchanged.  Restated as written.
        bc = _band_cutoff(freqs)
        Hm = H.copy()
        Hm[bc:] = H[bc]
        D = Hm @ G.T
    elif method == MAGLS:                                                 # getBinDecoder_MAGLS :525-623
        bc = _band_cutoff(freqs)
        D = np.zeros((nB, 2, Y.shape[0]), np.complex128)
        D[:bc + 1] = H[:bc + 1] @ G.T
        for b in range(bc + 1, nB):
            Hmod = D[b - 1] @ Y
            ph = np.arctan2(Hmod.imag, Hmod.real)
            D[b] = (np.abs(H[b]) * np.exp(1j * ph)) @ G.T
    elif method == SPR:                                                   # getBinDecoder_SPR :332-430
        if tdesign_deg is None:
            raise ValueError("SPR needs the t-design of degree 2 * order (RefProducers.tdesign)")
        nh, _ = np_spr_order(dirs_deg, nD, weights)
        assert nh >= order, "Input order exceeds the modal order of the spatial grid"
        wspr = np.full(nD, 1.0 / nD) if weights is None else np.asarray(weights, np.float64) / (4.0 * math.pi)   # :347-353
        Ynh = np_rsh(nh, dirs_deg).astype(np.float32).astype(np.float64)
        Ytd = np_rsh(nh, tdesign_deg).astype(np.float32).astype(np.float64)
        K = Ytd.shape[1]
        Htd = (H * wspr) @ (Ynh.T @ Ytd)                                  # HRTFs interpolated to the t-design, :399-412
        D = Htd @ Ytd[:Y.shape[0]].T / K                                  # projected on the SH of the decoding order, :413-420
    else:
        raise ValueError("unknown method")
    if maxRE:                                                             # saf_hoa.c:427-445
        D = D * np_maxre(order)[None, None, :]
    if diffCM:                                                            # applyDiffCovMatching, saf_hoa.c:497-604
        for b in range(nB - 1):
            Cref = (H[b] * w) @ H[b].conj().T
            Hamb = D[b] @ Y
            Camb = (Hamb * w) @ Hamb.conj().T
            Cref[np.diag_indices(2)] = Cref[np.diag_indices(2)].real
            Camb[np.diag_indices(2)] = Camb[np.diag_indices(2)].real
            X = np.linalg.cholesky(Cref).conj().T                        # upper factor, X^H X = C
            Xa = np.linalg.cholesky(Camb).conj().T
            U, _, Vh = np.linalg.svd(Xa.conj().T @ X)
            M = np.linalg.solve(Xa, Vh.conj().T @ (U.conj().T @ X))
            D[b] = M.conj().T @ D[b]
    return D


def np_decoder_filters(hrtfs, dirs_deg, fftSize, fs, method, order, itd_s=None, weights=None, diffCM=0, maxRE=0, tdesign_deg=None):
    """getBinauralAmbiDecoderFilters, saf_hoa.c:452-497: decoding matrix per bin, then one inverse real FFT per
    (ear, SH channel) -> FLAT 2 x nSH x fftSize = the matrixConv filter layout nCHout x nCHin x length_h."""
    nB = fftSize // 2 + 1
    freqs = (np.arange(nB, dtype=np.float32) * np.float32(fs) / np.float32(fftSize))   # getUniformFreqVector
    D = np_decoder_mtx(hrtfs, dirs_deg, method, order, freqs, itd_s, weights, diffCM, maxRE, tdesign_deg)
    Dt = np.transpose(D, (1, 2, 0)).copy()
    Dt[..., 0] = Dt[..., 0].real          # kiss_fftri uses only the real parts of DC and Nyquist (kiss_fftr.c:137-138)
    Dt[..., -1] = Dt[..., -1].real
    return np.fft.irfft(Dt, n=fftSize, axis=-1)


# =====================================================================================================================
#  restatement: image-source simulator (fp32 geometry bit-for-bit, accumulation in fp64)
# =====================================================================================================================

def np_shreal_recur_f32(order: int, azi: np.ndarray, incl: np.ndarray) -> np.ndarray:
    """getSHreal_recur, saf_sh.c:255-330 (+ unnorm_legendreP_recur :129-182): the fp32 recurrences as written."""
    f = np.float32
    azi = np.asarray(azi, f); x = np.cos(np.asarray(incl, f)).astype(f)
    nD = azi.size
    Y = np.zeros(((order + 1) ** 2, nD), f)
    Y[0] = f(1.0) / f(math.sqrt(4.0 * math.pi))
    sq4pi = f(3.544907701811032)
    fact = [f(float(math.factorial(i))) for i in range(2 * order + 2)]
    leg1 = np.zeros((order + 1, nD), f); leg2 = np.zeros((order + 1, nD), f)
    x2 = (x * x).astype(f)
    for n in range(1, order + 1):
        leg = np.zeros((order + 1, nD), f)
        if n == 1:
            leg[0] = x; leg[1] = np.sqrt(f(1.0) - x2)
        elif n == 2:
            leg[0] = (f(3.0) * x2 - f(1.0)) / f(2.0)
            leg[1] = x * f(3.0) * np.sqrt(f(1.0) - x2)
            leg[2] = f(3.0) * (f(1.0) - x2)
        else:
            k = 2 * n - 1
            df = f(1.0)
            for kk in range(1, (k + 1) // 2 + 1):
                df = f(df * f(2.0 * kk - 1.0))
            leg[n] = df * np.power(f(1.0) - x2, f(n / 2.0), dtype=f)
            leg[n - 1] = f(k) * x * leg1[n - 1]
            for m in range(n - 1):
                leg[m] = ((f(k) * x * leg1[m]) - (f(n + m - 1) * leg2[m])) / f(n - m)
        Nn0 = f(np.sqrt(f(2.0 * n + 1.0)))
        i0 = n * n
        for m in range(n + 1):
            if m == 0:
                Y[i0 + n] = Nn0 / sq4pi * leg[0]
            else:
                Nnm = f(Nn0 * np.sqrt(f(2.0) * fact[n - m] / fact[n + m], dtype=f))
                Y[i0 + n - m] = Nnm / sq4pi * leg[m] * np.sin(f(m) * azi).astype(f)
                Y[i0 + n + m] = Nnm / sq4pi * leg[m] * np.cos(f(m) * azi).astype(f)
        leg2 = leg1; leg1 = leg
    return Y


def np_ims_rir(room, abs_wall, n_bands, c_ms, fs, src, rec, sh_order, maxN=-1, maxTime_s=-1.0):
    """One source / receiver pair through ims_shoebox_computeEchograms + ims_shoebox_renderRIRs
    (saf_reverb.c:184-295; saf_reverb_internal.c:269-397 coreInitT, :399-521 coreInitN, :523-574 receiver SH module,
    :576-638 absorption module, :640-710 renderRIR).  Geometry in fp32 exactly as written (the tap index of every image
    source must match bit for bit); the taps are accumulated in fp64.

    As written, renderRIR convolves every band's RIR with its filterbank FIR into a scratch buffer that is never read
    (:697-698 write `temp`, :701-702 sum the UNFILTERED `rir_bands`), so the result is the plain sum over bands of the
    band echograms.  Restated as written."""
    f = np.float32
    room = np.asarray(room, f); aw = np.asarray(abs_wall, f).reshape(n_bands, 6)
    src = np.asarray(src, f); rec = np.asarray(rec, f)
    c_ms = f(c_ms); fs = f(fs)
    # computeEchograms :206-212 flips y, coreInit :288-294 moves the origin to the room centre
    s2 = np.array([src[0], room[1] - src[1], src[2]], f); r2 = np.array([rec[0], room[1] - rec[1], rec[2]], f)
    so = np.array([s2[0] - room[0] / f(2), room[1] / f(2) - s2[1], s2[2] - room[2] / f(2)], f)
    ro = np.array([r2[0] - room[0] / f(2), room[1] / f(2) - r2[1], r2[2] - room[2] / f(2)], f)
    if maxTime_s > 0:
        d_max = f(f(maxTime_s) * c_ms)
        Nx = int(d_max / room[0] + f(1.0)); Ny = int(d_max / room[1] + f(1.0)); Nz = int(d_max / room[2] + f(1.0))
        ii, jj, kk = np.meshgrid(np.arange(-Nx, Nx + 1), np.arange(-Ny, Ny + 1), np.arange(-Nz, Nz + 1), indexing="ij")
        # lattice order of the reference: i fastest, then j, then k
        I = ii.transpose(2, 1, 0).ravel(); J = jj.transpose(2, 1, 0).ravel(); K = kk.transpose(2, 1, 0).ravel()
    else:
        N = int(maxN)
        ii, jj, kk = np.meshgrid(np.arange(-N, N + 1), np.arange(-N, N + 1), np.arange(-N, N + 1), indexing="ij")
        I = ii.transpose(2, 1, 0).ravel(); J = jj.transpose(2, 1, 0).ravel(); K = kk.transpose(2, 1, 0).ravel()
        keep = (np.abs(I) + np.abs(J) + np.abs(K)) <= N
        I, J, K = I[keep], J[keep], K[keep]
    sgn = lambda a: np.where(a % 2 == 0, f(1.0), f(-1.0)).astype(f)
    sx = ((I.astype(f) * room[0]).astype(f) + (sgn(I) * so[0]).astype(f)).astype(f) - ro[0]
    sy = ((J.astype(f) * room[1]).astype(f) + (sgn(J) * so[1]).astype(f)).astype(f) - ro[1]
    sz = ((K.astype(f) * room[2]).astype(f) + (sgn(K) * so[2]).astype(f)).astype(f) - ro[2]
    d = np.sqrt((((sx * sx).astype(f) + (sy * sy).astype(f)).astype(f) + (sz * sz).astype(f)).astype(f)).astype(f)
    if maxTime_s > 0:
        keep = d < d_max
        I, J, K, sx, sy, sz, d = I[keep], J[keep], K[keep], sx[keep], sy[keep], sz[keep], d[keep]
    t = (d / c_ms).astype(f)
    att = np.where(d <= f(1.0), f(1.0), (f(1.0) / d).astype(f)).astype(f)
    # receiver module :556-569: unitCart2sph + getSHreal_recur
    nSH = (sh_order + 1) ** 2
    if sh_order == 0:
        g = np.ones((1, d.size), f)
    else:
        hyp = np.sqrt((sx * sx + sy * sy).astype(f)).astype(f)
        azi = np.arctan2(sy, sx).astype(f)
        elev = np.arctan2(sz, hyp).astype(f)
        incl = (f(math.pi) / f(2.0) - elev).astype(f)
        g = np_shreal_recur_f32(sh_order, azi, incl)
    # absorption module :576-638
    tot = np.zeros(d.size)
    for b in range(n_bands):
        r = np.sqrt(f(1.0) - aw[b]).astype(f)
        def axis(o, r0, r1):
            a = np.abs(o)
            ev = (np.power(r0, (a / 2.0).astype(f)) * np.power(r1, (a / 2.0).astype(f))).astype(f)
            op = (np.power(r0, np.ceil(a / 2.0).astype(f)) * np.power(r1, np.floor(a / 2.0).astype(f))).astype(f)
            on = (np.power(r0, np.floor(a / 2.0).astype(f)) * np.power(r1, np.ceil(a / 2.0).astype(f))).astype(f)
            return np.where(o % 2 == 0, ev, np.where(o > 0, op, on)).astype(f)
        tot += ((axis(I, r[0], r[1]) * axis(J, r[2], r[3])).astype(f) * axis(K, r[4], r[5])).astype(f).astype(np.float64)
    endtime = t.max()
    length = int(f(endtime * fs) + f(1.0)) + 1
    idx = ((t * fs).astype(f) + f(0.5)).astype(f).astype(np.int64)
    rir = np.zeros((nSH, length))
    vals = g.astype(np.float64) * att.astype(np.float64)[None, :] * tot[None, :]
    for ch in range(nSH):
        np.add.at(rir[ch], idx, vals[ch])
    return rir, idx
