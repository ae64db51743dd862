"""TEST INFRASTRUCTURE ONLY -- ctypes access to the two CPU checkers.

* ``load_oracle()``     our plain-C restatement (``oracle/safconv_oracle.c``)
* ``load_reference()``  the unmodified reference convolver compiled in place from
                        ``/root/reference`` into ``oracle/_ref/`` (see ``oracle/Makefile``)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``spatial_audio_framework_b200`` / ``libsafconv_b200.so``) never does.

Reference API being wrapped:
``/root/reference/framework/modules/saf_utilities/saf_utility_matrixConv.h:55-190``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_ORACLE_SO = HERE / "_build" / "libsafconv_oracle.so"
_REF_DIR = HERE / "_ref"

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def build(verbose: bool = False) -> None:
    """Compile the oracle restatement and (if /root/reference exists) oracle/_ref."""
    out = subprocess.run(["make", "-C", str(HERE), "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


_oracle_lib = None


def load_oracle():
    global _oracle_lib
    if _oracle_lib is not None:
        return _oracle_lib
    if not _ORACLE_SO.exists():
        build()
    lib = C.CDLL(str(_ORACLE_SO))
    lib.orc_matrixConv_create.restype = C.c_void_p
    lib.orc_matrixConv_create.argtypes = [C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_matrixConv_apply.argtypes = [C.c_void_p, _f32p, _f32p]
    lib.orc_matrixConv_destroy.argtypes = [C.c_void_p]
    lib.orc_multiConv_create.restype = C.c_void_p
    lib.orc_multiConv_create.argtypes = [C.c_int, _f32p, C.c_int, C.c_int, C.c_int]
    lib.orc_multiConv_apply.argtypes = [C.c_void_p, _f32p, _f32p]
    lib.orc_multiConv_destroy.argtypes = [C.c_void_p]
    lib.orc_TVConv_create.restype = C.c_void_p
    lib.orc_TVConv_create.argtypes = [C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_TVConv_apply.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
    lib.orc_TVConv_destroy.argtypes = [C.c_void_p]
    lib.orc_rfft_forward.argtypes = [C.c_int, _f32p, _f32p]
    lib.orc_rfft_backward.argtypes = [C.c_int, _f32p, _f32p]
    lib.orc_truth_matrix.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_long,
                                     C.POINTER(C.c_int), C.c_int, C.c_long, C.c_long, _f64p]
    lib.orc_truth_multi.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_long,
                                    C.POINTER(C.c_int), C.c_int, C.c_long, C.c_long, _f64p]
    for f in (lib.orc_fftconv, lib.orc_fftfilt):
        f.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p]
        f.restype = None
    _oracle_lib = lib
    return lib


_ref_lib = None
_ref_kind = None


def reference_available() -> bool:
    return any(_REF_DIR.glob("libsaf_ref_conv_*.so"))


def load_reference():
    """Load oracle/_ref (OpenBLAS-linked variant first, level-1 shim variant otherwise).

    Returns (lib, kind) with kind in {"openblas", "l1shim"}.
    """
    global _ref_lib, _ref_kind
    if _ref_lib is not None:
        return _ref_lib, _ref_kind
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")   # the reference convolver is single-threaded
    err = None
    for kind in ("openblas", "l1shim"):
        p = _REF_DIR / f"libsaf_ref_conv_{kind}.so"
        if not p.exists():
            continue
        try:
            lib = C.CDLL(str(p))
        except OSError as e:  # e.g. OpenBLAS missing on this machine
            err = e
            continue
        vpp = C.POINTER(C.c_void_p)
        lib.saf_matrixConv_create.argtypes = [vpp, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.saf_matrixConv_apply.argtypes = [C.c_void_p, _f32p, _f32p]
        lib.saf_matrixConv_destroy.argtypes = [vpp]
        lib.saf_multiConv_create.argtypes = [vpp, C.c_int, _f32p, C.c_int, C.c_int, C.c_int]
        lib.saf_multiConv_apply.argtypes = [C.c_void_p, _f32p, _f32p]
        lib.saf_multiConv_destroy.argtypes = [vpp]
        lib.saf_TVConv_create.argtypes = [vpp, C.c_int, C.POINTER(_f32p), C.c_int, C.c_int, C.c_int, C.c_int]
        lib.saf_TVConv_apply.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
        lib.saf_TVConv_destroy.argtypes = [vpp]
        lib.saf_rfft_create.argtypes = [vpp, C.c_int]
        lib.saf_rfft_forward.argtypes = [C.c_void_p, _f32p, _f32p]
        lib.saf_rfft_backward.argtypes = [C.c_void_p, _f32p, _f32p]
        lib.saf_rfft_destroy.argtypes = [vpp]
        for f in ("saf_matrixConv_create", "saf_matrixConv_apply", "saf_matrixConv_destroy",
                  "saf_multiConv_create", "saf_multiConv_apply", "saf_multiConv_destroy",
                  "saf_TVConv_create", "saf_TVConv_apply", "saf_TVConv_destroy",
                  "saf_rfft_create", "saf_rfft_forward", "saf_rfft_backward", "saf_rfft_destroy"):
            getattr(lib, f).restype = None
        _ref_lib, _ref_kind = lib, kind
        return lib, kind
    raise RuntimeError(f"no loadable reference library under {_REF_DIR} ({err})")


# ----------------------------------------------------------------------------
# numpy-level wrappers with one common interface:  obj.apply(in[nIn,hop]) -> out[nOut,hop]
# ----------------------------------------------------------------------------

class _Conv:
    nIn: int
    nOut: int
    hop: int

    def apply(self, x: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def run(self, x: np.ndarray) -> np.ndarray:
        """Process a whole signal x[nIn, T] (T multiple of hop) block by block -> y[nOut, T]."""
        nblk = x.shape[1] // self.hop
        y = np.empty((self.nOut, nblk * self.hop), np.float32)
        for b in range(nblk):
            blk = np.ascontiguousarray(x[:, b * self.hop:(b + 1) * self.hop])
            y[:, b * self.hop:(b + 1) * self.hop] = self.apply(blk)
        return y


class OracleMatrixConv(_Conv):
    def __init__(self, hop, H, usePart=1):
        H = np.ascontiguousarray(H, np.float32)
        self.nOut, self.nIn, self.len = H.shape
        self.hop = hop
        self._lib = load_oracle()
        self._h = self._lib.orc_matrixConv_create(hop, _fp(H), self.len, self.nIn, self.nOut, usePart)

    def apply(self, x):
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty((self.nOut, self.hop), np.float32)
        self._lib.orc_matrixConv_apply(self._h, _fp(x), _fp(y))
        return y

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orc_matrixConv_destroy(self._h)
            self._h = None


class OracleMultiConv(_Conv):
    def __init__(self, hop, H, usePart=1):
        H = np.ascontiguousarray(H, np.float32)
        self.nOut, self.len = H.shape
        self.nIn = self.nOut
        self.hop = hop
        self._lib = load_oracle()
        self._h = self._lib.orc_multiConv_create(hop, _fp(H), self.len, self.nOut, usePart)

    def apply(self, x):
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty((self.nOut, self.hop), np.float32)
        self._lib.orc_multiConv_apply(self._h, _fp(x), _fp(y))
        return y

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orc_multiConv_destroy(self._h)
            self._h = None


class OracleTVConv:
    """H: [nIRs, nOut, len]; one input channel."""

    def __init__(self, hop, H, initIdx=0):
        H = np.ascontiguousarray(H, np.float32)
        self.nIRs, self.nOut, self.len = H.shape
        self.hop = hop
        self._lib = load_oracle()
        self._h = self._lib.orc_TVConv_create(hop, _fp(H), self.len, self.nIRs, self.nOut, initIdx)

    def apply(self, x, irIdx):
        x = np.ascontiguousarray(x, np.float32).reshape(-1)
        y = np.empty((self.nOut, self.hop), np.float32)
        self._lib.orc_TVConv_apply(self._h, _fp(x), _fp(y), int(irIdx))
        return y

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orc_TVConv_destroy(self._h)
            self._h = None


class RefMatrixConv(_Conv):
    """The compiled reference: saf_matrixConv_* (saf_utility_matrixConv.c:49-236)."""

    def __init__(self, hop, H, usePart=1):
        H = np.ascontiguousarray(H, np.float32)
        self.nOut, self.nIn, self.len = H.shape
        self.hop = hop
        self._lib, self.kind = load_reference()
        self._h = C.c_void_p()
        self._lib.saf_matrixConv_create(C.byref(self._h), hop, _fp(H), self.len, self.nIn, self.nOut, usePart)

    def apply(self, x):
        x = np.ascontiguousarray(x, np.float32)
        y = np.zeros((self.nOut, self.hop), np.float32)
        self._lib.saf_matrixConv_apply(self._h, _fp(x), _fp(y))
        return y

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.saf_matrixConv_destroy(C.byref(self._h))
            self._h = None


class RefMultiConv(_Conv):
    """The compiled reference: saf_multiConv_* (saf_utility_matrixConv.c:257-414)."""

    def __init__(self, hop, H, usePart=1):
        H = np.ascontiguousarray(H, np.float32)
        self.nOut, self.len = H.shape
        self.nIn = self.nOut
        self.hop = hop
        self._lib, self.kind = load_reference()
        self._h = C.c_void_p()
        self._lib.saf_multiConv_create(C.byref(self._h), hop, _fp(H), self.len, self.nOut, usePart)

    def apply(self, x):
        x = np.ascontiguousarray(x, np.float32)
        y = np.zeros((self.nOut, self.hop), np.float32)
        self._lib.saf_multiConv_apply(self._h, _fp(x), _fp(y))
        return y

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.saf_multiConv_destroy(C.byref(self._h))
            self._h = None


class RefTVConv:
    """The compiled reference: saf_TVConv_* (saf_utility_matrixConv.c:441-620)."""

    def __init__(self, hop, H, initIdx=0):
        H = np.ascontiguousarray(H, np.float32)
        self.nIRs, self.nOut, self.len = H.shape
        self.hop = hop
        self._H = H
        self._lib, self.kind = load_reference()
        rows = (_f32p * self.nIRs)(*[_fp(H[i]) for i in range(self.nIRs)])
        self._h = C.c_void_p()
        self._lib.saf_TVConv_create(C.byref(self._h), hop, rows, self.len, self.nIRs, self.nOut, initIdx)

    def apply(self, x, irIdx):
        x = np.ascontiguousarray(x, np.float32).reshape(-1)
        y = np.zeros((self.nOut, self.hop), np.float32)
        self._lib.saf_TVConv_apply(self._h, _fp(x), _fp(y), int(irIdx))
        return y

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.saf_TVConv_destroy(C.byref(self._h))
            self._h = None


def ref_rfft(N: int, x: np.ndarray):
    """saf_rfft_forward/backward round trip pieces from the compiled reference (saf_utility_fft.c:531-753)."""
    lib, _ = load_reference()
    h = C.c_void_p()
    lib.saf_rfft_create(C.byref(h), N)
    x = np.ascontiguousarray(x, np.float32)
    X = np.zeros(2 * (N // 2 + 1), np.float32)
    lib.saf_rfft_forward(h, _fp(x), _fp(X))
    xb = np.zeros(N, np.float32)
    lib.saf_rfft_backward(h, _fp(X), _fp(xb))
    lib.saf_rfft_destroy(C.byref(h))
    return X, xb


def oracle_rfft(N: int, x: np.ndarray):
    lib = load_oracle()
    x = np.ascontiguousarray(x, np.float32)
    X = np.zeros(2 * (N // 2 + 1), np.float32)
    lib.orc_rfft_forward(N, _fp(x), _fp(X))
    xb = np.zeros(N, np.float32)
    lib.orc_rfft_backward(N, _fp(X), _fp(xb))
    return X, xb


def oracle_fftconv(x: np.ndarray, h: np.ndarray, filt: bool = False) -> np.ndarray:
    """fftconv / fftfilt restatement (saf_utility_fft.c:157-228): x[nCH, x_len], h[nCH, h_len]."""
    lib = load_oracle()
    x = np.ascontiguousarray(x, np.float32); h = np.ascontiguousarray(h, np.float32)
    nCH, xl = x.shape; hl = h.shape[1]
    y = np.zeros((nCH, xl if filt else xl + hl - 1), np.float32)
    (lib.orc_fftfilt if filt else lib.orc_fftconv)(_fp(x), _fp(h), xl, hl, nCH, _fp(y))
    return y


def ref_fftconv(x: np.ndarray, h: np.ndarray, filt: bool = False) -> np.ndarray:
    """The compiled reference's own fftconv / fftfilt."""
    lib, _ = load_reference()
    x = np.ascontiguousarray(x, np.float32); h = np.ascontiguousarray(h, np.float32)
    nCH, xl = x.shape; hl = h.shape[1]
    y = np.zeros((nCH, xl if filt else xl + hl - 1), np.float32)
    fn = lib.fftfilt if filt else lib.fftconv
    fn.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p]; fn.restype = None
    fn(_fp(x), _fp(h), xl, hl, nCH, _fp(y))
    return y


def truth_matrix(H: np.ndarray, x: np.ndarray, outs, n0: int, n1: int) -> np.ndarray:
    """fp64 direct convolution y[outs, n0:n1] for H[nOut,nIn,len], x[nIn,T]."""
    lib = load_oracle()
    H = np.ascontiguousarray(H, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    nOut, nIn, ln = H.shape
    outs = np.ascontiguousarray(outs, np.int32)
    y = np.zeros((len(outs), n1 - n0), np.float64)
    lib.orc_truth_matrix(_fp(H), ln, nIn, nOut, _fp(x), x.shape[1],
                         outs.ctypes.data_as(C.POINTER(C.c_int)), len(outs), n0, n1,
                         y.ctypes.data_as(_f64p))
    return y


def truth_multi(H: np.ndarray, x: np.ndarray, chans, n0: int, n1: int) -> np.ndarray:
    lib = load_oracle()
    H = np.ascontiguousarray(H, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    nCH, ln = H.shape
    chans = np.ascontiguousarray(chans, np.int32)
    y = np.zeros((len(chans), n1 - n0), np.float64)
    lib.orc_truth_multi(_fp(H), ln, nCH, _fp(x), x.shape[1],
                        chans.ctypes.data_as(C.POINTER(C.c_int)), len(chans), n0, n1,
                        y.ctypes.data_as(_f64p))
    return y
