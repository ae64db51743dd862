"""ctypes binding of libsafconv_b200.so (declarations: include/safconv_b200.h)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libsafconv_b200.so"

_f32p = C.POINTER(C.c_float)
_vpp = C.POINTER(C.c_void_p)


class SafConvError(RuntimeError):
    pass


class SafConvInfo(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("hopSize", C.c_int), ("length_h", C.c_int), ("nCHin", C.c_int),
        ("nCHout", C.c_int), ("nOutLocal", C.c_int), ("outBegin", C.c_int),
        ("fftSize", C.c_int), ("nBinsPacked", C.c_int), ("numFilterBlocks", C.c_int),
        ("macGrid", C.c_int), ("macStages", C.c_int), ("macThreads", C.c_int),
        ("maxBatch", C.c_int), ("device", C.c_int),
        ("bytesFilters", C.c_size_t), ("bytesDelayLine", C.c_size_t),
        ("algBytesPerBlock", C.c_double), ("macAlgBytesPerBlock", C.c_double),
    ]


# every symbol include/safconv_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "saf_matrixConv_create", "saf_matrixConv_destroy", "saf_matrixConv_apply",
    "saf_multiConv_create", "saf_multiConv_destroy", "saf_multiConv_apply",
    "saf_TVConv_create", "saf_TVConv_destroy", "saf_TVConv_apply",
    "safconv_last_error", "safconv_last_error_string", "safconv_version", "safconv_set_device",
    "safconv_matrixConv_create_shard", "safconv_matrixConv_create_from_shard", "safconv_multiConv_create_shard",
    "safconv_matrixConv_create_multi", "safconv_multiConv_create_multi", "safconv_multi_get_devices", "safconv_multi_get_shard",
    "safconv_apply_device", "safconv_apply_device_blocks",
    "safconv_set_stream", "safconv_get_stream", "safconv_synchronize", "safconv_reset_state",
    "safconv_get_info", "safconv_enable_kernel_timing", "safconv_get_kernel_times", "safconv_get_kernel_totals",
    "safconv_set_option",
    "safconv_render_offline", "safconv_render_offline_segment", "safconv_render_offline_device",
    "safconv_render_offline_segment_device",
    "safconv_get_offline_times",
    "safconv_fftconv", "safconv_fftfilt", "fftconv", "fftfilt",
    "safconv_rfft_forward", "safconv_rfft_backward",
    "safconv_set_true_mode0",
    "saf_rfft_create", "saf_rfft_destroy", "saf_rfft_forward", "saf_rfft_backward",
    "safconv_rfft_batch", "safconv_rfft_last_error", "safconv_rfft_last_error_string", "safconv_rfft_get_factors",
]


def build(verbose: bool = False) -> Path:
    """Compile libsafconv_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", str(PKG_DIR / "csrc")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0 or not LIB_PATH.exists():
        raise SafConvError("building libsafconv_b200.so failed")
    return LIB_PATH


_lib = None


def lib():
    """Load the product library.  Fails loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SafConvError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is the only implementation; there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    L.saf_matrixConv_create.argtypes = [_vpp, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.saf_matrixConv_destroy.argtypes = [_vpp]
    L.saf_matrixConv_apply.argtypes = [C.c_void_p, _f32p, _f32p]
    L.saf_multiConv_create.argtypes = [_vpp, C.c_int, _f32p, C.c_int, C.c_int, C.c_int]
    L.saf_multiConv_destroy.argtypes = [_vpp]
    L.saf_multiConv_apply.argtypes = [C.c_void_p, _f32p, _f32p]
    L.saf_TVConv_create.argtypes = [_vpp, C.c_int, C.POINTER(_f32p), C.c_int, C.c_int, C.c_int, C.c_int]
    L.saf_TVConv_destroy.argtypes = [_vpp]
    L.saf_TVConv_apply.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
    for f in ("saf_matrixConv_create", "saf_matrixConv_destroy", "saf_matrixConv_apply",
              "saf_multiConv_create", "saf_multiConv_destroy", "saf_multiConv_apply",
              "saf_TVConv_create", "saf_TVConv_destroy", "saf_TVConv_apply"):
        getattr(L, f).restype = None
    L.safconv_last_error.argtypes = [C.c_void_p]
    L.safconv_last_error.restype = C.c_int
    L.safconv_last_error_string.argtypes = [C.c_void_p]
    L.safconv_last_error_string.restype = C.c_char_p
    L.safconv_version.restype = C.c_char_p
    L.safconv_set_device.argtypes = [C.c_int]
    L.safconv_matrixConv_create_shard.argtypes = [_vpp, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.safconv_matrixConv_create_shard.restype = None
    L.safconv_matrixConv_create_from_shard.argtypes = [_vpp, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.safconv_matrixConv_create_from_shard.restype = None
    L.safconv_multiConv_create_shard.argtypes = [_vpp, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.safconv_multiConv_create_shard.restype = None
    _ip = C.POINTER(C.c_int)
    L.safconv_matrixConv_create_multi.argtypes = [_vpp, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _ip, C.c_int]
    L.safconv_matrixConv_create_multi.restype = None
    L.safconv_multiConv_create_multi.argtypes = [_vpp, C.c_int, _f32p, C.c_int, C.c_int, _ip, C.c_int]
    L.safconv_multiConv_create_multi.restype = None
    L.safconv_multi_get_devices.argtypes = [C.c_void_p, _ip, C.c_int]
    L.safconv_multi_get_shard.argtypes = [C.c_void_p, C.c_int]
    L.safconv_multi_get_shard.restype = C.c_void_p
    L.safconv_apply_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.safconv_apply_device_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.safconv_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.safconv_get_stream.argtypes = [C.c_void_p]
    L.safconv_get_stream.restype = C.c_void_p
    L.safconv_synchronize.argtypes = [C.c_void_p]
    L.safconv_reset_state.argtypes = [C.c_void_p]
    L.safconv_get_info.argtypes = [C.c_void_p, C.POINTER(SafConvInfo)]
    L.safconv_enable_kernel_timing.argtypes = [C.c_void_p, C.c_int]
    L.safconv_get_kernel_times.argtypes = [C.c_void_p, _f32p, C.POINTER(C.c_int)]
    L.safconv_get_kernel_totals.argtypes = [C.c_void_p, _f32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.safconv_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.safconv_debug_plan_size.restype = C.c_int
    L.safconv_render_offline.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
    L.safconv_render_offline_segment.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int]
    L.safconv_render_offline_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.safconv_render_offline_segment_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.safconv_get_offline_times.argtypes = [C.c_void_p, _f32p]
    L.saf_rfft_create.argtypes = [_vpp, C.c_int]
    L.saf_rfft_destroy.argtypes = [_vpp]
    L.saf_rfft_forward.argtypes = [C.c_void_p, _f32p, C.c_void_p]
    L.saf_rfft_backward.argtypes = [C.c_void_p, C.c_void_p, _f32p]
    for f in ("saf_rfft_create", "saf_rfft_destroy", "saf_rfft_forward", "saf_rfft_backward"):
        getattr(L, f).restype = None
    L.safconv_rfft_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p]
    L.safconv_rfft_last_error.argtypes = [C.c_void_p]
    L.safconv_rfft_last_error_string.argtypes = [C.c_void_p]
    L.safconv_rfft_last_error_string.restype = C.c_char_p
    L.safconv_rfft_get_factors.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
    L.safconv_debug_fft_factors.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int]
    L.safconv_debug_convolver_rfft.argtypes = [C.c_int, C.c_int, _f32p, _f32p, C.c_int]
    _lib = L
    return L


def version() -> str:
    return lib().safconv_version().decode()


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


class _Base:
    """Common part of the three convolver wrappers (handle lifetime, errors, extension calls)."""

    _destroy_name = ""

    def __init__(self):
        self._lib = lib()
        self._h = C.c_void_p()

    # -- lifetime ---------------------------------------------------------------------------
    def _check_created(self, what: str):
        if not self._h:
            msg = self._lib.safconv_last_error_string(None).decode()
            raise SafConvError(f"{what} failed: {msg}")

    def destroy(self):
        if getattr(self, "_h", None):
            getattr(self._lib, self._destroy_name)(C.byref(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def _raise_if_error(self):
        code = self._lib.safconv_last_error(self._h)
        if code:
            raise SafConvError(self._lib.safconv_last_error_string(self._h).decode())

    # -- extension surface ------------------------------------------------------------------
    @property
    def handle(self):
        return self._h

    def info(self) -> SafConvInfo:
        i = SafConvInfo()
        if self._lib.safconv_get_info(self._h, C.byref(i)):
            raise SafConvError("safconv_get_info failed")
        return i

    def multi_devices(self):
        """Devices of a multi-GPU handle ([] for a single-device handle)."""
        buf = (C.c_int * 16)()
        n = self._lib.safconv_multi_get_devices(self._h, buf, 16)
        return [buf[i] for i in range(n)]

    def shard_info(self, i: int) -> SafConvInfo:
        """safconv_get_info of device i's shard of a multi-GPU handle."""
        sh = self._lib.safconv_multi_get_shard(self._h, int(i))
        info = SafConvInfo()
        if not sh or self._lib.safconv_get_info(C.c_void_p(sh), C.byref(info)):
            raise SafConvError("not a multi-GPU handle / no such shard")
        return info

    def set_stream(self, cuda_stream_ptr: int | None):
        self._lib.safconv_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0))

    def synchronize(self):
        if self._lib.safconv_synchronize(self._h):
            self._raise_if_error()

    def reset_state(self):
        if self._lib.safconv_reset_state(self._h):
            self._raise_if_error()

    def set_option(self, name: str, value: int):
        if self._lib.safconv_set_option(self._h, name.encode(), int(value)):
            raise SafConvError(f"unknown option {name}")

    def apply_device(self, d_in_ptr: int, d_out_ptr: int, n_blocks: int = 1):
        """Enqueue n_blocks blocks on device pointers (no host sync)."""
        if self._lib.safconv_apply_device_blocks(self._h, C.c_void_p(d_in_ptr), C.c_void_p(d_out_ptr), n_blocks):
            self._raise_if_error()

    def enable_kernel_timing(self, n_blocks: int):
        """Record CUDA events around the kernels of the next n_blocks blocks (0 disables)."""
        if self._lib.safconv_enable_kernel_timing(self._h, int(n_blocks)):
            self._raise_if_error()

    def kernel_times_ms(self):
        """(avg ms [fwd FFT, MAC, iFFT+OLA], number of blocks averaged); restarts the event ring."""
        ms = (C.c_float * 3)()
        n = C.c_int(0)
        if self._lib.safconv_get_kernel_times(self._h, ms, C.byref(n)):
            self._raise_if_error()
        return [ms[0], ms[1], ms[2]], n.value

    def kernel_totals_ms(self):
        """(total ms [fwd FFT, MAC, iFFT+OLA], launch groups, blocks covered); restarts the recording."""
        ms = (C.c_float * 3)()
        g, n = C.c_int(0), C.c_int(0)
        if self._lib.safconv_get_kernel_totals(self._h, ms, C.byref(g), C.byref(n)):
            self._raise_if_error()
        return [ms[0], ms[1], ms[2]], g.value, n.value

    def run(self, x: np.ndarray) -> np.ndarray:
        """Block-by-block processing of a whole signal x[nIn, T] -> y[nOut, T] via the host API."""
        hop = self.hop
        nblk = x.shape[1] // hop
        y = np.empty((self.nOutLocal, nblk * hop), np.float32)
        for b in range(nblk):
            y[:, b * hop:(b + 1) * hop] = self.apply(np.ascontiguousarray(x[:, b * hop:(b + 1) * hop]))
        return y


class MatrixConv(_Base):
    """saf_matrixConv_create/apply/destroy (reference saf_utility_matrixConv.h:55-86).

    H: [nCHout, nCHin, length_h] float32.  ``shard=(outBegin, outCount)`` builds a handle that owns
    only those output channels (safconv_matrixConv_create_shard).
    """

    _destroy_name = "saf_matrixConv_destroy"

    def __init__(self, hopSize: int, H: np.ndarray, usePartFLAG: int = 1, shard=None, device: int | None = None,
                 _from_shard=None, devices=None):
        super().__init__()
        H = np.ascontiguousarray(H, np.float32)
        self.nCHout, self.nCHin, self.length_h = H.shape
        self.hop = int(hopSize)
        if device is not None:
            self._lib.safconv_set_device(int(device))
        if devices is not None:
            # one handle over several GPUs (safconv_matrixConv_create_multi); used with the unchanged apply / destroy
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            self.nOutLocal = self.nCHout
            self._lib.safconv_matrixConv_create_multi(C.byref(self._h), self.hop, _fp(H), self.length_h,
                                                      self.nCHin, self.nCHout, devs, len(devices))
        elif _from_shard is not None:
            nOutTotal, ob = _from_shard
            self.nOutLocal, self.nCHout = H.shape[0], int(nOutTotal)
            self._lib.safconv_matrixConv_create_from_shard(C.byref(self._h), self.hop, _fp(H), self.length_h,
                                                           self.nCHin, self.nCHout, int(ob), self.nOutLocal)
        elif shard is None:
            self.nOutLocal = self.nCHout
            self._lib.saf_matrixConv_create(C.byref(self._h), self.hop, _fp(H), self.length_h,
                                            self.nCHin, self.nCHout, int(usePartFLAG))
        else:
            ob, oc = shard
            self.nOutLocal = oc
            self._lib.safconv_matrixConv_create_shard(C.byref(self._h), self.hop, _fp(H), self.length_h,
                                                      self.nCHin, self.nCHout, int(ob), int(oc))
        self._check_created("saf_matrixConv_create")

    @classmethod
    def from_shard(cls, hopSize: int, Hshard: np.ndarray, nCHoutTotal: int, outBegin: int, device: int | None = None):
        """Hshard: [outCount, nCHin, length_h] = this rank's output channels only (safconv_matrixConv_create_from_shard)."""
        return cls(hopSize, Hshard, 1, device=device, _from_shard=(nCHoutTotal, outBegin))

    def render_offline(self, x: np.ndarray) -> np.ndarray:
        """Whole signal x[nCHin, nFrames*hop] -> y[nOutLocal, nFrames*hop] on the tensor-core offline path."""
        x = np.ascontiguousarray(x, np.float32)
        nfr = x.shape[1] // self.hop
        assert x.shape == (self.nCHin, nfr * self.hop)
        y = np.empty((self.nOutLocal, nfr * self.hop), np.float32)
        if self._lib.safconv_render_offline(self._h, _fp(x), _fp(y), nfr):
            self._raise_if_error()
        return y

    def render_offline_host(self, in_ptr: int, out_ptr: int, n_frames: int, n_halo: int = 0):
        """Host pointers (ideally page-locked): in [nCHin][(n_halo+n_frames)*hop] -> out [nOutLocal][n_frames*hop]."""
        fp = C.POINTER(C.c_float)
        if self._lib.safconv_render_offline_segment(self._h, C.cast(in_ptr, fp), C.cast(out_ptr, fp), n_frames, n_halo):
            self._raise_if_error()

    def render_offline_device(self, d_in_ptr: int, d_out_ptr: int, n_frames: int):
        if self._lib.safconv_render_offline_device(self._h, C.c_void_p(d_in_ptr), C.c_void_p(d_out_ptr), n_frames):
            self._raise_if_error()

    def render_offline_segment_device(self, d_in_ptr: int, d_out_ptr: int, n_frames: int, n_halo: int):
        if self._lib.safconv_render_offline_segment_device(self._h, C.c_void_p(d_in_ptr), C.c_void_p(d_out_ptr),
                                                           n_frames, n_halo):
            self._raise_if_error()

    def offline_times_ms(self):
        ms = (C.c_float * 3)()
        if self._lib.safconv_get_offline_times(self._h, ms):
            self._raise_if_error()
        return [ms[0], ms[1], ms[2]]

    def apply(self, inputSigs: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(inputSigs, np.float32)
        assert x.shape == (self.nCHin, self.hop)
        y = np.empty((self.nOutLocal, self.hop), np.float32)
        self._lib.saf_matrixConv_apply(self._h, _fp(x), _fp(y))
        self._raise_if_error()
        return y


class MultiConv(_Base):
    """saf_multiConv_create/apply/destroy (reference saf_utility_matrixConv.h:109-136).  H: [nCH, length_h]."""

    _destroy_name = "saf_multiConv_destroy"

    def __init__(self, hopSize: int, H: np.ndarray, usePartFLAG: int = 1, shard=None, device: int | None = None,
                 devices=None):
        super().__init__()
        H = np.ascontiguousarray(H, np.float32)
        self.nCH, self.length_h = H.shape
        self.hop = int(hopSize)
        if device is not None:
            self._lib.safconv_set_device(int(device))
        if devices is not None:
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            self.nOutLocal = self.nCH
            self._lib.safconv_multiConv_create_multi(C.byref(self._h), self.hop, _fp(H), self.length_h, self.nCH,
                                                     devs, len(devices))
        elif shard is None:
            self.nOutLocal = self.nCH
            self._lib.saf_multiConv_create(C.byref(self._h), self.hop, _fp(H), self.length_h, self.nCH, int(usePartFLAG))
        else:
            cb, cc = shard
            self.nOutLocal = cc
            self._lib.safconv_multiConv_create_shard(C.byref(self._h), self.hop, _fp(H), self.length_h,
                                                     self.nCH, int(cb), int(cc))
        self.nCHin = self.nOutLocal
        self._check_created("saf_multiConv_create")

    def apply(self, inputSigs: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(inputSigs, np.float32)
        assert x.shape == (self.nOutLocal, self.hop)
        y = np.empty((self.nOutLocal, self.hop), np.float32)
        self._lib.saf_multiConv_apply(self._h, _fp(x), _fp(y))
        self._raise_if_error()
        return y


class TVConv(_Base):
    """saf_TVConv_create/apply/destroy (reference saf_utility_matrixConv.h:157-190).  H: [nIRs, nCHout, length_h]."""

    _destroy_name = "saf_TVConv_destroy"

    def __init__(self, hopSize: int, H: np.ndarray, initIdx: int = 0, device: int | None = None):
        super().__init__()
        H = np.ascontiguousarray(H, np.float32)
        self.nIRs, self.nCHout, self.length_h = H.shape
        self.nOutLocal = self.nCHout
        self.hop = int(hopSize)
        self._H = H
        if device is not None:
            self._lib.safconv_set_device(int(device))
        rows = (_f32p * self.nIRs)(*[_fp(H[i]) for i in range(self.nIRs)])
        self._lib.saf_TVConv_create(C.byref(self._h), self.hop, rows, self.length_h, self.nIRs, self.nCHout, int(initIdx))
        self._check_created("saf_TVConv_create")

    def apply(self, inputSigs: np.ndarray, irIdx: int) -> np.ndarray:
        x = np.ascontiguousarray(inputSigs, np.float32).reshape(-1)
        assert x.shape == (self.hop,)
        y = np.empty((self.nCHout, self.hop), np.float32)
        self._lib.saf_TVConv_apply(self._h, _fp(x), _fp(y), int(irIdx))
        self._raise_if_error()
        return y


def fftconv(x: np.ndarray, h: np.ndarray, filt: bool = False) -> np.ndarray:
    """fftconv / fftfilt (reference saf_utility_fft.h:86-113): x[nCH, x_len], h[nCH, h_len] ->
    y[nCH, x_len + h_len - 1] (fftconv) or y[nCH, x_len] (fftfilt, filt=True)."""
    L = lib()
    x = np.ascontiguousarray(x, np.float32)
    h = np.ascontiguousarray(h, np.float32)
    nCH, xl = x.shape
    assert h.shape[0] == nCH
    hl = h.shape[1]
    y = np.empty((nCH, xl if filt else xl + hl - 1), np.float32)
    fn = L.safconv_fftfilt if filt else L.safconv_fftconv
    fn.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    fn.restype = C.c_int
    rc = fn(_fp(x), _fp(h), xl, hl, nCH, _fp(y))
    if rc:
        L.safconv_last_error_string.restype = C.c_char_p
        raise RuntimeError("fftconv failed (%d): %s" % (rc, L.safconv_last_error_string(None).decode()))
    return y


def fftfilt(x: np.ndarray, h: np.ndarray) -> np.ndarray:
    return fftconv(x, h, filt=True)


def rfft_forward(x: np.ndarray) -> np.ndarray:
    """x[batch, N] real -> X[batch, N/2+1] complex64 (saf_rfft_forward conventions, power-of-two N)."""
    L = lib()
    x = np.ascontiguousarray(np.atleast_2d(x), np.float32)
    nb, N = x.shape
    X = np.empty((nb, N // 2 + 1, 2), np.float32)
    L.safconv_rfft_forward.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    rc = L.safconv_rfft_forward(N, nb, _fp(x), _fp(X))
    if rc:
        raise RuntimeError("rfft_forward failed (%d)" % rc)
    return X[..., 0] + 1j * X[..., 1]


def rfft_backward(X: np.ndarray) -> np.ndarray:
    """X[batch, N/2+1] complex -> x[batch, N] real, scaled by 1/N (saf_rfft_backward conventions)."""
    L = lib()
    X = np.atleast_2d(np.asarray(X, np.complex64))
    nb, nbins = X.shape
    N = 2 * (nbins - 1)
    Xi = np.ascontiguousarray(np.stack([X.real, X.imag], -1), np.float32)
    x = np.empty((nb, N), np.float32)
    L.safconv_rfft_backward.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    rc = L.safconv_rfft_backward(N, nb, _fp(Xi), _fp(x))
    if rc:
        raise RuntimeError("rfft_backward failed (%d)" % rc)
    return x


class RFFT:
    """saf_rfft_create / forward / backward / destroy (reference saf_utility_fft.h:240-276): any even N, resident plan."""

    def __init__(self, N: int, device: int | None = None):
        self._lib = lib()
        self.N = int(N)
        self._h = C.c_void_p()
        if device is not None:
            self._lib.safconv_set_device(int(device))
        self._lib.saf_rfft_create(C.byref(self._h), self.N)
        if not self._h:
            raise SafConvError("saf_rfft_create failed: " + self._lib.safconv_last_error_string(None).decode())

    def factors(self):
        buf = (C.c_int * 32)()
        n = self._lib.safconv_rfft_get_factors(self._h, buf, 32)
        return [buf[i] for i in range(n)]

    def forward(self, x: np.ndarray) -> np.ndarray:
        """x [N] or [batch, N] real -> complex64 [N/2+1] or [batch, N/2+1]"""
        x2 = np.ascontiguousarray(np.atleast_2d(x), np.float32)
        nb = x2.shape[0]
        assert x2.shape[1] == self.N
        X = np.empty((nb, self.N // 2 + 1, 2), np.float32)
        if nb == 1:
            self._lib.saf_rfft_forward(self._h, _fp(x2), X.ctypes.data_as(C.c_void_p))
            rc = self._lib.safconv_rfft_last_error(self._h)
        else:
            rc = self._lib.safconv_rfft_batch(self._h, 0, nb, _fp(x2), _fp(X))
        if rc:
            raise SafConvError(self._lib.safconv_rfft_last_error_string(self._h).decode())
        Xc = X[..., 0] + 1j * X[..., 1]
        return Xc[0] if np.ndim(x) == 1 else Xc

    def backward(self, X: np.ndarray) -> np.ndarray:
        X2 = np.atleast_2d(np.asarray(X, np.complex64))
        nb = X2.shape[0]
        assert X2.shape[1] == self.N // 2 + 1
        Xi = np.ascontiguousarray(np.stack([X2.real, X2.imag], -1), np.float32)
        x = np.empty((nb, self.N), np.float32)
        if nb == 1:
            self._lib.saf_rfft_backward(self._h, Xi.ctypes.data_as(C.c_void_p), _fp(x))
            rc = self._lib.safconv_rfft_last_error(self._h)
        else:
            rc = self._lib.safconv_rfft_batch(self._h, 1, nb, _fp(Xi), _fp(x))
        if rc:
            raise SafConvError(self._lib.safconv_rfft_last_error_string(self._h).decode())
        return x[0] if np.ndim(X) == 1 else x

    def destroy(self):
        if getattr(self, "_h", None):
            self._lib.saf_rfft_destroy(C.byref(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def convolver_rfft(x: np.ndarray, inverse: bool = False) -> np.ndarray:
    """The convolvers' own power-of-two FFT cores on their own (test entry point; N = 64..16384)."""
    L = lib()
    if not inverse:
        x = np.ascontiguousarray(np.atleast_2d(x), np.float32)
        nb, N = x.shape
        X = np.empty((nb, N // 2 + 1, 2), np.float32)
        rc = L.safconv_debug_convolver_rfft(N, nb, _fp(x), _fp(X), 0)
        if rc:
            raise RuntimeError("convolver rfft failed (%d)" % rc)
        return X[..., 0] + 1j * X[..., 1]
    X = np.atleast_2d(np.asarray(x, np.complex64))
    nb, nbins = X.shape
    N = 2 * (nbins - 1)
    Xi = np.ascontiguousarray(np.stack([X.real, X.imag], -1), np.float32)
    out = np.empty((nb, N), np.float32)
    rc = L.safconv_debug_convolver_rfft(N, nb, _fp(Xi), _fp(out), 1)
    if rc:
        raise RuntimeError("convolver rfft failed (%d)" % rc)
    return out
