"""ctypes mirror of the filter-producer entry points of libsafconv_b200.so (include/safconv_b200.h, "Filter producers"):
the reference's ``getBinauralAmbiDecoderFilters`` / ``getBinauralAmbiDecoderMtx``
(``/root/reference/framework/modules/saf_hoa/saf_hoa.h:401-471``) and ``ims_shoebox_*``
(``/root/reference/framework/modules/saf_reverb/saf_reverb.h:93-230``).  Binding only: every number comes from the CUDA
library; without it (or without a device) the calls raise ``SafConvError``.  Used by tests and ``bench.py``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import SafConvError, lib

_f32p = C.POINTER(C.c_float)
_vpp = C.POINTER(C.c_void_p)

# BINAURAL_AMBI_DECODER_METHODS, saf_hoa.h:131-171
DECODER_DEFAULT, DECODER_LS, DECODER_LSDIFFEQ, DECODER_SPR, DECODER_TA, DECODER_MAGLS = range(6)

PRODUCER_SYMBOLS = [
    "safconv_matrixConv_create_device", "safconv_register_tdesign",
    "safconv_getBinauralAmbiDecoderMtx", "safconv_getBinauralAmbiDecoderFilters",
    "getBinauralAmbiDecoderMtx", "getBinauralAmbiDecoderFilters", "safconv_binauralDecoder_create_matrixConv",
    "safconv_ims_shoebox_create", "safconv_ims_shoebox_destroy", "safconv_ims_shoebox_computeEchograms",
    "safconv_ims_shoebox_renderRIRs", "safconv_ims_shoebox_setRoomDimensions", "safconv_ims_shoebox_setWallAbsCoeffs",
    "safconv_ims_shoebox_addSource", "safconv_ims_shoebox_addReceiverSH", "safconv_ims_shoebox_updateSource",
    "safconv_ims_shoebox_updateReceiver", "safconv_ims_shoebox_removeSource", "safconv_ims_shoebox_removeReceiver",
    "ims_shoebox_create", "ims_shoebox_destroy", "ims_shoebox_computeEchograms", "ims_shoebox_renderRIRs",
    "ims_shoebox_setRoomDimensions", "ims_shoebox_setWallAbsCoeffs", "ims_shoebox_addSource", "ims_shoebox_addReceiverSH",
    "ims_shoebox_updateSource", "ims_shoebox_updateReceiver", "ims_shoebox_removeSource", "ims_shoebox_removeReceiver",
    "safconv_ims_get_rir", "safconv_ims_get_rir_device", "safconv_ims_get_num_images", "safconv_ims_create_matrixConv",
]

_bound = False


def _L():
    global _bound
    L = lib()
    if _bound:
        return L
    L.safconv_getBinauralAmbiDecoderFilters.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                                        _f32p, _f32p, C.c_int, C.c_int, _f32p]
    L.safconv_getBinauralAmbiDecoderMtx.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p,
                                                    C.c_int, C.c_int, C.c_void_p]
    L.getBinauralAmbiDecoderFilters.argtypes = L.safconv_getBinauralAmbiDecoderFilters.argtypes
    L.getBinauralAmbiDecoderFilters.restype = None
    L.safconv_binauralDecoder_create_matrixConv.argtypes = [_vpp, C.c_int, C.c_void_p, _f32p, C.c_int, C.c_int, C.c_float,
                                                            C.c_int, C.c_int, _f32p, C.c_int, C.c_int]
    L.safconv_matrixConv_create_device.argtypes = [_vpp, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.safconv_matrixConv_create_device.restype = None
    for pre in ("safconv_ims_shoebox_", "ims_shoebox_"):
        getattr(L, pre + "create").argtypes = [_vpp, _f32p, _f32p, C.c_float, C.c_int, C.c_float, C.c_float]
        getattr(L, pre + "destroy").argtypes = [_vpp]
        getattr(L, pre + "computeEchograms").argtypes = [C.c_void_p, C.c_int, C.c_float]
        getattr(L, pre + "renderRIRs").argtypes = [C.c_void_p, C.c_int]
        getattr(L, pre + "setRoomDimensions").argtypes = [C.c_void_p, _f32p]
        getattr(L, pre + "setWallAbsCoeffs").argtypes = [C.c_void_p, _f32p]
        getattr(L, pre + "addSource").argtypes = [C.c_void_p, _f32p, C.c_void_p]
        getattr(L, pre + "addReceiverSH").argtypes = [C.c_void_p, C.c_int, _f32p, C.c_void_p]
        getattr(L, pre + "updateSource").argtypes = [C.c_void_p, C.c_int, _f32p]
        getattr(L, pre + "updateReceiver").argtypes = [C.c_void_p, C.c_int, _f32p]
        getattr(L, pre + "removeSource").argtypes = [C.c_void_p, C.c_int]
        getattr(L, pre + "removeReceiver").argtypes = [C.c_void_p, C.c_int]
        for f in ("create", "destroy", "computeEchograms", "renderRIRs", "setRoomDimensions", "setWallAbsCoeffs",
                  "updateSource", "updateReceiver", "removeSource", "removeReceiver"):
            getattr(L, pre + f).restype = None
    L.safconv_ims_get_rir.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(_f32p), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.safconv_ims_get_rir_device.argtypes = L.safconv_ims_get_rir.argtypes
    L.safconv_ims_get_num_images.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.safconv_ims_create_matrixConv.argtypes = [C.c_void_p, C.c_int, C.c_int, _vpp]
    _bound = True
    return L


def _err(what):
    L = lib()
    msg = L.safconv_last_error_string(None)
    return SafConvError(f"{what}: {msg.decode() if msg else 'unknown error'}")


def _f(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(_f32p)


def decoder_filters(hrtfs, dirs_deg, fftSize, fs, method, order, itd_s=None, weights=None, diffCM=0, maxRE=0,
                    reference_name=False):
    """getBinauralAmbiDecoderFilters: hrtfs (fftSize/2+1) x 2 x nDirs complex -> 2 x (order+1)^2 x fftSize float32."""
    L = _L()
    H = np.ascontiguousarray(hrtfs, np.complex64)
    d = _f(dirs_deg)
    nB, nE, nD = H.shape
    if nE != 2 or nB != fftSize // 2 + 1 or d.shape != (nD, 2):
        raise ValueError("hrtfs must be (fftSize/2+1, 2, nDirs), dirs_deg (nDirs, 2)")
    itd, w = _f(itd_s), _f(weights)
    out = np.zeros((2, (order + 1) ** 2, fftSize), np.float32)
    if reference_name:
        L.getBinauralAmbiDecoderFilters(H.ctypes.data_as(C.c_void_p), _p(d), nD, fftSize, float(fs), int(method), int(order),
                                        _p(itd), _p(w), int(diffCM), int(maxRE), _p(out))
        if L.safconv_last_error(None):
            raise _err("getBinauralAmbiDecoderFilters")
        return out
    rc = L.safconv_getBinauralAmbiDecoderFilters(H.ctypes.data_as(C.c_void_p), _p(d), nD, fftSize, float(fs), int(method),
                                                 int(order), _p(itd), _p(w), int(diffCM), int(maxRE), _p(out))
    if rc:
        raise _err("getBinauralAmbiDecoderFilters")
    return out


def decoder_mtx(hrtfs, dirs_deg, method, order, freqs=None, itd_s=None, weights=None, diffCM=0, maxRE=0):
    """getBinauralAmbiDecoderMtx: hrtfs nBands x 2 x nDirs complex -> nBands x 2 x (order+1)^2 complex64."""
    L = _L()
    H = np.ascontiguousarray(hrtfs, np.complex64)
    d = _f(dirs_deg)
    nB, _, nD = H.shape
    fr, itd, w = _f(freqs), _f(itd_s), _f(weights)
    out = np.zeros((nB, 2, (order + 1) ** 2), np.complex64)
    rc = L.safconv_getBinauralAmbiDecoderMtx(H.ctypes.data_as(C.c_void_p), _p(d), nD, nB, int(method), int(order),
                                             _p(fr), _p(itd), _p(w), int(diffCM), int(maxRE), out.ctypes.data_as(C.c_void_p))
    if rc:
        raise _err("getBinauralAmbiDecoderMtx")
    return out


def register_tdesign(degree, dirs_deg):
    """safconv_register_tdesign: the t-design the SPR decoder of order degree / 2 projects on (dirs_deg: [nPoints, 2])"""
    L = _L()
    d = _f(dirs_deg)
    L.safconv_register_tdesign.argtypes = [C.c_int, _f32p, C.c_int]
    if L.safconv_register_tdesign(int(degree), _p(d), d.shape[0]):
        raise _err("safconv_register_tdesign")


def apply_raw(handle, x, nOut):
    """one saf_matrixConv_apply on a raw handle: x (nIn, hop) float32 -> (nOut, hop)"""
    L = lib()
    x = np.ascontiguousarray(x, np.float32)
    y = np.zeros((nOut, x.shape[1]), np.float32)
    L.saf_matrixConv_apply(handle, x.ctypes.data_as(_f32p), y.ctypes.data_as(_f32p))
    if L.safconv_last_error(handle):
        raise SafConvError(L.safconv_last_error_string(handle).decode())
    return y


def destroy_raw(handle):
    lib().saf_matrixConv_destroy(C.byref(handle))


def decoder_matrixconv(hop, hrtfs, dirs_deg, fftSize, fs, method, order, weights=None, diffCM=0, maxRE=0):
    """safconv_binauralDecoder_create_matrixConv -> raw convolver handle (ctypes void*); destroy with
    ``lib().saf_matrixConv_destroy``."""
    L = _L()
    H = np.ascontiguousarray(hrtfs, np.complex64)
    d = _f(dirs_deg)
    w = _f(weights)
    h = C.c_void_p()
    rc = L.safconv_binauralDecoder_create_matrixConv(C.byref(h), int(hop), H.ctypes.data_as(C.c_void_p), _p(d), H.shape[2],
                                                     int(fftSize), float(fs), int(method), int(order), _p(w), int(diffCM), int(maxRE))
    if rc or not h:
        raise _err("safconv_binauralDecoder_create_matrixConv")
    return h


class ImsShoebox:
    """The reference's ims_shoebox_* scene (RIR path) on the GPU."""

    def __init__(self, room, abs_wall, lowest_band, n_bands, c_ms, fs, reference_names=False):
        self.L = _L()
        self.pre = "ims_shoebox_" if reference_names else "safconv_ims_shoebox_"
        self.h = C.c_void_p()
        room = _f(room)
        aw = _f(abs_wall)
        self._call("create", C.byref(self.h), _p(room), _p(aw), float(lowest_band), int(n_bands), float(c_ms), float(fs))
        if not self.h:
            raise _err("ims_shoebox_create")

    def _call(self, name, *args):
        return getattr(self.L, self.pre + name)(*args)

    def add_source(self, xyz):
        sid = self._call("addSource", self.h, _p(_f(xyz)), None)
        if sid < 0:
            raise _err("ims_shoebox_addSource")
        return sid

    def add_receiver_sh(self, order, xyz):
        rid = self._call("addReceiverSH", self.h, int(order), _p(_f(xyz)), None)
        if rid < 0:
            raise _err("ims_shoebox_addReceiverSH")
        return rid

    def update_source(self, sid, xyz):
        self._call("updateSource", self.h, int(sid), _p(_f(xyz)))

    def update_receiver(self, rid, xyz):
        self._call("updateReceiver", self.h, int(rid), _p(_f(xyz)))

    def remove_source(self, sid):
        self._call("removeSource", self.h, int(sid))

    def remove_receiver(self, rid):
        self._call("removeReceiver", self.h, int(rid))

    def set_room(self, room):
        self._call("setRoomDimensions", self.h, _p(_f(room)))

    def set_abs(self, abs_wall):
        self._call("setWallAbsCoeffs", self.h, _p(_f(abs_wall)))

    def compute_echograms(self, maxN, maxTime_s):
        self._call("computeEchograms", self.h, int(maxN), float(maxTime_s))
        if self.L.safconv_last_error(None):
            raise _err("ims_shoebox_computeEchograms")

    def render_rirs(self, frac=0):
        self._call("renderRIRs", self.h, int(frac))
        if self.L.safconv_last_error(None):
            raise _err("ims_shoebox_renderRIRs")

    def rir(self, rid, sid):
        p = _f32p(); n = C.c_int(); ch = C.c_int()
        if self.L.safconv_ims_get_rir(self.h, int(rid), int(sid), C.byref(p), C.byref(n), C.byref(ch)):
            raise _err("safconv_ims_get_rir")
        return np.ctypeslib.as_array(p, shape=(ch.value, n.value)).copy()

    def num_images(self, rid, sid):
        return self.L.safconv_ims_get_num_images(self.h, int(rid), int(sid))

    def matrixconv(self, rid, hop):
        """safconv_ims_create_matrixConv -> raw convolver handle (ctypes void*)."""
        h = C.c_void_p()
        rc = self.L.safconv_ims_create_matrixConv(self.h, int(rid), int(hop), C.byref(h))
        if rc or not h:
            raise _err("safconv_ims_create_matrixConv")
        return h

    def destroy(self):
        if self.h:
            self._call("destroy", C.byref(self.h))
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
