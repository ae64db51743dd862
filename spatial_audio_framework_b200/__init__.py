"""spatial_audio_framework_b200 -- host-side mirror of the reference convolver API.

The product is ``libsafconv_b200.so`` (C host layer + hand-written sm_100a CUDA kernels,
``csrc/``), whose exported symbols are a drop-in for the reference's
``saf_utility_matrixConv.h`` (``/root/reference/framework/modules/saf_utilities/
saf_utility_matrixConv.h:55-190``).  This package only *binds* that C ABI with
ctypes so that tests and ``bench.py`` can call it; it contains no compute path of
its own and there is no CPU fallback: importing works anywhere, but creating a
convolver without the built library or without a CUDA device raises.
"""
from ._capi import (  # noqa: F401
    LIB_PATH,
    SafConvError,
    MatrixConv,
    MultiConv,
    RFFT,
    convolver_rfft,
    TVConv,
    build,
    fftconv,
    fftfilt,
    lib,
    rfft_backward,
    rfft_forward,
    version,
)
from . import synth  # noqa: F401
from . import sharding  # noqa: F401
from . import producers  # noqa: F401

__all__ = ["LIB_PATH", "SafConvError", "MatrixConv", "MultiConv", "TVConv", "RFFT", "convolver_rfft", "build", "fftconv", "fftfilt", "rfft_forward", "rfft_backward", "lib", "version", "synth", "sharding", "producers"]
