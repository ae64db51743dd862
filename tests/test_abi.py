"""The C-ABI library loads without a GPU and exports every symbol include/safconv_b200.h declares."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    txt = (ROOT / "include" / "safconv_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b((?:saf|safconv|ims_shoebox)_\w+|fftconv|fftfilt|getBinauralAmbiDecoder\w+)\s*\(", txt)
    return sorted(set(names))


def test_header_symbols_all_exported(saf):
    lib = saf.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/safconv_b200.h but not exported"
    # and the binding's own list agrees with the header
    from spatial_audio_framework_b200._capi import EXPORTED_SYMBOLS
    from spatial_audio_framework_b200.producers import PRODUCER_SYMBOLS
    assert sorted(EXPORTED_SYMBOLS + PRODUCER_SYMBOLS) == syms


def test_reference_signatures_are_drop_in(saf):
    """Same six (+3 TVConv) names as the reference header saf_utility_matrixConv.h:55-190."""
    for s in ["saf_matrixConv_create", "saf_matrixConv_destroy", "saf_matrixConv_apply",
              "saf_multiConv_create", "saf_multiConv_destroy", "saf_multiConv_apply",
              "saf_TVConv_create", "saf_TVConv_destroy", "saf_TVConv_apply"]:
        assert s in declared_symbols()


def test_version_and_null_safety(saf):
    lib = saf.lib()
    assert b"sm_100a" in lib.safconv_version()
    h = C.c_void_p()
    lib.saf_matrixConv_destroy(C.byref(h))      # destroy(NULL handle) is a no-op (reference .c:140)
    lib.saf_multiConv_destroy(C.byref(h))
    lib.saf_TVConv_destroy(C.byref(h))
    x = np.zeros(8, np.float32)
    p = x.ctypes.data_as(C.POINTER(C.c_float))
    lib.saf_matrixConv_apply(None, p, p)        # apply on NULL handle: no-op, no crash
    lib.saf_multiConv_apply(None, p, p)


def test_invalid_arguments_leave_null_handle(saf):
    lib = saf.lib()
    H = np.zeros((2, 2, 16), np.float32)
    hp = H.ctypes.data_as(C.POINTER(C.c_float))
    for args in [(0, hp, 16, 2, 2, 1), (64, hp, 0, 2, 2, 1), (64, hp, 16, 0, 2, 1), (64, hp, 16, 2, 0, 1),
                 (64, None, 16, 2, 2, 1), (16384, None, 16, 2, 2, 1),
                 (16385, hp, 1, 2, 2, 1)]:      # hop > 8192 goes to the big-FFT engine, which needs an even numOvrlpAddBlocks * hop (here 1 x 16385)
        h = C.c_void_p(123)
        lib.saf_matrixConv_create(C.byref(h), *args)
        assert not h.value
        assert lib.safconv_last_error(None) == 1
        assert b"invalid" in lib.safconv_last_error_string(None)


def test_no_cpu_fallback_without_device(saf):
    """Without a CUDA device create() must fail loudly (NULL handle + error), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(saf.SafConvError, match="no usable CUDA device"):
        saf.MatrixConv(64, np.zeros((1, 1, 64), np.float32))
    with pytest.raises(saf.SafConvError, match="no usable CUDA device"):
        saf.MultiConv(64, np.zeros((1, 64), np.float32))


def test_multi_gpu_create_argument_checks_and_no_device(saf, monkeypatch):
    """safconv_matrixConv_create_multi: bad arguments are rejected before any device work; without a CUDA device it
    fails loudly like the single-device create; the extension calls reject a NULL / foreign handle."""
    import torch
    lib = saf.lib()
    H = np.zeros((2, 2, 16), np.float32)
    hp = H.ctypes.data_as(C.POINTER(C.c_float))
    devs = (C.c_int * 2)(0, 1)
    for args in [(64, None, 16, 2, 2, devs, 2), (64, hp, 16, 2, 2, None, 2), (64, hp, 16, 2, 2, devs, 0),
                 (64, hp, 16, 2, 2, devs, 17), (0, hp, 16, 2, 2, devs, 2), (64, hp, 0, 2, 2, devs, 2)]:
        h = C.c_void_p(123)
        lib.safconv_matrixConv_create_multi(C.byref(h), *args)
        assert not h.value
        assert lib.safconv_last_error(None) == 1 and b"invalid" in lib.safconv_last_error_string(None)
    assert lib.safconv_multi_get_devices(None, devs, 2) == 0
    assert not lib.safconv_multi_get_shard(None, 0)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(saf.SafConvError, match="no usable CUDA device"):
        saf.MatrixConv(64, H, devices=[0, 1])
    with pytest.raises(saf.SafConvError, match="no usable CUDA device"):
        saf.MultiConv(64, H[0], devices=[0])
    monkeypatch.setenv("SAFCONV_DEVICES", "all")           # the env route falls through to the ordinary failure
    with pytest.raises(saf.SafConvError, match="no usable CUDA device"):
        saf.MatrixConv(64, H)


def test_helpers_fail_loudly_without_device(saf):
    """fftconv / fftfilt / rfft have no CPU path either: error code + message, output untouched; bad arguments are
    rejected before any device work."""
    import torch
    import spatial_audio_framework_b200 as pkg
    lib = saf.lib()
    x = np.ones((2, 100), np.float32); h = np.ones((2, 9), np.float32)
    fp = C.POINTER(C.c_float)
    lib.safconv_fftconv.restype = C.c_int
    lib.safconv_rfft_forward.restype = C.c_int
    lib.safconv_last_error_string.restype = C.c_char_p
    y = np.full((2, 108), 7.0, np.float32)
    # invalid arguments
    assert lib.safconv_fftconv(x.ctypes.data_as(fp), h.ctypes.data_as(fp), 0, 9, 2, y.ctypes.data_as(fp)) == 1
    X = np.zeros((1, 49, 2), np.float32)
    assert lib.safconv_rfft_forward(97, 1, x.ctypes.data_as(fp), X.ctypes.data_as(fp)) == 1      # odd sizes are illegal (reference .c:542)
    hf = C.c_void_p(9)
    lib.saf_rfft_create(C.byref(hf), 97)
    assert not hf.value and b"even" in lib.safconv_last_error_string(None)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc = lib.safconv_fftconv(x.ctypes.data_as(fp), h.ctypes.data_as(fp), 100, 9, 2, y.ctypes.data_as(fp))
    assert rc != 0 and b"CUDA device" in lib.safconv_last_error_string(None)
    assert np.all(y == 7.0)
    with pytest.raises(RuntimeError):
        pkg.fftconv(x, h)
    with pytest.raises(RuntimeError):
        pkg.rfft_forward(np.zeros((1, 64), np.float32))
    with pytest.raises(saf.SafConvError, match="no usable CUDA device"):
        pkg.RFFT(1280)


def test_product_does_not_reference_oracle():
    """The product sources must not import/link anything under oracle/."""
    pkg = ROOT / "spatial_audio_framework_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.c")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")) + list(pkg.rglob("Makefile")):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, p


def test_header_is_strict_c99_and_cxx17_and_the_c_example_builds(tmp_path):
    """include/safconv_b200.h under -std=c99 / -std=c++17 -Wall -Wextra -pedantic -Werror, and the C example of the
    producer -> convolver flow (examples/producers_to_convolver.c) built against the library.  Without a CUDA device
    the example must stop at its first create with the library's error string (exit code 2), not crash."""
    import shutil
    import subprocess
    import torch
    if not shutil.which("gcc") or not shutil.which("g++"):
        pytest.skip("no host compiler")
    lib_dir = ROOT / "spatial_audio_framework_b200"
    link = ["-L" + str(lib_dir), "-lsafconv_b200", f"-Wl,-rpath,{lib_dir}", "-lm"]
    exe = tmp_path / "example"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + str(ROOT / "include"),
                    str(ROOT / "examples" / "producers_to_convolver.c"), "-o", str(exe)] + link, check=True)
    cpp = tmp_path / "h.cpp"
    cpp.write_text('#include "safconv_b200.h"\nint main() { return safconv_version() ? 0 : 1; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + str(ROOT / "include"),
                    str(cpp), "-o", str(tmp_path / "hcpp")] + link, check=True)
    assert subprocess.run([str(tmp_path / "hcpp")]).returncode == 0
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "energy at the ears" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 2 and "no usable CUDA device" in r.stderr
