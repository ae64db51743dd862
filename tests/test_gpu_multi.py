"""GPU tests of the multi-GPU handle (csrc/safconv_multi.c) and of several live handles on one device.

safconv_matrixConv_create_multi / safconv_multiConv_create_multi build ONE handle that the unchanged drop-in calls
saf_matrixConv_apply / saf_multiConv_apply drive over several GPUs (output channels sharded, reference
saf_utility_matrixConv.c:218-234).  Checked against the oracle with both transports ("host": every device reads /
writes the page-locked block directly; "nccl": broadcast in, send/recv gather out).  With one visible GPU the same
code still runs with a single worker (and a one-rank NCCL communicator); the G > 1 cases skip.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import TOL_MAXABS_FS, TOL_REL_L2, err_metrics

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def check(y, ref, what=""):
    ma, l2 = err_metrics(y, ref)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, f"{what}: max-abs/fs {ma:.3g}, rel-L2 {l2:.3g}"


def devices_for(g):
    if g > n_gpus():
        pytest.skip(f"needs {g} GPUs")
    return list(range(g))


# hop, L, nIn, nOut, blocks
MULTI_GPU_MATRIX = [
    (1024, 12000, 16, 8, 14),     # C4-like: look-ahead apply (P = 12) on every shard
    (256, 700, 5, 9, 20),         # ragged shards (9 outputs), P = 3
    (256, 1024, 4, 2, 12),        # C1: shards take the fused small-problem kernel
    (2048, 512, 32, 40, 4),       # unit-test shape: one partition, no look-ahead
    (2048, 300000, 8, 4, 5),      # long filters (P = 147), few channels
    (8192, 20000, 40, 6, 3),      # input block 1.3 MB: copy-engine path instead of zero-copy
]


@pytest.mark.parametrize("transport", ["host", "nccl"])
@pytest.mark.parametrize("g", [1, 2, 3])
@pytest.mark.parametrize("hop,L,nIn,nOut,nblk", MULTI_GPU_MATRIX)
def test_multi_gpu_matrixconv_vs_oracle(saf, orc, hop, L, nIn, nOut, nblk, g, transport):
    devs = devices_for(g)
    rng = np.random.default_rng(hop + L + nIn + 7 * g)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H, 1, devices=devs)
    assert mc.multi_devices() == devs[:min(g, nOut)]
    if transport == "nccl":
        mc.set_option("transport", 1)
    y = mc.run(x)                      # the unchanged saf_matrixConv_apply, pageable numpy buffers
    check(y, ref, f"multi-GPU matrixConv G={g} {transport}")
    info = mc.info()
    assert (info.nCHout, info.nOutLocal, info.nCHin, info.hopSize) == (nOut, nOut, nIn, hop)
    assert sum(mc.shard_info(i).nOutLocal for i in range(len(mc.multi_devices()))) == nOut
    # reset: the same input gives the same output again (to rounding: the look-ahead apply picks its regime by timing)
    mc.reset_state()
    check(mc.run(x), ref, f"multi-GPU matrixConv G={g} {transport}, after reset")
    mc.destroy()


@pytest.mark.parametrize("transport", ["host", "nccl"])
@pytest.mark.parametrize("g", [1, 2])
def test_multi_gpu_multiconv_vs_oracle(saf, orc, g, transport):
    devs = devices_for(g)
    hop, L, nCH, nblk = 512, 4096, 37, 8
    rng = np.random.default_rng(99 + g)
    H = rng.uniform(-1, 1, (nCH, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nCH, hop * nblk)).astype(np.float32)
    ref = orc.OracleMultiConv(hop, H, 1).run(x)
    mc = saf.MultiConv(hop, H, 1, devices=devs)
    if transport == "nccl":
        mc.set_option("transport", 1)
    check(mc.run(x), ref, f"multi-GPU multiConv G={g} {transport}")
    mc.destroy()


@pytest.mark.parametrize("g", [1, 2])
def test_multi_gpu_pinned_buffers_transport_switch_and_sleeping_workers(saf, orc, g):
    """Page-locked caller buffers are used directly; the transport can change between blocks (same per-device
    state); workers that went to sleep (spin time 0) wake up for the next block."""
    import time
    import torch
    devs = devices_for(g)
    hop, L, nIn, nOut, nblk = 512, 5000, 6, 6, 18
    rng = np.random.default_rng(5)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H, 1, devices=devs)
    mc.set_option("worker_spin_us", 0)
    xin = torch.empty((nIn, hop), dtype=torch.float32).pin_memory()
    yout = torch.empty((nOut, hop), dtype=torch.float32).pin_memory()
    fp = C.POINTER(C.c_float)
    y = np.empty_like(ref)
    for b in range(nblk):
        if b == 6:
            mc.set_option("transport", 1)
        if b == 12:
            mc.set_option("transport", 0)
        xin.copy_(torch.from_numpy(x[:, b * hop:(b + 1) * hop]))
        saf.lib().saf_matrixConv_apply(mc.handle, C.cast(xin.data_ptr(), fp), C.cast(yout.data_ptr(), fp))
        assert saf.lib().safconv_last_error(mc.handle) == 0, saf.lib().safconv_last_error_string(mc.handle)
        y[:, b * hop:(b + 1) * hop] = yout.numpy()
        if b % 5 == 0:
            time.sleep(0.01)
    check(y, ref, f"multi-GPU pinned G={g}")
    mc.destroy()


def test_multi_gpu_equals_single_gpu_bitwise(saf):
    """Sharding does not change a single bit when the shard geometry matches: compare against single-device SHARD
    handles of the same ranges (look-ahead off on both sides: it picks its regime -- and summation order -- by timing)."""
    devs = devices_for(2)
    hop, L, nIn, nOut, nblk = 1024, 9000, 8, 6, 10
    rng = np.random.default_rng(11)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    mc = saf.MatrixConv(hop, H, 1, devices=devs)
    mc.set_option("lookahead", 0)
    y = mc.run(x)
    mc.destroy()
    parts = []
    for ob in (0, 3):
        sh = saf.MatrixConv(hop, H, 1, shard=(ob, 3), device=0)
        sh.set_option("lookahead", 0)
        parts.append(sh.run(x))
        sh.destroy()
    assert np.array_equal(y, np.concatenate(parts, 0))


def test_multi_gpu_env_makes_plain_create_multi(saf, orc, monkeypatch):
    """SAFCONV_DEVICES: a SAF host gets a multi-GPU handle from the unchanged saf_matrixConv_create."""
    devices_for(2)
    monkeypatch.setenv("SAFCONV_DEVICES", "0,1")
    hop, L, nIn, nOut, nblk = 256, 2000, 3, 4, 9
    rng = np.random.default_rng(3)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    mc = saf.MatrixConv(hop, H, 1)
    assert mc.multi_devices() == [0, 1]
    check(mc.run(x), orc.OracleMatrixConv(hop, H, 1).run(x), "SAFCONV_DEVICES")
    mc.destroy()


def test_multi_gpu_create_errors(saf):
    H = np.zeros((2, 2, 8), np.float32)
    with pytest.raises(saf.SafConvError):
        saf.MatrixConv(64, H, 1, devices=[0, 0])            # repeated device
    with pytest.raises(saf.SafConvError):
        saf.MatrixConv(64, H, 1, devices=[n_gpus()])        # no such device
    mc = saf.MatrixConv(64, H, 1, devices=[0])
    with pytest.raises(saf.SafConvError):
        mc.set_option("no_such_option", 1)
    mc.destroy()


def test_two_live_handles_with_different_plans(saf, orc):
    """Kernel attributes are global per (function, device): handles that share a kernel instantiation but need
    different amounts of dynamic shared memory must not break each other (the smaller handle used to
    lower the limit and every later launch of the larger one failed)."""
    rng = np.random.default_rng(21)
    cases = [(1024, 9000, 64, 64),     # R = 8, 150 KB MAC pipeline
             (1024, 9000, 1, 64),      # R = 8, much smaller stages
             (8192, 9000, 2, 3),       # FFT work arrays > 48 KB
             (4096, 9000, 2, 3)]
    convs = []
    for hop, L, nIn, nOut in cases:
        H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
        x = rng.uniform(-1, 1, (nIn, hop * 3)).astype(np.float32)
        convs.append((saf.MatrixConv(hop, H, 1), orc.OracleMatrixConv(hop, H, 1).run(x), x))
    for _ in range(2):                                   # A (large) is applied AFTER B (small) was created and applied
        for mc, ref, x in convs:
            mc.reset_state()
            check(mc.run(x), ref, f"hop {mc.hop}")
    Hm = rng.uniform(-1, 1, (4, 3000)).astype(np.float32)
    xm = rng.uniform(-1, 1, (4, 8192 * 2)).astype(np.float32)
    big, small = saf.MultiConv(8192, Hm, 1), saf.MultiConv(256, Hm, 1)
    check(small.run(xm), orc.OracleMultiConv(256, Hm, 1).run(xm), "multi small")
    check(big.run(xm), orc.OracleMultiConv(8192, Hm, 1).run(xm), "multi big")
    for mc, _, _ in convs:
        mc.destroy()


def test_error_is_not_sticky(saf):
    """A failed call does not poison later, correct calls."""
    rng = np.random.default_rng(2)
    H = rng.uniform(-1, 1, (3, 2, 300)).astype(np.float32)
    tv = saf.TVConv(128, H, 0)
    x = rng.uniform(-1, 1, 128).astype(np.float32)
    with pytest.raises(saf.SafConvError):
        tv.apply(x, 99)
    tv.apply(x, 1)                                        # raises if the error were sticky
    tv.destroy()


def test_many_rows_in_filter_transform(saf, orc):
    """More than 65535 (IR x output) rows: the create-time transform indexes its CTAs linearly."""
    hop, L, nIRs, nOut = 64, 64, 2200, 32                 # 70400 rows
    rng = np.random.default_rng(8)
    H = rng.uniform(-1, 1, (nIRs, nOut, L)).astype(np.float32)
    tv, ref = saf.TVConv(hop, H, 0), orc.OracleTVConv(hop, H, 0)
    ys, rs = [], []
    for ir in (0, nIRs - 1, nIRs - 1, 1100, 3):
        x = rng.uniform(-1, 1, hop).astype(np.float32)
        ys.append(tv.apply(x, ir)); rs.append(ref.apply(x, ir))
    check(np.concatenate(ys, 1), np.concatenate(rs, 1), "TVConv, 70400 rows")
    tv.destroy()
