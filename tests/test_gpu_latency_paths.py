"""GPU parity of the round-2 latency paths, through the C ABI against the oracle:

* the thread-block-cluster kernel that serves small matrix problems in one launch (csrc/safconv_kernels.cu,
  small_cluster_kernel<R>): every R = M/32, odd / non-power-of-two hops, more inputs than one round of warps, ragged
  output ownership, with and without the host-visible completion word;
* the look-ahead apply's knobs: depth 2 (two tail passes queued, two-partition head pass), K3 adding the newest
  partitions itself, tail passes that leave 0 / 64 SMs free -- all must give the oracle's result (and, for a fixed
  setting, the same bits on every run).

Tolerance (north_star): max abs error <= 1e-5 of full scale AND relative L2 <= 1e-6 vs the reference.
"""
import ctypes as C
import time

import numpy as np
import pytest

from conftest import TOL_MAXABS_FS, TOL_REL_L2, err_metrics

pytestmark = pytest.mark.gpu


def check(y, ref, what=""):
    ma, l2 = err_metrics(y, ref)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, f"{what}: max-abs/fs {ma:.3g}, rel-L2 {l2:.3g}"
    return ma, l2


# hop, L, nIn, nOut, blocks                      M = nextpow2(max(32, hop))
CLUSTER_CASES = [
    (32, 100, 3, 2, 30),        # M = 64  (R = 2)
    (50, 333, 5, 3, 17),        # M = 64, non-power-of-two hop
    (128, 512, 25, 2, 12),      # M = 128 (R = 4): configs[1]
    (101, 400, 7, 5, 11),       # M = 128, odd hop (scalar input loads)
    (256, 1024, 4, 2, 12),      # M = 256 (R = 8): configs[0]
    (200, 900, 40, 6, 9),       # M = 256, 40 inputs over 8 CTAs, 6 outputs
    (512, 2000, 9, 17, 7),      # M = 512 (R = 16), 17 outputs: ragged ownership (3 / 2 per CTA)
    (1024, 3000, 6, 4, 6),      # M = 1024 (R = 32)
    (700, 700, 2, 1, 8),        # M = 1024, one partition, one output
    (64, 64, 1, 1, 9),          # a single unit of work: cluster of 2 CTAs
    (128, 300, 128, 2, 5),      # 128 inputs = the kernel's maximum (16 warps x 8 CTAs)
]


@pytest.mark.parametrize("flag_wait", [1, 0])
@pytest.mark.parametrize("hop,L,nIn,nOut,nblk", CLUSTER_CASES)
def test_small_cluster_kernel_vs_oracle(saf, orc, hop, L, nIn, nOut, nblk, flag_wait):
    rng = np.random.default_rng(hop * 7 + L + nIn + nOut)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H, 1)
    mc.set_option("flag_wait", flag_wait)
    y = mc.run(x)
    check(y, ref, "one-launch latency kernel")
    # state handling: reset and run again -> the same bits; then the three-kernel path on the same handle continues the
    # SAME delay line / overlap tails (both write the shared ring layout)
    mc.reset_state()
    assert np.array_equal(mc.run(x), y)
    mc.reset_state()
    half = (nblk // 2) * hop
    y1 = mc.run(x[:, :half])
    mc.set_option("small_fused", 0)
    y2 = mc.run(x[:, half:])
    check(np.concatenate([y1, y2], 1), ref, "latency kernel, then K1 -> K2 -> K3 on the same state")
    mc.destroy()


@pytest.mark.parametrize("depth,head_in_k3,reserve", [(1, 0, 32), (2, 0, 32), (2, 1, 32), (1, 1, 0), (1, 0, 64), (2, 0, 0)])
@pytest.mark.parametrize("hop,L,nIn,nOut,nblk", [(1024, 12000, 16, 8, 14), (512, 9000, 5, 70, 12), (256, 700, 3, 2, 9)])
def test_lookahead_knobs_vs_oracle(saf, orc, monkeypatch, hop, L, nIn, nOut, nblk, depth, head_in_k3, reserve):
    """Back-to-back calls (throughput regime) and paced calls (latency regime: the GPU is idle when the block arrives)
    for every look-ahead setting; P = 3 with depth 2 leaves a one-partition tail pass."""
    monkeypatch.setenv("SAFCONV_LA_DEPTH", str(depth))
    monkeypatch.setenv("SAFCONV_HEAD_IN_K3", str(head_in_k3))
    monkeypatch.setenv("SAFCONV_TAIL_RESERVE_SMS", str(reserve))
    rng = np.random.default_rng(hop + L + depth)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H, 1)
    mc.set_option("small_fused", 0)                      # force the look-ahead sequence for the small shape too
    y = mc.run(x)
    check(y, ref, "look-ahead, back to back")
    mc.reset_state()
    yp = np.empty_like(y)
    for b in range(nblk):
        yp[:, b * hop:(b + 1) * hop] = mc.apply(np.ascontiguousarray(x[:, b * hop:(b + 1) * hop]))
        time.sleep(0.003)                                # every pending tail pass has finished: latency regime
    check(yp, ref, "look-ahead, paced")
    mc.destroy()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("hop,L,nIn,nOut,nblk", [(128, 512, 25, 2, 14), (256, 1024, 4, 2, 12), (50, 333, 5, 3, 17),
                                                   (512, 2000, 9, 17, 9), (1024, 3000, 6, 4, 6)])
def test_resident_latency_kernel(saf, orc, hop, L, nIn, nOut, nblk):
    """Option "resident_us": the cluster kernel stays on its SMs and serves one block per doorbell (no launch per call).
    Same results as the one-launch path bit for bit (same code, other load instructions); it leaves on its idle timer and
    comes back with the next block; every other call on the handle stops it first; destroy while it polls is clean."""
    import torch
    rng = np.random.default_rng(hop + nIn * 3 + nOut)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    plain = saf.MatrixConv(hop, H, 1)
    y0 = plain.run(x)
    plain.destroy()
    mc = saf.MatrixConv(hop, H, 1)
    mc.set_option("resident_us", 3000)                   # 3 ms idle timer
    y = mc.run(x)                                        # pageable numpy blocks: staged through the handle's pinned buffers
    check(y, ref, "resident kernel")
    assert np.array_equal(y, y0)
    # idle longer than the timer between blocks: the kernel has left, the next call starts a new one
    mc.reset_state()                                     # (stops the resident kernel, zeroes the state)
    yp = np.empty_like(y)
    for b in range(nblk):
        yp[:, b * hop:(b + 1) * hop] = mc.apply(np.ascontiguousarray(x[:, b * hop:(b + 1) * hop]))
        if b % 3 == 0:
            time.sleep(0.02)
    assert np.array_equal(yp, y0)
    # page-locked caller buffers, other calls in between (device-pointer block, option change)
    mc.reset_state()
    n1 = nblk // 2
    fp = C.POINTER(C.c_float)
    parts = []
    for b in range(n1):
        xb = torch.from_numpy(np.ascontiguousarray(x[:, b * hop:(b + 1) * hop])).pin_memory()
        yb = torch.empty((nOut, hop)).pin_memory()
        saf.lib().saf_matrixConv_apply(mc.handle, C.cast(xb.data_ptr(), fp), C.cast(yb.data_ptr(), fp))     # zero-copy: the kernel reads / writes these
        assert saf.lib().safconv_last_error(mc.handle) == 0
        parts.append(yb.numpy().copy())
    d_in = torch.from_numpy(np.ascontiguousarray(x[:, n1 * hop:(n1 + 1) * hop])).cuda()
    d_out = torch.empty((nOut, hop), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    mc.apply_device(d_in.data_ptr(), d_out.data_ptr(), 1)     # stops the resident kernel, runs on the stream
    mc.synchronize()
    parts.append(d_out.cpu().numpy())
    mc.set_option("flag_wait", 0)
    for b in range(n1 + 1, nblk):
        parts.append(mc.apply(np.ascontiguousarray(x[:, b * hop:(b + 1) * hop])))
    check(np.concatenate(parts, 1), ref, "resident / device-pointer / resident")
    mc.destroy()                                         # while the kernel is polling
