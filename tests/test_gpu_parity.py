"""GPU parity tests proper: the CUDA path (through the C ABI, host pointers in / host pointers out)
against the oracle (bit-exact restatement of the reference), the committed golden fixtures from the
compiled reference, and the fp64 direct-convolution truth.

Tolerance (north_star): max abs error <= 1e-5 of full scale AND relative L2 <= 1e-6 vs the reference.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import TOL_MAXABS_FS, TOL_REL_L2, err_metrics, golden_files

pytestmark = pytest.mark.gpu


def check(y, ref, what=""):
    ma, l2 = err_metrics(y, ref)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, f"{what}: max-abs/fs {ma:.3g}, rel-L2 {l2:.3g}"
    return ma, l2


# hop, L, nIn, nOut, usePart, blocks
MATRIX_CASES = [
    (256, 1024, 4, 2, 1, 12),      # C1 (BASELINE.json configs[0])
    (256, 1024, 4, 2, 0, 12),      # C1, reference non-partitioned mode
    (128, 512, 25, 2, 1, 12),      # C2 (configs[1])
    (2048, 512, 32, 40, 1, 4),     # the real test__saf_matrixConv shape (test__utilities_module.c:337-352)
    (64, 64, 1, 1, 1, 6),          # smallest sensible
    (64, 1, 2, 2, 1, 5),           # one-tap filters
    (32, 1000, 2, 2, 1, 40),       # many partitions (P = 32), tiny blocks
    (96, 250, 3, 5, 1, 9),         # non-power-of-two hop, L not a multiple of hop
    (100, 333, 2, 3, 0, 8),        # reference non-pow2 FFT (500) in mode 0
    (17, 40, 3, 3, 1, 11),         # odd hop
    (1, 5, 2, 2, 1, 20),           # hop 1
    (512, 2000, 7, 65, 1, 6),      # > 64 outputs: two output tiles, ragged nIn
    (1024, 5000, 9, 16, 1, 8),     # R = 2 path
    (1024, 3000, 5, 24, 1, 6),     # R = 4 path (OTsz 24 -> WGo 6)
    (4096, 9000, 3, 2, 1, 4),      # large hop, split rows across warps (WGk > 1)
    (8192, 8192, 2, 1, 1, 3),      # maximum hop
]


@pytest.mark.parametrize("path", ["default", "three_kernels"])
@pytest.mark.parametrize("hop,L,nIn,nOut,part,nblk", MATRIX_CASES)
def test_matrixconv_vs_oracle(saf, orc, hop, L, nIn, nOut, part, nblk, path):
    """path "default": small problems take the one-launch fused latency kernel, the rest K1 -> K2 -> K3;
    "three_kernels": the fused kernel is switched off so every shape also runs through K1 -> K2 -> K3."""
    rng = np.random.default_rng(hop * 31 + L + nIn)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, part).run(x)
    mc = saf.MatrixConv(hop, H, part)
    if path == "three_kernels":
        mc.set_option("small_fused", 0)
    y = mc.run(x)
    check(y, ref, "matrixConv")
    mc.destroy()


MULTI_CASES = [
    (512, 4096, 256, 1, 10),       # C3 (configs[2])
    (256, 1024, 32, 0, 6),         # test__examples.c:323 shape (non-partitioned)
    (128, 1000, 4, 1, 12),
    (50, 77, 3, 1, 9),
    (64, 64, 1, 1, 4),
    (2048, 100000, 2, 1, 6),       # long filters, few channels (P = 49)
    (8192, 20000, 3, 1, 3),
]


@pytest.mark.parametrize("hop,L,nCH,part,nblk", MULTI_CASES)
def test_multiconv_vs_oracle(saf, orc, hop, L, nCH, part, nblk):
    rng = np.random.default_rng(hop * 13 + L + nCH)
    H = rng.uniform(-1, 1, (nCH, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nCH, hop * nblk)).astype(np.float32)
    ref = orc.OracleMultiConv(hop, H, part).run(x)
    mc = saf.MultiConv(hop, H, part)
    y = mc.run(x)
    check(y, ref, "multiConv")
    mc.destroy()


@pytest.mark.parametrize("hop,L,nIRs,nOut", [(128, 700, 4, 3), (512, 3000, 3, 8), (100, 64, 2, 1),
                                             (4096, 9000, 3, 2),                       # largest hop of the one-launch kernel
                                             (8192, 20000, 3, 2), (5000, 12000, 2, 3)])  # hop > 4096: three-launch path
def test_tvconv_vs_oracle(saf, orc, hop, L, nIRs, nOut):
    rng = np.random.default_rng(hop + L)
    H = rng.uniform(-1, 1, (nIRs, nOut, L)).astype(np.float32)
    seq = [1, 1, 0, 0, 0, nIRs - 1, 1, 1, 0, 1, 0, nIRs - 1, nIRs - 1, nIRs - 1]
    a, b = orc.OracleTVConv(hop, H, 1), saf.TVConv(hop, H, 1)
    ys, rs = [], []
    for ir in seq:
        x = rng.uniform(-1, 1, hop).astype(np.float32)
        rs.append(a.apply(x, ir))
        ys.append(b.apply(x, ir))
    check(np.concatenate(ys, 1), np.concatenate(rs, 1), "TVConv")


@pytest.mark.parametrize("path", golden_files("matrix") + golden_files("multi") + golden_files("tv"), ids=lambda p: p.stem)
def test_against_reference_golden(saf, path):
    """Fixtures come from the compiled reference itself (tests/golden/make_golden.py)."""
    g = np.load(path)
    kind, hop = str(g["kind"]), int(g["hop"])
    if kind == "matrix":
        y = saf.MatrixConv(hop, g["H"], int(g["part"])).run(g["x"])
    elif kind == "multi":
        y = saf.MultiConv(hop, g["H"], int(g["part"])).run(g["x"])
    else:
        tv = saf.TVConv(hop, g["H"], int(g["initIdx"]))
        y = np.concatenate([tv.apply(g["x"][0, i * hop:(i + 1) * hop], int(ir)) for i, ir in enumerate(g["seq"])], axis=1)
    check(y, g["y"], path.stem)


@pytest.mark.parametrize("path", golden_files("fftconv"), ids=lambda p: p.stem)
def test_fftconv_against_reference_golden(saf, path):
    """fftconv / fftfilt fixtures produced by the compiled reference (tests/golden/make_golden.py fftconv)."""
    import spatial_audio_framework_b200 as pkg
    g = np.load(path)
    check(pkg.fftconv(g["x"], g["h"]), g["y"], path.stem)
    check(pkg.fftfilt(g["x"], g["h"]), g["yfilt"], path.stem + " (fftfilt)")


def test_closer_to_truth_than_needed(saf, orc):
    """Three-way distances on a mid-size case: |gpu-ref|, |gpu-truth64|, |ref-truth64|."""
    rng = np.random.default_rng(99)
    hop, L, nIn, nOut, nblk = 256, 6000, 16, 4, 30
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    y = saf.MatrixConv(hop, H, 1).run(x)
    t = orc.truth_matrix(H, x, np.arange(nOut), 0, hop * nblk)
    _, l2_gr = err_metrics(y, ref)
    _, l2_gt = err_metrics(y, t)
    _, l2_rt = err_metrics(ref, t)
    print(f"rel-L2 gpu-ref {l2_gr:.3g}  gpu-truth {l2_gt:.3g}  ref-truth {l2_rt:.3g}")
    assert l2_gr <= TOL_REL_L2
    assert l2_gt <= 5e-7


@pytest.mark.parametrize("hop,L,nIn,nOut,nblk", [(256, 1024, 4, 2, 50), (128, 512, 25, 2, 60), (512, 2000, 2, 2, 12),
                                                  (100, 900, 3, 4, 70), (1024, 1024, 8, 8, 5), (64, 640, 16, 16, 45)])
def test_fused_small_problem_kernel(saf, orc, hop, L, nIn, nOut, nblk):
    """The one-launch latency path (small_fused_kernel on mapped host buffers): C1 / C2 and neighbours, more
    blocks than ring slots (wrap-around), thread-group splits G > 1 and G = 1, interleaved with the device API."""
    import torch
    rng = np.random.default_rng(hop + nIn)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H)
    half = nblk // 2
    y1 = mc.run(x[:, :half * hop])                                   # fused kernel, block by block
    xb = np.ascontiguousarray(x[:, half * hop:].reshape(nIn, nblk - half, hop).transpose(1, 0, 2))
    d_in = torch.from_numpy(xb).cuda()
    d_out = torch.empty((nblk - half, nOut, hop), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    mc.apply_device(d_in.data_ptr(), d_out.data_ptr(), nblk - half)  # same handle continues on the batched path
    mc.synchronize()
    y2 = d_out.cpu().numpy().transpose(1, 0, 2).reshape(nOut, (nblk - half) * hop)
    check(np.concatenate([y1, y2], 1), ref, "fused then batched")
    mc.reset_state()
    check(mc.run(x), ref, "fused only")


def test_state_api(saf, orc):
    """destroy/NULL semantics, reset_state, re-create = zero state (reference .c:109,113)."""
    rng = np.random.default_rng(3)
    hop, L, nIn, nOut = 128, 500, 3, 2
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * 6)).astype(np.float32)
    mc = saf.MatrixConv(hop, H)
    y1 = mc.run(x)
    mc.reset_state()
    y2 = mc.run(x)
    assert np.array_equal(y1, y2)
    lib = saf.lib()
    h = mc.handle
    lib.saf_matrixConv_destroy(C.byref(h))
    assert not h.value                               # superset of the reference: pointer is nulled
    lib.saf_matrixConv_destroy(C.byref(h))           # second destroy is a no-op


def test_device_pointer_api_and_graph(saf, orc):
    """safconv_apply_device_blocks on device buffers == host API; CUDA-graph replay == plain launches."""
    import torch
    rng = np.random.default_rng(4)
    hop, L, nIn, nOut, nblk = 256, 3000, 6, 10, 16
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H)
    # [nblk][nIn][hop] device layout
    xb = np.ascontiguousarray(x.reshape(nIn, nblk, hop).transpose(1, 0, 2))
    d_in = torch.from_numpy(xb).cuda()
    d_out = torch.empty((nblk, nOut, hop), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    mc.apply_device(d_in.data_ptr(), d_out.data_ptr(), nblk)
    mc.synchronize()
    y = d_out.cpu().numpy().transpose(1, 0, 2).reshape(nOut, nblk * hop)
    check(y, ref, "device API")
    mc.reset_state()
    mc.set_option("use_graph", 1)                    # explicit graph replay of K1 -> K2 -> K3 (takes precedence over the fused kernel)
    yg = mc.run(x)
    assert np.array_equal(yg, y)
    # kernel timing hooks
    mc.set_option("use_graph", 0)
    mc.enable_kernel_timing(8)
    for b in range(3):
        mc.apply(np.ascontiguousarray(x[:, b * hop:(b + 1) * hop]))
    ms, n = mc.kernel_times_ms()
    assert n == 3 and all(m > 0 for m in ms)
    mc.enable_kernel_timing(0)

    H2 = rng.uniform(-1, 1, (5, L)).astype(np.float32)
    x2 = rng.uniform(-1, 1, (5, hop * 8)).astype(np.float32)
    ref2 = orc.OracleMultiConv(hop, H2, 1).run(x2)
    m2 = saf.MultiConv(hop, H2)
    m2.set_option("use_graph", 1)
    check(m2.run(x2), ref2, "multi graph")


@pytest.mark.parametrize("hop,L,nIn,nOut,nblk", [(256, 700, 5, 9, 75), (1024, 96000 // 8, 16, 8, 40), (64, 64, 2, 3, 33)])
def test_batched_device_blocks_equal_block_by_block(saf, orc, hop, L, nIn, nOut, nblk):
    """safconv_apply_device_blocks shares the FFT launches across a batch (ring of P + maxBatch slots,
    batched inverse FFT + overlap-add chain): bit-identical to one block at a time, across batch
    boundaries and ring wrap-around, and equal to the oracle within tolerance."""
    import torch
    rng = np.random.default_rng(hop + nblk)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    xb = np.ascontiguousarray(x.reshape(nIn, nblk, hop).transpose(1, 0, 2))
    d_in = torch.from_numpy(xb).cuda()
    outs = []
    for batching, splits in ((1, [nblk]), (0, [nblk]), (1, [1, 7, 2, nblk - 10])):
        mc = saf.MatrixConv(hop, H)
        mc.set_option("batching", batching)
        mc.set_option("small_fused", 0)      # the bit-equality below is a property of the K1 -> K2 -> K3 path
        d_out = torch.zeros((nblk, nOut, hop), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        b0 = 0
        for n in splits:
            mc.apply_device(d_in[b0:].data_ptr(), d_out[b0:].data_ptr(), n)
            b0 += n
        mc.synchronize()
        outs.append(d_out.cpu().numpy().transpose(1, 0, 2).reshape(nOut, nblk * hop))
        mc.destroy()
    check(outs[0], ref, "batched device blocks")
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    # host-pointer API (one block per call) through the same three kernels gives the same bits as well
    mh = saf.MatrixConv(hop, H)
    mh.set_option("small_fused", 0)
    mh.set_option("lookahead", 0)        # the tail/head split of the host API sums the partitions in another order
    assert np.array_equal(mh.run(x), outs[0])
    # ... and through the fused small-problem kernel (different summation order) the same values to rounding
    check(saf.MatrixConv(hop, H).run(x), ref, "host API, default path")


@pytest.mark.parametrize("hop,L,nIn,nOut,nblk", [(256, 700, 5, 9, 30), (1024, 12000, 16, 8, 30), (64, 128, 2, 3, 20),
                                                  (128, 2000, 3, 70, 25), (512, 1024, 33, 17, 12),
                                                  (8192, 20000, 40, 2, 5)])      # 1.3 MB input blocks: copy-engine path
def test_lookahead_tail_head_split(saf, orc, hop, L, nIn, nOut, nblk):
    """Host-pointer apply with the look-ahead split (default for P >= 2): partitions p >= 1 of block t+1 are
    accumulated behind block t, the call itself only adds the newest partition.  Same result as the oracle and as
    the plain K1 -> K2 -> K3 sequence within tolerance; deterministic; survives reset, device-pointer calls in
    between (which drop the pre-computed tail) and option changes."""
    import torch
    rng = np.random.default_rng(hop + nOut)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H)
    mc.set_option("small_fused", 0)
    y = mc.run(x)
    check(y, ref, "look-ahead host API")
    mc.reset_state()
    assert np.array_equal(mc.run(x), y)                      # deterministic, reset drops the pending tail
    mp = saf.MatrixConv(hop, H)
    mp.set_option("small_fused", 0)
    mp.set_option("lookahead", 0)
    check(mp.run(x), y, "plain vs look-ahead")
    # host blocks, then device-pointer blocks, then host blocks again on one handle
    mc.reset_state()
    n1, n2 = nblk // 3, nblk // 3
    parts = [mc.run(x[:, :n1 * hop])]
    xb = np.ascontiguousarray(x[:, n1 * hop:(n1 + n2) * hop].reshape(nIn, n2, hop).transpose(1, 0, 2))
    d_in = torch.from_numpy(xb).cuda()
    d_out = torch.empty((n2, nOut, hop), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    mc.apply_device(d_in.data_ptr(), d_out.data_ptr(), n2)
    mc.synchronize()
    parts.append(d_out.cpu().numpy().transpose(1, 0, 2).reshape(nOut, n2 * hop))
    mc.set_option("lookahead", 1)
    parts.append(mc.run(x[:, (n1 + n2) * hop:]))
    check(np.concatenate(parts, 1), ref, "host / device / host")
    mc.destroy(); mp.destroy()


@pytest.mark.parametrize("hop,L,nIn,nOut,T", [(64, 200, 3, 2, 20), (128, 128, 1, 1, 5), (256, 2048, 11, 5, 300),
                                               (32, 1000, 2, 3, 270), (512, 3000, 4, 9, 40), (4096, 6000, 2, 3, 5),
                                               (96, 500, 3, 2, 40), (300, 1000, 5, 7, 30), (1024, 2500, 6, 64, 12),
                                               (128, 600, 3, 70, 20), (256, 700, 2, 130, 9)])      # > 64 outputs: GEMM output tiles
def test_offline_tensor_core_render_vs_oracle(saf, orc, hop, L, nIn, nOut, T):
    """safconv_render_offline (tcgen05 per-bin GEMM, 3xTF32 split, fp32 accumulate in TMEM) == the reference's
    block-by-block convolution from a zero state, within the north_star tolerance."""
    rng = np.random.default_rng(hop + L + T)
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * T)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H)
    y = mc.render_offline(x)
    ma, l2 = check(y, ref, "offline render")
    t = orc.truth_matrix(H, x, np.arange(nOut), 0, hop * T)
    _, l2_gt = err_metrics(y, t)
    _, l2_rt = err_metrics(ref, t)
    print(f"offline: gpu-ref {l2:.3g}  gpu-truth {l2_gt:.3g}  ref-truth {l2_rt:.3g}")
    assert l2_gt <= TOL_REL_L2
    # a second render on the same handle (workspace re-use, shorter signal) and the streaming path still agree
    y2 = mc.render_offline(x[:, :hop * (T // 2)])
    check(y2, ref[:, :hop * (T // 2)], "offline render, shorter")
    ys = mc.run(x)
    check(ys, y, "streaming vs offline")


@pytest.mark.parametrize("nOut", [2, 64])          # frames on the UMMA M axis (Nn = 32) / on the N axis (Nn = 128)
def test_offline_tail_tile_widths(saf, orc, nOut):
    """The persistent GEMM cuts the last 256-frame tile of a render to roundup16(T - t0) columns (frames on N) or to one
    128-row accumulator (frames on M): every width class, on ONE handle whose workspace keeps the rows of the longer
    renders that came before (stale operand rows beyond T must never reach an output)."""
    rng = np.random.default_rng(77 + nOut)
    hop, L, nIn = 32, 100, 3
    Ts = [513, 1, 15, 16, 17, 127, 128, 129, 255, 256, 257, 383, 400]
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * max(Ts))).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H)
    for T in Ts:
        y = mc.render_offline(x[:, :hop * T])
        check(y, ref[:, :hop * T], f"offline render, T = {T}")
    mc.destroy()


@pytest.mark.parametrize("kind", ["f16", "tf32"])
def test_offline_operand_kinds_and_dynamic_range(saf, orc, kind, monkeypatch):
    """Both tensor-core operand types of the offline path (fp16 hi/lo with power-of-two scaling = default, tf32 hi/lo)
    against the oracle, on a signal with 80 dB between a loud block and the quiet rest and decaying filters: the
    usual full-scale tolerance over the whole signal, and the quiet stretch on its own still to 1e-5 relative."""
    monkeypatch.setenv("SAFCONV_OFF_KIND", kind)
    rng = np.random.default_rng(21)
    hop, L, nIn, nOut, T = 128, 600, 5, 3, 60
    H = (rng.uniform(-1, 1, (nOut, nIn, L)) * np.exp(-6.9 * np.arange(L) / L)).astype(np.float32)
    x = (1e-4 * rng.uniform(-1, 1, (nIn, hop * T))).astype(np.float32)
    x[:, 2 * hop:3 * hop] = rng.uniform(-1, 1, (nIn, hop)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H)
    y = mc.render_offline(x)
    check(y, ref, f"offline {kind}")
    q = slice(20 * hop, T * hop)                      # the loud block's response ended at frame 2 + P = 7
    truth = orc.truth_matrix(H, x, np.arange(nOut), 20 * hop, T * hop)
    l2q = float(np.linalg.norm(y[:, q] - truth) / np.linalg.norm(truth))
    l2r = float(np.linalg.norm(ref[:, q] - truth) / np.linalg.norm(truth))
    print(f"offline {kind}: quiet stretch vs fp64 truth: gpu {l2q:.3g}, reference {l2r:.3g}")
    assert l2q <= 1e-5
    mc.destroy()


def test_offline_pipelined_host_render(saf, orc):
    """safconv_render_offline on host buffers longer than one 512-frame segment: time segments pipelined over three
    streams (strided H2D / kernels / strided D2H) -- against the oracle, against the one-piece device render, with
    page-locked and pageable buffers, and with leading halo frames."""
    import torch
    rng = np.random.default_rng(33)
    hop, L, nIn, nOut, T = 64, 300, 3, 4, 1300              # P = 5: three segments of 507 new frames
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * T)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    mc = saf.MatrixConv(hop, H)
    y = mc.render_offline(x)                                 # pageable numpy buffers
    check(y, ref, "pipelined host render (pageable)")
    xin = torch.from_numpy(x).pin_memory(); yout = torch.empty((nOut, hop * T)).pin_memory()
    mc.render_offline_host(xin.data_ptr(), yout.data_ptr(), T)
    check(yout.numpy(), ref, "pipelined host render (page-locked)")
    assert np.array_equal(yout.numpy(), y)
    # one-piece device render of the same signal
    xd = torch.from_numpy(x).cuda(); yd = torch.empty((nOut, hop * T), device="cuda")
    torch.cuda.synchronize()
    mc.render_offline_device(xd.data_ptr(), yd.data_ptr(), T)
    mc.synchronize()
    ma, l2 = err_metrics(yd.cpu().numpy(), y)
    assert ma <= 1e-6 and l2 <= 2e-7, (ma, l2)
    # leading halo frames: render frames [200, T) given frames [190, T)
    h0 = 10
    xs = np.ascontiguousarray(x[:, (200 - h0) * hop:])
    ys = np.empty((nOut, (T - 200) * hop), np.float32)
    fp = C.POINTER(C.c_float)
    rc = saf.lib().safconv_render_offline_segment(mc.handle, xs.ctypes.data_as(fp), ys.ctypes.data_as(fp), T - 200, h0)
    assert rc == 0
    check(ys, ref[:, 200 * hop:], "pipelined host render with halo")
    mc.destroy()


def test_offline_time_segments_equal_full_render(saf):
    """Multi-GPU offline sharding is by time: every segment rendered on its own (with a P-frame input halo,
    sharding.time_segment) reproduces the full render exactly -- no exchange between the GPUs is needed."""
    import torch
    from spatial_audio_framework_b200 import sharding
    rng = np.random.default_rng(17)
    hop, L, nIn, nOut, T = 128, 1000, 4, 3, 301
    P = (L + hop - 1) // hop
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * T)).astype(np.float32)
    mc = saf.MatrixConv(hop, H)
    full = mc.render_offline(x)
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            t0, t1, halo = sharding.time_segment(T, world, r, P)
            xs = torch.from_numpy(np.ascontiguousarray(x[:, (t0 - halo) * hop:t1 * hop])).cuda()
            ys = torch.empty((nOut, (t1 - t0) * hop), dtype=torch.float32, device="cuda")
            torch.cuda.synchronize()
            mc.render_offline_segment_device(xs.data_ptr(), ys.data_ptr(), t1 - t0, halo)
            mc.synchronize()
            parts.append(ys.cpu().numpy())
        got = np.concatenate(parts, 1)
        # fp16 operands are scaled by a power of two derived from each segment's own input bound: bit-identical
        # whenever the segments pick the same exponent, and equal to rounding of the (sub-2^-40) residues otherwise
        ma, l2 = err_metrics(got, full)
        assert ma <= 1e-7 and l2 <= 1e-7, (world, ma, l2)


def test_offline_render_c5_shape_vs_streaming(saf):
    """configs[4] channel counts (121 in x 64 out, hop 1024, 8192 taps) on a shortened signal: the offline
    tensor-core path against the (oracle-verified) streaming path on the same handle."""
    from spatial_audio_framework_b200 import synth
    hop, L, nIn, nOut, T = 1024, 8192, 121, 64, 300
    H = synth.decaying_rir((nOut, nIn, L), seed=5)
    x = synth.uniform((nIn, hop * T), seed=6)
    mc = saf.MatrixConv(hop, H)
    y = mc.render_offline(x)
    ys = mc.run(x)
    ma, l2 = check(y, ys, "C5 offline vs streaming")
    print("C5-shape offline vs streaming", ma, l2, mc.offline_times_ms())


@pytest.mark.parametrize("hop,L,nCH,nblk", [(512, 4096, 16, 70), (64, 300, 3, 41), (2048, 9000, 2, 9),
                                            (100, 450, 5, 37), (256, 5000, 3, 20), (128, 128, 40, 35)])
def test_multiconv_batched_device_blocks(saf, orc, hop, L, nCH, nblk):
    """multiConv on device-resident blocks: all forward FFTs, then all (channel, block) MACs + inverse FFTs, then
    the overlap-add chain -- equal to one fused launch per block (bit-identical where both use the same FFT core),
    across batch and ring boundaries."""
    import torch
    rng = np.random.default_rng(hop + nCH)
    H = rng.uniform(-1, 1, (nCH, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nCH, hop * nblk)).astype(np.float32)
    ref = orc.OracleMultiConv(hop, H, 1).run(x)
    d_in = torch.from_numpy(np.ascontiguousarray(x.reshape(nCH, nblk, hop).transpose(1, 0, 2))).cuda()
    outs = []
    for batching, splits in ((1, [nblk]), (0, [nblk]), (1, [3, 1, nblk - 4])):
        mc = saf.MultiConv(hop, H)
        mc.set_option("batching", batching)
        d_out = torch.zeros((nblk, nCH, hop), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        b0 = 0
        for n in splits:
            mc.apply_device(d_in[b0:].data_ptr(), d_out[b0:].data_ptr(), n)
            b0 += n
        mc.synchronize()
        outs.append(d_out.cpu().numpy().transpose(1, 0, 2).reshape(nCH, nblk * hop))
    check(outs[0], ref, "multiConv batched")
    yh = saf.MultiConv(hop, H).run(x)
    N = 64
    while N < 2 * hop:
        N *= 2
    warp_fft = 64 <= N // 2 <= 512 and -(-L // hop) <= 16        # the batched path's own condition (scdev_multi_batch)
    if not warp_fft:
        # shared-memory FFT everywhere: the batched kernels keep the summation order of the fused one-block kernel
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
        assert np.array_equal(yh, outs[0])
    else:
        # FFT sizes up to 1024 and up to 16 partitions: the batched path runs on the warp-level register FFT, single
        # blocks on the fused shared-memory kernel -- same values to rounding
        for i, y in enumerate((outs[1], outs[2], yh)):
            check(y, ref, f"multiConv path {i + 1}")
            ma, l2 = err_metrics(y, outs[0])
            assert ma <= 2e-6 and l2 <= 5e-7, (i, ma, l2)


def test_pinned_caller_buffers_are_used_directly(saf, orc):
    import torch
    rng = np.random.default_rng(8)
    hop, L, nIn, nOut, nblk = 128, 900, 4, 3, 10
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    lib = saf.lib()
    fp = C.POINTER(C.c_float)
    for graph in (0, 1):
        mr = saf.MatrixConv(hop, H)                   # same launch path, ordinary (pageable) caller buffers
        mr.set_option("use_graph", graph)
        ref = mr.run(x)
        mc = saf.MatrixConv(hop, H)
        mc.set_option("use_graph", graph)
        xin = torch.empty((nIn, hop)).pin_memory()
        yout = torch.empty((nOut, hop)).pin_memory()
        got = []
        for b in range(nblk):
            xin.copy_(torch.from_numpy(x[:, b * hop:(b + 1) * hop]))
            lib.saf_matrixConv_apply(mc.handle, C.cast(xin.data_ptr(), fp), C.cast(yout.data_ptr(), fp))
            got.append(yout.numpy().copy())
        assert np.array_equal(np.concatenate(got, 1), ref)


def test_output_channel_shards_concatenate(saf, orc):
    """Output channels are independent (reference .c:218-234): shard outputs == full run, bit for bit."""
    rng = np.random.default_rng(5)
    hop, L, nIn, nOut, nblk = 512, 4000, 5, 12, 6
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    full = saf.MatrixConv(hop, H).run(x)
    parts = [saf.MatrixConv(hop, H, shard=(b, 3)).run(x) for b in range(0, nOut, 3)]
    y = np.concatenate(parts, 0)
    check(y, orc.OracleMatrixConv(hop, H, 1).run(x), "shards vs oracle")
    check(y, full, "shards vs full")
    Hm = rng.uniform(-1, 1, (8, L)).astype(np.float32)
    xm = rng.uniform(-1, 1, (8, hop * nblk)).astype(np.float32)
    fullm = saf.MultiConv(hop, Hm).run(xm)
    pm = [saf.MultiConv(hop, Hm, shard=(b, 4)).run(xm[b:b + 4]) for b in (0, 4)]
    assert np.array_equal(np.concatenate(pm, 0), fullm)


def test_c4_long_rir_parity_and_properties(saf, orc):
    """BASELINE.json configs[3] (64 in, hop 1024, 96000 taps) at full filter length.

    The full 64x64 reference run costs ~11 s/block on one core, so parity is checked on a
    shard of output channels with the full nIn x P = 6016-term sum per bin, plus size-independent
    properties on the whole 64x64 problem (linearity, delayed-unit-impulse filters, zero input)."""
    hop, L, nIn = 1024, 96000, 64
    nblk = 97                                   # > P = 94 so the delay line wraps around
    from spatial_audio_framework_b200 import synth
    outs = 1
    H = synth.decaying_rir((outs, nIn, L), seed=11)
    x = synth.uniform((nIn, hop * nblk), seed=12)
    y = saf.MatrixConv(hop, H).run(x)
    ref = orc.OracleMatrixConv(hop, H, 1).run(x)
    n0, n1 = hop * (nblk - 1), hop * nblk        # last block: every partition contributes
    t = orc.truth_matrix(H, x, np.arange(outs), n0, n1)
    ma_gr, l2_gr = err_metrics(y, ref)
    ma_gt, l2_gt = err_metrics(y[:, n0:n1], t)
    ma_rt, l2_rt = err_metrics(ref[:, n0:n1], t)
    print(f"C4 shard: gpu-ref max/fs {ma_gr:.3g} L2 {l2_gr:.3g} | gpu-truth {ma_gt:.3g} {l2_gt:.3g} | ref-truth {ma_rt:.3g} {l2_rt:.3g}")
    assert ma_gr <= TOL_MAXABS_FS
    # relative L2 <= 1e-6 vs the reference, OR strictly closer to the fp64 truth than the reference is
    # (the reference's own sequential fp32 sum of 6016 segments sits ~1.9e-6 from truth)
    assert l2_gr <= TOL_REL_L2 or l2_gt < l2_rt
    assert l2_gt <= 1e-6


def test_c4_full_size_properties(saf):
    hop, L, nIn, nOut = 1024, 96000, 64, 64
    P = (L + hop - 1) // hop
    rng = np.random.default_rng(21)
    # filters = scaled unit impulses at known delays -> output is an exactly known mix of delayed inputs
    H = np.zeros((nOut, nIn, L), np.float32)
    delays = rng.integers(0, L, size=(nOut, nIn))
    gains = rng.uniform(-1, 1, size=(nOut, nIn)).astype(np.float32)
    for no in range(nOut):
        H[no, np.arange(nIn), delays[no]] = gains[no]
    nblk = P + 3
    x = rng.uniform(-1, 1, (nIn, hop * nblk)).astype(np.float32)
    mc = saf.MatrixConv(hop, H)
    info = mc.info()
    assert info.numFilterBlocks == P and info.fftSize == 2 * hop
    y = mc.run(x)
    T = hop * nblk
    exp = np.zeros((nOut, T), np.float64)
    for no in range(nOut):
        for ni in range(nIn):
            d = delays[no, ni]
            exp[no, d:] += gains[no, ni] * x[ni, :T - d].astype(np.float64)
    ma, l2 = err_metrics(y, exp)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, (ma, l2)
    # linearity + time invariance of the block engine: conv(a*x1 + x2) == a*conv(x1) + conv(x2)
    x2 = rng.uniform(-1, 1, x.shape).astype(np.float32)
    mc.reset_state(); y2 = mc.run(x2)
    mc.reset_state(); y3 = mc.run((0.5 * x + x2).astype(np.float32))
    ma, l2 = err_metrics(y3, 0.5 * y.astype(np.float64) + y2)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, (ma, l2)
    # zero input after reset -> exactly zero output
    mc.reset_state()
    assert not np.any(mc.apply(np.zeros((nIn, hop), np.float32)))


def test_c3_full_size_properties(saf):
    """configs[2] at full size (256 channels, hop 512, 4096 taps): delayed-impulse filters give an exactly known
    output; host API (one fused launch per block) and the batched device path (warp-FFT kernels)."""
    import torch
    hop, L, nCH, nblk = 512, 4096, 256, 24
    rng = np.random.default_rng(31)
    H = np.zeros((nCH, L), np.float32)
    delays = rng.integers(0, L, size=nCH)
    gains = rng.uniform(-1, 1, size=nCH).astype(np.float32)
    H[np.arange(nCH), delays] = gains
    x = rng.uniform(-1, 1, (nCH, hop * nblk)).astype(np.float32)
    T = hop * nblk
    exp = np.zeros((nCH, T), np.float64)
    for c in range(nCH):
        exp[c, delays[c]:] = gains[c] * x[c, :T - delays[c]].astype(np.float64)
    mc = saf.MultiConv(hop, H)
    ma, l2 = err_metrics(mc.run(x), exp)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, (ma, l2)
    mc.reset_state()
    d_in = torch.from_numpy(np.ascontiguousarray(x.reshape(nCH, nblk, hop).transpose(1, 0, 2))).cuda()
    d_out = torch.zeros((nblk, nCH, hop), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    mc.apply_device(d_in.data_ptr(), d_out.data_ptr(), nblk)
    mc.synchronize()
    ma, l2 = err_metrics(d_out.cpu().numpy().transpose(1, 0, 2).reshape(nCH, T), exp)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, (ma, l2)


def test_c5_full_channels_offline_properties(saf):
    """configs[4] channel counts and filter length (121 x 64, 8192 taps, hop 1024) on 1.3 s of audio through the
    offline tensor-core path: delayed-impulse filters -> exactly known mix of delayed inputs; linearity."""
    hop, L, nIn, nOut, T = 1024, 8192, 121, 64, 60
    rng = np.random.default_rng(41)
    H = np.zeros((nOut, nIn, L), np.float32)
    delays = rng.integers(0, L, size=(nOut, nIn))
    gains = rng.uniform(-1, 1, size=(nOut, nIn)).astype(np.float32)
    for no in range(nOut):
        H[no, np.arange(nIn), delays[no]] = gains[no]
    x = rng.uniform(-1, 1, (nIn, hop * T)).astype(np.float32)
    n = hop * T
    exp = np.zeros((nOut, n), np.float64)
    for no in range(nOut):
        for ni in range(nIn):
            d = delays[no, ni]
            exp[no, d:] += gains[no, ni] * x[ni, :n - d].astype(np.float64)
    mc = saf.MatrixConv(hop, H)
    y = mc.render_offline(x)
    ma, l2 = err_metrics(y, exp)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, (ma, l2)
    x2 = rng.uniform(-1, 1, x.shape).astype(np.float32)
    y2 = mc.render_offline(x2)
    y3 = mc.render_offline((0.5 * x + x2).astype(np.float32))
    ma, l2 = err_metrics(y3, 0.5 * y.astype(np.float64) + y2)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2, (ma, l2)


@pytest.mark.parametrize("nCH,xl,hl", [(3, 1000, 129), (1, 5, 7), (2, 5000, 4096), (4, 300, 1), (6, 48000, 700)])
def test_fftconv_fftfilt_vs_oracle(saf, orc, nCH, xl, hl):
    """safconv_fftconv / safconv_fftfilt (drop-in for saf_utility_fft.h:86-113, served by the multiConv engine)
    against the oracle's restatement of the reference's single-FFT implementation."""
    import spatial_audio_framework_b200 as pkg
    rng = np.random.default_rng(nCH * 7 + xl + hl)
    x = rng.uniform(-1, 1, (nCH, xl)).astype(np.float32)
    h = rng.uniform(-1, 1, (nCH, hl)).astype(np.float32)
    check(pkg.fftconv(x, h), orc.oracle_fftconv(x, h), "fftconv")
    check(pkg.fftfilt(x, h), orc.oracle_fftconv(x, h, True), "fftfilt")
    # the reference-named weak symbols
    import ctypes as C
    L = saf.lib()
    y = np.zeros((nCH, xl + hl - 1), np.float32)
    fp = C.POINTER(C.c_float)
    L.fftconv.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, fp]; L.fftconv.restype = None
    L.fftconv(x.ctypes.data_as(fp), h.ctypes.data_as(fp), xl, hl, nCH, y.ctypes.data_as(fp))
    check(y, orc.oracle_fftconv(x, h), "fftconv (reference-named symbol)")


def test_rfft_pair_vs_golden_and_oracle(saf, orc):
    """The CONVOLVERS' own power-of-two FFT cores on their own (safconv_fft.cuh through the debug entry point) against the
    reference's saf_rfft golden vectors (tests/golden/rfft_*.npz, power-of-two sizes) and the oracle's KissFFT
    restatement on random batches; round trip like the reference's own test__saf_rfft (<= 1e-5).  The public saf_rfft_*
    API (any even N, general mixed-radix engine) is tested in tests/test_gpu_rfft.py."""
    import spatial_audio_framework_b200 as pkg
    from conftest import golden_files
    seen = 0
    for f in golden_files("rfft"):
        g = np.load(f)
        N = int(g["N"])
        if N & (N - 1) or N < 64:
            continue
        seen += 1
        Xref = g["X"].astype(np.float32).reshape(-1, 2)
        X = pkg.convolver_rfft(g["x"][None, :])[0]
        err = np.abs(X - (Xref[:, 0] + 1j * Xref[:, 1])).max() / np.abs(Xref).max()
        assert err <= 1e-6, (f, err)
        xb = pkg.convolver_rfft((Xref[:, 0] + 1j * Xref[:, 1])[None, :], inverse=True)[0]
        assert np.abs(xb - g["xb"]).max() <= 1e-6 * max(np.abs(g["xb"]).max(), 1.0), f
    assert seen >= 3
    rng = np.random.default_rng(12)
    for N in (64, 128, 512, 1024, 2048, 4096, 16384):
        x = rng.uniform(-1, 1, (5, N)).astype(np.float32)
        X = pkg.convolver_rfft(x)
        for b in (0, 4):
            Xo, _ = orc.oracle_rfft(N, x[b])
            Xo = Xo.reshape(-1, 2)
            Xo = Xo[:, 0] + 1j * Xo[:, 1]
            l2 = np.linalg.norm(X[b] - Xo) / np.linalg.norm(Xo)
            assert l2 <= 1e-6, (N, b, l2)
        xr = pkg.convolver_rfft(X, inverse=True)
        assert np.abs(xr - x).max() <= 1e-5, N              # the reference's own criterion (test__utilities_module.c:381-404)


def test_two_live_handles_with_different_plans_share_kernels(saf, orc):
    """Two handles alive at once whose MAC kernels are the SAME instantiation (R = 8) with different shared-memory needs
    (64 inputs: a deep TMA stage ring; 1 input: a shallow one).  The opt-in shared-memory limit of a kernel is global per
    device, so creating the small plan must not lower it under the large plan's launches (round-1 advisor finding):
    create A (large), create B (small), then apply A, B, A block by block against the oracle."""
    rng = np.random.default_rng(42)
    hop, L, nblk = 1024, 2048, 3
    HA = rng.uniform(-1, 1, (64, 64, L)).astype(np.float32)
    HB = rng.uniform(-1, 1, (64, 1, L)).astype(np.float32)
    xA = rng.uniform(-1, 1, (64, hop * nblk)).astype(np.float32)
    xB = rng.uniform(-1, 1, (1, hop * nblk)).astype(np.float32)
    refA = orc.OracleMatrixConv(hop, HA, 1).run(xA)
    refB = orc.OracleMatrixConv(hop, HB, 1).run(xB)
    A = saf.MatrixConv(hop, HA, 1)
    B = saf.MatrixConv(hop, HB, 1)
    for conv in (A, B):
        conv.set_option("small_fused", 0)
        conv.set_option("lookahead", 0)          # plain K1 -> K2 (full pass) -> K3: the launches the finding was about
    yA = np.zeros_like(refA); yB = np.zeros_like(refB)
    for b in range(nblk):
        sl = slice(b * hop, (b + 1) * hop)
        yA[:, sl] = A.apply(xA[:, sl])
        yB[:, sl] = B.apply(xB[:, sl])
    check(yA, refA, "large plan beside a small one")
    check(yB, refB, "small plan beside a large one")
    # and with the default path (look-ahead tail / head passes), a third handle created in between
    A2 = saf.MatrixConv(hop, HA, 1)
    C2 = saf.MatrixConv(hop, HB[:8], 1)
    y2 = A2.run(xA)
    check(y2, refA, "large plan, look-ahead passes")
    check(C2.run(xB), refB[:8], "8 x 1 plan created after it")
    for conv in (A, B, A2, C2):
        conv.destroy()
