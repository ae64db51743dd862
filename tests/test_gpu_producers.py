"""GPU parity tests of the filter producers, through the C ABI with host pointers:
``getBinauralAmbiDecoderFilters`` / ``getBinauralAmbiDecoderMtx`` and ``ims_shoebox_*`` of libsafconv_b200.so against

* the committed golden outputs of the compiled, unmodified reference (tests/golden/producers_*.npz),
* the numpy fp64 restatement of the reference algorithms (oracle/producers.py, pinned on the reference by
  tests/test_producers_cpu.py) as ground truth, on inputs the goldens do not hold,
* size-independent properties at a configs[3]-sized room response.

Tolerance.  Decoder filters: the north_star bar (max-abs <= 1e-5 of full scale, rel-L2 <= 1e-6) against the fp64 truth;
against the reference's fp32 LAPACK result the same bar OR "at least as close to the truth as the reference is" (the
reference itself is up to 1e-6 (MagLS: 2e-5) from the truth).  MagLS chains one fp32 decoder per band into the next
band's phases, so its own bar against the truth is 5e-6.  Image-source RIRs: same image count, same occupied taps, values to the north_star bar (observed: 1e-7 ... 4e-7 rel-L2).
"""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN, TOL_MAXABS_FS, TOL_REL_L2, err_metrics

pytestmark = pytest.mark.gpu

from oracle import producers as PR  # noqa: E402
from spatial_audio_framework_b200 import synth  # noqa: E402


def gold(name):
    return np.load(GOLDEN / f"producers_{name}.npz")


def check_decoder(mine, ref, truth, magls, what):
    ma_t, l2_t = err_metrics(mine, truth)
    ma_r, l2_r = err_metrics(mine, ref)
    ma_rt, l2_rt = err_metrics(ref, truth)
    print(f"{what}: vs truth {l2_t:.2e} / {ma_t:.2e}; vs reference {l2_r:.2e} / {ma_r:.2e}; reference vs truth {l2_rt:.2e}")
    bar = 5e-6 if magls else TOL_REL_L2
    assert l2_t <= bar and ma_t <= TOL_MAXABS_FS, f"{what}: {l2_t:.3g} / {ma_t:.3g} from the fp64 truth"
    assert (l2_r <= TOL_REL_L2 and ma_r <= TOL_MAXABS_FS) or l2_t <= l2_rt, f"{what}: {l2_r:.3g} from the reference, {l2_t:.3g} from truth"


DEC_CASES = [(m, dc, mr) for m in (PR.DEFAULT, PR.LS, PR.LSDIFFEQ, PR.TA, PR.MAGLS) for dc in (0, 1) for mr in (0, 1)]


@pytest.mark.parametrize("m,dc,mr", DEC_CASES)
def test_decoder_filters_vs_reference_golden(saf, m, dc, mr):
    g = gold("decoder")
    H, d, itd, fftSize, fs, order = g["hrtfs"], g["dirs_deg"], g["itd_s"], int(g["fftSize"]), float(g["fs"]), int(g["order"])
    mine = saf.producers.decoder_filters(H, d, fftSize, fs, m, order, itd, None, dc, mr)
    truth = PR.np_decoder_filters(H, d, fftSize, fs, m, order, itd, None, dc, mr)
    check_decoder(mine, g[f"f_m{m}_dc{dc}_mr{mr}"], truth, m == PR.MAGLS, f"method {m} diffCM {dc} maxRE {mr}")


def test_decoder_filters_weights_orders_and_reference_name(saf):
    g = gold("decoder")
    H, d, itd, fftSize, fs = g["hrtfs"], g["dirs_deg"], g["itd_s"], int(g["fftSize"]), float(g["fs"])
    for key, kw in (("f_m1_weights", dict(method=PR.LS, order=3, weights=g["weights"], diffCM=1, maxRE=1)),
                    ("f_m5_o1", dict(method=PR.MAGLS, order=1)),
                    ("f_m2_o5", dict(method=PR.LSDIFFEQ, order=5, maxRE=1))):
        args = (H, d, fftSize, fs, kw["method"], kw["order"], itd, kw.get("weights"), kw.get("diffCM", 0), kw.get("maxRE", 0))
        mine = saf.producers.decoder_filters(*args)
        check_decoder(mine, g[key], PR.np_decoder_filters(*args), kw["method"] == PR.MAGLS, key)
        # the reference's own symbol name (weak export) is the same code
        again = saf.producers.decoder_filters(*args, reference_name=True)
        assert np.array_equal(mine, again)


needs_tdesign = pytest.mark.skipif(not PR.producers_reference_available(),
                                   reason="SAF's t-design tables are not in this repo: the tests take them from oracle/_ref/libsaf_ref_producers.so")


@needs_tdesign
def test_spr_decoder_vs_reference_golden_and_truth(saf):
    """BINAURAL_DECODER_SPR (saf_hoa_internal.c:332-430): condition numbers on the device pick the interpolation order, the
    t-design of degree 2 * order is handed over with safconv_register_tdesign"""
    g = gold("decoder")
    R = PR.load_producers_reference()
    H, d, itd, fftSize, fs, order = g["hrtfs"], g["dirs_deg"], g["itd_s"], int(g["fftSize"]), float(g["fs"]), int(g["order"])
    td = R.tdesign(2 * order)
    saf.producers.register_tdesign(2 * order, td)
    for dc in (0, 1):
        for mr in (0, 1):
            mine = saf.producers.decoder_filters(H, d, fftSize, fs, PR.SPR, order, itd, None, dc, mr)
            truth = PR.np_decoder_filters(H, d, fftSize, fs, PR.SPR, order, itd, None, dc, mr, tdesign_deg=td)
            check_decoder(mine, g[f"f_m3_dc{dc}_mr{mr}"], truth, False, f"SPR diffCM {dc} maxRE {mr}")
    # weights (the projection uses weights / 4 pi, the condition check the raw weights), another order
    w = g["weights"] * np.float32(4 * np.pi)
    saf.producers.register_tdesign(2, R.tdesign(2))
    mine = saf.producers.decoder_filters(H, d, fftSize, fs, PR.SPR, 1, itd, w, 0, 0)
    truth = PR.np_decoder_filters(H, d, fftSize, fs, PR.SPR, 1, itd, w, 0, 0, tdesign_deg=R.tdesign(2))
    check_decoder(mine, g["f_m3_o1_weights"], truth, False, "SPR order 1, weights")


@needs_tdesign
@pytest.mark.parametrize("nD,order,fftSize", [(836, 5, 256), (64, 1, 64), (400, 7, 128), (50, 2, 32), (12, 3, 16)])
def test_spr_decoder_other_grids_vs_live_reference(saf, nD, order, fftSize):
    """interpolation orders up to 20 (441 SH channels at 836 directions), an irregular grid whose condition numbers cross 100
    early, and a grid that cannot carry the requested order (the reference asserts there; here: error)"""
    R = PR.load_producers_reference()
    if nD == 50:
        rng = np.random.default_rng(0)
        d = np.stack([rng.uniform(-180, 180, nD), np.degrees(np.arcsin(rng.uniform(-1, 1, nD)))], 1).astype(np.float32)
        H, _, itd = synth.synthetic_hrtfs(nD, fftSize, 48000.0, seed=nD)
    else:
        H, d, itd = synth.synthetic_hrtfs(nD, fftSize, 48000.0, seed=nD)
    td = R.tdesign(2 * order)
    saf.producers.register_tdesign(2 * order, td)
    nh, conds = PR.np_spr_order(d, nD)
    if nh < order:
        with pytest.raises(saf.SafConvError, match="modal order"):
            saf.producers.decoder_filters(H, d, fftSize, 48000.0, PR.SPR, order, itd)
        return
    mine = saf.producers.decoder_filters(H, d, fftSize, 48000.0, PR.SPR, order, itd, None, 0, 1)
    truth = PR.np_decoder_filters(H, d, fftSize, 48000.0, PR.SPR, order, itd, None, 0, 1, tdesign_deg=td)
    ref = R.decoder_filters(H, d, fftSize, 48000.0, PR.SPR, order, itd, None, 0, 1)
    check_decoder(mine, ref, truth, False, f"SPR nD {nD} order {order} (Nh {nh})")


@pytest.mark.parametrize("order,nD,fftSize,m,dc,mr", [
    (7, 836, 512, PR.LS, 1, 1),          # 64 SH channels, a KU100-sized grid, non-trivial flags
    (10, 1202, 256, PR.LSDIFFEQ, 0, 1),  # the largest supported order (121 channels)
    (4, 400, 1000, PR.MAGLS, 1, 0),      # non-power-of-two fftSize (general-size inverse FFT), ~470 chained bands
    (2, 64, 6, PR.TA, 0, 0),             # tiny: 4 bands
    (0, 12, 64, PR.MAGLS, 0, 1),         # order 0
])
def test_decoder_filters_vs_truth_other_shapes(saf, order, nD, fftSize, m, dc, mr):
    H, d, itd = synth.synthetic_hrtfs(nD, fftSize, 48000.0, seed=order + nD)
    mine = saf.producers.decoder_filters(H, d, fftSize, 48000.0, m, order, itd, None, dc, mr)
    truth = PR.np_decoder_filters(H, d, fftSize, 48000.0, m, order, itd, None, dc, mr)
    ma, l2 = err_metrics(mine, truth)
    print(f"order {order} nD {nD} fft {fftSize} method {m}: {l2:.2e} / {ma:.2e}")
    assert l2 <= (5e-6 if m == PR.MAGLS else TOL_REL_L2) and ma <= TOL_MAXABS_FS, (ma, l2)
    if PR.producers_reference_available() and order <= 7:
        ref = PR.load_producers_reference().decoder_filters(H, d, fftSize, 48000.0, m, order, itd, None, dc, mr)
        check_decoder(mine, ref, truth, m == PR.MAGLS, "live reference")


@pytest.mark.parametrize("order,nD,fftSize", [(7, 836, 1024), (3, 146, 128), (10, 1202, 256), (1, 12, 64), (5, 2702, 512)])
def test_magls_cluster_kernel_vs_single_cta_kernel(saf, monkeypatch, order, nD, fftSize):
    """the MagLS band recurrence on a thread-block cluster (direction slices of Y / G in the shared memory of 8 CTAs, partial
    decoders exchanged through distributed shared memory) against the one-CTA kernel and the fp64 truth"""
    H, d, itd = synth.synthetic_hrtfs(nD, fftSize, 48000.0, seed=nD)
    monkeypatch.setenv("SAFCONV_MAGLS_CLUSTER", "1")
    a = saf.producers.decoder_filters(H, d, fftSize, 48000.0, PR.MAGLS, order, itd)
    monkeypatch.setenv("SAFCONV_MAGLS_CLUSTER", "0")
    b = saf.producers.decoder_filters(H, d, fftSize, 48000.0, PR.MAGLS, order, itd)
    truth = PR.np_decoder_filters(H, d, fftSize, 48000.0, PR.MAGLS, order, itd)
    (ma_a, l2_a), (ma_b, l2_b), (ma_ab, l2_ab) = err_metrics(a, truth), err_metrics(b, truth), err_metrics(a, b)
    print(f"order {order} nD {nD} fft {fftSize}: cluster {l2_a:.2e}, one CTA {l2_b:.2e} from truth; cluster vs one CTA {l2_ab:.2e}")
    assert l2_a <= 5e-6 and l2_b <= 5e-6 and l2_ab <= 5e-6


def test_decoder_mtx_arbitrary_bands(saf):
    """getBinauralAmbiDecoderMtx with a caller-supplied band grid (the hybrid-filterbank use of ambi_bin.c:284-305)"""
    H, d, itd = synth.synthetic_hrtfs(300, 264, 48000.0, seed=5)          # 133 "bands"
    freqs = np.sort(np.random.default_rng(1).uniform(0, 24000, H.shape[0])).astype(np.float32)
    for m, dc, mr in ((PR.LS, 0, 0), (PR.TA, 1, 1), (PR.MAGLS, 0, 0), (PR.LSDIFFEQ, 1, 0)):
        mine = saf.producers.decoder_mtx(H, d, m, 3, freqs, itd, None, dc, mr)
        truth = PR.np_decoder_mtx(H, d, m, 3, freqs, itd, None, dc, mr)
        ma, l2 = err_metrics(np.stack([mine.real, mine.imag]), np.stack([truth.real, truth.imag]))
        assert l2 <= (5e-6 if m == PR.MAGLS else TOL_REL_L2), (m, ma, l2)


def test_decoder_errors_on_device(saf):
    H, d, itd = synth.synthetic_hrtfs(9, 16)             # 9 directions cannot resolve order 3 (16 SH): singular Gram matrix
    with pytest.raises(saf.SafConvError, match="singular"):
        saf.producers.decoder_filters(H, d, 16, 48000.0, PR.LS, 3)
    with pytest.raises(saf.SafConvError, match="SPR"):
        saf.producers.decoder_filters(H, d, 16, 48000.0, PR.SPR, 9)         # no t-design of degree 18 registered
    # and the next good call is clean
    saf.producers.decoder_filters(H, d, 16, 48000.0, PR.LS, 1)


def test_decoder_straight_into_a_convolver(saf, orc):
    """safconv_binauralDecoder_create_matrixConv: filters designed, partitioned and transformed on the device == a
    convolver created from the downloaded filters (bit for bit), and == the reference convolver on those filters."""
    P = saf.producers
    H, d, itd = synth.synthetic_hrtfs(240, 512, 48000.0, seed=11)
    order, hop = 3, 128
    filt = P.decoder_filters(H, d, 512, 48000.0, PR.MAGLS, order, itd, None, 1, 1)
    h = P.decoder_matrixconv(hop, H, d, 512, 48000.0, PR.MAGLS, order, None, 1, 1)
    mc = saf.MatrixConv(hop, filt, 1)
    x = np.random.default_rng(2).uniform(-1, 1, (16, hop * 9)).astype(np.float32)
    ref = orc.OracleMatrixConv(hop, filt, 1).run(x)
    y = np.concatenate([P.apply_raw(h, x[:, b * hop:(b + 1) * hop], 2) for b in range(9)], 1)
    assert np.array_equal(y, mc.run(x))
    ma, l2 = err_metrics(y, ref)
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2
    P.destroy_raw(h); mc.destroy()


# ------------------------------------------------------------------------------------------------------------------
#  image sources
# ------------------------------------------------------------------------------------------------------------------
IMS_CASES = {   # the cases of tests/golden/make_golden_producers.py
    "t_o0": (0, -1, 0.10, 7, [5.1, 6.0, 1.1], [8.8, 5.5, 1.0]),
    "t_o3": (3, -1, 0.08, 7, [2.1, 1.0, 1.3], [8.8, 5.5, 0.9]),
    "t_o7": (7, -1, 0.04, 2, [4.4, 3.0, 1.4], [3.3, 2.5, 1.7]),
    "n_o2": (2, 4, -1.0, 7, [6.4, 4.0, 1.3], [1.0, 6.5, 2.0]),
    "n_o5_direct": (5, 0, -1.0, 7, [8.5, 5.0, 1.8], [8.8, 5.5, 0.9]),
}


def check_rir(rir, ref, what=""):
    assert rir.shape == ref.shape, f"{what}: shape {rir.shape} vs {ref.shape}"
    assert np.array_equal(rir[0] != 0, ref[0] != 0), f"{what}: different taps are occupied"      # omni channel: all terms positive
    ma, l2 = err_metrics(rir, ref)
    assert l2 <= TOL_REL_L2 and ma <= TOL_MAXABS_FS, f"{what}: {l2:.3g} / {ma:.3g}"


@pytest.fixture(params=["windows", "global_atomics"])
def render_path(request, monkeypatch):
    """both render kernels: ims_window_kernel (default: taps of a window in shared memory) and ims_render_kernel (lattice
    scan with fp64 atomics in HBM, SAFCONV_IMS_WINDOWS=0)"""
    monkeypatch.setenv("SAFCONV_IMS_WINDOWS", "1" if request.param == "windows" else "0")
    return request.param


@pytest.mark.parametrize("names", [False, True])
@pytest.mark.parametrize("name", sorted(IMS_CASES))
def test_ims_rir_vs_reference_golden(saf, name, names, render_path):
    g = gold("ims")
    order, maxN, maxT, nB, src, rec = IMS_CASES[name]
    s = saf.producers.ImsShoebox(synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL[:nB], 125.0, nB, 343.0, 48e3, reference_names=names)
    sid = s.add_source(src); rid = s.add_receiver_sh(order, rec)
    s.compute_echograms(maxN, maxT); s.render_rirs(0)
    assert s.num_images(rid, sid) == g[f"{name}_times"].size
    check_rir(s.rir(rid, sid), g[f"{name}_rir"], name)
    s.destroy()


def test_ims_reference_unit_test_sequence(saf):
    """test/src/test__reverb_module.c:27-96: add / remove / re-add sources (ID assignment), move a source and the
    receiver ten times, render after every move -- IDs and the final RIRs of all three active sources vs the reference"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mgp", GOLDEN / "make_golden_producers.py")
    # the sequence is defined once, next to the golden generator; it needs the oracle only for its own main()
    mgp = importlib.util.module_from_spec(spec); spec.loader.exec_module(mgp)
    g = gold("ims")
    ids, rirs = mgp.ims_unit_test_sequence(saf.producers.ImsShoebox)
    assert ids == list(g["ut_ids"])
    assert len(rirs) == 3
    for (rid, sid), r in rirs.items():
        check_rir(r, g[f"ut_rir_r{rid}_s{sid}"], f"unit test r{rid} s{sid}")


def test_ims_many_pairs_one_batch_and_refresh_rules(saf, render_path):
    """2 receivers (orders 1 and 4) x 3 sources rendered by one renderRIRs call; then the refresh rules of
    saf_reverb.c:184-257 in order-limited mode: only what changed is rendered again"""
    room, aw = [6.0, 5.0, 2.8], synth.IMS_TEST_ABS_WALL[:5]
    s = saf.producers.ImsShoebox(room, aw, 125.0, 5, 343.0, 48e3)
    srcs = [[1.2, 3.3, 1.1], [4.0, 4.1, 2.0], [2.5, 0.7, 0.4]]
    recs = [(1, [4.4, 1.7, 1.6]), (4, [1.0, 1.0, 1.0])]
    sids = [s.add_source(p) for p in srcs]
    rids = [s.add_receiver_sh(o, p) for o, p in recs]
    s.compute_echograms(6, -1.0); s.render_rirs(0)
    for (o, rp), rid in zip(recs, rids):
        for sp, sid in zip(srcs, sids):
            ref, _ = PR.np_ims_rir(room, aw, 5, 343.0, 48e3, sp, rp, o, 6, -1.0)
            check_rir(s.rir(rid, sid), ref, f"rec order {o}")
    before = {(r, k): s.rir(r, k) for r in rids for k in sids}
    # nothing changed -> nothing is rendered again (the host copies stay the same objects' content), then move one source
    s.compute_echograms(6, -1.0); s.render_rirs(0)
    s.update_source(sids[1], [4.0, 4.1, 2.1])
    s.compute_echograms(6, -1.0); s.render_rirs(0)
    for r, (o, rp) in zip(rids, recs):
        for k, sp in zip(sids, srcs):
            now = s.rir(r, k)
            if k == sids[1]:
                ref, _ = PR.np_ims_rir(room, aw, 5, 343.0, 48e3, [4.0, 4.1, 2.1], rp, o, 6, -1.0)
                check_rir(now, ref, "moved source")
            else:
                assert np.array_equal(now, before[(r, k)])
    # new walls -> everything again
    aw2 = aw * np.float32(0.5)
    s.set_abs(aw2); s.compute_echograms(6, -1.0); s.render_rirs(0)
    ref, _ = PR.np_ims_rir(room, aw2, 5, 343.0, 48e3, srcs[0], recs[1][1], 4, 6, -1.0)
    check_rir(s.rir(rids[1], sids[0]), ref, "new absorption")
    # errors: fractional delays (unimplemented upstream too), unknown ids, an echogram that cannot hold the direct path
    with pytest.raises(saf.SafConvError, match="fractional"):
        s.render_rirs(1)
    with pytest.raises(saf.SafConvError):
        s.rir(rids[0], 99)
    s.compute_echograms(-1, 1e-4)
    with pytest.raises(saf.SafConvError, match="empty"):
        s.render_rirs(0)
    s.destroy()


def test_ims_broadband_and_order_10(saf, render_path):
    """nBands = 1 (the reference's own renderRIRs crashes there) and the largest receiver order, vs the restatement"""
    room, aw = [7.0, 4.0, 3.5], synth.IMS_TEST_ABS_WALL[3:4]
    s = saf.producers.ImsShoebox(room, aw, 125.0, 1, 343.0, 44100.0)
    sid = s.add_source([1.0, 1.5, 1.2]); rid = s.add_receiver_sh(10, [5.5, 2.5, 1.8])
    s.compute_echograms(-1, 0.06); s.render_rirs(0)
    ref, _ = PR.np_ims_rir(room, aw, 1, 343.0, 44100.0, [1.0, 1.5, 1.2], [5.5, 2.5, 1.8], 10, -1, 0.06)
    check_rir(s.rir(rid, sid), ref, "order 10 broadband")
    s.destroy()


def test_ims_long_response_properties_and_checksum(saf, render_path):
    """A 1 s response (1.7 M lattice points, ~0.9 M image sources, 48 001 taps, 16 channels): tap positions from an
    independent numpy pass over the lattice, the omni channel's checksum, silence before the direct sound."""
    room, nB = synth.IMS_TEST_ROOM, 7
    src, rec, order = [5.1, 6.0, 1.1], [8.8, 5.5, 0.9], 3
    s = saf.producers.ImsShoebox(room, synth.IMS_TEST_ABS_WALL, 125.0, nB, 343.0, 48e3)
    sid = s.add_source(src); rid = s.add_receiver_sh(order, rec)
    s.compute_echograms(-1, 1.0); s.render_rirs(0)
    rir = s.rir(rid, sid)
    ref, idx = PR.np_ims_rir(room, synth.IMS_TEST_ABS_WALL, nB, 343.0, 48e3, src, rec, order, -1, 1.0)
    assert s.num_images(rid, sid) == idx.size and idx.size > 500000
    assert rir.shape == ref.shape and rir.shape[0] == 16 and rir.shape[1] >= 47990
    first = idx.min()
    assert not rir[:, :first].any() and rir[0, first] > 0
    assert (rir[0] >= 0).all()                                     # omni channel: every term is positive
    assert abs(rir[0].astype(np.float64).sum() - ref[0].sum()) <= 1e-6 * ref[0].sum()
    check_rir(rir, ref, "1 s response")
    s.destroy()


def test_ims_both_render_kernels_agree_on_a_full_length_bank(saf, monkeypatch):
    """4 sources x a 7th-order receiver x 2 s (configs[3]'s filter length: 96 001 taps, 64 channels, 25.8 M image sources):
    the windowed kernel and the lattice-scan kernel find the same images and agree to fp32 rounding of the same fp64 sums"""
    rng = np.random.default_rng(0)
    pos = [[rng.uniform(0.5, 9.5), rng.uniform(0.5, 6.5), rng.uniform(0.5, 2.5)] for _ in range(4)]
    out = {}
    for path in ("1", "0"):
        monkeypatch.setenv("SAFCONV_IMS_WINDOWS", path)
        s = saf.producers.ImsShoebox(synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL, 125.0, 7, 343.0, 48e3)
        sids = [s.add_source(q) for q in pos]
        rid = s.add_receiver_sh(7, [8.8, 5.5, 0.9])
        s.compute_echograms(-1, 2.0); s.render_rirs(0)
        out[path] = ([s.num_images(rid, k) for k in sids], [s.rir(rid, k) for k in sids])
        s.destroy()
    assert out["1"][0] == out["0"][0] and sum(out["1"][0]) > 25_000_000
    for a, b in zip(out["1"][1], out["0"][1]):
        assert a.shape == b.shape == (64, 96001)
        assert np.allclose(a, b, rtol=3e-7, atol=0) and (a == b).mean() > 0.99


def test_ims_straight_into_a_convolver(saf, orc):
    """safconv_ims_create_matrixConv: 3 sources -> 9 SH channels, bank assembled on the device == convolver made from the
    downloaded RIRs"""
    P = saf.producers
    room, aw = [6.0, 5.0, 2.8], synth.IMS_TEST_ABS_WALL
    s = P.ImsShoebox(room, aw, 125.0, 7, 343.0, 48e3)
    srcs = [[1.2, 3.3, 1.1], [4.0, 4.1, 2.0], [2.5, 0.7, 0.4]]
    sids = [s.add_source(p) for p in srcs]
    s.remove_source(sids[1]); sids[1] = s.add_source(srcs[1])        # slot order stays source order
    rid = s.add_receiver_sh(2, [4.4, 1.7, 1.6])
    s.compute_echograms(-1, 0.05); s.render_rirs(0)
    rirs = [s.rir(rid, k) for k in sids]
    L = max(r.shape[1] for r in rirs)
    bank = np.zeros((9, 3, L), np.float32)
    for k, r in enumerate(rirs):
        bank[:, k, :r.shape[1]] = r
    hop = 256
    h = s.matrixconv(rid, hop)
    mc = saf.MatrixConv(hop, bank, 1)
    x = np.random.default_rng(4).uniform(-1, 1, (3, hop * 12)).astype(np.float32)
    y = np.concatenate([P.apply_raw(h, x[:, b * hop:(b + 1) * hop], 9) for b in range(12)], 1)
    assert np.array_equal(y, mc.run(x))
    ma, l2 = err_metrics(y, orc.OracleMatrixConv(hop, bank, 1).run(x))
    assert ma <= TOL_MAXABS_FS and l2 <= TOL_REL_L2
    P.destroy_raw(h); mc.destroy(); s.destroy()
