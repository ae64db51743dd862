"""Generate tests/golden/producers_*.npz from the COMPILED, UNMODIFIED reference producers
(oracle/_ref/libsaf_ref_producers.so, built by `make -C oracle ref` in the build container):

    python tests/golden/make_golden_producers.py

* producers_sh.npz       getRSH / getSHreal_recur / getMaxREweights values
                         (/root/reference/framework/modules/saf_hoa/saf_hoa.c:118-150, 235-266; saf_sh.c:255-330)
* producers_decoder.npz  getBinauralAmbiDecoderFilters (saf_hoa.c:452-497) on a seeded synthetic HRTF set, every
                         method x {plain, max-rE, covariance matching, both} (SPR: outputs only -- the t-design it projects on
                         is fetched from the compiled reference at test time, RefProducers.tdesign)
* producers_ims.npz      ims_shoebox_* (saf_reverb.c): the scene and the add / remove / move sequence of the reference's
                         own unit test (test/src/test__reverb_module.c:27-96, which asserts nothing), plus time-limited and
                         order-limited echograms at SH orders 0, 3, 7
The inputs are stored next to the outputs, so the fixtures also serve machines without the reference tree (the GPU box).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import producers as PR  # noqa: E402
from spatial_audio_framework_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def ims_unit_test_sequence(make_scene):
    """test__reverb_module.c:27-96 against any scene implementation; returns (ids, {(rid, sid): rir})."""
    s = make_scene(synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL, 125.0, 7, 343.0, 48e3)
    src = {1: [5.1, 6.0, 1.1], 2: [2.1, 1.0, 1.3], 3: [4.4, 3.0, 1.4], 4: [6.4, 4.0, 1.3], 5: [8.5, 5.0, 1.8]}
    ids = []
    id1 = s.add_source(src[1]); id2 = s.add_source(src[2]); rid = s.add_receiver_sh(3, [8.8, 5.5, 0.9])
    ids += [id1, id2, rid]
    s.remove_source(0)
    id3 = s.add_source(src[3]); id4 = s.add_source(src[4]); id5 = s.add_source(src[5])
    ids += [id3, id4, id5]
    s.remove_source(id3); s.remove_source(id4)
    id4 = s.add_source(src[4])
    ids.append(id4)
    mov_s = list(src[1]); mov_r = [8.8, 5.5, 0.9]
    for i in range(10):
        mov_s[1] = np.float32(2.0) + np.float32(i) / np.float32(10.0)
        mov_r[0] = np.float32(3.0) + np.float32(i) / np.float32(10.0)
        s.update_source(id4, mov_s); s.update_receiver(rid, mov_r)
        s.compute_echograms(-1, 0.05)
        s.render_rirs(0)
    active = sorted({id2, id4, id5})
    rirs = {(rid, sid): s.rir(rid, sid) for sid in active}
    s.destroy()
    return ids, rirs


IMS_CASES = [
    # name,        order, maxN, maxTime_s, nBands, src,               rec
    ("t_o0",       0,     -1,   0.10,      7,      [5.1, 6.0, 1.1],   [8.8, 5.5, 1.0]),
    ("t_o3",       3,     -1,   0.08,      7,      [2.1, 1.0, 1.3],   [8.8, 5.5, 0.9]),
    ("t_o7",       7,     -1,   0.04,      2,      [4.4, 3.0, 1.4],   [3.3, 2.5, 1.7]),
    ("n_o2",       2,     4,    -1.0,      7,      [6.4, 4.0, 1.3],   [1.0, 6.5, 2.0]),
    ("n_o5_direct", 5,    0,    -1.0,      7,      [8.5, 5.0, 1.8],   [8.8, 5.5, 0.9]),
]


def main():
    R = PR.load_producers_reference()
    rng = np.random.default_rng(7)
    # --- SH ---
    nD = 200
    dirs_deg = np.stack([rng.uniform(-180, 180, nD), np.degrees(np.arcsin(rng.uniform(-1, 1, nD)))], 1).astype(np.float32)
    dirs_deg[:4] = [[0, 90], [0, -90], [45, 0], [-180, 30]]           # poles and axis points
    dirs_rad = np.stack([np.radians(dirs_deg[:, 0]), np.pi / 2 - np.radians(dirs_deg[:, 1])], 1).astype(np.float32)
    np.savez_compressed(OUT / "producers_sh.npz", kind="producers_sh", dirs_deg=dirs_deg, dirs_rad=dirs_rad,
                        rsh10=R.rsh(10, dirs_deg), rsh3=R.rsh(3, dirs_deg),
                        recur7=R.shreal_recur(7, dirs_rad), recur10=R.shreal_recur(10, dirs_rad), recur2=R.shreal_recur(2, dirs_rad),
                        maxre=np.stack([np.pad(R.maxre(o), (0, 121 - (o + 1) ** 2)) for o in range(11)]))
    # --- decoder ---
    fftSize, fs, order, n_dirs = 128, 48000.0, 3, 146
    H, dirs, itd = synth.synthetic_hrtfs(n_dirs, fftSize, fs)
    w = (np.full(n_dirs, 1.0 / n_dirs) * (1.0 + 0.05 * np.cos(np.radians(dirs[:, 1])))).astype(np.float32)
    w /= w.sum()
    out = {}
    for m in (PR.DEFAULT, PR.LS, PR.LSDIFFEQ, PR.TA, PR.MAGLS):
        for dc in (0, 1):
            for mr in (0, 1):
                out[f"f_m{m}_dc{dc}_mr{mr}"] = R.decoder_filters(H, dirs, fftSize, fs, m, order, itd, None, dc, mr)
    for dc in (0, 1):
        for mr in (0, 1):
            out[f"f_m3_dc{dc}_mr{mr}"] = R.decoder_filters(H, dirs, fftSize, fs, PR.SPR, order, itd, None, dc, mr)
    out["f_m3_o1_weights"] = R.decoder_filters(H, dirs, fftSize, fs, PR.SPR, 1, itd, w * np.float32(4 * np.pi), 0, 0)
    out["f_m1_weights"] = R.decoder_filters(H, dirs, fftSize, fs, PR.LS, order, itd, w, 1, 1)
    out["f_m5_o1"] = R.decoder_filters(H, dirs, fftSize, fs, PR.MAGLS, 1, itd, None, 0, 0)
    out["f_m2_o5"] = R.decoder_filters(H, dirs, fftSize, fs, PR.LSDIFFEQ, 5, itd, None, 0, 1)
    np.savez_compressed(OUT / "producers_decoder.npz", kind="producers_decoder", hrtfs=H, dirs_deg=dirs, itd_s=itd, weights=w,
                        fftSize=fftSize, fs=fs, order=order, **out)
    # --- image sources ---
    ims = {}
    ids, rirs = ims_unit_test_sequence(R.ims)
    ims["ut_ids"] = np.array(ids, np.int32)
    for (rid, sid), r in rirs.items():
        ims[f"ut_rir_r{rid}_s{sid}"] = r
    for name, order, maxN, maxT, nB, src, rec in IMS_CASES:
        s = R.ims(synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL[:nB], 125.0, nB, 343.0, 48e3)
        sid = s.add_source(src); rid = s.add_receiver_sh(order, rec)
        s.compute_echograms(maxN, maxT); s.render_rirs(0)
        ims[f"{name}_rir"] = s.rir(rid, sid)
        ims[f"{name}_times"] = s.echogram_times(rid, sid)
        s.destroy()
    np.savez_compressed(OUT / "producers_ims.npz", kind="producers_ims", **ims)
    for f in ("producers_sh.npz", "producers_decoder.npz", "producers_ims.npz"):
        print(f, (OUT / f).stat().st_size)


if __name__ == "__main__":
    main()
