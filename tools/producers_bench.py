"""Timing of the filter producers on the GPU box (not part of bench.py's contract line; `python bench.py` carries a bounded
version of the same figures in secondary.producers).

    python tools/producers_bench.py [--sources 64] [--order 7] [--max-time 2.0] [--cpu]

* decoder: getBinauralAmbiDecoderFilters, order 7 (64 SH channels), 836 directions, fftSize 1024, every built method
* ims:     ims_shoebox_renderRIRs of `--sources` sources x one SH receiver, `--max-time` seconds at 48 kHz
           (64 sources, order 7, 2.0 s = the 64 x 64 x 96 000 bank of BASELINE.json configs[3]), then
           safconv_ims_create_matrixConv (bank -> convolver without leaving the device) and one block through it
* --cpu:   the compiled reference (oracle/_ref/libsaf_ref_producers.so) on ONE core on a bounded sample of the same work
Wall-clock around the synchronous C calls (each ends with a stream synchronisation).
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import spatial_audio_framework_b200 as saf  # noqa: E402
from spatial_audio_framework_b200 import synth  # noqa: E402

P = saf.producers


def timed(f, n=3):
    best = 1e30
    out = None
    for _ in range(n):
        t = time.perf_counter(); out = f(); best = min(best, time.perf_counter() - t)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sources", type=int, default=64)
    ap.add_argument("--order", type=int, default=7)
    ap.add_argument("--max-time", type=float, default=2.0)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    res = {}

    # ---- decoder design ----
    H, d, itd = synth.synthetic_hrtfs(836, 1024, 48000.0)
    dec = {}
    P.decoder_filters(H, d, 1024, 48000.0, P.DECODER_LS, 1)          # context / module load
    for name, m, dc, mr in (("LS", P.DECODER_LS, 0, 0), ("LSDIFFEQ", P.DECODER_LSDIFFEQ, 0, 0), ("TA", P.DECODER_TA, 0, 0),
                            ("MAGLS", P.DECODER_MAGLS, 0, 0), ("MAGLS+diffCM+maxRE", P.DECODER_MAGLS, 1, 1)):
        t, _ = timed(lambda: P.decoder_filters(H, d, 1024, 48000.0, m, 7, itd, None, dc, mr))
        dec[name] = {"gpu_ms": round(t * 1e3, 3)}
    if a.cpu:
        from oracle import producers as PR
        R = PR.load_producers_reference()
        for name, m, dc, mr in (("LS", PR.LS, 0, 0), ("MAGLS+diffCM+maxRE", PR.MAGLS, 1, 1)):
            t, _ = timed(lambda: R.decoder_filters(H, d, 1024, 48000.0, m, 7, itd, None, dc, mr), 1)
            dec[name]["reference_1core_ms"] = round(t * 1e3, 1)
    res["decoder"] = {"workload": "order 7, 836 directions, fftSize 1024 (513 bands)", **dec}

    # ---- image sources ----
    room, aw = synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL
    rng = np.random.default_rng(0)
    s = P.ImsShoebox(room, aw, 125.0, 7, 343.0, 48e3)
    sids = [s.add_source([rng.uniform(0.5, 9.5), rng.uniform(0.5, 6.5), rng.uniform(0.5, 2.5)]) for _ in range(a.sources)]
    rid = s.add_receiver_sh(a.order, [8.8, 5.5, 0.9])
    s.compute_echograms(-1, 0.05); s.render_rirs(0)               # warm-up (module load)
    t0 = time.perf_counter()
    s.compute_echograms(-1, a.max_time); s.render_rirs(0)
    t_render = time.perf_counter() - t0
    images = sum(s.num_images(rid, k) for k in sids)
    r0 = s.rir(rid, sids[0])
    nch = (a.order + 1) ** 2
    t0 = time.perf_counter()
    h = s.matrixconv(rid, 1024)
    t_conv = time.perf_counter() - t0
    x = rng.uniform(-1, 1, (a.sources, 1024)).astype(np.float32)
    y = P.apply_raw(h, x, nch)
    # first block of the convolver == the first 1024 taps of the bank applied to the block's first samples (spot check)
    want = sum(np.convolve(x[k], s.rir(rid, sids[k])[0][:1024])[:1024] for k in range(min(a.sources, 4)))
    if a.sources <= 4:
        assert np.abs(y[0] - want).max() <= 1e-4 * max(1.0, np.abs(want).max())
    P.destroy_raw(h)
    ims = {"workload": f"{a.sources} sources x 1 receiver of SH order {a.order} ({nch} ch), {a.max_time} s at 48 kHz, 7 bands, room 10 x 7 x 3 m",
           "rir_shape": list(r0.shape), "image_sources": int(images), "gpu_render_s": round(t_render, 4),
           "images_per_s": round(images / t_render, 1), "taps_per_s": round(images * nch / t_render, 1),
           "bank_to_convolver_s": round(t_conv, 4)}
    if a.cpu:
        from oracle import producers as PR
        R = PR.load_producers_reference()
        rs = R.ims(room, aw, 125.0, 7, 343.0, 48e3)
        sid = rs.add_source([5.1, 6.0, 1.1]); rrid = rs.add_receiver_sh(a.order, [8.8, 5.5, 0.9])
        tcpu = min(0.4, a.max_time)
        t0 = time.perf_counter(); rs.compute_echograms(-1, tcpu); rs.render_rirs(0); t_ref = time.perf_counter() - t0
        n_ref = rs.echogram_times(rrid, sid).size
        rs.destroy()
        ims["reference_1core"] = {"sample": f"1 source, {tcpu} s", "image_sources": int(n_ref), "s": round(t_ref, 3),
                                  "images_per_s": round(n_ref / t_ref, 1)}
    s.destroy()
    res["ims"] = ims
    line = json.dumps(res)
    print(line)
    if a.out:
        Path(a.out).write_text(line + "\n")


if __name__ == "__main__":
    main()
