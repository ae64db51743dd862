"""Accuracy / timing of the offline tensor-core path vs the streaming path for several promotion intervals
(SAFCONV_OFF_FLUSH).  Benchmark tooling, not part of the product path."""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import spatial_audio_framework_b200 as saf
from spatial_audio_framework_b200 import synth
import oracle as O

hop, L, nIn, nOut, T = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (1024, 8192, 121, 64, 300))]
H = synth.decaying_rir((nOut, nIn, L), seed=5)
x = synth.uniform((nIn, hop * T), seed=6)
outs = np.arange(min(nOut, 2))
n0, n1 = hop * (T - 2), hop * T
truth = O.truth_matrix(H, x, outs, n0, n1)
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
ys = None
for fl in (1, 2, 4, 8, 100000):
    os.environ["SAFCONV_OFF_FLUSH"] = str(fl)
    mc = saf.MatrixConv(hop, H)
    if ys is None:
        ys = mc.run(x)
    y = mc.render_offline(x)
    y = mc.render_offline(x)
    ms = mc.offline_times_ms()
    print(json.dumps(dict(flush=fl, offline_vs_stream=rel(y, ys), offline_vs_truth=rel(y[outs][:, n0:n1], truth),
                          stream_vs_truth=rel(ys[outs][:, n0:n1], truth), ms_fft=ms[0], ms_gemm=ms[1], ms_ifft=ms[2])))
    mc.destroy()
