"""Device timeline of the look-ahead apply (SAFCONV_TRACE=1): back-to-back calls on the C4 workload.  Debug tooling."""
import os, sys, time, ctypes as C
os.environ["SAFCONV_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spatial_audio_framework_b200 as saf
import bench
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C4"]
H = bench.filters_for(w, 0, w["nOut"])
mc = saf.MatrixConv(w["hop"], H); del H
lib = saf.lib(); fp = C.POINTER(C.c_float)
xin = torch.rand((w["nIn"], w["hop"])).pin_memory(); yout = torch.empty((w["nOut"], w["hop"])).pin_memory()
xp, yp = C.cast(xin.data_ptr(), fp), C.cast(yout.data_ptr(), fp)
ts = []
pace = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0      # seconds between calls (0 = back to back)
for i in range(14):
    if pace: time.sleep(pace)
    t0 = time.perf_counter(); lib.saf_matrixConv_apply(mc.handle, xp, yp); ts.append(1e6 * (time.perf_counter() - t0))
print("host us per call:", [round(t) for t in ts])
