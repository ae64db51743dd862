"""Device timeline of back-to-back look-ahead calls (SAFCONV_TIMELINE=n).  Debug tooling.  usage: timeline.py WORKLOAD [n]"""
import os, sys, time, ctypes as C
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
os.environ["SAFCONV_TIMELINE"] = str(n)
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
f = lib.saf_matrixConv_apply
for i in range(n + 4):
    f(mc.handle, xp, yp)
