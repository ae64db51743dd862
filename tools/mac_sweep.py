"""Sweep the MAC kernel's pipeline knobs (SAFCONV_MAC_STAGES x SAFCONV_MAC_STAGE_KB) on a workload and print
the CUDA-event duration of K2 and the achieved H+FDL GB/s.  Benchmark tooling, not part of the product path."""
import os, sys, json, itertools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import spatial_audio_framework_b200 as saf
import bench

wname = sys.argv[1] if len(sys.argv) > 1 else "C4"
nout = int(sys.argv[2]) if len(sys.argv) > 2 else None
w = dict(bench.WORKLOADS[wname])
if nout:
    w["nOut"] = nout
H = bench.filters_for(w, 0, w["nOut"])
B = 16
x = torch.rand((B, w["nIn"], w["hop"]), device="cuda") * 2 - 1
y = torch.empty((B, w["nOut"], w["hop"]), device="cuda")
combos = [(4, 32), (6, 32), (3, 64), (8, 16), (12, 16), (2, 96), (5, 40), (4, 48), (16, 8), (6, 16)]
for ns, kb in combos:
    os.environ["SAFCONV_MAC_STAGES"] = str(ns)
    os.environ["SAFCONV_MAC_STAGE_KB"] = str(kb)
    try:
        c = saf.MatrixConv(w["hop"], H, 1)
    except Exception as e:
        print(ns, kb, "create failed:", str(e)[:80]); continue
    info = c.info()
    for _ in range(2):
        c.apply_device(x.data_ptr(), y.data_ptr(), B)
    c.synchronize()
    c.enable_kernel_timing(3 * B)
    for _ in range(3):
        c.apply_device(x.data_ptr(), y.data_ptr(), B)
    c.synchronize()
    ms, n = c.kernel_times_ms()
    print(json.dumps(dict(workload=wname, nOut=w["nOut"], stages=ns, stage_kb=kb, mac_ms=ms[1], fft_ms=ms[0], ifft_ms=ms[2],
                          GBps=info.macAlgBytesPerBlock / ms[1] / 1e6)))
    c.destroy()
