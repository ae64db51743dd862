"""Where does the per-block latency of the small real-time configs go?  (benchmark tooling)"""
import sys, time, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import spatial_audio_framework_b200 as saf
import bench

def p50(f, n=2000, warm=200):
    for _ in range(warm): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.percentile(ts, 50)), 1e6 * float(np.percentile(ts, 99))

lib = saf.lib()
fp = C.POINTER(C.c_float)
for wn in sys.argv[1:] or ["C1", "C2"]:
    w = bench.WORKLOADS[wn]
    H = bench.filters_for(w, 0, w["nOut"])
    for fused in (1, 0):
        mc = saf.MatrixConv(w["hop"], H)
        mc.set_option("small_fused", fused)
        xin = torch.rand((w["nIn"], w["hop"])).pin_memory(); yout = torch.empty((w["nOut"], w["hop"])).pin_memory()
        xp, yp = C.cast(xin.data_ptr(), fp), C.cast(yout.data_ptr(), fp)
        host = p50(lambda: lib.saf_matrixConv_apply(mc.handle, xp, yp))
        xn = np.random.rand(w["nIn"], w["hop"]).astype(np.float32); yn = np.empty((w["nOut"], w["hop"]), np.float32)
        xnp, ynp = xn.ctypes.data_as(fp), yn.ctypes.data_as(fp)
        pageable = p50(lambda: lib.saf_matrixConv_apply(mc.handle, xnp, ynp))
        xd = xin.cuda(); yd = torch.empty((w["nOut"], w["hop"]), device="cuda")
        def dev():
            lib.safconv_apply_device_blocks(mc.handle, C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()), 1)
            lib.safconv_synchronize(mc.handle)
        devlat = p50(dev)
        # kernel-only: CUDA events on the handle's stream around 200 back-to-back single-block launches
        st = torch.cuda.Stream(); mc.set_stream(st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record()
            for _ in range(200):
                lib.safconv_apply_device_blocks(mc.handle, C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()), 1)
            e1.record()
        torch.cuda.synchronize()
        print(f"{wn} fused={fused}: host-API pinned p50/p99 {host[0]:.1f}/{host[1]:.1f} us | pageable {pageable[0]:.1f}/{pageable[1]:.1f} us | "
              f"device-ptr + sync {devlat[0]:.1f}/{devlat[1]:.1f} us | back-to-back device launches {1e3 * e0.elapsed_time(e1) / 200:.1f} us each")
        mc.destroy()
# floor: empty-stream synchronise and a trivial torch kernel + sync
x = torch.zeros(32, device="cuda")
print("torch add_ + synchronize p50/p99 us:", p50(lambda: (x.add_(1), torch.cuda.synchronize())))

# TVConv (reference saf_utility_matrixConv.h:157-190): 1 input, nOut outputs, IR set switched every few blocks
hop, L, nIRs, nOut = 512, 24000, 16, 4
Htv = (np.random.default_rng(1).uniform(-1, 1, (nIRs, nOut, L)) * np.exp(-6.9 * np.arange(L) / L)).astype(np.float32)
tv = saf.TVConv(hop, Htv, 0)
xtv = np.random.rand(hop).astype(np.float32)
state = {"i": 0}
def tv_call():
    state["i"] += 1
    tv.apply(xtv, (state["i"] // 7) % nIRs)
print("TVConv 1 x %d, hop %d, %d taps, %d IR sets: host-API p50/p99 us:" % (nOut, hop, L, nIRs), p50(tv_call))
