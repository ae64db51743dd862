#!/usr/bin/env python
"""bench.py -- headline benchmark of the SAF multichannel-convolution hot path on B200.

Workload (BASELINE.json `metric` is quoted on configs[3]): saf_matrixConv 64-in x 64-out, hop 1024,
96000-tap synthetic (exponentially decaying noise) RIRs, output channels sharded over N GPUs.
One "step" = `--blocks` consecutive hop-sized blocks through the hot path (K1 input FFT ->
K2 filter-streaming MAC -> K3 inverse FFT + overlap-add).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU code (oracle/_ref) on the host cores

Prints ONE JSON line (rank 0).  `value` = output-channel*samples per second, whole job, inputs resident
in HBM; `e2e` = the same through the reference-facing host-pointer API (H2D / D2H inside the timed region);
`roofline` = the MAC kernel's algorithmic (H + delay line) bytes / its CUDA-event duration against the
measured HBM copy peak; `cpu_baseline` = the compiled reference timed on this box's host cores on a
bounded sample of the same workload.

PyTorch is used for plumbing only (device buffers, events, torch.distributed/NCCL); every kernel on the
path is ours (libsafconv_b200.so).  oracle/ is touched only by the cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: kind, nIn, nOut, hop, L
    "C4": dict(kind="matrix", nIn=64, nOut=64, hop=1024, L=96000,
               desc="saf_matrixConv 64-in x 64-out, hop 1024, 96000-tap RIRs (BASELINE.json configs[3])"),
    "C4s": dict(kind="matrix", nIn=64, nOut=64, hop=1024, L=8192,
                desc="C4 scaled down to 8192 taps (debug only)"),
    "C4g8": dict(kind="matrix", nIn=64, nOut=8, hop=1024, L=96000,
                 desc="one GPU's share of configs[3] at 8 GPUs: 64-in x 8-out, hop 1024, 96000 taps (debug: per-block call overhead at 62 us of filter stream)"),
    "UT": dict(kind="matrix", nIn=32, nOut=40, hop=2048, L=512,
               desc="test__saf_matrixConv shape 32x40, hop 2048, 512 taps"),
    "C3": dict(kind="multi", nIn=256, nOut=256, hop=512, L=4096,
               desc="saf_multiConv 256 ch, hop 512, 4096 taps (BASELINE.json configs[2])"),
    "C5": dict(kind="offline", nIn=121, nOut=64, hop=1024, L=8192, seconds=60.0,
               desc="offline batched render: 121 SH x 64 out, 8192 taps, hop 1024, 60 s of audio in one call (BASELINE.json configs[4])"),
    "C4o": dict(kind="offline", nIn=64, nOut=64, hop=1024, L=96000, seconds=60.0,
                desc="configs[3] filters (64x64, 96000 taps) rendered offline: 60 s of audio in one call through the tensor-core path"),
    "C5s": dict(kind="offline", nIn=121, nOut=64, hop=1024, L=8192, seconds=6.0,
                desc="C5 on a 6 s signal (debug)"),
    "C2": dict(kind="matrix", nIn=25, nOut=2, hop=128, L=512,
               desc="saf_matrixConv 25x2, hop 128, 512 taps (BASELINE.json configs[1])"),
    "C1": dict(kind="matrix", nIn=4, nOut=2, hop=256, L=1024,
               desc="saf_matrixConv 4x2, hop 256, 1024 taps (BASELINE.json configs[0])"),
}
METRIC = "convolved out-ch*samples/sec"
UNIT = "out-ch*samples/s"


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def filters_for(w, out_begin, out_count, seed=0x5AF0C0DE):
    """Deterministic per-output-channel filters so that every rank can build its own shard."""
    from spatial_audio_framework_b200 import synth
    if w["kind"] in ("matrix", "offline"):
        H = np.empty((out_count, w["nIn"], w["L"]), np.float32)
        for i in range(out_count):
            H[i] = synth.decaying_rir((w["nIn"], w["L"]), seed=seed + out_begin + i)
    else:
        H = np.empty((out_count, w["L"]), np.float32)
        for i in range(out_count):
            H[i] = synth.decaying_rir((w["L"],), seed=seed + out_begin + i)
    return H


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def pct(a, q):
    return float(np.percentile(a, q)) if len(a) else None


def call_loop(fn, hdl, xin, yout, n, in_stride=0, out_stride=0, nbuf=1, period=None):
    """n synchronous drop-in calls fn(hdl, in, out) on ctypes float pointers (buffer b % nbuf of a ring of blocks);
    period != None paces the calls like an audio callback.  Returns (per-call ms list, total seconds)."""
    import ctypes as C
    fp = C.POINTER(C.c_float)
    ins = [C.cast(xin + (b % nbuf) * in_stride, fp) for b in range(nbuf)]
    outs = [C.cast(yout + (b % nbuf) * out_stride, fp) for b in range(nbuf)]
    lat = []
    pc = time.perf_counter
    t_begin = pc()
    t_next = t_begin + (period or 0.0)
    for k in range(n):
        if period:
            while pc() < t_next:
                pass
            t_next += period
        t0 = pc()
        fn(hdl, ins[k % nbuf], outs[k % nbuf])
        lat.append(1e3 * (pc() - t0))
    return lat, pc() - t_begin


def host_api_latency(conv, kind, x_host, y_host, nbuf, in_elems, out_elems, hop, blocks=2000, warm=100, paced_calls=200):
    """p50 / p99 of the synchronous host-pointer call (>= 2000 blocks after 100 warm-up), with the caller's
    buffers page-locked and pageable, and paced (GPU idle between calls)."""
    lib, hdl = conv._lib, conv.handle
    fn = lib.saf_matrixConv_apply if kind == "matrix" else lib.saf_multiConv_apply
    out = {}
    call_loop(fn, hdl, x_host.data_ptr(), y_host.data_ptr(), warm, in_elems * 4, out_elems * 4, nbuf)
    lat, dt = call_loop(fn, hdl, x_host.data_ptr(), y_host.data_ptr(), blocks, in_elems * 4, out_elems * 4, nbuf)
    out["blocks"] = blocks
    out["warmup_blocks"] = warm
    out["pinned"] = {"p50_ms": pct(lat, 50), "p99_ms": pct(lat, 99), "blocks_per_s": blocks / dt}
    # pageable caller buffers (the reference's hosts pass malloc'd frames: matrixconv.c:137-149)
    xp = np.array(x_host.numpy()[0:1]).reshape(-1).copy()
    yp = np.empty(out_elems, np.float32)
    call_loop(fn, hdl, xp.ctypes.data, yp.ctypes.data, warm)
    lat, dt = call_loop(fn, hdl, xp.ctypes.data, yp.ctypes.data, blocks)
    out["pageable"] = {"p50_ms": pct(lat, 50), "p99_ms": pct(lat, 99), "blocks_per_s": blocks / dt}
    if paced_calls:
        period = min(hop / 48000.0, 0.005)
        lat, _ = call_loop(fn, hdl, x_host.data_ptr(), y_host.data_ptr(), paced_calls + 20, in_elems * 4, out_elems * 4, nbuf, period)
        lat = lat[20:]
        out["paced"] = {"p50_ms": pct(lat, 50), "p99_ms": pct(lat, 99), "period_ms": 1e3 * period, "calls": paced_calls,
                        "note": "real-time block period hop/48000 s, capped at 5 ms (the GPU is idle long before)"}
    if kind == "matrix" and hop <= 1024:
        # opt-in resident latency kernel (no launch per call: the host rings a doorbell in page-locked memory); only small
        # matrix problems have one -- for anything else the option changes nothing
        conv.set_option("resident_us", 20000)
        call_loop(fn, hdl, x_host.data_ptr(), y_host.data_ptr(), warm, in_elems * 4, out_elems * 4, nbuf)
        lat, dt = call_loop(fn, hdl, x_host.data_ptr(), y_host.data_ptr(), blocks, in_elems * 4, out_elems * 4, nbuf)
        out["resident"] = {"p50_ms": pct(lat, 50), "p99_ms": pct(lat, 99), "blocks_per_s": blocks / dt, "idle_timeout_us": 20000,
                           "note": "safconv_set_option(h, \"resident_us\", 20000): same synchronous saf_matrixConv_apply, page-locked buffers"}
        if paced_calls:
            period = min(hop / 48000.0, 0.005)
            lat, _ = call_loop(fn, hdl, x_host.data_ptr(), y_host.data_ptr(), paced_calls + 20, in_elems * 4, out_elems * 4, nbuf, period)
            lat = lat[20:]
            out["resident"]["paced_p50_ms"] = pct(lat, 50)
            out["resident"]["paced_p99_ms"] = pct(lat, 99)
        conv.set_option("resident_us", 0)
    return out


def parity_check(conv, w, B, dev, blocks=6):
    """--check: the batched device path that `value` times (safconv_apply_device_blocks from a reset state) against the
    oracle on output channel 0, same seeded input.  Only the first `blocks` blocks are compared (the CPU checker needs
    ~0.2 s per block and output channel at C4)."""
    import torch
    import oracle as O
    hop, nIn = w["hop"], w["nIn"]
    rng = np.random.default_rng(4321)
    x = rng.uniform(-1, 1, (B, nIn, hop)).astype(np.float32)
    conv.synchronize()
    conv.reset_state()
    d_in = torch.from_numpy(x).to(dev)
    d_out = torch.empty((B, conv.nOutLocal, hop), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    conv.apply_device(d_in.data_ptr(), d_out.data_ptr(), B)
    conv.synchronize()
    y = d_out[:blocks, 0].cpu().numpy().reshape(-1)
    H0 = filters_for(w, 0, 1)
    if w["kind"] == "matrix":
        ref = O.OracleMatrixConv(hop, H0, 1)
        r = np.concatenate([ref.apply(x[b])[0] for b in range(blocks)])
    else:
        ref = O.OracleMultiConv(hop, H0, 1)
        r = np.concatenate([ref.apply(np.ascontiguousarray(x[b, 0:1]))[0] for b in range(blocks)])
    fs = float(np.abs(r).max())
    conv.reset_state()
    return {"parity_rel_l2": float(np.linalg.norm(y - r) / np.linalg.norm(r)),
            "parity_max_abs_fs": float(np.abs(y - r).max() / fs),
            "what": f"output channel 0, first {blocks} of {B} blocks of one batched step from a reset state vs the CPU oracle (same input)",
            "tolerance": "rel L2 <= 1e-6, max abs <= 1e-5 of full scale"}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference (CPU) leg: the compiled, unmodified reference convolver (oracle/_ref), one handle per thread,
# each owning a disjoint slice of output channels of the same problem
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(w, steps, warmup, threads=None, blocks_per_step=None, ch_per_thread=1):
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    import oracle as O
    try:
        O.load_reference()
        kind = "reference"
        mk_matrix, mk_multi = O.RefMatrixConv, O.RefMultiConv
    except Exception:
        kind = "port"
        mk_matrix, mk_multi = O.OracleMatrixConv, O.OracleMultiConv
    cores = threads or os.cpu_count() or 1
    total_out = w["nOut"]
    cores = max(1, min(cores, total_out // ch_per_thread))
    hop, nIn = w["hop"], w["nIn"]
    convs, xs = [], []
    rng = np.random.default_rng(1)
    for t in range(cores):
        H = filters_for(w, t * ch_per_thread, ch_per_thread)
        if w["kind"] in ("matrix", "offline"):
            convs.append(mk_matrix(hop, H, 1))
            xs.append(rng.uniform(-1, 1, (nIn, hop)).astype(np.float32))
        else:
            convs.append(mk_multi(hop, H, 1))
            xs.append(rng.uniform(-1, 1, (ch_per_thread, hop)).astype(np.float32))

    def one_step():
        def work(i):
            for _ in range(blocks_per_step or 1):
                convs[i].apply(xs[i])
        th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()

    # bound the sample: blocks per step chosen so that one step is ~0.25 s of CPU work (1 block for C4)
    if blocks_per_step is None:
        blocks_per_step = 1
        t0 = time.perf_counter()
        one_step()
        t1 = time.perf_counter() - t0
        blocks_per_step = int(max(1, min(4096, 0.25 / max(t1, 1e-6))))
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    units = cores * ch_per_thread * hop * blocks_per_step * steps
    sample = (f"{cores} threads x {ch_per_thread} of {total_out} output channels each (independent per reference .c:218-234), "
              f"full {nIn}-input x {w['L']}-tap filters, {blocks_per_step} block(s)/step x {steps} steps after {warmup} warm-up; "
              f"{'compiled unmodified reference (KissFFT + ' + O.load_reference()[1] + ' level-1 BLAS)' if kind == 'reference' else 'oracle port'}")
    return dict(value=units / dt, unit=UNIT, cores=cores, kind=kind, sample=sample, seconds=dt,
                ms_per_step=1e3 * dt / steps)


def workload_config(w, world):
    """`config` of the JSON line: what defines the workload, identical in both arms (the driver compares the two dicts);
    everything that describes HOW an arm ran it goes to `config_detail`."""
    hop, L = w["hop"], w["L"]
    P = int(np.ceil(np.float32(L) / np.float32(hop)))
    M = 32
    while M < hop:
        M *= 2
    per_gpu_mb = ((w["nOut"] + world - 1) // world) * w["nIn"] * P * M * 8 / 1e6
    return {"workload": w["desc"], "nIn": w["nIn"], "nOut": w["nOut"], "hop": hop, "length_h": L, "partitions": P,
            "l2": "inputs larger than L2: %.0f MB of filter spectra streamed per block per GPU (L2 = 126 MB)" % per_gpu_mb
                  if per_gpu_mb > 126 else "filter spectra (%.1f MB per GPU) fit in L2: L2 is not flushed between blocks, the real-time use keeps them there" % per_gpu_mb,
            "filters": "exponentially decaying uniform noise (-60 dB at the last tap), seeded per output channel"}


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(w, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, max(1, args.gpus)),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm, offline workload (configs[4]): one step = one safconv_render_offline of the whole signal
# ------------------------------------------------------------------------------------------------
def run_offline_arm(args, w):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure_offline(args, w, world, rank, local, dist)
    if line is not None:
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def measure_offline(args, w, world, rank, local, dist):
    """One offline workload (configs[4]) on this rank; returns the JSON line on rank 0, None elsewhere."""
    import torch
    import spatial_audio_framework_b200 as saf
    from spatial_audio_framework_b200 import sharding
    dev = torch.device("cuda", local)
    hop, nIn, nOut = w["hop"], w["nIn"], w["nOut"]
    T = int(np.ceil(w["seconds"] * 48000.0 / hop))
    P = int(np.ceil(np.float32(w["L"]) / np.float32(hop)))
    # offline rendering shards the signal in TIME (each rank: all output channels of its own stretch of audio,
    # plus a P-frame input halo); nothing is exchanged between GPUs
    t0, t1, halo = sharding.time_segment(T, world, rank, P)
    Tr = t1 - t0
    H = filters_for(w, 0, nOut)
    conv = saf.MatrixConv(hop, H, 1, device=local)
    del H
    info = conv.info()
    stream = torch.cuda.Stream(device=dev)
    conv.set_stream(stream.cuda_stream)
    g = torch.Generator(device="cpu").manual_seed(1234)
    x_full = torch.rand((nIn, T * hop), generator=g) * 2 - 1                      # same signal on every rank
    x_host = x_full[:, (t0 - halo) * hop:t1 * hop].contiguous().pin_memory()      # this rank's stretch + halo
    del x_full
    y_host = torch.empty((nOut, Tr * hop), dtype=torch.float32).pin_memory()
    x_dev = torch.empty((nIn, (Tr + halo) * hop), dtype=torch.float32, device=dev)
    y_dev = torch.empty((nOut, Tr * hop), dtype=torch.float32, device=dev)
    x_dev.copy_(x_host)

    def step_device():
        conv.render_offline_segment_device(x_dev.data_ptr(), y_dev.data_ptr(), Tr, halo)

    def step_host():
        # the host-buffer API: page-locked input -> page-locked output, segments pipelined over three streams
        conv.render_offline_host(x_host.data_ptr(), y_host.data_ptr(), Tr, halo)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    kms = np.zeros(3)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(args.steps):
        step_device()
        kms += np.array(conv.offline_times_ms())        # syncs the stream; per-kernel events of this render
    with torch.cuda.stream(stream):
        e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    wall1 = time.time()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(wall0, wall1) if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    units = float(nOut) * T * hop * args.steps
    value = units / (ms_total * 1e-3)
    kms /= args.steps

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    step_host()
    if dist:
        dist.barrier()
    t0w = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    if dist:
        dist.barrier()
    dt = time.perf_counter() - t0w
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = {"value": float(nOut) * T * hop * e2e_steps / float(t.item()), "unit": UNIT,
           "h2d_bytes_per_step": int(nIn * (T + P * (world - 1)) * hop * 4), "d2h_bytes_per_step": int(nOut * T * hop * 4),
           "steps": e2e_steps,
           "api": "per rank: safconv_render_offline_segment on page-locked host buffers (its stretch of the signal + halo): time segments, H2D / kernels / D2H pipelined over three streams"}

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    bf16 = float(peaks.get("bf16_tflops", 1590.0))
    alg_flops = 8.0 * nOut * P * nIn * (hop + 1) * (Tr + halo)     # complex MACs as real flops, this rank's launch
    gemm_ms = float(kms[1])
    achieved = alg_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    f16 = os.environ.get("SAFCONV_OFF_KIND", "f16") != "tf32"
    tc_peak = bf16 if f16 else bf16 / 2.0
    roofline = {"bound": "tensor",
                "kernel": "offline_gemm_kernel (tcgen05 kind::%s, hi/lo split operands, 3 MMAs per product for fp32 accuracy)" % ("f16" if f16 else "tf32"),
                "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak,
                "issued_frac": 3.0 * achieved / tc_peak,
                "peak_source": ("f16 dense = the measured bf16 peak (MEASURED_PEAKS.json bf16_tflops)" if f16 else
                                "tf32 dense = half of the measured bf16 peak (MEASURED_PEAKS.json bf16_tflops)") if peaks else "fallback 1590",
                "traffic": None, "alg_flops_per_launch": alg_flops, "issued_flops_per_launch": 3.0 * alg_flops,
                "avg_launch_ms": gemm_ms, "launches_timed": args.steps, "rank": 0,
                "kernel_ms_per_render": {"forward_fft": float(kms[0]), "gemm": gemm_ms, "ifft_ola": float(kms[2])}}
    if rank != 0:
        return None
    cpu = None
    if world == 1 and not args.no_cpu:
        c = cpu_reference_run(w, steps=args.cpu_steps, warmup=1)
        cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
    conv.destroy()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (%s x3 split tensor-core products, fp32 accumulate)" % ("f16" if f16 else "tf32"), "data": "synthetic",
        "config": {"workload": w["desc"], "nIn": nIn, "nOut": nOut, "hop": hop, "length_h": w["L"], "frames": T,
                   "partitions": P,
                   "sharding": (f"time: {world} GPUs x ~{Tr} frames (+{P}-frame input halo), no exchange between GPUs"
                                if world > 1 else "single GPU"),
                   "l2": "inputs larger than L2: %.1f GB of operands per render" % ((nIn * (Tr + halo) * hop * 4 * 3 + info.bytesFilters * 2) / 1e9)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": (6 if f16 else 5) * args.steps, "roofline": roofline, "cpu_baseline": cpu,
        "realtime_factor_48k": (T * hop * args.steps / (ms_total * 1e-3)) / 48000.0,
    }
    return line


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def make_conv(saf, w, H, devices=None, device=None, shard=None):
    if w["kind"] == "matrix":
        if shard is not None:
            return saf.MatrixConv.from_shard(w["hop"], H, w["nOut"], shard, device=device)
        return saf.MatrixConv(w["hop"], H, 1, device=device, devices=devices)
    return saf.MultiConv(w["hop"], H, 1, device=device, devices=devices)


def e2e_multi_gpu(args, w, world, x_host, y_host, B):
    """N > 1, rank 0: the SAME per-block synchronous drop-in call as at N = 1, on ONE handle that spans all N GPUs
    (safconv_matrixConv_create_multi: single host process, a worker thread per device, no torch in the data path)."""
    import spatial_audio_framework_b200 as saf
    hop, nIn, nOut = w["hop"], w["nIn"], w["nOut"]
    t0 = time.perf_counter()
    H = filters_for(w, 0, nOut)
    conv = make_conv(saf, w, H, devices=list(range(world)))
    del H
    create_s = time.perf_counter() - t0
    lib, hdl = conv._lib, conv.handle
    fn = lib.saf_matrixConv_apply if w["kind"] == "matrix" else lib.saf_multiConv_apply
    xin, yout = x_host.data_ptr(), y_host.data_ptr()
    sx, sy = nIn * hop * 4, nOut * hop * 4
    blocks = max(args.e2e_blocks, B)
    res = {}
    for name, tr in (("host", 0), ("nccl", 1)):
        try:
            conv.set_option("transport", tr)
        except Exception as ex:                       # NCCL missing: say so, keep the other transport
            res[name] = {"unavailable": str(ex)}
            continue
        conv.set_option("worker_spin_us", 200)
        call_loop(fn, hdl, xin, yout, 100, sx, sy, B)
        lat, dt = call_loop(fn, hdl, xin, yout, blocks, sx, sy, B)
        if lib.safconv_last_error(hdl):
            raise SystemExit("bench.py: multi-GPU apply failed: " + lib.safconv_last_error_string(hdl).decode())
        r = {"value": float(nOut) * hop * blocks / dt, "blocks": blocks, "warmup_blocks": 100,
             "block_latency_ms_p50": pct(lat, 50), "block_latency_ms_p99": pct(lat, 99)}
        period = min(hop / 48000.0, 0.005)
        pl, _ = call_loop(fn, hdl, xin, yout, 220, sx, sy, B, period)
        r["block_latency_paced_ms_p50"], r["block_latency_paced_ms_p99"] = pct(pl[20:], 50), pct(pl[20:], 99)
        conv.set_option("worker_spin_us", int(2e6 * period))          # workers keep spinning through the period
        pl, _ = call_loop(fn, hdl, xin, yout, 220, sx, sy, B, period)
        r["block_latency_paced_spinning_workers_ms_p50"], r["block_latency_paced_spinning_workers_ms_p99"] = pct(pl[20:], 50), pct(pl[20:], 99)
        r["paced_period_ms"] = 1e3 * period
        res[name] = r
    shards = [conv.shard_info(i) for i in range(len(conv.multi_devices()))]
    out = {"devices": conv.multi_devices(), "create_seconds": create_s,
           "outputs_per_device": [int(i.nOutLocal) for i in shards], "transports": res}
    conv.destroy()
    return out


def latency_config(name, args, dev):
    """secondary block: one BASELINE.json config through the host-pointer drop-in call (p50 / p99) + device-resident rate"""
    import torch
    import spatial_audio_framework_b200 as saf
    w = WORKLOADS[name]
    hop, nIn, nOut = w["hop"], w["nIn"], w["nOut"]
    H = filters_for(w, 0, nOut)
    conv = make_conv(saf, w, H, device=dev.index)
    info = conv.info()
    nbuf = 8
    g = torch.Generator(device="cpu").manual_seed(77)
    x_host = (torch.rand((nbuf, nIn, hop), generator=g) * 2 - 1).pin_memory()
    y_host = torch.empty((nbuf, nOut, hop), dtype=torch.float32).pin_memory()
    lat = host_api_latency(conv, w["kind"], x_host, y_host, nbuf, nIn * hop, nOut * hop, hop,
                           blocks=args.lat_blocks, warm=100)
    # device-resident batched rate
    B = 32
    d_in = (torch.rand((B, nIn, hop), generator=g) * 2 - 1).to(dev)
    d_out = torch.empty((B, nOut, hop), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    conv.set_stream(stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        conv.apply_device(d_in.data_ptr(), d_out.data_ptr(), B)
    steps = 20
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(steps):
            conv.apply_device(d_in.data_ptr(), d_out.data_ptr(), B)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (steps * B)
    conv.set_stream(None)
    chk = parity_check(conv, w, B, dev, blocks=B) if args.check else None
    cpu = None
    if not args.no_cpu:
        c = cpu_reference_run(w, steps=3, warmup=1, threads=1, ch_per_thread=nOut)
        cpu = {"value": c["value"], "unit": UNIT, "cores": 1, "kind": c["kind"], "ms_per_block": 1e3 * hop * nOut / c["value"],
               "sample": "the whole problem on ONE core, as the reference ships (single-threaded): " + c["sample"]}
    conv.destroy()
    small = info.bytesFilters <= (64 << 20)
    return {"workload": w["desc"], "host_api": lat,
            "device_resident": {"us_per_block": 1e3 * ms, "value": nOut * hop / (ms * 1e-3), "unit": UNIT, "blocks_per_step": B},
            "bound": ("latency (launch + PCIe round trip); filters and delay line are L2-resident: an HBM fraction is not a roofline here"
                      if small else "hbm"),
            "alg_bytes_per_block": float(info.algBytesPerBlock), "filter_bytes": int(info.bytesFilters),
            "parity": chk, "cpu_baseline": cpu}


def tvconv_latency(args, dev):
    """secondary block: saf_TVConv_apply (reference saf_utility_matrixConv.c:546-620) through the host-pointer drop-in call,
    the impulse-response set switching every 7 blocks (cross-fades), page-locked caller buffers"""
    import ctypes as C
    import torch
    import spatial_audio_framework_b200 as saf
    from spatial_audio_framework_b200 import synth
    out = {}
    fp = C.POINTER(C.c_float)
    for name, hop, L, nIRs, nOut, calls in (("hop512", 512, 24000, 16, 4, args.lat_blocks), ("hop8192", 8192, 24000, 16, 4, 300)):
        H = np.stack([synth.decaying_rir((nOut, L), seed=100 + i) for i in range(nIRs)])
        tv = saf.TVConv(hop, H, 0, device=dev.index)
        x = (torch.rand((hop,)) * 2 - 1).pin_memory()
        y = torch.empty((nOut, hop), dtype=torch.float32).pin_memory()
        xin, yout = C.cast(x.data_ptr(), fp), C.cast(y.data_ptr(), fp)
        fn, hdl = tv._lib.saf_TVConv_apply, tv.handle
        lat = []
        for k in range(calls + 100):
            t0 = time.perf_counter()
            fn(hdl, xin, yout, (k // 7) % nIRs)
            lat.append(1e3 * (time.perf_counter() - t0))
        if tv._lib.safconv_last_error(hdl):
            raise RuntimeError(tv._lib.safconv_last_error_string(hdl).decode())
        lat = lat[100:]
        out[name] = {"workload": f"saf_TVConv 1 x {nOut}, hop {hop}, {L} taps, {nIRs} IR sets switched every 7 blocks",
                     "p50_ms": pct(lat, 50), "p99_ms": pct(lat, 99), "blocks": calls, "warmup_blocks": 100,
                     "value": nOut * hop / (1e-3 * float(np.mean(lat))), "unit": UNIT}
        tv.destroy()
    return out


def producers_block(args, dev):
    """secondary block: the filter producers in front of the convolver -- the binaural Ambisonic decoder
    design (reference saf_hoa.c:452-497) and the shoebox image-source RIR bank (saf_reverb.c:184-295) -- timed through their
    C-ABI calls (wall clock; every call ends in a stream synchronisation), checked against the numpy fp64 restatement
    (oracle/producers.py) and timed against the compiled reference on ONE core on a bounded sample of the same work."""
    import spatial_audio_framework_b200 as saf
    from spatial_audio_framework_b200 import synth
    from oracle import producers as PR
    P = saf.producers
    saf.lib().safconv_set_device(dev.index)
    out = {}

    def best(f, n=3):
        b, r = 1e30, None
        for _ in range(n):
            t0 = time.perf_counter(); r = f(); b = min(b, time.perf_counter() - t0)
        return b, r

    # decoder design: order 7 (64 SH channels -> 2 ears), a KU100-sized grid, 513 bands
    H, d, itd = synth.synthetic_hrtfs(836, 1024, 48000.0)
    P.decoder_filters(H, d, 1024, 48000.0, P.DECODER_LS, 1)
    dec = {"workload": "getBinauralAmbiDecoderFilters: order 7, 836 directions, fftSize 1024 (513 bands)"}
    for name, m, dc, mr in (("LS", PR.LS, 0, 0), ("MAGLS_diffCM_maxRE", PR.MAGLS, 1, 1)):
        t, f = best(lambda: P.decoder_filters(H, d, 1024, 48000.0, m, 7, itd, None, dc, mr))
        truth = PR.np_decoder_filters(H, d, 1024, 48000.0, m, 7, itd, None, dc, mr)
        ent = {"gpu_ms": 1e3 * t, "parity_rel_l2_vs_fp64": float(np.linalg.norm(f - truth) / np.linalg.norm(truth))}
        if PR.producers_reference_available():
            R = PR.load_producers_reference()
            tr, fr = best(lambda: R.decoder_filters(H, d, 1024, 48000.0, m, 7, itd, None, dc, mr), 1)
            ent["reference_1_core_ms"] = 1e3 * tr
            ent["reference_rel_l2_vs_fp64"] = float(np.linalg.norm(fr - truth) / np.linalg.norm(truth))
        dec[name] = ent
    out["decoder"] = dec

    # image sources: the 64 x 64 x 96 000 bank of configs[3] = 64 sources, one 7th-order receiver, 2 s at 48 kHz
    nsrc, order, tmax = 64, 7, 2.0
    rng = np.random.default_rng(0)
    s = P.ImsShoebox(synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL, 125.0, 7, 343.0, 48e3)
    pos = [[rng.uniform(0.5, 9.5), rng.uniform(0.5, 6.5), rng.uniform(0.5, 2.5)] for _ in range(nsrc)]
    sids = [s.add_source(q) for q in pos]
    rec = [8.8, 5.5, 0.9]
    rid = s.add_receiver_sh(order, rec)
    s.compute_echograms(-1, 0.05); s.render_rirs(0)
    t0 = time.perf_counter()
    s.compute_echograms(-1, tmax); s.render_rirs(0)
    t_render = time.perf_counter() - t0
    images = sum(s.num_images(rid, k) for k in sids)
    t0 = time.perf_counter()
    h = s.matrixconv(rid, 1024)
    t_conv = time.perf_counter() - t0
    P.destroy_raw(h)
    # parity on a bounded sample: the first 0.25 s of source 0 against the restatement (same image set -> same taps)
    s2 = P.ImsShoebox(synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL, 125.0, 7, 343.0, 48e3)
    sid2 = s2.add_source(pos[0]); rid2 = s2.add_receiver_sh(order, rec)
    s2.compute_echograms(-1, 0.25); s2.render_rirs(0)
    r2 = s2.rir(rid2, sid2)
    ref, idx = PR.np_ims_rir(synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL, 7, 343.0, 48e3, pos[0], rec, order, -1, 0.25)
    ims = {"workload": f"ims_shoebox_renderRIRs: {nsrc} sources x one SH receiver of order {order} (64 ch), {tmax} s at 48 kHz "
                       "(the 64 x 64 x 96000 filter bank of configs[3]), 7 absorption bands, room 10 x 7 x 3 m",
           "image_sources": int(images), "render_s": t_render, "image_sources_per_s": images / t_render,
           "tap_updates_per_s": images * 64 / t_render, "bank_to_convolver_s": t_conv,
           "parity_sample": "source 0, first 0.25 s", "parity_same_image_count": bool(s2.num_images(rid2, sid2) == idx.size),
           "parity_same_taps": bool(np.array_equal(r2[0] != 0, ref[0] != 0)),
           "parity_rel_l2": float(np.linalg.norm(r2 - ref) / np.linalg.norm(ref))}
    s2.destroy(); s.destroy()
    if PR.producers_reference_available():
        R = PR.load_producers_reference()
        rs = R.ims(synth.IMS_TEST_ROOM, synth.IMS_TEST_ABS_WALL, 125.0, 7, 343.0, 48e3)
        sidr = rs.add_source(pos[0]); ridr = rs.add_receiver_sh(order, rec)
        t0 = time.perf_counter(); rs.compute_echograms(-1, 0.4); rs.render_rirs(0); tr = time.perf_counter() - t0
        nr = int(rs.echogram_times(ridr, sidr).size)
        rs.destroy()
        ims["reference_1_core"] = {"sample": "1 source, 0.4 s", "image_sources": nr, "s": tr, "image_sources_per_s": nr / tr}
    out["ims"] = ims
    return out


def run_own_arm(args, w):
    if w["kind"] == "offline":
        return run_offline_arm(args, w)
    import torch
    import spatial_audio_framework_b200 as saf
    from spatial_audio_framework_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    idle = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idle = dist.new_group(backend="gloo")           # ranks that have nothing to do wait here WITHOUT spinning
    dev = torch.device("cuda", local)

    hop, nIn, nOut, B = w["hop"], w["nIn"], w["nOut"], args.blocks
    ob, oc = sharding.shard_range(nOut, world, rank)
    H = filters_for(w, ob, oc)
    t_c0 = time.perf_counter()
    if w["kind"] == "matrix":
        conv = saf.MatrixConv(hop, H, 1, device=local) if world == 1 else \
            saf.MatrixConv.from_shard(hop, H, nOut, ob, device=local)
    else:
        conv = saf.MultiConv(hop, H, 1, device=local)
    create_s = time.perf_counter() - t_c0
    del H
    info = conv.info()
    stream = torch.cuda.Stream(device=dev)
    conv.set_stream(stream.cuda_stream)

    g = torch.Generator(device="cpu").manual_seed(1234)
    x_host = (torch.rand((B, nIn, hop), generator=g) * 2 - 1).pin_memory()          # one step of input, pinned
    y_host = torch.empty((B, nOut, hop), dtype=torch.float32).pin_memory()
    engine = sharding.ShardedStep(conv, w["kind"], nIn, nOut, hop, B, world, rank, dev, stream, dist)

    # ---------------- device-resident timing (value) ----------------
    engine.load_input(x_host)                      # inputs (both double-buffer slots) resident in HBM before the timed region
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        engine.step_device(prefetch=True)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    conv.enable_kernel_timing(B * args.steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(args.steps):
            engine.step_device(prefetch=True)
        engine.drain()
        e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    wall1 = time.time()
    ms_total = e0.elapsed_time(e1)
    ktot, ngroups, nkblocks = conv.kernel_totals_ms()
    kms = [t / max(nkblocks, 1) for t in ktot]
    conv.enable_kernel_timing(0)
    clocks = sampler.stop(wall0, wall1) if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    per_rank = None
    if dist:                                         # per-rank step time and MAC time: rules out imbalance between shards
        mine = torch.tensor([ms_total / (args.steps * B), kms[1]], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_block": [float(a[0]) for a in allr], "mac_ms_per_block": [float(a[1]) for a in allr]}
    ms_total = ms_total_max
    units = float(nOut) * hop * B * args.steps
    value = units / (ms_total * 1e-3)

    # ---------------- parity of what was just timed ----------------
    check = None
    if args.check and rank == 0:
        conv.set_stream(None)
        check = parity_check(conv, w, B, dev)
        conv.set_stream(stream.cuda_stream)

    # ---------------- end-to-end through the host-pointer API ----------------
    # the SAME contract at every N: one synchronous saf_matrixConv_apply(handle, in, out) per block on host buffers
    e2e = None
    if world == 1:
        conv.set_stream(None)
        lat = host_api_latency(conv, w["kind"], x_host, y_host, B, nIn * hop, nOut * hop, hop,
                               blocks=max(args.e2e_blocks, B), warm=100, paced_calls=200)
        conv.set_stream(stream.cuda_stream)
        e2e = {"value": float(nOut) * hop * lat["pinned"]["blocks_per_s"], "unit": UNIT,
               "h2d_bytes_per_step": int(B * nIn * hop * 4), "d2h_bytes_per_step": int(B * nOut * hop * 4),
               "blocks": lat["blocks"], "warmup_blocks": lat["warmup_blocks"],
               "api": ("saf_matrixConv_apply" if w["kind"] == "matrix" else "saf_multiConv_apply") +
                      " (one synchronous host-pointer call per block, page-locked caller buffers)",
               "block_latency_ms_p50": lat["pinned"]["p50_ms"], "block_latency_ms_p99": lat["pinned"]["p99_ms"],
               "block_latency_paced_ms_p50": lat["paced"]["p50_ms"], "block_latency_paced_ms_p99": lat["paced"]["p99_ms"],
               "paced_period_ms": lat["paced"]["period_ms"],
               "pageable_caller_buffers": {"value": float(nOut) * hop * lat["pageable"]["blocks_per_s"],
                                           "block_latency_ms_p50": lat["pageable"]["p50_ms"],
                                           "block_latency_ms_p99": lat["pageable"]["p99_ms"],
                                           "note": "malloc'd in / out frames like the reference's hosts (matrixconv.c:137-149): staged through the handle's pinned buffers"}}
    else:
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            mg = e2e_multi_gpu(args, w, world, x_host, y_host, B)
            best = max((k for k in mg["transports"] if "value" in mg["transports"][k]), key=lambda k: mg["transports"][k]["value"])
            r = mg["transports"][best]
            e2e = {"value": r["value"], "unit": UNIT,
                   "h2d_bytes_per_step": int(B * nIn * hop * 4) * (world if best == "host" else 1),
                   "d2h_bytes_per_step": int(B * nOut * hop * 4),
                   "blocks": r["blocks"], "warmup_blocks": r["warmup_blocks"], "transport": best,
                   "api": ("saf_matrixConv_apply" if w["kind"] == "matrix" else "saf_multiConv_apply") +
                          f" on ONE handle over {world} GPUs (safconv_matrixConv_create_multi): one synchronous host-pointer call per block, "
                          "single host process, one worker thread per device, page-locked caller buffers -- the same contract as N = 1",
                   "block_latency_ms_p50": r["block_latency_ms_p50"], "block_latency_ms_p99": r["block_latency_ms_p99"],
                   "block_latency_paced_ms_p50": r["block_latency_paced_ms_p50"], "block_latency_paced_ms_p99": r["block_latency_paced_ms_p99"],
                   "paced_period_ms": r["paced_period_ms"], "multi_gpu_handle": mg}
        dist.barrier(group=idle)                       # gloo: blocking socket wait, the other ranks do not burn host cores

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = measured_peak_gbs()
    # one MAC launch streams the filters once per block for all blocks of its launch group
    blocks_per_launch = nkblocks / max(ngroups, 1)
    mac_ms = ktot[1] / max(ngroups, 1)
    mac_bytes = float(info.macAlgBytesPerBlock) * blocks_per_launch
    achieved = mac_bytes / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0
    # DRAM traffic of the same kernel from the committed ncu --set full capture (per block; a launch covers blocks_per_launch)
    traffic, traffic_src = args.traffic, "--traffic" if args.traffic else None
    if traffic is None and world == 1:
        for tf in ("r02_traffic.json", "r01_traffic.json"):
            try:
                tj = json.loads((ROOT / "profiles" / tf).read_text()).get(args.workload)
                if tj:
                    traffic = float(tj["dram_bytes_per_block"]) * blocks_per_launch
                    traffic_src = f"profiles/{tf} (ncu dram__bytes_read.sum + dram__bytes_write.sum per block x blocks per launch)"
                    break
            except Exception:
                pass
    small = w["kind"] == "multi" or info.bytesFilters <= (64 << 20)
    roofline = {"bound": "hbm", "kernel": "mac_kernel (K2 filter-streaming complex MAC)" if w["kind"] == "matrix" else "multi_mac_ifft_w_kernel (per-channel MAC with a register sliding window + warp-level inverse FFT, batched)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": traffic, "traffic_source": traffic_src, "alg_bytes_per_launch": mac_bytes, "avg_launch_ms": mac_ms,
                "launches_timed": ngroups, "blocks_per_launch": blocks_per_launch, "rank": 0,
                "whole_block": {"alg_bytes": float(info.algBytesPerBlock),
                                "achieved_GBps": float(info.algBytesPerBlock) * B * args.steps / (ms_total * 1e-3) / 1e9},
                "kernel_ms_per_block": {"input_fft": kms[0], "mac": kms[1], "ifft_ola": kms[2]}}
    roofline["whole_block"]["frac"] = roofline["whole_block"]["achieved_GBps"] / peak
    if per_rank:
        roofline["per_rank"] = per_rank
    if small:
        roofline["bound"] = "l2"
        roofline["note"] = ("the filter set fits in L2 (126 MB) and the batched blocks of a step re-use it from L2 / registers: DRAM "
                            "traffic is far below the algorithmic bytes, so this is NOT an HBM roofline (frac can exceed 1); the bound "
                            "is L2 bandwidth / launch latency -- see profiles/ for the L2 counters")
    else:
        roofline["note"] = ("peak is the measured COPY bandwidth (read + write); this kernel is a read-only stream, which runs "
                            "above a copy on HBM3e -- see frac_of_nominal and traffic (ncu DRAM bytes = 0.986 x algorithmic)")
    roofline["peak_nominal"] = 8000.0
    roofline["frac_of_nominal"] = achieved / 8000.0

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        c = cpu_reference_run(w, steps=args.cpu_steps, warmup=1)
        cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu["host_threads"] = os.cpu_count()
        cpu["cpu_model"] = cpu_model()
        c1 = cpu_reference_run(w, steps=3, warmup=1, threads=1)
        cpu["as_shipped_1_core"] = {"value": c1["value"], "unit": UNIT, "cores": 1,
                                    "sample": "the reference convolver is single-threaded (SURVEY.md 8d): " + c1["sample"]}
        cpu["note"] = ("GPU / CPU ratios depend on the HOST of the box (the all-threads figure scales with its core count); "
                       "they say nothing about kernel quality -- the roofline fraction does")

    # ---------------- secondary block: the other BASELINE.json configs, bounded ----------------
    secondary = None
    if world == 1 and args.secondary and args.workload == "C4":
        conv.destroy()
        del engine
        torch.cuda.empty_cache()
        secondary = {}
        for name in ("C1", "C2", "C3", "UT"):
            try:
                secondary[name] = latency_config(name, args, dev)
            except Exception as ex:
                secondary[name] = {"failed": repr(ex)}
        try:
            secondary["TVConv"] = tvconv_latency(args, dev)
        except Exception as ex:
            secondary["TVConv"] = {"failed": repr(ex)}
        try:
            a5 = argparse.Namespace(**vars(args))
            a5.steps, a5.warmup, a5.e2e_steps, a5.cpu_steps = 5, 3, 2, 3
            secondary["C5"] = measure_offline(a5, WORKLOADS["C5"], 1, 0, local, None)
        except Exception as ex:
            secondary["C5"] = {"failed": repr(ex)}
        try:
            secondary["producers"] = producers_block(args, dev)
        except Exception as ex:
            secondary["producers"] = {"failed": repr(ex)}

    # matrix: per launch group (= one step of B blocks): forward FFT, MAC, inverse FFT, overlap-add chain
    launches_per_step = (4 if B > 1 else 3) * ((B + int(info.maxBatch) - 1) // int(info.maxBatch)) if w["kind"] == "matrix" else B
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, world),
        "config_detail": {"blocks_per_step": B, "partitions": int(info.numFilterBlocks), "filter_spectra_bytes_this_gpu": int(info.bytesFilters),
                          "sharding": f"output channels over {world} GPU(s), {oc} per GPU; value: one process per GPU, input batch NCCL-broadcast, output shards all-gathered; e2e: one multi-GPU handle in rank 0" if world > 1 else "single GPU",
                          "create_seconds_rank0": create_s},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline, "cpu_baseline": cpu,
        "ms_per_block": ms_total / (args.steps * B),
        "realtime_factor_48k": (hop * B * args.steps / (ms_total * 1e-3)) / 48000.0,
    }
    if check:
        line["parity_rel_l2"] = check["parity_rel_l2"]
        line["parity"] = check
    if secondary:
        line["secondary"] = secondary
    emit(line)
    if dist:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything except the final JSON line goes to stderr (NCCL / torchrun banners print to fd 1)."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--blocks", type=int, default=32, help="hop-sized blocks per step")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--e2e-blocks", type=int, default=2000, help="synchronous host-pointer calls timed for e2e / p50 (after 100 warm-up calls)")
    ap.add_argument("--lat-blocks", type=int, default=2000, help="calls per latency measurement of the secondary configs")
    ap.add_argument("--no-check", dest="check", action="store_false", help="skip the parity check of the timed path against the oracle")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false",
                    help="skip the secondary block (C1/C2/C3/UT latency and the offline C5 render) of the default N=1 line")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per MAC launch from the committed ncu capture (profiles/), echoed into roofline.traffic")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "b200" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, w)
    else:
        run_own_arm(args, w)


if __name__ == "__main__":
    main()
