"""N>1 path on CPU: world_size-2 (and 3, ragged) `gloo` runs of the output-channel sharding exchange
(spatial_audio_framework_b200/sharding.py): broadcast input batch -> per-rank shard -> all-gather.

The per-rank compute here is the CPU checker (oracle) on the rank's own output channels -- test
infrastructure standing in for the CUDA kernels, which have their own parity tests; what is under test is
the host logic: shard ranges, buffer rotation over consecutive steps (convolver state carries across
steps), gather layout and reassembly to [B][nOut][hop].
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, nIn, nOut, hop, L, B, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import oracle as O
    from spatial_audio_framework_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        if kind == "matrix":
            H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
        else:
            H = rng.uniform(-1, 1, (nOut, L)).astype(np.float32)
        x = rng.uniform(-1, 1, (steps, B, nIn, hop)).astype(np.float32)
        ob, oc = sharding.shard_range(nOut, world, rank)
        if oc > 0:
            conv = O.OracleMatrixConv(hop, H[ob:ob + oc], 1) if kind == "matrix" else O.OracleMultiConv(hop, H[ob:ob + oc], 1)

        def compute(xt, yt):
            xn = xt.numpy()
            for b in range(B):
                blk = xn[b] if kind == "matrix" else xn[b, ob:ob + oc]
                yt[b] = torch.from_numpy(conv.apply(np.ascontiguousarray(blk)))

        eng = sharding.ShardedStep(None, kind, nIn, nOut, hop, B, world, rank, torch.device("cpu"), None, dist, compute_fn=compute)
        outs = []
        for s in range(steps):
            # only rank 0 holds the real input; the others must receive it through the broadcast
            xh = torch.from_numpy(x[s]) if rank == 0 else torch.zeros((B, nIn, hop))
            yh = torch.empty((B, nOut, hop))
            if s % 2 == 0:
                eng.step_host(xh, yh)
                if rank == 0:
                    outs.append(yh.numpy().copy())
            else:
                buf = eng.t % 2
                if rank == 0:
                    eng.x[buf].copy_(xh)
                eng.step_device()
                outs.append(eng.last_output().numpy().copy())
        if rank == 0:
            q.put(np.stack(outs))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run(world, kind, nIn, nOut, hop, L, B, steps):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, nIn, nOut, hop, L, B, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


@pytest.mark.parametrize("world,kind,nIn,nOut", [(2, "matrix", 3, 4), (3, "matrix", 2, 5), (2, "multi", 6, 6), (4, "matrix", 2, 3)])
def test_sharded_exchange_matches_single_process(world, kind, nIn, nOut):
    import oracle as O
    hop, L, B, steps = 64, 200, 3, 4
    got = _run(world, kind, nIn, nOut, hop, L, B, steps)           # [steps][B][nOut][hop]
    rng = np.random.default_rng(7)
    if kind == "matrix":
        H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
        ref = O.OracleMatrixConv(hop, H, 1)
    else:
        H = rng.uniform(-1, 1, (nOut, L)).astype(np.float32)
        ref = O.OracleMultiConv(hop, H, 1)
    x = rng.uniform(-1, 1, (steps, B, nIn, hop)).astype(np.float32)
    for s in range(steps):
        for b in range(B):
            exp = ref.apply(np.ascontiguousarray(x[s, b]))
            assert np.array_equal(got[s, b], exp), (s, b)


def test_shard_range_properties():
    from spatial_audio_framework_b200.sharding import shard_range
    for n in (1, 2, 5, 64, 65, 121):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            cs = [c for _, c in spans]
            assert max(cs) - min(cs) <= 1


def test_time_segment_properties_and_halo_is_sufficient():
    """Offline time sharding (host logic): segments tile the signal, and a P-frame halo reproduces the full
    block convolution exactly (checked with the CPU oracle as the per-segment compute)."""
    import oracle as O
    from spatial_audio_framework_b200.sharding import time_segment
    rng = np.random.default_rng(3)
    hop, L, nIn, nOut, T = 64, 300, 2, 2, 41
    P = (L + hop - 1) // hop
    H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
    x = rng.uniform(-1, 1, (nIn, hop * T)).astype(np.float32)
    full = O.OracleMatrixConv(hop, H, 1).run(x)
    for world in (1, 2, 5, 8):
        segs = [time_segment(T, world, r, P) for r in range(world)]
        assert segs[0][0] == 0 and segs[-1][1] == T
        for (a0, a1, _), (b0, _, hb) in zip(segs, segs[1:]):
            assert a1 == b0 and hb == min(P, b0)
        parts = []
        for t0, t1, halo in segs:
            y = O.OracleMatrixConv(hop, H, 1).run(np.ascontiguousarray(x[:, (t0 - halo) * hop:t1 * hop]))
            parts.append(y[:, halo * hop:])
        got = np.concatenate(parts, 1)
        # frames whose history is complete inside the halo are bit-identical to the full run
        assert np.array_equal(got, full), world
