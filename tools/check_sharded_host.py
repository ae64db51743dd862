"""torchrun check: ShardedStep.step_host (sub-batch pipeline over NCCL) == one full handle on rank 0.  Debug tooling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import spatial_audio_framework_b200 as saf
from spatial_audio_framework_b200 import sharding

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
hop, L, nIn, nOut, B = 256, 3000, 6, 8, 16
rng = np.random.default_rng(5)
H = rng.uniform(-1, 1, (nOut, nIn, L)).astype(np.float32)
ob, oc = sharding.shard_range(nOut, world, rank)
conv = saf.MatrixConv.from_shard(hop, H[ob:ob + oc], nOut, ob, device=local)
stream = torch.cuda.Stream(device=dev)
conv.set_stream(stream.cuda_stream)
eng = sharding.ShardedStep(conv, "matrix", nIn, nOut, hop, B, world, rank, dev, stream, dist)
full = saf.MatrixConv(hop, H, 1, device=local) if rank == 0 else None
ok = True
for step in range(4):
    x = rng.uniform(-1, 1, (B, nIn, hop)).astype(np.float32)          # same on every rank (same seed)
    xh = torch.from_numpy(x).pin_memory(); yh = torch.zeros((B, nOut, hop)).pin_memory()
    eng.step_host(xh, yh)
    if rank == 0:
        ref = np.stack([full.apply(x[b]) for b in range(B)])
        err = float(np.abs(yh.numpy() - ref).max() / np.abs(ref).max())
        print(f"step {step}: S={getattr(eng, 'S', 1)} max-abs/fs {err:.3g}")
        ok = ok and err <= 1e-5
dist.barrier()
if rank == 0:
    print("sharded host step:", "OK" if ok else "MISMATCH")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
