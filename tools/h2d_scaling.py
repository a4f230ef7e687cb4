"""tools/h2d_scaling.py — concurrent pinned host -> device bandwidth per rank (run under torchrun with 1, 2, 4, 8 ranks): the
platform term of the end-to-end number.  Every rank copies the same amount from its own pinned buffer at the same time."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gb = 4
h = torch.empty(gb * (1 << 27), dtype=torch.float64).pin_memory()
h.fill_(1.0)
d = torch.empty_like(h, device="cuda")
d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
best = 0.0
for _ in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    best = max(best, gb * 1.073741824 / (time.perf_counter() - t0))
t = torch.tensor([best], device="cuda")
if world > 1:
    lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    su = t.clone(); dist.all_reduce(su, op=dist.ReduceOp.SUM)
    if rank == 0:
        print("H2D, %d concurrent ranks, %d GiB each from pinned memory: slowest rank %.1f GB/s, aggregate %.1f GB/s" % (world, gb, lo.item(), su.item()), flush=True)
    dist.destroy_process_group()
else:
    print("H2D, 1 rank, %d GiB from pinned memory: %.1f GB/s" % (gb, best), flush=True)
