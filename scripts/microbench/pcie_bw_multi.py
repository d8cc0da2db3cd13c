"""Pinned D2H / H2D copy bandwidth with all ranks copying at once (torchrun --nproc-per-node N): tells whether the GPUs
of the box share their host link.  Prints per-rank and aggregate GB/s."""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
n = 1536 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, f in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    f()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    for _ in range(4):
        f()
    torch.cuda.synchronize()
    gbs = torch.tensor([4 * n / (time.perf_counter() - t) / 1e9], device="cuda")
    if world > 1:
        g = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(g, gbs)
    else:
        g = [gbs]
    if rank == 0:
        v = [float(x) for x in g]
        print(name, "per rank", ["%.1f" % x for x in v], "aggregate %.1f GB/s" % sum(v), flush=True)
if world > 1:
    dist.destroy_process_group()
