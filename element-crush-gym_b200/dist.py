"""Multi-GPU: one process per GPU, boards sharded by global board index, no collective in the step
path.  The only communication is the optional reduction of episode statistics / MCTS visit counts
(the reference's counterpart is the flat result list returned through multiprocessing.Pool,
util/multiprocessingAutoBatcher.py:19-34)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int):
    """Contiguous split of [0, total) like batch_data (multiprocessingAutoBatcher.py:37-43): the first
    total % world ranks get one extra.  Returns (first_global_index, count)."""
    base, extra = divmod(int(total), int(world))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def bind_to_gpu_numa_node(index: int) -> bool:
    """Pin this process to the CPUs next to GPU `index` (NVML's ideal CPU affinity) before any pinned host buffer is
    allocated: pages of cudaHostAlloc memory land on the allocating thread's NUMA node, and a D2H stream that crosses
    the socket interconnect runs at a fraction of the PCIe rate (bench.py e2e with 2+ ranks on a 2-socket host).
    Returns False (and changes nothing) when NVML or sched_setaffinity is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def init_from_env(backend: str = None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_* (torchrun); no-op when single."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
            if os.environ.get("ECG_NUMA_BIND", "0") == "1":  # opt-in: the pool's boxes are single-node VMs (no effect, r04e)
                bind_to_gpu_numa_node(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """All-reduce the int64[6] vector of BatchedBoards.episode_stats():
    [sum, n, min, max, wins, sum_sq] -> sums are added, min/max reduced with MIN/MAX."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    sums = stats[[0, 1, 4, 5]].clone()
    mn = stats[2:3].clone()
    mx = stats[3:4].clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    out = stats.clone()
    out[0], out[1], out[4], out[5] = sums[0], sums[1], sums[2], sums[3]
    out[2], out[3] = mn[0], mx[0]
    return out


def reduce_visit_counts(visits: torch.Tensor, reward_sums: torch.Tensor):
    """Sum per-root-child visit counts and reward sums over ranks (MCTS backprop, abc/mcts.py:105-107)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in (visits, reward_sums):
            if t.numel():
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return visits, reward_sums


def stats_dict(stats: torch.Tensor) -> dict:
    s = [int(x) for x in stats.cpu().tolist()]
    n = max(s[1], 1)
    mean = s[0] / n
    var = max(s[5] / n - mean * mean, 0.0)
    return {"episodes": s[1], "mean": mean, "min": s[2], "max": s[3], "std": var ** 0.5, "wins": s[4]}
