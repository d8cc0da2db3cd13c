"""Host-side multi-rank logic on CPU: sharding by global board index and the statistics / visit-count
reductions, run as 2 gloo ranks (the N>1 path of bench.py uses the same functions over NCCL)."""
import os
import sys

import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_range_matches_reference_batch_data():
    import ecg_b200 as E
    # util/multiprocessingAutoBatcher.py:37-43: first n % cpus workers get one extra task
    for total, world in ((10, 4), (16, 8), (7, 8), (1 << 24, 8), (5, 1)):
        parts = [E.dist.shard_range(total, world, r) for r in range(world)]
        assert sum(c for _, c in parts) == total
        base, extra = divmod(total, world)
        assert [c for _, c in parts] == [base + 1 if r < extra else base for r in range(world)]
        pos = 0
        for first, count in parts:
            assert first == pos
            pos += count


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import ecg_b200 as E
    r, w, _ = E.dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # per-rank episode statistics [sum, n, min, max, wins, sum_sq]
    scores = torch.arange(10 * rank, 10 * rank + 10, dtype=torch.int64)
    stats = torch.tensor([scores.sum(), scores.numel(), scores.min(), scores.max(), (scores >= 15).sum(),
                          (scores * scores).sum()], dtype=torch.int64)
    red = E.dist.reduce_stats(stats)
    visits = torch.full((144,), rank + 1, dtype=torch.int64)
    rsum = torch.full((144,), 10 * (rank + 1), dtype=torch.int64)
    E.dist.reduce_visit_counts(visits, rsum)
    q.put((rank, red.tolist(), int(visits[0]), int(rsum[0]), E.dist.stats_dict(red)))
    tdist.destroy_process_group()


def test_two_rank_gloo_reductions():
    world, port = 2, 29611
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allsc = torch.arange(0, 20, dtype=torch.int64)
    want = [int(allsc.sum()), 20, 0, 19, int((allsc >= 15).sum()), int((allsc * allsc).sum())]
    for rank, red, v, rs, sd in out:
        assert red == want
        assert v == 3 and rs == 30
        assert sd["episodes"] == 20 and sd["min"] == 0 and sd["max"] == 19 and abs(sd["mean"] - 9.5) < 1e-12
