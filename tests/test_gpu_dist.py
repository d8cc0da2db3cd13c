"""Two ranks on two GPUs over NCCL: the sharded Philox batch gives the unsharded result, episode statistics and MCTS
reward sums reduce over NVLink.  Skipped on a one-GPU box (the world-size-2 gloo test covers the host logic there)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
KEY, N, MOVES = 0xFEEDFACE, 4096 + 37, 6


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as tdist
    import ecg_b200 as E
    r, w, local = E.dist.init_from_env("nccl")
    dev = torch.device("cuda", local)
    first, count = E.dist.shard_range(N, w, r)
    bb = E.BatchedBoards(E.BoardConfig(seed=3), count, MOVES, device=dev, key=KEY, board0=first, env_goal=120)
    total = bb.rollout()
    stats = E.dist.stats_dict(E.dist.reduce_stats(bb.episode_stats()))
    st = E.BoardV2(4, E.BoardConfig(seed=3), device=dev)
    m = E.BatchedRolloutMCTS(st, 3, 3, False, leaves=2048 + 5, key=9)
    action, value, policies = m()
    q.put((r, first, count, total.cpu().numpy(), bb.array.cpu().numpy(), stats, action, m._root.reward, m._root.visits,
           m.env_steps))
    tdist.barrier()
    tdist.destroy_process_group()


def test_two_gpus_nccl_sharding_and_reductions():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import ecg_b200 as E
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29631, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # the unsharded run on one GPU
    ref = E.BatchedBoards(E.BoardConfig(seed=3), N, MOVES, device="cuda:0", key=KEY, env_goal=120)
    rt = ref.rollout().cpu().numpy()
    assert np.array_equal(np.concatenate([o[3] for o in out]), rt)
    assert np.array_equal(np.concatenate([o[4] for o in out]), ref.array.cpu().numpy())
    want = E.dist.stats_dict(ref.episode_stats())
    for o in out:
        assert o[5] == want  # both ranks hold the reduced statistics
    # MCTS: identical trees on both ranks (the reduced reward sums drive the same selections)
    assert out[0][6:9] == out[1][6:9] and out[0][9] == out[1][9] > 0
    one = E.BatchedRolloutMCTS(E.BoardV2(4, E.BoardConfig(seed=3), device="cuda:0"), 3, 3, False, leaves=2048 + 5, key=9)
    a1, _, _ = one()
    assert (a1, one._root.reward, one._root.visits) == out[0][6:9]  # and the same as one GPU rolling out all leaves
