"""samplerTasks.py / util/multiprocessingAutoBatcher.py call sites on the batched engine (sampler.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _sampler():
    from importlib import import_module
    return import_module("element-crush-gym_b200.sampler")


def _reference_batch_data(data_size, cpus_available):
    # util/multiprocessingAutoBatcher.py:37-43 with cpu_count() = cpus_available
    cpus = min(data_size, cpus_available)
    per = int(data_size / cpus)
    under = data_size - per * cpus
    return [per + 1 if i < under else per for i in range(cpus)]


def test_batch_data_single_process():
    S = _sampler()
    assert S.batch_data(7) == _reference_batch_data(7, 1) == [7]
    assert S.batch_data(3, "f") == [("f", 3)]


def _fake_task(count, first=0, seeds=None, scale=1):
    # stands in for a GPU task: one "reward" per run, a function of the run's global index and seed
    return [scale * (1000 * (first + i) + (int(seeds[i]) if seeds is not None else 0)) for i in range(count)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import ecg_b200 as E
    E.dist.init_from_env(backend="gloo")
    S = _sampler()
    sizes = S.batch_data(7)
    a = S.async_pbar_auto_batcher(_fake_task, 7, seeds=list(range(1, 8)), scale=2)
    b = S.async_pbar_auto_batcher(_fake_task, 1)  # fewer runs than ranks: the last rank plays nothing
    q.put((rank, sizes, a, b))
    tdist.destroy_process_group()


def test_auto_batcher_two_gloo_ranks():
    world, port = 2, 29633
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [2 * (1000 * i + (i + 1)) for i in range(7)]
    for rank, sizes, a, b in out:
        assert sizes == _reference_batch_data(7, 2) == [4, 3]
        assert a == want  # rank order = the reference's worker order, every rank holds the whole list
        assert b == [0]


def test_seed_rules():
    S = _sampler()
    np.random.seed(5)
    s = S._draw_seeds(1000)
    assert s.dtype == np.int64 and (s > 0).all() and (s < 2 ** 31 - 1).all()
    with pytest.raises(ValueError):
        S._episodes(None, [0, 1], None, "replay")  # BoardConfig would replace seed 0
    with pytest.raises(ValueError):
        S._episodes(3, [1, 2], None, "replay")
    assert S._episodes(None, None, None, "philox") == (1, None, True)


# ------------------------------------------------------------------ GPU

@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ecg_b200
    return ecg_b200


@pytest.mark.gpu
def test_random_task_matches_reference_episodes(E):
    """samplerTasks.random_task with explicit seeds == the reference-generated episodes (tests/golden/episodes_*)."""
    S = _sampler()
    d = np.load(os.path.join(GOLDEN, "episodes_9x9x6.npz"))
    seeds, moves = d["seeds"].astype(np.int64), int(d["moves"])
    want = d["rewards"].sum(axis=1).tolist()
    assert S.random_task(seeds=seeds, moves=moves) == want
    assert S.async_pbar_auto_batcher(S.random_task, len(seeds), seeds=seeds, moves=moves) == want
    one = S.random_task(seeds=seeds[3:4], moves=moves)
    assert one == [want[3]]
    d6 = np.load(os.path.join(GOLDEN, "episodes_6x6x4.npz"))
    cfg = E.BoardConfig(rows=6, columns=6, types=4)
    assert S.random_task(seeds=d6["seeds"].astype(np.int64), moves=int(d6["moves"]), cfg=cfg) == \
        d6["rewards"].sum(axis=1).tolist()


@pytest.mark.gpu
def test_random_task_call_shapes(E):
    S = _sampler()
    r = S.random_task()  # the reference's zero-argument call: one episode, a plain int
    assert isinstance(r, int) and r >= 0
    assert S.random_task(0) == []
    a = S.random_task(1000, refill="philox", key=77)
    assert len(a) == 1000 and a == S.random_task(1000, refill="philox", key=77)
    # Philox episodes are keyed by the global episode index: any split gives the same list
    assert S.random_task(600, refill="philox", key=77) + S.random_task(400, refill="philox", key=77, first=600) == a
    assert S.async_pbar_auto_batcher(S.random_task, 1000, refill="philox", key=77) == a
    assert 200 < np.mean(a) < 500  # 20-move random episodes at 9x9x6: 345 +- 144
    # large requests are played in slices: same list
    keep = S.PHILOX_SLICE, S.REPLAY_SLICE, S.GREEDY_SLICE
    try:
        S.PHILOX_SLICE, S.REPLAY_SLICE, S.GREEDY_SLICE = 256, 16, 16
        assert S.random_task(1000, refill="philox", key=77) == a
        seeds = np.arange(1, 41)
        r1 = S.random_task(seeds=seeds, moves=5)
        g1 = S.greedy_test(seeds=seeds, moves=3)
        gp = S.greedy_test(40, refill="philox", key=9, moves=3)
    finally:
        S.PHILOX_SLICE, S.REPLAY_SLICE, S.GREEDY_SLICE = keep
    assert S.random_task(seeds=seeds, moves=5) == r1
    assert S.greedy_test(seeds=seeds, moves=3) == g1
    assert S.greedy_test(40, refill="philox", key=9, moves=3) == gp


@pytest.mark.gpu
def test_greedy_test_matches_reference_episodes(E):
    """samplerTasks.greedy_test (:17-22) with explicit seeds == episodes generated by the unmodified reference
    (scripts/gen_golden_sampler.py): every greedy action and the final reward."""
    S = _sampler()
    d = np.load(os.path.join(GOLDEN, "sampler_greedy.npz"))
    for k in range(int(d["cases"])):
        R, Cc, T = (int(x) for x in d[f"shape{k}"])
        cfg = E.BoardConfig(rows=R, columns=Cc, types=T)
        rewards, actions = S.greedy_test(seeds=d[f"seeds{k}"], moves=int(d[f"moves{k}"]), cfg=cfg,
                                         return_actions=True)
        assert np.array_equal(actions, d[f"actions{k}"])
        assert rewards == d[f"rewards{k}"].tolist()
    g = S.greedy_test(64, refill="philox", key=3)
    r = S.random_task(64, refill="philox", key=3)
    assert np.mean(g) > np.mean(r)  # one-step lookahead beats random play


@pytest.mark.gpu
def test_mcts_task_smoke(E):
    S = _sampler()
    r = S.mcts_task(seeds=[5], moves=3, simulations=12, leaves=256)
    assert isinstance(r, list) and len(r) == 1 and r[0] >= 0
    # the MCTS-played episode is at least as good as nothing: three moves always score
    assert r[0] >= 3 * 6
