"""samplerTasks.py (:9-32) and util/multiprocessingAutoBatcher.py (:37-56) for the batched engine.

The reference samples episodes one Python object at a time -- `async_pbar_auto_batcher(random_task, n)` splits n calls
of a zero-argument task over a `multiprocessing.Pool` (main.py:244-252) and returns the flat list of final rewards.
Here a task plays ALL the episodes of a call in lockstep on the GPU, and the batcher splits n over the ranks of the
process group (one process per GPU) instead of over CPU workers; the list it returns has the same length and meaning.

    random_task()      one episode of uniformly random legal actions  -> int   (samplerTasks.py:9-14)
    random_task(n)     n such episodes in one rollout kernel          -> list[int]
    greedy_test(n)     n greedy episodes (boardv2.py:209-218 per move) -> list[int]  (samplerTasks.py:17-22)
    mcts_task(n)       n episodes played by the MCTS                   -> list[int]  (samplerTasks.py:25-32)
    async_pbar_auto_batcher(task, n, **kw) -> list[int] of length n on every rank

Seeds.  The reference's tasks build `BoardConfig()`, whose seed is drawn from numpy's global RNG (boardConfig.py:31);
`seeds=` passes them explicitly (one per episode), which is what makes an episode reproducible -- and, in
refill="replay" (the default, the reference's own RNG semantics), bit-identical to the reference's episode for that seed.
refill="philox" is the throughput mode: episodes are keyed by (key, global episode index), independent of the sharding.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import dist as ecg_dist
from .boards import BatchedBoards
from .config import BoardConfig
from .mcts import BatchedRolloutMCTS
from .state import BoardV2


def batch_data(data_size: int, insert=None):
    """multiprocessingAutoBatcher.py:37-43 verbatim in meaning: sizes of the per-worker batches for `cpus` workers --
    here workers = ranks.  Kept for callers that import it; the engine itself uses dist.shard_range."""
    world = _world()[1]
    cpus = max(1, min(int(data_size), world))
    per = int(data_size) // cpus
    under = int(data_size) - per * cpus
    sizes = [per + 1 if i < under else per for i in range(cpus)]
    return [(insert, s) for s in sizes] if insert else sizes


def _world():
    d = torch.distributed
    if d.is_available() and d.is_initialized():
        return d.get_rank(), d.get_world_size()
    return 0, 1


def _draw_seeds(n: int) -> np.ndarray:
    """BoardConfig.__post_init__ (:31): `seed or np.random.randint(0, 2 ** 31 - 1)` -- a drawn 0 is drawn again."""
    s = np.random.randint(0, 2 ** 31 - 1, size=n).astype(np.int64)
    while (s == 0).any():
        z = s == 0
        s[z] = np.random.randint(0, 2 ** 31 - 1, size=int(z.sum()))
    return s


def _episodes(samples, seeds, cfg, refill):
    """-> (n, seeds int64 [n] | None, scalar result wanted)"""
    single = samples is None and seeds is None
    if seeds is not None:
        seeds = np.asarray(seeds, dtype=np.int64).reshape(-1)
        if samples is not None and int(samples) != seeds.size:
            raise ValueError("samples and len(seeds) differ")
        if (seeds == 0).any():
            raise ValueError("seed 0 cannot be asked for: BoardConfig replaces it by a random one (boardConfig.py:31)")
        n = seeds.size
    else:
        n = 1 if samples is None else int(samples)
        if refill == "replay":
            seeds = _draw_seeds(n)
    if n < 0:
        raise ValueError("samples must be >= 0")
    return n, seeds, single


# Episodes per launch.  Replay mode keeps one MT19937 stream per episode in HBM (stream_len words: 32 KB at the
# default 8192), and a greedy move expands every legal child of every board (~17 per board at 9x9x6).
REPLAY_SLICE = 1 << 16
PHILOX_SLICE = 1 << 24
GREEDY_SLICE = 1 << 16


def _slices(n: int, refill: str, per: int = None):
    per = per or (REPLAY_SLICE if refill == "replay" else PHILOX_SLICE)
    return [(lo, min(lo + per, n)) for lo in range(0, n, per)]


def _cfg(cfg: Optional[BoardConfig]) -> BoardConfig:
    return cfg if cfg is not None else BoardConfig()


def _boards(cfg, n, moves, seeds, refill, key, first, device, stream_len):
    if refill == "replay":
        return BatchedBoards(cfg, n, moves, device=device, refill="replay", seeds=seeds, stream_len=stream_len)
    return BatchedBoards(cfg, n, moves, device=device, refill="philox", key=key, board0=first)


def random_task(samples: Optional[int] = None, *, cfg: BoardConfig = None, moves: int = 20,
                seeds: Sequence[int] = None, refill: str = "replay", key: int = None, first: int = 0, device=None,
                stream_len: int = 8192):
    """samplerTasks.random_task: `state = BoardV2(20, BoardConfig()); np.random.seed(cfg.seed); while not terminal:
    state = state.apply_action(np.random.choice(state.legal_actions))`, for `samples` episodes in ONE rollout kernel
    (ecg_rollout).  Returns the final cumulative rewards (an int for the reference's zero-argument call)."""
    cfg = _cfg(cfg)
    n, seeds, single = _episodes(samples, seeds, cfg, refill)
    out = []
    for lo, hi in _slices(n, refill):
        bb = _boards(cfg, hi - lo, moves, None if seeds is None else seeds[lo:hi], refill, key, first + lo, device,
                     stream_len)
        bb.rollout()
        _check(bb)
        out.extend(bb.reward.tolist())
    return out[0] if single else out


def greedy_test(samples: Optional[int] = None, *, cfg: BoardConfig = None, moves: int = 20,
                seeds: Sequence[int] = None, refill: str = "replay", key: int = None, first: int = 0, device=None,
                stream_len: int = 8192, return_actions: bool = False):
    """samplerTasks.greedy_test: every move is BoardV2.greedy_action (the legal action with the largest one-step
    reward, first maximum wins) -- per move ONE expansion launch over every legal child of every board."""
    cfg = _cfg(cfg)
    n, seeds, single = _episodes(samples, seeds, cfg, refill)
    out, acts = [], []
    for lo, hi in _slices(n, refill, per=GREEDY_SLICE):
        bb = _boards(cfg, hi - lo, moves, None if seeds is None else seeds[lo:hi], refill, key, first + lo, device,
                     stream_len)
        taken = []
        seen = torch.zeros_like(bb.status)  # status bytes of all moves, OR-ed on the device (no sync in the loop)
        for _ in range(moves):
            a = bb.greedy_action()  # -1: no legal action (capped shuffle loop) -> the step is a flagged no-op
            bb.apply_action(a)
            seen |= bb.status
            if return_actions:
                taken.append(a.clone())
        bb.status.copy_(seen)
        _check(bb)
        out.extend(bb.reward.tolist())
        if return_actions:
            acts.append(torch.stack(taken, dim=1).cpu().numpy())
    res = out[0] if single else out
    if return_actions:
        return res, (np.concatenate(acts) if acts else np.zeros((0, moves), dtype=np.int32))
    return res


def mcts_task(samples: Optional[int] = None, *, cfg: BoardConfig = None, moves: int = 20,
              seeds: Sequence[int] = None, simulations: int = 100, leaves: int = 1 << 14, refill: str = "philox",
              key: int = 0x5EED, first: int = 0, device=None):
    """samplerTasks.mcts_task: `mcts = MCTS(state, 2, 100, False); while not terminal: action, _, _ = mcts(); state =
    state.apply_action(action)`.  The tree is the reference's host tree, sequential per episode; every simulation rolls
    out `leaves` episodes on the GPU(s) (BatchedRolloutMCTS)."""
    cfg0 = _cfg(cfg)
    single = samples is None and seeds is None
    n = len(seeds) if seeds is not None else (1 if samples is None else int(samples))
    out = []
    for e in range(n):
        c = cfg0 if seeds is None else BoardConfig(seed=int(seeds[e]), rows=cfg0.rows, columns=cfg0.columns,
                                                    types=cfg0.types)
        if seeds is None and e:  # a fresh BoardConfig() per episode, like the reference's task
            c = BoardConfig(rows=cfg0.rows, columns=cfg0.columns, types=cfg0.types)
        state = BoardV2(moves, c, device=device)
        mcts = BatchedRolloutMCTS(state, 2, simulations, False, False, leaves=leaves, key=key + first + e,
                                  refill=refill)
        while not state.is_terminal:
            action, _, _ = mcts()
            state = state.apply_action(action)
        out.append(int(state.reward))
    return out[0] if single else out


def _check(bb: BatchedBoards):
    from . import _native as N
    bad = N.ST_STREAM_OVERFLOW | N.ST_BAD_CELL
    st = int(torch.bitwise_and(bb.status, bad).max().item()) if bb.n else 0
    if st & N.ST_STREAM_OVERFLOW:
        raise RuntimeError("a replay stream ran out: pass a larger stream_len")
    if st:
        raise RuntimeError(f"board status {st:#x}")


def async_pbar_auto_batcher(func: Callable, data_size: int, *, seeds: Sequence[int] = None, **kw) -> List[int]:
    """util/multiprocessingAutoBatcher.async_pbar_auto_batcher(func, data_size): `data_size` runs of `func`, split like
    batch_data over the workers -- here the ranks of the process group (torchrun, one per GPU; a single process runs
    them all).  `func(count, first=..., **kw)` plays this rank's share; the flat result list (rank order = the
    reference's worker order) is returned on every rank.  `seeds` (one per run) are sharded with the runs."""
    rank, world = _world()
    if func is mcts_task:
        # its rollouts are already sharded over the ranks (BatchedRolloutMCTS reduces the leaves' rewards over the
        # group): every rank walks the same trees together, so the episodes themselves are not split
        if seeds is not None:
            kw["seeds"] = seeds
        return list(func(int(data_size), **kw)) if seeds is None else list(func(**kw))
    first, count = ecg_dist.shard_range(int(data_size), world, rank)
    if seeds is not None:
        seeds = np.asarray(seeds, dtype=np.int64).reshape(-1)
        if seeds.size != int(data_size):
            raise ValueError("seeds must hold one seed per run")
        kw["seeds"] = seeds[first:first + count]
    mine = list(func(count, first=first, **kw)) if count else []
    if world == 1:
        return mine
    parts = [None] * world
    torch.distributed.all_gather_object(parts, mine)
    return [x for p in parts for x in p]
