"""Self-play data for the policy/value net: the caller side of the stepping path.

Mirrors dataset.py of the reference:
    mcts_task  (dataset.py:16-43)   one MCTS-driven episode after the other -> observations / policies / values
    Dataset    (dataset.py:46-241)  sample (with the reference's pickle cache "(rows, cols, types).ds"),
                                    with_mirroring / with_type_switching / with_batching, get_split
with the rollouts of every MCTS simulation batched on the GPU (BatchedRolloutMCTS) and both augmentations done
on the packed device boards by ecg_augment (csrc: augment_kernel); the policy of a mirrored sample is the
original policy re-indexed by the fixed action permutation `mirror_actions(cfg)`.

Reference behaviour kept on purpose:
  * policies[a] = p for (a, p) in zip(state.legal_actions, policy_logits) (dataset.py:31-33): the i-th ASCENDING
    legal action receives the visit share of the i-th EXPANDED child (children expand from the largest action
    down, standard/mcts.py:33); `pair_policies="by_action"` gives the share to the action that earned it instead.
  * values: the final episode reward, repeated once per move (dataset.py:41).
  * sample() rounds the size up to a multiple of 20 (dataset.py:66) and get_split divides the values by the
    largest value of the data set (dataset.py:193).
Deviation: the reference's type_switch (dataset.py:114-176) renames tokens through 7-letter permutations that
also produce the non-existent type `types + 1` and adds mega_token to every special cell (values outside the
closed cell set, SURVEY.md 8a invariants); here a type switch is a permutation of the plain types 1..types and
leaves empty cells and special tokens untouched.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from itertools import islice, permutations
from pickle import dump, load

import numpy as np
import torch

from . import _native as N
from .boards import BatchedBoards, _ptr, _stream
from .config import BoardConfig
from .mcts import BatchedRolloutMCTS
from .state import BoardV2


def mirror_actions(cfg: BoardConfig) -> np.ndarray:
    """perm[a] = the action that swaps the left-right mirrored cell pair of action a (dataset.py:99-107)."""
    perm = np.empty(cfg.action_space, dtype=np.int64)
    for a in range(cfg.action_space):
        (r1, c1), (r2, c2) = cfg.decode(a)
        perm[a] = cfg.encode((r1, cfg.columns - 1 - c1), (r2, cfg.columns - 1 - c2))
    return perm


def mirror_policies(cfg: BoardConfig, policies: torch.Tensor) -> torch.Tensor:
    """[..., action_space] -> mirrored_policy[perm[a]] = policy[a]"""
    perm = torch.as_tensor(mirror_actions(cfg), device=policies.device)
    out = torch.zeros_like(policies)
    out[..., perm] = policies
    return out


def augment_boards(boards: BatchedBoards, mirror: bool = False, type_perm=None) -> BatchedBoards:
    """A new batch holding the mirrored and / or type-permuted boards (scores, moves etc. are copied)."""
    out = boards.clone()
    perm = None
    if type_perm is not None:
        perm = bytes(int(t) for t in type_perm)
        if len(perm) != boards.cfg.types:
            raise ValueError("type_perm needs one entry per type")
    N.check(boards.L.ecg_augment(C.byref(boards.nat), _ptr(boards.boards), _ptr(out.boards), int(bool(mirror)), perm,
                                 boards.n, _stream(boards.device)), "ecg_augment")
    out._mask_valid = False
    return out


def mcts_task(cfg: BoardConfig, moves: int, count: int, *, simulations: int = 256, leaves: int = 4096,
              key: int = 0x5EED, device=None, pair_policies: str = "reference", callback=None):
    """dataset.py:16-43: whole episodes until more than `count` samples exist.
    Returns {'observations': [int64 [R, C]], 'policies': [float64 [A]], 'values': [int]}."""
    data = {"observations": [], "policies": [], "values": []}
    n = 0
    episode = 0
    while n <= count:  # dataset.py:26-27 (`if count > batch_size: break`, tested once per episode)
        state = BoardV2(moves, cfg, device=device)
        mcts = BatchedRolloutMCTS(state, 3, simulations, False, False, leaves=leaves, key=key + episode)
        while not state.is_terminal:
            action, _, shares = mcts()
            policies = np.zeros((cfg.action_space,))
            if pair_policies == "reference":
                for a, p in zip(state.legal_actions, shares):
                    policies[a] = p
            else:
                for a, p in zip(mcts.last_root_actions, shares):
                    policies[a] = p
            data["observations"].append(state.array)
            data["policies"].append(policies)
            state = state.apply_action(action)
            n += 1
            if callback:
                callback()
        data["values"].extend([state.reward] * moves)
        episode += 1
    return data


class Dataset:
    """dataset.py:46-241 with device-side augmentation."""

    def __init__(self, cfg: BoardConfig, moves: int = 20, *, simulations: int = 256, leaves: int = 4096,
                 key: int = 0x5EED, device=None, pair_policies: str = "reference"):
        self.cfg = cfg
        self.moves = moves
        self.simulations, self.leaves, self.key, self.device = simulations, leaves, key, device
        self.pair_policies = pair_policies
        self._size = 0
        self._mirroring = False
        self._batching = 1
        self._type_switching = False
        self._type_switching_limit = -1
        self.dataset = {"observations": [], "policies": [], "values": []}

    @property
    def cache_file(self) -> str:
        return str((*self.cfg.shape, self.cfg.types)) + ".ds"  # dataset.py:67

    def sample(self, size, caching=True, directory="."):
        size = 20 * math.ceil(size / 20)
        file = os.path.join(directory, self.cache_file)
        if caching and os.path.isfile(file) and len(self.dataset["values"]) == 0:
            with open(file, "rb") as fh:
                self.dataset = load(fh)
        missing = size - len(self.dataset["values"])
        if missing > 0:
            batch = mcts_task(self.cfg, self.moves, missing - 1, simulations=self.simulations, leaves=self.leaves,
                              key=self.key + len(self.dataset["values"]), device=self.device,
                              pair_policies=self.pair_policies)
            for k, v in batch.items():
                self.dataset[k].extend(v)
            if caching:
                with open(file, "wb") as fh:
                    dump(self.dataset, fh)
        self._size = size
        return self

    def with_mirroring(self, should_mirror):
        self._mirroring = should_mirror
        return self

    def with_batching(self, batch_size):
        self._batching = batch_size
        return self

    def with_type_switching(self, should_switch, switch_limit):
        self._type_switching = should_switch
        self._type_switching_limit = min(switch_limit, math.factorial(self.cfg.types))
        return self

    # ---- device-side assembly of the (augmented) sample set
    def tensors(self):
        """-> (observations int64 [M, R, C], policies float64 [M, A], values float64 [M]) on the device, in the
        reference's order: the samples, then every type-switched copy, then the mirror image of all of them."""
        n = self._size
        obs = np.stack(self.dataset["observations"][:n]).astype(np.int64)
        pol = torch.as_tensor(np.stack(self.dataset["policies"][:n]))
        val = torch.as_tensor(np.asarray(self.dataset["values"][:n], dtype=np.float64))
        base = BatchedBoards(self.cfg, n, self.moves, device=self.device, refill="philox", key=0, arrays=obs)
        dev = base.device
        pol, val = pol.to(dev), val.to(dev)
        parts = [base]
        if self._type_switching:
            limit = self._type_switching_limit
            if limit <= 0:
                limit = math.factorial(self.cfg.types)
            perms = islice(permutations(range(1, self.cfg.types + 1)), 1, limit)  # skip the identity
            parts += [augment_boards(base, type_perm=p) for p in perms]
        pols = [pol] * len(parts)
        vals = [val] * len(parts)
        if self._mirroring:
            mp = mirror_policies(self.cfg, pol)
            parts += [augment_boards(b, mirror=True) for b in list(parts)]
            pols += [mp] * (len(parts) - len(pols))
            vals += [val] * (len(parts) - len(vals))
        return torch.cat([b.array for b in parts]), torch.cat(pols), torch.cat(vals)

    def get_split(self, split=0.8, generator: torch.Generator = None):
        if not (0 < split < 1):
            raise ValueError("Split value must be between 0 and 1.")
        obs, pol, val = self.tensors()
        val = val / max(self.dataset["values"])  # dataset.py:193
        idx = torch.randperm(obs.shape[0], device=obs.device, generator=generator)
        obs, pol, val = obs[idx], pol[idx], val[idx]
        cut = int(obs.shape[0] * split)

        def batchify(o, p, v):
            k = self._batching
            return [{"observations": o[i:i + k], "policies": p[i:i + k], "values": v[i:i + k]}
                    for i in range(0, o.shape[0], k)]

        return batchify(obs[:cut], pol[:cut], val[:cut]), batchify(obs[cut:], pol[cut:], val[cut:])
