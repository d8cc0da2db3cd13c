"""Standard MCTS with GPU-batched random rollouts (BASELINE.json configs[4]).

Host-side tree exactly as mctslib/abc/mcts.py:33-130 + mctslib/standard/mcts.py:9-42 (UCB1 with
c = node.state.n_actions, expansion pops the largest untried action, most-visited child is played, the tree
is re-rooted at it), but the simulation step evaluates the expanded node with `leaves` independent random
rollouts in ONE kernel launch (ecg_rollout) instead of one Python rollout; back-propagation then adds
`leaves` visits and the sum of their final rewards.  With several ranks every rank rolls out its share and the
(visits, reward sum) pair is all-reduced over NCCL (dist.reduce_visit_counts) -- the only communication.

Rollouts use the Philox mode: like the reference's non-deterministic MCTS (standard/mcts.py:15 seeds from
`random`), results are statistically, not bit-wise, comparable with the reference; against the CPU oracle's
Philox episodes they are bit-exact (tests/test_gpu_mcts.py).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import dist as ecg_dist
from .boards import BatchedBoards
from .state import BoardV2


class Node:
    """mctslib.abc.BaseNode + mctslib.standard.Node"""

    def __init__(self, state: BoardV2, parent: Optional["Node"] = None):
        self.state = state.clone()
        self.parent = parent
        self.children: dict = {}
        self.visits = 0
        self.reward = 0
        self.untried_actions = list(state.legal_actions)

    @property
    def is_fully_expanded(self) -> bool:
        return len(self.untried_actions) == 0

    def expand(self) -> "Node":
        action = self.untried_actions.pop()  # standard/mcts.py:33
        child = Node(self.state.apply_action(action), self)
        self.children[action] = child
        return child

    @property
    def policies(self):
        return [child.visits / self.visits for child in self.children.values()]

    def update(self, reward_sum, n: int = 1):
        self.visits += n
        self.reward += reward_sum

    @property
    def exploitation(self):
        return self.reward / self.visits

    @property
    def exploration(self):
        return math.sqrt(math.log(self.parent.visits) / (1 + self.visits))

    def ucb1(self, c: float):
        if self.visits == 0:
            return float("inf")
        return self.exploitation + c * self.exploration

    def best_child(self, c):
        return max(self.children.values(), key=lambda child: child.ucb1(c))


def replicate(state: BoardV2, n: int, *, key: int, board0: int, device=None) -> BatchedBoards:
    """n copies of one board as a Philox-mode batch (copy i is global board board0 + i)."""
    src = state._b
    out = BatchedBoards(state.cfg, n, state.n_actions, device=device or src.device, refill="philox", key=key,
                        board0=board0, _empty=True)
    words = state.cfg.native.board_words
    tile = src.boards[: 32 * words].view(words // 4, 32, 4)  # [chunk, lane, 4 words]; the board sits in lane 0
    full = tile[:, :1, :].expand(words // 4, 32, 4).reshape(-1)  # the same board in all 32 lanes
    out.boards.view(-1, 32 * words)[:] = full
    return out


class BatchedRolloutMCTS:
    """Same call shape as mctslib.standard.mcts.MCTS: `action, value, policies = mcts()`."""

    def __init__(self, state: BoardV2, exploration_weight: float, simulations: int, verbose: bool = False,
                 deterministic: bool = False, *, leaves: int = 1 << 20, key: int = 0x5EED):
        self._root = Node(state)
        self._simulations = simulations
        self._verbose = verbose
        self._exploration_weight = exploration_weight  # stored, unused -- like abc/mcts.py:80,95
        self.deterministic = deterministic
        self.leaves = int(leaves)
        self.key = int(key)
        self._sim_counter = 0
        self.env_steps = 0
        self._root.expand()  # abc/mcts.py:82

    def rollout(self, state: BoardV2):
        """-> (sum of final rewards, number of rollouts) over all ranks"""
        rank, world = 0, 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        first, count = ecg_dist.shard_range(self.leaves, world, rank)
        base = self._sim_counter * self.leaves
        self._sim_counter += 1
        dev = state._b.device
        if count > 0 and not state.is_terminal:
            batch = replicate(state, count, key=self.key, board0=base + first)
            total = batch.rollout()
            rsum = total.sum() + state.reward * count
            self.env_steps += int(batch.rollout_steps.sum().item())
        else:
            rsum = torch.tensor(state.reward * count, dtype=torch.int64, device=dev)
        visits = torch.tensor([count], dtype=torch.int64, device=dev)
        rsum = rsum.reshape(1).to(torch.int64)
        ecg_dist.reduce_visit_counts(visits, rsum)
        return int(rsum.item()), int(visits.item())

    def __call__(self):
        node = self._root
        for _ in range(self._simulations):
            while not node.state.is_terminal and node.is_fully_expanded:  # selection, abc/mcts.py:94-95
                node = node.best_child(node.state.n_actions)
            if not node.state.is_terminal and not node.is_fully_expanded:  # expansion
                node = node.expand()
            reward_sum, n = self.rollout(node.state)  # simulation (batched)
            while node is not None:  # backpropagation
                node.update(reward_sum, n)
                node = node.parent
            node = self._root
        action, best_child = max(self._root.children.items(), key=lambda kv: kv[1].visits)
        policies = self._root.policies
        self.last_root_actions = list(self._root.children.keys())  # the action behind each entry of `policies`
        node = self._root
        while not node.state.is_terminal and node.is_fully_expanded:
            node = node.best_child(0)
        value = node.state.reward
        best_child.parent = None  # tree reuse, abc/mcts.py:123-124
        self._root = best_child
        return action, value, policies
