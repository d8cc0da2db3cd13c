"""Standard MCTS with GPU-batched random rollouts (BASELINE.json configs[4]).

Host-side tree exactly as mctslib/abc/mcts.py:33-130 + mctslib/standard/mcts.py:9-42 (UCB1 with
c = node.state.n_actions, expansion pops the largest untried action, most-visited child is played, the tree
is re-rooted at it; pinned against the reference's own tree by tests/golden/mcts_tree.json), but the simulation
step evaluates the expanded node with `leaves` random rollouts instead of one Python rollout; back-propagation
then adds `leaves` visits and the sum of their final rewards.  With several ranks every rank rolls out its share
and the (reward sum, env-steps) pair is all-reduced over NCCL (dist.reduce_visit_counts) -- the only
communication.  The tree policy needs every simulation's value before the next selection, so there is exactly one
device-to-host read per simulation (the reduced reward sum); visit counts are known on the host.

Two rollout dynamics:

refill="philox" (default, the throughput mode): `leaves` INDEPENDENT episodes in one ecg_rollout launch, refill
    tiles and picks from Philox substreams.  This is NOT the reference's rollout distribution: the reference reseeds
    the refill generator with cfg.seed at every step (boardv2.py:46), which makes its steps cascade longer --
    measured on 9x9x6, 20-move random episodes: 1.79-1.81 cascade iterations and 24.4 reward per step with the
    reference's semantics against 1.55 and 17.3 with i.i.d. Philox tiles; episode reward 476 +- 231 against
    345 +- 144.  Bit-exact against the CPU oracle's Philox episodes (tests/test_gpu_mcts.py), not comparable
    with the reference's values.

refill="replay" (the reference's dynamics, A12): mctslib/standard/mcts.py:14-19 seeds numpy from Python's `random`,
    draws the FIRST action with it, and from then on every pick and every refill comes from the MT19937 stream of
    cfg.seed, restarted by each apply_action -- a rollout is a deterministic function of its first action.  The
    engine therefore plays every legal first action once (one expansion launch + one ecg_rollout launch on the
    shared MT(cfg.seed) stream, bit-exact with the reference's episodes) and draws the `leaves` first picks
    uniformly over them; `deterministic=True` takes the first pick from MT(cfg.seed) as the reference intended
    (its `state.seed` does not exist, standard/mcts.py:15).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

import ctypes as C

from . import _native as N
from . import dist as ecg_dist
from .boards import BatchedBoards, _ptr, _stream
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


class _LeafBuffers:
    """device buffers of one rank's share of the leaves, allocated once per search: a simulation only refills them
    (one copy of the leaf board into every lane, one fill of moves_left) and launches ecg_rollout"""

    def __init__(self, cfg, count: int, device):
        self.count = count
        L = N.lib()
        nat = cfg.native
        self.boards = torch.zeros(L.ecg_boards_bytes(C.byref(nat), count) // 4, dtype=torch.int32, device=device)
        self.moves_left = torch.zeros(count, dtype=torch.int32, device=device)
        self.total = torch.zeros(count, dtype=torch.int64, device=device)
        self.steps = torch.zeros(count, dtype=torch.int32, device=device)
        self.scratch = torch.empty(count + 1, dtype=torch.int32, device=device)  # work list of the two-kernel rollout


def replicate(state: BoardV2, n: int, *, key: int, board0: int, device=None) -> BatchedBoards:
    """n copies of one board as a Philox-mode batch (copy i is global board board0 + i)."""
    out = BatchedBoards(state.cfg, n, state.n_actions, device=device or state.device, refill="philox", key=key,
                        board0=board0, _empty=True)
    words = state.cfg.native.board_words
    tile = state._boards.view(words // 4, 32, 4)  # [chunk, lane, 4 words]; the board sits in lane 0
    full = tile[:, :1, :].expand(words // 4, 32, 4).reshape(-1)  # the same board in all 32 lanes
    out.boards.view(-1, 32 * words)[:] = full
    return out


class BatchedRolloutMCTS:
    """Same call shape as mctslib.standard.mcts.MCTS: `action, value, policies = mcts()`."""

    def __init__(self, state: BoardV2, exploration_weight: float, simulations: int, verbose: bool = False,
                 deterministic: bool = False, *, leaves: int = 1 << 20, key: int = 0x5EED, refill: str = "philox"):
        if refill not in ("philox", "replay"):
            raise ValueError("refill must be 'philox' or 'replay'")
        self._root = Node(state)
        self._simulations = simulations
        self._verbose = verbose
        self._exploration_weight = exploration_weight  # stored, unused -- like abc/mcts.py:80,95
        self.deterministic = deterministic
        self.leaves = int(leaves)
        self.key = int(key)
        self.refill = refill
        self._sim_counter = 0
        self.env_steps = 0  # env-steps simulated on the devices so far, over all ranks
        self._leaf = None   # _LeafBuffers of this rank
        self._root.expand()  # abc/mcts.py:82

    def _first_action_returns(self, state: BoardV2):
        """refill="replay": final cumulative reward of the reference's rollout for every legal first action of
        `state`, int64 [L] on the device (standard/mcts.py:16-18 over boardv2.py:46), and the env-steps simulated."""
        legal = state.legal_actions
        dev = state.device
        n = len(legal)
        acts = torch.tensor(legal, dtype=torch.int32).to(dev)
        boards, mask, res = state._step(acts, n, torch.zeros(n, dtype=torch.int32, device=dev))
        rest = BatchedBoards(state.cfg, n, state.n_actions - 1, device=dev, refill="replay", _empty=True)
        rest.stream, rest.stream_len, rest.stream_stride = state._stream, state.stream_len, 0  # MT(cfg.seed), shared
        rest.two_kernel_step = False  # rollouts run in the exact kernel: no tile tables needed
        rest.stream_pos = res[3].contiguous()  # each episode continues where its first step left the stream
        rest.boards = boards
        total = rest.rollout()
        return total + res[0].to(torch.int64) + int(state.reward), (rest.rollout_steps.sum() + n).to(torch.int64)

    def rollout(self, state: BoardV2):
        """-> (sum of final rewards, number of rollouts) over all ranks"""
        rank, world = 0, 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        first, count = ecg_dist.shard_range(self.leaves, world, rank)
        sim = self._sim_counter
        self._sim_counter += 1
        dev = state.device
        steps = torch.zeros((), dtype=torch.int64, device=dev)
        if count == 0 or state.is_terminal or (self.refill == "replay" and not state.legal_actions):
            rsum = torch.tensor(int(state.reward) * count, dtype=torch.int64, device=dev)
        elif self.refill == "philox":
            if self._leaf is None or self._leaf.count != count:
                self._leaf = _LeafBuffers(state.cfg, count, dev)
            lb, nat = self._leaf, state.cfg.native
            words = nat.board_words
            tile = state._boards.view(words // 4, 32, 4)  # [chunk, lane, 4 words]; the board sits in lane 0
            lb.boards.view(-1, 32 * words)[:] = tile[:, :1, :].expand(words // 4, 32, 4).reshape(-1)
            lb.moves_left.fill_(state.n_actions)
            rf = N.Refill()
            rf.mode, rf.philox_key, rf.board0, rf.step_ctr = N.REFILL_PHILOX, self.key, sim * self.leaves + first, 0
            N.check(N.lib().ecg_rollout_scratch(C.byref(nat), C.byref(rf), _ptr(lb.boards), _ptr(lb.moves_left),
                                                _ptr(lb.total), _ptr(lb.steps), None, _ptr(lb.scratch), count,
                                                _stream(dev)), "ecg_rollout")
            rsum = lb.total.sum() + int(state.reward) * count
            steps = lb.steps.sum().to(torch.int64)
        else:
            returns, steps = self._first_action_returns(state)  # [L]; every rank plays the same few episodes
            if rank:
                steps = torch.zeros_like(steps)  # counted once
            n = returns.numel()
            if self.deterministic:  # np.random.seed(cfg.seed); np.random.choice(legal): one fixed first pick
                import numpy as np
                picks = torch.zeros(n, dtype=torch.int64, device=dev)
                picks[int(np.random.RandomState(state.cfg.seed).choice(n))] = count
            else:  # the reference seeds this pick from Python's `random`: uniform over the legal set
                import numpy as np
                g = np.random.Generator(np.random.PCG64((self.key * 1000003 + sim * 8191 + rank) & (2 ** 63 - 1)))
                picks = torch.from_numpy(g.multinomial(count, [1.0 / n] * n).astype(np.int64)).to(dev)
            rsum = (picks * returns).sum()
        # one reduction, one device-to-host read per simulation: (reward sum, env-steps); visits are known on the host
        both = torch.stack([rsum.to(torch.int64).reshape(()), steps.reshape(())])
        ecg_dist.reduce_visit_counts(both, both[:0])
        rsum_h, steps_h = both.tolist()
        self.env_steps += steps_h
        return rsum_h, self.leaves

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
