"""BoardV2-shaped single-board view (match3tile/boardv2.py:11-226) over the batched engine, so code
written against the reference's `State` ABC (mctslib/abc/mcts.py:8-30) -- BaseMCTS, Node.expand,
MCTS.rollout, samplerTasks.random_task/greedy_test -- runs unmodified with the GPU doing the stepping.

Functional like the reference: apply_action never mutates; it returns a new state that owns its board.
The reference's RNG semantics are kept (np.random.seed(cfg.seed) restarts the refill stream at every
step, boardv2.py:46) by replaying the MT19937 stream of cfg.seed, generated on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from .boards import BatchedBoards
from .config import BoardConfig


class BoardV2:
    def __init__(self, n_actions: int, cfg: BoardConfig = None, array=None, *, device=None, stream_len: int = 8192,
                 _boards: BatchedBoards = None):
        self.cfg = cfg if cfg is not None else BoardConfig()
        self.n_actions = n_actions
        self._reward = 0
        if _boards is not None:
            self._b = _boards
        else:
            arrays = None if array is None else np.asarray(array, dtype=np.int64)[None]
            self._b = BatchedBoards(self.cfg, 1, n_actions, device=device, refill="replay", seeds=[self.cfg.seed],
                                    stream_len=stream_len, arrays=arrays)
        self._actions = []

    @property
    def array(self) -> np.ndarray:
        return self._b.array[0].cpu().numpy()

    @property
    def legal_actions(self):
        if len(self._actions) == 0:  # boardv2.py:33: cached only when non-empty
            self._actions = self._b.legal_actions[0]
        return self._actions

    def clone(self) -> "BoardV2":
        c = BoardV2(self.n_actions, self.cfg, _boards=self._b.clone())
        c._reward = self._reward
        c._actions = self._actions  # shared, like boardv2.py:40
        return c

    def apply_action(self, action) -> "BoardV2":
        if self.is_terminal:  # boardv2.py:44-45
            return self
        if int(action) not in self.cfg.actions:  # boardv2.py:48
            raise KeyError(action)
        nb = self._b.clone()
        nb.moves_left.fill_(self.n_actions)
        nb.apply_action(torch.tensor([int(action)], dtype=torch.int32))
        nxt = BoardV2(self.n_actions - 1, self.cfg, _boards=nb)
        nxt._reward = self._reward + int(nb.step_reward[0].item())
        nxt.last_cascades = int(nb.cascades[0].item())
        nxt.last_status = int(nb.status[0].item())
        return nxt

    @property
    def greedy_action(self):
        best_action, highest = None, -1  # boardv2.py:209-218
        for action in self.legal_actions:
            r = self.apply_action(action).reward
            if r > highest:
                highest, best_action = r, action
        return best_action

    @property
    def is_terminal(self) -> bool:
        return self.n_actions < 1

    @property
    def reward(self):
        return self._reward
