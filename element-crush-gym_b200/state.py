"""BoardV2-shaped single-board view (match3tile/boardv2.py:11-226) over the batched engine, so code
written against the reference's `State` ABC (mctslib/abc/mcts.py:8-30) -- BaseMCTS, Node.expand,
MCTS.rollout, samplerTasks.random_task/greedy_test -- runs unmodified with the GPU doing the stepping.

Functional like the reference: apply_action never mutates; it returns a new state that owns its board.
A state is two small immutable device buffers (one packed board, one packed legal mask), so clone() shares
them and apply_action is ONE ecg_step launch (out of place, the action read from a device-resident table)
followed by ONE device-to-host read of (reward, cascades, status, words drawn).

The reference's RNG semantics are kept: np.random.seed(cfg.seed) restarts the refill stream at every step
(boardv2.py:46) -- the device replays the MT19937 stream of cfg.seed -- and, because the reference draws from
the GLOBAL numpy generator, apply_action leaves the host generator exactly where the reference leaves it
(reseeded and advanced by the words the step drew), so an unmodified `np.random.choice(state.legal_actions)`
loop (samplerTasks.py:9-14, mctslib/standard/mcts.py:16-18) picks the same actions as with the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from .boards import BatchedBoards, _ptr, _stream
from .config import BoardConfig

_streams: dict = {}   # (seed, device, stream_len) -> raw MT19937 words of np.random.seed(seed), on the device
_action_ids: dict = {}  # (device, action_space) -> int32 arange: ecg_step reads the action from here, no upload


def _shared_stream(seed: int, device: torch.device, stream_len: int) -> torch.Tensor:
    key = (int(seed), str(device), int(stream_len))
    t = _streams.get(key)
    if t is None:
        sd = torch.tensor([int(seed) & 0xFFFFFFFF], dtype=torch.int64).to(device).to(torch.int32)
        t = torch.empty(stream_len, dtype=torch.int32, device=device)
        N.check(N.lib().ecg_mt19937_stream(_ptr(sd), _ptr(t), stream_len, 1, _stream(device)), "ecg_mt19937_stream")
        if len(_streams) > 64:
            _streams.clear()
        _streams[key] = t
    return t


def _actions_table(device: torch.device, action_space: int) -> torch.Tensor:
    key = (str(device), int(action_space))
    t = _action_ids.get(key)
    if t is None:
        t = _action_ids[key] = torch.arange(action_space, dtype=torch.int32, device=device)
    return t


class BoardV2:
    def __init__(self, n_actions: int, cfg: BoardConfig = None, array=None, *, device=None, stream_len: int = 8192,
                 _packed=None):
        if not torch.cuda.is_available():
            raise N.EcgError("BoardV2 needs a CUDA device (no CPU fallback)")
        self.cfg = cfg if cfg is not None else BoardConfig()
        self.n_actions = n_actions
        self._reward = 0
        self._actions = []
        self._L = N.lib()
        self._nat = self.cfg.native
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.stream_len = int(stream_len)
        self._stream = _shared_stream(self.cfg.seed, self.device, self.stream_len)
        self.stream_pos = 0  # words drawn by the step that produced this state (since its last reseed)
        self.last_cascades = 0
        self.last_status = 0
        if _packed is not None:
            self._boards, self._mask, self._mask_valid = _packed
            return
        self._boards = torch.zeros(N.TILE * self._nat.board_words, dtype=torch.int32, device=self.device)
        self._mask = torch.zeros(N.TILE * self._nat.mask_words, dtype=torch.int32, device=self.device)
        self._mask_valid = False
        if array is not None:  # boardv2.py:17-18
            a = torch.as_tensor(np.asarray(array, dtype=np.int64)[None]).to(self.device).contiguous()
            if tuple(a.shape) != (1, self.cfg.rows, self.cfg.columns):
                raise ValueError(f"expected shape {(self.cfg.rows, self.cfg.columns)}")
            st = torch.zeros(1, dtype=torch.uint8, device=self.device)
            N.check(self._L.ecg_pack(C.byref(self._nat), _ptr(a), 8, _ptr(self._boards), _ptr(st), 1,
                                     _stream(self.device)), "ecg_pack")
            if int(st.item()) & N.ST_BAD_CELL:
                raise ValueError("cell value outside {0, 1..min(type_mask, 11), h_line, v_line, bomb, mega_token}")
        else:  # boardv2.py:20-27: np.random.seed(cfg.seed), draw, redraw matched cells
            rf = self._refill(None)
            N.check(self._L.ecg_init_boards(C.byref(self._nat), C.byref(rf), _ptr(self._boards), None, 1,
                                            _stream(self.device)), "ecg_init_boards")

    # ------------------------------------------------------------------ plumbing
    def _refill(self, pos_out) -> N.Refill:
        rf = N.Refill()
        rf.mode = N.REFILL_REPLAY
        rf.stream = self._stream.data_ptr()
        rf.stream_len = self.stream_len
        rf.stream_stride = 0
        if pos_out is not None:
            rf.stream_pos = pos_out.data_ptr()
        return rf

    def _ensure_mask(self):
        if not self._mask_valid:
            N.check(self._L.ecg_legal_mask(C.byref(self._nat), _ptr(self._boards), _ptr(self._mask), 1,
                                           _stream(self.device)), "ecg_legal_mask")
            self._mask_valid = True
        return self._mask

    @property
    def _b(self) -> BatchedBoards:
        """a one-board BatchedBoards COPY of this state (replay mode, stream of cfg.seed)"""
        b = BatchedBoards(self.cfg, 1, self.n_actions, device=self.device, refill="replay", seeds=[self.cfg.seed],
                          stream_len=self.stream_len, _empty=True)
        b.stream, b.stream_len, b.stream_stride = self._stream, self.stream_len, 0
        if b.two_kernel_step:
            b._build_tiles(1)
        b.stream_pos = torch.full((1,), self.stream_pos, dtype=torch.int32, device=self.device)
        b.boards.copy_(self._boards)
        b.mask.copy_(self._ensure_mask())
        b._mask_valid = True
        b.score.fill_(int(self._reward))
        return b

    # ------------------------------------------------------------------ State ABC / BoardV2 surface
    @property
    def array(self) -> np.ndarray:
        out = torch.empty((1, self.cfg.rows, self.cfg.columns), dtype=torch.int64, device=self.device)
        N.check(self._L.ecg_unpack(C.byref(self._nat), _ptr(self._boards), _ptr(out), 8, 1, _stream(self.device)),
                "ecg_unpack")
        return out[0].cpu().numpy()

    @property
    def legal_actions(self):
        if len(self._actions) == 0:  # boardv2.py:33: cached only when non-empty
            out = torch.empty((1, self.cfg.action_space), dtype=torch.bool, device=self.device)
            N.check(self._L.ecg_unpack_mask(C.byref(self._nat), _ptr(self._ensure_mask()), _ptr(out), 1,
                                            _stream(self.device)), "ecg_unpack_mask")
            self._actions = torch.nonzero(out[0].cpu(), as_tuple=False).flatten().tolist()
        return self._actions

    def clone(self) -> "BoardV2":
        c = BoardV2(self.n_actions, self.cfg, device=self.device, stream_len=self.stream_len,
                    _packed=(self._boards, self._mask, self._mask_valid))  # immutable buffers: shared
        c._reward = self._reward
        c._actions = self._actions  # shared, like boardv2.py:40
        c.stream_pos, c.last_cascades, c.last_status = self.stream_pos, self.last_cascades, self.last_status
        return c

    def _step(self, actions: torch.Tensor, n: int, src: torch.Tensor = None):
        """ecg_step of this board with n explicit actions (n > 1: the same board for every action, src = zeros).
        -> (boards, mask, res) with res int32 [4, n]: reward, cascades, status, words drawn."""
        dev = self.device
        boards = torch.empty(-(-n // N.TILE) * N.TILE * self._nat.board_words, dtype=torch.int32, device=dev)
        mask = torch.empty(-(-n // N.TILE) * N.TILE * self._nat.mask_words, dtype=torch.int32, device=dev)
        res = torch.zeros((4, n), dtype=torch.int32, device=dev)
        st8 = torch.zeros(n, dtype=torch.uint8, device=dev) if n > 1 else None
        io = N.StepIO()
        io.boards_in = self._boards.data_ptr()
        io.boards_out = boards.data_ptr()
        io.actions = actions.data_ptr()
        io.reward = res[0].data_ptr()
        io.cascades = res[1].data_ptr()
        io.status = st8.data_ptr() if st8 is not None else res[2].data_ptr()  # one byte into a zeroed little-endian word
        io.mask_out = mask.data_ptr()
        io.env_goal = 2 ** 31 - 1
        if src is not None:
            io.src_index = src.data_ptr()
        rf = self._refill(res[3])
        N.check(self._L.ecg_step(C.byref(self._nat), C.byref(rf), C.byref(io), n, _stream(dev)), "ecg_step")
        if st8 is not None:
            res[2].copy_(st8)
        return boards, mask, res

    def apply_action(self, action) -> "BoardV2":
        if self.is_terminal:  # boardv2.py:44-45
            return self
        if int(action) not in self.cfg.actions:  # boardv2.py:48
            raise KeyError(action)
        table = _actions_table(self.device, self.cfg.action_space)
        boards, mask, res = self._step(table[int(action):], 1)
        # the new state's legal set rides along: (reward, cascades, status, words drawn) and the action-ordered
        # legal bytes leave the device in ONE copy -- the one synchronisation of the step
        A = self.cfg.action_space
        out = torch.empty(16 + A, dtype=torch.uint8, device=self.device)
        out[:16].view(torch.int32).copy_(res.reshape(-1))
        N.check(self._L.ecg_unpack_mask(C.byref(self._nat), _ptr(mask), C.c_void_p(out.data_ptr() + 16), 1,
                                        _stream(self.device)), "ecg_unpack_mask")
        host = out.cpu()
        reward, cascades, status, drawn = host[:16].view(torch.int32).tolist()
        nxt = BoardV2(self.n_actions - 1, self.cfg, device=self.device, stream_len=self.stream_len,
                      _packed=(boards, mask, True))
        nxt._actions = torch.nonzero(host[16:], as_tuple=False).flatten().tolist()
        nxt._reward = self._reward + reward
        nxt.stream_pos, nxt.last_cascades, nxt.last_status = drawn, cascades, status
        # the reference drew from the global generator after np.random.seed(cfg.seed) (boardv2.py:46, :172;
        # boardFunctions.py:17-22): leave the host generator in the same state
        np.random.seed(self.cfg.seed)
        if drawn:
            np.random.randint(0, 2 ** 32, size=drawn, dtype=np.uint32)  # one raw MT19937 word each
        return nxt

    @property
    def greedy_action(self):
        """argmax over legal actions of the one-step cumulative reward, first maximum wins (boardv2.py:209-218);
        every legal child in ONE launch."""
        legal = self.legal_actions
        if self.is_terminal or not legal:
            # the reference loops over legal_actions with apply_action returning self on a terminal board
            return legal[0] if legal and self._reward > -1 else None
        acts = torch.tensor(legal, dtype=torch.int32).to(self.device)
        src = torch.zeros(len(legal), dtype=torch.int32, device=self.device)
        _, _, res = self._step(acts, len(legal), src)
        rewards = res[0].tolist()
        best_action, highest = None, -1
        for a, r in zip(legal, rewards):
            if self._reward + r > highest:
                highest, best_action = self._reward + r, a
        # like the reference, which steps every child with np.random.seed(cfg.seed): the last child's draws remain
        np.random.seed(self.cfg.seed)
        drawn = int(res[3][-1].item())
        if drawn:
            np.random.randint(0, 2 ** 32, size=drawn, dtype=np.uint32)
        return best_action

    @property
    def is_terminal(self) -> bool:
        return self.n_actions < 1

    @property
    def reward(self):
        return self._reward
