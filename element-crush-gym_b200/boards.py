"""BatchedBoards: N match-3 boards resident on one GPU, stepped in lockstep by libecg.so.

This is the batched counterpart of the reference's `BoardV2` (match3tile/boardv2.py:11-226): the same
vocabulary (`array`, `legal_actions`, `apply_action`, `reward`, `n_actions`, `is_terminal`, `clone`,
`greedy_action`) with a leading board dimension.  PyTorch tensors are only the device buffers; all
board logic runs in the CUDA kernels behind the C-ABI (include/ecg.h).
"""
from __future__ import annotations

import ctypes as C
import os
import secrets

import torch

from . import _native as N
from .config import BoardConfig

_I32_MAX = 2 ** 31 - 1


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class BatchedBoards:
    """num_boards boards of one BoardConfig on `device`.

    refill="philox": refill tiles, shuffles and random picks come from Philox4x32-10 keyed by
        (key, board0 + i, step) -- results do not depend on how the batch is sharded over GPUs.
    refill="replay": board i replays numpy's legacy MT19937 stream of seeds[i] exactly like the
        reference (np.random.seed(cfg.seed) at the top of every apply_action, boardv2.py:46), which
        makes trajectories bit-identical to BoardV2 / samplerTasks.random_task.  One seed = one stream
        shared by all boards (every state of one BoardConfig); stream_index[i] picks among len(seeds) streams.
    """

    def __init__(self, cfg: BoardConfig, num_boards: int, n_actions: int = 20, *, device=None, refill="philox",
                 key=None, seeds=None, board0: int = 0, stream_len: int = 4096, arrays=None,
                 env_goal: int = _I32_MAX, stream_index=None, _empty=False):
        if not torch.cuda.is_available():
            raise N.EcgError("BatchedBoards needs a CUDA device (no CPU fallback)")
        self.L = N.lib()
        self.cfg = cfg
        self.nat = cfg.native
        self.n = int(num_boards)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.env_goal = int(env_goal)
        self.refill_mode = refill
        self.board0 = int(board0)
        self.step_ctr = 0
        self.key = int(key if key is not None else (cfg.seed if refill == "philox" else 0)) & (2 ** 64 - 1)
        self.stream = None
        self.stream_pos = None
        self.stream_index = None  # replay: int32 [n], board i replays stream stream_index[i] (children of expand())
        self.tiles = self.tile_wpos = None  # replay: per-stream tile tables (ecg_replay_tiles) of the two-kernel step
        self.stream_len = 0
        self.stream_stride = 0
        self._moves_bound = max(int(n_actions), 0)  # host-side upper bound of moves_left (no device read needed)
        dev = self.device
        nb = self.L.ecg_boards_bytes(C.byref(self.nat), self.n) // 4
        nm = self.L.ecg_masks_bytes(C.byref(self.nat), self.n) // 4
        self.boards = torch.zeros(nb, dtype=torch.int32, device=dev)
        self.mask = torch.zeros(nm, dtype=torch.int32, device=dev)
        self._mask_valid = False
        self.moves_left = torch.full((self.n,), int(n_actions), dtype=torch.int32, device=dev)
        self.score = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self.step_reward = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self.cascades = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self.flags = torch.zeros(self.n, dtype=torch.uint8, device=dev)
        self.status = torch.zeros(self.n, dtype=torch.uint8, device=dev)
        self.last_actions = torch.full((self.n,), -1, dtype=torch.int32, device=dev)
        # work list of the two-kernel step (ecg_step_io.scratch); ECG_SINGLE_KERNEL=1 keeps the one-kernel step
        self.two_kernel_step = (os.environ.get("ECG_SINGLE_KERNEL", "0") != "1" and
                                (refill == "philox" or stream_len < 65535))
        self._scratch = torch.empty(self.n + 1, dtype=torch.int32, device=dev) if self.two_kernel_step else None
        if _empty:
            return
        if refill == "replay":
            if seeds is None:
                seeds = [cfg.seed] * 1  # one shared stream: every board replays cfg.seed (MCTS leaves do)
            seeds_t = torch.as_tensor(seeds, dtype=torch.int64).reshape(-1)
            if stream_index is not None:  # any number of streams, board i replays stream stream_index[i]
                self.stream_index = torch.as_tensor(stream_index).to(device=dev, dtype=torch.int32).contiguous()
                if self.stream_index.numel() != self.n:
                    raise ValueError("stream_index must hold one stream id per board")
            elif seeds_t.numel() not in (1, self.n):
                raise ValueError("seeds must hold one seed (shared stream) or one per board (or pass stream_index)")
            self.seeds = seeds_t.clone()
            sd = seeds_t.to(dev).to(torch.int32)
            self.stream_len = int(stream_len)
            self.stream = torch.empty(sd.numel() * self.stream_len, dtype=torch.int32, device=dev)
            N.check(self.L.ecg_mt19937_stream(_ptr(sd), _ptr(self.stream), self.stream_len, sd.numel(), _stream(dev)),
                    "ecg_mt19937_stream")
            self.stream_stride = self.stream_len if (sd.numel() == self.n and self.n > 1) or stream_index is not None else 0
            if self.two_kernel_step:
                self._build_tiles(sd.numel())
            self.stream_pos = torch.zeros(self.n, dtype=torch.int32, device=dev)
        elif refill != "philox":
            raise ValueError("refill must be 'philox' or 'replay'")
        if arrays is not None:
            self.set_arrays(arrays)
        else:
            rf = self._refill()
            N.check(self.L.ecg_init_boards(C.byref(self.nat), C.byref(rf), _ptr(self.boards), _ptr(self.status),
                                           self.n, _stream(dev)), "ecg_init_boards")

    # ------------------------------------------------------------------ plumbing
    def _build_tiles(self, n_streams: int):
        """tile tables of the replay streams for this config's `types` (the two-kernel replay step reads its refill
        tiles from them instead of rejection-sampling raw words)"""
        dev = self.device
        tw = int(self.L.ecg_replay_tiles_words(self.stream_len))
        self.tiles = torch.empty(n_streams * tw, dtype=torch.int32, device=dev)
        self.tile_wpos = torch.empty(n_streams * (self.stream_len + 1), dtype=torch.int16, device=dev)
        N.check(self.L.ecg_replay_tiles(_ptr(self.stream), self.stream_len, self.cfg.types, _ptr(self.tiles),
                                        _ptr(self.tile_wpos), n_streams, _stream(dev)), "ecg_replay_tiles")

    def _refill(self, step_ctr=None) -> N.Refill:
        rf = N.Refill()
        if self.refill_mode == "philox":
            rf.mode = N.REFILL_PHILOX
            rf.philox_key = self.key
            rf.board0 = self.board0
            rf.step_ctr = (self.step_ctr if step_ctr is None else step_ctr) & 0xFFFFFFFF
        else:
            rf.mode = N.REFILL_REPLAY
            rf.stream = self.stream.data_ptr()
            rf.stream_len = self.stream_len
            rf.stream_stride = self.stream_stride
            rf.stream_pos = self.stream_pos.data_ptr()
            if self.stream_index is not None:
                rf.stream_index = self.stream_index.data_ptr()
            if self.tiles is not None and self.two_kernel_step:
                rf.tiles = self.tiles.data_ptr()
                rf.tile_wpos = self.tile_wpos.data_ptr()
        return rf

    def set_arrays(self, arrays):
        """Load boards from [N, rows, cols] cell values (BoardV2.array convention)."""
        a = torch.as_tensor(arrays)
        if a.dtype not in (torch.int64, torch.uint8):
            a = a.to(torch.int64)
        a = a.to(self.device).contiguous()
        if tuple(a.shape) != (self.n, self.cfg.rows, self.cfg.columns):
            raise ValueError(f"expected shape {(self.n, self.cfg.rows, self.cfg.columns)}, got {tuple(a.shape)}")
        N.check(self.L.ecg_pack(C.byref(self.nat), _ptr(a), a.element_size(), _ptr(self.boards), _ptr(self.status),
                                self.n, _stream(self.device)), "ecg_pack")
        self._mask_valid = False
        if bool((self.status & N.ST_BAD_CELL).any()):
            raise ValueError("cell value outside {0, 1..min(type_mask, 11), h_line, v_line, bomb, mega_token}")

    # ------------------------------------------------------------------ BoardV2 surface
    @property
    def array(self) -> torch.Tensor:
        """int64 [N, rows, cols]: BoardV2.array of every board"""
        return self.observe(torch.int64)

    def observe(self, dtype=torch.uint8, out=None) -> torch.Tensor:
        if dtype not in (torch.uint8, torch.int64):
            raise ValueError("dtype must be torch.uint8 or torch.int64")
        if out is None:
            out = torch.empty((self.n, self.cfg.rows, self.cfg.columns), dtype=dtype, device=self.device)
        N.check(self.L.ecg_unpack(C.byref(self.nat), _ptr(self.boards), _ptr(out), out.element_size(), self.n,
                                  _stream(self.device)), "ecg_unpack")
        return out

    def packed_mask(self) -> torch.Tensor:
        """packed legal mask of the current boards (computed by the step kernel; recomputed only after
        set_arrays / construction)"""
        if not self._mask_valid:
            N.check(self.L.ecg_legal_mask(C.byref(self.nat), _ptr(self.boards), _ptr(self.mask), self.n,
                                          _stream(self.device)), "ecg_legal_mask")
            self._mask_valid = True
        return self.mask

    def legal_mask(self, out=None) -> torch.Tensor:
        """bool [N, action_space]; row i is the membership vector of BoardV2.legal_actions of board i"""
        m = self.packed_mask()
        if out is None:
            out = torch.empty((self.n, self.cfg.action_space), dtype=torch.bool, device=self.device)
        N.check(self.L.ecg_unpack_mask(C.byref(self.nat), _ptr(m), _ptr(out), self.n, _stream(self.device)),
                "ecg_unpack_mask")
        return out

    @property
    def legal_actions(self):
        """list (per board) of ascending action lists, like BoardV2.legal_actions (boardv2.py:31-35)"""
        m = self.legal_mask().cpu()
        return [torch.nonzero(row, as_tuple=False).flatten().tolist() for row in m]

    def random_action(self, out=None) -> torch.Tensor:
        """int32 [N]: a uniformly random legal action per board (README's env.board.random_action(),
        = np.random.choice(state.legal_actions), samplerTasks.py:13); -1 where no action is legal"""
        m = self.packed_mask()
        if out is None:
            out = torch.empty(self.n, dtype=torch.int32, device=self.device)
        rf = self._refill()
        N.check(self.L.ecg_random_action(C.byref(self.nat), C.byref(rf), _ptr(m), _ptr(out), None, self.n,
                                         _stream(self.device)), "ecg_random_action")
        return out

    def apply_action(self, actions=None) -> "BatchedBoards":
        """One lockstep BoardV2.apply_action (boardv2.py:43-207) on every board, IN PLACE.
        actions: int tensor/array [N]; None = each board plays a uniformly random legal action.
        Afterwards: step_reward, score (cumulative = BoardV2.reward), cascades, moves_left, flags, status,
        last_actions and the legal mask describe the new state."""
        io = N.StepIO()
        if actions is None:
            io.mask_in = self.packed_mask().data_ptr()
            keep = None
        else:
            keep = torch.as_tensor(actions).to(device=self.device, dtype=torch.int32).contiguous()
            if keep.numel() != self.n:
                raise ValueError("one action per board")
            io.actions = keep.data_ptr()
        io.boards_in = io.boards_out = self.boards.data_ptr()
        io.actions_out = self.last_actions.data_ptr()
        io.moves_left = self.moves_left.data_ptr()
        io.reward = self.step_reward.data_ptr()
        io.score = self.score.data_ptr()
        io.cascades = self.cascades.data_ptr()
        io.mask_out = self.mask.data_ptr()
        io.flags = self.flags.data_ptr()
        io.status = self.status.data_ptr()
        io.env_goal = self.env_goal
        if self._scratch is not None:
            io.scratch = self._scratch.data_ptr()
        rf = self._refill()
        N.check(self.L.ecg_step(C.byref(self.nat), C.byref(rf), C.byref(io), self.n, _stream(self.device)), "ecg_step")
        self._mask_valid = True
        self.step_ctr += 1
        self._moves_bound = max(self._moves_bound - 1, 0)
        return self

    def rollout(self) -> torch.Tensor:
        """Play uniformly random legal actions until every board is terminal, in ONE kernel
        (MCTS.rollout, mctslib/standard/mcts.py:14-19; samplerTasks.random_task :9-14).
        Returns int64 [N] points collected; score / moves_left / boards are updated in place."""
        total = torch.empty(self.n, dtype=torch.int64, device=self.device)
        steps = torch.empty(self.n, dtype=torch.int32, device=self.device)
        rf = self._refill()
        scratch = self._scratch if self.refill_mode == "philox" else None  # work list of the two-kernel rollout
        N.check(self.L.ecg_rollout_scratch(C.byref(self.nat), C.byref(rf), _ptr(self.boards), _ptr(self.moves_left),
                                           _ptr(total), _ptr(steps), _ptr(self.status), _ptr(scratch), self.n,
                                           _stream(self.device)), "ecg_rollout")
        self.step_ctr += self._moves_bound  # no board plays more moves than that: no device read, no sync
        self._moves_bound = 0
        self.score += total.to(torch.int32)
        self.moves_left -= steps
        self._mask_valid = False
        self.rollout_steps = steps
        return total

    @property
    def reward(self) -> torch.Tensor:
        """cumulative reward per board (BoardV2.reward, boardv2.py:224-226)"""
        return self.score.to(torch.int64)

    @property
    def n_actions(self) -> torch.Tensor:
        return self.moves_left

    @property
    def is_terminal(self) -> torch.Tensor:
        """boardv2.py:220-222"""
        return self.moves_left < 1

    def clone(self) -> "BatchedBoards":
        c = BatchedBoards(self.cfg, self.n, device=self.device, refill=self.refill_mode, key=self.key,
                          board0=self.board0, env_goal=self.env_goal, _empty=True)
        for name in ("boards", "mask", "moves_left", "score", "step_reward", "cascades", "flags", "status",
                     "last_actions"):
            getattr(c, name).copy_(getattr(self, name))
        c._mask_valid = self._mask_valid
        c.step_ctr = self.step_ctr
        c.stream, c.stream_len, c.stream_stride = self.stream, self.stream_len, self.stream_stride  # read-only, shared
        c.stream_index = self.stream_index  # read-only, shared
        c.tiles, c.tile_wpos = self.tiles, self.tile_wpos
        c.stream_pos = None if self.stream_pos is None else self.stream_pos.clone()
        c._moves_bound = self._moves_bound
        return c

    def expand(self):
        """Every legal (board, action) pair stepped once, in ONE kernel (Node.expand for all children,
        mctslib/standard/mcts.py:31-42; the inner loop of greedy_action, boardv2.py:209-218).
        Returns (children: BatchedBoards with one board per pair, parent: int64 [P], action: int32 [P]);
        pairs are ordered by (board, ascending action).  A child is stepped with its parent's refill stream / Philox
        id, like the reference where every child is stepped with np.random.seed(cfg.seed); in replay mode it keeps
        replaying that stream afterwards (stream_index) from the position its step reached (stream_pos)."""
        legal = self.legal_mask() & (self.moves_left >= 1)[:, None]  # terminal boards have no children (boardv2.py:44)
        pairs = torch.nonzero(legal, as_tuple=False)  # sorted by (board, action)
        parent = pairs[:, 0].contiguous()
        action = pairs[:, 1].to(torch.int32).contiguous()
        p = int(parent.numel())
        child = BatchedBoards(self.cfg, p, device=self.device, refill=self.refill_mode, key=self.key,
                              board0=self.board0, env_goal=self.env_goal, _empty=True)
        child.stream, child.stream_len, child.stream_stride = self.stream, self.stream_len, self.stream_stride
        child.tiles, child.tile_wpos = self.tiles, self.tile_wpos
        child.step_ctr = self.step_ctr + 1
        child._moves_bound = max(self._moves_bound - 1, 0)
        src = parent.to(torch.int32).contiguous()
        if self.refill_mode == "replay":
            child.stream_pos = torch.zeros(p, dtype=torch.int32, device=self.device)
            if self.stream_stride:  # per-board streams: child j replays the stream of its parent
                child.stream_index = src if self.stream_index is None else self.stream_index[parent].contiguous()
        if p == 0:
            return child, parent, action
        io = N.StepIO()
        io.boards_in = self.boards.data_ptr()
        io.boards_out = child.boards.data_ptr()
        io.actions = action.data_ptr()
        io.src_index = src.data_ptr()
        io.actions_out = child.last_actions.data_ptr()
        io.reward = child.step_reward.data_ptr()
        io.cascades = child.cascades.data_ptr()
        io.mask_out = child.mask.data_ptr()
        io.status = child.status.data_ptr()
        io.env_goal = self.env_goal
        rf = self._refill()
        if self.refill_mode == "replay":
            # every child restarts the stream (np.random.seed at the top of apply_action); with explicit actions
            # stream_pos is output only: the words the child's step consumed, where its next random pick continues
            rf.stream_pos = child.stream_pos.data_ptr()
        N.check(self.L.ecg_step(C.byref(self.nat), C.byref(rf), C.byref(io), p, _stream(self.device)), "ecg_step")
        child.moves_left.copy_(self.moves_left[parent] - 1)
        child.score.copy_(self.score[parent] + child.step_reward)
        child._mask_valid = True
        return child, parent, action

    def greedy_action(self) -> torch.Tensor:
        """argmax over legal actions of the one-step reward, first maximum wins (boardv2.py:209-218); -1 for a
        board without legal action.  One expansion kernel + a segmented arg-max."""
        child, parent, action = self.expand()
        best = torch.full((self.n,), -1, dtype=torch.int32, device=self.device)
        if parent.numel() == 0:
            return best
        r = child.step_reward.to(torch.int64)
        top = torch.full((self.n,), -1, dtype=torch.int64, device=self.device)
        top.scatter_reduce_(0, parent, r, reduce="amax", include_self=True)
        is_top = r == top[parent]
        # first maximum = smallest action among the maxima
        big = torch.full((self.n,), 1 << 30, dtype=torch.int64, device=self.device)
        cand = torch.where(is_top, action.to(torch.int64), torch.full_like(r, 1 << 30))
        big.scatter_reduce_(0, parent, cand, reduce="amin", include_self=True)
        has = big < (1 << 30)
        best[has] = big[has].to(torch.int32)
        return best

    def observe_onehot(self, channels: int = None, dtype=torch.float32, out=None) -> torch.Tensor:
        """[N, rows, cols, channels] one-hot of the cell values = nnx.one_hot(board.array, channels) with
        channels = 2 ** (ceil(log2(types)) + 2) (elementCrush.py:66,92); values >= channels give all zeros."""
        import math
        if channels is None:
            channels = 2 ** (int(math.ceil(math.log2(self.cfg.types))) + 2)
        kind = {torch.uint8: 0, torch.float32: 1, torch.bfloat16: 2, torch.float16: 3}[dtype]
        if out is None:
            out = torch.empty((self.n, self.cfg.rows, self.cfg.columns, channels), dtype=dtype, device=self.device)
        N.check(self.L.ecg_observe_onehot(C.byref(self.nat), _ptr(self.boards), _ptr(out), channels, kind, self.n,
                                          _stream(self.device)), "ecg_observe_onehot")
        return out

    # ------------------------------------------------------------------ statistics
    def episode_stats(self) -> torch.Tensor:
        """int64[6] on device: sum(score), n, min, max, wins, sum(score^2) (main.py:240-267 sample())"""
        out = torch.tensor([0, 0, 2 ** 63 - 1, -2 ** 63, 0, 0], dtype=torch.int64, device=self.device)
        N.check(self.L.ecg_episode_stats(_ptr(self.score), _ptr(self.flags), _ptr(out), self.n, _stream(self.device)),
                "ecg_episode_stats")
        return out


def fresh_key() -> int:
    return secrets.randbits(63)
