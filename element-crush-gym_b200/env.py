"""Match3Env contract (match3tile/env.py:8-66) over the batched engine.

The reference's Match3Env is stale (it calls BoardV2 with a `seed=` keyword that no longer exists and
unpacks apply_action's result), so the *contract* is taken from env.py's text and
the *semantics* from BoardV2:
    step(action)  -> (obs, move_score, done, won, {})        env.py:48-56
    reset(seed)   -> (obs, {})                               env.py:58-65
    score += move_score; moves_taken += 1; won = score >= env_goal; done = won or moves_taken == num_moves
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _native as N
from .boards import BatchedBoards
from .config import BoardConfig


class BatchedMatch3Env:
    """num_envs independent Match3Env instances stepped in lockstep on one GPU.

    Same keyword arguments as the reference constructor (env.py:11-20) plus num_envs / device / refill.
    refill="philox" (default): env i uses Philox substreams (seed, board0 + i).
    refill="replay": env i is the reference env with seed `seed + i` (numpy MT19937 replay, bit-exact).
    """
    metadata = {"render_modes": ["human"], "render_fps": 60, "animation_speed": 1}

    def __init__(self, num_envs: int, width: int = 9, height: int = 9, num_types: int = 6, num_moves: int = 20,
                 env_goal: int = 500, seed: int = None, render_mode: str = None, *, device=None, refill="philox",
                 board0: int = 0, stream_len: int = 4096, obs_dtype=torch.uint8):
        assert width >= 3 and height >= 3, "Board size too small: min size: 3x3"  # env.py:27
        assert render_mode is None, "rendering (pygame) is outside the engine's scope"
        self.num_envs = int(num_envs)
        self.seed = seed if seed is not None else random.randint(1, 2 ** 31 - 1)
        self.width, self.height, self.num_types = width, height, num_types
        self.env_goal, self.num_moves = env_goal, num_moves
        self.action_space = height * (width - 1) + width * (height - 1)  # env.py:36
        self.render_mode = None
        self.device, self.refill, self.board0, self.stream_len = device, refill, board0, stream_len
        self.obs_dtype = obs_dtype
        self.cfg = BoardConfig(seed=self.seed, rows=height, columns=width, types=num_types)
        self.board: BatchedBoards = None
        self._make_board()

    def _make_board(self):
        kw = dict(device=self.device, refill=self.refill, board0=self.board0, env_goal=self.env_goal)
        if self.refill == "replay":
            kw.update(seeds=[(self.seed + i) & 0xFFFFFFFF for i in range(self.num_envs)], stream_len=self.stream_len)
        else:
            kw.update(key=self.seed)
        self.board = BatchedBoards(self.cfg, self.num_envs, self.num_moves, **kw)

    # -- bookkeeping views (env.py:34)
    @property
    def score(self):
        return self.board.score

    @property
    def moves_taken(self):
        return self.num_moves - self.board.moves_left

    def init(self):
        return self.board.observe(self.obs_dtype)

    def reset(self, seed=None):
        """env.py:58-65.  `(1 + self.seed) % 2**32 - 1` parses as ((1 + seed) % 2**32) - 1 == seed, i.e. a
        reset without an explicit seed replays the same boards; kept."""
        if seed is not None:
            self.seed = seed
            self.cfg = BoardConfig(seed=self.seed, rows=self.height, columns=self.width, types=self.num_types)
        self._make_board()
        return self.board.observe(self.obs_dtype), {}

    def step(self, actions=None):
        """actions: [num_envs] ints (tensor / ndarray / list) or None for board.random_action().
        Returns device tensors (obs [N,H,W], reward int32 [N], done bool [N], won bool [N], info)."""
        b = self.board
        b.apply_action(actions)
        obs = b.observe(self.obs_dtype)
        done = (b.flags & N.FLAG_DONE) != 0
        won = (b.flags & N.FLAG_WON) != 0
        return obs, b.step_reward, done, won, {}

    def render(self):
        return None


class HostStepper:
    """The public call with HOST buffers: actions come from pinned host memory, observation, reward, done
    and won go back to pinned host memory every step.  The batch is cut into chunks whose H2D copy, step
    kernel and D2H copies run on separate CUDA streams so PCIe transfers overlap the kernels (2^24 boards: 5.2e8 /
    5.6e8 / 5.8e8 env-steps/s with 2 / 8 / 32 chunks, 8.4e8 / 9.3e8 / 1.07e9 with the 4-bit observation, r06).
    bench.py's e2e number is measured through this class.

    obs_format="uint8" (default, the env contract): observation uint8 [N, H, W] of cell VALUES, reward int32,
        actions int32 -- 4 B in and H*W + 6 B out per board and step.
    obs_format="nibbles": the same information in about half the PCIe bytes (PCIe is the whole e2e cost: the step
        kernel is 20x faster than the copies): observation uint8 [N, ceil(H*W/2)] of 4-bit cell CODES
        (ecg_unpack_nibbles; `decode_obs` turns it into the uint8 form on the host when wanted), reward int16
        (saturating), actions int16.
    host_expand=True (with obs_format="uint8"): the same outputs as the default, byte for byte, but the observation
        crosses PCIe as 4-bit codes and a pool of host threads in libecg.so (ecg_host_expander_*) widens every chunk
        to uint8 cell values as soon as its copy has landed, while later chunks are still in flight:
        H*W/2 + 6 B out per board and step instead of H*W + 6."""

    def __init__(self, env: BatchedMatch3Env, chunks: int = 16, obs_format: str = "uint8", host_expand: bool = False,
                 expand_threads: int = None, expand_pieces: int = 8):
        if obs_format not in ("uint8", "nibbles"):
            raise ValueError("obs_format must be 'uint8' or 'nibbles'")
        if host_expand and obs_format != "uint8":
            raise ValueError("host_expand widens the 4-bit observation to the uint8 form: obs_format must be 'uint8'")
        self.env = env
        self.obs_format = obs_format
        self.host_expand = bool(host_expand)
        self._expander = None
        n = env.num_envs
        self.n = n
        dev = env.board.device
        self.dev = dev
        chunks = max(1, min(chunks, n // N.TILE or 1))
        per = -(-n // chunks)
        per = -(-per // N.TILE) * N.TILE  # whole tiles, so chunk views of the packed buffers stay aligned
        self.bounds = [(lo, min(lo + per, n)) for lo in range(0, n, per)]
        self.streams = [torch.cuda.Stream(device=dev) for _ in self.bounds]
        R, Cc = env.height, env.width
        pin = dict(pin_memory=True)
        nib = obs_format == "nibbles"
        small = torch.int16 if nib else torch.int32
        obs_shape = (n, (R * Cc + 1) // 2) if nib else (n, R, Cc)
        self.h_actions = torch.zeros(n, dtype=small, **pin)
        self.h_obs = torch.zeros(obs_shape, dtype=torch.uint8, **pin)
        self.h_reward = torch.zeros(n, dtype=small, **pin)
        self.h_done = torch.zeros(n, dtype=torch.bool, **pin)
        self.h_won = torch.zeros(n, dtype=torch.bool, **pin)
        self.d_actions = torch.zeros(n, dtype=torch.int32, device=dev)
        self.d_actions_in = torch.zeros(n, dtype=small, device=dev) if nib else self.d_actions
        self.d_reward = torch.zeros(n, dtype=small, device=dev) if nib else None
        self.d_tmp32 = torch.zeros(n, dtype=torch.int32, device=dev) if nib else None
        self.d_obs = None if self.host_expand else torch.zeros(obs_shape, dtype=torch.uint8, device=dev)
        # done / won as 0/1 bytes, split from the flags byte on the device: a host-side pass over N flags per step would
        # cost more than the copy (and torchrun pins the host side to one thread)
        self.d_done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.d_won = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.scratch = (torch.empty(n + len(self.bounds), dtype=torch.int32, device=dev)
                        if env.board.two_kernel_step else None)
        a_bytes = 2 if nib else 4
        self.h2d_bytes = n * a_bytes
        self.d2h_bytes = n * (self.h_obs[0].numel() + a_bytes + 2)
        self.action_d2h_bytes = n * a_bytes  # random_action(): the pick the host sends back in
        if self.host_expand:
            import ctypes as C
            import os
            nb = (R * Cc + 1) // 2
            self.d_nib = torch.zeros((n, nb), dtype=torch.uint8, device=dev)
            self.h_nib = torch.zeros((n, nb), dtype=torch.uint8, **pin)
            # a chunk's 4-bit codes cross PCIe in `expand_pieces` copies, each followed by an event: finer pieces are
            # widened sooner (2^24 boards, 32 chunks, 14 threads: 1 / 4 / 8 / 16 pieces -> 24.4 / 20.6 / 19.7 / 19.7 ms)
            self.expand_sub = max(1, int(expand_pieces))
            self.events = [torch.cuda.Event() for _ in range(len(self.bounds) * self.expand_sub)]
            if expand_threads is None:
                local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
                # two cores stay free for the thread that queues the chunks and the driver's own threads; past ~12
                # threads the widening is bound by host memory bandwidth anyway (r10: 8 / 12 / 14 / 16 threads ->
                # 24.8 / 22.2 / 21.2 / 22.3 ms per 2^24-board step on a 16-core host)
                expand_threads = max(1, ((os.cpu_count() or 1) - 2) // max(local, 1))
            self.expand_threads = int(expand_threads)
            self._expander = C.c_void_p(env.board.L.ecg_host_expander_create(self.expand_threads))
            if not self._expander:
                raise N.EcgError("ecg_host_expander_create failed: " + env.board.L.ecg_last_error().decode())
            self.d2h_bytes = n * (nb + a_bytes + 2)

    def close(self):
        if self._expander:
            self.env.board.L.ecg_host_expander_destroy(self._expander)
            self._expander = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chunk_call(self, lo, hi, fn_step):
        b = self.env.board
        nat = b.nat
        bw, mw = nat.board_words, nat.mask_words
        io = N.StepIO()
        boards = b.boards[lo * bw:]
        mask = b.mask[lo * mw:]
        io.boards_in = io.boards_out = boards.data_ptr()
        io.actions = self.d_actions[lo:].data_ptr()
        io.actions_out = b.last_actions[lo:].data_ptr()
        io.moves_left = b.moves_left[lo:].data_ptr()
        io.reward = b.step_reward[lo:].data_ptr()
        io.score = b.score[lo:].data_ptr()
        io.cascades = b.cascades[lo:].data_ptr()
        io.mask_out = mask.data_ptr()
        io.flags = b.flags[lo:].data_ptr()
        io.status = b.status[lo:].data_ptr()
        io.env_goal = b.env_goal
        if self.scratch is not None:  # k-th chunk [lo, hi): hi - lo + 1 entries of the work list from lo + k
            io.scratch = self.scratch[lo + self.bounds.index((lo, hi)):].data_ptr()
        rf = b._refill()
        rf.board0 = b.board0 + lo
        if b.refill_mode == "replay":
            if b.stream_index is not None:
                rf.stream_index = b.stream_index[lo:].data_ptr()
            elif b.stream_stride:  # chunk [lo, hi) of per-board streams (and of their tile tables)
                rf.stream = b.stream.data_ptr() + 4 * lo * b.stream_stride
                if rf.tiles:
                    rf.tiles = b.tiles.data_ptr() + 4 * lo * (b.tiles.numel() // b.n)
                    rf.tile_wpos = b.tile_wpos.data_ptr() + 2 * lo * (b.stream_len + 1)
            rf.stream_pos = b.stream_pos[lo:].data_ptr()
        return io, rf

    def step(self, actions_host: torch.Tensor = None):
        """actions_host: pinned int32 (int16 with obs_format="nibbles") [N]; defaults to self.h_actions.  Returns the
        pinned host tensors (obs, reward, done bool [N], won bool [N], {}) after a full sync."""
        import ctypes as C
        env, b = self.env, self.env.board
        L = b.L
        nib = self.obs_format == "nibbles"
        src = self.h_actions if actions_host is None else actions_host
        cur = torch.cuda.current_stream(self.dev)
        for k, ((lo, hi), st) in enumerate(zip(self.bounds, self.streams)):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                self.d_actions_in[lo:hi].copy_(src[lo:hi], non_blocking=True)
                if nib:
                    self.d_actions[lo:hi].copy_(self.d_actions_in[lo:hi])  # int16 -> int32 on the device
                io, rf = self._chunk_call(lo, hi, None)
                sp = C.c_void_p(st.cuda_stream)
                N.check(L.ecg_step(C.byref(b.nat), C.byref(rf), C.byref(io), hi - lo, sp), "ecg_step")
                boards = C.c_void_p(b.boards[lo * b.nat.board_words:].data_ptr())
                if nib:
                    N.check(L.ecg_unpack_nibbles(C.byref(b.nat), boards, C.c_void_p(self.d_obs[lo:].data_ptr()),
                                                 hi - lo, sp), "ecg_unpack_nibbles")
                    torch.clamp(b.step_reward[lo:hi], max=32767, out=self.d_tmp32[lo:hi])  # saturate, then narrow
                    self.d_reward[lo:hi].copy_(self.d_tmp32[lo:hi])
                    self.h_reward[lo:hi].copy_(self.d_reward[lo:hi], non_blocking=True)
                elif self.host_expand:
                    N.check(L.ecg_unpack_nibbles(C.byref(b.nat), boards, C.c_void_p(self.d_nib[lo:].data_ptr()),
                                                 hi - lo, sp), "ecg_unpack_nibbles")
                    sub = self.expand_sub
                    for j in range(sub):  # pieces of a chunk: each is widened as soon as it has landed
                        slo = lo + ((hi - lo) * j // sub) // 16 * 16
                        shi = hi if j + 1 == sub else lo + ((hi - lo) * (j + 1) // sub) // 16 * 16
                        if shi <= slo:
                            continue
                        self.h_nib[slo:shi].copy_(self.d_nib[slo:shi], non_blocking=True)
                        ev = self.events[k * sub + j]
                        ev.record(st)
                        N.check(L.ecg_host_expander_submit(self._expander, C.byref(b.nat),
                                                           C.c_void_p(self.h_nib[slo:].data_ptr()),
                                                           C.c_void_p(self.h_obs[slo:].data_ptr()), shi - slo,
                                                           C.c_void_p(ev.cuda_event), 1), "ecg_host_expander_submit")
                    self.h_reward[lo:hi].copy_(b.step_reward[lo:hi], non_blocking=True)
                else:
                    N.check(L.ecg_unpack(C.byref(b.nat), boards, C.c_void_p(self.d_obs[lo:].data_ptr()), 1, hi - lo,
                                         sp), "ecg_unpack")
                    self.h_reward[lo:hi].copy_(b.step_reward[lo:hi], non_blocking=True)
                if not self.host_expand:
                    self.h_obs[lo:hi].copy_(self.d_obs[lo:hi], non_blocking=True)
                torch.bitwise_and(b.flags[lo:hi], N.FLAG_DONE, out=self.d_done[lo:hi])
                torch.bitwise_right_shift(b.flags[lo:hi], 1, out=self.d_won[lo:hi])  # FLAG_WON == 2, the top flag
                self.h_done[lo:hi].copy_(self.d_done[lo:hi].view(torch.bool), non_blocking=True)
                self.h_won[lo:hi].copy_(self.d_won[lo:hi].view(torch.bool), non_blocking=True)
        for st in self.streams:
            cur.wait_stream(st)
        cur.synchronize()
        if self.host_expand:
            N.check(L.ecg_host_expander_wait(self._expander), "ecg_host_expander_wait")
        b._mask_valid = True
        b.step_ctr += 1
        b._moves_bound = max(b._moves_bound - 1, 0)
        return self.h_obs, self.h_reward, self.h_done, self.h_won, {}

    def random_action(self) -> torch.Tensor:
        """board.random_action() with the result copied to pinned host memory (int32 / int16 [N])"""
        a = self.env.board.random_action(out=self.d_actions)
        if self.obs_format == "nibbles":
            self.d_actions_in.copy_(a)
            a = self.d_actions_in
        self.h_actions.copy_(a, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self.h_actions

    def decode_obs(self, obs_nibbles=None) -> np.ndarray:
        """host side, optional: the nibble observation -> uint8 [N, H, W] of cell values (BoardV2.array as uint8)"""
        cfg = self.env.cfg
        x = (self.h_obs if obs_nibbles is None else obs_nibbles).numpy()
        lut = np.array(list(range(12)) + [cfg.h_line, cfg.v_line, cfg.bomb, cfg.mega_token], dtype=np.uint8)
        cells = np.empty((x.shape[0], x.shape[1] * 2), dtype=np.uint8)
        cells[:, 0::2] = lut[x & 15]
        cells[:, 1::2] = lut[x >> 4]
        R, Cc = self.env.height, self.env.width
        return cells[:, :R * Cc].reshape(-1, R, Cc)


class Match3Env:
    """Single-environment drop-in with the reference's signature and return types (env.py:8-66):
    numpy int64 observation, Python int reward, Python bools.  Runs the reference's own RNG semantics
    (numpy MT19937 replay of `seed`), so episodes are bit-identical to BoardV2 driven the same way."""
    metadata = BatchedMatch3Env.metadata

    def __init__(self, width: int = 9, height: int = 9, num_types: int = 6, num_moves: int = 20, env_goal: int = 500,
                 seed: int = None, render_mode: str = None, *, device=None, refill="replay"):
        self._env = BatchedMatch3Env(1, width, height, num_types, num_moves, env_goal, seed, render_mode,
                                     device=device, refill=refill, obs_dtype=torch.int64)
        self.width, self.height, self.num_types = width, height, num_types
        self.env_goal, self.num_moves = env_goal, num_moves
        self.action_space = self._env.action_space
        self.render_mode = None

    class _Board:
        def __init__(self, outer):
            self._o = outer

        @property
        def array(self):
            return self._o._env.board.array[0].cpu().numpy()

        @property
        def legal_actions(self):
            return self._o._env.board.legal_actions[0]

        def random_action(self) -> int:
            return int(self._o._env.board.random_action()[0].item())

    @property
    def board(self):
        return Match3Env._Board(self)

    @property
    def seed(self):
        return self._env.seed

    @property
    def score(self) -> int:
        return int(self._env.board.score[0].item())

    @property
    def moves_taken(self) -> int:
        return int(self._env.moves_taken[0].item())

    def init(self) -> np.ndarray:
        return self.board.array

    def step(self, action: int):
        self.actions = self.board.legal_actions  # env.py:49
        obs, reward, done, won, info = self._env.step([int(action)])
        return obs[0].cpu().numpy(), int(reward[0].item()), bool(done[0].item()), bool(won[0].item()), info

    def reset(self, seed=None):
        obs, info = self._env.reset(seed)
        return obs[0].cpu().numpy(), info

    def render(self):
        return None
