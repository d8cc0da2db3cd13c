"""ctypes front-end of the CPU ORACLE (oracle/ecg_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(element-crush-gym_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libecg_oracle.so")

ST_TERMINAL, ST_STREAM_OVERFLOW, ST_SHUFFLE_CAP, ST_BAD_ACTION, ST_NO_LEGAL, ST_BAD_CELL, ST_CASCADE_CAP = 1, 2, 4, 8, 16, 32, 64


class Cfg(C.Structure):
    _fields_ = [
        ("rows", C.c_int), ("cols", C.c_int), ("types", C.c_int), ("bits", C.c_int),
        ("type_mask", C.c_int64), ("special_type_mask", C.c_int64), ("h_line", C.c_int64),
        ("v_line", C.c_int64), ("bomb", C.c_int64), ("mega_token", C.c_int64),
        ("action_space", C.c_int),
    ]


class Rng(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("seed", C.c_uint32), ("mt", C.c_uint32 * 624), ("mti", C.c_int),
        ("raw", C.c_void_p), ("raw_len", C.c_int64), ("key", C.c_uint32 * 2), ("board", C.c_uint32 * 2),
        ("step", C.c_uint32), ("blk", C.c_uint32 * 4), ("blk_idx", C.c_int64), ("pos", C.c_int64),
        ("overflow", C.c_int), ("dig_x", C.c_uint32), ("dig_left", C.c_uint32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc); returns the .so path."""
    src = os.path.join(_HERE, "ecg_oracle.c")
    hdr = os.path.join(_HERE, "ecg_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ecgo_random_episode.restype = C.c_int64
        L.ecgo_philox_episode.restype = C.c_int64
        L.ecgo_rng_u32.restype = C.c_uint32
        L.ecgo_rng_below.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """Reference semantics for one board shape (rows, cols, types)."""

    def __init__(self, rows=9, cols=9, types=6):
        self.L = lib()
        self.cfg = Cfg()
        self.L.ecgo_cfg_init(C.byref(self.cfg), rows, cols, types)
        self.rows, self.cols, self.types = rows, cols, types
        self.A = self.cfg.action_space

    # -- config / codec (boardConfig.py)
    def decode(self, action):
        out = (C.c_int * 4)()
        self.L.ecgo_decode(C.byref(self.cfg), int(action), out)
        return (out[0], out[1]), (out[2], out[3])

    def encode(self, t1, t2):
        return self.L.ecgo_encode(C.byref(self.cfg), t1[0], t1[1], t2[0], t2[1])

    # -- rng
    def rng_mt(self, seed):
        r = Rng()
        self.L.ecgo_rng_init_mt(C.byref(r), C.c_uint32(seed))
        return r

    def rng_replay(self, raw):
        raw = np.ascontiguousarray(raw, dtype=np.uint32)
        r = Rng()
        self.L.ecgo_rng_init_replay(C.byref(r), _p(raw), C.c_int64(raw.size))
        r._keep = raw
        return r

    def rng_philox(self, key, board, step):
        r = Rng()
        self.L.ecgo_rng_init_philox(C.byref(r), C.c_uint64(key), C.c_uint64(board), C.c_uint32(step))
        return r

    @staticmethod
    def mt_raw(seed, n):
        out = np.empty(n, dtype=np.uint32)
        lib().ecgo_mt_raw(C.c_uint32(seed), _p(out), C.c_int64(n))
        return out

    @staticmethod
    def philox(ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        lib().ecgo_philox4x32_10(c, k, o)
        return [int(x) for x in o]

    # -- board functions (boardFunctions.py)
    def legal_actions(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.int64)
        out = (C.c_int * max(self.A, 1))()
        n = self.L.ecgo_legal_actions(C.byref(self.cfg), _p(arr), out)
        return [out[i] for i in range(n)]

    def matches_and_spawn(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.int64)
        mask = np.zeros((self.rows, self.cols), dtype=np.uint8)
        spawn = np.zeros((self.rows, self.cols), dtype=np.int32)
        n = self.L.ecgo_matches_and_spawn(C.byref(self.cfg), _p(arr), _p(mask), _p(spawn))
        return mask.astype(bool), spawn, n

    def shuffle(self, rng, arr):
        arr = np.ascontiguousarray(arr, dtype=np.int64).copy()
        self.L.ecgo_shuffle(C.byref(self.cfg), C.byref(rng), _p(arr))
        return arr

    # -- state (boardv2.py)
    def init_board(self, rng):
        out = np.zeros((self.rows, self.cols), dtype=np.int64)
        self.L.ecgo_init_board(C.byref(self.cfg), C.byref(rng), _p(out))
        return out

    def apply_action(self, rng, arr, action):
        """-> (next_board, step_reward, cascades, draws, status)"""
        arr = np.ascontiguousarray(arr, dtype=np.int64)
        out = np.zeros_like(arr)
        reward = C.c_int64(0)
        casc = C.c_int(0)
        draws = C.c_int(0)
        st = self.L.ecgo_apply_action(C.byref(self.cfg), C.byref(rng), _p(arr), int(action), _p(out),
                                      C.byref(reward), C.byref(casc), C.byref(draws))
        return out, int(reward.value), int(casc.value), int(draws.value), int(st)

    # -- batches
    def legal_mask_batch(self, boards):
        boards = np.ascontiguousarray(boards, dtype=np.int64)
        n = boards.shape[0]
        mask = np.zeros((n, self.A), dtype=np.uint8)
        self.L.ecgo_legal_mask_batch(C.byref(self.cfg), _p(boards), _p(mask), C.c_int64(n))
        return mask.astype(bool)

    def step_batch(self, boards, actions, *, mode, raw=None, raw_stride=0, key=0, board0=0, step_ctr=0,
                   moves_left=None, want_legal=True):
        """mode: 'mt' (raw = per-board seeds), 'replay' (raw = u32 streams), 'philox'.
        -> dict(boards, reward, cascades, status, legal)"""
        boards = np.ascontiguousarray(boards, dtype=np.int64)
        n = boards.shape[0]
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        m = {"mt": 0, "replay": 1, "philox": 2}[mode]
        raw_len = 0
        if raw is not None:
            raw = np.ascontiguousarray(raw, dtype=np.uint32)
            if m == 0:
                raw_stride, raw_len = 1, 0
            elif raw.ndim == 2:
                raw_stride, raw_len = raw.shape[1], raw.shape[1]
            else:
                raw_stride, raw_len = 0, raw.size
        ml = None if moves_left is None else np.ascontiguousarray(moves_left, dtype=np.int32)
        out = np.zeros_like(boards)
        reward = np.zeros(n, dtype=np.int64)
        casc = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint8)
        legal = np.zeros((n, self.A), dtype=np.uint8) if want_legal else None
        self.L.ecgo_step_batch(C.byref(self.cfg), m, _p(raw), C.c_int64(raw_stride), C.c_int64(raw_len),
                               C.c_uint64(key), C.c_uint64(board0), C.c_uint32(step_ctr), _p(boards), _p(actions),
                               _p(ml), _p(out), _p(reward), _p(casc), _p(status), _p(legal), C.c_int64(n))
        return {"boards": out, "reward": reward, "cascades": casc, "status": status,
                "legal": None if legal is None else legal.astype(bool)}

    def random_episode(self, seed, n_moves=20):
        """samplerTasks.py:9-14 on an explicit seed -> (total reward, steps, final board)"""
        steps = C.c_int64(0)
        fb = np.zeros((self.rows, self.cols), dtype=np.int64)
        r = self.L.ecgo_random_episode(C.byref(self.cfg), C.c_uint32(seed), n_moves, C.byref(steps), _p(fb))
        return int(r), int(steps.value), fb

    def random_episode_batch(self, seeds, n_moves=20):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = seeds.size
        reward = np.zeros(n, dtype=np.int64)
        steps = np.zeros(n, dtype=np.int64)
        self.L.ecgo_random_episode_batch(C.byref(self.cfg), _p(seeds), n_moves, _p(reward), _p(steps), C.c_int64(n))
        return reward, steps

    def philox_episode_batch(self, boards, key, board0=0, n_moves=20, step0=0, inplace=False):
        boards = np.ascontiguousarray(boards, dtype=np.int64)
        if not inplace:
            boards = boards.copy()
        n = boards.shape[0]
        reward = np.zeros(n, dtype=np.int64)
        steps = np.zeros(n, dtype=np.int64)
        self.L.ecgo_philox_episode_batch(C.byref(self.cfg), C.c_uint64(key), C.c_uint64(board0), C.c_uint32(step0), n_moves,
                                         _p(boards), _p(reward), _p(steps), C.c_int64(n))
        return boards, reward, steps

    @staticmethod
    def max_threads():
        return lib().ecgo_max_threads()

    @staticmethod
    def set_threads(n):
        lib().ecgo_set_threads(int(n))
