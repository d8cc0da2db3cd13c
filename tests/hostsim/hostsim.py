"""TEST-ONLY: ctypes front-end of the host build of the kernel core (tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "..", "..", "element-crush-gym_b200", "csrc")
SIZES = tuple(range(4, 17))


def _lib_path(size):
    return os.path.join(_HERE, f"libecg_hostsim_{size}.so")


def _stale(size):
    lib = _lib_path(size)
    srcs = [os.path.join(_HERE, "hostsim.cpp"), os.path.join(_CSRC, "ecg_core.cuh"), os.path.join(_CSRC, "ecg_bits.cuh")]
    return not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs)


def _compile(size):
    tmp = _lib_path(size) + f".{os.getpid()}.tmp"
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wno-unknown-pragmas", f"-DHS_SIZE={size}",
                    "-o", tmp, os.path.join(_HERE, "hostsim.cpp")], check=True)
    os.replace(tmp, _lib_path(size))


def build(sizes=SIZES, force=False):
    """one library per board size, the stale ones compiled in parallel"""
    todo = [n for n in sizes if force or _stale(n)]
    if todo:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            list(ex.map(_compile, todo))
    return [_lib_path(n) for n in sizes]


_libs = {}


def lib(size):
    if size not in _libs:
        build((size,))
        L = C.CDLL(_lib_path(size))
        L.hs_handoffs.restype = C.c_int64
        L.hs_class_errors.restype = C.c_int64
        _libs[size] = L
    return _libs[size]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class HostSim:
    def __init__(self, rows, cols, types):
        self.rows, self.cols, self.types = rows, cols, types
        self.A = rows * (cols - 1) * 2
        self.L = lib(rows)

    def step(self, boards, actions, *, mode, raw=None, key=0, board0=0, step_ctr=0, moves_left=None, two_pass=False,
             pooled=False):
        """two_pass: the two-kernel step on the host -- common-case build first, exact build on a hand-off;
        pooled (Philox): the common-case pass cut the way the pooled step kernel cuts it (BEGIN, step_iter<DEFER_LEGAL>
        per pass with cascade_class checked against the pending match, legal swaps at FINISH)"""
        boards = np.ascontiguousarray(boards, dtype=np.int64)
        n = boards.shape[0]
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        m = {"replay": 1, "philox": 2}[mode] + (2 if two_pass else 0)
        if pooled:
            assert mode == "philox"
            m = 6
        stride = rawlen = 0
        if raw is not None:
            raw = np.ascontiguousarray(raw, dtype=np.uint32)
            stride, rawlen = (raw.shape[1], raw.shape[1]) if raw.ndim == 2 else (0, raw.size)
        ml = None if moves_left is None else np.ascontiguousarray(moves_left, dtype=np.int32)
        out = np.zeros_like(boards)
        reward = np.zeros(n, dtype=np.int64)
        casc = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint8)
        legal = np.zeros((n, self.A), dtype=np.uint8)
        words = np.zeros(n, dtype=np.uint32)
        if mode == "replay":
            self.L.hs_set_words(_p(words))
        rc = self.L.hs_step(self.rows, self.cols, self.types, m, _p(raw), C.c_int64(stride), C.c_int64(rawlen),
                            C.c_uint64(key), C.c_uint64(board0), C.c_uint32(step_ctr), _p(boards), _p(actions), _p(ml),
                            _p(out), _p(reward), _p(casc), _p(status), _p(legal), C.c_int64(n))
        assert rc == 0
        return {"boards": out, "reward": reward, "cascades": casc, "status": status, "legal": legal.astype(bool),
                "handoffs": int(self.L.hs_handoffs()), "words": words, "class_errors": int(self.L.hs_class_errors())}

    def legal(self, boards):
        boards = np.ascontiguousarray(boards, dtype=np.int64)
        n = boards.shape[0]
        legal = np.zeros((n, self.A), dtype=np.uint8)
        assert self.L.hs_legal(self.rows, self.cols, self.types, _p(boards), _p(legal), C.c_int64(n)) == 0
        return legal.astype(bool)

    def matches(self, boards):
        boards = np.ascontiguousarray(boards, dtype=np.int64)
        n = boards.shape[0]
        mask = np.zeros(boards.shape, dtype=np.uint8)
        spawn = np.zeros(boards.shape, dtype=np.int32)
        assert self.L.hs_matches(self.rows, self.cols, self.types, _p(boards), _p(mask), _p(spawn), C.c_int64(n)) == 0
        return mask.astype(bool), spawn

    def init(self, *, mode, n, raw=None, key=0, board0=0):
        m = {"replay": 1, "philox": 2}[mode]
        stride = rawlen = 0
        if raw is not None:
            raw = np.ascontiguousarray(raw, dtype=np.uint32)
            stride, rawlen = (raw.shape[1], raw.shape[1]) if raw.ndim == 2 else (0, raw.size)
        out = np.zeros((n, self.rows, self.cols), dtype=np.int64)
        status = np.zeros(n, dtype=np.uint8)
        assert self.L.hs_init(self.rows, self.cols, self.types, m, _p(raw), C.c_int64(stride), C.c_int64(rawlen),
                              C.c_uint64(key), C.c_uint64(board0), _p(out), _p(status), C.c_int64(n)) == 0
        return out, status
