"""ctypes binding of libecg.so (include/ecg.h).  There is no CPU fallback: if the CUDA library is
missing or fails to load, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ECG_LIB selects another build of the same library (e.g. the -DECG_PROFILE_PHASES one used for ncu)
LIB_PATH = os.environ.get("ECG_LIB") or os.path.join(_HERE, "lib", "libecg.so")

ST_TERMINAL, ST_STREAM_OVERFLOW, ST_SHUFFLE_CAP, ST_BAD_ACTION = 1, 2, 4, 8
ST_NO_LEGAL, ST_BAD_CELL, ST_CASCADE_CAP = 16, 32, 64
FLAG_DONE, FLAG_WON = 1, 2
REFILL_REPLAY, REFILL_PHILOX = 1, 2
TILE = 32


class EcgError(RuntimeError):
    pass


class Config(C.Structure):
    """struct ecg_config"""
    _fields_ = [(n, C.c_int32) for n in (
        "rows", "cols", "types", "bits", "type_mask", "special_type_mask", "h_line", "v_line", "bomb",
        "mega_token", "action_space", "board_words", "mask_words", "reserved")]


class Refill(C.Structure):
    """struct ecg_refill"""
    _fields_ = [("mode", C.c_int32), ("stream_len", C.c_int32), ("stream", C.c_void_p),
                ("stream_stride", C.c_int64), ("stream_pos", C.c_void_p), ("philox_key", C.c_uint64),
                ("board0", C.c_uint64), ("step_ctr", C.c_uint32), ("reserved", C.c_uint32),
                ("stream_index", C.c_void_p), ("tiles", C.c_void_p), ("tile_wpos", C.c_void_p)]


class StepIO(C.Structure):
    """struct ecg_step_io"""
    _fields_ = [("boards_in", C.c_void_p), ("boards_out", C.c_void_p), ("actions", C.c_void_p),
                ("mask_in", C.c_void_p), ("actions_out", C.c_void_p), ("moves_left", C.c_void_p),
                ("reward", C.c_void_p), ("score", C.c_void_p), ("cascades", C.c_void_p),
                ("mask_out", C.c_void_p), ("flags", C.c_void_p), ("status", C.c_void_p),
                ("env_goal", C.c_int32), ("reserved", C.c_int32), ("src_index", C.c_void_p),
                ("scratch", C.c_void_p)]


EXPORTS = {
    # name: (restype, argtypes)
    "ecg_version": (C.c_int, []),
    "ecg_sizeof": (C.c_int, [C.c_int]),
    "ecg_last_error": (C.c_char_p, []),
    "ecg_launch_count": (C.c_int64, []),
    "ecg_config_init": (C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.c_int]),
    "ecg_boards_bytes": (C.c_int64, [C.POINTER(Config), C.c_int64]),
    "ecg_masks_bytes": (C.c_int64, [C.POINTER(Config), C.c_int64]),
    "ecg_pack": (C.c_int, [C.POINTER(Config), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ecg_unpack": (C.c_int, [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "ecg_unpack_nibbles": (C.c_int, [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ecg_unpack_mask": (C.c_int, [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ecg_mt19937_stream": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "ecg_replay_tiles_words": (C.c_int64, [C.c_int32]),
    "ecg_replay_tiles": (C.c_int, [C.c_void_p, C.c_int32, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ecg_init_boards": (C.c_int, [C.POINTER(Config), C.POINTER(Refill), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ecg_legal_mask": (C.c_int, [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ecg_random_action": (C.c_int, [C.POINTER(Config), C.POINTER(Refill), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_void_p]),
    "ecg_step": (C.c_int, [C.POINTER(Config), C.POINTER(Refill), C.POINTER(StepIO), C.c_int64, C.c_void_p]),
    "ecg_step_mark_event": (C.c_int, [C.c_void_p]),
    "ecg_rollout": (C.c_int, [C.POINTER(Config), C.POINTER(Refill), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_int64, C.c_void_p]),
    "ecg_rollout_scratch": (C.c_int, [C.POINTER(Config), C.POINTER(Refill), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ecg_episode_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ecg_augment": (C.c_int, [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int64, C.c_void_p]),
    "ecg_observe_onehot": (C.c_int, [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    "ecg_host_expand_nibbles": (C.c_int, [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_int64]),
    "ecg_host_expander_create": (C.c_void_p, [C.c_int]),
    "ecg_host_expander_submit": (C.c_int, [C.c_void_p, C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                           C.c_int]),
    "ecg_host_expander_wait": (C.c_int, [C.c_void_p]),
    "ecg_host_expander_destroy": (None, [C.c_void_p]),
}

_lib = None


def lib():
    """Load libecg.so (once).  Raises EcgError when the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EcgError(f"{LIB_PATH} not found: build it with `python element-crush-gym_b200/build.py` "
                           "(nvcc, sm_100a). This engine has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)  # AttributeError = the library does not match include/ecg.h
            fn.restype, fn.argtypes = res, args
        for which, st in enumerate((Config, Refill, StepIO)):  # ECG_SIZEOF_*: a stale stub must not reach ecg_step
            if L.ecg_sizeof(which) != C.sizeof(st):
                raise EcgError(f"{LIB_PATH}: sizeof({st.__name__}) is {L.ecg_sizeof(which)} in the library, "
                               f"{C.sizeof(st)} in _native.py -- rebuild the library")
        _lib = L
    return _lib


def check(rc, what="ecg call"):
    if rc != 0:
        raise EcgError(f"{what} failed ({rc}): {lib().ecg_last_error().decode()}")


def make_config(rows, cols, types) -> Config:
    cfg = Config()
    check(lib().ecg_config_init(C.byref(cfg), rows, cols, types), "ecg_config_init")
    return cfg
