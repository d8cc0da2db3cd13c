"""element-crush-gym_b200: B200-native batched stepping engine for Element-Crush-Gym's match-3 env.

Host-side mirror of the reference interface for the hot path only:
    BoardConfig                       match3tile/boardConfig.py
    BoardV2 (single-board State)      match3tile/boardv2.py, mctslib/abc/mcts.py:8-30
    BatchedBoards                     N x BoardV2 in lockstep on one GPU
    Match3Env / BatchedMatch3Env      match3tile/env.py
    dist                              sharding + NCCL reduction of statistics
    BatchedRolloutMCTS                mctslib/standard/mcts.py with GPU-batched rollouts
    dataset / Dataset                 dataset.py (self-play samples, mirror / type-switch augmentation)
    sampler                           samplerTasks.py + util/multiprocessingAutoBatcher.py (episodes sharded over ranks)
All board logic runs in libecg.so (hand-written CUDA for sm_100a behind include/ecg.h); there is no CPU path.
The directory name is not a Python identifier: import it with
    importlib.import_module("element-crush-gym_b200")      # or `import ecg_b200` (alias module at the repo root)
"""
from . import _native
from ._native import EcgError, FLAG_DONE, FLAG_WON, ST_BAD_ACTION, ST_BAD_CELL, ST_CASCADE_CAP, ST_NO_LEGAL, \
    ST_SHUFFLE_CAP, ST_STREAM_OVERFLOW, ST_TERMINAL
from .config import BoardConfig
from .boards import BatchedBoards, fresh_key
from .env import BatchedMatch3Env, HostStepper, Match3Env
from .state import BoardV2
from . import dist
from .mcts import BatchedRolloutMCTS
from . import dataset
from .dataset import Dataset
from . import sampler

__all__ = ["BoardConfig", "BoardV2", "BatchedBoards", "BatchedMatch3Env", "Match3Env", "HostStepper", "dist",
           "EcgError", "fresh_key", "BatchedRolloutMCTS", "dataset", "Dataset", "sampler"]
