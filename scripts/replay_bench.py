#!/usr/bin/env python
"""The reference's dynamics at scale: lockstep steps in ECG_REFILL_REPLAY mode (every step restarts the MT19937
stream of cfg.seed, boardv2.py:46) on Philox-drawn distinct boards.
    python scripts/replay_bench.py [boards] [steps] [streams: 1 = one shared stream | m = m distinct streams]"""
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

E = importlib.import_module("element-crush-gym_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
streams = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
seed = 20261019
cfg = E.BoardConfig(seed=seed)
src = E.BatchedBoards(cfg, n, 1 << 30, key=99)
kw = dict(seeds=[seed]) if streams == 1 else dict(seeds=[seed + i for i in range(streams)],
                                                 stream_index=torch.arange(n, device=src.device).remainder(streams))
b = E.BatchedBoards(cfg, n, 1 << 30, refill="replay", stream_len=2048, **kw)
b.boards.copy_(src.boards)
b._mask_valid = False
b.packed_mask()
for _ in range(3):
    b.apply_action(None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    b.apply_action(None)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"mode": "replay", "streams": streams, "boards": n, "ms_per_step": ms, "env_steps_per_s": n / ms * 1e3,
                  "mean_cascades": float(b.cascades.float().mean().item()),
                  "mean_reward": float(b.step_reward.float().mean().item()),
                  "overflow": int(((b.status & E.ST_STREAM_OVERFLOW) != 0).sum().item()),
                  "lib": os.path.basename(E._native.LIB_PATH)}))
