#!/usr/bin/env python
"""Differential run at scale: the CUDA path against the CPU oracle on natural play, every board size 4..16, both refill
modes, lockstep steps (two-kernel step) AND whole-episode rollouts (two-kernel rollout in Philox mode).  Compares, per
step, boards, step rewards, cascade counts, status and legal masks (lockstep) and final boards / episode rewards
(rollouts).  Prints one JSON line per configuration and a total; exit code 1 on any mismatch.
    python scripts/gpu_fuzz_vs_oracle.py [boards_per_config] [moves]
The oracle is test infrastructure: this script is a checker, like tests/."""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle.oracle import Oracle  # noqa: E402

E = importlib.import_module("element-crush-gym_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
moves = int(sys.argv[2]) if len(sys.argv) > 2 else 12
KEY = 0xD1FFE7E57
SHAPES = [(4, 4), (5, 4), (6, 4), (7, 5), (8, 5), (9, 6), (10, 6), (11, 7), (12, 7), (13, 9), (14, 7), (15, 8), (16, 8),
          (9, 3), (9, 8), (6, 11)]
np_ = lambda t: t.cpu().numpy()  # noqa: E731
bad = 0
total_steps = 0
t0 = time.time()
for rows, types in SHAPES:
    o = Oracle(rows, rows, types)
    cfg = E.BoardConfig(seed=11, rows=rows, columns=rows, types=types)
    nn = n if rows <= 12 else n // 2
    # ---- Philox: lockstep steps vs step_batch, then rollouts vs philox_episode_batch
    bb = E.BatchedBoards(cfg, nn, moves, key=KEY, board0=1000)
    boards = np_(bb.array)
    roll = bb.clone()
    mism = 0
    for t in range(moves):
        bb.apply_action(None)
    final, total, steps = o.philox_episode_batch(np_(roll.array), KEY, 1000, moves)
    mism += int((np_(bb.array) != final).any(axis=(1, 2)).sum()) + int((np_(bb.reward) != total).sum())
    tot = roll.rollout()
    mism += int((np_(roll.array) != final).any(axis=(1, 2)).sum()) + int((np_(tot) != total).sum())
    mism += int((np_(roll.rollout_steps) != steps).sum())
    mism += int((np_(bb.legal_mask()) != o.legal_mask_batch(final)).any(axis=1).sum())
    total_steps += 2 * int(steps.sum())
    # ---- replay: per-board MT19937 streams, every step against step_batch
    seeds = np.arange(1, nn // 8 + 1, dtype=np.int64)
    # few types cascade for hundreds of iterations: a stream that runs out makes the NEXT pick overflow on the engine's
    # side only (the oracle is handed the action), so give those shapes streams that never run out
    sl = 32768 if types <= 3 else 4096
    seeds = seeds[: len(seeds) // (4 if types <= 3 else 1)]
    rb = E.BatchedBoards(cfg, len(seeds), moves, refill="replay", seeds=seeds, stream_len=sl)
    raw = np.stack([Oracle.mt_raw(int(s), sl) for s in seeds])
    rboards = np_(rb.array)
    rr = rb.clone()
    rm = 0
    for t in range(moves):
        rb.apply_action(None)
        a = np_(rb.last_actions)
        res = o.step_batch(rboards, a, mode="replay", raw=raw)
        rm += int((np_(rb.array) != res["boards"]).any(axis=(1, 2)).sum()) + int((np_(rb.step_reward) != res["reward"]).sum())
        # a board without a legal move is a no-op on both sides; the engine says NO_LEGAL, the oracle (handed action
        # -1) BAD_ACTION
        want_status = np.where(a < 0, E.ST_NO_LEGAL, res["status"])
        rm += int((np_(rb.cascades) != res["cascades"]).sum()) + int((np_(rb.status) != want_status).sum())
        rm += int((np_(rb.legal_mask()) != res["legal"]).any(axis=1).sum())
        rboards = res["boards"]
        total_steps += len(seeds)
    rt = rr.rollout()
    rm += int((np_(rr.array) != rboards).any(axis=(1, 2)).sum()) + int((np_(rt) != np_(rb.reward)).sum())
    total_steps += len(seeds) * moves
    bad += mism + rm
    print(json.dumps({"shape": f"{rows}x{rows}x{types}", "philox_boards": nn, "replay_boards": len(seeds), "moves": moves,
                      "philox_mismatches": mism, "replay_mismatches": rm}), flush=True)
print(json.dumps({"total_env_steps_compared": total_steps, "mismatches": bad, "seconds": round(time.time() - t0, 1)}))
sys.exit(1 if bad else 0)
