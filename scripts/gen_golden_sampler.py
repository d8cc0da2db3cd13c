#!/usr/bin/env python
"""Generate tests/golden/sampler_greedy.npz by running the UNMODIFIED reference (read-only import from
/root/reference): samplerTasks.greedy_test (:17-22) with explicit seeds.  samplerTasks itself cannot be imported here
(it imports mctslib.nn -> jax), so its five-line loop is restated below around the reference's own BoardV2.
Authoring container only; the vectors are committed.  Usage: python scripts/gen_golden_sampler.py [--out tests/golden]
"""
import argparse
import os
import sys

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from match3tile.boardConfig import BoardConfig  # noqa: E402
from match3tile.boardv2 import BoardV2  # noqa: E402

CASES = [((9, 9, 6), 20, list(range(1, 9))), ((6, 6, 4), 12, list(range(101, 107))), ((12, 12, 7), 6, [7, 8])]


def greedy_episode(shape, moves, seed):
    R, C, T = shape
    state = BoardV2(moves, BoardConfig(seed=seed, rows=R, columns=C, types=T))
    np.random.seed(state.cfg.seed)
    actions = []
    while not state.is_terminal:
        a = state.greedy_action
        actions.append(a)
        state = state.apply_action(a)
    return actions, int(state.reward), np.asarray(state.array)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    a = ap.parse_args()
    out = {}
    for k, (shape, moves, seeds) in enumerate(CASES):
        acts, rewards, finals = [], [], []
        for s in seeds:
            ac, r, f = greedy_episode(shape, moves, s)
            acts.append(ac)
            rewards.append(r)
            finals.append(f)
            print(shape, s, r, flush=True)
        out[f"shape{k}"] = np.asarray(shape, dtype=np.int32)
        out[f"moves{k}"] = np.int32(moves)
        out[f"seeds{k}"] = np.asarray(seeds, dtype=np.int64)
        out[f"actions{k}"] = np.asarray(acts, dtype=np.int16)
        out[f"rewards{k}"] = np.asarray(rewards, dtype=np.int64)
        out[f"final{k}"] = np.asarray(finals, dtype=np.int8)
    out["cases"] = np.int32(len(CASES))
    np.savez_compressed(os.path.join(a.out, "sampler_greedy.npz"), **out)


if __name__ == "__main__":
    main()
