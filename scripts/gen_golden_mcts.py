#!/usr/bin/env python
"""Golden vectors for the host tree: run the UNMODIFIED reference
mctslib.standard.mcts.MCTS (read-only import from /root/reference) on reference BoardV2 states with
`rollout` replaced by a deterministic function of the state, and record what every `mcts()` call returns
(action, value, policies) plus the root's children (actions in insertion order, visits, reward sums).
That pins UCB1 with c = node.state.n_actions (abc/mcts.py:95), pop-largest expansion (standard/mcts.py:33),
the policies order, the c = 0 value descent and the tree re-use between calls (abc/mcts.py:123-124).
Authoring container only; the vectors are committed as tests/golden/mcts_tree.json.
Usage: python scripts/gen_golden_mcts.py [--out tests/golden/mcts_tree.json]"""
import argparse
import json
import os
import sys

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from match3tile.boardConfig import BoardConfig  # noqa: E402
from match3tile.boardv2 import BoardV2  # noqa: E402
from mctslib.standard.mcts import MCTS  # noqa: E402


def stub_value(state):
    """deterministic stand-in for MCTS.rollout: a function of the state's cells and cumulative reward only
    (tests/test_gpu_mcts.py evaluates the same formula on the engine's BoardV2)"""
    a = np.asarray(state.array, dtype=np.int64)
    w = np.arange(1, a.size + 1, dtype=np.int64).reshape(a.shape)
    return int(state.reward) + int((a * w).sum() % 1009)


class StubMCTS(MCTS):
    def rollout(self, state):
        return stub_value(state)


CASES = [  # (rows, types, seed, moves, simulations, calls)
    (9, 6, 1, 6, 40, 3), (9, 6, 2, 5, 150, 2), (9, 6, 12345, 4, 400, 2),
    (6, 4, 3, 6, 60, 3), (6, 4, 7, 3, 200, 3), (12, 7, 5, 4, 80, 2),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mcts_tree.json"))
    args = ap.parse_args()
    out = []
    for rows, types, seed, moves, sims, calls in CASES:
        cfg = BoardConfig(seed=seed, rows=rows, columns=rows, types=types)
        state = BoardV2(moves, cfg)
        np.random.seed(cfg.seed)
        mcts = StubMCTS(state, 3.0, sims, False)
        rec = {"rows": rows, "types": types, "seed": seed, "moves": moves, "simulations": sims, "calls": []}
        for _ in range(calls):
            root = mcts._root
            action, value, policies = mcts()
            rec["calls"].append({
                "action": int(action), "value": int(value), "policies": [float(p) for p in policies],
                "root_visits": int(root.visits), "root_reward": int(root.reward),
                "child_actions": [int(a) for a in root.children.keys()],
                "child_visits": [int(c.visits) for c in root.children.values()],
                "child_rewards": [int(c.reward) for c in root.children.values()],
            })
            state = state.apply_action(action)
            if state.is_terminal:
                break
        out.append(rec)
        print(rows, types, seed, "done", [c["action"] for c in rec["calls"]], flush=True)
    with open(args.out, "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
