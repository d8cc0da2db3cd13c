#!/usr/bin/env python
"""Time the UNMODIFIED reference's Python/NumPy path (read-only import from /root/reference) on this machine's cores:
BASELINE.md section 3 / SURVEY.md 8d config 1.  Authoring container only -- the reference is pure Python and cannot
travel to the GPU box, so bench.py's cpu_baseline there is the C oracle port ("kind": "port") and this script's output
(profiles/python_reference_authoring_container.json) is the reference's own number, on other cores.

One task = samplerTasks.random_task (samplerTasks.py:9-14; re-typed here because samplerTasks imports jax):
BoardV2(20, BoardConfig()), np.random.seed(cfg.seed), apply_action(np.random.choice(legal_actions)) to terminal.
(a) single process; (b) all cores through the reference's own util.multiprocessingAutoBatcher.async_pbar_auto_batcher.
Usage: python scripts/time_python_reference.py [episodes_single] [episodes_pool]"""
import json
import os
import sys
import time

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from match3tile.boardConfig import BoardConfig  # noqa: E402
from match3tile.boardv2 import BoardV2  # noqa: E402
from util.multiprocessingAutoBatcher import async_pbar_auto_batcher  # noqa: E402


def random_task():
    state = BoardV2(20, BoardConfig())
    np.random.seed(state.cfg.seed)
    while not state.is_terminal:
        state = state.apply_action(np.random.choice(state.legal_actions))
    return state.reward


def main():
    n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
    t = time.perf_counter()
    rewards = [random_task() for _ in range(n1)]
    dt1 = time.perf_counter() - t
    t = time.perf_counter()
    pooled = async_pbar_auto_batcher(random_task, n2)
    dt2 = time.perf_counter() - t
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out = {"what": "reference Python/NumPy path, samplerTasks.random_task (20-move 9x9x6 episodes), authoring container",
           "cpu": cpu, "cores": os.cpu_count(),
           "single_process": {"episodes": n1, "env_steps": 20 * n1, "seconds": dt1, "env_steps_per_s": 20 * n1 / dt1,
                              "mean_episode_reward": float(np.mean(rewards))},
           "all_cores_reference_pool": {"episodes": n2, "env_steps": 20 * n2, "seconds": dt2,
                                        "env_steps_per_s": 20 * n2 / dt2, "workers": min(n2, os.cpu_count()),
                                        "mean_episode_reward": float(np.mean(pooled)),
                                        "note": "includes the Pool's process start-up, as the reference's sample() does"}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
