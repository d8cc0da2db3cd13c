#!/usr/bin/env python
"""Throughput of the fused rollout kernel (MCTS.rollout / random_task for every board in ONE launch).
   python scripts/rollout_bench.py [boards] [moves] [rows] [types]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

E = importlib.import_module("element-crush-gym_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
moves = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 9
types = int(sys.argv[4]) if len(sys.argv) > 4 else 6
cfg = E.BoardConfig(seed=5, rows=rows, columns=rows, types=types)
base = E.BatchedBoards(cfg, n, moves, key=99)
for rep in range(3):
    b = base.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tot = b.rollout()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = int(b.rollout_steps.sum().item())
    print(f"rollout {rows}x{rows}x{types}: {n} boards x {moves} moves: {ms:.2f} ms  {steps / ms * 1e3:.3e} env-steps/s  "
          f"mean episode reward {tot.float().mean().item():.1f}")
