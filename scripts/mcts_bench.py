#!/usr/bin/env python
"""BASELINE.json configs[4] / SURVEY.md 8d config 5: mctslib standard MCTS on 9x9x6 (20 moves) with GPU-batched
random rollouts, `leaves` rollouts per simulation, (visits, reward sum) all-reduced over NCCL when launched under
torchrun.  Prints one JSON line: leaves/s, env-steps/s, time per simulation.

    python scripts/mcts_bench.py [--leaves 1048576] [--sims 32] [--moves 20]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/mcts_bench.py
"""
import argparse
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

E = importlib.import_module("element-crush-gym_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--leaves", type=int, default=1 << 20)
ap.add_argument("--sims", type=int, default=32)
ap.add_argument("--moves", type=int, default=20)
ap.add_argument("--refill", default="philox", choices=["philox", "replay"])
a = ap.parse_args()
rank, world, local = E.dist.init_from_env()
torch.cuda.set_device(local)
cfg = E.BoardConfig(seed=7)
state = E.BoardV2(a.moves, cfg, device=f"cuda:{local}")
m = E.BatchedRolloutMCTS(state, 3, 4, False, leaves=a.leaves, key=1234, refill=a.refill)
m()  # warm-up: 4 simulations (also re-roots the tree, like the reference's move loop)
m._simulations = a.sims
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
steps0 = m.env_steps
t0 = time.perf_counter()
action, value, policies = m()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
steps = m.env_steps - steps0  # already summed over the ranks
tmax = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
if rank == 0:
    dt = float(tmax.item())
    print(json.dumps({"config": "mctslib standard MCTS, 9x9x6, GPU-batched rollouts", "n_gpus": world,
                      "leaves_per_simulation": a.leaves, "simulations": a.sims, "moves": a.moves,
                      "ms_per_simulation": dt / a.sims * 1e3, "leaves_per_s": a.leaves * a.sims / dt,
                      "rollout_env_steps_per_s": steps / dt, "refill": a.refill, "action": int(action), "value": int(value),
                      "root_children": len(policies),
                      "reduction": "NCCL all-reduce of the reward sum per simulation (visit counts are known on the host)" if world > 1 else "none"}))
if world > 1:
    torch.distributed.destroy_process_group()
