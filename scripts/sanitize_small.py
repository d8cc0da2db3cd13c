#!/usr/bin/env python
"""Small invocation of every step-path kernel for compute-sanitizer (memcheck / racecheck / initcheck):
two-kernel Philox step (hand-off list: atomicAdd + work list), one-kernel replay step, whole-episode rollout,
chunked in-place HostStepper (both observation formats), expand() through src_index, the pack / unpack / mask kernels.
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
Prints one line per section; results are checked against nothing here (the -m gpu tests do that)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

E = importlib.import_module("element-crush-gym_b200")
dev = torch.device("cuda", 0)
for shape in ((9, 9, 6), (16, 16, 8)):
    cfg = E.BoardConfig(seed=5, rows=shape[0], columns=shape[1], types=shape[2])
    n = 2048 + 13
    b = E.BatchedBoards(cfg, n, 6, device=dev, key=77)
    for _ in range(3):
        b.apply_action(None)                       # two-kernel step, random picks
    b.apply_action(b.random_action())              # explicit actions
    b.observe(torch.uint8), b.observe(torch.int64), b.legal_mask()
    c = b.clone()
    c.rollout()                                    # whole episodes in one kernel
    child, parent, action = b.expand()             # src_index
    child.apply_action(None)
    torch.cuda.synchronize()
    print(shape, "philox ok: hand-offs of the last step", int(b._scratch[0]), "children", child.n, flush=True)
    r = E.BatchedBoards(cfg, 515, 5, device=dev, refill="replay", seeds=list(range(1, 516)), stream_len=2048)
    r.apply_action(None)
    ch, _, _ = r.expand()
    ch.apply_action(None)
    ch.rollout()
    s = E.BatchedBoards(cfg, 300, 5, device=dev, refill="replay", seeds=[9])
    s.apply_action(None)
    s.rollout()
    torch.cuda.synchronize()
    print(shape, "replay ok", flush=True)
env = E.BatchedMatch3Env(1000 + 7, seed=3, num_moves=5, env_goal=50, device=dev)
for fmt in ("uint8", "nibbles"):
    hs = E.HostStepper(env, chunks=5, obs_format=fmt)
    for _ in range(2):
        hs.step(hs.random_action())
    print("HostStepper", fmt, "ok", flush=True)
st = E.BoardV2(4, E.BoardConfig(seed=3), device=dev)
st2 = st.apply_action(st.legal_actions[0])
_ = st2.greedy_action, st2.array
m = E.BatchedRolloutMCTS(st, 3, 3, False, leaves=512)
m()
m = E.BatchedRolloutMCTS(st, 3, 3, False, leaves=64, refill="replay")
m()
torch.cuda.synchronize()
print("BoardV2 / MCTS ok", flush=True)
