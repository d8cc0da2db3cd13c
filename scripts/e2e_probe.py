#!/usr/bin/env python
"""Where the end-to-end (host buffer) step spends its time: python scripts/e2e_probe.py [boards]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
E = importlib.import_module("element-crush-gym_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
env = E.BatchedMatch3Env(n, 9, 9, 6, num_moves=1 << 30, seed=12345, refill="philox")
env.board.packed_mask()
def timeit(f, reps=4):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3
for chunks in (4, 8, 16, 32):
    hs = E.HostStepper(env, chunks=chunks)
    a = hs.random_action()
    print(f"chunks={chunks:3d}: random_action {timeit(hs.random_action):6.2f} ms   step {timeit(lambda: hs.step(a)):6.2f} ms")
    del hs
b = env.board
print(f"observe(u8) kernel: {timeit(lambda: b.observe(torch.uint8)):6.2f} ms;  step kernel: {timeit(lambda: b.apply_action(None)):6.2f} ms")
