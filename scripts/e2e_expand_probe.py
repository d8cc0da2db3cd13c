#!/usr/bin/env python
"""HostStepper(host_expand=True): chunks x host threads sweep.  python scripts/e2e_expand_probe.py [boards]"""
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
for chunks in (16, 32, 64):
    for threads in (12, 14):
        hs = E.HostStepper(env, chunks=chunks, host_expand=True, expand_threads=threads)
        a = hs.random_action()
        print(f"chunks={chunks:3d} threads={threads:2d}: random_action {timeit(hs.random_action):6.2f} ms   step {timeit(lambda: hs.step(a)):6.2f} ms", flush=True)
        hs.close()
        del hs
hs = E.HostStepper(env, chunks=32, obs_format="nibbles")
a = hs.random_action()
print(f"nibbles, 32 chunks: random_action {timeit(hs.random_action):6.2f} ms   step {timeit(lambda: hs.step(a)):6.2f} ms")
