#!/usr/bin/env python
"""Host-side widening of the 4-bit observation (ecg_host_expander_*) alone, no GPU: 2^24 9x9 boards, pieces queued
without events.  usage: python scripts/microbench/host_expand.py [threads ...]   (ECG_EXPAND_NT=0: cached stores)"""
import ctypes as C
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

E = importlib.import_module("element-crush-gym_b200")
N = E._native
L = N.lib()
cfg = N.make_config(9, 9, 6)
n = 1 << 24
nib = np.random.default_rng(1).integers(0, 256, size=(n, 41), dtype=np.uint8)
out = np.zeros((n, 81), dtype=np.uint8)
for t in [int(a) for a in sys.argv[1:]] or [4, 8, 16]:
    x = L.ecg_host_expander_create(t)
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        for k in range(32):
            lo, hi = n * k // 32, n * (k + 1) // 32
            L.ecg_host_expander_submit(x, C.byref(cfg), nib[lo:].ctypes.data, out[lo:].ctypes.data, hi - lo, None, 4)
        L.ecg_host_expander_wait(x)
        best = min(best, time.perf_counter() - t0)
    L.ecg_host_expander_destroy(x)
    print(f"threads {t}: {best * 1e3:.1f} ms per 2^24 boards, {n * 122 / best / 1e9:.1f} GB/s in+out, "
          f"NT={os.environ.get('ECG_EXPAND_NT', '1')}", flush=True)
