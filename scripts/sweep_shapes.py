#!/usr/bin/env python
"""BASELINE.json configs[3]: board-shape sweep (6x6x4, 9x9x6, 12x12x7, 16x16x8) of the lockstep step kernel
(Philox refill, random legal action, legal mask each step), with the roofline of every shape: algorithmic bytes per
env-step x env-steps/s against the measured HBM peak.
    python scripts/sweep_shapes.py [boards] [steps] [rows:types,rows:types,...]"""
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

E = importlib.import_module("element-crush-gym_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
shapes = ((6, 4), (9, 6), (12, 7), (16, 8))
if len(sys.argv) > 3:
    shapes = tuple(tuple(int(x) for x in p.split(":")) for p in sys.argv[3].split(","))
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
out = []
for rows, types in shapes:
    cfg = E.BoardConfig(seed=5, rows=rows, columns=rows, types=types)
    b = E.BatchedBoards(cfg, n, 1 << 30, key=99)
    b.packed_mask()
    for _ in range(4):
        b.apply_action(None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        b.apply_action(None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    A = cfg.action_space
    bytes_per_step = 2 * ((rows * rows + 1) // 2) + 17 + (A + 7) // 8
    r = {"shape": f"{rows}x{rows}x{types}", "boards": n, "ms_per_step": ms, "env_steps_per_s": n / ms * 1e3,
         "algorithmic_bytes_per_step": bytes_per_step, "hbm_gbs_algorithmic": n * bytes_per_step / ms / 1e6,
         "roofline_frac": n * bytes_per_step / ms / 1e6 / peak, "hbm_peak_gbs": peak,
         "handed_off_to_exact_kernel": int(b._scratch[0].item()) if b._scratch is not None else None,
         "lib": os.path.basename(E._native.LIB_PATH),
         "mean_cascades": float(b.cascades.float().mean().item()),
         "flagged": int((b.status != 0).sum().item())}
    out.append(r)
    print(json.dumps(r))
    del b
