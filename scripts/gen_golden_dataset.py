#!/usr/bin/env python
"""Generate tests/golden/dataset_mirror.npz from the UNMODIFIED reference (read-only import from /root/reference):
Dataset.mirror (dataset.py:86-112) applied to random boards (special tokens included) and sparse random policies,
for every engine shape, plus the action permutation it implies.  Authoring container only; vectors are committed."""
import io
import os
import sys
from contextlib import redirect_stderr

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from match3tile.boardConfig import BoardConfig  # noqa: E402
from dataset import Dataset  # noqa: E402

SHAPES = [(9, 9, 6), (6, 6, 4), (12, 12, 7), (16, 16, 8), (5, 5, 2)]
out = {}
rng = np.random.default_rng(20261018)
for R, Cc, T in SHAPES:
    cfg = BoardConfig(seed=1, rows=R, columns=Cc, types=T)
    n = 24
    specials = np.array([cfg.h_line, cfg.v_line, cfg.bomb, cfg.mega_token])
    obs, pol = [], []
    for _ in range(n):
        b = rng.integers(1, T + 1, size=(R, Cc)).astype(np.int64)
        k = int(rng.integers(0, 5))
        for _ in range(k):
            b[rng.integers(0, R), rng.integers(0, Cc)] = specials[rng.integers(0, 4)]
        p = np.zeros(cfg.action_space)
        idx = rng.choice(cfg.action_space, size=int(rng.integers(1, 12)), replace=False)
        p[idx] = rng.integers(1, 100, size=len(idx)) / 128.0
        obs.append(b)
        pol.append(p)
    data = {"observations": list(obs), "policies": list(pol), "values": list(range(n))}
    with redirect_stderr(io.StringIO()):  # tqdm bar
        got = Dataset(cfg).with_mirroring(True).mirror(data)
    assert len(got["values"]) == 2 * n
    perm = np.empty(cfg.action_space, dtype=np.int64)
    for a in range(cfg.action_space):
        (r1, c1), (r2, c2) = cfg.decode(a)
        perm[a] = cfg.encode((r1, Cc - 1 - c1), (r2, Cc - 1 - c2))
    tag = f"{R}x{Cc}x{T}"
    out[f"obs_{tag}"] = np.stack(obs)
    out[f"pol_{tag}"] = np.stack(pol)
    out[f"mobs_{tag}"] = np.stack(got["observations"][n:])
    out[f"mpol_{tag}"] = np.stack(got["policies"][n:])
    out[f"mval_{tag}"] = np.asarray(got["values"][n:])
    out[f"perm_{tag}"] = perm
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "dataset_mirror.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items() if k.startswith("perm")})
