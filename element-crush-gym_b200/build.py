"""Build libecg.so (sm_100a) in-tree with nvcc: one object per board size, compiled in parallel,
plus the C-ABI object; linked into element-crush-gym_b200/lib/libecg.so.

    python element-crush-gym_b200/build.py [--force] [--ptxas-v] [--sizes 9,6]
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libecg.so")
SIZES = tuple(range(4, 17))  # every square size the reference can run (boardConfig.decode needs columns >= 4)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--diag-suppress", "177"]
DEPS = ["ecg_bits.cuh", "ecg_core.cuh", "ecg_ops.h", os.path.join("..", "..", "include", "ecg.h")]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, log):
    p = subprocess.run(cmd, capture_output=True, text=True)
    if log or p.returncode:
        sys.stderr.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return p.stdout + p.stderr


def build(force=False, ptxas_v=False, sizes=SIZES, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    deps = [os.path.join(CSRC, d) for d in DEPS]
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if ptxas_v else []
    jobs = []
    objs = []
    for n in SIZES:
        obj = os.path.join(OBJDIR, f"shape_{n}.o")
        objs.append(obj)
        src = os.path.join(CSRC, "ecg_shape_kernels.cu")
        if n in sizes and (force or _stale(obj, deps + [src])):
            jobs.append([nvcc, *NVCC_FLAGS, *extra, f"-DECG_SIZE={n}", "-c", src, "-o", obj])
    api_obj = os.path.join(OBJDIR, "api.o")
    api_src = os.path.join(CSRC, "ecg_api.cu")
    if force or _stale(api_obj, deps + [api_src]):
        jobs.append([nvcc, *NVCC_FLAGS, *extra, "-c", api_src, "-o", api_obj])
    logs = []
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            logs = list(ex.map(lambda c: _run(c, verbose or ptxas_v), jobs))
    missing = [o for o in objs if not os.path.exists(o)]
    if missing:
        raise RuntimeError(f"missing objects (build all sizes once): {missing}")
    if jobs or force or _stale(LIB, objs + [api_obj]):
        _run([nvcc, "-shared", "-o", LIB, api_obj, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], verbose)
    return LIB, logs


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--ptxas-v", action="store_true")
    ap.add_argument("--sizes", default=",".join(map(str, SIZES)))
    a = ap.parse_args()
    lib, _ = build(a.force, a.ptxas_v, tuple(int(x) for x in a.sizes.split(",")), verbose=True)
    print(lib)
