#!/usr/bin/env python
"""Static instruction footprint of the frequently executed code of a kernel, from an ncu report
(--set full --import-source on): how many SASS instructions run in >= X% of warp-loop trips, attributed to
the enclosing function of ecg_core.cuh.  Usage: scripts/ncu_hotset.py report.ncu-rep [trips]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    sec = hdr = cur = None
    recs = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            sec = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 10:
            continue
        if r[2] == "-":
            cur = (sec, int(r[0]))
            continue
        try:
            a = int(r[2], 16)
        except ValueError:
            continue
        if a not in recs:
            recs[a] = (cur, int(r[hdr.index("Instructions Executed")] or 0))
    mx = max(e for _, e in recs.values())
    trips = float(sys.argv[2]) if len(sys.argv) > 2 else None
    if trips is None:  # the ballot at the top of the loop runs once per trip per warp
        cnt = collections.Counter(e for _, e in recs.values() if e > 0)
        trips = max((e for e, c in cnt.items() if c >= 8), default=mx)
    lines = open(os.path.join(ROOT, "element-crush-gym_b200", "csrc", "ecg_core.cuh")).read().split("\n")

    def fn_of(line):
        for i in range(min(line, len(lines)) - 1, 0, -1):
            m = re.match(r"\s*ECG_(?:HD|PHASE|HD_NOINLINE)\s+.*?(\w+)\(", lines[i])
            if m:
                return m.group(1)
        return "?"

    ctx = None
    tot = {0.5: collections.Counter(), 0.1: collections.Counter(), 0.01: collections.Counter()}
    dyn = collections.Counter()
    for a in sorted(recs):
        cur, e = recs[a]
        if cur and cur[0] == "ecg_core.cuh":
            ctx = "core:" + fn_of(cur[1])
        elif cur and cur[0] == "ecg_shape_kernels.cu":
            ctx = "kernel"
        dyn[ctx] += e
        for th in tot:
            if e >= th * trips:
                tot[th][ctx] += 1
    print(f"static instructions: {len(recs)}; trips (warp-level) ~ {trips:.0f}")
    for th in (0.5, 0.1, 0.01):
        n = sum(tot[th].values())
        print(f"executed in >= {th * 100:.0f}% of trips: {n} instructions = {n * 16 / 1024:.1f} KB")
    print("per function (>= 10% of trips | dynamic instructions per trip):")
    for k, v in tot[0.1].most_common():
        print(f"  {v:5d}  {dyn[k] / trips:7.1f}  {k}")


if __name__ == "__main__":
    main()
