#!/usr/bin/env python
"""Straight-line segments of a kernel's SASS with their execution counts, from an ncu report (--set full
--import-source on): consecutive instructions with the same (executions, thread-executions) form one segment; printed
are executions per 32 boards, active threads per execution and warp-instructions per 32 boards of every segment above a
cost threshold.  A segment that runs more often than the trip count at partial occupancy is divergent code that could
have run once (this is how the twice-per-trip Philox block of r03 was found).
Usage: scripts/ncu_segments.py report.ncu-rep n_boards [min_cost]"""
import csv
import io
import subprocess
import sys


def main():
    rep, n = sys.argv[1], int(sys.argv[2])
    min_cost = float(sys.argv[3]) if len(sys.argv) > 3 else 12.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ie, it, isrc = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
    nw = n / 32
    segs = []
    for i, r in enumerate(rows[2:]):
        if len(r) <= it:
            continue
        e, t = int(r[ie] or 0), int(r[it] or 0)
        if segs and segs[-1][2] == e and segs[-1][3] == t:
            segs[-1][1] = i
            segs[-1][4] += 1
        else:
            segs.append([i, i, e, t, 1, r[isrc].strip().split(";")[0][:44]])
    total = sum(s[2] * s[4] for s in segs) / nw
    print(f"{rows[0][1][:100]}\nwarp-instructions per 32 boards: {total:.0f}")
    print(f"{'sass index':>13s} {'len':>4s} {'exec/32b':>8s} {'active':>6s} {'winstr/32b':>10s}  first instruction")
    for a, b, e, t, ln, src in segs:
        cost = e * ln / nw
        if cost >= min_cost:
            print(f"{a:6d}-{b:6d} {ln:4d} {e / nw:8.3f} {t / max(e, 1):6.1f} {cost:10.1f}  {src}")


if __name__ == "__main__":
    main()
