#!/usr/bin/env python
"""Per-phase instruction breakdown of a kernel from an ncu report taken on the -DECG_PROFILE_PHASES build
(major phases kept out of line).  Usage: scripts/ncu_phases.py report.ncu-rep [n_boards]"""
import bisect
import collections
import csv
import io
import re
import subprocess
import sys


def export(rep, what):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", what],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
    rows = export(rep, "sass")
    hdr = rows[1]
    ia, isrc = hdr.index("Address"), hdr.index("Source")
    ie, it = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    ins, targets = [], set()
    for r in rows[2:]:
        try:
            a = int(r[ia], 16)
        except ValueError:
            continue
        ins.append((a, r[isrc], int(r[ie] or 0), int(r[it] or 0)))
        m = re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", r[isrc])
        if m:
            targets.add(int(m.group(1), 16))
    starts = sorted(targets | {ins[0][0]})
    seg = collections.defaultdict(lambda: [0, 0, 0])
    segof = {}
    for a, s, e, t in ins:
        k = starts[bisect.bisect_right(starts, a) - 1]
        segof[a] = k
        seg[k][0] += e
        seg[k][1] += t
        seg[k][2] += 1
    both = export(rep, "cuda,sass")
    addr2line, cur, sec = {}, None, None
    for r in both:
        if not r:
            continue
        if r[0] == "File Path":
            sec = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if len(r) > 3 and r[2] == "-":
            cur = (sec, r[0], r[1].strip()[:60])
        elif len(r) > 3:
            try:
                addr2line.setdefault(int(r[2], 16), cur)
            except ValueError:
                pass
    nw = n / 32
    tot = sum(v[0] for v in seg.values())
    tt = sum(v[1] for v in seg.values())
    print(f"warp-instructions per 32 boards: {tot / nw:.0f}; avg active threads {tt / tot:.1f}")
    for k, v in sorted(seg.items(), key=lambda kv: -kv[1][0]):
        c = collections.Counter(addr2line.get(a) for a, _, _, _ in ins if segof[a] == k and addr2line.get(a))
        names = "; ".join(f"{x[0]}:{x[1]} {x[2]}" for x, _ in c.most_common(2))
        print(f"{v[0] / tot * 100:6.2f}% ({v[0] / nw:6.0f}/32 boards) thr {v[1] / max(v[0], 1):5.1f} static {v[2]:5d}  {names}")


if __name__ == "__main__":
    main()
