#!/usr/bin/env python
"""Time / instruction share per source region of a kernel from an ncu report (--set full --import-source on).
Every SASS instruction is attributed to the nearest preceding instruction that maps to a line of ecg_core.cuh or
ecg_shape_kernels.cu (the inlined bit-board helpers of ecg_bits.cuh carry no context of their own); lines are then
grouped into the regions below.  Usage: scripts/ncu_regions.py report.ncu-rep n_boards"""
import bisect, collections, csv, io, re, subprocess, sys

REGIONS = [  # (name, file, first line, last line) in ecg_core.cuh ("K" = ecg_shape_kernels.cu)
    ("philox", "core", 34, 131), ("spawn helpers / crossing", "core", 227, 566), ("find_matches", "core", 567, 611),
    ("legal_swaps", "core", 612, 655), ("pick / select", "core", 656, 800), ("trigger_specials", "core", 801, 865),
    ("gravity", "core", 866, 884), ("refill", "core", 885, 965), ("shuffle", "core", 966, 1015),
    ("special_pair", "core", 1016, 1055), ("step_begin (swap)", "core", 1056, 1125), ("step_iter", "core", 1126, 1200),
]


def export(rep, what):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", what],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, n = sys.argv[1], int(sys.argv[2])
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
            cur = (sec, int(r[0]))
        elif len(r) > 3:
            try:
                addr2line.setdefault(int(r[2], 16), cur)
            except ValueError:
                pass
    rows = export(rep, "sass")
    hdr = rows[1]
    ia, ie, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    isamp = hdr.index("# Samples")
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    anchor = None
    for r in rows[2:]:
        try:
            a = int(r[ia], 16)
        except ValueError:
            continue
        l = addr2line.get(a)
        if l and l[0] in ("ecg_core.cuh", "ecg_shape_kernels.cu"):
            anchor = l
        if anchor is None:
            continue
        v = agg[anchor]
        v[0] += int(r[ie] or 0)
        v[1] += int(r[it] or 0)
        v[2] += int(r[isamp] or 0)
        v[3] += 1
    byline = "--lines" in sys.argv
    out = collections.defaultdict(lambda: [0, 0, 0, 0])
    for (f, line), v in agg.items():
        if byline:
            key = f"{f}:{line}"
        elif f == "ecg_shape_kernels.cu":
            key = "kernel glue (cursor, load/store, state machine)"
        else:
            key = next((nm for nm, _, lo, hi in REGIONS if lo <= line <= hi), f"core:{line}")
        for i in range(4):
            out[key][i] += v[i]
    tot = [sum(v[i] for v in out.values()) for i in range(4)]
    nw = n / 32
    print(f"warp-instr per 32 boards {tot[0] / nw:.0f}; thread-instr per board {tot[1] / n:.0f}; "
          f"avg active threads {tot[1] / tot[0]:.1f}; samples {tot[2]}")
    print(f"{'region':48s} {'time%':>6s} {'instr%':>6s} {'winstr/32b':>10s} {'active':>6s} {'static':>6s}")
    for k, v in sorted(out.items(), key=lambda kv: -kv[1][2])[:60 if byline else 99]:
        print(f"{k:48s} {v[2] / tot[2] * 100:6.2f} {v[0] / tot[0] * 100:6.2f} {v[0] / nw:10.0f} "
              f"{v[1] / max(v[0], 1):6.1f} {v[3]:6d}")


if __name__ == "__main__":
    main()
