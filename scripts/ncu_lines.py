#!/usr/bin/env python
"""SASS listing of a kernel from an ncu report (--import-source on, built with -lineinfo) in address order, every
instruction with the source line ncu attributes it to, its executions per 32 boards and active threads -- the map from
the segments of scripts/ncu_segments.py back to the lines of ecg_core.cuh / ecg_shape_kernels.cu.
Usage: scripts/ncu_lines.py report.ncu-rep n_boards [first_sass_index last_sass_index]"""
import csv
import io
import os
import subprocess
import sys


def main():
    rep, n = sys.argv[1], int(sys.argv[2])
    lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    by_addr = {}
    cur_file, cur_line, hdr = "", "", None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = os.path.basename(r[1])
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 8:
            if r[0]:
                cur_line = r[0]
            elif r[2].startswith("0x"):
                ie, it = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
                by_addr[int(r[2], 16)] = (cur_file, cur_line, r[3].strip(), int(r[ie] or 0), int(r[it] or 0))
    nw = n / 32
    for i, a in enumerate(sorted(by_addr)):
        if lo <= i <= hi:
            f, ln, ins, e, t = by_addr[a]
            print(f"{i:5d} {f}:{ln:<5s} {e / nw:7.3f} {t / max(e, 1):5.1f}  {ins[:70]}")


if __name__ == "__main__":
    main()
