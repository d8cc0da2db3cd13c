#!/usr/bin/env python
"""Time / instruction share per source region of a kernel from an ncu report (--set full --import-source on).
Every SASS instruction is attributed to the nearest preceding instruction that maps to a line of ecg_core.cuh or
ecg_shape_kernels.cu (the inlined bit-board helpers of ecg_bits.cuh carry no context of their own); lines are then
grouped into the regions below.  Usage: scripts/ncu_regions.py report.ncu-rep n_boards"""
import bisect, collections, csv, io, re, subprocess, sys

import os

CORE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "element-crush-gym_b200", "csrc", "ecg_core.cuh")
GROUPS = {  # function of ecg_core.cuh -> region (anything else keeps its own name)
    "philox4x32_10": "philox", "init": "philox / digits", "u32": "philox / digits", "digit": "philox / digits",
    "below": "philox / digits", "seek": "philox / digits", "preset_block": "philox / digits",
    "philox_pick": "philox", "derive": "find_matches", "eq_at": "find_matches", "add_spawn": "crossing / rare spawns",
    "add_spawn_disjoint": "find_matches", "run_right": "crossing / rare spawns", "run_down": "crossing / rare spawns",
    "mega_spawns": "crossing / rare spawns", "corner_spawns": "crossing / rare spawns",
    "merged_spawns": "crossing / rare spawns", "scan_order_matches": "crossing / rare spawns",
    "straight_spawn": "crossing / rare spawns", "single_cross_matches": "crossing / rare spawns",
    "crossing_matches": "crossing / rare spawns", "long_run_spawns": "crossing / rare spawns",
    "swaps_to_actions": "pick / select", "mask_count": "pick / select", "mask_select": "pick / select",
    "swaps_count": "pick / select", "swaps_select_bit": "pick / select", "action_of_swap": "pick / select",
    "swaps_select": "pick / select", "decode_action": "pick / select", "fill_rows_with_any": "trigger_specials",
    "fill_cols_with_any": "trigger_specials", "shuffle_rows_impl": "shuffle", "shuffle_rows": "shuffle",
    "special_pair_impl": "special_pair", "step_begin_at": "swap (step_begin)", "step_begin": "swap (step_begin)",
    "step_iter": "step_iter (clear, points, glue)",
}


def core_functions():
    """[(first line, name)] of the functions defined in ecg_core.cuh, in file order"""
    out = []
    pat = re.compile(r"^\s*(?:ECG_\w+|static|inline)\b[^;=]*?\b(\w+)\s*\([^;]*$")
    for i, line in enumerate(open(CORE), 1):
        m = pat.match(line)
        if m and not line.lstrip().startswith("//"):
            out.append((i, m.group(1)))
    return out


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
    funcs = core_functions()
    starts = [f[0] for f in funcs]
    out = collections.defaultdict(lambda: [0, 0, 0, 0])
    for (f, line), v in agg.items():
        if byline:
            key = f"{f}:{line}"
        elif f == "ecg_shape_kernels.cu":
            key = "kernel glue (cursor, load/store, state machine)"
        else:
            k = bisect.bisect_right(starts, line) - 1
            fn = funcs[k][1] if k >= 0 else f"core:{line}"
            key = GROUPS.get(fn, fn)
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
