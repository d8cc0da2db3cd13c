#!/usr/bin/env python
"""Key counters of the kernels in ncu reports (--set full) as one JSON object: scripts/ncu_summary.py name=report.ncu-rep ..."""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]
STALLS = "smsp__average_warp_latency_issue_stalled_"


def summary(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, val = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, val):
        if h in KEEP or h == "Kernel Name":
            d[h] = f"{v} {u}".strip()
        elif h.startswith(STALLS) and h.endswith(".ratio"):
            d.setdefault("stall_cycles_per_issued_instruction", {})[h[len(STALLS):-6]] = round(float(v or 0), 3)
    st = d.get("stall_cycles_per_issued_instruction")
    if st:
        d["stall_cycles_per_issued_instruction"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:10])
    return d


if __name__ == "__main__":
    res = {}
    for a in sys.argv[1:]:
        name, rep = a.split("=", 1)
        res[name] = summary(rep)
    print(json.dumps(res, indent=1))
