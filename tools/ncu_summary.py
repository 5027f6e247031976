#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the small CSV kept under profiles/ (one column per kernel, first captured
launch of each).  Usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep > profiles/rNN_ncu_full_X_summary.csv"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels, seen = [], set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name not in seen:
            seen.add(name); kernels.append((name, r))
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit"] + [k for k, _ in kernels])
    for m in METRICS:
        if m in idx:
            w.writerow([m, units[idx[m]]] + [r[idx[m]] for _, r in kernels])


if __name__ == "__main__":
    main()
