#!/usr/bin/env python
"""Prints the metrics of an .ncu-rep (one kernel) that the profiles/ summaries quote.

usage: ncu_summary.py report.ncu-rep  [> profiles/rNN_..._ncu_summary.txt]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {k: i for i, k in enumerate(hdr)}
    print(f"Kernel Name [] = {vals[col['Kernel Name']]}")
    for k in KEYS:
        if k in col:
            print(f"{k} [{units[col[k]]}] = {vals[col[k]]}")
    cyc = float(vals[col["sm__cycles_elapsed.avg"]].replace(",", ""))
    counts = {}
    for op in ("dfma", "dmul", "dadd"):
        k = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"
        counts[op] = float(vals[col[k]].replace(",", "")) * cyc
    print("\nFP64 thread instructions (per_cycle_elapsed x elapsed cycles):")
    for op in ("dfma", "dmul", "dadd"):
        print(f"  {op} = {counts[op]:.4e}")
    flops = counts["dadd"] + counts["dmul"] + 2 * counts["dfma"]
    print(f"executed FP64 flops (dadd + dmul + 2*dfma) = {flops:.4e}")
    warp_fp64 = (counts["dfma"] + counts["dmul"] + counts["dadd"]) / 32
    inst = float(vals[col["smsp__inst_executed.sum"]].replace(",", ""))
    print(f"FP64 share of executed warp instructions = {warp_fp64 / inst:.3f}")


if __name__ == "__main__":
    main()
