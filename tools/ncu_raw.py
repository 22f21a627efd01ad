#!/usr/bin/env python3
"""Print a fixed set of raw counters (+ the stall reasons above 0.15) of the first kernel in an ncu report:
    python tools/ncu_raw.py rep.ncu-rep [kernel regex]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"] + (["--kernel-name", f"regex:{sys.argv[2]}"] if len(sys.argv) > 2 else [])
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_active.min", "sm__cycles_elapsed.avg",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.avg", "smsp__inst_executed.max", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size"]
for k in want:
    if k in hdr:
        i = hdr.index(k)
        print(k, r[i], units[i])
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        try:
            if float(r[i]) > 0.15:
                print("  stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), r[i])
        except ValueError:
            pass
