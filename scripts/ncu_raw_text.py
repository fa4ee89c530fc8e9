#!/usr/bin/env python3
"""Raw-page metrics of every kernel in an .ncu-rep as text (the subset profiles/ keeps), one block per launch.
usage: ncu_raw_text.py report.ncu-rep > out.txt"""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes", "dram__throughput", "gpu__time_duration", "launch__", "l1tex__data_bank_conflicts", "lts__t_sector_hit_rate",
        "sm__cycles_elapsed.avg", "sm__inst_executed.avg.pct", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma",
        "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_xu", "sm__pipe_alu_cycles_active.avg.pct",
        "sm__pipe_fp64_cycles_active.avg.pct", "sm__throughput", "sm__warps_active", "smsp__issue_active.avg", "smsp__warps_eligible",
        "smsp__average_warps_issue_stalled")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    m = dict(zip(hdr, r))
    print("== kernel %s  grid %s block %s" % (m.get("Kernel Name"), m.get("Grid Size"), m.get("Block Size")))
    for i, h in enumerate(hdr):
        if h.startswith(KEEP):
            print("%-100s %s %s" % (h, r[i], units[i]))
    print()
