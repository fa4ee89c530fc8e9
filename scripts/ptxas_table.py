#!/usr/bin/env python3
"""Registers / spills of every kernel from `python fixedpointldpc_b200/build.py -v` output (ptxas -v)."""
import re
import subprocess
import sys

t = open(sys.argv[1] if len(sys.argv) > 1 else "/tmp/build_v.log").read()
for b in re.split(r"ptxas info\s+: Compiling entry function '", t)[1:]:
    name = b.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    m = re.search(r"Used (\d+) registers", b)
    sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
    print(dem[:170].replace("ldpc::", ""), "| regs", m.group(1) if m else None, "| spill", sp.groups() if sp else None)
