#!/usr/bin/env python3
"""Where the local-memory spills of a kernel sit: STL/LDL per code region between BAR.SYNC / CALL / RET markers.
usage: sass_spills.py <lib.so> <mangled-name-substring>"""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, fn = None, {}
for line in names.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        fn[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
        fn[cur].append(line)
for name, lines in fn.items():
    if pat not in name:
        continue
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    print("==", dem[:140], "instructions", len(lines))
    seg, start, stl, ldl = 0, 0, 0, 0
    for i, l in enumerate(lines + ["BAR.SYNC end"]):
        if "STL" in l:
            stl += 1
        if "LDL" in l:
            ldl += 1
        if "BAR.SYNC" in l or i == len(lines):
            print("  region %2d  instr %6d..%6d (%5d)  STL %3d  LDL %3d" % (seg, start, i, i - start, stl, ldl))
            seg, start, stl, ldl = seg + 1, i + 1, 0, 0
