#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` dump: executed warp-instructions per SASS opcode and per
source line, plus stall samples.  Usage: ncu -i X.ncu-rep --page source --csv | ncu_source_summary.py"""
import csv
import sys
from collections import Counter, defaultdict

rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
ops, stalls_by_op = Counter(), Counter()
total, samples = 0, 0
sass = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    try:
        ex = int(r[col["Instructions Executed"]])
        smp = int(r[col["# Samples"]])
    except ValueError:
        continue
    parts = src.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.rstrip(";")
    ops[op] += ex
    stalls_by_op[op] += smp
    total += ex
    samples += smp
    sass.append((ex, smp, src, r))
print("total warp-instructions executed: %d, samples %d" % (total, samples))
ALU = ("LOP3", "VIMNMX", "SHF", "IADD3", "PRMT", "VIADD", "ISETP", "SEL", "LEA", "PLOP3", "POPC", "FLO", "IABS", "VIADDMNMX", "IMNMX", "BMSK", "SGXT")
FMA = ("IMAD", "IDP", "FFMA", "FMUL", "FADD")
LSU = ("LDS", "STS", "LDG", "STG", "ATOM", "RED", "LDL", "STL", "LDC", "SHFL")
cls = Counter()
for op, n in ops.items():
    k = "other"
    for name, grp in (("alu", ALU), ("fma", FMA), ("lsu", LSU)):
        if any(op.startswith(g) for g in grp):
            k = name
            break
    if op.startswith("IMAD.HI"):
        k = "fma(half)"
    cls[k] += n
print("by pipe class:", {k: "%.1f%%" % (100.0 * v / total) for k, v in cls.most_common()})
print("%-28s %12s %7s %9s" % ("opcode", "executed", "share", "samples%"))
for op, n in ops.most_common(28):
    print("%-28s %12d %6.1f%% %8.1f%%" % (op, n, 100.0 * n / total, 100.0 * stalls_by_op[op] / max(1, samples)))
if "--stalls" in sys.argv:
    names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    agg = Counter()
    for ex, smp, src, r in sass:
        for n in names:
            try:
                agg[n] += int(r[col[n]])
            except ValueError:
                pass
    tot = sum(agg.values())
    print("stall samples:", {k: "%.1f%%" % (100.0 * v / tot) for k, v in agg.most_common(10)})
