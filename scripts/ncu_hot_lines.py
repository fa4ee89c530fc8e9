#!/usr/bin/env python3
"""The SASS lines of an `ncu --page source --csv` dump that collect the most stall samples, with their executed counts
and the two largest stall reasons.   usage: zcat X_source.csv.gz | ncu_hot_lines.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
out, total = [], 0
for idx, r in enumerate(rows[hdr_i + 1:]):
    if len(r) < len(hdr):
        continue
    try:
        ex, smp = int(r[col["Instructions Executed"]]), int(r[col["# Samples"]])
    except ValueError:
        continue
    total += smp
    st = sorted(((int(r[col[n]] or 0), n[6:]) for n in names), reverse=True)[:2]
    out.append((smp, idx, r[col["Address"]], ex, " ".join(r[col["Source"]].split()), st))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
print("samples %d; top %d lines" % (total, n))
for smp, idx, addr, ex, src, st in sorted(out, reverse=True)[:n]:
    print("%5.2f%%  sass#%5d  exec %10d  %-70s %s" % (100.0 * smp / max(1, total), idx, ex, src[:70], ", ".join("%s %d" % (b, a) for a, b in st if a)))
