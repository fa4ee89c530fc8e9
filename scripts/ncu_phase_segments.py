#!/usr/bin/env python3
"""Split an `ncu --page source --csv` dump of the decode kernel at its BAR.SYNC instructions (phase boundaries in
SASS order) and report, per code region: share of executed instructions, share of stall samples, ALU fraction and
the dominant stall reasons.  Usage: ncu -i X.ncu-rep --page source --csv | ncu_phase_segments.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
ALU = ("LOP3", "VIMNMX", "SHF", "IADD3", "PRMT", "VIADD", "ISETP", "SEL", "LEA", "PLOP3", "POPC", "FLO")
seg, segs = 0, {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    try:
        ex = int(r[col["Instructions Executed"]])
        smp = int(r[col["# Samples"]])
    except ValueError:
        continue
    d = segs.setdefault(seg, {"ex": 0, "smp": 0, "n": 0, "alu": 0, "stalls": {}})
    d["ex"] += ex
    d["smp"] += smp
    d["n"] += 1
    parts = src.split()
    op = parts[1] if src.startswith("@") and len(parts) > 1 else parts[0]
    if any(op.startswith(x) for x in ALU):
        d["alu"] += ex
    for n in names:
        try:
            d["stalls"][n] = d["stalls"].get(n, 0) + int(r[col[n]])
        except ValueError:
            pass
    if "BAR.SYNC" in src:
        seg += 1
tot_s = sum(d["smp"] for d in segs.values())
tot_e = sum(d["ex"] for d in segs.values())
for k, d in segs.items():
    if d["ex"] < tot_e * 0.002 and d["smp"] < tot_s * 0.002:
        continue
    top = sorted(d["stalls"].items(), key=lambda x: -x[1])[:6]
    ts = sum(d["stalls"].values()) or 1
    print("seg %2d sass#%5d exec %5.1f%% samples %5.1f%%  alu-frac %.2f  rel-ipc %.2f | %s" % (
        k, d["n"], 100 * d["ex"] / tot_e, 100 * d["smp"] / tot_s, d["alu"] / max(1, d["ex"]),
        (d["ex"] / tot_e) / (d["smp"] / tot_s) if d["smp"] else 0,
        ", ".join("%s %.0f%%" % (n[6:], 100 * v / ts) for n, v in top)))
