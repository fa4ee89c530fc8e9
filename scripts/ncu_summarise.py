#!/usr/bin/env python3
"""Turn one `ncu --set full --import-source on` capture of the decode kernel into the files kept under profiles/:

    scripts/ncu_summarise.py gpurun_out/prof_wifi_r01j.ncu-rep wifi v4 --frames 16384 --iters 30

writes profiles/r01/ncu_raw_<code>_<tag>.txt (the raw-page metrics that matter here), profiles/r01/
ncu_source_summary_<code>_<tag>.txt (executed warp-instructions per SASS opcode and pipe class, stall samples,
per-phase segments) and updates the code's entry in profiles/r01/traffic.json, which bench.py reads for
`roofline.traffic` and the ALU-pipe roofline."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("dram__bytes", "gpu__time_duration", "l1tex__data_bank_conflicts", "l1tex__data_pipe_lsu_wavefronts",
        "l1tex__t_requests_pipe_lsu_mem_local", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit", "launch__",
        "lts__t_sector_hit_rate", "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__inst_executed",
        "sm__issue_active", "sm__pipe_", "sm__throughput", "sm__warps_active", "smsp__average_warps_issue_stalled",
        "smsp__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active", "smsp__warps_eligible",
        "TPC.TriageCompute.sm__inst_executed_pipe_alu")
ALU = ("LOP3", "VIMNMX", "SHF", "IADD3", "PRMT", "VIADD", "ISETP", "SEL", "LEA", "PLOP3", "POPC", "FLO", "IABS",
       "VIADDMNMX", "IMNMX", "BMSK", "SGXT")


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("code")
    ap.add_argument("tag")
    ap.add_argument("--frames", type=int, required=True, help="frames decoded by the captured launch")
    ap.add_argument("--iters", type=float, default=30.0, help="average iterations per frame in the captured launch")
    ap.add_argument("--round", default="r02")
    ap.add_argument("--outdir", default=None, help="write here instead of profiles/<round> (e.g. gpurun_out/summaries on the GPU box)")
    args = ap.parse_args()
    outdir = args.outdir or os.path.join(ROOT, "profiles", args.round)
    os.makedirs(outdir, exist_ok=True)

    raw = ncu(args.report, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    metric = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    with open(os.path.join(outdir, "ncu_raw_%s_%s.txt" % (args.code, args.tag)), "w") as fh:
        fh.write("# %s, kernel %s\n" % (os.path.basename(args.report), metric.get("Kernel Name", ("?", ""))[0]))
        for h in hdr:
            if h.startswith(KEEP):
                fh.write("%-100s %s %s\n" % (h, metric[h][0], metric[h][1]))

    src_csv = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True, text=True,
                             check=True).stdout
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(outdir, "ncu_source_summary_%s_%s.txt" % (args.code, args.tag)), "w") as fh:
        fh.write(subprocess.run([sys.executable, os.path.join(here, "ncu_source_summary.py"), "--stalls"], input=src_csv,
                                capture_output=True, text=True).stdout)
        fh.write("\n# code regions split at BAR.SYNC (scripts/ncu_phase_segments.py), in SASS order; the two large ones are\n"
                 "# the variable phase and the check phase, the ones before them the refill, the last one starts with the\n"
                 "# wait at the barrier that ends the check phase\n")
        fh.write(subprocess.run([sys.executable, os.path.join(here, "ncu_phase_segments.py")], input=src_csv,
                                capture_output=True, text=True).stdout)

    rows = list(csv.reader(io.StringIO(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {n: i for i, n in enumerate(rows[hi])}
    alu = total = 0
    for r in rows[hi + 1:]:
        if len(r) < len(rows[hi]):
            continue
        try:
            ex = int(r[col["Instructions Executed"]])
        except ValueError:
            continue
        parts = r[col["Source"]].split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        total += ex
        if op.startswith(ALU):
            alu += ex

    def num(name):
        v, u = metric[name]
        x = float(v.replace(",", ""))
        return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)

    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    path = os.path.join(outdir, "traffic.json")
    try:
        with open(path) as fh:
            doc = json.load(fh)
    except Exception:
        doc = {}
    doc["_comment"] = ("per code, from one `ncu --set full` capture of the packed decode kernel (profiles/<round>/ncu_raw_<code>_<tag>.txt): "
                       "dram__bytes_read.sum + dram__bytes_write.sum divided by the frames of the launch; pipe utilisation; "
                       "executed warp-instructions per frame-iteration (all / ALU pipe, from the SASS source page)")
    fi = args.frames * args.iters
    doc[args.code] = {
        "dram_bytes_per_frame": round(dram / args.frames, 1), "frames_in_capture": args.frames, "avg_iters": args.iters,
        "kernel_ms": num("gpu__time_duration.sum"),
        # over the ELAPSED cycles of the launch (idle SM cycles and the tail included): what bench.py scales
        "issue_elapsed_pct": num("sm__inst_executed.avg.pct_of_peak_sustained_elapsed"),
        "alu_elapsed_pct": num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "issue_per_cycle": num("smsp__issue_active.avg.per_cycle_active"),
        "warp_inst_per_frame_iter": round(total / fi, 1), "alu_warp_inst_per_frame_iter": round(alu / fi, 1),
        "source": "profiles/%s/ncu_raw_%s_%s.txt" % (args.round, args.code, args.tag),
    }
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print(json.dumps(doc[args.code]))


if __name__ == "__main__":
    main()
