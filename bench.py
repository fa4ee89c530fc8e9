#!/usr/bin/env python3
"""Headline benchmark: decoded info Gbit/s per B200 @30 iterations (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--code wifi|a5|c79|a24] [--impl reference]

A "step" is one pass of the decode path over one batch of synthetic frames.  The default
workload is BASELINE.json configs[1]: the 802.11n n=1944 rate-1/2 code (H_802.11_IndZero),
FRAC_WIDTH=4 quantised LLRs, MAX_ITER=30, at an Eb/N0 (0 dB) where every frame runs all 30
iterations under the reference's own early-termination rule, so `value` is
the "@30 iters" number with nothing skipped; the 2 dB operating point with early termination
is reported beside it as `operating_point`.

One process per GPU (torchrun for N > 1): frames are independent, so every rank decodes its own
shard (weak scaling) and NCCL only all-reduces the iteration / frame counters.

`--impl reference` times the reference's own CPU decoder (oracle/_ref/libref_<code>.so, the
unmodified sources of /root/reference compiled by oracle/build_ref.py) on all host cores, one
process per core (the reference is not re-entrant), on a bounded sample of the same frames.
If that binary is absent the plain-C port (oracle/ldpc_oracle.c) is timed instead.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # code: (name, Eb/N0 where all frames run 30 iterations, operating Eb/N0, channel rate, frames per step)
    "wifi": ("802.11n n=1944 R=1/2 (H_802.11_IndZero), FRAC_WIDTH=4, MAX_ITER=30", 0.0, 2.0, 0.5, 1 << 17),
    "a5": ("array p=47 r=5 n=2209 (H_array_p47_r5_forward), FRAC_WIDTH=4, MAX_ITER=30", 2.0, 4.5, None, 1 << 17),
    "c79": ("array cut79 n=2212 (H2212_316_array_cut79), FRAC_WIDTH=4, MAX_ITER=30", 2.0, 4.5, None, 1 << 17),
    "a24": ("array p=47 r=24 n=2209 (H_array_p47_r24_forward), FRAC_WIDTH=4, MAX_ITER=30", 3.0, 6.0, None, 1 << 15),
}
PRECHECK = {"wifi": False, "a5": True, "c79": False, "a24": True}  # decode_general_fp vs decode_fixpoint
# frames per host core for the cpu_baseline leg of the GPU arm: about 10 s of the reference decoder at 30 iterations
CPU_SAMPLE_PER_CORE = {"wifi": 4096, "a5": 3072, "c79": 3072, "a24": 384}


def channel_rate(code_name, code):
    import fixedpointldpc_b200 as fp
    r = WORKLOADS[code_name][3]
    if r is not None:
        return r
    if code_name == "c79":
        return fp.codes.INFO_BITS["c79"] / code.n
    return code.rate  # ROM::getRate, ArrayLDPCMacro.h:60


def make_frames(torch, code, rate, ebn0_db, frames, seed, device, dtype):
    """All-zero codeword over BPSK/AWGN, quantised like DecodeTrial (PerfTest.cpp:159-170):
    LLR_fp = int(2*snr*(1 + N(0, sigma)) * 2^4), snr = 2*10^(dB/10)*R, sigma = sqrt(1/snr)."""
    snr = 2.0 * 10.0 ** (ebn0_db / 10.0) * rate
    sigma = math.sqrt(1.0 / snr)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((frames, code.n), dtype=dtype, device=device)
    chunk = 1 << 14
    for s in range(0, frames, chunk):
        e = min(frames, s + chunk)
        z = torch.randn((e - s, code.n), generator=g, device=device, dtype=torch.float64)
        llr = 2.0 * snr * (1.0 + sigma * z) * 16.0
        out[s:e] = torch.trunc(llr).to(dtype)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.thread = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.02)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    code_name, llr, fixpoint, use_ref, tables = args
    from oracle import pyoracle as po
    if use_ref:
        ref = po.Reference(code_name)
        if code_name == "wifi":
            pass
        ref.set_tables(po.Tables(*tables))
        t0 = time.perf_counter()
        iters = ref.decode_many(llr, fixpoint)
        dt = time.perf_counter() - t0
    else:
        orc = po.Oracle(po.Tables(*tables))
        t0 = time.perf_counter()
        iters = orc.decode_many(llr, precheck=fixpoint)
        dt = time.perf_counter() - t0
    return dt, iters


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_decode(code_name, code, llr_host, cores):
    """Decode llr_host [frames][n] int32 on `cores` processes; returns (seconds, iters, kind)."""
    import multiprocessing as mp
    from oracle import pyoracle as po
    from oracle import build_ref
    build_ref.build_oracle()
    use_ref = po.reference_available(code_name)
    vdeg, cdeg, vlist, clist = code.tables()
    tables = (code.n, code.m, vdeg, cdeg, vlist, clist)
    parts = np.array_split(llr_host, cores)
    jobs = [(code_name, np.ascontiguousarray(p), PRECHECK[code_name], use_ref, tables) for p in parts if len(p)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(len(jobs)) as pool:
        pool.map(_cpu_worker, [(code_name, j[1][:1], j[2], j[3], j[4]) for j in jobs])  # page in, warm up
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    iters = np.concatenate([r[1] for r in res])
    return wall, iters, ("reference" if use_ref else "port")


def run_reference_arm(args):
    import fixedpointldpc_b200 as fp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name, ebn0, _, _, _ = WORKLOADS[args.code]
    code = fp.codes.NAMED[args.code]()
    k = fp.codes.INFO_BITS[args.code]
    rate = channel_rate(args.code, code)
    cores = os.cpu_count() or 1
    import torch
    per_core = args.cpu_frames_per_core
    frames = per_core * cores
    llr = make_frames(torch, code, rate, ebn0, frames, 20261018, "cpu", torch.int32).numpy()
    times, total_it = [], 0
    for step in range(args.warmup + args.steps):
        wall, iters, kind = cpu_decode(args.code, code, llr, cores)
        if step >= args.warmup:
            times.append(wall)
            total_it += int(iters.sum())
    t = sum(times)
    value = frames * args.steps * k / t / 1e9
    line = {
        "impl": "reference", "metric": "decoded info Gbit/s @30 iters", "value": value, "unit": "Gbit/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": name, "ebn0_db": ebn0, "frames_per_step": frames,
                   "avg_iters": total_it / (frames * args.steps)},
        "cpu_baseline": {"value": value, "unit": "Gbit/s", "cores": cores, "kind": kind, "cpu": cpu_model(),
                         "sample": "%d frames per step (%d per core), same channel as the GPU arm" % (frames, per_core),
                         "frames_per_s": frames * args.steps / t},
        "e2e": {"value": value, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import fixedpointldpc_b200 as fp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    name, ebn0, ebn0_op, _, frames = WORKLOADS[args.code]
    if args.frames:
        frames = args.frames
    code = fp.codes.NAMED[args.code]()
    k = fp.codes.INFO_BITS[args.code]
    rate = channel_rate(args.code, code)
    dec = fp.Decoder(code, max_iter=30, precheck=PRECHECK[args.code], device=local, precision=args.precision,
                     threads=args.threads, frames_per_cta=args.frames_per_cta)
    # a dedicated non-default stream: the C ABI treats a NULL stream as "the decoder's own stream",
    # and torch.cuda.Event only sees work on the stream it is recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    # inputs resident in HBM; one distinct batch per rank (weak scaling)
    llr16 = make_frames(torch, code, rate, ebn0, frames, 20261018 + rank, dev, torch.int16)
    iters = torch.zeros(frames, dtype=torch.int32, device=dev)
    bits = torch.zeros((frames, code.nw32), dtype=torch.int32, device=dev)

    def step_device(llr):
        dec.decode_device(llr.data_ptr(), 16, frames, iters.data_ptr(), bits.data_ptr(), None, None,
                          stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = dec.stats()["kernel_launches"]
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), dec.stats()["kernel_launches"] - l0

    with ClockSampler(local) as clocks:
        ms_total, launches = timed(lambda: step_device(llr16), args.steps, args.warmup)
    clk = clocks.summary()
    counters = torch.stack([iters.clamp(min=0).sum().to(torch.int64), torch.tensor(frames, device=dev)])
    if world > 1:
        dist.all_reduce(counters)  # the only data-path-adjacent collective: iteration / frame counters
    total_frames = int(counters[1].item())
    avg_iters = float(counters[0].item()) / total_frames
    fallback = dec.stats()["fallback_frames"]
    ms_step = ms_total / args.steps
    fps = total_frames / (ms_step * 1e-3)
    value = fps * k / 1e9

    # operating point (early termination active), same engine, per rank
    llr_op = make_frames(torch, code, rate, ebn0_op, frames, 777 + rank, dev, torch.int16)
    ms_op, _ = timed(lambda: step_device(llr_op), max(1, args.steps // 2), 1)
    ms_op /= max(1, args.steps // 2)
    it_op = float(iters.clamp(min=0).sum().item()) / frames
    del llr_op

    # end to end through the reference-facing C-ABI call on pinned HOST buffers (int32 like `const int *LLR`)
    e2e_frames = min(frames, args.e2e_frames)
    h_llr = torch.empty((e2e_frames, code.n), dtype=torch.int32).pin_memory()
    h_llr.copy_(llr16[:e2e_frames].to(torch.int32).cpu())
    h_iters = torch.empty(e2e_frames, dtype=torch.int32).pin_memory()
    h_bits = torch.empty((e2e_frames, code.nw32), dtype=torch.int32).pin_memory()

    def step_host():
        dec.decode_raw(h_llr.data_ptr(), e2e_frames, h_iters.data_ptr(), h_bits.data_ptr())

    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, args.steps // 2)
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_frames * e2e_steps * k / float(t_e2e.item()) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    algo_bytes = 2 * code.n + (code.n + 7) // 8 + 4  # int16 LLR in + packed bits + iteration count out
    per_gpu_fps = fps / world
    achieved = per_gpu_fps * algo_bytes / 1e9
    # DRAM bytes per launch and the utilisation of the bound resource, from the committed ncu capture of this kernel
    # (profiles/r01/traffic.json, written by scripts/ncu_summarise.py)
    traffic, cap = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01", "traffic.json")) as fh:
            cap = json.load(fh)[args.code]
        traffic = cap["dram_bytes_per_frame"] * frames
    except Exception:
        pass
    # algorithmic integer work: 18 scalar ops per necessary sxor, 3 per edge, 1 per variable
    _, cdeg, _, _ = code.tables()
    sxors = int((3 * cdeg - 6).sum())
    ops_iter = 18 * sxors + 3 * code.edges + code.n
    sm_mhz = clk["sm_mhz"] or float(peaks.get("sm_max_mhz", 1965.0))
    # The bound resource is the ALU pipe (LOP3 / SHF / VIMNMX / VIADD / PRMT issue slots, 64 lanes per clock and SM).
    # ncu measured its utilisation for one launch of this kernel; the live figure is that utilisation scaled by
    # (frame-iterations/s now) / (frame-iterations/s of the captured launch) -- same kernel, same work per frame-iteration.
    alu_peak = 148 * 64 * sm_mhz * 1e6 / 1e12
    alu = None
    if cap and cap.get("kernel_ms"):
        cap_rate = cap["frames_in_capture"] * cap["avg_iters"] / (cap["kernel_ms"] * 1e-3)
        frac = cap["alu_pipe_pct"] / 100.0 * (per_gpu_fps * avg_iters) / cap_rate
        alu = {"bound": "ALU pipe issue slots (int32 LOP3/SHF/VIMNMX/VIADD/PRMT)", "achieved": frac * alu_peak,
               "peak": alu_peak, "unit": "T thread-instr/s", "frac": frac,
               "peak_source": "148 SMs x 64 ALU lanes x median SM clock under load (profiles/microbench/pipes_r01_b200.txt)",
               "ncu_capture": {k: cap[k] for k in ("alu_pipe_pct", "fma_pipe_pct", "lsu_pipe_pct", "issue_per_cycle",
                                                   "warp_inst_per_frame_iter", "alu_warp_inst_per_frame_iter", "source") if k in cap},
               "algorithmic_int_ops_per_frame_iter": ops_iter,
               "algorithmic_int_ops_per_s": per_gpu_fps * avg_iters * ops_iter}

    line = {
        "metric": "decoded info Gbit/s @30 iters", "value": value, "unit": "Gbit/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16x2 (int32 re-decode of guarded frames)",
        "data": "synthetic",
        "config": {"workload": name, "code": args.code, "ebn0_db": ebn0, "frames_per_step_per_gpu": frames,
                   "avg_iters": avg_iters, "info_bits": k, "early_termination": "reference rule, never met at this Eb/N0",
                   "l2_policy": "inputs larger than L2 (%.0f MB int16 LLRs per step)" % (frames * code.n * 2 / 1e6),
                   "fallback_frames": fallback, "precision": args.precision},
        "frames_per_s": fps,
        "operating_point": {"ebn0_db": ebn0_op, "avg_iters": it_op, "value": world * frames / (ms_op * 1e-3) * k / 1e9,
                            "unit": "Gbit/s", "frames_per_s": world * frames / (ms_op * 1e-3)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": algo_bytes * frames,
                     "peak_source": peak_src, "algorithmic_bytes_per_frame": algo_bytes,
                     "note": "state is smem-resident; HBM is touched once per frame; the binding resource is the ALU pipe, see roofline_int"},
        "roofline_int": alu,
        "e2e": {"value": e2e_value, "unit": "Gbit/s", "h2d_bytes_per_step": e2e_frames * code.n * 4,
                "d2h_bytes_per_step": e2e_frames * (4 + code.nw32 * 4), "frames_per_step": e2e_frames,
                "api": "ldpc_decode_batch (host int32 LLR in, iters + packed bits out)"},
        "gpu_launches": launches, "clocks": clk,
    }
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = min(frames, (args.cpu_baseline_frames_per_core or CPU_SAMPLE_PER_CORE[args.code]) * cores)
        llr_host = llr16[:sample].to(torch.int32).cpu().numpy()
        wall, cpu_iters, kind = cpu_decode(args.code, code, llr_host, cores)
        step_device(llr16)
        torch.cuda.synchronize()
        gpu_iters = iters[:sample].cpu().numpy()
        line["cpu_baseline"] = {"value": sample * k / wall / 1e9, "unit": "Gbit/s", "cores": cores, "kind": kind,
                                "cpu": cpu_model(),
                                "sample": "first %d frames of the step's batch, one process per core" % sample,
                                "frames_per_s": sample / wall,
                                "iteration_count_mismatches_vs_gpu": int((gpu_iters != cpu_iters).sum())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--code", default="wifi", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per step per GPU (default per workload)")
    ap.add_argument("--precision", type=int, default=0, choices=[0, 16, 32])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--frames-per-cta", type=int, default=0)
    ap.add_argument("--e2e-frames", type=int, default=1 << 16)
    ap.add_argument("--cpu-frames-per-core", type=int, default=256, help="--impl reference: frames per core and step")
    ap.add_argument("--cpu-baseline-frames-per-core", type=int, default=0,
                    help="GPU arm: frames per core of the cpu_baseline leg (default: about 10 s of CPU work)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least three warm-up steps
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
