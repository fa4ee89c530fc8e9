#!/usr/bin/env python3
"""Headline benchmark: decoded info Gbit/s per B200 @30 iterations (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--code wifi|a5|c79|a24] [--only] [--impl reference]

A "step" is one pass of the decode path over one batch of synthetic frames.  The headline workload is
BASELINE.json configs[1]: the 802.11n n=1944 rate-1/2 code (H_802.11_IndZero), FRAC_WIDTH=4 quantised LLRs,
MAX_ITER=30, at an Eb/N0 (0 dB) where every frame runs all 30 iterations under the reference's own
early-termination rule, so `value` is the "@30 iters" number with nothing skipped; the 2 dB
operating point with early termination is reported beside it as `operating_point`.  The other three named codes
(array p47 r5, cut79, array p47 r24) are measured the same way in the same run and reported under `codes`
(`--only` skips them), each with its own roofline, end-to-end and CPU-reference figures.

One process per GPU (torchrun for N > 1): frames are independent, so every rank decodes its own shard (weak
scaling); NCCL all-reduces the iteration / frame counters after the timed region, and the Monte-Carlo leg (`mc`)
has its counter all-reduce inside the timed region.

`--impl reference` times the reference's own CPU decoder (oracle/_ref/libref_<code>.so, the unmodified sources
of /root/reference compiled by oracle/build_ref.py) on all host cores, one process per core (the reference is not
re-entrant), on a bounded sample of the same workload.  It does not import the product package.  If that binary
is absent the plain-C port (oracle/ldpc_oracle.c) is timed instead.
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

METRIC = "decoded info Gbit/s @30 iters"
WORKLOADS = {
    # code: (name, Eb/N0 where all frames run 30 iterations, operating Eb/N0, frames per step and GPU)
    "wifi": ("802.11n n=1944 R=1/2 (H_802.11_IndZero), FRAC_WIDTH=4, MAX_ITER=30", 0.0, 2.0, 1 << 17),
    "a5": ("array p=47 r=5 n=2209 (H_array_p47_r5_forward), FRAC_WIDTH=4, MAX_ITER=30", 2.0, 4.5, 1 << 17),
    "c79": ("array cut79 n=2212 (H2212_316_array_cut79), FRAC_WIDTH=4, MAX_ITER=30", 2.0, 4.5, 1 << 17),
    "a24": ("array p=47 r=24 n=2209 (H_array_p47_r24_forward), FRAC_WIDTH=4, MAX_ITER=30", 3.0, 6.0, 1 << 15),
}
PRECHECK = {"wifi": False, "a5": True, "c79": False, "a24": True}  # decode_general_fp vs decode_fixpoint
# frames per host core for the cpu_baseline leg of the GPU arm (headline: about 10 s of the reference decoder at
# 30 iterations; the other codes about 5 s each)
CPU_SAMPLE_PER_CORE = {"wifi": 4096, "a5": 1536, "c79": 1536, "a24": 192}
N_OF = {"wifi": 1944, "a5": 2209, "c79": 2212, "a24": 2209}


def workload_config(code_name):
    """The `config` object: the workload only, identical in both arms (what a run did beyond that is in `run`)."""
    from oracle import named_codes as nc
    name, ebn0, _, frames = WORKLOADS[code_name]
    return {"workload": name, "code": code_name, "ebn0_db": ebn0, "max_iter": 30, "info_bits": nc.INFO_BITS[code_name],
            "frames_per_step_per_gpu": frames,
            "early_termination": "reference rule (syndrome after every iteration), never met at this Eb/N0",
            "l2_policy": "inputs larger than L2 (%.0f MB of int16 LLRs per step)" % (frames * N_OF[code_name] * 2 / 1e6)}


def make_frames(torch, n, rate, ebn0_db, frames, seed, device, dtype):
    """All-zero codeword over BPSK/AWGN, quantised like DecodeTrial (PerfTest.cpp:159-170):
    LLR_fp = int(2*snr*(1 + N(0, sigma)) * 2^4), snr = 2*10^(dB/10)*R, sigma = sqrt(1/snr)."""
    snr = 2.0 * 10.0 ** (ebn0_db / 10.0) * rate
    sigma = math.sqrt(1.0 / snr)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((frames, n), dtype=dtype, device=device)
    chunk = 1 << 14
    for s in range(0, frames, chunk):
        e = min(frames, s + chunk)
        z = torch.randn((e - s, n), generator=g, device=device, dtype=torch.float64)
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
# CPU reference (oracle/_ref, else the port): the checker side only -- nothing here touches the product
# ------------------------------------------------------------------------------------------------
_WORKER_CACHE = {}


def _cpu_worker(args):
    code_name, llr, fixpoint, use_ref = args
    from oracle import named_codes as nc
    from oracle import pyoracle as po
    key = (code_name, use_ref)
    if key not in _WORKER_CACHE:  # built by the warm-up call, outside the timed map
        tables = nc.tables(code_name)
        if use_ref:
            dec = po.Reference(code_name)
            dec.set_tables(tables)
        else:
            dec = po.Oracle(tables)
        _WORKER_CACHE[key] = dec
    dec = _WORKER_CACHE[key]
    t0 = time.perf_counter()
    iters = dec.decode_many(llr, fixpoint) if use_ref else dec.decode_many(llr, precheck=fixpoint)
    return time.perf_counter() - t0, iters


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class CpuPool:
    """One process per host core (the reference keeps its scratch in function statics, ArrayLDPC_Decoder.cpp:21-37)."""

    def __init__(self, cores):
        import multiprocessing as mp
        from oracle import build_ref
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores)
        self._built = build_ref.build_oracle

    def decode(self, code_name, llr_host):
        """Decode llr_host [frames][n] int32; returns (seconds, iters, kind)."""
        from oracle import pyoracle as po
        use_ref = po.reference_available(code_name)
        if not use_ref:
            self._built()
        parts = [np.ascontiguousarray(p) for p in np.array_split(llr_host, self.cores) if len(p)]
        fix = PRECHECK[code_name]
        self.pool.map(_cpu_worker, [(code_name, p[:1], fix, use_ref) for p in parts])  # page in, warm up
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, [(code_name, p, fix, use_ref) for p in parts])
        wall = time.perf_counter() - t0
        return wall, np.concatenate([r[1] for r in res]), ("reference" if use_ref else "port")

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import named_codes as nc
    name, ebn0, _, _ = WORKLOADS[args.code]
    k = nc.INFO_BITS[args.code]
    cores = os.cpu_count() or 1
    per_core = args.cpu_frames_per_core
    frames = per_core * cores
    llr = make_frames(torch, N_OF[args.code], nc.channel_rate(args.code), ebn0, frames, 20261018, "cpu", torch.int32).numpy()
    pool = CpuPool(cores)
    times, total_it, kind = [], 0, "reference"
    for step in range(args.warmup + args.steps):
        wall, iters, kind = pool.decode(args.code, llr)
        if step >= args.warmup:
            times.append(wall)
            total_it += int(iters.sum())
    pool.close()
    t = sum(times)
    value = frames * args.steps * k / t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gbit/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args.code),
        "avg_iters": total_it / (frames * args.steps),
        "cpu_baseline": {"value": value, "unit": "Gbit/s", "cores": cores, "kind": kind, "cpu": cpu_model(),
                         "sample": "%d frames per step (%d per core) of the workload's channel, one process per core" % (frames, per_core),
                         "frames_per_s": frames * args.steps / t},
        "e2e": {"value": value, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def load_captures():
    """Per-code figures of the committed ncu captures (profiles/<round>/traffic.json, written by
    scripts/ncu_summarise.py); later rounds override earlier ones."""
    caps = {}
    for rnd in ("r01", "r02"):
        try:
            with open(os.path.join(ROOT, "profiles", rnd, "traffic.json")) as fh:
                for k, v in json.load(fh).items():
                    if isinstance(v, dict):
                        caps[k] = v
        except Exception:
            pass
    return caps


def roofline_int(cap, per_gpu_fps, avg_iters, sm_mhz, ops_iter):
    """The bound resource is instruction issue: the kernel is a stream of dependent integer instructions on
    shared-memory-resident state, and the ALU pipe (LOP3 / SHF / PRMT / VIADD.16x2 at 64 lanes per clock and SM)
    carries most of them.  ncu measured both utilisations, over the elapsed cycles of one launch of this kernel
    (`sm__inst_executed...pct_of_peak_sustained_elapsed`, `sm__pipe_alu_cycles_active...elapsed`); the live figure is
    that percentage scaled by (frame-iterations/s now) / (frame-iterations/s of the captured launch) -- same kernel,
    same instructions per frame-iteration."""
    if not cap or not cap.get("kernel_ms") or "issue_elapsed_pct" not in cap:
        return None
    cap_rate = cap["frames_in_capture"] * cap["avg_iters"] / (cap["kernel_ms"] * 1e-3)
    scale = (per_gpu_fps * avg_iters) / cap_rate
    frac = cap["issue_elapsed_pct"] / 100.0 * scale
    peak = 148 * 4 * sm_mhz * 1e6 / 1e12  # warp-instructions issued per second: 4 schedulers per SM, one per clock
    return {"bound": "issue slots (warp instructions issued per cycle and scheduler)", "achieved": frac * peak, "peak": peak,
            "unit": "T warp-instr/s", "frac": frac,
            "alu_pipe_frac": cap["alu_elapsed_pct"] / 100.0 * scale,
            "rate_scale_vs_capture": scale,
            "peak_source": "148 SMs x 4 schedulers x median SM clock under load",
            "ncu_capture": {k: cap[k] for k in ("issue_elapsed_pct", "alu_elapsed_pct", "fma_pipe_pct", "lsu_pipe_pct",
                                                "warp_inst_per_frame_iter", "kernel_ms", "frames_in_capture", "avg_iters",
                                                "source") if k in cap},
            "algorithmic_int_ops_per_frame_iter": ops_iter,
            "algorithmic_int_ops_per_s": per_gpu_fps * avg_iters * ops_iter}


def bench_code(ctx, code_name, steps, warmup, headline):
    """All figures of one code on this rank's GPU; collective calls inside are made by every rank."""
    torch, dist, fp, args = ctx["torch"], ctx["dist"], ctx["fp"], ctx["args"]
    world, rank, local, dev, stream = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"], ctx["stream"]
    name, ebn0, ebn0_op, frames = WORKLOADS[code_name]
    if args.frames:
        frames = args.frames
    code = fp.codes.NAMED[code_name]()
    k = fp.codes.INFO_BITS[code_name]
    rate = ctx["nc"].channel_rate(code_name)
    dec = fp.Decoder(code, max_iter=30, precheck=PRECHECK[code_name], device=local, precision=args.precision,
                     threads=args.threads, frames_per_cta=args.frames_per_cta)

    # inputs resident in HBM; one distinct batch per rank (weak scaling)
    llr16 = make_frames(torch, code.n, rate, ebn0, frames, 20261018 + rank, dev, torch.int16)
    iters = torch.zeros(frames, dtype=torch.int32, device=dev)
    bits = torch.zeros((frames, code.nw32), dtype=torch.int32, device=dev)

    def step_device(llr):
        dec.decode_device(llr.data_ptr(), 16, frames, iters.data_ptr(), bits.data_ptr(), None, None, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps, nwarm):
        for _ in range(nwarm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = dec.stats()["kernel_launches"]
        e0.record(stream)
        for _ in range(nsteps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), dec.stats()["kernel_launches"] - l0

    with ClockSampler(local) as clocks:
        ms_total, launches = timed(lambda: step_device(llr16), steps, warmup)
    clk = clocks.summary()
    counters = torch.stack([iters.clamp(min=0).sum().to(torch.int64), torch.tensor(frames, device=dev)])
    if world > 1:
        dist.all_reduce(counters)  # after the timed region: iteration / frame counters of all ranks
    total_frames = int(counters[1].item())
    avg_iters = float(counters[0].item()) / total_frames
    fallback = dec.stats()["fallback_frames"]
    ms_step = ms_total / steps
    fps = total_frames / (ms_step * 1e-3)
    value = fps * k / 1e9

    # operating point (early termination active), same engine, per rank
    llr_op = make_frames(torch, code.n, rate, ebn0_op, frames, 777 + rank, dev, torch.int16)
    op_steps = max(2, steps // 2)
    ms_op, _ = timed(lambda: step_device(llr_op), op_steps, 1)
    ms_op /= op_steps
    it_op = float(iters.clamp(min=0).sum().item()) / frames
    del llr_op

    # end to end through the reference-facing C-ABI calls on pinned HOST buffers: the int32 layout of the reference's
    # `const int *LLR` (ldpc_decode_batch) and the int16 entry (ldpc_decode_batch_i16); iterations + packed bits out
    e2e_frames = min(frames, args.e2e_frames)
    h_iters = torch.empty(e2e_frames, dtype=torch.int32).pin_memory()
    h_bits = torch.empty((e2e_frames, code.nw32), dtype=torch.int32).pin_memory()
    e2e_steps = max(2, steps // 2)
    e2e = {}
    for llr_bits, dt in ((32, torch.int32), (16, torch.int16)):
        h_llr = torch.empty((e2e_frames, code.n), dtype=dt).pin_memory()
        h_llr.copy_(llr16[:e2e_frames].to(dt).cpu())

        def step_host():
            dec.decode_raw(h_llr.data_ptr(), e2e_frames, h_iters.data_ptr(), h_bits.data_ptr(), llr_bits=llr_bits)

        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        torch.cuda.synchronize()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        v = world * e2e_frames * e2e_steps * k / float(t_e2e.item()) / 1e9
        e2e[llr_bits] = {"value": v, "unit": "Gbit/s", "h2d_bytes_per_step": e2e_frames * code.n * llr_bits // 8,
                         "d2h_bytes_per_step": e2e_frames * (4 + code.nw32 * 4), "frames_per_step": e2e_frames,
                         "steps": e2e_steps, "frac_of_device_rate": v / value,
                         "api": "ldpc_decode_batch (host int32 LLR in, iters + packed bits out)" if llr_bits == 32 else
                                "ldpc_decode_batch_i16 (host int16 LLR in, iters + packed bits out)"}
        del h_llr

    peaks = ctx["peaks"]
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    algo_bytes = 2 * code.n + (code.n + 7) // 8 + 4  # int16 LLR in + packed bits + iteration count out
    per_gpu_fps = fps / world
    achieved = per_gpu_fps * algo_bytes / 1e9
    cap = ctx["captures"].get(code_name)
    traffic = cap["dram_bytes_per_frame"] * frames if cap and "dram_bytes_per_frame" in cap else None
    # algorithmic integer work: 18 scalar ops per necessary sxor, 3 per edge, 1 per variable
    _, cdeg, _, _ = code.tables()
    ops_iter = 18 * int((3 * cdeg - 6).sum()) + 3 * code.edges + code.n
    sm_mhz = clk["sm_mhz"] or float(peaks.get("sm_max_mhz", 1965.0))

    out = {
        "value": value, "unit": "Gbit/s", "frames_per_s": fps, "ms_per_step": ms_step, "steps": steps, "warmup": warmup,
        "avg_iters": avg_iters, "config": workload_config(code_name),
        "run": {"frames_per_step_per_gpu": frames, "fallback_frames": fallback, "precision": args.precision,
                "threads": dec.stats()["threads"], "frames_per_cta": dec.stats()["frames_per_cta"], "grid": dec.stats()["grid"],
                "resident_ctas_per_sm": dec.stats()["resident_ctas_per_sm"],
                "launch_smem_bytes": dec.stats()["launch_smem_bytes"]},
        "operating_point": {"ebn0_db": ebn0_op, "avg_iters": it_op, "value": world * frames / (ms_op * 1e-3) * k / 1e9,
                            "unit": "Gbit/s", "frames_per_s": world * frames / (ms_op * 1e-3),
                            "frame_iterations_per_s": world * frames / (ms_op * 1e-3) * it_op,
                            "frac_of_30it_frame_iteration_rate": (frames / (ms_op * 1e-3) * it_op) / (per_gpu_fps * avg_iters)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": algo_bytes * frames,
                     "peak_source": peak_src, "algorithmic_bytes_per_frame": algo_bytes,
                     "note": "state is smem-resident; HBM is touched once per frame; the binding resource is instruction issue, see roofline_int"},
        "roofline_int": roofline_int(cap, per_gpu_fps, avg_iters, sm_mhz, ops_iter),
        "e2e": e2e[32], "e2e_i16": e2e[16],
        "gpu_launches": launches, "clocks": clk,
    }
    if headline:
        out["dtype"] = "int16x2 (int32 re-decode of guarded frames)"
    if world == 1 and not args.no_cpu and rank == 0:
        pool = ctx["cpu_pool"]()
        sample = min(frames, (args.cpu_baseline_frames_per_core or CPU_SAMPLE_PER_CORE[code_name]) * pool.cores)
        llr_host = llr16[:sample].to(torch.int32).cpu().numpy()
        wall, cpu_iters, kind = pool.decode(code_name, llr_host)
        step_device(llr16)
        torch.cuda.synchronize()
        gpu_iters = iters[:sample].cpu().numpy()
        out["cpu_baseline"] = {"value": sample * k / wall / 1e9, "unit": "Gbit/s", "cores": pool.cores, "kind": kind,
                               "cpu": cpu_model(),
                               "sample": "first %d frames of the step's batch, one process per core" % sample,
                               "frames_per_s": sample / wall,
                               "iteration_count_mismatches_vs_gpu": int((gpu_iters != cpu_iters).sum())}
    dec.close()
    del llr16, iters, bits
    torch.cuda.empty_cache()
    return out


def bench_mc(ctx, steps):
    """Monte-Carlo leg (BASELINE config 5): array p47 r5, channel generated in the kernel, only counters leave the GPU;
    every round ends with the all-reduce of the counters, inside the timed region."""
    torch, dist, fp = ctx["torch"], ctx["dist"], ctx["fp"]
    world, rank, local, dev, stream = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"], ctx["stream"]
    code = fp.codes.NAMED["a5"]()
    dec = fp.Decoder(code, max_iter=30, precheck=True, device=local)
    ebn0, batch = 5.0, 1 << 18
    snr = 2 * 10 ** (ebn0 / 10) * code.rate
    local_c = torch.zeros(4, dtype=torch.int64, device=dev)
    total_c = torch.zeros(4, dtype=torch.int64, device=dev)

    def one_round(r):
        local_c.zero_()
        dec.mc_run_device(batch, snr, local_c.data_ptr(), cuda_stream=stream.cuda_stream, stream=fp.STREAM_PHILOX,
                          seed=20261019, first_frame=(r * world + rank) * batch)
        if world > 1:
            dist.all_reduce(local_c)  # NCCL: frames, frame errors, bit errors, iteration sum of this round
        total_c.add_(local_c)

    for r in range(2):
        one_round(r)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_c.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for r in range(steps):
        one_round(2 + r)
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total = world * batch * steps
    tc = [int(x) for x in total_c.tolist()]
    dec.close()
    return {"code": "a5", "ebn0_db": ebn0, "counters": {"frames": tc[0], "frame_errors": tc[1], "bit_errors": tc[2], "iter_sum": tc[3]}, "stream": "philox", "frames_per_round_per_gpu": batch, "rounds": steps,
            "frames_per_s": total / (float(ms.item()) * 1e-3), "collective": "all-reduce of 4 int64 counters per round (NCCL), inside the timed region" if world > 1 else "none (one GPU)",
            "ms_per_round": float(ms.item()) / steps}


def bench_encoder(ctx, steps):
    """Encoder leg: FP_Encoder::encode batched on the GPU for the array p47 r5 generator derived from H.
    Device-timed on resident message words, and end to end through ldpc_encode_batch (host bytes in, packed codewords
    out); every codeword of one batch is checked against H on the host (numpy), a sample against the host encoder."""
    import numpy as np
    torch, fp = ctx["torch"], ctx["fp"]
    dev, stream, local = ctx["dev"], ctx["stream"], ctx["local"]
    code = fp.codes.NAMED["a5"]()
    gen = fp.Generator(code=code)
    frames, kw, nw = 1 << 18, (gen.k + 31) // 32, (gen.n + 31) // 32
    rng = np.random.default_rng(20261019)
    info = rng.integers(0, 256, (frames, (gen.k + 7) // 8), dtype=np.uint8)
    if gen.k % 8:
        info[:, -1] &= (1 << (gen.k % 8)) - 1
    words = np.zeros((frames, kw * 4), np.uint8)
    words[:, :info.shape[1]] = info
    d_info = torch.from_numpy(words.view(np.int32).reshape(frames, kw)).to(dev)
    d_cw = torch.zeros((frames, nw), dtype=torch.int32, device=dev)
    for _ in range(3):
        gen.encode_batch_device(d_info.data_ptr(), frames, d_cw.data_ptr(), device=local, cuda_stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        gen.encode_batch_device(d_info.data_ptr(), frames, d_cw.data_ptr(), device=local, cuda_stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # all codewords of the batch against H; a sample against the host encoder (FP_Encoder::encode on the CPU)
    cw = d_cw.cpu().numpy().view(np.uint32)
    checked = 1 << 14
    bits = np.unpackbits(cw[:checked].view(np.uint8), axis=1, bitorder="little")[:, :gen.n]
    _, cdeg, _, clist = code.tables()
    clist = np.asarray(clist).reshape(code.m, -1)
    syn = np.zeros((checked, code.m), np.uint8)
    for k in range(clist.shape[1]):
        syn ^= np.where(k < np.asarray(cdeg), bits[:, np.maximum(clist[:, k], 0)], 0).astype(np.uint8)
    host_mismatch = sum(int((gen.encode(info[f].tobytes()) != bits[f]).any()) for f in range(0, checked, checked // 64))
    t0 = time.perf_counter()
    out = gen.encode_batch(info[: 1 << 16], device=local)
    e2e_s = time.perf_counter() - t0
    bytes_per_frame = 4 * (kw + nw)
    return {"code": "a5", "api": "ldpc_encode_batch_device (resident message words -> packed codewords)", "frames_per_launch": frames,
            "ms_per_launch": ms, "frames_per_s": frames / (ms * 1e-3), "value": frames * gen.k / (ms * 1e-3) / 1e9, "unit": "info Gbit/s",
            "hbm_GBps": frames * bytes_per_frame / (ms * 1e-3) / 1e9, "algorithmic_bytes_per_frame": bytes_per_frame,
            "codewords_checked_against_H": checked, "codewords_failing_H": int(syn.any(axis=1).sum()),
            "host_encoder_mismatches_in_64": host_mismatch,
            "e2e": {"api": "ldpc_encode_batch (host bytes in, packed codewords out, pageable memory)", "frames": 1 << 16,
                    "value": (1 << 16) * gen.k / e2e_s / 1e9, "unit": "info Gbit/s", "identical_to_device_run": bool((out == cw[: 1 << 16]).all())}}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import fixedpointldpc_b200 as fp
    from oracle import named_codes as nc  # tables / rates of the workloads for the CPU checker leg; no compute

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
    # a dedicated non-default stream: the C ABI treats a NULL stream as "the decoder's own stream",
    # and torch.cuda.Event only sees work on the stream it is recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    pool_box = {}

    def cpu_pool():
        if "p" not in pool_box:
            pool_box["p"] = CpuPool(os.cpu_count() or 1)
        return pool_box["p"]

    ctx = {"torch": torch, "dist": dist, "fp": fp, "nc": nc, "args": args, "world": world, "rank": rank, "local": local,
           "dev": dev, "stream": stream, "peaks": peaks, "captures": load_captures(), "cpu_pool": cpu_pool}

    head = bench_code(ctx, args.code, args.steps, args.warmup, True)
    others = {}
    if not args.only:
        for c in ("a5", "c79", "a24", "wifi"):
            if c != args.code:
                others[c] = bench_code(ctx, c, max(3, args.steps // 4), 3, False)
        mc = bench_mc(ctx, max(3, args.steps // 4))
        enc = bench_encoder(ctx, max(3, args.steps // 4)) if rank == 0 else None
    if "p" in pool_box:
        pool_box["p"].close()
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": head.pop("dtype"), "data": "synthetic"}
        for key in ("config", "avg_iters", "run", "frames_per_s", "operating_point", "roofline", "roofline_int", "e2e", "e2e_i16",
                    "gpu_launches", "clocks", "cpu_baseline"):
            if key in head:
                line[key] = head[key]
        if not args.only:
            line["codes"] = others
            line["mc"] = mc
            line["encoder"] = enc
            line["gpu_launches"] = head["gpu_launches"]  # launches inside the headline's timed region
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
    ap.add_argument("--code", default="wifi", choices=sorted(WORKLOADS), help="headline code")
    ap.add_argument("--only", action="store_true", help="measure the headline code only (no `codes` / `mc` blocks)")
    ap.add_argument("--frames", type=int, default=0, help="frames per step per GPU (default per workload)")
    ap.add_argument("--precision", type=int, default=0, choices=[0, 16, 32])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--frames-per-cta", type=int, default=0)
    ap.add_argument("--e2e-frames", type=int, default=1 << 17, help="frames per end-to-end call (capped at the step's frames)")
    ap.add_argument("--cpu-frames-per-core", type=int, default=256, help="--impl reference: frames per core and step")
    ap.add_argument("--cpu-baseline-frames-per-core", type=int, default=0,
                    help="GPU arm: frames per core of the cpu_baseline leg (default: about 10 s of CPU work)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least three warm-up steps
    if args.frames or args.threads or args.frames_per_cta:
        args.only = True  # launch-shape experiments concern one code
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
