#!/usr/bin/env python3
"""Full-state parity at the scale SURVEY.md 8(c) asks for: per code and Eb/N0 point, F frames decoded by the CUDA
engine (through the C ABI, parity-mode outputs on) and by the CPU oracle on all host cores; iteration counts,
decoded bits, posteriors and the final variable-to-check messages (EdgeRAM order) are compared frame by frame
through CRC-32 digests of every array.  Frames that hit decode_fixpoint's pre-check (iters == 0) keep stale
posteriors / EdgeRAM in the reference (quirk Q6) and are compared on iterations and bits only.

    python scripts/parity_at_scale.py [--frames 100000] [--codes wifi a5 c79 a24] > profiles/r02/parity_at_scale.txt   (also run by tests/test_gpu_parity_at_scale.py)
"""
import argparse
import multiprocessing as mp
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# operating point, a point where every frame runs MAX_ITER, and a high-SNR point (pre-check hits, large channel values)
POINTS = {"wifi": (2.0, 0.5, 9.0), "a5": (4.5, 2.0, 8.0), "c79": (4.5, 2.0, 8.0), "a24": (6.0, 3.0, 9.0)}
PRECHECK = {"wifi": False, "a5": True, "c79": False, "a24": True}


def digest(iters, bits, post, edge, mask):
    """[frames][4] uint32: iterations, crc(bits), crc(post), crc(valid EdgeRAM words); the last two 0 when iters == 0."""
    out = np.zeros((len(iters), 4), np.uint32)
    for f in range(len(iters)):
        out[f, 0] = np.uint32(int(iters[f]) & 0xffffffff)
        out[f, 1] = zlib.crc32(np.ascontiguousarray(bits[f]).tobytes())
        if int(iters[f]) > 0:
            out[f, 2] = zlib.crc32(np.ascontiguousarray(post[f], np.int32).tobytes())
            out[f, 3] = zlib.crc32(np.ascontiguousarray(edge[f][mask], np.int32).tobytes())
    return out


def cpu_worker(args):
    tables, llr, precheck = args
    from oracle import pyoracle as po
    t = po.Tables(*tables)
    orc = po.Oracle(t)
    mask = np.arange(t.dc_max)[:, None] < t.cdeg[None, :]
    out = np.zeros((len(llr), 4), np.uint32)
    for f, x in enumerate(llr):
        it, bits, post, edge = orc.decode(x, precheck=precheck)
        packed = np.packbits(bits.astype(np.uint8), bitorder="little")
        pad = (-len(packed)) % 4
        packed = np.concatenate([packed, np.zeros(pad, np.uint8)]).view(np.uint32)
        out[f:f + 1] = digest([it], [packed], [post], [edge], mask)
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=100000)
    ap.add_argument("--codes", nargs="+", default=["wifi", "a5", "c79", "a24"])
    args = ap.parse_args(argv)
    import fixedpointldpc_b200 as fp
    from conftest import channel_frames
    from oracle import build_ref
    build_ref.build_oracle()
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    bad = 0
    with ctx.Pool(cores) as pool:
        for name in args.codes:
            code = fp.codes.NAMED[name]()
            rate = fp.codes.INFO_BITS[name] / code.n
            vdeg, cdeg, vlist, clist = code.tables()
            tables = (code.n, code.m, vdeg, cdeg, vlist, clist)
            mask = np.arange(code.dc_max)[:, None] < np.asarray(cdeg)[None, :]
            frames = args.frames if name != "a24" else max(1000, args.frames // 10)
            for snr_db in POINTS[name]:
                llr = channel_frames(code.n, rate, snr_db, frames, seed=int(snr_db * 10) + 1000)
                t0 = time.time()
                parts = np.array_split(llr, cores * 4)
                job = pool.map_async(cpu_worker, [(tables, p, PRECHECK[name]) for p in parts if len(p)])
                dec = fp.Decoder(code, precheck=PRECHECK[name])
                gpu = np.zeros((frames, 4), np.uint32)
                step = 2000 if name != "a24" else 500
                for s in range(0, frames, step):
                    out = dec.decode(llr[s:s + step], want_post=True, want_v2c=True)
                    gpu[s:s + step] = digest(out["iters"], out["bits"], out["post"], out["v2c"], mask)
                fallback = dec.stats()["fallback_frames"]
                dec.close()
                t_gpu = time.time() - t0
                cpu = np.concatenate(job.get())
                diff = (gpu != cpu)
                nbad = int(diff.any(axis=1).sum())
                bad += nbad
                it = gpu[:, 0].astype(np.int64)
                print("%-4s Eb/N0 %.1f dB  frames %6d  mismatching frames %d  (iters %d, bits %d, posteriors %d, messages %d)  "
                      "avg iterations %.2f, %d at 30, %d at 0 (pre-check), int32 re-decodes %d, %.0f s"
                      % (name, snr_db, frames, nbad, diff[:, 0].sum(), diff[:, 1].sum(), diff[:, 2].sum(), diff[:, 3].sum(),
                         it.mean(), (it == 30).sum(), (it == 0).sum(), fallback, time.time() - t0), flush=True)
    print("TOTAL mismatching frames: %d" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
