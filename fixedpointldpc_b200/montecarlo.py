"""Monte-Carlo BER/FER driver: the frame loop of ArrayLDPC_Debug_Wifi / ArrayLDPC_Debug / ArrayLDPC_PerfTest
(PerfTest.cpp:97-137, 276-313, 491-514) sharded over the GPUs of one box.

Frames are independent and identified by a global index g (the noise of frame g depends on (seed, g) only), so
rank r of R simulates the blocks  g in [(round*R + r)*B, (round*R + r + 1)*B)  and the only communication is
  * one all-reduce(sum) of four counters per round (frames, frame errors, bit errors, iteration sum), and
  * one all-gather of the last round's per-frame error counts, to cut the run at exactly the frame on which the
    reference's sequential `while(pckerror < 100)` would have stopped.
The result therefore does not depend on R or B, and with the reference noise stream it equals the reference's
own printout.  torch.distributed provides the plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np


class Shards:
    """Block-cyclic partition of the global frame index space."""

    def __init__(self, world, rank, batch):
        assert 0 <= rank < world and batch > 0
        self.world, self.rank, self.batch = world, rank, batch

    def first_frame(self, round_index, rank=None):
        rank = self.rank if rank is None else rank
        return (round_index * self.world + rank) * self.batch

    def frames_per_round(self):
        return self.world * self.batch


def sequential_stop(frame_err, target_frame_errors, max_frames=None):
    """Apply the reference's loop `while(pckerror < target)` to per-frame error counts given in frame order.
    Returns (bit_errors, frame_errors, frames, reached)."""
    fe = np.asarray(frame_err).astype(np.int64)
    if max_frames is not None:
        fe = fe[:max_frames]
    hits = np.flatnonzero(fe > 0)
    if target_frame_errors and len(hits) >= target_frame_errors:
        stop = hits[target_frame_errors - 1] + 1
        return int(fe[:stop].sum()), int(target_frame_errors), int(stop), True
    return int(fe.sum()), int(len(hits)), int(len(fe)), False


def run_point(simulate, shards, target_frame_errors=100, max_frames=None, dist=None, device="cpu"):
    """One Eb/N0 point.

    simulate(first_frame, frames) -> (frame_err uint16[frames], iter_sum) for this rank's block.
    dist: torch.distributed (initialised) or None for a single process.
    Returns dict(bit_errors, frame_errors, frames, iter_sum, rounds, reached)."""
    import torch
    world = shards.world
    totals = np.zeros(3, np.int64)  # bit errors, frame errors, frames -- complete rounds only
    iter_sum = 0
    rnd = 0
    while True:
        ferr, its = simulate(shards.first_frame(rnd), shards.batch)
        ferr = np.ascontiguousarray(ferr, np.uint16)
        local = torch.tensor([int(ferr.astype(np.int64).sum()), int((ferr > 0).sum()), len(ferr), int(its)],
                             dtype=torch.int64, device=device)
        if dist is not None and world > 1:
            dist.all_reduce(local)  # the tiny counter all-reduce (NCCL over NVLink on the GPU box)
        rb, rf, rn, ri = (int(x) for x in local.tolist())
        need = target_frame_errors - totals[1] if target_frame_errors else None
        frames_left = None if max_frames is None else max_frames - totals[2]
        last = (need is not None and rf >= need) or (frames_left is not None and rn >= frames_left)
        if not last:
            totals += (rb, rf, rn)
            iter_sum += ri
            rnd += 1
            continue
        # final round: gather the per-frame counts and stop on the exact frame
        mine = torch.from_numpy(ferr.astype(np.int32)).to(device)
        if dist is not None and world > 1:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            allerr = torch.cat(parts).cpu().numpy()
        else:
            allerr = mine.cpu().numpy()
        b, f, n, reached = sequential_stop(allerr, need, frames_left)
        totals += (b, f, n)
        iter_sum += ri  # iterations of the whole last round (the reference does not print them)
        return {"bit_errors": int(totals[0]), "frame_errors": int(totals[1]), "frames": int(totals[2]),
                "iter_sum": int(iter_sum), "rounds": rnd + 1, "reached": bool(reached)}


def reference_print(res, n):
    """The two lines every driver prints (PerfTest.cpp:136-137): note BER divides by CWD_LENGTH (quirk Q10)."""
    return "%g %g %d\n FER: %g BER: %g" % (res["bit_errors"], res["frame_errors"], res["frames"],
                                             res["frame_errors"] / res["frames"],
                                             res["bit_errors"] / res["frames"] / n)


def gpu_simulator(dec, snr, **mc_kwargs):
    def simulate(first_frame, frames):
        out = dec.mc_run(frames, snr, first_frame=first_frame, **mc_kwargs)
        return out["frame_err"], out["iter_sum"]
    return simulate


def main(argv=None):
    import torch
    import torch.distributed as dist
    import fixedpointldpc_b200 as fp

    ap = argparse.ArgumentParser(description="BER/FER point(s) on all GPUs of the box (launch with torchrun for >1 GPU)")
    ap.add_argument("--code", default="a5", choices=sorted(fp.codes.NAMED))
    ap.add_argument("--ebn0", type=float, nargs="+", default=[4.5])
    ap.add_argument("--rate", type=float, default=None, help="rate in snr = 2*10^(dB/10)*R (default: the drivers' choice)")
    ap.add_argument("--frame-errors", type=int, default=100)
    ap.add_argument("--max-frames", type=int, default=None)
    ap.add_argument("--batch", type=int, default=1 << 17, help="frames per GPU per round")
    ap.add_argument("--stream", default="philox", choices=["philox", "reference"])
    ap.add_argument("--seed", type=int, default=123456789)
    ap.add_argument("--precheck", type=int, default=None)
    args = ap.parse_args(argv)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    code = fp.codes.NAMED[args.code]()
    precheck = (args.code in ("a5", "a24")) if args.precheck is None else bool(args.precheck)
    dec = fp.Decoder(code, precheck=precheck, device=local)
    rate = args.rate if args.rate is not None else (0.5 if args.code == "wifi" else
                                                    (code.rate if args.code in ("a5", "a24") else fp.codes.INFO_BITS[args.code] / code.n))
    stream = fp.STREAM_REFERENCE if args.stream == "reference" else fp.STREAM_PHILOX
    shards = Shards(world, rank, args.batch)
    for db in args.ebn0:
        snr = 2 * 10 ** (db / 10) * rate
        t0 = time.time()
        res = run_point(gpu_simulator(dec, snr, stream=stream, seed=args.seed), shards, args.frame_errors,
                        args.max_frames, dist if world > 1 else None, dev)
        if rank == 0:
            res.update(ebn0_db=db, seconds=time.time() - t0, gpus=world, code=args.code,
                       fer=res["frame_errors"] / res["frames"], ber_ref=res["bit_errors"] / res["frames"] / code.n)
            print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
