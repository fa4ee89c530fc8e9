#!/usr/bin/env python3
"""Host-to-device copy ceiling of the box with N processes copying at once (what bounds the end-to-end path at 8 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/h2d_ceiling.py

Every rank copies a pinned 512 MiB buffer to its GPU 20 times (cudaMemcpyAsync on one stream, like the host pipeline of
ldpc_decode_batch), all ranks at once; rank 0 prints per-rank and aggregate GB/s, alone (1 rank active) and together."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes, reps = 512 << 20, 20
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(rank + 1)
    devbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def timed(active):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if active:
            for _ in range(reps):
                devbuf.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gbs = torch.tensor([nbytes * reps / dt / 1e9 if active else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            parts = [torch.zeros_like(gbs) for _ in range(world)]
            dist.all_gather(parts, gbs)
            return [float(p.item()) for p in parts]
        return [float(gbs.item())]

    timed(True)
    together = timed(True)
    alone = timed(rank == 0)
    if rank == 0:
        print(json.dumps({"ranks": world, "bytes_per_copy": nbytes, "copies": reps,
                          "alone_gbs": alone[0], "together_gbs_per_rank": together, "together_gbs_total": sum(together)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
