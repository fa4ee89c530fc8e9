"""Host-side Monte-Carlo logic on CPU: sharding, the sequential stopping rule, and the world_size-2 path
over gloo (the GPU box runs the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest


def _fake_errors(first_frame, frames):
    """Deterministic per-frame error counts as a function of the global frame index."""
    g = np.arange(first_frame, first_frame + frames, dtype=np.int64)
    h = (g * 2654435761 + 12345) % 1009
    return np.where(h < 9, (h % 7) + 1, 0).astype(np.uint16)


def _simulate(first_frame, frames):
    e = _fake_errors(first_frame, frames)
    return e, int(3 * frames)


def test_sequential_stop():
    from fixedpointldpc_b200.montecarlo import sequential_stop
    fe = [0, 2, 0, 0, 5, 1, 0, 3]
    assert sequential_stop(fe, 3) == (8, 3, 6, True)
    assert sequential_stop(fe, 4) == (11, 4, 8, True)
    assert sequential_stop(fe, 5) == (11, 4, 8, False)
    assert sequential_stop(fe, 0, max_frames=5) == (7, 2, 5, False)
    assert sequential_stop([], 1) == (0, 0, 0, False)


def test_single_process_equals_sequential_loop():
    from fixedpointldpc_b200.montecarlo import Shards, run_point
    want_b = want_f = want_n = 0
    g = 0
    while want_f < 100:
        e = int(_fake_errors(g, 1)[0])
        want_n += 1
        if e:
            want_f += 1
            want_b += e
        g += 1
    for batch in (64, 1000, 4096):
        res = run_point(_simulate, Shards(1, 0, batch), 100)
        assert (res["bit_errors"], res["frame_errors"], res["frames"]) == (want_b, want_f, want_n)
    res = run_point(_simulate, Shards(1, 0, 500), 0, max_frames=1234)
    assert res["frames"] == 1234 and res["bit_errors"] == int(_fake_errors(0, 1234).sum())


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from fixedpointldpc_b200.montecarlo import Shards, run_point
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = run_point(_simulate, Shards(world, rank, batch), 100, dist=dist)
    q.put((rank, res))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [257, 2048])
def test_two_ranks_over_gloo_match_single_process(batch):
    import torch.multiprocessing as mp
    from fixedpointldpc_b200.montecarlo import Shards, run_point
    single = run_point(_simulate, Shards(1, 0, 512), 100)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert (got[r]["bit_errors"], got[r]["frame_errors"], got[r]["frames"]) == \
               (single["bit_errors"], single["frame_errors"], single["frames"])


def test_reference_print_format():
    from fixedpointldpc_b200.montecarlo import reference_print
    txt = reference_print({"bit_errors": 2732, "frame_errors": 100, "frames": 393214}, 1944)
    assert txt == "2732 100 393214\n FER: 0.000254314 BER: 3.57401e-06"   # wifi_results_4_4_2dB_30iter.txt:3-4
