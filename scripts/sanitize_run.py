#!/usr/bin/env python3
"""Workload for compute-sanitizer (racecheck / memcheck / synccheck) on the decode kernel:

    compute-sanitizer --tool racecheck python scripts/sanitize_run.py

Covers the packed and the int32 kernel, a long queue (slots claim one frame ahead) and a short one, the
pre-check, parity-mode outputs and Monte-Carlo mode, on the 802.11 code (irregular, balanced work order) and the
array code (regular)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fixedpointldpc_b200 as fp  # noqa: E402
from conftest import channel_frames  # noqa: E402

SCALE = float(os.environ.get("SANITIZE_SCALE", "1"))  # shrink the frame counts (the tools slow the kernels 20 - 100 x)
os.environ.setdefault("LDPC_NO_FEED", "1")  # the fed launch waits for copies queued after it: not under a tool that serialises

for name, snr, precheck, frames in (("wifi", 2.0, False, int(12000 * SCALE)), ("a5", 4.5, True, int(6000 * SCALE))):
    code = fp.codes.NAMED[name]()
    rate = fp.codes.INFO_BITS[name] / code.n
    llr = channel_frames(code.n, rate, snr, frames, seed=5)
    for precision in (16, 32):
        dec = fp.Decoder(code, precision=precision, precheck=precheck)
        out = dec.decode(llr if precision == 16 else llr[:max(64, int(1500 * SCALE))])
        small = dec.decode(llr[:40], want_post=True, want_v2c=True)
        assert (small["iters"] == out["iters"][:40]).all()
        print(name, precision, "iterations", np.bincount(out["iters"])[:8], flush=True)
        dec.close()
    dec = fp.Decoder(code, precheck=precheck)
    snr_lin = 2 * 10 ** (snr / 10) * rate
    res = dec.mc_run(3000, snr_lin, stream=fp.STREAM_PHILOX, seed=3)
    print(name, "mc", {k: res[k] for k in ("frames", "frame_errors", "bit_errors", "iter_sum")}, flush=True)
    res = dec.mc_run(300, snr_lin, stream=fp.STREAM_REFERENCE, seed=123456789, pin_index=np.arange(0, 200, 3), pin_value=112)
    print(name, "mc reference stream + pins", {k: res[k] for k in ("frames", "frame_errors", "bit_errors", "iter_sum")}, flush=True)
    dec.close()
print("done")
