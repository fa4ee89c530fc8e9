#!/usr/bin/env python3
"""One launch each of the kernels beside the decode-from-memory path, for ncu:
encode_kernel (65 536 random messages of the array p47 r5 code), channel_kernel (4 096 frames of the reference noise
stream), and one Monte-Carlo launch of the decode kernel (2^17 frames at 5 dB, channel generated in the refill step).

    ncu --set full --import-source on --clock-control none -k regex:"encode_kernel|channel_kernel|decode_kernel" \
        -o gpurun_out/prof_aux python scripts/ncu_aux.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fixedpointldpc_b200 as fp  # noqa: E402


def main():
    code = fp.codes.array_p47_r5()
    gen = fp.Generator(code=code)
    rng = np.random.default_rng(1)
    info = rng.integers(0, 256, (65536, (gen.k + 7) // 8), dtype=np.uint8)
    t0 = time.time()
    cw = gen.encode_batch(info)
    t_enc = time.time() - t0
    dec = fp.Decoder(code, precheck=True)
    snr = 2 * 10 ** (5.0 / 10) * code.rate
    llr = dec.mc_channel(4096, snr, stream=fp.STREAM_REFERENCE, seed=123456789)
    t0 = time.time()
    out = dec.mc_run(1 << 17, snr, stream=fp.STREAM_PHILOX, seed=7, want_frame_err=False)
    t_mc = time.time() - t0
    print("encode_batch %.3f s (65536 messages, first word %08x), channel %s, mc_run %.3f s: %s" %
          (t_enc, int(cw[0, 0]), llr.shape, t_mc, {k: out[k] for k in ("frames", "frame_errors", "bit_errors", "iter_sum")}))


if __name__ == "__main__":
    main()
