#!/usr/bin/env python3
"""Golden vectors of the reference's floating-point decoder FP_Decoder::decode_general(const double *)
(ArrayLDPC_Decoder.cpp:735-933), dumped from the reference's own objects (oracle/_ref/libref_<variant>.so, compiled
unmodified from /root/reference by oracle/build_ref.py).  Run where /root/reference exists:

    python tests/golden/make_golden_f64.py        -> tests/golden/reference_f64.npz

Channel: the drivers' line LLR = 2*snr*(1 - 2c + N(0, sigma)) (PerfTest.cpp:112) WITHOUT the quantiser, all-zero
codeword, numpy RNG with a fixed seed (the frames themselves are stored, so the generator does not matter)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, named_codes as nc, pyoracle as po  # noqa: E402

CASES = [("wifi", 2.0, 5), ("wifi", 1.0, 3), ("a5", 4.5, 4), ("a5", 3.0, 2), ("c79", 4.0, 2)]


def main():
    build_ref.build_ref()
    out = {}
    rng = np.random.default_rng(20261019)
    for name, db, frames in CASES:
        t = nc.tables(name)
        ref = po.Reference(name)
        ref.set_tables(t)
        snr = 2 * 10 ** (db / 10) * nc.channel_rate(name)
        llr = 2 * snr * (1 + np.sqrt(1 / snr) * rng.standard_normal((frames, t.n)))
        iters, bits, post = [], [], []
        for x in llr:
            it, b, p, _ = ref.decode_general(x)
            iters.append(it); bits.append(np.packbits(b.astype(np.uint8))); post.append(p)
        tag = "%s_%sdB" % (name, str(db).replace(".", "p"))
        out[tag + "_llr"] = llr
        out[tag + "_iters"] = np.array(iters, np.int32)
        out[tag + "_bits"] = np.array(bits)
        out[tag + "_post"] = np.array(post)
        print(tag, iters)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_f64.npz"), **out)


if __name__ == "__main__":
    main()
