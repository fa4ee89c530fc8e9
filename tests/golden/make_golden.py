#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference itself (run where /root/reference exists).

Every vector comes from the UNMODIFIED reference objects (oracle/_ref/libref_<variant>.so, see
oracle/build_ref.py) driven exactly like the reference's drivers:
  wifi : ArrayLDPC_Debug_Wifi flow (PerfTest.cpp:23-140), message PerfTest.cpp:33, 2 dB, decode_general_fp
  a5   : ArrayLDPC_Debug flow (PerfTest.cpp:217-316), message :221-224, 4.5 dB, decode_fixpoint,
         plus ArrayLDPC_PerfTest flow (all-zero codeword, 2 dB) frames that run 30 iterations
  a24  : DecodeTrial-style all-zero frames at 3 dB and 6 dB, decode_fixpoint
  c79  : all-zero frames at 2 dB and 4.5 dB, decode_general_fp (tables from H2212_316_array_cut79.txt)
plus the sxor grid, RNG / Normal known answers, parsed tables and the reference's printed transcripts
(captured from oracle/_ref/wrapper_<variant>, the reference's own console program).
Run:  python tests/golden/make_golden.py     (about 15 s; the 565 s wifi transcript run is optional: --transcript)
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
WIFI_MSG = b"OMG  how long   dd   should this string be to make it 243".ljust(122, b"\0")
A5_MSG = (b"OMG how long should this string be to make it 248, just imagine that. "
          b"\t\t\t\t\t\t\t\t   I guess it's still not long enough. Let's see. This is a testing string "
          b"\t\t\t\t\t\t\t\t\tfor a lot of characters so that we have some random bit stream that's"
          b"\t\t\t\t\t\t\t\t\tcorrect").ljust(248, b"\0")


def valid_edges(t, edge):
    mask = t.cdeg[None, :] > np.arange(edge.shape[0])[:, None]
    return np.where(mask, edge, 0).astype(np.int32)


def frames_block(ref, t, decode, codeword, snr, sigma, count):
    llr, iters, bits, post, edge = [], [], [], [], []
    for _ in range(count):
        x = ref.channel_frame(codeword, snr, sigma)
        it, b, p, e = decode(x)
        llr.append(x); iters.append(it); bits.append(b.copy()); post.append(p.copy())
        edge.append(valid_edges(t, e[:t.dc_max]))
    return dict(llr=np.array(llr, np.int16), iters=np.array(iters, np.int32), bits=np.packbits(np.array(bits, np.uint8), axis=1),
                post=np.array(post, np.int32), edge=np.array(edge, np.int16))


def main():
    if not os.path.isdir(REF):
        raise SystemExit("needs /root/reference")
    g = {}
    # ---- scalar known answers
    r = po.Reference("wifi")
    g["sxor_grid_m300_300"] = r.sxor_grid(-300, 300).astype(np.int16)
    g["sxor_grid_digest_m1024_1024"] = np.frombuffer(po.fnv1a64(r.sxor_grid(-1024, 1024)).encode(), np.uint8)
    big = np.array([[x, y, r.sxor(x, y)] for x in (-40000, -4770, -1000, -257, -256, -255, -1, 0, 1, 255, 256, 257, 1000, 4770, 40000, 2 ** 20 + 3)
                    for y in (-2 ** 20 - 7, -4000, -1000, -300, -60, -1, 0, 1, 10, 60, 255, 256, 300, 4000, 65535, 2 ** 21)], np.int64)
    g["sxor_big"] = big
    r.put_seed(123456789)
    g["random_first"] = np.array([r.random() for _ in range(1000)])
    r.put_seed(123456789)
    g["normal_first"] = np.array([r.normal(0.0, 1.0) for _ in range(1000)])
    r.put_seed(1)
    for _ in range(10000):
        r.random()
    g["seed_after_10000_from_1"] = np.array([r.get_seed()], np.int64)  # rngs.cpp:42 CHECK = 399268537

    # ---- wifi: Debug_Wifi flow
    r.read_h(REF)
    t = r.get_tables()
    g["wifi_vdeg"], g["wifi_cdeg"], g["wifi_vlist"], g["wifi_clist"] = t.vdeg, t.cdeg, t.vlist.astype(np.int16), t.clist.astype(np.int16)
    assert r.encoder_open(os.path.join(REF, "H_802.11_IndZerog.txt")) == 0
    cw, idx = r.encode(WIFI_MSG)
    assert int(cw.sum()) == 675, cw.sum()  # codeword weight 675
    g["wifi_codeword"], g["wifi_info_index"] = cw.astype(np.uint8), idx.astype(np.int16)
    r.set_info(WIFI_MSG, idx)
    r.put_seed(123456789)
    snr = 2 * 10 ** (2.0 / 10) * 0.5
    blk = frames_block(r, t, r.decode_general_fp, cw, snr, np.sqrt(1 / snr), 24)
    for k, v in blk.items():
        g["wifi_2dB_" + k] = v
    ber = []
    r.put_seed(123456789)
    for i in range(24):
        r.decode_general_fp(r.channel_frame(cw, snr, np.sqrt(1 / snr)))
        ber.append(r.calculate_ber())
    g["wifi_2dB_ber"] = np.array(ber, np.int32)
    r.put_seed(987654321)
    snr = 2 * 10 ** (0.5 / 10) * 0.5
    blk = frames_block(r, t, r.decode_general_fp, cw, snr, np.sqrt(1 / snr), 8)
    for k, v in blk.items():
        g["wifi_0p5dB_" + k] = v

    # ---- a5: Debug flow (decode_fixpoint) + PerfTest flow
    r5 = po.Reference("a5")
    t5 = po.read_alist_a(os.path.join(REF, "H_array_p47_r5_forward.txt"))
    r5.set_tables(t5)
    assert r5.encoder_open(os.path.join(REF, "codes", "G_array_forward.txt")) == 0
    cw5, idx5 = r5.encode(A5_MSG)
    assert len(A5_MSG) == 248 and int(cw5.sum()) == 964, cw5.sum()  # codeword weight 964
    g["a5_codeword"], g["a5_info_index"] = cw5.astype(np.uint8), idx5.astype(np.int16)
    g["a5_rate"] = np.array([r5.rate()])
    r5.put_seed(123456789)
    snr = 2 * 10 ** (4.5 / 10) * r5.rate()
    blk = frames_block(r5, t5, r5.decode_fixpoint, cw5, snr, np.sqrt(1 / snr), 24)
    for k, v in blk.items():
        g["a5_4p5dB_" + k] = v
    r5.put_seed(123456789)
    snr = 2 * 10 ** (2.0 / 10) * r5.rate()
    blk = frames_block(r5, t5, r5.decode_fixpoint, None, snr, np.sqrt(1 / snr), 6)
    for k, v in blk.items():
        g["a5_2dB_" + k] = v
    # pre-check hits (decode_fixpoint returns 0): high SNR
    r5.put_seed(42)
    snr = 2 * 10 ** (9.0 / 10) * r5.rate()
    blk = frames_block(r5, t5, r5.decode_fixpoint, cw5, snr, np.sqrt(1 / snr), 6)
    g["a5_9dB_llr"], g["a5_9dB_iters"], g["a5_9dB_bits"] = blk["llr"], blk["iters"], blk["bits"]

    # ---- a24
    r24 = po.Reference("a24")
    t24 = po.read_alist_a(os.path.join(REF, "codes", "H_array_p47_r24_forward.txt"))
    r24.set_tables(t24)
    g["a24_rate"] = np.array([r24.rate()])
    for tag, db, cnt in (("3dB", 3.0, 2), ("6dB", 6.0, 4)):
        r24.put_seed(2024)
        snr = 2 * 10 ** (db / 10) * r24.rate()
        blk = frames_block(r24, t24, r24.decode_fixpoint, None, snr, np.sqrt(1 / snr), cnt)
        for k, v in blk.items():
            g["a24_%s_%s" % (tag, k)] = v

    # ---- c79
    r79 = po.Reference("c79")
    t79 = po.read_format_c(os.path.join(REF, "H2212_316_array_cut79.txt"))
    r79.set_tables(t79)
    g["c79_cdeg"], g["c79_clist"] = t79.cdeg, t79.clist.astype(np.int16)
    for tag, db, cnt in (("2dB", 2.0, 4), ("4p5dB", 4.5, 12)):
        r79.put_seed(79)
        snr = 2 * 10 ** (db / 10) * 1899.0 / 2212.0
        blk = frames_block(r79, t79, r79.decode_general_fp, None, snr, np.sqrt(1 / snr), cnt)
        for k, v in blk.items():
            g["c79_%s_%s" % (tag, k)] = v

    np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **g)
    print("wrote reference_vectors.npz with %d arrays" % len(g))

    # ---- transcripts of the reference's own console program
    lines = []
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink(os.path.join(REF, "codes", "G_array_forward.txt"), os.path.join(tmp, "G_array_forward.txt"))
        out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "wrapper_a5"), "2", "2", "1", "t.csv"], cwd=tmp,
                             capture_output=True, text=True).stdout
        lines.append("# wrapper_a5 a b c d  -> ArrayLDPC_PerfTest(2,2,1,\"test.csv\") (Wrapper.cpp:27)\n" + out)
    # the drivers the reference's own main leaves unreachable, through the dispatching main of
    # oracle/ref_driver_main.cpp: ArrayLDPC_Debug_Shorten (PerfTest.cpp:318-431, prints every return value),
    # ArrayLDPC_Debug (:217-316), ArrayLDPC_TimeTrial (:520-607)
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink(os.path.join(REF, "codes", "G_array_forward.txt"), os.path.join(tmp, "G_array_forward.txt"))
        with open(os.path.join(OUT, "reference_drivers.txt"), "w") as fh:
            for args in ("shorten 36", "shorten 200", "debug", "timetrial 2 300", "timetrial 6 500"):
                out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "wrapper_a5")] + args.split(), cwd=tmp,
                                     capture_output=True, text=True).stdout
                fh.write("# wrapper_a5 %s\n" % args + out)
    if "--transcript" in sys.argv:
        with tempfile.TemporaryDirectory() as tmp:
            for f in ("H_802.11_IndZero.txt", "H_802.11_IndZerog.txt"):
                os.symlink(os.path.join(REF, f), os.path.join(tmp, f))
            out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "wrapper_wifi")], cwd=tmp, input="2\n",
                                 capture_output=True, text=True).stdout
            lines.append("# echo 2 | wrapper_wifi -> ArrayLDPC_Debug_Wifi (Wrapper.cpp:32); compare wifi_results_4_4_2dB_30iter.txt\n" + out)
        with open(os.path.join(OUT, "reference_transcripts.txt"), "w") as fh:
            fh.write("\n".join(lines))
    else:
        print(lines[0])


if __name__ == "__main__":
    main()
