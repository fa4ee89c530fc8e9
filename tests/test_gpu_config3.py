"""BASELINE config 3 as a pinned workload: the shortened array code H2212_316_array_cut79 (Format C file; no reader,
generator or driver exists for it in the reference), with a generator derived from H, the shortening rule of
ArrayLDPC_Debug_Shorten (PerfTest.cpp:362-366 zero the message, :410-414 pin the known positions to LLR 7*2^4), and a
three-point Eb/N0 sweep -- the GPU counters must EQUAL the counters of the sequential CPU oracle run on the reference's
own noise stream (same seed, same frames, same stopping frame)."""
import os

import numpy as np
import pytest

from conftest import tables_of

pytestmark = pytest.mark.gpu


def test_cut79_shortened_sweep_counters_equal_oracle(fp, po, tmp_path):
    code = fp.codes.cut79()
    t = tables_of(code)
    gen = fp.Generator(code=code)                       # GF(2) elimination on H (the reference ships no G for this code)
    assert gen.k == fp.codes.INFO_BITS["c79"] == 1899
    path = os.path.join(tmp_path, "G_cut79.txt")
    gen.save(path)
    og = po.read_format_b(path)                         # the oracle's own Format-B reader
    assert (og.info_index == gen.info_index).all()
    rng = np.random.default_rng(79)
    short_len = 300                                     # known (zero) information bits
    info = bytearray(rng.integers(0, 256, (gen.k + 7) // 8, dtype=np.uint8).tobytes())
    for i in range(short_len // 8 + 1):                 # zero whole bytes like the reference's loop over InfoStream
        info[i] = 0
    info = bytes(info)
    orc = po.Oracle(t)
    cw = orc.encode(og, info)
    assert (gen.encode(info) == cw).all()               # host encoder of the product == oracle encoder
    assert orc.decode(np.where(cw > 0, -100, 100).astype(np.int32))[0] == 1  # H c = 0: converges at once
    true_info = orc.set_info_bit(info, gen.k)
    pins = gen.info_index[:short_len].astype(np.int32)
    assert (cw[pins] == 0).all()
    rate = (gen.k - short_len) / code.n
    ndev = min(2, fp.device_count())
    decs = [fp.Decoder(code, device=d) for d in range(ndev)]
    grp = fp.McGroup(decs)
    seen_errors = 0
    for db, target, cap in ((3.8, 20, 1500), (4.2, 12, 2500), (4.6, 6, 4000)):   # 48 / 216 / 1318 frames
        snr = 2 * 10 ** (db / 10) * rate
        sigma = np.sqrt(1 / snr)
        # the reference's loop, sequentially, on the oracle (every driver starts from the default seed)
        orc.seed.value = 123456789
        frames = bit_errors = frame_errors = iter_sum = 0
        hist = np.zeros(32, np.int64)
        while frame_errors < target and frames < cap:
            llr = orc.channel_frame(cw, code.n, snr, sigma)
            llr[pins] = 7 * 16
            it, bits, _, _ = orc.decode(llr)
            e = orc.calculate_ber(bits, gen.info_index, true_info)
            frames += 1; bit_errors += e; frame_errors += e > 0; iter_sum += it; hist[it] += 1
        res = grp.run(snr, sigma=sigma, target_block_errors=target, max_frames=cap, frames_per_round=333,
                      stream=fp.STREAM_REFERENCE, seed=123456789, codeword=cw.astype(np.uint8),
                      info_index=gen.info_index, pin_index=pins, pin_value=7 * 16)
        assert (res["errors"], res["block_errors"], res["frames"], res["iter_sum"]) == (bit_errors, frame_errors, frames, iter_sum), (db, res)
        assert (res["iter_hist"] == hist).all()
        seen_errors += frame_errors
    assert seen_errors > 0
    grp.close()
    for d in decs:
        d.close()


def test_console_sweep_on_shortened_cut79_equals_oracle(fp, po, tmp_path):
    """`ldpc_wrapper_c79 sweep 3.8 4.6 0.4 out.csv 10 300`: the console program's shortened sweep (all-zero codeword,
    generator derived from H, first 300 information positions pinned, reference noise stream continuing from point to
    point like the reference's file-static RNG state) against the same loop run sequentially on the CPU oracle."""
    import subprocess
    from conftest import ROOT
    code = fp.codes.cut79()
    t = tables_of(code)
    gen = fp.Generator(code=code)
    short_len, target = 300, 10
    pins = gen.info_index[:short_len].astype(np.int32)
    rate = (gen.k - short_len) / code.n
    exe = os.path.join(ROOT, "fixedpointldpc_b200", "ldpc_wrapper_c79")
    res = subprocess.run([exe, "sweep", "3.8", "4.6", "0.4", "out.csv", str(target), str(short_len)], capture_output=True,
                         text=True, cwd=str(tmp_path), timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    rows = open(os.path.join(tmp_path, "out.csv")).read().strip().split("\n")[1:]
    orc = po.Oracle(t)
    orc.seed.value = 123456789                      # rngs.cpp:45; never re-seeded between points
    zeros = np.zeros(gen.k, np.int32)
    db, want = 3.8, []
    while db <= 4.6 + 1e-9:
        snr = 2 * 10.0 ** (db / 10) * rate
        sigma = np.sqrt(1 / snr)
        frames = bit_errors = frame_errors = iter_sum = 0
        while frame_errors < target:
            llr = orc.channel_frame(None, code.n, snr, sigma)
            llr[pins] = 7 * 16
            it, bits, _, _ = orc.decode(llr)
            e = orc.calculate_ber(bits, gen.info_index, zeros)
            frames += 1; bit_errors += e; frame_errors += e > 0; iter_sum += it
        want.append((frames, frame_errors, bit_errors, iter_sum))
        db += 0.4
    assert len(rows) == len(want) == 3
    for row, (frames, fe, be, its) in zip(rows, want):
        cols = row.split(",")
        assert (int(cols[1]), int(cols[2]), int(cols[3])) == (frames, fe, be), (row, frames, fe, be)
        assert abs(float(cols[6]) - its / frames) < 1e-4 * max(1.0, its / frames)
