"""Monte-Carlo mode (fused channel + decode + calculateBER) on the GPU."""
import numpy as np
import pytest

from conftest import tables_of

pytestmark = pytest.mark.gpu

WIFI_SNR_2DB = 2 * 10 ** (2.0 / 10) * 0.5


def test_reference_stream_channel_is_bit_exact(fp, golden, po):
    """LDPC_STREAM_REFERENCE regenerates the reference's Lehmer/Odeh-Evans noise in parallel: the first frames
    of the ArrayLDPC_Debug_Wifi flow (default seed, PerfTest.cpp:108-120) must come out bit for bit."""
    code = fp.codes.wifi_1944_r12()
    dec = fp.Decoder(code)
    cw = golden["wifi_codeword"]
    llr = dec.mc_channel(24, WIFI_SNR_2DB, stream=fp.STREAM_REFERENCE, seed=123456789, codeword=cw)
    assert (llr == golden["wifi_2dB_llr"]).all()
    # skip-ahead: frames 10..23 generated on their own
    part = dec.mc_channel(14, WIFI_SNR_2DB, stream=fp.STREAM_REFERENCE, seed=123456789, codeword=cw, first_frame=10)
    assert (part == golden["wifi_2dB_llr"][10:]).all()
    # a long stretch against the oracle's sequential generator (also pins the FMA-free arithmetic)
    orc = po.Oracle(None)
    orc.seed.value = 123456789
    want = np.array([orc.channel_frame(cw.astype(np.int32), code.n, WIFI_SNR_2DB, np.sqrt(1 / WIFI_SNR_2DB)) for _ in range(400)])
    assert (dec.mc_channel(400, WIFI_SNR_2DB, stream=fp.STREAM_REFERENCE, seed=123456789, codeword=cw) == want).all()
    dec.close()


def test_reference_stream_simulation_matches_golden_flow(fp, golden):
    code = fp.codes.wifi_1944_r12()
    dec = fp.Decoder(code)
    out = dec.mc_run(24, WIFI_SNR_2DB, stream=fp.STREAM_REFERENCE, seed=123456789, codeword=golden["wifi_codeword"],
                     info_index=golden["wifi_info_index"], want_iters=True)
    assert (out["iters"] == golden["wifi_2dB_iters"]).all()
    assert (out["frame_err"] == golden["wifi_2dB_ber"]).all()
    assert out["frames"] == 24 and out["iter_sum"] == golden["wifi_2dB_iters"].sum()
    # array code, decode_fixpoint flow (ArrayLDPC_Debug, PerfTest.cpp:276-311)
    a5 = fp.codes.array_p47_r5()
    dec5 = fp.Decoder(a5, precheck=True)
    snr = 2 * 10 ** (4.5 / 10) * a5.rate
    out = dec5.mc_run(24, snr, stream=fp.STREAM_REFERENCE, seed=123456789, codeword=golden["a5_codeword"],
                      info_index=golden["a5_info_index"], want_iters=True)
    assert (out["iters"] == golden["a5_4p5dB_iters"]).all()
    llr = dec5.mc_channel(24, snr, stream=fp.STREAM_REFERENCE, seed=123456789, codeword=golden["a5_codeword"])
    assert (llr == golden["a5_4p5dB_llr"]).all()
    dec.close(); dec5.close()


def test_wifi_results_transcript_reproduced_on_gpu(fp, golden):
    """wifi_results_4_4_2dB_30iter.txt: `2732 100 393214` (bit errors, frame errors, frames) for
    ArrayLDPC_Debug_Wifi at 2 dB with the default seed, stopping at the 100th frame error."""
    code = fp.codes.wifi_1944_r12()
    dec = fp.Decoder(code)
    frame_errors = bit_errors = counter = 0
    batch, first = 1 << 16, 0
    done = False
    while not done:
        out = dec.mc_run(batch, WIFI_SNR_2DB, stream=fp.STREAM_REFERENCE, seed=123456789, first_frame=first,
                         codeword=golden["wifi_codeword"], info_index=golden["wifi_info_index"])
        for e in out["frame_err"]:          # the reference's sequential stopping rule (PerfTest.cpp:97,131-134)
            counter += 1
            if e:
                frame_errors += 1
                bit_errors += int(e)
                if frame_errors == 100:
                    done = True
                    break
        first += batch
        assert first < 1 << 20
    assert (bit_errors, frame_errors, counter) == (2732, 100, 393214)
    assert "%g %g" % (frame_errors / counter, bit_errors / counter / 1944) == "0.000254314 3.57401e-06"
    dec.close()


def test_philox_stream_statistics_and_determinism(fp):
    code = fp.codes.cut79()
    dec = fp.Decoder(code)
    snr = 2 * 10 ** (3.0 / 10) * 0.8585
    sigma = np.sqrt(1 / snr)
    a = dec.mc_channel(512, snr, stream=fp.STREAM_PHILOX, seed=77)
    b = dec.mc_channel(512, snr, stream=fp.STREAM_PHILOX, seed=77)
    assert (a == b).all()
    c = dec.mc_channel(100, snr, stream=fp.STREAM_PHILOX, seed=77, first_frame=200)
    assert (c == a[200:300]).all()                      # frame g depends on (seed, g) only
    assert (dec.mc_channel(64, snr, stream=fp.STREAM_PHILOX, seed=78) != a[:64]).mean() > 0.9
    z = (a.astype(np.float64) + 0.5 * np.sign(a)) / (2 * snr * 16) - 1.0   # undo the quantiser (mid-point)
    z /= sigma
    nsamp = z.size
    assert abs(z.mean()) < 5 / np.sqrt(nsamp) + 2e-3
    assert abs(z.var() - 1.0) < 0.01
    assert abs((z ** 4).mean() - 3.0) < 0.05
    assert abs(np.corrcoef(z[:, :-1].ravel(), z[:, 1:].ravel())[0, 1]) < 5 / np.sqrt(nsamp)
    assert (np.abs(z) > 4).mean() == pytest.approx(6.33e-5, rel=0.5)
    dec.close()


@pytest.mark.parametrize("name,precheck,snr_db", [("wifi", False, 1.4), ("a5", True, 3.6)])
def test_simulation_equals_decode_of_its_own_channel(fp, po, name, precheck, snr_db):
    """mc_run == decode_batch(mc_channel) + host-side calculateBER, with a non-zero codeword and info positions."""
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    rng = np.random.default_rng(4)
    # a valid codeword is not needed for the bookkeeping identity; use random bits and random info positions
    cw = rng.integers(0, 2, code.n).astype(np.uint8)
    info = np.sort(rng.choice(code.n, fp.codes.INFO_BITS[name], replace=False)).astype(np.int32)
    snr = 2 * 10 ** (snr_db / 10) * fp.codes.INFO_BITS[name] / code.n
    dec = fp.Decoder(code, precheck=precheck)
    frames = 600
    sim = dec.mc_run(frames, snr, stream=fp.STREAM_PHILOX, seed=5, first_frame=1000, codeword=cw, info_index=info,
                     want_iters=True)
    llr = dec.mc_channel(frames, snr, stream=fp.STREAM_PHILOX, seed=5, first_frame=1000, codeword=cw)
    out = dec.decode(llr)
    bits = fp.unpack_bits(out["bits"], code.n)
    errs = (bits[:, info] != cw[info][None, :]).sum(axis=1)
    assert (sim["iters"] == out["iters"]).all()
    assert (sim["frame_err"] == errs).all()
    assert sim["frames"] == frames and sim["bit_errors"] == errs.sum() and sim["frame_errors"] == (errs > 0).sum()
    assert sim["iter_sum"] == out["iters"].sum()
    # spot check the decode itself against the oracle
    orc = po.Oracle(t)
    for f in range(0, frames, 97):
        assert orc.decode(llr[f], precheck=precheck)[0] == out["iters"][f]
    dec.close()


def test_shortening_pins(fp, golden):
    """ArrayLDPC_Debug_Shorten (PerfTest.cpp:410-414): the first short_len info positions are forced to 7*16."""
    code = fp.codes.array_p47_r5()
    dec = fp.Decoder(code, precheck=True)
    info = golden["a5_info_index"].astype(np.int32)
    snr = 2 * 10 ** (4.5 / 10) * (1978.0 - 976.0) / 2209.0
    llr = dec.mc_channel(8, snr, stream=fp.STREAM_REFERENCE, seed=123456789, pin_index=info[:36], pin_value=7 * 16)
    base = dec.mc_channel(8, snr, stream=fp.STREAM_REFERENCE, seed=123456789)
    assert (llr[:, info[:36]] == 112).all()
    rest = np.setdiff1d(np.arange(code.n), info[:36])
    assert (llr[:, rest] == base[:, rest]).all()
    sim = dec.mc_run(8, snr, stream=fp.STREAM_REFERENCE, seed=123456789, pin_index=info[:36], pin_value=112,
                     info_index=info, want_iters=True)
    assert (sim["iters"] == dec.decode(llr)["iters"]).all()
    dec.close()


def test_driver_run_point_reproduces_transcript(fp, golden):
    """The sharded driver (one rank here) with the reference stream prints the reference's own two lines."""
    from fixedpointldpc_b200.montecarlo import Shards, gpu_simulator, reference_print, run_point
    code = fp.codes.wifi_1944_r12()
    dec = fp.Decoder(code)
    sim = gpu_simulator(dec, WIFI_SNR_2DB, stream=fp.STREAM_REFERENCE, seed=123456789,
                        codeword=golden["wifi_codeword"], info_index=golden["wifi_info_index"])
    res = run_point(sim, Shards(1, 0, 50000), 100)
    assert reference_print(res, code.n) == "2732 100 393214\n FER: 0.000254314 BER: 3.57401e-06"
    dec.close()


@pytest.mark.parametrize("name", ["wifi", "a5"])
def test_gpu_encoder_matches_host_encoder_and_satisfies_h(fp, po, golden, name):
    """Batched bit-packed popcount encoder == FP_Encoder::encode (ArrayLDPC_Encoder.cpp:160-225) message by message."""
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    parity = np.setdiff1d(np.arange(code.n), golden[name + "_info_index"].astype(np.int64)).astype(np.int32)
    gen = fp.Generator(code=code, parity_cols=parity)
    rng = np.random.default_rng(8)
    kb = (gen.k + 7) // 8
    msgs = rng.integers(0, 256, size=(300, kb), dtype=np.uint8)
    msgs[0] = 0
    msgs[1] = 255
    packed = gen.encode_batch(msgs)
    bits = fp.unpack_bits(packed, code.n)
    for f in range(0, 300, 7):
        assert (bits[f] == gen.encode(bytes(msgs[f]))).all()
    # fixed driver message reproduces the reference's codeword
    fixed = {"wifi": b"OMG  how long   dd   should this string be to make it 243".ljust(122, b"\0")}.get(name)
    if fixed is not None:
        one = fp.unpack_bits(gen.encode_batch(np.frombuffer(fixed, np.uint8)[None, :]), code.n)[0]
        assert (one == golden["wifi_codeword"]).all()
    # every codeword satisfies H
    syn = np.zeros((300, t.m), np.int32)
    for k in range(t.dc_max):
        col = t.clist[:, k]
        ok = col >= 0
        syn[:, ok] ^= bits[:, col[ok]]
    assert (syn == 0).all()
    # linearity: enc(a ^ b) == enc(a) ^ enc(b)
    x = gen.encode_batch(msgs[10:20] ^ msgs[20:30])
    assert (x == (packed[10:20] ^ packed[20:30])).all()


def test_random_message_simulation_with_per_frame_codewords(fp, golden):
    """Random messages per frame: encode on the GPU, send each frame's own codeword, count errors against it."""
    import torch
    code = fp.codes.array_p47_r5()
    parity = np.setdiff1d(np.arange(code.n), golden["a5_info_index"].astype(np.int64)).astype(np.int32)
    gen = fp.Generator(code=code, parity_cols=parity)
    frames, kw = 512, (gen.k + 31) // 32
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    info = torch.randint(0, 2 ** 31 - 1, (frames, kw), generator=g, device="cuda", dtype=torch.int32)
    info[:, -1] &= (1 << (gen.k % 32)) - 1
    cws = torch.zeros((frames, code.nw32), dtype=torch.int32, device="cuda")
    gen.encode_batch_device(info.data_ptr(), frames, cws.data_ptr())
    torch.cuda.synchronize()
    dec = fp.Decoder(code, precheck=True)
    snr = 2 * 10 ** (4.2 / 10) * code.rate
    sim = dec.mc_run(frames, snr, stream=fp.STREAM_PHILOX, seed=9, d_codewords=cws.data_ptr(),
                     info_index=golden["a5_info_index"], want_iters=True)
    llr = dec.mc_channel(frames, snr, stream=fp.STREAM_PHILOX, seed=9, d_codewords=cws.data_ptr())
    out = dec.decode(llr)
    sent = fp.unpack_bits(cws.cpu().numpy().view(np.uint32), code.n)
    got = fp.unpack_bits(out["bits"], code.n)
    idx = golden["a5_info_index"].astype(np.int64)
    errs = (got[:, idx] != sent[:, idx]).sum(axis=1)
    assert (sim["iters"] == out["iters"]).all() and (sim["frame_err"] == errs).all()
    # hard decisions of the noiseless part agree with the codeword sign convention (bit 1 -> negative LLR)
    assert ((llr < 0) == (sent == 1)).mean() > 0.9
    dec.close()


def test_array_debug_flow_20000_frames_known_answer(fp, golden):
    """ArrayLDPC_Debug flow (A5, 4.5 dB, decode_fixpoint, reference noise from the default seed):
    the first 20 000 frames hold 993 frame errors, 27 011 info-bit errors and the recorded iteration histogram."""
    code = fp.codes.array_p47_r5()
    dec = fp.Decoder(code, precheck=True)
    snr = 2 * 10 ** (4.5 / 10) * code.rate
    out = dec.mc_run(20000, snr, stream=fp.STREAM_REFERENCE, seed=123456789, codeword=golden["a5_codeword"],
                     info_index=golden["a5_info_index"], want_iters=True)
    assert out["frame_errors"] == 993 and out["bit_errors"] == 27011
    hist = np.bincount(out["iters"], minlength=31)
    assert [hist[1], hist[2], hist[3], hist[4], hist[5], hist[30]] == [8, 920, 4697, 5037, 3115, 1022]
    dec.close()


def test_philox_fer_agrees_with_reference_stream_within_confidence(fp, golden):
    """Different noise stream, same statistics: the frame error rate at 2 dB (802.11) measured with the Philox
    stream must be compatible with the reference's 100 / 393 214 (two-sample z-test, |z| < 3)."""
    code = fp.codes.wifi_1944_r12()
    dec = fp.Decoder(code)
    frames = 400000
    out = dec.mc_run(frames, WIFI_SNR_2DB, stream=fp.STREAM_PHILOX, seed=2026, codeword=golden["wifi_codeword"],
                     info_index=golden["wifi_info_index"], want_frame_err=False)
    p1, n1, p2, n2 = out["frame_errors"] / frames, frames, 100 / 393214, 393214
    pooled = (out["frame_errors"] + 100) / (n1 + n2)
    z = (p1 - p2) / np.sqrt(pooled * (1 - pooled) * (1 / n1 + 1 / n2))
    assert abs(z) < 3, (out["frame_errors"], z)
    # bit errors per failed frame are of the reference's order (2732 / 100)
    assert 10 < out["bit_errors"] / max(1, out["frame_errors"]) < 60
    dec.close()


def test_array_philox_rates_agree_with_reference_stream_within_confidence(fp, golden):
    """BASELINE config 5 runs the array p47 r5 waterfall on the Philox stream; its rates must be compatible with the
    reference noise stream's known answer at 4.5 dB (993 frame errors in 20 000 frames):
    two-sample z-test on the frame error rate (|z| < 3.5), mean iteration count within 1 %, bit errors per failed
    frame of the reference's order (27 011 / 993)."""
    code = fp.codes.array_p47_r5()
    dec = fp.Decoder(code, precheck=True)
    snr = 2 * 10 ** (4.5 / 10) * code.rate
    ref = dec.mc_run(20000, snr, stream=fp.STREAM_REFERENCE, seed=123456789, codeword=golden["a5_codeword"],
                     info_index=golden["a5_info_index"], want_frame_err=False)
    assert ref["frame_errors"] == 993
    frames = 400000
    out = dec.mc_run(frames, snr, stream=fp.STREAM_PHILOX, seed=55, codeword=golden["a5_codeword"],
                     info_index=golden["a5_info_index"], want_frame_err=False)
    p1, n1, p2, n2 = out["frame_errors"] / frames, frames, 993 / 20000, 20000
    pooled = (out["frame_errors"] + 993) / (n1 + n2)
    z = (p1 - p2) / np.sqrt(pooled * (1 - pooled) * (1 / n1 + 1 / n2))
    assert abs(z) < 3.5, (out["frame_errors"], z)
    assert abs(out["iter_sum"] / frames - ref["iter_sum"] / 20000) < 0.01 * ref["iter_sum"] / 20000 + 0.05
    assert 0.8 * 27011 / 993 < out["bit_errors"] / out["frame_errors"] < 1.2 * 27011 / 993
    dec.close()
