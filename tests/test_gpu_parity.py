"""Parity tests proper: the CUDA engine through the C ABI vs the C oracle and the golden vectors.

Bit-exact bar (integer path): iteration counts, decoded bits, posteriors and the final
variable-to-check messages in the reference's EdgeRAM [slot][check] order."""
import numpy as np
import pytest

from conftest import channel_frames, tables_of, valid_mask

pytestmark = pytest.mark.gpu

CASES = [  # code, Eb/N0 points, frames
    ("wifi", (2.0, 0.5), 96),
    ("a5", (4.5, 2.0), 64),
    ("c79", (4.5, 2.0), 64),
    ("a24", (6.0, 3.0), 10),
]


def _expect(orc, llr, precheck):
    return [orc.decode(x, precheck=precheck) for x in llr]


def _compare(fp, code, t, out, want, check_state=True):
    bits = fp.unpack_bits(out["bits"], code.n)
    mask = valid_mask(t)
    for f, (it, b, post, edge) in enumerate(want):
        assert out["iters"][f] == it, ("iters", f, out["iters"][f], it)
        assert (bits[f] == b).all(), ("bits", f)
        if check_state and it > 0:  # iters == 0: the reference leaves post / EdgeRAM stale (quirk Q6)
            assert (out["post"][f] == post).all(), ("post", f)
            assert (out["v2c"][f][mask] == edge[mask]).all(), ("v2c", f)
            assert (out["v2c"][f][~mask] == 0).all()


@pytest.mark.parametrize("name,snrs,frames", CASES)
@pytest.mark.parametrize("precision", [32, 16, 0])
@pytest.mark.parametrize("precheck", [False, True])
def test_decode_matches_oracle(fp, po, name, snrs, frames, precision, precheck):
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    orc = po.Oracle(t)
    rate = fp.codes.INFO_BITS[name] / code.n
    dec = fp.Decoder(code, precheck=precheck, precision=precision)
    for snr in snrs:
        llr = channel_frames(code.n, rate, snr, frames, seed=int(snr * 100) + frames)
        out = dec.decode(llr, want_post=True, want_v2c=True)
        _compare(fp, code, t, out, _expect(orc, llr, precheck))
    assert dec.stats()["kernel_launches"] > 0
    dec.close()


@pytest.mark.parametrize("name,tag,precheck", [("wifi", "wifi_2dB", False), ("wifi", "wifi_0p5dB", False),
                                               ("a5", "a5_4p5dB", True), ("a5", "a5_2dB", True), ("a24", "a24_3dB", True),
                                               ("a24", "a24_6dB", True), ("c79", "c79_2dB", False), ("c79", "c79_4p5dB", False)])
@pytest.mark.parametrize("precision", [32, 0])
def test_decode_matches_reference_golden(fp, golden, name, tag, precheck, precision):
    """Vectors dumped from the reference's own FP_Decoder (tests/golden/make_golden.py)."""
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    mask = valid_mask(t)
    llr = golden[tag + "_llr"].astype(np.int32)
    dec = fp.Decoder(code, precheck=precheck, precision=precision)
    out = dec.decode(llr, want_post=True, want_v2c=True)
    bits = fp.unpack_bits(out["bits"], code.n)
    want_bits = np.unpackbits(golden[tag + "_bits"], axis=1)[:, :code.n]
    assert (out["iters"] == golden[tag + "_iters"]).all()
    assert (bits == want_bits).all()
    for f in range(len(llr)):
        if out["iters"][f] > 0:
            assert (out["post"][f] == golden[tag + "_post"][f]).all()
            assert (out["v2c"][f][mask] == golden[tag + "_edge"][f].astype(np.int32)[mask]).all()
    dec.close()


def test_precheck_hits_golden(fp, golden):
    code = fp.codes.array_p47_r5()
    llr = golden["a5_9dB_llr"].astype(np.int32)
    for precision in (32, 16):
        dec = fp.Decoder(code, precheck=True, precision=precision)
        out = dec.decode(llr)
        assert (out["iters"] == golden["a5_9dB_iters"]).all() and (golden["a5_9dB_iters"] == 0).any()
        assert (fp.unpack_bits(out["bits"], code.n) == np.unpackbits(golden["a5_9dB_bits"], axis=1)[:, :code.n]).all()
        # without the pre-check the same frames take one iteration (decode_general_fp, quirk Q6)
        dec2 = fp.Decoder(code, precheck=False, precision=precision)
        out2 = dec2.decode(llr)
        assert (out2["iters"][golden["a5_9dB_iters"] == 0] == 1).all()
        dec.close(); dec2.close()


def test_edge_cases(fp, po):
    code = fp.codes.cut79()
    t = tables_of(code)
    orc = po.Oracle(t)
    dec = fp.Decoder(code)
    # empty batch
    out = dec.decode(np.zeros((0, code.n), np.int32))
    assert out["iters"].shape == (0,)
    # one frame, odd frame counts (ragged against the two-frames-per-word packing and the slot count)
    for frames in (1, 3, 11, 257):
        llr = channel_frames(code.n, 0.86, 3.0, frames, seed=frames)
        out = dec.decode(llr, want_post=True, want_v2c=True)
        _compare(fp, code, t, out, _expect(orc, llr, False))
    # all-zero LLRs: sgn(0) = -1, bit(0) = 1 (quirk Q3); ties everywhere
    llr = np.zeros((4, code.n), np.int32)
    out = dec.decode(llr, want_post=True, want_v2c=True)
    _compare(fp, code, t, out, _expect(orc, llr, False))
    # +-1 only, and a constant frame
    rng = np.random.default_rng(0)
    llr = rng.choice(np.array([-1, 1], np.int32), size=(6, code.n))
    llr[5] = 7
    out = dec.decode(llr, want_post=True, want_v2c=True)
    _compare(fp, code, t, out, _expect(orc, llr, False))
    dec.close()


@pytest.mark.parametrize("name", ["wifi", "a5"])
def test_guard_fallback_is_exact(fp, po, name):
    """Values outside the packed kernel's 13-bit range: auto precision must re-decode them exactly in int32,
    packed-only precision must flag them (iters == -1) instead of returning a wrong answer."""
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    orc = po.Oracle(t)
    rate = fp.codes.INFO_BITS[name] / code.n
    llr = channel_frames(code.n, rate, 2.5, 24, seed=5)
    llr[::3] *= 40          # |LLR| up to ~2e4: beyond int16 guard, far below the int32 domain limit
    llr[1::3] *= 9          # borderline: grows past 2^13 during decoding
    want = _expect(orc, llr, False)
    dec = fp.Decoder(code, precision=0)
    out = dec.decode(llr, want_post=True, want_v2c=True)
    _compare(fp, code, t, out, want)
    assert dec.stats()["fallback_frames"] >= 8
    dec16 = fp.Decoder(code, precision=16)
    out16 = dec16.decode(llr)
    flagged = out16["iters"] < 0
    assert flagged[::3].all()
    ok = ~flagged
    assert (out16["iters"][ok] == np.array([w[0] for w in want])[ok]).all()
    dec.close(); dec16.close()


def test_max_iter_and_small_configs(fp, po):
    code = fp.codes.wifi_1944_r12()
    t = tables_of(code)
    llr = channel_frames(code.n, 0.5, 1.0, 40, seed=3)
    for max_iter in (1, 2, 7):
        orc = po.Oracle(t, max_iter=max_iter)
        dec = fp.Decoder(code, max_iter=max_iter)
        out = dec.decode(llr, want_post=True, want_v2c=True)
        _compare(fp, code, t, out, _expect(orc, llr, False))
        dec.close()
    # launch-shape overrides must not change results
    orc = po.Oracle(t)
    want = _expect(orc, llr, False)
    for threads, slots in ((128, 2), (512, 3), (1024, 0)):
        dec = fp.Decoder(code, threads=threads, frames_per_cta=slots)
        _compare(fp, code, t, dec.decode(llr, want_post=True, want_v2c=True), want)
        dec.close()


def test_generic_kernel_on_irregular_random_code(fp, po):
    """A code that matches none of the exact instantiations (falls to the DC<=64 / DV<=32 bucket)."""
    rng = np.random.default_rng(11)
    n, m = 600, 200
    rows = []
    for c in range(m):
        d = int(rng.integers(3, 20))
        rows.append(np.sort(rng.choice(n, d, replace=False)))
    dc = max(len(r) for r in rows)
    clist = np.full((m, dc), -1, np.int32)
    for c, r in enumerate(rows):
        clist[c, :len(r)] = r
    code = fp.Code.from_checks(n, [len(r) for r in rows], clist)
    t = tables_of(code)
    orc = po.Oracle(t)
    llr = channel_frames(n, 0.6, 3.0, 40, seed=9)
    for precision in (32, 0):
        dec = fp.Decoder(code, precision=precision)
        _compare(fp, code, t, dec.decode(llr, want_post=True, want_v2c=True), _expect(orc, llr, False))
        dec.close()


def test_full_size_properties(fp, po):
    """BASELINE-size batch (2^17 frames of the 802.11 code): size-independent properties.
    (1) the result does not depend on batch composition / slot scheduling: a shuffled batch gives the
        shuffled result; (2) every frame with iters < 30 satisfies H; (3) a sample equals the oracle;
    (4) int16 device input == int32 host input."""
    import torch
    code = fp.codes.wifi_1944_r12()
    t = tables_of(code)
    frames = 1 << 17
    llr = channel_frames(code.n, 0.5, 1.6, frames, seed=2026)
    dec = fp.Decoder(code)
    out = dec.decode(llr)
    perm = np.random.default_rng(1).permutation(frames)
    out_p = dec.decode(llr[perm])
    assert (out_p["iters"] == out["iters"][perm]).all() and (out_p["bits"] == out["bits"][perm]).all()
    bits = fp.unpack_bits(out["bits"][:4096], code.n)
    conv = out["iters"][:4096] < 30
    syn = np.zeros((4096, t.m), np.int32)
    for k in range(t.dc_max):
        col = t.clist[:, k]
        ok = col >= 0
        syn[:, ok] ^= bits[:, col[ok]]
    assert (syn[conv] == 0).all() and conv.sum() > 1000
    orc = po.Oracle(t)
    sample = np.random.default_rng(2).choice(frames, 64, replace=False)
    for f in sample:
        it, b, _, _ = orc.decode(llr[f])
        assert out["iters"][f] == it and (fp.unpack_bits(out["bits"][f:f + 1], code.n)[0] == b).all()
    d_llr = torch.from_numpy(llr.astype(np.int16)).cuda()
    d_iters = torch.zeros(frames, dtype=torch.int32, device="cuda")
    d_bits = torch.zeros((frames, code.nw32), dtype=torch.int32, device="cuda")
    dec.decode_device(d_llr.data_ptr(), 16, frames, d_iters.data_ptr(), d_bits.data_ptr())
    dec.sync()
    assert (d_iters.cpu().numpy() == out["iters"]).all()
    assert (d_bits.cpu().numpy().view(np.uint32) == out["bits"]).all()
    dec.close()


@pytest.mark.parametrize("name", ["wifi", "a5", "c79"])
def test_guard_boundary_values(fp, po, name):
    """Channel values and messages right at the packed kernel's range limits (8100 for LLRs, 2^13 for messages,
    the +-24000 posterior clamp): auto precision must stay bit exact, whichever kernel ends up decoding a frame."""
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    orc = po.Oracle(t)
    rng = np.random.default_rng(13)
    frames = 40
    llr = rng.integers(-8200, 8201, size=(frames, code.n)).astype(np.int32)
    llr[0] = 8099; llr[1] = -8099; llr[2] = 8100; llr[3] = -8100; llr[4] = 8191; llr[5] = -8192
    llr[6] = rng.choice(np.array([8099, -8099, 8100, -8100], np.int32), size=code.n)
    llr[7:20] = rng.integers(-2100, 2101, size=(13, code.n))       # grows past 2^13 while decoding (dv up to 11)
    llr[20:30] = np.abs(rng.integers(1500, 8099, size=(10, code.n)))  # all positive: posteriors pile up -> clamp
    want = _expect(orc, llr, False)
    dec = fp.Decoder(code, precision=0)
    out = dec.decode(llr, want_post=True, want_v2c=True)
    _compare(fp, code, t, out, want)
    st = dec.stats()
    assert 0 < st["fallback_frames"] <= frames
    dec.close()


@pytest.mark.parametrize("name,snr_db,frames,scale", [("wifi", 1.5, 1 << 16, 1), ("a5", 3.8, 1 << 15, 1), ("c79", 3.2, 1 << 15, 1),
                                                      ("a24", 5.6, 1 << 11, 1), ("wifi", 2.0, 1 << 14, 6), ("a5", 4.0, 1 << 13, 5)])
def test_packed_kernel_equals_int32_kernel_at_scale(fp, name, snr_db, frames, scale):
    """The int16x2 kernel (+ guard fallback) against the exact int32 kernel on tens of thousands of frames:
    every iteration count and every decoded bit must agree (the int32 kernel itself is pinned to the oracle above)."""
    code = fp.codes.NAMED[name]()
    rate = fp.codes.INFO_BITS[name] / code.n
    llr = channel_frames(code.n, rate, snr_db, frames, seed=frames + scale) * scale
    d_auto = fp.Decoder(code, precision=0, precheck=(name in ("a5", "a24")))
    d_32 = fp.Decoder(code, precision=32, precheck=(name in ("a5", "a24")))
    a = d_auto.decode(llr)
    b = d_32.decode(llr)
    assert (a["iters"] == b["iters"]).all()
    assert (a["bits"] == b["bits"]).all()
    assert len(np.unique(a["iters"])) > 3
    if scale > 1:
        assert d_auto.stats()["fallback_frames"] > 0
    d_auto.close(); d_32.close()


@pytest.mark.parametrize("name,snr_db,precheck", [("wifi", 1.8, False), ("a5", 4.2, True), ("c79", 4.0, False)])
def test_one_long_launch_equals_many_short_ones(fp, name, snr_db, precheck):
    """A long queue makes every slot take its next frame (and prefetch its channel values) one frame early; a
    short one does not.  40 000 frames in one device launch and the same frames in launches of 1 500 must agree
    frame for frame (iteration counts and decoded bits), pre-check hits included."""
    import torch
    code = fp.codes.NAMED[name]()
    rate = fp.codes.INFO_BITS[name] / code.n
    frames, small = 40000, 1500
    llr = channel_frames(code.n, rate, snr_db, frames, seed=77)
    if precheck:
        llr[::97] = np.abs(llr[::97]) + 1  # all-zero codeword, no channel error, no zero LLR (a zero decides 1): decode_fixpoint returns 0
    dec = fp.Decoder(code, precheck=precheck)
    d_llr = torch.from_numpy(llr.astype(np.int16)).cuda()

    def run(first, count):
        it = torch.full((count,), -7, dtype=torch.int32, device="cuda")
        bits = torch.zeros((count, code.nw32), dtype=torch.int32, device="cuda")
        dec.decode_device(d_llr[first:first + count].data_ptr(), 16, count, it.data_ptr(), bits.data_ptr())
        dec.sync()
        return it.cpu().numpy(), bits.cpu().numpy()

    it_long, bits_long = run(0, frames)
    assert it_long.min() >= 0 and len(np.unique(it_long)) > 3
    if precheck:
        assert (it_long[::97] == 0).all()
    for first in range(0, frames, small * 9):  # a sample of short launches
        it_s, bits_s = run(first, min(small, frames - first))
        assert (it_s == it_long[first:first + len(it_s)]).all()
        assert (bits_s == bits_long[first:first + len(it_s)]).all()
    dec.close()


@pytest.mark.parametrize("name", ["wifi", "a5"])
def test_int16_host_entry_equals_int32_host_entry(fp, po, name):
    """ldpc_decode_batch_i16 (half the host->device bytes) == ldpc_decode_batch on the widened values == the oracle,
    including the chunked host pipeline (more frames than one chunk) and a frame that leaves the packed guard range."""
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    rate = fp.codes.INFO_BITS[name] / code.n
    llr = channel_frames(code.n, rate, 2.0 if name == "wifi" else 4.0, 9000, seed=99)
    llr[7] *= 40         # leaves the 13-bit guard range (clipped to int16): exact int32 re-decode from the int16 input
    llr = np.clip(llr, -32768, 32767)
    dec = fp.Decoder(code, precheck=(name == "a5"))
    a = dec.decode(llr, want_post=True)
    b = dec.decode_i16(llr.astype(np.int16), want_post=True)
    assert (a["iters"] == b["iters"]).all() and (a["bits"] == b["bits"]).all() and (a["post"] == b["post"]).all()
    orc = po.Oracle(t)
    for f in (0, 7, 8999):
        it, bits, post, _ = orc.decode(llr[f], precheck=(name == "a5"))
        assert it == b["iters"][f] and (fp.unpack_bits(b["bits"][f:f + 1], code.n)[0] == bits).all()
        if it > 0:
            assert (post == b["post"][f]).all()
    assert dec.stats()["fallback_frames"] >= 1
    dec.close()


@pytest.mark.parametrize("name,bits16", [("wifi", False), ("wifi", True), ("a5", True)])
def test_fed_host_pipeline_equals_chunked_pipeline(fp, po, name, bits16):
    """Long host batches run as ONE persistent launch fed by chunked copies (arrival mark polled by the kernel, results
    copied out per finished chunk); short ones as chunked launches.  Same results, including frames that leave the packed
    guard range (re-decoded by the int32 kernel after the fed launch, results copied again) and the oracle's."""
    import os
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    rate = fp.codes.INFO_BITS[name] / code.n
    frames = 40000
    llr = channel_frames(code.n, rate, 2.0 if name == "wifi" else 4.0, frames, seed=4242)
    for f in (5, 20000, frames - 1):
        llr[f] = np.clip(llr[f] * 40, -32768, 32767)
    dec = fp.Decoder(code, precheck=(name == "a5"))
    run = (lambda: dec.decode_i16(llr.astype(np.int16))) if bits16 else (lambda: dec.decode(llr))
    os.environ.pop("LDPC_NO_FEED", None)
    a = run()
    os.environ["LDPC_FEED_BATCH"] = "15000"      # several fed launches per call (the default batch is 2^18 frames)
    try:
        c = run()
    finally:
        os.environ.pop("LDPC_FEED_BATCH", None)
    os.environ["LDPC_NO_FEED"] = "1"
    try:
        b = run()
    finally:
        os.environ.pop("LDPC_NO_FEED", None)
    assert (a["iters"] == b["iters"]).all() and (a["bits"] == b["bits"]).all()
    assert (c["iters"] == b["iters"]).all() and (c["bits"] == b["bits"]).all()
    assert dec.stats()["fallback_frames"] >= 3
    orc = po.Oracle(t)
    for f in (0, 5, 1234, 20000, frames - 1):
        it, bits, _, _ = orc.decode(llr[f], precheck=(name == "a5"))
        assert it == a["iters"][f] and (fp.unpack_bits(a["bits"][f:f + 1], code.n)[0] == bits).all()
    dec.close()
