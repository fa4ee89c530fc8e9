"""Pins the C oracle (oracle/ldpc_oracle.c) against the reference's own compiled objects.

Runs wherever oracle/_ref/libref_<variant>.so exists (built from /root/reference by
oracle/build_ref.py; the binaries travel to the GPU box).  CPU only."""
import os

import numpy as np
import pytest

from conftest import REFERENCE_DIR, needs_reference, valid_mask

VARIANTS = ("wifi", "a5", "a24", "c79")


def _ref(po, name):
    if not po.reference_available(name):
        pytest.skip("oracle/_ref/libref_%s.so not built" % name)
    return po.Reference(name)


def test_sxor_matches_reference_grid(po):
    ref = _ref(po, "wifi")
    orc = po.Oracle(None)
    grid = ref.sxor_grid(-140, 140)
    mine = np.array([[orc.sxor(x, y) for y in range(-140, 141)] for x in range(-140, 141)], np.int32)
    assert (grid == mine).all()
    rng = np.random.default_rng(1)
    for x, y in rng.integers(-2 ** 24, 2 ** 24, size=(4000, 2)):
        assert ref.sxor(int(x), int(y)) == orc.sxor(int(x), int(y))


def test_rng_and_normal_streams(po):
    ref = _ref(po, "wifi")
    orc = po.Oracle(None)
    ref.put_seed(123456789)
    orc.seed.value = 123456789
    assert [ref.random() for _ in range(500)] == [orc.random() for _ in range(500)]
    assert [ref.normal(0.5, 2.0) for _ in range(500)] == [orc.normal(0.5, 2.0) for _ in range(500)]
    assert ref.get_seed() == orc.seed.value


@pytest.mark.parametrize("name,snr_db,frames", [("wifi", 2.0, 30), ("wifi", 0.5, 6), ("a5", 4.5, 30), ("a5", 2.0, 4),
                                                ("a24", 6.0, 5), ("a24", 3.0, 2), ("c79", 4.5, 20), ("c79", 2.0, 4)])
def test_decode_general_fp_state_matches_reference(po, fp, name, snr_db, frames):
    ref = _ref(po, name)
    code = fp.codes.NAMED[name]()
    from conftest import tables_of
    t = tables_of(code)
    ref.set_tables(t)
    orc = po.Oracle(t)
    rate = fp.codes.INFO_BITS[name] / code.n
    snr = 2 * 10 ** (snr_db / 10) * rate
    ref.put_seed(4242)
    mask = valid_mask(t)
    for _ in range(frames):
        llr = ref.channel_frame(None, snr, np.sqrt(1 / snr))
        it, bits, post, edge = ref.decode_general_fp(llr)
        it2, bits2, post2, edge2 = orc.decode(llr)
        assert it == it2
        assert (bits == bits2).all() and (post == post2).all()
        assert (edge[:t.dc_max][mask] == edge2[mask]).all()
        if name in ("a5", "a24"):  # decode_fixpoint == decode_general_fp on the forward array H
            it3, bits3, post3, edge3 = ref.decode_fixpoint(llr)
            it4, bits4, post4, edge4 = orc.decode(llr, precheck=True)
            assert it3 == it4 == it and (bits3 == bits4).all()
            assert (post3 == post4).all() and (edge3[:t.dc_max][mask] == edge4[mask]).all()


def test_precheck_returns_zero_and_keeps_stale_state(po, fp):
    ref = _ref(po, "a5")
    code = fp.codes.array_p47_r5()
    from conftest import tables_of
    t = tables_of(code)
    ref.set_tables(t)
    orc = po.Oracle(t)
    state = None
    ref.put_seed(7)
    snr = 2 * 10 ** (9.0 / 10) * code.rate
    hits = 0
    for i in range(12):
        # alternate a noisy frame (decodes) and a clean one (pre-check hit)
        s = snr if i % 2 else 2 * 10 ** (3.5 / 10) * code.rate
        llr = ref.channel_frame(None, s, np.sqrt(1 / s))
        it, bits, post, edge = ref.decode_fixpoint(llr)
        it2, bits2, post2, edge2 = orc.decode(llr, precheck=True, state=state)
        state = (bits2, post2, edge2)
        assert it == it2 and (bits == bits2).all() and (post == post2).all() and (edge[:t.dc_max] == edge2).all()
        hits += it == 0
        assert ref.hard_decision(llr) == (0 if it == 0 else 1)
    assert hits >= 3


@needs_reference
def test_channel_encoder_and_ber_match_reference(po):
    ref = _ref(po, "wifi")
    ref.read_h(REFERENCE_DIR)
    t = ref.get_tables()
    t2 = po.read_alist_a(os.path.join(REFERENCE_DIR, "H_802.11_IndZero.txt"))
    assert (t.vlist == t2.vlist).all() and (t.clist == t2.clist).all() and (t.vdeg == t2.vdeg).all() and (t.cdeg == t2.cdeg).all()
    gen = po.read_format_b(os.path.join(REFERENCE_DIR, "H_802.11_IndZerog.txt"))
    assert ref.encoder_open(os.path.join(REFERENCE_DIR, "H_802.11_IndZerog.txt")) == 0
    orc = po.Oracle(t)
    msg = bytes(np.random.default_rng(3).integers(0, 256, 122, dtype=np.uint8))
    cw, idx = ref.encode(msg)
    assert (cw == orc.encode(gen, msg)).all() and (idx == gen.info_index).all()
    ref.set_info(msg, idx)
    true_info = orc.set_info_bit(msg, 972)
    ref.put_seed(99)
    orc.seed.value = 99
    snr = 2 * 10 ** (1.6 / 10) * 0.5
    for _ in range(10):
        a = ref.channel_frame(cw, snr, np.sqrt(1 / snr))
        b = orc.channel_frame(cw, 1944, snr, np.sqrt(1 / snr))
        assert (a == b).all()
        _, bits, _, _ = ref.decode_general_fp(a)
        assert ref.calculate_ber() == orc.calculate_ber(bits, idx, true_info)


@needs_reference
def test_array_generator_matches_reference(po):
    ref = _ref(po, "a5")
    path = os.path.join(REFERENCE_DIR, "codes", "G_array_forward.txt")
    gen = po.read_format_b(path)
    assert ref.encoder_open(path) == 0
    orc = po.Oracle(None)
    msg = bytes(np.random.default_rng(5).integers(0, 256, 248, dtype=np.uint8))
    cw, idx = ref.encode(msg)
    assert (cw == orc.encode(gen, msg)).all() and (idx == gen.info_index).all()
    # the codeword satisfies the forward H
    t = po.read_alist_a(os.path.join(REFERENCE_DIR, "H_array_p47_r5_forward.txt"))
    syn = [int(cw[t.clist[c, :t.cdeg[c]]].sum() & 1) for c in range(t.m)]
    assert sum(syn) == 0
