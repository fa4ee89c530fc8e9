"""The C restatement of the floating-point decoder (oracle_decode_general_f64) against the reference's own
FP_Decoder::decode_general (oracle/_ref, where /root/reference exists) and against the committed golden vectors."""
import os

import numpy as np
import pytest

from conftest import ROOT, needs_reference
from oracle import named_codes as nc
from oracle import pyoracle as po

GOLDEN_F64 = os.path.join(ROOT, "tests", "golden", "reference_f64.npz")
TAGS = ["wifi_2p0dB", "wifi_1p0dB", "a5_4p5dB", "a5_3p0dB", "c79_4p0dB"]


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_f64_equals_golden_bit_for_bit(tag):
    """Same libm (glibc) on both sides, same operation order: the restatement reproduces the dumped doubles exactly."""
    g = np.load(GOLDEN_F64)
    name = tag.split("_")[0]
    t = nc.tables(name)
    orc = po.Oracle(t)
    for f, x in enumerate(g[tag + "_llr"]):
        it, bits, post, _ = orc.decode_f64(x)
        assert it == g[tag + "_iters"][f]
        assert (np.packbits(bits.astype(np.uint8)) == g[tag + "_bits"][f]).all()
        assert (post == g[tag + "_post"][f]).all()


@needs_reference
def test_oracle_f64_equals_reference_objects():
    rng = np.random.default_rng(5)
    for name, db in (("wifi", 1.5), ("a5", 4.0)):
        t = nc.tables(name)
        ref = po.Reference(name)
        ref.set_tables(t)
        orc = po.Oracle(t)
        snr = 2 * 10 ** (db / 10) * nc.channel_rate(name)
        for x in 2 * snr * (1 + np.sqrt(1 / snr) * rng.standard_normal((3, t.n))):
            a, b = ref.decode_general(x), orc.decode_f64(x)
            assert a[0] == b[0] and (a[1] == b[1]).all() and (a[2] == b[2]).all()
            mask = t.cdeg[None, :] > np.arange(t.dc_max)[:, None]
            assert (a[3][:t.dc_max, :t.m][mask] == b[3][mask]).all()
    for x, y in ((0.3, -2.0), (5.0, 5.0), (0.0, 1.0), (-1e-3, 40.0), (12.5, -0.25)):
        assert ref.sxor_f64(x, y) == orc.sxor_f64(x, y)
