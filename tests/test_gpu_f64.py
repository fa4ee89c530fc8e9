"""FP_Decoder::decode_general(const double *) on the GPU (ldpc_decode_batch_f64) against the reference's golden vectors
and the C restatement.

Tolerance (floating point): CUDA's log/exp differ from glibc's by at most an ulp or two per call, and the differences
are amplified through up to 30 iterations of message passing.  Bar: iteration counts and decoded bits EQUAL on the
golden frames; posteriors within 1e-9 relative (of max(|value|, 1)) for frames that converge, 1e-6 for frames that run
all 30 iterations without converging (their messages wander chaotically)."""
import os

import numpy as np
import pytest

from conftest import ROOT, tables_of, valid_mask

pytestmark = pytest.mark.gpu
GOLDEN_F64 = os.path.join(ROOT, "tests", "golden", "reference_f64.npz")
TAGS = ["wifi_2p0dB", "wifi_1p0dB", "a5_4p5dB", "a5_3p0dB", "c79_4p0dB"]


def _close(a, b, rel):
    return np.abs(a - b) <= rel * np.maximum(np.abs(b), 1.0)


@pytest.mark.parametrize("tag", TAGS)
def test_f64_decoder_matches_reference_golden(fp, po, tag):
    g = np.load(GOLDEN_F64)
    name = tag.split("_")[0]
    code = fp.codes.NAMED[name]()
    dec = fp.Decoder(code)
    llr = g[tag + "_llr"]
    out = dec.decode_f64(llr, want_post=True, want_v2c=True)
    assert (out["iters"] == g[tag + "_iters"]).all(), (out["iters"], g[tag + "_iters"])
    bits = fp.unpack_bits(out["bits"], code.n)
    want_bits = np.unpackbits(g[tag + "_bits"], axis=1)[:, :code.n]
    t = tables_of(code)
    orc = po.Oracle(t)
    mask = valid_mask(t)
    for f in range(len(llr)):
        rel = 1e-9 if out["iters"][f] < 30 else 1e-6
        assert (bits[f] == want_bits[f]).all()
        assert _close(out["post"][f], g[tag + "_post"][f], rel).all(), np.abs(out["post"][f] - g[tag + "_post"][f]).max()
        _, _, _, edge = orc.decode_f64(llr[f])
        assert _close(out["v2c"][f][mask], edge[mask], rel).all()
    dec.close()


def test_f64_decoder_vs_fixed_point_loss(fp, po):
    """What this decoder is for: on the same noise the fixed-point decoder (FRAC_WIDTH 4, approximate sxor) needs
    somewhat more iterations than the floating-point one and decodes the same bits on frames both converge on."""
    code = fp.codes.array_p47_r5()
    rng = np.random.default_rng(11)
    snr = 2 * 10 ** (4.5 / 10) * code.rate
    llr = 2 * snr * (1 + np.sqrt(1 / snr) * rng.standard_normal((64, code.n)))
    dec = fp.Decoder(code)
    a = dec.decode_f64(llr)
    b = dec.decode((llr * 16).astype(np.int32))
    both = (a["iters"] < 30) & (b["iters"] < 30)
    assert both.sum() >= 50 and (a["bits"][both] == b["bits"][both]).all()
    fa, fb = float(a["iters"][both].mean()), float(b["iters"][both].mean())
    assert fa <= fb < fa + 3.0, (fa, fb)
    # batch == frame by frame, and the facade-sized call (one frame)
    one = dec.decode_f64(llr[3:4])
    assert one["iters"][0] == a["iters"][3] and (one["post"][0] == a["post"][3]).all()
    dec.close()
