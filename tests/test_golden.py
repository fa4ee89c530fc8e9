"""The C oracle against the committed golden vectors (tests/golden/reference_vectors.npz, produced
from the reference's own objects by tests/golden/make_golden.py).  CPU only; runs on the GPU box too."""
import numpy as np
import pytest

from conftest import tables_of, valid_mask

BLOCKS = [("wifi", "wifi_2dB", False), ("wifi", "wifi_0p5dB", False), ("a5", "a5_4p5dB", True), ("a5", "a5_2dB", True),
          ("a24", "a24_3dB", True), ("a24", "a24_6dB", True), ("c79", "c79_2dB", False), ("c79", "c79_4p5dB", False)]


def test_sxor_golden(golden, po):
    orc = po.Oracle(None)
    grid = golden["sxor_grid_m300_300"]
    for x in range(-300, 301, 7):
        for y in range(-300, 301):
            assert orc.sxor(x, y) == grid[x + 300, y + 300]
    for x, y, r in golden["sxor_big"]:
        assert orc.sxor(int(x), int(y)) == r
    # spot values and the non-associativity witness recorded in SURVEY.md 8(c)
    for x, y, r in [(5, -3, -1), (0, 7, 0), (1, 1, 1), (2, 2, 1), (16, 16, 8), (40, 3, 2), (100, 27, 27), (-20, 20, -10),
                    (128, 128, 128), (200, 60, 69), (255, 1, 11), (257, 1, 1), (300, 10, 8), (1000, -1000, -990),
                    (4770, -4000, -3990)]:
        assert orc.sxor(x, y) == r
    assert orc.sxor(orc.sxor(9, 14), 20) == 2 and orc.sxor(9, orc.sxor(14, 20)) == 3


def test_rng_golden(golden, po):
    orc = po.Oracle(None)
    orc.seed.value = 123456789
    assert np.array_equal(np.array([orc.random() for _ in range(1000)]), golden["random_first"])
    orc.seed.value = 123456789
    assert np.array_equal(np.array([orc.normal(0.0, 1.0) for _ in range(1000)]), golden["normal_first"])
    orc.seed.value = 1
    for _ in range(10000):
        orc.random()
    assert orc.seed.value == 399268537 == int(golden["seed_after_10000_from_1"][0])  # rngs.cpp:42 CHECK


@pytest.mark.parametrize("name,tag,precheck", BLOCKS)
def test_decode_golden(golden, po, fp, name, tag, precheck):
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    orc = po.Oracle(t)
    mask = valid_mask(t)
    llr = golden[tag + "_llr"].astype(np.int32)
    bits = np.unpackbits(golden[tag + "_bits"], axis=1)[:, :code.n]
    for f in range(len(llr)):
        it, b, post, edge = orc.decode(llr[f], precheck=precheck)
        assert it == golden[tag + "_iters"][f]
        assert (b == bits[f]).all()
        if it > 0:
            assert (post == golden[tag + "_post"][f]).all()
            assert (edge[mask] == golden[tag + "_edge"][f].astype(np.int32)[mask]).all()


def test_precheck_golden(golden, po, fp):
    code = fp.codes.array_p47_r5()
    orc = po.Oracle(tables_of(code))
    llr = golden["a5_9dB_llr"].astype(np.int32)
    want = golden["a5_9dB_iters"]
    assert (want == 0).sum() >= 2
    bits = np.unpackbits(golden["a5_9dB_bits"], axis=1)[:, :code.n]
    for f in range(len(llr)):
        it, b, _, _ = orc.decode(llr[f], precheck=True)
        assert it == want[f] and (b == bits[f]).all()


def test_wifi_channel_and_ber_golden(golden, po, fp):
    """Debug_Wifi flow (PerfTest.cpp:97-135): channel from the default seed, decode, calculateBER."""
    code = fp.codes.wifi_1944_r12()
    orc = po.Oracle(tables_of(code))
    cw = golden["wifi_codeword"].astype(np.int32)
    idx = golden["wifi_info_index"].astype(np.int32)
    snr = 2 * 10 ** (2.0 / 10) * 0.5
    orc.seed.value = 123456789
    for f in range(8):
        llr = orc.channel_frame(cw, code.n, snr, np.sqrt(1 / snr))
        assert (llr == golden["wifi_2dB_llr"][f]).all()
        it, bits, _, _ = orc.decode(llr)
        assert it == golden["wifi_2dB_iters"][f]
        assert orc.calculate_ber(bits, idx, cw[idx]) == golden["wifi_2dB_ber"][f]
    assert list(golden["wifi_2dB_iters"][:8]) == [12, 9, 14, 10, 9, 6, 10, 7]
    assert list(golden["a5_4p5dB_iters"][:8]) == [5, 4, 8, 4, 10, 5, 4, 4]
    assert list(golden["wifi_2dB_llr"][0][:8]) == [-115, -95, -28, -71, 94, 6, -33, -40]
    assert list(golden["a5_4p5dB_llr"][0][:8]) == [-277, 81, 201, 124, -84, -240, 192, -143]
