"""Monte-Carlo point through the multi-GPU entry of the C ABI (ldpc_mc_group_*: one host thread per device inside the
library, counters all-reduced over NCCL once per round, per-frame results only in the final round)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WIFI_SNR_2DB = 2 * 10 ** (2.0 / 10) * 0.5


def _decoders(fp, code, count, **kw):
    return [fp.Decoder(code, device=d, **kw) for d in range(count)]


def test_group_reproduces_wifi_results_transcript(fp, golden):
    """`2732 100 393214` (wifi_results_4_4_2dB_30iter.txt) through ldpc_mc_group_run on every GPU count the box offers
    (1, and 2 when there are two): the counters do not depend on the number of devices or the round size."""
    code = fp.codes.wifi_1944_r12()
    ndev = fp.device_count()
    for count, per_round in [(1, 1 << 16)] + ([(2, 1 << 15), (2, 50000)] if ndev >= 2 else []):
        decs = _decoders(fp, code, count)
        grp = fp.McGroup(decs)
        res = grp.run(WIFI_SNR_2DB, target_block_errors=100, frames_per_round=per_round, stream=fp.STREAM_REFERENCE,
                      seed=123456789, codeword=golden["wifi_codeword"], info_index=golden["wifi_info_index"])
        assert (res["errors"], res["block_errors"], res["frames"], res["reached"]) == (2732, 100, 393214, 1), (count, res)
        assert res["devices"] == count and res["iter_hist"].sum() == 393214
        assert (res["iter_hist"] * np.arange(32)).sum() == res["iter_sum"]
        grp.close()
        for d in decs:
            d.close()


def test_group_rules_and_iteration_log(fp, golden):
    """max_frames rule, the PerfTest iteration-count rule (quirk Q9: `3000 100 100` at 2 dB) and the per-frame log."""
    a5 = fp.codes.array_p47_r5()
    ndev = min(2, fp.device_count())
    decs = _decoders(fp, a5, ndev, precheck=True)
    grp = fp.McGroup(decs)
    snr = 2 * 10 ** (2.0 / 10) * a5.rate
    res = grp.run(snr, target_block_errors=100, count_iterations=True, frames_per_round=64, stream=fp.STREAM_REFERENCE,
                  seed=123456789, want_iters=200)
    assert (res["errors"], res["block_errors"], res["frames"]) == (3000, 100, 100)   # ArrayLDPC_PerfTest(2 dB)
    assert (res["iters"] == 30).all() and len(res["iters"]) == 100
    # ArrayLDPC_Debug flow at 4.5 dB, first 24 frames == the golden iteration list, whatever the round size
    snr = 2 * 10 ** (4.5 / 10) * a5.rate
    res = grp.run(snr, target_block_errors=0, max_frames=24, frames_per_round=5, stream=fp.STREAM_REFERENCE, seed=123456789,
                  codeword=golden["a5_codeword"], info_index=golden["a5_info_index"], want_iters=24)
    assert res["frames"] == 24 and (res["iters"] == golden["a5_4p5dB_iters"]).all()
    assert res["iter_sum"] == golden["a5_4p5dB_iters"].sum()
    single = decs[0].mc_run(24, snr, stream=fp.STREAM_REFERENCE, seed=123456789, codeword=golden["a5_codeword"],
                            info_index=golden["a5_info_index"])
    assert res["errors"] == single["bit_errors"] and res["block_errors"] == single["frame_errors"]
    grp.close()
    for d in decs:
        d.close()


def test_group_rejects_bad_arguments(fp):
    code = fp.codes.cut79()
    d = fp.Decoder(code)
    with pytest.raises(fp.LdpcError):
        fp.McGroup([d, d])   # two decoders on one device
    grp = fp.McGroup([d])
    with pytest.raises(fp.LdpcError):
        grp.run(1.0, target_block_errors=0, max_frames=0)   # no stopping rule
    grp.close()
    d.close()
