"""The reference-compatible C++ facade (include/ArrayLDPCMacro.h, PerfTest.h) and console program on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, tables_of, valid_mask

pytestmark = pytest.mark.gpu

PKG = os.path.join(ROOT, "fixedpointldpc_b200")
WIFI_MSG = b"OMG  how long   dd   should this string be to make it 243".ljust(122, b"\0")


def _dump_case(tmp, fp, golden, name, tags):
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    code.save(os.path.join(tmp, "H.txt"))
    llr = np.concatenate([golden[tag + "_llr"].astype(np.int32) for tag in tags])
    iters = np.concatenate([golden[tag + "_iters"] for tag in tags]).astype(np.int32)
    bits = np.concatenate([np.unpackbits(golden[tag + "_bits"], axis=1)[:, :code.n] for tag in tags]).astype(np.uint8)
    # posteriors / EdgeRAM images: golden where the reference dump has them, otherwise from the (pinned) C oracle
    from oracle import pyoracle
    orc = pyoracle.Oracle(t)
    mask = valid_mask(t)
    post = np.zeros((len(llr), code.n), np.int32)
    edge = np.zeros((len(llr), t.dc_max, t.m), np.int32)
    for f in range(len(llr)):
        it, _, pp, ee = orc.decode(llr[f])
        if iters[f] > 0:
            assert it == iters[f]
        post[f] = pp
        edge[f] = np.where(mask, ee, 0)
    for arr, fname in ((llr, "llr.bin"), (iters, "iters.bin"), (bits, "bits.bin"), (post, "post.bin"), (edge, "edge.bin"),
                       (t.cdeg.astype(np.int32), "cdeg.bin")):
        np.ascontiguousarray(arr).tofile(os.path.join(tmp, fname))


def _build_check(tmp, variant):
    exe = os.path.join(tmp, "facade_check")
    cmd = ["g++", "-O1", "-std=c++17", "-DLDPC_CODE_VARIANT=%d" % variant, "-I", os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "facade_check.cpp"), "-L", PKG, "-lldpc_b200", "-Wl,-rpath," + PKG]
    subprocess.run(cmd, check=True)
    return exe


@pytest.mark.parametrize("name,variant,mode,tags", [
    ("wifi", 0, "general", ["wifi_2dB", "wifi_0p5dB"]),
    ("a5", 1, "fixpoint", ["a5_4p5dB", "a5_9dB", "a5_2dB", "a5_9dB"]),
    ("a5", 1, "general", ["a5_4p5dB"]),
    ("c79", 3, "general", ["c79_4p5dB", "c79_2dB"]),
    ("a24", 2, "fixpoint", ["a24_6dB", "a24_3dB"]),
    ("a24", 2, "general", ["a24_6dB"]),
])
def test_fp_decoder_class_frame_by_frame(tmp_path, fp, golden, name, variant, mode, tags):
    tmp = str(tmp_path)
    _dump_case(tmp, fp, golden, name, tags)
    exe = _build_check(tmp, variant)
    res = subprocess.run([exe, tmp, mode], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "mismatches 0" in res.stdout


@pytest.mark.parametrize("name,variant,tags", [("wifi", 0, ["wifi_2p0dB", "wifi_1p0dB"]), ("a5", 1, ["a5_4p5dB", "a5_3p0dB"])])
def test_fp_decoder_class_floating_point_decoder(tmp_path, fp, name, variant, tags):
    """decode_general(const double *) of the facade class against the reference's golden doubles."""
    tmp = str(tmp_path)
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_f64.npz"))
    code = fp.codes.NAMED[name]()
    code.save(os.path.join(tmp, "H.txt"))
    np.concatenate([g[t + "_llr"] for t in tags]).astype(np.float64).tofile(os.path.join(tmp, "llr64.bin"))
    np.concatenate([g[t + "_post"] for t in tags]).astype(np.float64).tofile(os.path.join(tmp, "post64.bin"))
    np.concatenate([g[t + "_iters"] for t in tags]).astype(np.int32).tofile(os.path.join(tmp, "iters.bin"))
    np.concatenate([np.unpackbits(g[t + "_bits"], axis=1)[:, :code.n] for t in tags]).astype(np.uint8).tofile(os.path.join(tmp, "bits.bin"))
    exe = _build_check(tmp, variant)
    res = subprocess.run([exe, tmp, "double"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "mismatches 0" in res.stdout


def test_facade_per_frame_latency(tmp_path, fp, golden):
    """The reference's drivers decode one frame per call; this is what such a caller pays per call through the facade
    (one engine call for decode_fixpoint: pre-check and decode in the same launch).  Recorded, bounded loosely."""
    tmp = str(tmp_path)
    _dump_case(tmp, fp, golden, "a5", ["a5_4p5dB"])
    exe = _build_check(tmp, 1)
    res = subprocess.run([exe, tmp, "latency"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    print(res.stdout)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "facade_latency.txt"), "w") as fh:
            fh.write(res.stdout)
    us = [float(l.split(":")[1].split()[0]) for l in res.stdout.strip().split("\n")]
    assert len(us) == 2 and max(us) < 5000.0


def _wifi_files(tmp, fp, golden):
    code = fp.codes.wifi_1944_r12()
    code.save(os.path.join(tmp, "H_802.11_IndZero.txt"))
    parity = np.setdiff1d(np.arange(code.n), golden["wifi_info_index"].astype(np.int64)).astype(np.int32)
    gen = fp.Generator(code=code, parity_cols=parity)
    assert (gen.encode(WIFI_MSG) == golden["wifi_codeword"]).all()
    gen.save(os.path.join(tmp, "H_802.11_IndZerog.txt"))


def test_console_program_reproduces_wifi_results(tmp_path, fp, golden):
    """`echo 2 | wrapper` == wifi_results_4_4_2dB_30iter.txt (the reference's only published output)."""
    tmp = str(tmp_path)
    _wifi_files(tmp, fp, golden)
    res = subprocess.run([os.path.join(PKG, "ldpc_wrapper_wifi")], input="2\n", capture_output=True, text=True, cwd=tmp,
                         timeout=600)
    assert res.returncode == 0, res.stderr
    assert res.stdout == "EbNo in dB? SNR is 2 dB\n2732 100 393214\n FER: 0.000254314 BER: 3.57401e-06\n"


def test_console_program_array_perftest(tmp_path, fp, golden):
    """ArrayLDPC_PerfTest(2 dB) -> `3000 100 100`, (6 dB) -> `142 100 100`, and the two output
    files are created empty like the reference does."""
    tmp = str(tmp_path)
    exe = os.path.join(PKG, "ldpc_wrapper_a5")
    res = subprocess.run([exe, "2", "2", "1", "test.csv"], capture_output=True, text=True, cwd=tmp, timeout=600)
    assert res.stdout == "3000 100 100\n FER: 1 BER: 0.0135808\n", res.stdout + res.stderr
    assert os.path.getsize(os.path.join(tmp, "test.csv")) == 0 and os.path.getsize(os.path.join(tmp, "test.csv_log.txt")) == 0
    res = subprocess.run([exe, "6", "6", "1", "t6.csv"], capture_output=True, text=True, cwd=tmp, timeout=600)
    assert res.stdout.startswith("142 100 100\n"), res.stdout


def test_console_program_array_debug_and_trials(tmp_path, fp, golden):
    tmp = str(tmp_path)
    code = fp.codes.array_p47_r5()
    parity = np.setdiff1d(np.arange(code.n), golden["a5_info_index"].astype(np.int64)).astype(np.int32)
    fp.Generator(code=code, parity_cols=parity).save(os.path.join(tmp, "G_array_forward.txt"))
    exe = os.path.join(PKG, "ldpc_wrapper_a5")
    # every driver's console output against what the reference's own binary prints (tests/golden/reference_drivers.txt,
    # captured from oracle/_ref/wrapper_a5): the reference stream is process-wide state, each run is a fresh process
    from test_reference_drivers import driver_sections
    ref = driver_sections()
    res = subprocess.run([exe, "debug"], capture_output=True, text=True, cwd=tmp, timeout=900)
    assert res.returncode == 0, res.stderr
    # ArrayLDPC_Debug (PerfTest.cpp:217-316): `2515 100 2108`; the reference also dumps the message bits (the debug
    # overload of FP_Encoder::encode), which the facade does not print
    assert res.stdout.split("\n")[0] == "SNR is 7.03061 dB" == ref["debug"].split("\n")[0]
    assert res.stdout.split("\n")[-3:] == ref["debug"].split("\n")[-3:] and "2515 100 2108" in res.stdout
    for args in ("timetrial 2 300", "timetrial 6 500", "shorten 36", "shorten 200"):
        res = subprocess.run([exe] + args.split(), capture_output=True, text=True, cwd=tmp, timeout=900)
        assert res.returncode == 0 and res.stdout == ref[args], (args, res.stdout[-300:], res.stderr[-300:])
    res = subprocess.run([exe, "decodetrial", "4.5", "20000"], capture_output=True, text=True, cwd=tmp, timeout=900)
    assert "bits per second for decoder" in res.stdout and "equivalent SNR is: 7.03061" in res.stdout
    res = subprocess.run([exe, "sweep", "4.0", "5.0", "0.5", "sweep.csv", "50"], capture_output=True, text=True, cwd=tmp,
                         timeout=900)
    rows = open(os.path.join(tmp, "sweep.csv")).read().strip().split("\n")
    assert rows[0].startswith("EbN0_dB,frames") and len(rows) == 4
    log = open(os.path.join(tmp, "sweep.csv_log.txt")).read().strip().split("\n")   # iteration histogram per point
    assert len(log) == 3 and all(sum(int(x) for x in l.split(" iterations ")[1].split()) == int(l.split()[3]) for l in log)
    fers = [float(r.split(",")[4]) for r in rows[1:]]
    assert fers[0] > fers[1] > fers[2] > 0
