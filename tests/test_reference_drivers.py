"""The committed driver transcripts are what the reference's own console program prints (where /root/reference exists)."""
import os
import subprocess

from conftest import ROOT, REFERENCE_DIR, needs_reference


def driver_sections():
    """{argument string: stdout} of oracle/_ref/wrapper_a5 (tests/golden/reference_drivers.txt, make_golden.py)."""
    text = open(os.path.join(ROOT, "tests", "golden", "reference_drivers.txt")).read()
    out = {}
    for block in text.split("# wrapper_a5 ")[1:]:
        head, body = block.split("\n", 1)
        out[head.strip()] = body
    return out


@needs_reference
def test_driver_transcripts_are_the_reference_output(tmp_path):
    exe = os.path.join(ROOT, "oracle", "_ref", "wrapper_a5")
    os.symlink(os.path.join(REFERENCE_DIR, "codes", "G_array_forward.txt"), os.path.join(tmp_path, "G_array_forward.txt"))
    for args, want in driver_sections().items():
        got = subprocess.run([exe] + args.split(), cwd=tmp_path, capture_output=True, text=True, timeout=600).stdout
        assert got == want, args


def test_driver_transcripts_shape():
    sec = driver_sections()
    assert sorted(sec) == ["debug", "shorten 200", "shorten 36", "timetrial 2 300", "timetrial 6 500"]
    assert sec["shorten 36"].endswith("10922 100 100\n FER: 1 BER: 0.0494432\n")
    assert sec["shorten 200"].endswith("10092 100 100\n FER: 1 BER: 0.0456858\n")
    assert sec["timetrial 2 300"].startswith("9000 300 300\n")
