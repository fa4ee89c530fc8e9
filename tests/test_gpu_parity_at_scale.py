"""Full-state parity at the scale SURVEY.md 8(c) asks for (scripts/parity_at_scale.py): 100 000 frames per code at three
Eb/N0 points (array p47 r24: 10 000), CUDA engine through the C ABI with parity-mode outputs against the CPU oracle on
all host cores; iteration counts, decoded bits, posteriors and final messages compared frame by frame through CRC-32
digests.  About 100 s on the B200 box (16 host cores); LDPC_SCALE_FRAMES shrinks it."""
import os
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
def test_full_state_parity_at_scale(capsys):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))  # (a plain import: the worker processes of its pool import it too)
    import parity_at_scale as mod
    frames = os.environ.get("LDPC_SCALE_FRAMES", "100000")
    rc = mod.main(["--frames", frames])
    out = capsys.readouterr().out
    assert rc == 0 and "TOTAL mismatching frames: 0" in out, out
    with open(os.path.join(ROOT, "gpurun_out", "parity_at_scale_last.txt") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else os.devnull, "w") as fh:
        fh.write(out)
