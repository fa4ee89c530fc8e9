import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
REFERENCE_DIR = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU known-answer run, opt in with LDPC_SLOW=1")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Compile the product library and the C oracle once per session (seconds; nvcc cross-compiles)."""
    from fixedpointldpc_b200 import build as product_build
    from oracle import build_ref
    product_build.build()
    build_ref.build_oracle()
    build_ref.build_ref()  # no-op without /root/reference


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def fp():
    import fixedpointldpc_b200
    return fixedpointldpc_b200


def tables_of(code):
    from oracle import pyoracle
    vdeg, cdeg, vlist, clist = code.tables()
    return pyoracle.Tables(code.n, code.m, vdeg, cdeg, vlist, clist)


def channel_frames(n, rate, snr_db, count, seed, codeword=None):
    """BPSK/AWGN + the reference's quantiser (PerfTest.cpp:112-119), numpy RNG."""
    rng = np.random.default_rng(seed)
    snr = 2 * 10 ** (snr_db / 10) * rate
    sigma = np.sqrt(1 / snr)
    tx = 1.0 if codeword is None else 1.0 - 2.0 * np.asarray(codeword, np.float64)
    return (2 * snr * (tx + sigma * rng.standard_normal((count, n))) * 16).astype(np.int32)


def valid_mask(t):
    return t.cdeg[None, :] > np.arange(t.dc_max)[:, None]


needs_reference = pytest.mark.skipif(not os.path.isdir(REFERENCE_DIR), reason="/root/reference not present")
