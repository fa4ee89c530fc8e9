"""B200-native fixed-point LDPC decode engine (drop-in for the decode path of tyc85/FixedPointLDPC).

The product is `libldpc_b200.so` (hand-written sm_100a CUDA kernels behind the C ABI declared in
include/ldpc_capi.h) plus the C++ facade that re-creates the reference's FP_Decoder / PerfTest
names.  This package is the thin ctypes binding used by the tests and the benchmark.
"""
from .capi import (Code, Decoder, Generator, LdpcError, McGroup, device_count, load_library, unpack_bits,  # noqa: F401
                   FMT_A, FMT_AUTO, FMT_C, STREAM_PHILOX, STREAM_REFERENCE)
from . import codes  # noqa: F401
