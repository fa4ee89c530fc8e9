"""Parity-check tables of the four named codes, in pure numpy.  TEST INFRASTRUCTURE ONLY.

The reference arm of bench.py and the CPU checkers need the codes without touching the product library
(libldpc_b200.so) and without /root/reference (absent on the GPU box), so the constructions are restated here:

  a5, a24  class ROM (ArrayLDPCMacro.h:42-82) / codes/alist_from_arraycode.m:8: check a*p + t of row group a touches
           variable b*p + ((t + a*b) mod p) of every column group b ("forward" shift, the one decode_fixpoint
           addresses, ArrayLDPC_Decoder.cpp:474-476)
  c79      H2212_316_array_cut79.txt: p = 79, row multipliers 0,1,3,4, 28 selected column groups, backward shift
           
  wifi     H_802.11_IndZero.txt = the IEEE 802.11n n=1944 rate-1/2 Z=81 prototype matrix (first base row
           57 - - - 50 - 11 - 50 - 79 - 1 0 ...)

tests/test_named_codes.py checks every table against the reference's own file (where /root/reference exists) and
against the product's constructors.
"""
import numpy as np

from .pyoracle import Tables, tables_from_clist

CUT79_ROWS = (0, 1, 3, 4)
CUT79_COLS = (2, 6, 7, 14, 17, 18, 22, 26, 27, 30, 36, 37, 38, 46, 47, 49, 55, 56, 57, 58, 61, 62, 65, 66, 67,
              76, 77, 78)
_ = -1
WIFI_1944_R12 = (
    (57, _, _, _, 50, _, 11, _, 50, _, 79, _, 1, 0, _, _, _, _, _, _, _, _, _, _),
    (3, _, 28, _, 0, _, _, _, 55, 7, _, _, _, 0, 0, _, _, _, _, _, _, _, _, _),
    (30, _, _, _, 24, 37, _, _, 56, 14, _, _, _, _, 0, 0, _, _, _, _, _, _, _, _),
    (62, 53, _, _, 53, _, _, 3, 35, _, _, _, _, _, _, 0, 0, _, _, _, _, _, _, _),
    (40, _, _, 20, 66, _, _, 22, 28, _, _, _, _, _, _, _, 0, 0, _, _, _, _, _, _),
    (0, _, _, _, 8, _, 42, _, 50, _, _, 8, _, _, _, _, _, 0, 0, _, _, _, _, _),
    (69, 79, 79, _, _, _, 56, _, 52, _, _, _, 0, _, _, _, _, _, 0, 0, _, _, _, _),
    (65, _, _, _, 38, 57, _, _, 72, _, 27, _, _, _, _, _, _, _, _, 0, 0, _, _, _),
    (64, _, _, _, 14, 52, _, _, 30, _, _, 32, _, _, _, _, _, _, _, _, 0, 0, _, _),
    (_, 45, _, 70, 0, _, _, _, 77, 9, _, _, _, _, _, _, _, _, _, _, _, 0, 0, _),
    (2, 56, _, 57, 35, _, _, _, _, _, 12, _, _, _, _, _, _, _, _, _, _, _, 0, 0),
    (24, _, 61, _, 60, _, _, 27, 51, _, _, 16, 1, _, _, _, _, _, _, _, _, _, _, 0),
)
# k = n - rank(H): converts frames/s into info bit/s
INFO_BITS = {"a5": 1978, "a24": 1104, "c79": 1899, "wifi": 972}


def _from_rows(n, rows):
    dc = max(len(r) for r in rows)
    clist = np.full((len(rows), dc), -1, np.int32)
    for c, row in enumerate(rows):
        clist[c, :len(row)] = sorted(row)
    return tables_from_clist(n, np.array([len(r) for r in rows], np.int32), clist)


def array_code(p, row_mult, col_sel, backward=False):
    sign = -1 if backward else 1
    rows = [[b * p + (t + sign * a * g) % p for b, g in enumerate(col_sel)] for a in row_mult for t in range(p)]
    return _from_rows(p * len(col_sel), rows)


def qc_code(proto, z):
    rows = [[j * z + (t + s) % z for j, s in enumerate(prow) if s >= 0] for prow in proto for t in range(z)]
    return _from_rows(len(proto[0]) * z, rows)


def tables(name) -> Tables:
    if name == "a5":
        return array_code(47, range(5), range(47))
    if name == "a24":
        return array_code(47, range(24), range(47))
    if name == "c79":
        return array_code(79, CUT79_ROWS, CUT79_COLS, backward=True)
    if name == "wifi":
        return qc_code(WIFI_1944_R12, 81)
    raise KeyError(name)


def channel_rate(name):
    """The `Rate` the reference's drivers put into snr = 2*10^(dB/10)*Rate: 0.5 hard-coded for 802.11
    (PerfTest.cpp:62), ROM::getRate = 1 - (r*p - r + 1)/p^2 for the array codes (ArrayLDPCMacro.h:60,
    PerfTest.cpp:252,486), k/n for the cut code (no driver exists for it)."""
    if name == "wifi":
        return 0.5
    if name == "c79":
        return INFO_BITS["c79"] / 2212.0
    r = {"a5": 5, "a24": 24}[name]
    return 1.0 - (r * 47 - r + 1) / (47.0 * 47.0)
