"""Constructors for the codes BASELINE.json names, so nothing needs /root/reference at run time.

The array codes are generated from ROM's circulant-shift rule (ArrayLDPCMacro.h:42-82,
codes/alist_from_arraycode.m); the shortened p=79 code uses the row multipliers / column groups
recovered from H2212_316_array_cut79.txt.  The 802.11n matrix is the standard's
rate-1/2 Z=81 prototype matrix (IEEE 802.11n-2009 Annex R, Table R.3), which is what
H_802.11_IndZero.txt holds.  tests/test_codes.py checks each against the reference's file.
"""
import numpy as np

from .capi import Code

CUT79_ROWS = (0, 1, 3, 4)
CUT79_COLS = (2, 6, 7, 14, 17, 18, 22, 26, 27, 30, 36, 37, 38, 46, 47, 49, 55, 56, 57, 58, 61, 62, 65, 66, 67,
              76, 77, 78)

# IEEE 802.11n, n = 1944, rate 1/2, Z = 81; -1 = zero block, s = identity cyclically shifted by s
WIFI_1944_R12 = (
    (57, -1, -1, -1, 50, -1, 11, -1, 50, -1, 79, -1, 1, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1),
    (3, -1, 28, -1, 0, -1, -1, -1, 55, 7, -1, -1, -1, 0, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1),
    (30, -1, -1, -1, 24, 37, -1, -1, 56, 14, -1, -1, -1, -1, 0, 0, -1, -1, -1, -1, -1, -1, -1, -1),
    (62, 53, -1, -1, 53, -1, -1, 3, 35, -1, -1, -1, -1, -1, -1, 0, 0, -1, -1, -1, -1, -1, -1, -1),
    (40, -1, -1, 20, 66, -1, -1, 22, 28, -1, -1, -1, -1, -1, -1, -1, 0, 0, -1, -1, -1, -1, -1, -1),
    (0, -1, -1, -1, 8, -1, 42, -1, 50, -1, -1, 8, -1, -1, -1, -1, -1, 0, 0, -1, -1, -1, -1, -1),
    (69, 79, 79, -1, -1, -1, 56, -1, 52, -1, -1, -1, 0, -1, -1, -1, -1, -1, 0, 0, -1, -1, -1, -1),
    (65, -1, -1, -1, 38, 57, -1, -1, 72, -1, 27, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, -1, -1, -1),
    (64, -1, -1, -1, 14, 52, -1, -1, 30, -1, -1, 32, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, -1, -1),
    (-1, 45, -1, 70, 0, -1, -1, -1, 77, 9, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, -1),
    (2, 56, -1, 57, 35, -1, -1, -1, -1, -1, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0),
    (24, -1, 61, -1, 60, -1, -1, 27, 51, -1, -1, 16, 1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0),
)


def array_p47_r5():
    """configs[0]: H_array_p47_r5_forward.txt (n=2209, m=235)."""
    return Code.array(47, 5)


def array_p47_r24():
    """configs[3]: codes/H_array_p47_r24_forward.txt (n=2209, m=1128)."""
    return Code.array(47, 24)


def cut79():
    """configs[2]: H2212_316_array_cut79.txt (n=2212, m=316), backward shift."""
    return Code.array(79, 4, ncols=28, row_mult=CUT79_ROWS, col_sel=CUT79_COLS, backward=True)


def qc_checks(proto, z):
    """Check lists of a quasi-cyclic code: block (i,j) with shift s puts a one at row i*z+t,
    column j*z + (t+s) mod z."""
    proto = np.asarray(proto)
    rows = []
    for i in range(proto.shape[0]):
        for t in range(z):
            rows.append(sorted(j * z + (t + int(s)) % z for j, s in enumerate(proto[i]) if s >= 0))
    dc = max(len(r) for r in rows)
    clist = np.full((len(rows), dc), -1, np.int32)
    for r, row in enumerate(rows):
        clist[r, :len(row)] = row
    cdeg = np.array([len(r) for r in rows], np.int32)
    return proto.shape[1] * z, cdeg, clist


def wifi_1944_r12():
    """configs[1]: H_802.11_IndZero.txt (n=1944, m=972)."""
    n, cdeg, clist = qc_checks(WIFI_1944_R12, 81)
    return Code.from_checks(n, cdeg, clist)


NAMED = {"a5": array_p47_r5, "a24": array_p47_r24, "c79": cut79, "wifi": wifi_1944_r12}
# info length k = n - rank(H), used only to convert frames/s into info bit/s
INFO_BITS = {"a5": 1978, "a24": 1104, "c79": 1899, "wifi": 972}
