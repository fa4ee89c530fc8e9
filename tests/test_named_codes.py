"""oracle/named_codes.py (the checker side's pure-numpy tables) == the reference's files == the product's constructors."""
import os

import numpy as np
import pytest

from oracle import named_codes as nc
from oracle import pyoracle as po

REF = "/root/reference"
FILES = {"a5": ("H_array_p47_r5_forward.txt", po.read_alist_a), "a24": ("codes/H_array_p47_r24_forward.txt", po.read_alist_a),
         "wifi": ("H_802.11_IndZero.txt", po.read_alist_a), "c79": ("H2212_316_array_cut79.txt", po.read_format_c)}


def same(a, b):
    return (a.n, a.m) == (b.n, b.m) and (a.cdeg == b.cdeg).all() and (a.vdeg == b.vdeg).all() and \
        (a.clist[:, :a.dc_max] == b.clist[:, :a.dc_max]).all() and (a.vlist[:, :a.dv_max] == b.vlist[:, :a.dv_max]).all()


@pytest.mark.parametrize("name", sorted(FILES))
def test_named_tables_match_reference_files(name):
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    path, reader = FILES[name]
    assert same(nc.tables(name), reader(os.path.join(REF, path)))


@pytest.mark.parametrize("name", sorted(FILES))
def test_named_tables_match_product_constructors(name):
    import fixedpointldpc_b200 as fp
    code = fp.codes.NAMED[name]()
    vdeg, cdeg, vlist, clist = code.tables()
    assert same(nc.tables(name), po.Tables(code.n, code.m, vdeg, cdeg, vlist, clist))
    assert nc.INFO_BITS[name] == fp.codes.INFO_BITS[name]
    if name in ("a5", "a24"):
        assert abs(nc.channel_rate(name) - code.rate) < 1e-15
