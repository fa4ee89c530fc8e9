"""Host side of the product: loaders / constructors, the C-ABI surface, error behaviour.  CPU only."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REFERENCE_DIR, ROOT, needs_reference, tables_of


def _same(code, t):
    vdeg, cdeg, vlist, clist = code.tables()
    return ((code.n, code.m, code.edges) == (t.n, t.m, t.edges) and (vdeg == t.vdeg).all() and (cdeg == t.cdeg).all()
            and (vlist == t.vlist).all() and (clist == t.clist).all())


def test_library_exports_every_declared_symbol(fp):
    header = open(os.path.join(ROOT, "include", "ldpc_capi.h")).read()
    declared = set(re.findall(r"\b(ldpc_[a-z0-9_]+)\s*\(", header))
    declared -= {"ldpc_status"}
    from fixedpointldpc_b200 import capi
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None


def test_named_codes_match_golden_tables(fp, golden):
    vdeg, cdeg, vlist, clist = fp.codes.wifi_1944_r12().tables()
    assert (vdeg == golden["wifi_vdeg"]).all() and (cdeg == golden["wifi_cdeg"]).all()
    assert (vlist == golden["wifi_vlist"]).all() and (clist == golden["wifi_clist"]).all()
    _, cdeg, _, clist = fp.codes.cut79().tables()
    assert (cdeg == golden["c79_cdeg"]).all() and (clist == golden["c79_clist"]).all()
    a5 = fp.codes.array_p47_r5()
    assert (a5.n, a5.m, a5.edges, a5.dc_max, a5.dv_max) == (2209, 235, 11045, 47, 5)
    assert a5.rate == pytest.approx(float(golden["a5_rate"][0]), abs=0)  # ROM::getRate, ArrayLDPCMacro.h:60
    a24 = fp.codes.array_p47_r24()
    assert (a24.n, a24.m, a24.edges) == (2209, 1128, 53016)
    assert a24.rate == pytest.approx(float(golden["a24_rate"][0]), abs=0)


@needs_reference
def test_loaders_and_constructors_match_reference_files(fp, po):
    R = REFERENCE_DIR
    assert _same(fp.codes.array_p47_r5(), po.read_alist_a(R + "/H_array_p47_r5_forward.txt"))
    assert _same(fp.Code.load(R + "/H_array_p47_r5_forward.txt"), po.read_alist_a(R + "/H_array_p47_r5_forward.txt"))
    assert _same(fp.Code.array(47, 5, backward=True), po.read_alist_a(R + "/codes/H_array_p47_r5.txt"))
    assert _same(fp.codes.array_p47_r24(), po.read_alist_a(R + "/codes/H_array_p47_r24_forward.txt"))
    assert _same(fp.codes.cut79(), po.read_format_c(R + "/H2212_316_array_cut79.txt"))
    assert _same(fp.Code.load(R + "/H2212_316_array_cut79.txt"), po.read_format_c(R + "/H2212_316_array_cut79.txt"))
    assert _same(fp.Code.load(R + "/H_array_2209_235_old.txt", fp.FMT_C), po.read_format_c(R + "/H_array_2209_235_old.txt"))
    assert _same(fp.Code.load(R + "/H_802.11_IndZero.txt", fp.FMT_A), po.read_alist_a(R + "/H_802.11_IndZero.txt"))
    assert _same(fp.codes.wifi_1944_r12(), po.read_alist_a(R + "/H_802.11_IndZero.txt"))


def test_save_load_round_trip(fp, tmp_path):
    code = fp.codes.cut79()
    path = str(tmp_path / "h.txt")
    code.save(path)
    again = fp.Code.load(path)
    assert _same(again, tables_of(code))


def test_loader_errors(fp, tmp_path):
    from fixedpointldpc_b200 import capi
    with pytest.raises(fp.LdpcError) as e:
        fp.Code.load(str(tmp_path / "missing.txt"))
    assert e.value.status == capi.ERR_IO
    bad = tmp_path / "bad.txt"
    bad.write_text("4 2\n1 2\n1 1 1 1\n2 2\n0\n0\n1\n1\n0 1\n2 x\n")
    with pytest.raises(fp.LdpcError) as e:
        fp.Code.load(str(bad))
    assert e.value.status == capi.ERR_FORMAT
    short = tmp_path / "short.txt"
    short.write_text("4 2\n1 2\n1 1 1 1\n2 2\n0\n0\n1\n")
    with pytest.raises(fp.LdpcError) as e:
        fp.Code.load(str(short))
    assert e.value.status == capi.ERR_FORMAT
    empty = tmp_path / "empty.txt"
    empty.write_text("")
    with pytest.raises(fp.LdpcError):
        fp.Code.load(str(empty))
    # vlist that disagrees with clist
    incons = tmp_path / "incons.txt"
    incons.write_text("4 2\n1 2\n1 1 1 1\n2 2\n1\n0\n1\n1\n0 1\n2 3\n")
    with pytest.raises(fp.LdpcError) as e:
        fp.Code.load(str(incons), fp.FMT_A)
    assert e.value.status == capi.ERR_FORMAT
    with pytest.raises(fp.LdpcError):  # degree-1 check: the recursion needs d >= 2
        fp.Code.from_checks(3, [1, 2], np.array([[0, -1], [1, 2]], np.int32))
    with pytest.raises(fp.LdpcError):  # variable index out of range
        fp.Code.from_checks(3, [2, 2], np.array([[0, 3], [1, 2]], np.int32))


def test_unsorted_rows_are_sorted(fp):
    code = fp.Code.from_checks(4, [3, 2], np.array([[2, 0, 1], [3, 1, -1]], np.int32))
    _, cdeg, vlist, clist = code.tables()
    assert clist.tolist() == [[0, 1, 2], [1, 3, -1]] and vlist.tolist() == [[0, -1], [0, 1], [0, -1], [1, -1]]


def test_no_cpu_fallback_without_device(fp):
    """On a box without a GPU the decoder must refuse loudly (never decode on the CPU)."""
    from fixedpointldpc_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(fp.LdpcError) as e:
        fp.Decoder(fp.codes.array_p47_r5())
    assert e.value.status == capi.ERR_NO_DEVICE


def test_new_entry_points_reject_null_arguments_without_a_device():
    """Argument checks of the round-2 entries run before any CUDA call (no compute on a CPU-only box)."""
    import ctypes as C
    from fixedpointldpc_b200 import capi
    L = capi.load_library()
    err = C.c_int(0)
    assert not L.ldpc_mc_group_create(None, 0, C.byref(err)) and err.value == capi.ERR_ARG
    assert L.ldpc_mc_group_size(None) == 0
    assert L.ldpc_mc_group_run(None, None, None, None) == capi.ERR_ARG
    assert L.ldpc_mc_run_multi(None, 0, None, None, None) == capi.ERR_ARG
    assert L.ldpc_decode_batch_f64(None, None, 1, None, None, None, None) == capi.ERR_ARG
    assert L.ldpc_decode_batch_i16(None, None, 1, None, None, None, None) == capi.ERR_ARG
    assert L.ldpc_decoder_device(None) == capi.ERR_ARG and not L.ldpc_decoder_code(None)
    L.ldpc_mc_group_destroy(None)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "fixedpointldpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("test infrastructure", ""), f
