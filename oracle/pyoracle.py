"""ctypes front-end to the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.

  Oracle      -- oracle/_build/liboracle.so, the plain-C restatement (ldpc_oracle.c)
  Reference   -- oracle/_ref/libref_<variant>.so, the reference's own objects compiled
                 unmodified (present only if it was built where /root/reference exists)
  read_alist_a / read_format_c / read_format_b -- numpy readers restating the reference's
                 parsers (ArrayLDPC_Decoder.cpp:642-674, ArrayLDPC_Encoder.cpp:45-83)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class _Code(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("dc_max", C.c_int), ("dv_max", C.c_int),
                ("cdeg", C.c_void_p), ("clist", C.c_void_p), ("vdeg", C.c_void_p), ("vlist", C.c_void_p)]


class _Gen(C.Structure):
    _fields_ = [("n", C.c_int), ("rows", C.c_int), ("stride", C.c_int),
                ("flag", C.c_void_p), ("deg", C.c_void_p), ("mlist", C.c_void_p)]


class Tables:
    """Dense copies of the four ReadH tables (ArrayLDPCMacro.h:170-172); -1 pads short rows."""

    def __init__(self, n, m, vdeg, cdeg, vlist, clist):
        self.n, self.m = int(n), int(m)
        self.vdeg = np.ascontiguousarray(vdeg, dtype=np.int32)
        self.cdeg = np.ascontiguousarray(cdeg, dtype=np.int32)
        self.vlist = np.ascontiguousarray(vlist, dtype=np.int32)
        self.clist = np.ascontiguousarray(clist, dtype=np.int32)
        self.dv_max = self.vlist.shape[1]
        self.dc_max = self.clist.shape[1]
        self.edges = int(self.cdeg.sum())

    def c_struct(self):
        return _Code(self.n, self.m, self.dc_max, self.dv_max, self.cdeg.ctypes.data,
                     self.clist.ctypes.data, self.vdeg.ctypes.data, self.vlist.ctypes.data)


def _ints(path):
    with open(path) as fh:
        return np.array(fh.read().split(), dtype=np.int64)


def _ragged(tok, pos, degs, width):
    out = np.full((len(degs), width), -1, dtype=np.int32)
    for i, d in enumerate(degs):
        out[i, :d] = tok[pos:pos + d]
        pos += d
    return out, pos


def read_alist_a(path):
    """Format A, exactly what FP_Decoder::ReadH consumes (ArrayLDPC_Decoder.cpp:650-671)."""
    tok = _ints(path)
    n, m, dv, dc = (int(x) for x in tok[:4])
    pos = 4
    vdeg = tok[pos:pos + n].astype(np.int32); pos += n
    cdeg = tok[pos:pos + m].astype(np.int32); pos += m
    vlist, pos = _ragged(tok, pos, vdeg, max(dv, int(vdeg.max())))
    clist, pos = _ragged(tok, pos, cdeg, max(dc, int(cdeg.max())))
    return Tables(n, m, vdeg, cdeg, vlist, clist)


def read_format_c(path):
    """Format C (legacy, 1-based check lists only; no reader exists in the reference)."""
    tok = _ints(path)
    n, m = int(tok[0]), int(tok[1])
    cdeg = tok[2:2 + m].astype(np.int32)
    clist, _ = _ragged(tok - 1, 2 + m, cdeg, int(cdeg.max()))
    clist = np.where(clist < 0, -1, clist).astype(np.int32)
    for c in range(m):
        clist[c, :cdeg[c]] = np.sort(clist[c, :cdeg[c]])
    return tables_from_clist(n, cdeg, clist)


def tables_from_clist(n, cdeg, clist):
    m = len(cdeg)
    cols = [[] for _ in range(n)]
    for c in range(m):
        for k in range(cdeg[c]):
            cols[clist[c, k]].append(c)
    vdeg = np.array([len(x) for x in cols], dtype=np.int32)
    vlist = np.full((n, int(vdeg.max())), -1, dtype=np.int32)
    for v, lst in enumerate(cols):
        vlist[v, :len(lst)] = lst
    return Tables(n, m, vdeg, np.asarray(cdeg, dtype=np.int32), vlist, clist)


class Generator:
    def __init__(self, n, rows, flag, deg, mlist):
        self.n, self.rows = int(n), int(rows)
        self.flag = np.ascontiguousarray(flag, dtype=np.int32)
        self.deg = np.ascontiguousarray(deg, dtype=np.int32)
        self.mlist = np.ascontiguousarray(mlist, dtype=np.int32)
        self.info_index = np.flatnonzero(self.flag == 0).astype(np.int32)
        self.parity_index = np.flatnonzero(self.flag != 0).astype(np.int32)
        self.k = self.n - self.rows

    def c_struct(self):
        return _Gen(self.n, self.rows, self.mlist.shape[1], self.flag.ctypes.data,
                    self.deg.ctypes.data, self.mlist.ctypes.data)


def read_format_b(path):
    """Format B, what FP_Encoder::FP_Encoder consumes (ArrayLDPC_Encoder.cpp:45-83)."""
    tok = _ints(path)
    n, rows = int(tok[0]), int(tok[1])
    pos = 4
    flag = tok[pos:pos + n]; pos += n
    deg = tok[pos:pos + rows].astype(np.int32); pos += rows
    mlist, _ = _ragged(tok, pos, deg, int(deg.max()))
    return Generator(n, rows, flag, deg, mlist)


class Oracle:
    """The plain-C restatement."""

    def __init__(self, tables, max_iter=30):
        if not os.path.exists(ORACLE_LIB):
            raise RuntimeError("oracle not built: run python oracle/build_ref.py")
        self.lib = L = C.CDLL(ORACLE_LIB)
        self.t = tables
        self.max_iter = max_iter
        self._code = tables.c_struct() if tables is not None else None
        L.oracle_sxor.restype = C.c_int
        L.oracle_sxor.argtypes = [C.c_int, C.c_int]
        for name in ("oracle_decode_general_fp", "oracle_decode_fixpoint"):
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(_Code), _i32p, C.c_int, _i32p, _i32p, _i32p]
        L.oracle_decode_many.restype = C.c_long
        L.oracle_decode_many.argtypes = [C.POINTER(_Code), _i32p, C.c_long, C.c_int, C.c_int, _i32p]
        L.oracle_random.restype = C.c_double
        L.oracle_random.argtypes = [C.POINTER(C.c_long)]
        L.oracle_normal.restype = C.c_double
        L.oracle_normal.argtypes = [C.POINTER(C.c_long), C.c_double, C.c_double]
        L.oracle_channel_frame.restype = None
        L.oracle_channel_frame.argtypes = [C.POINTER(C.c_long), C.c_void_p, C.c_int, C.c_double,
                                           C.c_double, C.c_int, _i32p]
        L.oracle_encode.restype = None
        L.oracle_encode.argtypes = [C.POINTER(_Gen), C.c_char_p, C.c_int, _i32p]
        L.oracle_set_info_bit.restype = None
        L.oracle_set_info_bit.argtypes = [C.c_char_p, C.c_int, C.c_int, _i32p]
        L.oracle_calculate_ber.restype = C.c_int
        L.oracle_calculate_ber.argtypes = [_i32p, _i32p, _i32p, C.c_int]
        L.oracle_sxor_f64.restype = C.c_double
        L.oracle_sxor_f64.argtypes = [C.c_double, C.c_double]
        L.oracle_decode_general_f64.restype = C.c_int
        L.oracle_decode_general_f64.argtypes = [C.POINTER(_Code), _f64p, C.c_int, _i32p, _f64p, _f64p]
        self.seed = C.c_long(123456789)  # rngs.cpp:45 DEFAULT

    def sxor(self, x, y):
        return self.lib.oracle_sxor(int(x), int(y))

    def sxor_f64(self, x, y):
        return self.lib.oracle_sxor_f64(float(x), float(y))

    def decode_f64(self, llr):
        """decode_general(const double *): returns (iters, bits[n], post[n] float64, edge[dc_max][m] float64)."""
        t = self.t
        llr = np.ascontiguousarray(llr, dtype=np.float64)
        bits, post, edge = np.zeros(t.n, np.int32), np.zeros(t.n, np.float64), np.zeros((t.dc_max, t.m), np.float64)
        it = self.lib.oracle_decode_general_f64(C.byref(self._code), llr, self.max_iter, bits, post, edge.reshape(-1))
        return it, bits, post, edge

    def decode(self, llr, precheck=False, state=None):
        """Returns (iters, bits[n], post[n], edge[dc_max][m]).  `state` = (bits, post, edge)
        buffers to reuse so stale contents survive a pre-check hit, like the reference."""
        t = self.t
        llr = np.ascontiguousarray(llr, dtype=np.int32)
        if state is None:
            state = (np.zeros(t.n, np.int32), np.zeros(t.n, np.int32), np.zeros((t.dc_max, t.m), np.int32))
        bits, post, edge = state
        fn = self.lib.oracle_decode_fixpoint if precheck else self.lib.oracle_decode_general_fp
        it = fn(C.byref(self._code), llr, self.max_iter, bits, post, edge.reshape(-1))
        return it, bits, post, edge

    def decode_many(self, llr, precheck=False):
        llr = np.ascontiguousarray(llr, dtype=np.int32).reshape(-1, self.t.n)
        iters = np.zeros(len(llr), np.int32)
        self.lib.oracle_decode_many(C.byref(self._code), llr.reshape(-1), len(llr), self.max_iter,
                                    int(precheck), iters)
        return iters

    def random(self):
        return self.lib.oracle_random(C.byref(self.seed))

    def normal(self, mean, sd):
        return self.lib.oracle_normal(C.byref(self.seed), mean, sd)

    def channel_frame(self, codeword, n, snr, sigma, frac_width=4):
        out = np.zeros(n, np.int32)
        cw = None
        if codeword is not None:
            cw = np.ascontiguousarray(codeword, dtype=np.int32)
        self.lib.oracle_channel_frame(C.byref(self.seed), cw.ctypes.data if cw is not None else None,
                                      n, snr, sigma, frac_width, out)
        return out

    def encode(self, gen, info_bytes):
        out = np.zeros(gen.n, np.int32)
        g = gen.c_struct()
        self.lib.oracle_encode(C.byref(g), info_bytes, len(info_bytes), out)
        return out

    def set_info_bit(self, info_bytes, k):
        out = np.zeros(k + 8, np.int32)
        self.lib.oracle_set_info_bit(info_bytes, len(info_bytes), k, out)
        return out[:k]

    def calculate_ber(self, bits, info_index, true_info):
        return self.lib.oracle_calculate_ber(np.ascontiguousarray(bits, np.int32),
                                             np.ascontiguousarray(info_index, np.int32),
                                             np.ascontiguousarray(true_info, np.int32), len(true_info))


def reference_available(variant):
    return os.path.exists(os.path.join(REF_DIR, "libref_%s.so" % variant))


class Reference:
    """The reference's own compiled objects for one compile-time code variant."""

    def __init__(self, variant):
        path = os.path.join(REF_DIR, "libref_%s.so" % variant)
        if not os.path.exists(path):
            raise RuntimeError("reference build missing: " + path)
        self.lib = L = C.CDLL(path)
        dims = np.zeros(16, np.int32)
        L.ref_dims(dims.ctypes.data_as(C.c_void_p))
        (self.n, self.m, self.dc, self.dv, self.info_length, self.p, self.ncgrp, self.nvgrp,
         self.max_iter, self.frac_width, self.ram_depth) = (int(x) for x in dims[:11])
        L.ref_decode_general_fp.restype = C.c_int
        L.ref_decode_general_fp.argtypes = [_i32p, _i32p, _i32p, _i32p]
        L.ref_decode_fixpoint.restype = C.c_int
        L.ref_decode_fixpoint.argtypes = [_i32p, C.c_int, _i32p, _i32p, _i32p]
        L.ref_decode_many.restype = C.c_long
        L.ref_decode_many.argtypes = [_i32p, C.c_long, C.c_int, _i32p]
        L.ref_sxor.restype = C.c_int
        L.ref_sxor.argtypes = [C.c_int, C.c_int]
        L.ref_sxor_grid.argtypes = [C.c_int, C.c_int, _i32p]
        L.ref_rate.restype = C.c_double
        L.ref_random.restype = C.c_double
        L.ref_normal.restype = C.c_double
        L.ref_normal.argtypes = [C.c_double, C.c_double]
        L.ref_put_seed.argtypes = [C.c_long]
        L.ref_get_seed.restype = C.c_long
        L.ref_channel_frame.argtypes = [C.c_void_p, C.c_double, C.c_double, _i32p]
        L.ref_set_tables.argtypes = [_i32p, _i32p, _i32p, C.c_int, _i32p, C.c_int]
        L.ref_get_tables.argtypes = [_i32p, _i32p, _i32p, _i32p]
        L.ref_read_h.argtypes = [C.c_char_p]
        L.ref_encoder_open.restype = C.c_int
        L.ref_encoder_open.argtypes = [C.c_char_p]
        L.ref_encode.argtypes = [C.c_char_p, C.c_int, _i32p, _i32p]
        L.ref_set_info.argtypes = [C.c_char_p, C.c_int, _i32p]
        L.ref_calculate_ber.restype = C.c_int
        L.ref_hard_decision.restype = C.c_int
        L.ref_hard_decision.argtypes = [_i32p]
        if hasattr(L, "ref_decode_general"):
            L.ref_decode_general.restype = C.c_int
            L.ref_decode_general.argtypes = [_f64p, _i32p, _f64p, _f64p]
            L.ref_sxor_f64.restype = C.c_double
            L.ref_sxor_f64.argtypes = [C.c_double, C.c_double]

    def set_tables(self, t):
        assert (t.n, t.m) == (self.n, self.m) and t.dc_max <= self.dc and t.dv_max <= self.dv
        self.lib.ref_set_tables(t.vdeg, t.cdeg, t.vlist.reshape(-1), t.vlist.shape[1],
                                t.clist.reshape(-1), t.clist.shape[1])

    def read_h(self, directory):
        self.lib.ref_read_h(directory.encode())

    def get_tables(self):
        vdeg = np.zeros(self.n, np.int32); cdeg = np.zeros(self.m, np.int32)
        vlist = np.zeros((self.n, self.dv), np.int32); clist = np.zeros((self.m, self.dc), np.int32)
        self.lib.ref_get_tables(vdeg, cdeg, vlist.reshape(-1), clist.reshape(-1))
        return Tables(self.n, self.m, vdeg, cdeg, vlist, clist)

    def _bufs(self):
        return (np.zeros(self.n, np.int32), np.zeros(self.n, np.int32),
                np.zeros((self.dc, self.ram_depth), np.int32))

    def decode_general_fp(self, llr):
        bits, post, edge = self._bufs()
        it = self.lib.ref_decode_general_fp(np.ascontiguousarray(llr, np.int32), bits, post, edge.reshape(-1))
        return it, bits, post, edge

    def decode_fixpoint(self, llr, set_pcv=True):
        bits, post, edge = self._bufs()
        it = self.lib.ref_decode_fixpoint(np.ascontiguousarray(llr, np.int32), int(set_pcv), bits, post,
                                          edge.reshape(-1))
        return it, bits, post, edge

    def decode_general(self, llr):
        """FP_Decoder::decode_general(const double *): (iters, bits, post float64, edge float64 [dc][ram_depth])."""
        bits, post = np.zeros(self.n, np.int32), np.zeros(self.n, np.float64)
        edge = np.zeros((self.dc, self.ram_depth), np.float64)
        it = self.lib.ref_decode_general(np.ascontiguousarray(llr, np.float64), bits, post, edge.reshape(-1))
        return it, bits, post, edge

    def sxor_f64(self, x, y):
        return self.lib.ref_sxor_f64(float(x), float(y))

    def decode_many(self, llr, fixpoint):
        llr = np.ascontiguousarray(llr, np.int32).reshape(-1, self.n)
        iters = np.zeros(len(llr), np.int32)
        self.lib.ref_decode_many(llr.reshape(-1), len(llr), int(fixpoint), iters)
        return iters

    def sxor(self, x, y):
        return self.lib.ref_sxor(int(x), int(y))

    def sxor_grid(self, lo, hi):
        out = np.zeros((hi - lo + 1) ** 2, np.int32)
        self.lib.ref_sxor_grid(lo, hi, out)
        return out.reshape(hi - lo + 1, hi - lo + 1)

    def rate(self):
        return self.lib.ref_rate()

    def random(self):
        return self.lib.ref_random()

    def normal(self, mean, sd):
        return self.lib.ref_normal(mean, sd)

    def put_seed(self, x):
        self.lib.ref_put_seed(x)

    def get_seed(self):
        return self.lib.ref_get_seed()

    def channel_frame(self, codeword, snr, sigma):
        out = np.zeros(self.n, np.int32)
        cw = np.ascontiguousarray(codeword, np.int32) if codeword is not None else None
        self.lib.ref_channel_frame(cw.ctypes.data if cw is not None else None, snr, sigma, out)
        return out

    def encoder_open(self, path):
        return self.lib.ref_encoder_open(path.encode())

    def encode(self, info_bytes):
        cw = np.zeros(self.n, np.int32); idx = np.zeros(self.info_length, np.int32)
        self.lib.ref_encode(info_bytes, len(info_bytes), cw, idx)
        return cw, idx

    def set_info(self, info_bytes, info_index):
        self.lib.ref_set_info(info_bytes, len(info_bytes), np.ascontiguousarray(info_index, np.int32))

    def calculate_ber(self):
        return self.lib.ref_calculate_ber()

    def hard_decision(self, llr):
        return self.lib.ref_hard_decision(np.ascontiguousarray(llr, np.int32))


def fnv1a64(arr):
    """FNV-1a 64 over the little-endian int32 bytes of arr (digest used in SURVEY.md 8(c))."""
    h = 0xcbf29ce484222325
    for b in np.ascontiguousarray(arr, dtype="<i4").tobytes():
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h
