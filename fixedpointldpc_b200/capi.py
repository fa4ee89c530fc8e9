"""ctypes binding of libldpc_b200.so (include/ldpc_capi.h).

Thin by design: the product is the CUDA library; Python only moves pointers.  Importing this
module never falls back to a CPU implementation -- if the shared library is missing or no CUDA
device is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LDPC_B200_LIB: load another build of the same library (e.g. the -DLDPC_PHASE_TIMING one of scripts/phase_timing.sh)
LIB_PATH = os.environ.get("LDPC_B200_LIB") or os.path.join(HERE, "libldpc_b200.so")

FMT_AUTO, FMT_A, FMT_C = 0, 1, 3
OK, ERR_IO, ERR_FORMAT, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5, -6, -7

# every symbol include/ldpc_capi.h declares
EXPORTS = (
    "ldpc_strerror", "ldpc_last_error", "ldpc_code_load", "ldpc_code_from_checks", "ldpc_code_array",
    "ldpc_code_free", "ldpc_code_dims", "ldpc_code_tables", "ldpc_code_rate", "ldpc_code_save",
    "ldpc_decoder_cfg_default", "ldpc_decoder_create", "ldpc_decoder_destroy", "ldpc_decode_batch", "ldpc_decode_batch_i16",
    "ldpc_decode_batch_device", "ldpc_decoder_sync", "ldpc_decoder_get_stats", "ldpc_device_count",
    "ldpc_mc_run", "ldpc_mc_run_device", "ldpc_mc_channel", "ldpc_hard_decision_batch",
    "ldpc_decoder_device", "ldpc_decoder_code", "ldpc_decoder_max_iter", "ldpc_decode_batch_f64", "ldpc_mc_group_create", "ldpc_mc_group_size", "ldpc_mc_group_run", "ldpc_mc_group_destroy",
    "ldpc_mc_run_multi",
    "ldpc_encode_batch", "ldpc_encode_batch_device", "ldpc_gen_load", "ldpc_gen_from_code", "ldpc_gen_save", "ldpc_gen_free", "ldpc_gen_dims", "ldpc_gen_indices", "ldpc_gen_encode",
)
STREAM_PHILOX, STREAM_REFERENCE = 1, 2


class LdpcError(RuntimeError):
    def __init__(self, status, detail):
        self.status = status
        super().__init__("ldpc status %d (%s): %s" % (status, _strerror(status), detail))


class DecoderCfg(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("precheck", C.c_int), ("device", C.c_int), ("precision", C.c_int),
                ("threads", C.c_int), ("frames_per_cta", C.c_int)]


class DecoderStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("frames", C.c_uint64), ("fallback_frames", C.c_uint64),
                ("threads", C.c_int), ("threads32", C.c_int), ("frames_per_cta", C.c_int),
                ("frames_per_cta32", C.c_int), ("grid", C.c_int), ("smem_bytes", C.c_int),
                ("smem_bytes32", C.c_int),
                ("resident_ctas_per_sm", C.c_int), ("launch_smem_bytes", C.c_int)]


class McCfg(C.Structure):
    _fields_ = [("snr", C.c_double), ("sigma", C.c_double), ("stream", C.c_int), ("seed", C.c_uint64),
                ("first_frame", C.c_uint64), ("codeword", C.c_void_p), ("d_codewords", C.c_void_p),
                ("info_index", C.c_void_p),
                ("info_count", C.c_int), ("pin_index", C.c_void_p), ("pin_count", C.c_int), ("pin_value", C.c_int)]


class McCounters(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("frame_errors", C.c_uint64), ("bit_errors", C.c_uint64),
                ("iter_sum", C.c_uint64)]


class McStop(C.Structure):
    _fields_ = [("target_block_errors", C.c_uint64), ("max_frames", C.c_uint64), ("frames_per_round", C.c_size_t),
                ("count_iterations", C.c_int), ("iters_out", C.c_void_p), ("iters_cap", C.c_size_t)]


class McResult(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("block_errors", C.c_uint64), ("errors", C.c_uint64), ("iter_sum", C.c_uint64),
                ("iter_hist", C.c_uint64 * 32), ("rounds", C.c_uint64), ("reached", C.c_int), ("devices", C.c_int),
                ("seconds", C.c_double)]


_lib = None


def load_library():
    """dlopen the in-tree library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libldpc_b200.so is not built: run `python -m fixedpointldpc_b200.build` "
                           "(or __graft_entry__.build())")
    L = C.CDLL(LIB_PATH)
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    L.ldpc_strerror.restype = C.c_char_p
    L.ldpc_strerror.argtypes = [C.c_int]
    L.ldpc_last_error.restype = C.c_char_p
    L.ldpc_code_load.restype = vp
    L.ldpc_code_load.argtypes = [C.c_char_p, C.c_int, ip]
    L.ldpc_code_from_checks.restype = vp
    L.ldpc_code_from_checks.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, ip]
    L.ldpc_code_array.restype = vp
    L.ldpc_code_array.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, ip]
    L.ldpc_code_free.argtypes = [vp]
    L.ldpc_code_dims.argtypes = [vp, ip, ip, ip, ip, ip]
    L.ldpc_code_tables.argtypes = [vp, vp, vp, vp, vp]
    L.ldpc_code_rate.restype = C.c_double
    L.ldpc_code_rate.argtypes = [vp]
    L.ldpc_code_save.argtypes = [vp, C.c_char_p]
    L.ldpc_decoder_cfg_default.argtypes = [C.POINTER(DecoderCfg)]
    L.ldpc_decoder_create.restype = vp
    L.ldpc_decoder_create.argtypes = [vp, C.POINTER(DecoderCfg), ip]
    L.ldpc_decoder_destroy.argtypes = [vp]
    L.ldpc_decode_batch.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, vp]
    L.ldpc_decode_batch_i16.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, vp]
    L.ldpc_decode_batch_device.argtypes = [vp, vp, C.c_int, C.c_size_t, vp, vp, vp, vp, vp]
    L.ldpc_decoder_sync.argtypes = [vp]
    L.ldpc_decoder_get_stats.argtypes = [vp, C.POINTER(DecoderStats)]
    L.ldpc_device_count.restype = C.c_int
    L.ldpc_mc_run.argtypes = [vp, C.POINTER(McCfg), C.c_size_t, vp, vp, C.POINTER(McCounters)]
    L.ldpc_mc_run_device.argtypes = [vp, C.POINTER(McCfg), C.c_size_t, vp, vp, vp, vp]
    L.ldpc_mc_channel.argtypes = [vp, C.POINTER(McCfg), C.c_size_t, vp]
    if hasattr(L, "ldpc_mc_group_run"):  # (absent from older builds loaded through LDPC_B200_LIB for A/B runs)
        L.ldpc_decoder_device.argtypes = [vp]
        L.ldpc_decoder_code.restype = vp
        L.ldpc_decoder_code.argtypes = [vp]
        L.ldpc_decoder_max_iter.argtypes = [vp]
        L.ldpc_decode_batch_f64.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, vp]
        L.ldpc_mc_group_create.restype = vp
        L.ldpc_mc_group_create.argtypes = [C.POINTER(vp), C.c_int, ip]
        L.ldpc_mc_group_size.argtypes = [vp]
        L.ldpc_mc_group_run.argtypes = [vp, C.POINTER(McCfg), C.POINTER(McStop), C.POINTER(McResult)]
        L.ldpc_mc_group_destroy.argtypes = [vp]
        L.ldpc_mc_run_multi.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(McCfg), C.POINTER(McStop), C.POINTER(McResult)]
    L.ldpc_hard_decision_batch.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.ldpc_gen_load.restype = vp
    L.ldpc_gen_load.argtypes = [C.c_char_p, ip]
    L.ldpc_gen_free.argtypes = [vp]
    L.ldpc_gen_from_code.restype = vp
    L.ldpc_gen_from_code.argtypes = [vp, vp, C.c_int, ip]
    L.ldpc_gen_save.argtypes = [vp, C.c_char_p]
    L.ldpc_gen_dims.argtypes = [vp, ip, ip, ip]
    L.ldpc_gen_indices.argtypes = [vp, vp, vp]
    L.ldpc_gen_encode.argtypes = [vp, C.c_char_p, C.c_int, vp]
    L.ldpc_encode_batch.argtypes = [vp, C.c_int, vp, C.c_size_t, vp]
    L.ldpc_encode_batch_device.argtypes = [vp, C.c_int, vp, C.c_size_t, vp, vp]
    _lib = L
    return L


def _strerror(status):
    return load_library().ldpc_strerror(status).decode()


def _check(status):
    if status != OK:
        raise LdpcError(status, load_library().ldpc_last_error().decode())


def _ptr(arr):
    return None if arr is None else arr.ctypes.data


class Code:
    """Parity-check tables of one code (host).  Mirrors what FP_Decoder::ReadH fills
    (ArrayLDPC_Decoder.cpp:642-674) with runtime dimensions."""

    def __init__(self, handle):
        self._h = handle
        L = load_library()
        v = [C.c_int() for _ in range(5)]
        _check(L.ldpc_code_dims(self._h, *[C.byref(x) for x in v]))
        self.n, self.m, self.edges, self.dc_max, self.dv_max = (x.value for x in v)
        self.nw32 = (self.n + 31) // 32

    @classmethod
    def load(cls, path, fmt=FMT_AUTO):
        L = load_library()
        err = C.c_int()
        h = L.ldpc_code_load(os.fsencode(path), fmt, C.byref(err))
        if not h:
            _check(err.value)
        return cls(h)

    @classmethod
    def from_checks(cls, n, cdeg, clist):
        L = load_library()
        cdeg = np.ascontiguousarray(cdeg, np.int32)
        clist = np.ascontiguousarray(clist, np.int32)
        err = C.c_int()
        h = L.ldpc_code_from_checks(n, len(cdeg), _ptr(cdeg), _ptr(clist), clist.shape[1], C.byref(err))
        if not h:
            _check(err.value)
        return cls(h)

    @classmethod
    def array(cls, p, nrows, ncols=None, row_mult=None, col_sel=None, backward=False):
        L = load_library()
        ncols = p if ncols is None else ncols
        rm = None if row_mult is None else np.ascontiguousarray(row_mult, np.int32)
        cs = None if col_sel is None else np.ascontiguousarray(col_sel, np.int32)
        err = C.c_int()
        h = L.ldpc_code_array(p, nrows, _ptr(rm), ncols, _ptr(cs), int(backward), C.byref(err))
        if not h:
            _check(err.value)
        return cls(h)

    def tables(self):
        vdeg = np.zeros(self.n, np.int32); cdeg = np.zeros(self.m, np.int32)
        vlist = np.zeros((self.n, self.dv_max), np.int32); clist = np.zeros((self.m, self.dc_max), np.int32)
        _check(load_library().ldpc_code_tables(self._h, _ptr(vdeg), _ptr(cdeg), _ptr(vlist), _ptr(clist)))
        return vdeg, cdeg, vlist, clist

    @property
    def rate(self):
        return load_library().ldpc_code_rate(self._h)

    def save(self, path):
        _check(load_library().ldpc_code_save(self._h, os.fsencode(path)))

    def close(self):
        if self._h:
            load_library().ldpc_code_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Generator:
    """Format-B generator equations + FP_Encoder::encode (host side; ArrayLDPC_Encoder.cpp:34-225)."""

    def __init__(self, path=None, code=None, parity_cols=None):
        """Generator(path) parses Format B; Generator(code=..., parity_cols=...) derives the equations from H by
        GF(2) elimination (parity_cols fixes the solved-for columns, None = greedy from the last column)."""
        L = load_library()
        err = C.c_int()
        if path is not None:
            self._h = L.ldpc_gen_load(os.fsencode(path), C.byref(err))
        else:
            pc = None if parity_cols is None else np.ascontiguousarray(parity_cols, np.int32)
            self._h = L.ldpc_gen_from_code(code._h, _ptr(pc), 0 if pc is None else len(pc), C.byref(err))
        if not self._h:
            _check(err.value)
        v = [C.c_int() for _ in range(3)]
        _check(L.ldpc_gen_dims(self._h, *[C.byref(x) for x in v]))
        self.n, self.rows, self.k = (x.value for x in v)
        self.info_index = np.zeros(self.k, np.int32)
        self.parity_index = np.zeros(self.rows, np.int32)
        _check(L.ldpc_gen_indices(self._h, _ptr(self.info_index), _ptr(self.parity_index)))

    def save(self, path):
        _check(load_library().ldpc_gen_save(self._h, os.fsencode(path)))

    def encode_batch(self, info_bytes, device=0):
        """GPU encoder: info_bytes uint8 [frames][ceil(k/8)] (LSB first) -> packed codewords uint32 [frames][nw32]."""
        info = np.ascontiguousarray(info_bytes, np.uint8).reshape(-1, (self.k + 7) // 8)
        out = np.zeros((len(info), (self.n + 31) // 32), np.uint32)
        _check(load_library().ldpc_encode_batch(self._h, device, _ptr(info), len(info), _ptr(out)))
        return out

    def encode_batch_device(self, info_words_ptr, frames, codewords_ptr, device=0, cuda_stream=None):
        _check(load_library().ldpc_encode_batch_device(self._h, device, info_words_ptr, frames, codewords_ptr, cuda_stream))

    def encode(self, info_bytes):
        cw = np.zeros(self.n, np.uint8)
        _check(load_library().ldpc_gen_encode(self._h, info_bytes, len(info_bytes), _ptr(cw)))
        return cw

    def __del__(self):
        try:
            if self._h:
                load_library().ldpc_gen_free(self._h)
                self._h = None
        except Exception:
            pass


class Decoder:
    """Batched FP_Decoder::decode_general_fp (precheck=False) / decode_fixpoint (precheck=True)."""

    def __init__(self, code, max_iter=30, precheck=False, device=0, precision=0, threads=0, frames_per_cta=0):
        L = load_library()
        self.code = code
        cfg = DecoderCfg()
        L.ldpc_decoder_cfg_default(C.byref(cfg))
        cfg.max_iter, cfg.precheck, cfg.device, cfg.precision = max_iter, int(precheck), device, precision
        cfg.threads, cfg.frames_per_cta = threads, frames_per_cta
        err = C.c_int()
        self._h = L.ldpc_decoder_create(code._h, C.byref(cfg), C.byref(err))
        if not self._h:
            _check(err.value)

    def decode(self, llr, want_bits=True, want_post=False, want_v2c=False):
        """llr: int32 [frames][n] host array.  Returns dict(iters, bits, post, v2c)."""
        c = self.code
        llr = np.ascontiguousarray(llr, np.int32).reshape(-1, c.n)
        f = len(llr)
        out = {"iters": np.zeros(f, np.int32),
               "bits": np.zeros((f, c.nw32), np.uint32) if want_bits else None,
               "post": np.zeros((f, c.n), np.int32) if want_post else None,
               "v2c": np.zeros((f, c.dc_max, c.m), np.int32) if want_v2c else None}
        _check(load_library().ldpc_decode_batch(self._h, _ptr(llr), f, _ptr(out["iters"]), _ptr(out["bits"]),
                                                _ptr(out["post"]), _ptr(out["v2c"])))
        return out

    def decode_f64(self, llr, want_post=True, want_v2c=False):
        """FP_Decoder::decode_general(const double *): llr float64 [frames][n]; returns dict(iters, bits, post, v2c)."""
        c = self.code
        llr = np.ascontiguousarray(llr, np.float64).reshape(-1, c.n)
        f = len(llr)
        out = {"iters": np.zeros(f, np.int32), "bits": np.zeros((f, c.nw32), np.uint32),
               "post": np.zeros((f, c.n), np.float64) if want_post else None,
               "v2c": np.zeros((f, c.dc_max, c.m), np.float64) if want_v2c else None}
        _check(load_library().ldpc_decode_batch_f64(self._h, _ptr(llr), f, _ptr(out["iters"]), _ptr(out["bits"]),
                                                    _ptr(out["post"]), _ptr(out["v2c"])))
        return out

    def decode_raw(self, llr_ptr, frames, iters_ptr, bits_ptr=None, post_ptr=None, v2c_ptr=None, llr_bits=32):
        """Host-pointer call (pinned buffers for the end-to-end timing); llr_bits = 32 (`const int *LLR`) or 16."""
        fn = load_library().ldpc_decode_batch if llr_bits == 32 else load_library().ldpc_decode_batch_i16
        _check(fn(self._h, llr_ptr, frames, iters_ptr, bits_ptr, post_ptr, v2c_ptr))

    def decode_i16(self, llr, want_bits=True, want_post=False, want_v2c=False):
        """llr: int16 [frames][n] host array (ldpc_decode_batch_i16).  Returns dict(iters, bits, post, v2c)."""
        c = self.code
        llr = np.ascontiguousarray(llr, np.int16).reshape(-1, c.n)
        f = len(llr)
        out = {"iters": np.zeros(f, np.int32),
               "bits": np.zeros((f, c.nw32), np.uint32) if want_bits else None,
               "post": np.zeros((f, c.n), np.int32) if want_post else None,
               "v2c": np.zeros((f, c.dc_max, c.m), np.int32) if want_v2c else None}
        _check(load_library().ldpc_decode_batch_i16(self._h, _ptr(llr), f, _ptr(out["iters"]), _ptr(out["bits"]),
                                                    _ptr(out["post"]), _ptr(out["v2c"])))
        return out

    def decode_device(self, llr_ptr, llr_bits, frames, iters_ptr, bits_ptr=None, post_ptr=None, v2c_ptr=None,
                      stream=None):
        """Device-pointer call, asynchronous on `stream` (raw cudaStream_t value or None)."""
        _check(load_library().ldpc_decode_batch_device(self._h, llr_ptr, llr_bits, frames, iters_ptr, bits_ptr,
                                                       post_ptr, v2c_ptr, stream))

    def hard_decision(self, values):
        """FP_Decoder::hardDecision / checkPost_fp*: returns (fail[frames], bits[frames][nw32])."""
        c = self.code
        values = np.ascontiguousarray(values, np.int32).reshape(-1, c.n)
        fail = np.zeros(len(values), np.int32)
        bits = np.zeros((len(values), c.nw32), np.uint32)
        _check(load_library().ldpc_hard_decision_batch(self._h, _ptr(values), len(values), _ptr(fail), _ptr(bits)))
        return fail, bits

    def sync(self):
        _check(load_library().ldpc_decoder_sync(self._h))

    # ---- Monte-Carlo mode -------------------------------------------------------------------
    def _mc_cfg(self, snr, sigma=None, stream=STREAM_PHILOX, seed=1, first_frame=0, codeword=None, info_index=None,
                pin_index=None, pin_value=0, d_codewords=None):
        keep = []
        cfg = McCfg()
        cfg.snr = snr
        cfg.sigma = float(np.sqrt(1.0 / snr)) if sigma is None else sigma
        cfg.stream, cfg.seed, cfg.first_frame = stream, seed, first_frame
        if d_codewords is not None:
            cfg.d_codewords = d_codewords
        if codeword is not None:
            cw = np.ascontiguousarray(codeword, np.uint8); keep.append(cw); cfg.codeword = cw.ctypes.data
        if info_index is not None:
            ii = np.ascontiguousarray(info_index, np.int32); keep.append(ii)
            cfg.info_index, cfg.info_count = ii.ctypes.data, len(ii)
        if pin_index is not None and len(pin_index):
            pi = np.ascontiguousarray(pin_index, np.int32); keep.append(pi)
            cfg.pin_index, cfg.pin_count, cfg.pin_value = pi.ctypes.data, len(pi), pin_value
        return cfg, keep

    def mc_run(self, frames, snr, want_frame_err=True, want_iters=False, **kw):
        """One batch of the drivers' loop body (channel -> decode -> calculateBER) on the GPU.
        Returns dict(frames, frame_errors, bit_errors, iter_sum, frame_err[frames], iters[frames])."""
        cfg, keep = self._mc_cfg(snr, **kw)
        ferr = np.zeros(frames, np.uint16) if want_frame_err else None
        iters = np.zeros(frames, np.int32) if want_iters else None
        tot = McCounters()
        _check(load_library().ldpc_mc_run(self._h, C.byref(cfg), frames, _ptr(ferr), _ptr(iters), C.byref(tot)))
        del keep
        return {"frames": tot.frames, "frame_errors": tot.frame_errors, "bit_errors": tot.bit_errors,
                "iter_sum": tot.iter_sum, "frame_err": ferr, "iters": iters}

    def mc_run_device(self, frames, snr, counters_ptr, frame_err_ptr=None, iters_ptr=None, cuda_stream=None, **kw):
        cfg, keep = self._mc_cfg(snr, **kw)
        _check(load_library().ldpc_mc_run_device(self._h, C.byref(cfg), frames, frame_err_ptr, iters_ptr, counters_ptr,
                                                 cuda_stream))
        del keep

    def mc_channel(self, frames, snr, **kw):
        """The quantised LLRs [frames][n] mc_run would decode."""
        cfg, keep = self._mc_cfg(snr, **kw)
        out = np.zeros((frames, self.code.n), np.int32)
        _check(load_library().ldpc_mc_channel(self._h, C.byref(cfg), frames, _ptr(out)))
        del keep
        return out

    def stats(self):
        s = DecoderStats()
        _check(load_library().ldpc_decoder_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in DecoderStats._fields_}

    def close(self):
        if self._h:
            load_library().ldpc_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class McGroup:
    """One Monte-Carlo point on several GPUs of the box (ldpc_mc_group_*): one Decoder per device, one host thread per
    device inside the library, counters all-reduced over NCCL once per round."""

    def __init__(self, decoders):
        L = load_library()
        self.decoders = list(decoders)
        if len(self.decoders) > 1 and "LDPC_NCCL_LIB" not in os.environ:
            # one NCCL per process: prefer the one PyTorch ships (same soname as the system's), located without
            # importing torch
            import importlib.util
            try:
                spec = importlib.util.find_spec("nvidia.nccl")
                for root in (spec.submodule_search_locations if spec else []):
                    cand = os.path.join(root, "lib", "libnccl.so.2")
                    if os.path.exists(cand):
                        os.environ["LDPC_NCCL_LIB"] = cand
                        break
            except Exception:
                pass
        arr = (C.c_void_p * len(self.decoders))(*[d._h for d in self.decoders])
        err = C.c_int()
        self._h = L.ldpc_mc_group_create(arr, len(self.decoders), C.byref(err))
        if not self._h:
            _check(err.value)

    def run(self, snr, target_block_errors=100, max_frames=0, frames_per_round=0, count_iterations=False, want_iters=0, **kw):
        """Returns dict(frames, block_errors, errors, iter_sum, iter_hist, rounds, reached, devices, seconds[, iters])."""
        cfg, keep = self.decoders[0]._mc_cfg(snr, **kw)
        stop = McStop(target_block_errors, max_frames, frames_per_round, int(count_iterations), None, 0)
        log = None
        if want_iters:
            log = np.full(want_iters, -2, np.int32)
            stop.iters_out, stop.iters_cap = log.ctypes.data, want_iters
        res = McResult()
        _check(load_library().ldpc_mc_group_run(self._h, C.byref(cfg), C.byref(stop), C.byref(res)))
        del keep
        out = {k: getattr(res, k) for k in ("frames", "block_errors", "errors", "iter_sum", "rounds", "reached", "devices", "seconds")}
        out["iter_hist"] = np.array(list(res.iter_hist), np.int64)
        if log is not None:
            out["iters"] = log[:min(want_iters, out["frames"])]
        return out

    def close(self):
        if self._h:
            load_library().ldpc_mc_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unpack_bits(bits, n):
    """[frames][nw32] uint32 -> [frames][n] 0/1 ints (bit v%32 of word v/32)."""
    b = np.ascontiguousarray(bits, np.uint32)
    out = ((b[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(len(b), -1)
    return out[:, :n].astype(np.int32)


def device_count():
    return load_library().ldpc_device_count()
