"""markov-huffman-coding_b200 — host-side mirror of the reference's coder interface over the C ABI of libmh_gpu.so.

The reference is compiled C++ (its product has no Python); its real host driver here is the C++ CLI
(csrc/cli_main.cpp -> bin/markovhuffman). This module is the thin ctypes binding the tests and bench.py use to
drive the same C ABI (include/mh_gpu.h). Names follow the reference:

    i_coding_provider          -> CodingProvider          (src/coding.h:18-35)
      .compress / .decompress  -> .compress / .decompress (src/coding.cpp:61-160) — run on the GPU, host buffers in/out
      .write_coding_tree       -> .write_coding_tree      (src/huffman.cpp:83-85, src/markov_huffman.cpp:80-88)
      .get_type / .get_encoding / .decoding_lookup / .print_table + .print_tree (debug_dump)
    huffman_table(int*)        -> CodingProvider.from_counts(counts, order=0)
    markov_huffman_table(int*) -> CodingProvider.from_counts(counts, order=1)
    *_table(bitbuffer&)        -> CodingProvider.from_table_file(bytes)
    construct_table            -> Session.histogram / gpu_histogram            (src/main.cpp:29-39)

There is no CPU fallback: if libmh_gpu.so is missing the import fails; if no CUDA device is usable every GPU call
raises MhError(MH_ERR_NO_DEVICE / MH_ERR_CUDA).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmh_gpu.so")
CLI_PATH = os.path.join(_HERE, "bin", "markovhuffman")

ORDER_HUFFMAN = 0
ORDER_MARKOV = 1
PREV0 = 0x20

MH_OK = 0
MH_ERR_INVALID_ARG = -1
MH_ERR_CUDA = -2
MH_ERR_NO_DEVICE = -3
MH_ERR_CAPACITY = -4
MH_ERR_BAD_TABLE = -5
MH_ERR_CODE_TOO_LONG = -6
MH_ERR_BAD_HEADER = -7
MH_ERR_TYPE_MISMATCH = -8
MH_ERR_CORRUPT_STREAM = -9
MH_ERR_COUNT_WRAPPED = -10
MH_ERR_NOT_CONVERGED = -11
MH_ERR_WORKSPACE = -12

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libmh_gpu.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
        "`make -C markov-huffman-coding_b200`. There is no CPU fallback." % LIB_PATH)

_lib = ctypes.CDLL(LIB_PATH)
_vp, _u64, _i, _sz = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_size_t
_u8 = ctypes.c_uint8
_pp = ctypes.POINTER(ctypes.c_void_p)
_pu64 = ctypes.POINTER(ctypes.c_uint64)
_pi = ctypes.POINTER(ctypes.c_int)
_psz = ctypes.POINTER(ctypes.c_size_t)

# every symbol include/mh_gpu.h declares, with its signature (tests check this list against the header)
_SIGNATURES = {
    "mh_status_string": (ctypes.c_char_p, [_i]),
    "mh_last_error": (ctypes.c_char_p, []),
    "mh_device_count": (_i, []),
    "mh_device_memory": (_i, [_i, _pu64, _pu64]),
    "mh_version": (_i, []),
    "mh_tunable_set": (_i, [ctypes.c_char_p, ctypes.c_longlong]),
    "mh_tunable_get": (_i, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_longlong)]),
    "mh_pinned_alloc": (_vp, [_sz]),
    "mh_pinned_free": (None, [_vp]),
    "mh_table_from_counts": (_i, [_vp, _i, _pp]),
    "mh_table_from_bytes": (_i, [_vp, _sz, _pp]),
    "mh_table_serialize": (_i, [_vp, _vp, _sz, _psz]),
    "mh_table_order": (_i, [_vp]),
    "mh_table_context_empty": (_i, [_vp, _i]),
    "mh_table_code": (_i, [_vp, _i, _i, _vp, _pi]),
    "mh_table_max_code_bits": (_i, [_vp]),
    "mh_table_code_lengths": (_i, [_vp, _vp, _sz]),
    "mh_table_lookup": (_i, [_vp, _i, _i, _pi, _pi, _pi]),
    "mh_table_pair_lut": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "mh_table_debug_dump": (_i, [_vp, _vp, _sz, _psz]),
    "mh_table_destroy": (None, [_vp]),
    "mh_codebook_create": (_i, [_vp, _pp]),
    "mh_codebook_update": (_i, [_vp, _vp, _vp]),
    "mh_codebook_destroy": (None, [_vp]),
    "mh_codebook_create_empty": (_i, [_pp]),
    "mh_codebook_build_device": (_i, [_vp, _vp, _i, _vp]),
    "mh_codebook_download": (_i, [_vp, _vp, _vp, _vp]),
    "mh_dectable_create": (_i, [_vp, _pp]),
    "mh_dectable_update": (_i, [_vp, _vp, _vp]),
    "mh_dectable_destroy": (None, [_vp]),
    "mh_workspace_create": (_i, [_u64, _u64, _pp]),
    "mh_workspace_destroy": (None, [_vp]),
    "mh_gpu_histogram": (_i, [_vp, _u64, _u8, _i, _vp, _vp, _vp]),
    "mh_gpu_encode": (_i, [_vp, _u64, _u8, _vp, _u64, _vp, _u64, _vp, _vp, _vp]),
    "mh_gpu_decode": (_i, [_vp, _u64, _u64, _u8, _vp, _vp, _u64, _vp, _vp, _vp]),
    "mh_gpu_decode_shard": (_i, [_vp, ctypes.c_uint32, _u64, _u64, _i, _u8, ctypes.c_uint32, _i, _vp, _vp, _u64, _vp, _vp, _vp]),
    "mh_decode_subsequence_bits": (ctypes.c_uint32, [_i, _u64]),
    "mh_session_create": (_i, [_i, _u64, _pp]),
    "mh_session_create_sized": (_i, [_i, _u64, _u64, _pp]),
    "mh_session_fetch": (_i, [_vp, _vp, _u64, _pu64]),
    "mh_session_destroy": (None, [_vp]),
    "mh_session_compress": (_i, [_vp, _vp, _u64, _i, _vp, _u64, _pu64, _pp]),
    "mh_session_compress_with_table": (_i, [_vp, _vp, _vp, _u64, _vp, _u64, _pu64, _pu64]),
    "mh_session_decompress": (_i, [_vp, _vp, _vp, _u64, _vp, _u64, _pu64]),
    "mh_session_histogram": (_i, [_vp, _vp, _u64, _i, _vp]),
    "mh_synth_markov": (_i, [_vp, _u64, _u64, _u64, _vp, _u64, _vp]),
    "mh_synth_fibonacci": (_i, [_i, _u8, _u64, _u64, _vp, _u64, _vp]),
    "mh_comm_available": (_i, []),
    "mh_comm_unique_id": (_i, [_vp]),
    "mh_comm_create": (_i, [_i, _i, _i, _vp, _pp]),
    "mh_comm_create_local": (_i, [_i, _vp, _i, _pp]),
    "mh_comm_rank": (_i, [_vp]),
    "mh_comm_world": (_i, [_vp]),
    "mh_comm_device": (_i, [_vp]),
    "mh_comm_transport": (_i, [_vp]),
    "mh_comm_destroy": (None, [_vp]),
    "mh_comm_reserve": (_i, [_vp, _u64, _u64]),
    "mh_comm_stats": (_i, [_vp, _vp, _i, _i]),
    "mh_shard_local_bytes": (_u64, [_u64]),
    "mh_shard_payload_offset": (ctypes.c_uint32, []),
    "mh_sharded_compress": (_i, [_vp, _vp, _u64, _i, _vp, _u64, _vp, _pp, _i, _vp]),
    "mh_sharded_decompress": (_i, [_vp, _vp, _vp, _u64, _vp, _i, _vp, _u64, _pu64, _pu64, _vp]),
    "mh_sharded_compress_host": (_i, [_vp, _vp, _u64, _i, _vp, _u64, _pu64, _pp]),
    "mh_sharded_decompress_host": (_i, [_vp, _vp, _vp, _u64, _vp, _u64, _pu64]),
    "mh_kernel_launches": (_u64, []),
    "mh_profile_enable": (_i, [_i]),
    "mh_profile_report": (_i, [_vp, _sz, _psz]),
}
for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(_lib, _name)   # AttributeError here = the library does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args


class MhError(RuntimeError):
    def __init__(self, status, where=""):
        self.status = status
        detail = _lib.mh_last_error().decode() if status == MH_ERR_CUDA else ""
        super().__init__("%s: %s (%d)%s" % (where, _lib.mh_status_string(status).decode(), status, " — " + detail if detail else ""))


def _check(status, where):
    if status != MH_OK:
        raise MhError(status, where)


def device_count():
    return _lib.mh_device_count()


def device_memory(device=0):
    """(free bytes, total bytes) of a device; a host driver sizes its session from it."""
    free, total = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _check(_lib.mh_device_memory(int(device), ctypes.byref(free), ctypes.byref(total)), "mh_device_memory")
    return free.value, total.value


def tunable_set(name, value):
    """Experiment / test override of a library tunable (-1 restores the default)."""
    _check(_lib.mh_tunable_set(name.encode(), int(value)), "mh_tunable_set(%s)" % name)


def tunable_get(name):
    v = ctypes.c_longlong(0)
    _check(_lib.mh_tunable_get(name.encode(), ctypes.byref(v)), "mh_tunable_get(%s)" % name)
    return v.value


def kernel_launches():
    return _lib.mh_kernel_launches()


def profile_enable(on=True):
    _check(_lib.mh_profile_enable(1 if on else 0), "mh_profile_enable")


def profile_report():
    """{kernel: {"launches": n, "ms": total}} for the launches recorded since profile_enable()."""
    import json
    buf = ctypes.create_string_buffer(1 << 16)
    n = ctypes.c_size_t(0)
    _check(_lib.mh_profile_report(buf, len(buf), ctypes.byref(n)), "mh_profile_report")
    return json.loads(buf.value.decode())


def _as_buffer(data):
    """bytes-like / numpy uint8 -> (keepalive object, address, length) without copying when possible."""
    try:
        import numpy as np
    except ImportError:  # pragma: no cover
        np = None
    if np is not None and isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.uint8)
        return a, a.ctypes.data, a.size
    b = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    buf = (ctypes.c_uint8 * len(b)).from_buffer_copy(b) if len(b) else (ctypes.c_uint8 * 1)()
    return buf, ctypes.addressof(buf), len(b)


class CodingProvider:
    """The reference's i_coding_provider: one Huffman tree (order 0, `-h`) or 256 context trees (order 1)."""

    def __init__(self, handle):
        self._h = ctypes.c_void_p(handle)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.mh_table_destroy(h)
            self._h = None

    # ---- constructors ---------------------------------------------------------------------------------
    @classmethod
    def from_counts(cls, counts, order):
        """huffman_table(int*) / markov_huffman_table(int*). counts: 256 or 65536 non-negative ints (uint64)."""
        n = 65536 if order else 256
        arr = (ctypes.c_uint64 * n)(*[int(c) & 0xFFFFFFFFFFFFFFFF for c in counts])
        out = ctypes.c_void_p()
        _check(_lib.mh_table_from_counts(arr, int(order), ctypes.byref(out)), "mh_table_from_counts")
        return cls(out.value)

    @classmethod
    def from_counts_array(cls, counts_u64, order):
        """Same, from a contiguous numpy uint64 array (no per-element conversion)."""
        out = ctypes.c_void_p()
        _check(_lib.mh_table_from_counts(counts_u64.ctypes.data, int(order), ctypes.byref(out)), "mh_table_from_counts")
        return cls(out.value)

    @classmethod
    def from_table_file(cls, data):
        """huffman_table(bitbuffer&) / markov_huffman_table(bitbuffer&): load an encoding-table file image."""
        keep, addr, n = _as_buffer(data)
        out = ctypes.c_void_p()
        _check(_lib.mh_table_from_bytes(addr, n, ctypes.byref(out)), "mh_table_from_bytes")
        return cls(out.value)

    # ---- i_coding_provider surface --------------------------------------------------------------------
    def get_type(self):
        return _lib.mh_table_order(self._h)

    def empty(self, prev=0):
        return bool(_lib.mh_table_context_empty(self._h, prev))

    def write_coding_tree(self):
        n = ctypes.c_size_t(0)
        cap = 1 << 17
        buf = (ctypes.c_uint8 * cap)()
        _check(_lib.mh_table_serialize(self._h, buf, cap, ctypes.byref(n)), "mh_table_serialize")
        return bytes(buf[: n.value])

    def get_encoding(self, prev, c):
        """(length, '0101…' string) — length 0 means the symbol has no codeword."""
        bits = (ctypes.c_uint8 * 32)()
        ln = ctypes.c_int(0)
        _check(_lib.mh_table_code(self._h, prev, c, bits, ctypes.byref(ln)), "mh_table_code")
        return ln.value, "".join(str((bits[b // 8] >> (7 - b % 8)) & 1) for b in range(ln.value))

    def max_code_bits(self):
        return _lib.mh_table_max_code_bits(self._h)

    def code_lengths(self, dtype=None):
        """numpy array of all code lengths: [65536] indexed 256*prev + c (order 1) or [256] (order 0); uint64 unless
        another dtype is asked for (uint8 is what the library hands out)."""
        import numpy as np
        n = 65536 if self.get_type() else 256
        lens = np.zeros(n, dtype=np.uint8)
        _check(_lib.mh_table_code_lengths(self._h, lens.ctypes.data, n), "mh_table_code_lengths")
        return lens.astype(np.uint64 if dtype is None else dtype, copy=False)

    def decoding_lookup(self, prev, window):
        """(kind, value, depth): kind 0 null, 1 leaf, 2 internal node at depth 8."""
        k, v, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(_lib.mh_table_lookup(self._h, prev, window, ctypes.byref(k), ctypes.byref(v), ctypes.byref(d)), "mh_table_lookup")
        return k.value, v.value, d.value

    def pair_lut(self):
        """The decoder's two-symbol table: (table u32[rows*256], rank u8[256], live u8[64], len1 u8[ctx_rows*256], rows,
        ctx_rows) as numpy arrays, or None when the table has more than 63 live contexts."""
        import numpy as np
        table = np.zeros(64 * 256, dtype=np.uint32)
        maps = np.zeros(256 + 64 * 257, dtype=np.uint8)
        rows, ctx_rows = ctypes.c_uint32(0), ctypes.c_uint32(0)
        _check(_lib.mh_table_pair_lut(self._h, table.ctypes.data, maps.ctypes.data, ctypes.byref(rows), ctypes.byref(ctx_rows)), "mh_table_pair_lut")
        if rows.value == 0:
            return None
        return (table[: rows.value * 256], maps[:256], maps[256:320], maps[320 : 320 + ctx_rows.value * 256], rows.value, ctx_rows.value)

    def debug_dump(self):
        """print_table() followed by print_tree(): the `-g` output."""
        n = ctypes.c_size_t(0)
        _lib.mh_table_debug_dump(self._h, None, 0, ctypes.byref(n))
        buf = ctypes.create_string_buffer(max(1, n.value))
        _check(_lib.mh_table_debug_dump(self._h, buf, n.value, ctypes.byref(n)), "mh_table_debug_dump")
        return buf.raw[: n.value]

    def compress(self, data, session=None):
        """i_coding_provider::compress with this (given) table: the `-e` path. Returns header + payload."""
        s = session or default_session(len(data))
        return s.compress_with_table(self, data)[0]

    def decompress(self, stream, session=None):
        s = session or default_session(max(64, 8 * len(stream)))
        return s.decompress(self, stream)


class Session:
    """Owns a stream and device buffers for inputs up to max_input_bytes (mh_session)."""

    def __init__(self, max_input_bytes, device=0):
        out = ctypes.c_void_p()
        _check(_lib.mh_session_create(int(device), int(max_input_bytes), ctypes.byref(out)), "mh_session_create")
        self._h = out
        self.max_input_bytes = int(max_input_bytes)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mh_session_destroy(self._h)
            self._h = None

    __del__ = close

    def histogram(self, data, order):
        import numpy as np
        keep, addr, n = _as_buffer(data)
        counts = np.zeros(65536 if order else 256, dtype=np.uint64)
        _check(_lib.mh_session_histogram(self._h, addr, n, int(order), counts.ctypes.data), "mh_session_histogram")
        return counts

    def compress(self, data, order, out=None):
        """`markovhuffman in -o out [-h] -d table` without the file I/O: returns (stream bytes, CodingProvider)."""
        import numpy as np
        keep, addr, n = _as_buffer(data)
        cap = n + (n >> 3) + 4160
        buf = out if out is not None else np.empty(cap, dtype=np.uint8)
        out_len = ctypes.c_uint64(0)
        table = ctypes.c_void_p()
        _check(_lib.mh_session_compress(self._h, addr, n, int(order), buf.ctypes.data, buf.size, ctypes.byref(out_len), ctypes.byref(table)),
               "mh_session_compress")
        return buf[: out_len.value].tobytes(), CodingProvider(table.value)

    def compress_with_table(self, provider, data):
        import numpy as np
        keep, addr, n = _as_buffer(data)
        # a foreign table can expand the input: size for its longest codeword
        cap = max(n + (n >> 3), (n * max(8, provider.max_code_bits()) + 7) // 8) + 4160
        buf = np.empty(cap, dtype=np.uint8)
        out_len, dropped = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _check(_lib.mh_session_compress_with_table(self._h, provider._h, addr, n, buf.ctypes.data, buf.size, ctypes.byref(out_len), ctypes.byref(dropped)),
               "mh_session_compress_with_table")
        return buf[: out_len.value].tobytes(), dropped.value

    def decompress_into(self, provider, stream, capacity):
        """i_coding_provider::decompress straight into a host buffer of `capacity` bytes (one call; large streams are
        decoded in chunks with the copies of neighbouring chunks overlapping the kernels). Returns the bytes."""
        import numpy as np
        keep, addr, n = _as_buffer(stream)
        buf = np.empty(max(1, int(capacity)), dtype=np.uint8)
        out_len = ctypes.c_uint64(0)
        _check(_lib.mh_session_decompress(self._h, provider._h, addr, n, buf.ctypes.data, buf.size, ctypes.byref(out_len)), "mh_session_decompress")
        return buf[: out_len.value].tobytes()

    def decompress(self, provider, stream):
        import numpy as np
        keep, addr, n = _as_buffer(stream)
        out_len = ctypes.c_uint64(0)
        rc = _lib.mh_session_decompress(self._h, provider._h, addr, n, None, 0, ctypes.byref(out_len))   # decode on the device, learn the size
        if rc not in (MH_OK, MH_ERR_CORRUPT_STREAM):
            raise MhError(rc, "mh_session_decompress")
        need = out_len.value
        buf = np.empty(max(1, need), dtype=np.uint8)
        frc = _lib.mh_session_fetch(self._h, buf.ctypes.data, buf.size, ctypes.byref(out_len))
        if frc == MH_ERR_WORKSPACE:
            # the stream did not fit the session's device buffers: it was decoded in chunks and only counted, nothing
            # stays resident to fetch — decode again, chunk by chunk, into the host buffer
            rc = _lib.mh_session_decompress(self._h, provider._h, addr, n, buf.ctypes.data, buf.size, ctypes.byref(out_len))
            if rc not in (MH_OK, MH_ERR_CORRUPT_STREAM):
                raise MhError(rc, "mh_session_decompress")
        else:
            _check(frc, "mh_session_fetch")
        if rc != MH_OK:
            raise MhError(rc, "mh_session_decompress")
        return buf[: out_len.value].tobytes()


_default = None


def default_session(min_bytes):
    """A lazily grown process-wide session, for the one-shot helpers."""
    global _default
    need = max(int(min_bytes), 1 << 20)
    if _default is None or _default.max_input_bytes < need:
        if _default is not None:
            _default.close()
        _default = Session(need)
    return _default


def compress(data, order=ORDER_MARKOV):
    """One-shot `-d` path: (compressed file image, table file image)."""
    stream, provider = default_session(len(data)).compress(data, order)
    return stream, provider.write_coding_tree()


def decompress(stream, table_file):
    """One-shot `-x -e` path."""
    provider = CodingProvider.from_table_file(table_file)
    return provider.decompress(stream)


# ---- device-pointer entry points (ints are raw CUDA device addresses, e.g. torch_tensor.data_ptr()) --------
class Workspace:
    def __init__(self, max_input_bytes, max_payload_bytes):
        out = ctypes.c_void_p()
        _check(_lib.mh_workspace_create(int(max_input_bytes), int(max_payload_bytes), ctypes.byref(out)), "mh_workspace_create")
        self._h = out

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mh_workspace_destroy(self._h)
            self._h = None

    __del__ = close


class Codebook:
    def __init__(self, provider=None):
        out = ctypes.c_void_p()
        if provider is None:
            _check(_lib.mh_codebook_create_empty(ctypes.byref(out)), "mh_codebook_create_empty")
        else:
            _check(_lib.mh_codebook_create(provider._h, ctypes.byref(out)), "mh_codebook_create")
        self._h = out

    def build_device(self, d_counts, order, stream=0):
        """The encoder's tables straight from the device-resident histogram (no host round trip)."""
        _check(_lib.mh_codebook_build_device(self._h, d_counts, int(order), stream or None), "mh_codebook_build_device")

    def download(self):
        """(wide u64[65536], ctx u32[rows*256], meta u32[8]) as numpy arrays (synchronises the device)."""
        import numpy as np
        enc = np.zeros(65536, dtype=np.uint64)
        ctx = np.zeros(96 * 256, dtype=np.uint32)
        meta = np.zeros(8, dtype=np.uint32)
        _check(_lib.mh_codebook_download(self._h, enc.ctypes.data, ctx.ctypes.data, meta.ctypes.data), "mh_codebook_download")
        return enc, (ctx[: int(meta[0]) * 256] if meta[0] <= 96 else ctx[:0]), meta

    def update(self, provider, stream=0):
        _check(_lib.mh_codebook_update(self._h, provider._h, stream or None), "mh_codebook_update")

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mh_codebook_destroy(self._h)
            self._h = None

    __del__ = close


class DecodeTable:
    def __init__(self, provider):
        out = ctypes.c_void_p()
        _check(_lib.mh_dectable_create(provider._h, ctypes.byref(out)), "mh_dectable_create")
        self._h = out

    def update(self, provider, stream=0):
        _check(_lib.mh_dectable_update(self._h, provider._h, stream or None), "mh_dectable_update")

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mh_dectable_destroy(self._h)
            self._h = None

    __del__ = close


def gpu_histogram(d_in, n, prev0, order, d_counts, ws, stream=0):
    _check(_lib.mh_gpu_histogram(d_in, n, prev0, order, d_counts, ws._h, stream or None), "mh_gpu_histogram")


def gpu_encode(d_in, n, prev0, codebook, bit_base, d_out, out_capacity, d_result, ws, stream=0):
    _check(_lib.mh_gpu_encode(d_in, n, prev0, codebook._h, bit_base, d_out, out_capacity, d_result, ws._h, stream or None), "mh_gpu_encode")


def gpu_decode(d_bits, bit_base, n_bits, prev0, dectable, d_out, out_capacity, d_result, ws, stream=0):
    _check(_lib.mh_gpu_decode(d_bits, bit_base, n_bits, prev0, dectable._h, d_out, out_capacity, d_result, ws._h, stream or None), "mh_gpu_decode")


def synth_markov(trans_counts_u32, seed, seg_bytes, first_seg, d_out, n, stream=0):
    """trans_counts_u32: contiguous numpy uint32[65536] on the host; d_out: device address."""
    _check(_lib.mh_synth_markov(trans_counts_u32.ctypes.data, seed, seg_bytes, first_seg, d_out, n, stream or None), "mh_synth_markov")


def synth_fibonacci(k, base, seed, first_index, d_out, n, stream=0):
    _check(_lib.mh_synth_fibonacci(k, base, seed, first_index, d_out, n, stream or None), "mh_synth_fibonacci")


DECODE_WARM_UNIT = 8192


def gpu_decode_shard(d_bits, start_bit, n_bits, buf_bytes, exact_start, prev0, warm_bits, stream_end, dectable, d_out,
                     out_capacity, d_result, ws, stream=0):
    _check(_lib.mh_gpu_decode_shard(d_bits, start_bit, n_bits, buf_bytes, int(exact_start), prev0, warm_bits, int(stream_end),
                                    dectable._h, d_out, out_capacity, d_result, ws._h, stream or None), "mh_gpu_decode_shard")


def decode_subsequence_bits(order, n_bits):
    return _lib.mh_decode_subsequence_bits(int(order), int(n_bits))


# ---- one logical stream over several GPUs (mh_comm_* / mh_sharded_*) ------------------------------------------
MAX_SHARDS = 64
COMM_ID_BYTES = 128
STAT_NAMES = ("gather_us", "halo_us", "seam_us", "trees_us", "codebook_us", "dectable_us", "seam_rounds", "calls")


class ShardLayout(ctypes.Structure):
    """mh_shard_layout: how one stream is cut into `world` shards."""
    _fields_ = [("world", ctypes.c_int), ("order", ctypes.c_int), ("exact", ctypes.c_int), ("total_bits", ctypes.c_uint64),
                ("dropped", ctypes.c_uint64), ("bit_base", ctypes.c_uint64 * MAX_SHARDS), ("n_bits", ctypes.c_uint64 * MAX_SHARDS),
                ("prev0", ctypes.c_uint8 * MAX_SHARDS)]


def comm_available():
    return bool(_lib.mh_comm_available())


def comm_unique_id():
    buf = (ctypes.c_uint8 * COMM_ID_BYTES)()
    _check(_lib.mh_comm_unique_id(buf), "mh_comm_unique_id")
    return bytes(buf)


def shard_local_bytes(max_payload_bytes):
    return _lib.mh_shard_local_bytes(int(max_payload_bytes))


def shard_payload_offset():
    return _lib.mh_shard_payload_offset()


class Comm:
    """One rank of a multi-GPU group (mh_comm). Every rank makes the same compress / decompress calls."""

    def __init__(self, handle):
        self._h = ctypes.c_void_p(handle)

    @classmethod
    def create(cls, device, rank, world, unique_id):
        """One process per GPU: `unique_id` comes from rank 0's comm_unique_id(), handed round by the caller."""
        out = ctypes.c_void_p()
        buf = (ctypes.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id) if unique_id else None
        _check(_lib.mh_comm_create(int(device), int(rank), int(world), buf, ctypes.byref(out)), "mh_comm_create")
        return cls(out.value)

    @classmethod
    def create_local(cls, world, devices=None, use_nccl=1):
        """`world` ranks inside this process (each must then be driven by its own thread)."""
        out = (ctypes.c_void_p * world)()
        devs = (ctypes.c_int * world)(*devices) if devices is not None else None
        _check(_lib.mh_comm_create_local(int(world), devs, int(use_nccl), out), "mh_comm_create_local")
        return [cls(h) for h in out]

    rank = property(lambda self: _lib.mh_comm_rank(self._h))
    world = property(lambda self: _lib.mh_comm_world(self._h))
    device = property(lambda self: _lib.mh_comm_device(self._h))
    transport = property(lambda self: ("none", "nccl", "in-process")[_lib.mh_comm_transport(self._h)])

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mh_comm_destroy(self._h)
            self._h = None

    __del__ = close

    def reserve(self, max_shard_bytes, max_payload_bytes):
        _check(_lib.mh_comm_reserve(self._h, int(max_shard_bytes), int(max_payload_bytes)), "mh_comm_reserve")

    def stats(self, reset=False):
        buf = (ctypes.c_double * len(STAT_NAMES))()
        _check(_lib.mh_comm_stats(self._h, buf, len(STAT_NAMES), 1 if reset else 0), "mh_comm_stats")
        return dict(zip(STAT_NAMES, buf))

    def compress(self, d_in, n, order, d_local, local_cap, prepare_decode=False, stream=0):
        """Collective: (ShardLayout, CodingProvider). d_in / d_local are device addresses."""
        layout = ShardLayout()
        table = ctypes.c_void_p()
        _check(_lib.mh_sharded_compress(self._h, d_in, int(n), int(order), d_local, int(local_cap), ctypes.byref(layout), ctypes.byref(table),
                                        1 if prepare_decode else 0, stream or None), "mh_sharded_compress")
        return layout, CodingProvider(table.value)

    def compress_host(self, data_np, order, out_np):
        """Ranks of one process on shared host buffers (numpy uint8 arrays): returns (stream length, CodingProvider or None)."""
        out_len = ctypes.c_uint64(0)
        table = ctypes.c_void_p()
        _check(_lib.mh_sharded_compress_host(self._h, data_np.ctypes.data, data_np.size, int(order), out_np.ctypes.data, out_np.size,
                                             ctypes.byref(out_len), ctypes.byref(table)), "mh_sharded_compress_host")
        return out_len.value, (CodingProvider(table.value) if table.value else None)

    def decompress_host(self, provider, stream_np, out_np):
        out_len = ctypes.c_uint64(0)
        _check(_lib.mh_sharded_decompress_host(self._h, provider._h, stream_np.ctypes.data, stream_np.size, out_np.ctypes.data, out_np.size,
                                               ctypes.byref(out_len)), "mh_sharded_decompress_host")
        return out_len.value

    def decompress(self, provider, d_local, local_cap, layout, d_out, out_capacity, speculative=True, stream=0):
        """Collective: (symbols this rank wrote, symbols of the ranks before it)."""
        n_out, off = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _check(_lib.mh_sharded_decompress(self._h, provider._h, d_local, int(local_cap), ctypes.byref(layout), 1 if speculative else 0,
                                          d_out, int(out_capacity), ctypes.byref(n_out), ctypes.byref(off), stream or None), "mh_sharded_decompress")
        return n_out.value, off.value
