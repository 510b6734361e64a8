"""ctypes bindings for the test oracle (oracle/libmh_oracle.so) and, when present, the reference harness
(oracle/_ref/libmh_ref.so). TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = os.path.join(ORACLE_DIR, "libmh_oracle.so")
_REF = os.path.join(ORACLE_DIR, "_ref", "libmh_ref.so")
REF_STOCK = os.path.join(ORACLE_DIR, "_ref", "markovhuffman_stock")
REF_PATCHED = os.path.join(ORACLE_DIR, "_ref", "markovhuffman_patched")

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build():
    """Compile the C restatement (and the reference, when its sources are mounted). Building the checker is
    not using it."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle"])
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


def _load():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(ORACLE_DIR, "mh_oracle.c")):
        build()
    lib = ctypes.CDLL(_LIB)
    lib.mho_histogram.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint8, ctypes.c_int, ctypes.c_void_p]
    lib.mho_histogram.restype = None
    lib.mho_table_from_counts.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.mho_table_from_counts.restype = ctypes.c_void_p
    lib.mho_table_from_bytes.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.mho_table_from_bytes.restype = ctypes.c_void_p
    lib.mho_table_free.argtypes = [ctypes.c_void_p]
    lib.mho_table_free.restype = None
    lib.mho_table_write.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    lib.mho_table_write.restype = ctypes.c_long
    lib.mho_compress.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]
    lib.mho_compress.restype = ctypes.c_long
    lib.mho_decompress.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
    lib.mho_decompress.restype = ctypes.c_long
    lib.mho_encode_shard.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint8, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
    lib.mho_encode_shard.restype = ctypes.c_longlong
    lib.mho_payload_bits.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    lib.mho_payload_bits.restype = ctypes.c_uint64
    lib.mho_synth_markov.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64]
    lib.mho_synth_markov.restype = None
    lib.mho_synth_fibonacci.argtypes = [ctypes.c_int, ctypes.c_uint8, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64]
    lib.mho_synth_fibonacci.restype = None
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _buf(data):
    """bytes / bytearray / np.uint8 array -> (contiguous np array, void*)"""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    return a, a.ctypes.data_as(ctypes.c_void_p)


# ---- struct mirrors (only what the tests read) ---------------------------------------------------------
class _Node(ctypes.Structure):
    _fields_ = [("left", ctypes.c_int16), ("right", ctypes.c_int16), ("is_internal", ctypes.c_uint8),
                ("value", ctypes.c_uint8), ("weight", ctypes.c_int32), ("height", ctypes.c_int32), ("depth", ctypes.c_int32)]


class _Tree(ctypes.Structure):
    _fields_ = [("n_nodes", ctypes.c_int), ("root", ctypes.c_int), ("nodes", _Node * 511),
                ("code_len", ctypes.c_int32 * 256), ("code_bits", (ctypes.c_uint8 * 32) * 256), ("lut", ctypes.c_int16 * 256)]


class _Table(ctypes.Structure):
    _fields_ = [("markov", ctypes.c_int), ("trees", ctypes.POINTER(_Tree))]


class Table:
    """Owns an mho_table*."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("oracle could not build/load the table")
        self.h = handle
        self._t = ctypes.cast(handle, ctypes.POINTER(_Table)).contents

    def __del__(self):
        if getattr(self, "h", None):
            lib().mho_table_free(self.h)
            self.h = None

    @property
    def markov(self):
        return bool(self._t.markov)

    @classmethod
    def from_counts(cls, counts, markov):
        c = np.ascontiguousarray(np.asarray(counts).astype(np.int64).astype(np.int32))  # int32 wrap like the reference (F3)
        assert c.size == (65536 if markov else 256)
        return cls(lib().mho_table_from_counts(c.ctypes.data_as(ctypes.c_void_p), int(bool(markov))))

    @classmethod
    def from_bytes(cls, data):
        a, p = _buf(bytes(data))
        return cls(lib().mho_table_from_bytes(p, len(a)))

    def serialize(self):
        out = np.zeros(256 * (1 + 256 * 10 // 8 + 64) + 64, dtype=np.uint8)
        n = lib().mho_table_write(self.h, out.ctypes.data_as(ctypes.c_void_p), out.size)
        assert n >= 0
        return out[:n].tobytes()

    def code_lengths(self):
        """np.int32 [ntab, 256]"""
        ntab = 256 if self.markov else 1
        return np.array([list(self._t.trees[i].code_len) for i in range(ntab)], dtype=np.int32)

    def code(self, prev, c):
        """(length, bits as a '0'/'1' string)"""
        tr = self._t.trees[prev if self.markov else 0]
        ln = tr.code_len[c]
        bits = bytes(tr.code_bits[c])
        return ln, "".join(str((bits[b // 8] >> (7 - b % 8)) & 1) for b in range(ln))

    def lut(self, prev):
        """list of 256 (kind, value, depth): kind 0 null, 1 leaf, 2 internal (depth-8 node)"""
        tr = self._t.trees[prev if self.markov else 0]
        out = []
        for w in range(256):
            n = tr.lut[w]
            if n < 0:
                out.append((0, 0, 0))
            else:
                nd = tr.nodes[n]
                out.append((2 if nd.is_internal else 1, nd.value, nd.depth))
        return out

    def compress(self, data, return_dropped=False):
        a, p = _buf(data)
        cap = 1 + 32 * len(a) + 16 if len(a) < (1 << 16) else 1 + 8 * len(a) + 16
        out = np.empty(cap, dtype=np.uint8)
        dropped = ctypes.c_uint64(0)
        n = lib().mho_compress(self.h, p, len(a), out.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(dropped))
        assert n >= 1, n
        res = out[:n].tobytes()
        return (res, dropped.value) if return_dropped else res

    def decompress(self, stream, cap=None):
        a, p = _buf(stream)
        cap = cap if cap is not None else max(64, 8 * len(a))
        out = np.empty(cap, dtype=np.uint8)
        n = lib().mho_decompress(self.h, p, len(a), out.ctypes.data_as(ctypes.c_void_p), cap)
        if n < 0:
            raise ValueError("oracle decompress error %d" % n)
        return out[:n].tobytes()

    def encode_shard(self, data, prev0, bit_base):
        """(payload bytes whose first bit sits at bit (bit_base & 7), n_bits) — what one GPU of the sharded path writes."""
        a, p = _buf(data)
        cap = 32 * len(a) + 16
        out = np.zeros(cap, dtype=np.uint8)
        nbits = lib().mho_encode_shard(self.h, p, len(a), prev0, bit_base, out.ctypes.data_as(ctypes.c_void_p), cap)
        assert nbits >= 0
        return out[: ((bit_base & 7) + nbits + 7) // 8].tobytes(), int(nbits)

    def payload_bits(self, data):
        a, p = _buf(data)
        return lib().mho_payload_bits(self.h, p, len(a))


def histogram(data, markov, prev0=0x20):
    a, p = _buf(data)
    counts = np.zeros(65536 if markov else 256, dtype=np.int32)
    lib().mho_histogram(p, len(a), prev0, int(bool(markov)), counts.ctypes.data_as(ctypes.c_void_p))
    return counts


def compress_from_input(data, markov):
    """What `markovhuffman in -o out [-h] -d table` produces: (stream bytes, table bytes)."""
    t = Table.from_counts(histogram(data, markov), markov)
    return t.compress(data), t.serialize()


def synth_markov(trans_counts, seed, seg_bytes, first_seg, n):
    tc = np.ascontiguousarray(trans_counts, dtype=np.uint32)
    assert tc.size == 65536
    out = np.empty(n, dtype=np.uint8)
    lib().mho_synth_markov(tc.ctypes.data_as(ctypes.c_void_p), seed, seg_bytes, first_seg, out.ctypes.data_as(ctypes.c_void_p), n)
    return out.tobytes()


def synth_fibonacci(k, base, seed, first_index, n):
    out = np.empty(n, dtype=np.uint8)
    lib().mho_synth_fibonacci(k, base, seed, first_index, out.ctypes.data_as(ctypes.c_void_p), n)
    return out.tobytes()


# ---- the real reference, through oracle/_ref/libmh_ref.so (absent => None) ------------------------------
_ref = None


def ref():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF):
            return None
        r = ctypes.CDLL(_REF)
        r.ref_table_from_counts.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        r.ref_table_from_counts.restype = ctypes.c_long
        r.ref_codes_from_counts.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        r.ref_codes_from_counts.restype = ctypes.c_int
        r.ref_lut_from_counts.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        r.ref_lut_from_counts.restype = ctypes.c_int
        r.ref_compress_counts.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        r.ref_compress_counts.restype = ctypes.c_long
        r.ref_compress_table.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        r.ref_compress_table.restype = ctypes.c_long
        r.ref_decompress_table.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        r.ref_decompress_table.restype = ctypes.c_long
        _ref = r
    return _ref


def _i32(counts):
    return np.ascontiguousarray(np.asarray(counts).astype(np.int64).astype(np.int32))


def ref_table_from_counts(counts, markov):
    c = _i32(counts)
    out = np.zeros(1 << 17, dtype=np.uint8)
    n = ref().ref_table_from_counts(c.ctypes.data_as(ctypes.c_void_p), int(bool(markov)), out.ctypes.data_as(ctypes.c_void_p), out.size)
    assert n >= 0, n
    return out[:n].tobytes()


def ref_codes_from_counts(counts, markov):
    c = _i32(counts)
    ntab = 256 if markov else 1
    lens = np.zeros(ntab * 256, dtype=np.int32)
    bits = np.zeros(ntab * 256 * 32, dtype=np.uint8)
    ref().ref_codes_from_counts(c.ctypes.data_as(ctypes.c_void_p), int(bool(markov)), lens.ctypes.data_as(ctypes.c_void_p), bits.ctypes.data_as(ctypes.c_void_p))
    return lens.reshape(ntab, 256), bits.reshape(ntab, 256, 32)


def ref_lut_from_counts(counts, markov):
    c = _i32(counts)
    ntab = 256 if markov else 1
    kind = np.zeros(ntab * 256, dtype=np.uint8)
    value = np.zeros(ntab * 256, dtype=np.uint8)
    depth = np.zeros(ntab * 256, dtype=np.int32)
    ref().ref_lut_from_counts(c.ctypes.data_as(ctypes.c_void_p), int(bool(markov)), kind.ctypes.data_as(ctypes.c_void_p),
                              value.ctypes.data_as(ctypes.c_void_p), depth.ctypes.data_as(ctypes.c_void_p))
    return kind.reshape(ntab, 256), value.reshape(ntab, 256), depth.reshape(ntab, 256)


def ref_compress_counts(counts, markov, data):
    c = _i32(counts)
    a, p = _buf(data)
    out = np.empty(1 + 32 * len(a) + 64, dtype=np.uint8)
    n = ref().ref_compress_counts(c.ctypes.data_as(ctypes.c_void_p), int(bool(markov)), p, len(a), out.ctypes.data_as(ctypes.c_void_p), out.size)
    assert n >= 0, n
    return out[:n].tobytes()


def ref_compress_table(table, data):
    t, tp = _buf(table)
    a, p = _buf(data)
    out = np.empty(1 + 32 * len(a) + 64, dtype=np.uint8)
    n = ref().ref_compress_table(tp, len(t), p, len(a), out.ctypes.data_as(ctypes.c_void_p), out.size)
    assert n >= 0, n
    return out[:n].tobytes()


def ref_decompress_table(table, stream):
    t, tp = _buf(table)
    a, p = _buf(stream)
    out = np.empty(max(64, 8 * len(a)), dtype=np.uint8)
    n = ref().ref_decompress_table(tp, len(t), p, len(a), out.ctypes.data_as(ctypes.c_void_p), out.size)
    assert n >= 0, n
    return out[:n].tobytes()
