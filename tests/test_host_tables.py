"""Host logic of the product (tree building, code assignment, table file format) against the oracle and the
reference-built golden vectors. CPU only: no kernel is launched."""
import base64
import hashlib
import os
import re

import numpy as np
import pytest

import oracle_py as o
from conftest import ROOT, golden_input, load_golden
from mhlib import load

mh = load()
CASES = load_golden()
IDS = ["%s-%s" % (c["input"], c["mode"]) for c in CASES]


def _counts(data, markov):
    return o.histogram(data, markov).astype(np.int64).astype(np.uint64)   # oracle counts are only the INPUT here


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_table_file_matches_reference_golden(case):
    data = golden_input(case["input"])
    markov = case["mode"] == "markov"
    p = mh.CodingProvider.from_counts_array(np.ascontiguousarray(_counts(data, markov)), int(markov))
    table = p.write_coding_tree()
    assert hashlib.sha256(table).hexdigest() == case["table_sha256"]
    assert table == base64.b64decode(case["table_b64"])
    assert p.get_type() == int(markov)
    if table:
        q = mh.CodingProvider.from_table_file(table)          # loader is the writer's mirror (left subtree first)
        assert q.get_type() == int(markov)
        assert q.write_coding_tree() == table
        ot = o.Table.from_counts(o.histogram(data, markov), markov)
        for prev in ([0x20, data[0]] if markov and data else [0]):
            for c in range(256):
                assert q.get_encoding(prev, c) == p.get_encoding(prev, c) == ot.code(prev, c)
    else:
        with pytest.raises(mh.MhError):
            mh.CodingProvider.from_table_file(table)


def _count_vectors():
    rng = np.random.default_rng(3)
    vs = [("all_equal", np.full(256, 7)), ("two", np.bincount([1, 255], minlength=256)), ("one", np.bincount([9], minlength=256) * 3)]
    fib = np.zeros(256, dtype=np.int64); a, b = 1, 1
    for i in range(44):
        fib[100 + i] = a; a, b = b, a + b
    vs.append(("fib44", fib))
    big = rng.integers(1, 1 << 28, 256); big[32] = (1 << 31) - 5
    vs.append(("sum_wraps_int32", big))
    neg = rng.integers(0, 1 << 20, 256).astype(np.int64); neg[101] = (1 << 31) + 12345
    vs.append(("count_wraps_negative", neg))
    wide = rng.integers(0, 1 << 40, 256)
    wide[wide % (1 << 32) == 0] += 1
    vs.append(("counts_beyond_32_bits", wide))
    for k in range(8):
        vs.append(("rand%d" % k, rng.integers(0, 1 << rng.integers(1, 24), 256) * (rng.random(256) < rng.random())))
    return vs


@pytest.mark.parametrize("name,counts", _count_vectors(), ids=[n for n, _ in _count_vectors()])
def test_single_tree_vs_oracle(name, counts):
    """Tie-breaking, int32 wrap, LUT semantics: product host code == oracle (itself pinned to the reference)."""
    cu = np.ascontiguousarray(np.asarray(counts, dtype=np.int64).astype(np.uint64))
    p = mh.CodingProvider.from_counts_array(cu, 0)
    ot = o.Table.from_counts(counts, False)
    assert p.write_coding_tree() == ot.serialize()
    for c in range(256):
        assert p.get_encoding(0, c) == ot.code(0, c)
    for w, (kind, value, depth) in enumerate(ot.lut(0)):
        k2, v2, d2 = p.decoding_lookup(0, w)
        assert k2 == kind
        if kind:
            assert d2 == depth
        if kind == 1:
            assert v2 == value
    if o.ref() is not None:
        assert p.write_coding_tree() == o.ref_table_from_counts(counts, False)


def test_count_that_wraps_to_zero_is_refused():
    counts = np.zeros(256, dtype=np.uint64)
    counts[65] = 1 << 32
    counts[66] = 5
    with pytest.raises(mh.MhError) as e:
        mh.CodingProvider.from_counts_array(counts, 0)
    assert e.value.status == mh.MH_ERR_COUNT_WRAPPED


def test_malformed_table_files():
    good = base64.b64decode(next(c for c in CASES if c["input"] == "input_ipsum.txt" and c["mode"] == "markov")["table_b64"])
    for cut in (1, 5, 33, len(good) // 2):
        with pytest.raises(mh.MhError) as e:
            mh.CodingProvider.from_table_file(good[:cut])
        assert e.value.status == mh.MH_ERR_BAD_TABLE
    with pytest.raises(mh.MhError):
        mh.CodingProvider.from_table_file(b"\x40\x80")      # -h file whose root is a bare leaf: the writer never emits it
    with pytest.raises(mh.MhError):
        mh.CodingProvider.from_table_file(b"\x00" * 64)     # internal nodes forever: depth limit


def test_debug_dump_matches_reference_cli(tmp_path):
    """-g output (print_table + print_tree) byte for byte, when the reference binary is available."""
    import subprocess
    if not os.path.exists(o.REF_STOCK):
        pytest.skip("oracle/_ref not built")
    for name in ("input_b.txt", "input_ipsum.txt"):
        for markov in (True, False):
            src = os.path.join(ROOT, "tests/golden/inputs", name)
            out = subprocess.run([o.REF_STOCK, src, "-o", str(tmp_path / "o"), "-g"] + ([] if markov else ["-h"]),
                                 stdout=subprocess.PIPE, stderr=subprocess.PIPE).stdout
            p = mh.CodingProvider.from_counts_array(np.ascontiguousarray(_counts(golden_input(name), markov)), int(markov))
            assert p.debug_dump() == out


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports exactly what include/mh_gpu.h declares (no compute call is made)."""
    header = open(os.path.join(ROOT, "include/mh_gpu.h")).read()
    declared = set(re.findall(r"\b(mh_[a-z0-9_]+)\s*\(", header))
    declared -= {"mh_status"}
    import ctypes
    lib = ctypes.CDLL(mh.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libmh_gpu.so does not export %s" % name
    assert declared == set(mh._SIGNATURES), declared ^ set(mh._SIGNATURES)
    assert mh._lib.mh_status_string(mh.MH_ERR_BAD_HEADER) == b"Input appears corrupt"
    assert mh._lib.mh_version() >= 100


def test_no_device_fails_loudly():
    """Without a usable GPU every compute entry point must raise, never fall back to the CPU."""
    if mh.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(mh.MhError) as e:
        mh.Session(1 << 20)
    assert e.value.status in (mh.MH_ERR_NO_DEVICE, mh.MH_ERR_CUDA)
    with pytest.raises(mh.MhError):
        mh.compress(b"hello")
    with pytest.raises(mh.MhError) as e:
        mh.device_memory(0)
    assert e.value.status in (mh.MH_ERR_NO_DEVICE, mh.MH_ERR_CUDA)


def _pair_decode(pl, bits, n_bits, prev0, markov, n_symbols):
    """Decode with the two-symbol table exactly as the kernels do on their fast path (entry by entry, prefix rows
    included); flagged entries are not expected in these inputs."""
    table, rank, live, len1, rows, ctx_rows = pl
    out = bytearray()
    row = int(rank[prev0]) if markov else 0
    pos = 0
    while len(out) < n_symbols:
        w = 0
        for k in range(8):
            p = pos + k
            bit = (bits[p >> 3] >> (7 - (p & 7))) & 1 if p < n_bits else 0
            w = (w << 1) | bit
        e = int(table[row * 256 + w])
        assert not (e & 0x30), "flagged entry on a stream whose codewords are <= 16 bits"
        cnt = (e >> 6) & 15
        assert cnt in (0, 1, 2)
        if cnt >= 1:
            out.append((e >> 16) & 255)
            if row < ctx_rows:
                l1 = int(len1[row * 256 + w])
                assert 1 <= l1 <= (e & 15) - (1 if cnt == 2 else 0)
                if cnt == 1:
                    assert l1 == (e & 15)
        if cnt == 2:
            out.append((e >> 24) & 255)
        else:
            assert (e >> 24) == 0
        pos += e & 15
        row = (e >> 10) & 63
        assert row < rows
    return bytes(out[:n_symbols]), pos


@pytest.mark.parametrize("name", ["input_ipsum.txt", "input_wiki_cpp.txt", "input_b.txt"])
@pytest.mark.parametrize("markov", [True, False], ids=["markov", "huffman"])
def test_pair_table_decodes_like_the_oracle(name, markov):
    """flatten_pairlut (the decoder's shared-memory table): decoding the oracle's stream entry by entry through it gives
    back the input — pairs, single symbols and the prefix rows of 9..16-bit codewords."""
    data = golden_input(name)
    p = mh.CodingProvider.from_counts_array(np.ascontiguousarray(_counts(data, markov)), int(markov))
    pl = p.pair_lut()
    if markov and len(set(data) | {0x20}) > 63:
        assert pl is None            # too many live contexts for the 6-bit row field
        return
    assert pl is not None
    table, rank, live, len1, rows, ctx_rows = pl
    assert ctx_rows < rows <= 64
    stream, _ = o.compress_from_input(data, markov)
    payload = stream[1:]
    n_bits = len(payload) * 8 - (stream[0] & 7)      # header: 0 0 1 1 E R R R, R = padding bits (src/coding.cpp:88)
    got, pos = _pair_decode(pl, payload, n_bits, 0x20, markov, len(data))
    assert got == data
    assert pos == n_bits
    if markov:
        for c in range(256):
            r = int(rank[c])
            assert r <= ctx_rows
            if r < ctx_rows:
                assert int(live[r]) == c


def test_pair_table_absent_for_many_contexts():
    data = golden_input("input_wiki_cpp.html")
    p = mh.CodingProvider.from_counts_array(np.ascontiguousarray(_counts(data, True)), 1)
    assert p.pair_lut() is None      # 188 live contexts: the decoder keeps the 8-bit LUT
    q = mh.CodingProvider.from_counts_array(np.ascontiguousarray(_counts(data, False)), 0)
    assert q.pair_lut() is not None  # one tree: always available


def test_pair_table_entry_kinds_follow_the_8bit_lut():
    """With long codewords (counts spread over six orders of magnitude) the pair table runs out of prefix rows: every
    context-row entry must then be exactly one of leaf entry (LUT leaf), prefix entry or deep flag (LUT internal node at
    depth 8), null flag (no LUT entry), and the prefix rows must have gone to the heaviest deep nodes."""
    data = golden_input("input_ipsum.txt")
    counts = _counts(data, True).astype(np.float64)
    live = counts > 0
    counts[live] = np.maximum(1, (counts[live] ** 2.2)).astype(np.float64)      # stretch the distribution: deep trees
    counts = counts.astype(np.uint64)
    p = mh.CodingProvider.from_counts_array(np.ascontiguousarray(counts), 1)
    assert p.max_code_bits() > 12
    table, rank, live_ctx, len1, rows, ctx_rows = p.pair_lut()
    t = table.reshape(rows, 256)
    flagged_weight, prefix_weight = [], []
    n_prefix = 0
    for prev in range(256):
        r = int(rank[prev])
        if r >= ctx_rows:
            continue
        for w in range(256):
            kind, value, depth = p.decoding_lookup(prev, w)
            e = int(t[r, w])
            if kind == 0:
                assert e & 0x20
            elif kind == 1:
                assert not (e & 0x30) and ((e >> 6) & 15) in (1, 2) and ((e >> 16) & 255) == value and int(len1[r * 256 + w]) == depth
            else:
                assert (e & 0x10) or (((e >> 6) & 15) == 0 and (e & 63) == 8 and ctx_rows < ((e >> 10) & 63) < rows)
                if not (e & 0x10):
                    n_prefix += 1
    assert n_prefix == rows - ctx_rows - 1 > 0      # every prefix row is reached from exactly one deep node
    assert rows <= 64


def test_pair_table_prefix_rows_of_a_fibonacci_tree():
    """A Fibonacci-shaped tree (codewords up to ~30 bits): its depth-8 node gets a prefix row although not every codeword
    below it ends within 16 bits — those that do are one-symbol entries, the others deep flags naming the tree node reached
    and the context ([31:24] + bit 6, [23:16]; the next-row field stays 0: speculative lookups behind a flag stay in the table) — and a stream of <= 16-bit codewords decodes through the table without a flag."""
    fib = np.zeros(256, dtype=np.uint64)
    a, b = 1, 1
    for i in range(32):
        fib[65 + i] = a
        a, b = b, a + b
    p = mh.CodingProvider.from_counts_array(fib, 0)
    assert p.max_code_bits() > 24
    table, rank, live, len1, rows, ctx_rows = p.pair_lut()
    assert ctx_rows == 1 and rows == 3            # the context row, the null row, one prefix row
    t = table.reshape(rows, 256)
    prefixes = [w for w in range(256) if not (int(t[0, w]) & 0x30) and ((int(t[0, w]) >> 6) & 15) == 0]
    assert len(prefixes) == 1 and (int(t[0, prefixes[0]]) & 63) == 8 and ((int(t[0, prefixes[0]]) >> 10) & 63) == 2
    singles = flags = 0
    for w in range(256):
        e = int(t[2, w])
        if e & 0x10:
            flags += 1
            assert ((e >> 24) | ((e & 0x40) << 2)) < 2 * 32 - 1 and ((e >> 16) & 255) == 0 and not (e & 0x20) and ((e >> 10) & 63) == 0
        else:
            singles += 1
            assert ((e >> 6) & 15) == 1 and 1 <= (e & 15) <= 8 and ((e >> 10) & 63) == 0 and 65 <= ((e >> 16) & 255) < 97
    assert singles > 0 and flags > 0
    # symbols whose codewords stay within 16 bits: decoded through the table as the kernels' fast path does
    lens = p.code_lengths()
    short = [c for c in range(65, 97) if 0 < int(lens[c]) <= 16]
    rng = np.random.default_rng(3)
    data = bytes(rng.choice(short, 4000).astype(np.uint8))
    stream = o.Table.from_counts(fib.astype(np.int64), False).compress(data)
    payload = stream[1:]
    got, pos = _pair_decode((table, rank, live, len1, rows, ctx_rows), payload, len(payload) * 8 - (stream[0] & 7), 0x20, False, len(data))
    assert got == data
