"""Pins the oracle (oracle/mh_oracle.c): against the golden vectors generated from the stock reference build
(tests/golden/golden.json — SURVEY.md App. C plus edge cases) and, when oracle/_ref is present, against the
reference itself on random count vectors and inputs. CPU only."""
import base64
import hashlib

import numpy as np
import pytest

import oracle_py as o
from conftest import golden_input, load_golden

CASES = load_golden()
IDS = ["%s-%s" % (c["input"], c["mode"]) for c in CASES]


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_matches_reference_golden(case):
    data = golden_input(case["input"])
    markov = case["mode"] == "markov"
    stream, table = o.compress_from_input(data, markov)
    assert len(table) == case["table_bytes"]
    assert sha(table) == case["table_sha256"]
    assert table == base64.b64decode(case["table_b64"])
    assert len(stream) == case["stream_bytes"]
    assert "%02x" % stream[0] == case["header"]
    assert sha(stream) == case["stream_sha256"]
    if "stream_b64" in case:
        assert stream == base64.b64decode(case["stream_b64"])


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_roundtrip_through_table_file(case):
    """-x -e path: load the table FILE (not the in-memory trees), decode, compare with the input; and encoding
    with the loaded table equals encoding with the built one."""
    data = golden_input(case["input"])
    table = base64.b64decode(case["table_b64"])
    if len(table) == 0:
        with pytest.raises(ValueError):
            o.Table.from_bytes(table)
        return
    markov = case["mode"] == "markov"
    stream, _ = o.compress_from_input(data, markov)
    t = o.Table.from_bytes(table)
    assert t.markov == markov
    assert t.serialize() == table
    assert t.compress(data) == stream
    assert t.decompress(stream) == data


def test_appendix_c_hex_vectors():
    """The hand-checkable vectors of SURVEY.md App. C."""
    s, t = o.compress_from_input(b"aaaabbcd", True)
    assert s.hex() == "30f3"
    assert t.hex() == "8000000056" "1b0800000000000000" "0562b0d62b1d64b2" + "00" * 20
    s, t = o.compress_from_input(b"aaaabbcd", False)
    assert s.hex() == "3a0adc" and t.hex() == "5856258ec8"
    s, t = o.compress_from_input(b"aaaabbcdcb", True)
    assert s.hex() == "36f380"
    s, t = o.compress_from_input(b"aaaabbcdcb", False)
    assert s.hex() == "3d0afbc0" and t.hex() == "58562592c6"
    s, t = o.compress_from_input(b"", True)
    assert s.hex() == "30" and t.hex() == "80" + "00" * 32
    s, t = o.compress_from_input(b"", False)
    assert s.hex() == "38" and t == b""
    s, t = o.compress_from_input(b"Z", True)
    assert s.hex() == "3780" and len(t) == 35 and t.hex().startswith("8000000055aad0")
    s, t = o.compress_from_input(b"Z", False)
    assert s.hex() == "3f80" and t.hex() == "56ab40"


def test_header_errors():
    data = golden_input("input_b.txt")
    sm, tm = o.compress_from_input(data, True)
    sh, th = o.compress_from_input(data, False)
    with pytest.raises(ValueError, match="-3"):
        o.Table.from_bytes(tm).decompress(sh)      # coder type mismatch (src/coding.cpp:107-110)
    with pytest.raises(ValueError, match="-2"):
        o.Table.from_bytes(tm).decompress(b"\x80" + sm[1:])   # placeholder header never replaced (:103-106)


def test_unknown_symbol_is_dropped_like_reference():
    t = o.Table.from_counts(o.histogram(b"abracadabra", False), False)
    s_all = t.compress(b"abracadabra")
    s_x, dropped = t.compress(b"abrXacadabra", return_dropped=True)
    assert dropped == 1 and s_x == s_all


# ---------------------------------------------------------------------------------------------------------
# against the real reference (library harness over the reference's own objects)
# ---------------------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(o.ref() is None, reason="oracle/_ref not built (no /root/reference here)")


def _count_vectors():
    rng = np.random.default_rng(7)
    vs = []
    vs.append(("all_equal", np.full(256, 5)))
    vs.append(("ties_small", rng.integers(0, 3, 256)))
    vs.append(("two", np.bincount([65, 66], minlength=256)))
    vs.append(("one", np.bincount([200], minlength=256) * 9))
    fib = np.zeros(256, dtype=np.int64); a, b = 1, 1
    for i in range(44):
        fib[40 + i] = a; a, b = b, a + b
    vs.append(("fib44", fib))                      # codewords up to 43 bits
    vs.append(("geometric", (2.0 ** -np.arange(256) * 1e9).astype(np.int64)))
    big = rng.integers(1, 1 << 28, 256); big[32] = (1 << 31) - 5
    vs.append(("sum_wraps_int32", big))            # internal weights wrap (F3)
    neg = rng.integers(0, 1 << 20, 256).astype(np.int64); neg[101] = (1 << 31) + 12345
    vs.append(("count_wraps_negative", neg))
    for k in range(6):
        vs.append(("rand%d" % k, rng.integers(0, 1 << rng.integers(1, 20), 256) * (rng.random(256) < rng.random())))
    return vs


@needs_ref
@pytest.mark.parametrize("name,counts", _count_vectors(), ids=[n for n, _ in _count_vectors()])
def test_single_tree_vs_reference(name, counts):
    t = o.Table.from_counts(counts, False)
    assert t.serialize() == o.ref_table_from_counts(counts, False)
    lens, bits = o.ref_codes_from_counts(counts, False)
    assert np.array_equal(t.code_lengths(), lens)
    for c in range(256):
        ln, s = t.code(0, c)
        ref_s = "".join(str((int(bits[0, c, b // 8]) >> (7 - b % 8)) & 1) for b in range(ln))
        assert s == ref_s
    kind, value, depth = o.ref_lut_from_counts(counts, False)
    for w, (k, v, d) in enumerate(t.lut(0)):
        assert k == kind[0, w]
        if k:
            assert d == depth[0, w]
        if k == 1:
            assert v == value[0, w]


@needs_ref
def test_markov_random_vs_reference():
    rng = np.random.default_rng(11)
    for trial in range(3):
        k = [4, 40, 256][trial]
        syms = rng.choice(256, size=k, replace=False)
        data = bytes(syms[rng.integers(0, k, 20000) % (1 + rng.integers(0, k, 20000))].astype(np.uint8))
        for markov in (True, False):
            counts = o.histogram(data, markov)
            t = o.Table.from_counts(counts, markov)
            table = t.serialize()
            assert table == o.ref_table_from_counts(counts, markov)
            stream = t.compress(data)
            assert stream == o.ref_compress_counts(counts, markov, data)
            assert stream == o.ref_compress_table(table, data)
            assert o.ref_decompress_table(table, stream) == data
            assert t.decompress(stream) == data


@needs_ref
def test_long_codewords_vs_reference():
    """Fibonacci counts: codewords far beyond 8 bits exercise the depth-8 LUT entries and the tree walk."""
    data = o.synth_fibonacci(40, 48, 99, 0, 200000)
    for markov in (True, False):
        counts = o.histogram(data, markov)
        t = o.Table.from_counts(counts, markov)
        assert t.code_lengths().max() > 16
        stream = t.compress(data)
        assert stream == o.ref_compress_counts(counts, markov, data)
        assert o.ref_decompress_table(t.serialize(), stream) == data
        assert t.decompress(stream) == data


def test_synth_generators_are_deterministic_and_shaped():
    ipsum = golden_input("input_ipsum.txt")
    tc = o.histogram(ipsum, True).astype(np.uint32)
    a = o.synth_markov(tc, 42, 4096, 0, 3 * 4096 + 100)
    b = o.synth_markov(tc, 42, 4096, 1, 4096)
    assert a[4096:8192] == b                          # segments are independent of where generation starts
    assert set(a) <= set(ipsum)
    f = o.synth_fibonacci(40, 48, 1234, 0, 10000)
    g = o.synth_fibonacci(40, 48, 1234, 5000, 5000)
    assert f[5000:] == g
    assert min(f) >= 48 and max(f) < 88
