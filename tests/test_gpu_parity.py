"""Parity of the CUDA path against the oracle and the reference-built golden vectors. Everything goes through the
C ABI (include/mh_gpu.h) via the ctypes mirror. Bit-exact: integer / byte work has no tolerance."""
import base64
import hashlib

import numpy as np
import pytest

import oracle_py as o
from conftest import golden_input, load_golden
from mhlib import load

pytestmark = pytest.mark.gpu

mh = load()
CASES = load_golden()
IDS = ["%s-%s" % (c["input"], c["mode"]) for c in CASES]


@pytest.fixture(scope="module")
def session():
    s = mh.Session(72 << 20)
    yield s
    s.close()


@pytest.fixture(scope="module")
def ipsum_counts():
    return o.histogram(golden_input("input_ipsum.txt"), True).astype(np.uint32)


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.fixture
def tunables():
    """Set library tunables for one test (mh_tunable_set); everything goes back to its default afterwards."""
    touched = []

    def set_(name, value):
        touched.append(name)
        mh.tunable_set(name, int(value))

    yield set_
    for name in touched:
        mh.tunable_set(name, -1)


# ---- config 1: the reference's own corpus, -d then -x, both modes ---------------------------------------
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_corpus_compress_matches_reference_bytes(session, case):
    data = golden_input(case["input"])
    order = int(case["mode"] == "markov")
    stream, provider = session.compress(data, order)
    table = provider.write_coding_tree()
    assert len(table) == case["table_bytes"] and sha(table) == case["table_sha256"]
    assert len(stream) == case["stream_bytes"]
    assert "%02x" % stream[0] == case["header"]
    assert sha(stream) == case["stream_sha256"]
    if "stream_b64" in case:
        assert stream == base64.b64decode(case["stream_b64"])


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_corpus_extract_restores_input(session, case):
    data = golden_input(case["input"])
    table = base64.b64decode(case["table_b64"])
    if not table:
        pytest.skip("empty -h table cannot be loaded (neither can the reference)")
    markov = case["mode"] == "markov"
    stream = o.Table.from_bytes(table).compress(data)           # reference-identical stream (pinned by test_oracle)
    assert sha(stream) == case["stream_sha256"]
    provider = mh.CodingProvider.from_table_file(table)          # the -e path
    assert session.decompress(provider, stream) == data
    assert session.compress_with_table(provider, data)[0] == stream


@pytest.mark.parametrize("order", [0, 1])
def test_histogram_matches_oracle(session, order):
    rng = np.random.default_rng(5)
    for data in (golden_input("input_wiki_cpp.html"), golden_input("edge_random_64k.bin"), b"", b"x",
                 bytes(rng.integers(0, 256, 1 << 20, dtype=np.uint8)), bytes(rng.integers(97, 101, 300001, dtype=np.uint8))):
        got = session.histogram(data, order)
        want = o.histogram(data, bool(order)).astype(np.int64).astype(np.uint64)
        assert np.array_equal(got, want)


# ---- sizes around every internal boundary (16-byte loads, 4 KiB rounds, tiles, subsequences, chunks) -----
SIZES = [1, 2, 15, 16, 17, 31, 33, 255, 4095, 4096, 4097, 8191, 8192, 16383, 16384, 16385, 65536 + 3, 262144 + 17, 1048576 + 5]


@pytest.mark.parametrize("order", [0, 1])
def test_ragged_sizes_roundtrip_and_match_oracle(session, ipsum_counts, order):
    text = o.synth_markov(ipsum_counts, 77, 4096, 0, max(SIZES))
    for n in SIZES:
        data = text[:n]
        stream, provider = session.compress(data, order)
        want_stream, want_table = o.compress_from_input(data, bool(order))
        assert stream == want_stream, "n=%d" % n
        assert provider.write_coding_tree() == want_table
        assert session.decompress(provider, stream) == data, "n=%d" % n


# ---- configs 2-4 at sizes the oracle finishes in seconds --------------------------------------------------
@pytest.mark.parametrize("order", [0, 1])
def test_markov_text_8mib(session, ipsum_counts, order):
    data = o.synth_markov(ipsum_counts, 20261018, 65536, 0, 8 << 20)
    stream, provider = session.compress(data, order)
    want_stream, want_table = o.compress_from_input(data, bool(order))
    assert provider.write_coding_tree() == want_table
    assert sha(stream) == sha(want_stream)
    assert session.decompress(provider, stream) == data


@pytest.mark.parametrize("order", [0, 1])
def test_fibonacci_long_codewords_4mib(session, order):
    """Codewords beyond 8 bits: depth-8 LUT entries + tree walk in the decoder, 2-word pushes in the encoder."""
    data = o.synth_fibonacci(40, 48, 1234, 0, 4 << 20)
    stream, provider = session.compress(data, order)
    assert provider.max_code_bits() > 16
    want_stream, want_table = o.compress_from_input(data, bool(order))
    assert provider.write_coding_tree() == want_table
    assert sha(stream) == sha(want_stream)
    assert session.decompress(provider, stream) == data


@pytest.mark.parametrize("order", [0, 1])
def test_binary_data_all_256_symbols(session, order):
    rng = np.random.default_rng(99)
    skew = (rng.integers(0, 256, 3 << 20) * rng.integers(0, 256, 3 << 20) >> 8).astype(np.uint8)   # skewed, K = 256
    data = bytes(skew)
    stream, provider = session.compress(data, order)
    want_stream, want_table = o.compress_from_input(data, bool(order))
    assert provider.write_coding_tree() == want_table
    assert sha(stream) == sha(want_stream)
    assert session.decompress(provider, stream) == data


def test_codewords_longer_than_32_bits(session):
    """A foreign (-e) table with 44-deep Fibonacci codes: the encoder's long-code path, the decoder's deep walk."""
    fib = np.zeros(256, dtype=np.uint64); a, b = 1, 1
    for i in range(44):
        fib[60 + i] = a; a, b = b, a + b
    provider = mh.CodingProvider.from_counts_array(fib, 0)
    assert provider.max_code_bits() == 43
    rng = np.random.default_rng(1)
    data = bytes((60 + rng.integers(0, 44, 200000)).astype(np.uint8))      # uniform over symbols: long codes are common
    ot = o.Table.from_counts(fib.astype(np.int64), False)
    stream, dropped = session.compress_with_table(provider, data)
    assert dropped == 0
    assert stream == ot.compress(data)
    assert session.decompress(provider, stream) == data


def test_symbols_missing_from_a_foreign_table_are_dropped_like_the_reference(session):
    provider = mh.CodingProvider.from_counts_array(o.histogram(b"abracadabra" * 50, True).astype(np.uint64), 1)
    ot = o.Table.from_counts(o.histogram(b"abracadabra" * 50, True), True)
    data = b"abraXcadabra" * 3000 + b"ZZZZ" * 5000 + b"abra"
    stream, dropped = session.compress_with_table(provider, data)
    want, want_dropped = ot.compress(data, return_dropped=True)
    assert dropped == want_dropped > 0
    assert stream == want


def test_header_checks(session):
    data = golden_input("input_b.txt")
    sm, pm = session.compress(data, 1)
    sh_, ph = session.compress(data, 0)
    with pytest.raises(mh.MhError) as e:
        session.decompress(pm, sh_)
    assert e.value.status == mh.MH_ERR_TYPE_MISMATCH
    with pytest.raises(mh.MhError) as e:
        session.decompress(pm, b"\x80" + sm[1:])
    assert e.value.status == mh.MH_ERR_BAD_HEADER


def test_single_symbol_contexts(session):
    for data in (b"q" * 100000, b"ab" * 70000, b"Z"):
        for order in (0, 1):
            stream, provider = session.compress(data, order)
            assert stream == o.compress_from_input(data, bool(order))[0]
            assert session.decompress(provider, stream) == data


@pytest.mark.parametrize("fmt", ["1", "2"])
def test_encoder_table_formats_agree(session, ipsum_counts, fmt, tunables):
    """The encoder picks a table format from the codebook (u32 box in shared memory / u32 box in global memory /
    u64 wide entries). Force the two fallbacks on data that would normally take the first."""
    data = o.synth_markov(ipsum_counts, 5, 65536, 0, (3 << 20) + 77)
    want = {order: o.compress_from_input(data, bool(order))[0] for order in (0, 1)}
    tunables("enc_fmt", fmt)
    for order in (0, 1):
        assert session.compress(data, order)[0] == want[order]


def test_warp_private_encoder_agrees(session, ipsum_counts, tunables):
    """The opt-in encoder with warp-private tiles (enc_warp = 1: one scan warp, tile groups, per-warp staging) writes the
    same bytes: text at sizes around its 1 KiB tiles and 64-tile groups, with and without bulk-copied input, a
    Fibonacci stream (codewords beyond 16 bits escape to the wide table), dropped symbols (tiles of very few bits)."""
    text = o.synth_markov(ipsum_counts, 31, 65536, 0, (5 << 20) + 1029)
    want = {(n, order): o.compress_from_input(text[:n], bool(order))[0]
            for n in (1, 1023, 1024, 1025, 65536, 65537, (1 << 20) + 3, len(text)) for order in (0, 1)}
    fib = o.synth_fibonacci(40, 48, 7, 0, (2 << 20) + 5)
    want_fib = {order: o.compress_from_input(fib, bool(order))[0] for order in (0, 1)}
    provider = mh.CodingProvider.from_counts_array(o.histogram(b"abracadabra" * 50, True).astype(np.uint64), 1)
    holes = b"abraXcadabra" * 3000 + b"ZZZZ" * 5000 + b"abra"
    want_holes = o.Table.from_counts(o.histogram(b"abracadabra" * 50, True), True).compress(holes, return_dropped=True)
    tunables("enc_warp", "1")
    for tma in ("1", "0"):
        tunables("enc_tma", tma)
        for (n, order), stream in want.items():
            assert session.compress(text[:n], order)[0] == stream, "n=%d order=%d tma=%s" % (n, order, tma)
        for order in (0, 1):
            assert session.compress(fib, order)[0] == want_fib[order]
        assert session.compress_with_table(provider, holes) == want_holes


def test_optimistic_64_symbol_encoder_and_its_fail_over(session, ipsum_counts, tunables):
    """Device-built tables are encoded by an optimistic launch with 64 symbols per thread (30 KiB tiles, a staging area
    sized by what shared memory leaves, not by the worst case) and an ordinary launch queued behind it that leaves at
    once unless the first one declined or failed. Same bytes as the oracle: around the 30 KiB tiles, with the optimistic
    launch switched off (enc_spt = 32), when a tile's bits overflow the staging area (fail-over in mid-stream), and
    when the tables are declined up front (mean codeword above 5.5 bits)."""
    text = o.synth_markov(ipsum_counts, 64, 65536, 0, (3 << 20) + 77)
    sizes = (30719, 30720, 30721, 61440, 61445, 92160 + 63, len(text))
    want = {(n, order): o.compress_from_input(text[:n], bool(order))[0] for n in sizes for order in (0, 1)}
    for spt in ("-1", "32"):
        tunables("enc_spt", spt)
        for (n, order), stream in want.items():
            assert session.compress(text[:n], order)[0] == stream, "n=%d order=%d enc_spt=%s" % (n, order, spt)
    tunables("enc_spt", "-1")
    # one whole tile of rare symbols (14-bit codewords in -h mode: more than the staging area holds) inside a stream whose
    # mean codeword is about two bits: the optimistic launch fails in mid-stream
    rng = np.random.default_rng(15)
    ladder = np.array([2.0 ** -(i + 1) for i in range(7)])
    ladder[-1] *= 2
    body = rng.choice(np.arange(65, 72, dtype=np.uint8), size=4 << 20, p=ladder)   # codewords of 1..7 bits
    rare = np.tile(np.arange(16, 256, dtype=np.uint8), 128)
    rare = np.concatenate([rare[rare < 65], rare[rare > 71]])                       # 233 rare values: 14-bit codewords
    rare = np.tile(rare, 2)[:30720]
    body[3 * 30720: 4 * 30720] = rare                                               # ... filling the stream's fourth tile
    data = bytes(body)
    stream, provider = session.compress(data, 0)
    lens = provider.code_lengths()
    assert int(lens[rare].sum()) > (112 * 1024 - 32768 - 30720 - 64) * 8, "the tile was meant to overflow the staging area"
    assert float(np.dot(np.bincount(body, minlength=256), lens)) / len(data) < 5.5, "the tables were meant to be accepted"
    want_stream, want_table = o.compress_from_input(data, False)
    assert provider.write_coding_tree() == want_table
    assert sha(stream) == sha(want_stream)
    assert session.decompress(provider, stream) == data
    stream1, provider1 = session.compress(data, 1)
    assert sha(stream1) == sha(o.compress_from_input(data, True)[0])
    # uniform bytes: 8-bit codewords, declined up front
    noise = bytes(rng.integers(0, 256, 1 << 20, dtype=np.uint8))
    assert sha(session.compress(noise, 0)[0]) == sha(o.compress_from_input(noise, False)[0])


@pytest.mark.parametrize("threads", ["64", "128", "256"])
def test_small_write_ctas_keep_speculative_lookups_inside_the_table(session, threads, tunables):
    """D4 sizes its CTAs by the number of subsequences (small streams: small CTAs on every SM). With little shared
    memory behind the pair table a speculative lookup behind a flagged prefix-row entry must not leave the table: such
    entries keep their next-row field 0 (a -h table of HTML has them: 18-bit codewords)."""
    tunables("dec_write_threads", threads)
    data = golden_input("input_wiki_cpp.html")
    for order in (0, 1):
        stream, provider = session.compress(data, order)
        assert session.decompress(provider, stream) == data
    fib = o.synth_fibonacci(40, 48, 11, 0, (1 << 20) + 7)
    for order in (0, 1):
        stream, provider = session.compress(fib, order)
        assert session.decompress(provider, stream) == fib


@pytest.mark.parametrize("pair", ["1", "0"], ids=["pair-table", "lut8"])
@pytest.mark.parametrize("sub_bits", ["256", "1024", "8192"])
def test_decoder_subsequence_sizes_agree(ipsum_counts, sub_bits, pair, tunables):
    """The subsequence size only changes how the work is cut, never the bytes — through the two-symbol pair table
    (text has few live contexts) and, forced, through the reference's 8-bit LUT + tree walk."""
    tunables("dec_sub_bits_markov", sub_bits)
    tunables("dec_sub_bits_huffman", sub_bits)
    tunables("dec_pair", pair)
    s = mh.Session(8 << 20)
    try:
        text = o.synth_markov(ipsum_counts, 9, 4096, 0, (2 << 20) + 123)
        fib = o.synth_fibonacci(40, 48, 4321, 0, 1 << 20)
        for data in (text, fib):
            for order in (0, 1):
                stream, provider = s.compress(data, order)
                assert s.decompress(provider, stream) == data
    finally:
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("order", [1, 0], ids=["markov", "huffman"])
@pytest.mark.parametrize("kind", ["text", "fib"])
def test_session_streams_inputs_larger_than_its_buffer(ipsum_counts, order, kind):
    """an input that does not fit the session's device buffer goes through it in chunks — histogram counts add
    up with the byte before each chunk as its context, every chunk is encoded at its global bit offset and OR-merged at
    the byte it shares with its neighbour. Same bytes as one pass. Extraction of a stream larger than the buffer likewise."""
    n = 1_000_003
    data = o.synth_markov(ipsum_counts, 11, 4096, 0, n) if kind == "text" else o.synth_fibonacci(40, 48, 99, 0, n)
    want_stream, want_table = o.compress_from_input(data, bool(order))
    s = mh.Session(100_003)          # ten ragged chunks, seams in the middle of bytes
    try:
        assert np.array_equal(s.histogram(data, order), o.histogram(data, bool(order)).astype(np.int64).astype(np.uint64))
        stream, provider = s.compress(data, order)
        assert stream == want_stream
        assert provider.write_coding_tree() == want_table
        again, dropped = s.compress_with_table(provider, data)
        assert again == want_stream and dropped == 0
        # ... and back: the stream is larger than the compressed-side buffer too, so it is decoded in bit-range chunks,
        # each from the exact state (bit position, previous symbol) its predecessor ended in
        assert s.decompress(provider, stream) == data
    finally:
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", ["4096", "20000", "65536", "1000000"])
def test_pipelined_extract_into_a_host_buffer(ipsum_counts, chunk, tunables):
    """Extraction straight into a host buffer runs as a pipeline of bit-range chunks (H2D of chunk k + 1, decode of k and
    D2H of k - 1 overlap): the bytes do not depend on where the chunks are cut."""
    tunables("pipe_min_bytes", 1)
    tunables("pipe_chunk_bytes", chunk)
    s = mh.Session(3 << 20)
    try:
        for data in (o.synth_markov(ipsum_counts, 21, 4096, 0, (1 << 20) + 4321), o.synth_fibonacci(40, 48, 7, 0, 600_001)):
            for order in (1, 0):
                stream, provider = s.compress(data, order)
                assert s.decompress_into(provider, stream, len(data) + 5) == data
                with pytest.raises(mh.MhError) as e:
                    s.decompress_into(provider, stream, len(data) - 1)
                assert e.value.status == mh.MH_ERR_CAPACITY
    finally:
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", ["4096", "50000", "262144", "1048576"])
def test_pipelined_compress_from_a_host_buffer(ipsum_counts, chunk, tunables):
    """Compression of a large host buffer runs as a pipeline: the input travels in chunks while the histogram of the chunk
    before it runs (counts accumulate on the device), then the chunks are encoded at their global bit offsets while the
    payload of the chunk before travels back; bytes shared by two chunks are OR-merged. Same bytes as one pass, wherever
    the chunks are cut, with a built table and with a given one (-e)."""
    tunables("pipe_min_bytes", 1)
    tunables("enc_pipe_chunk_bytes", chunk)
    s = mh.Session(3 << 20)
    try:
        for data in (o.synth_markov(ipsum_counts, 23, 4096, 0, (1 << 20) + 4321), o.synth_fibonacci(40, 48, 8, 0, 600_001), b"", b"x"):
            for order in (1, 0):
                want_stream, want_table = o.compress_from_input(data, bool(order))
                stream, provider = s.compress(data, order)
                assert stream == want_stream
                assert provider.write_coding_tree() == want_table
                again, dropped = s.compress_with_table(provider, data)
                assert again == want_stream and dropped == 0
    finally:
        s.close()


@pytest.mark.gpu
def test_extract_larger_than_the_uncompressed_side_buffer(ipsum_counts):
    """The session never reallocates: a stream that decodes to more bytes than its uncompressed-side buffer holds is
    decoded in bit-range chunks that fit (halved until they do). The size query reports the size, mh_session_fetch
    answers MH_ERR_WORKSPACE (nothing resident), the call with a host buffer delivers the bytes."""
    import ctypes
    data = b"a" * 3_000_000 + o.synth_markov(ipsum_counts, 31, 4096, 0, 500_000)   # ~1 bit per symbol, then text
    big = mh.Session(4 << 20)
    stream, provider = big.compress(data, 1)
    big.close()
    assert len(stream) < 700_000
    h = ctypes.c_void_p()
    assert mh._lib.mh_session_create_sized(0, 300_000, len(stream) + 64, ctypes.byref(h)) == 0
    try:
        src = np.frombuffer(stream, dtype=np.uint8)
        n = ctypes.c_uint64(0)
        assert mh._lib.mh_session_decompress(h, provider._h, src.ctypes.data, src.size, None, 0, ctypes.byref(n)) == 0
        assert n.value == len(data)
        out = np.zeros(len(data), dtype=np.uint8)
        got = ctypes.c_uint64(1)
        assert mh._lib.mh_session_fetch(h, out.ctypes.data, out.size, ctypes.byref(got)) == mh.MH_ERR_WORKSPACE and got.value == 0
        assert mh._lib.mh_session_decompress(h, provider._h, src.ctypes.data, src.size, out.ctypes.data, out.size, ctypes.byref(n)) == 0
        assert n.value == len(data) and out.tobytes() == data
    finally:
        mh._lib.mh_session_destroy(h)


@pytest.mark.gpu
def test_payload_longer_than_2_pow_32_bits():
    """The reference's own decoder breaks at 2^31 payload bits (int length / int bi, src/coding.cpp:115,120).
    600 MB of K = 256 data code to ~4.8 Gbit: the stream must equal the oracle's (64-bit restatement) byte for byte
    and extract to the input — through the device API (one launch each) and through a session that has to chunk it."""
    import torch
    n = 600_000_000
    rng = np.random.default_rng(2026)
    data = rng.integers(0, 256, n, dtype=np.uint8)
    data[::3] &= 0x3F                      # skew the statistics a little: codewords of 6..10 bits, the LUT8 fallback included
    want_stream, want_table = o.compress_from_input(data, True)
    assert (len(want_stream) - 1) * 8 > (1 << 32)
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(data).to(dev)
    cap = n + n // 8 + 4096
    d_pay = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(n, dtype=torch.uint8, device=dev)
    d_counts = torch.zeros(65536, dtype=torch.int64, device=dev)
    d_res = torch.zeros(8, dtype=torch.int64, device=dev)
    ws = mh.Workspace(n, cap)
    mh.gpu_histogram(d_in.data_ptr(), n, 0x20, 1, d_counts.data_ptr(), ws)
    provider = mh.CodingProvider.from_counts_array(d_counts.cpu().numpy().view(np.uint64), 1)
    assert provider.write_coding_tree() == want_table
    book, dectab = mh.Codebook(provider), mh.DecodeTable(provider)
    mh.gpu_encode(d_in.data_ptr(), n, 0x20, book, 0, d_pay.data_ptr(), cap, d_res.data_ptr(), ws)
    bits = int(d_res[0].item())
    assert bits > (1 << 32) and int(d_res[2].item()) == 0
    nbytes = (bits + 7) // 8
    assert 1 + nbytes == len(want_stream)
    got = hashlib.sha256(bytes([0x30 | ((8 - bits % 8) % 8)]))
    got.update(memoryview(d_pay[:nbytes].cpu().numpy()))
    assert got.hexdigest() == sha(want_stream)
    mh.gpu_decode(d_pay.data_ptr(), 0, bits, 0x20, dectab, d_out.data_ptr(), n, d_res[4:].data_ptr(), ws)
    assert d_res[4:7].tolist() == [n, 0, 0]
    assert torch.equal(d_out, d_in)
    del d_pay, d_out, ws, book, dectab
    torch.cuda.empty_cache()
    # the host-buffer path with device buffers a quarter of the size: chunked encode at bit offsets beyond 2^32,
    # chunked extract from exact states at bit positions beyond 2^32
    s = mh.Session(160_000_000)
    try:
        stream, _ = s.compress(data, 1)
        assert sha(stream) == sha(want_stream)
        assert np.array_equal(np.frombuffer(s.decompress_into(provider, stream, n), dtype=np.uint8), data)
    finally:
        s.close()


@pytest.mark.gpu
def test_large_text_takes_the_8192_bit_subsequences_and_equals_the_oracle(ipsum_counts):
    """200 MB of Markov text is long enough for the decoder to pick its largest subsequences (8192 bits, what the 1 GiB
    bench runs) without any override; stream == oracle, extract == input, both coder types."""
    import torch
    n = 200_000_000
    dev = torch.device("cuda", 0)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    mh.synth_markov(ipsum_counts, 4242, 65536, 0, d_in.data_ptr(), n)
    torch.cuda.synchronize()
    data = d_in.cpu().numpy()
    s = mh.Session(n)
    try:
        for order in (1, 0):
            want_stream, want_table = o.compress_from_input(data, bool(order))
            assert mh.decode_subsequence_bits(order, (len(want_stream) - 1) * 8) == 8192
            stream, provider = s.compress(data, order)
            assert sha(stream) == sha(want_stream) and provider.write_coding_tree() == want_table
            assert np.array_equal(np.frombuffer(s.decompress_into(provider, stream, n), dtype=np.uint8), data)
    finally:
        s.close()
