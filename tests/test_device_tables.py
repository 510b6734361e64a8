"""The encoder's tables built on the device (csrc/mh_tables.cu: the reference's heap, tie-breaking, int32 weights and
code assignment, src/huffman.cpp:97-164 + src/min_pq.tpp, one warp per context) against the host-built ones, which
tests/test_host_tables.py pins to the oracle and to the reference's own classes. Bit for bit, on count vectors chosen
for ties, long codewords and wrapping weights, and on the corpus; then the encoder run from them."""
import numpy as np
import pytest

import oracle_py as o
from conftest import golden_input
from mhlib import load
from test_oracle import _count_vectors

mh = load()
pytestmark = pytest.mark.gpu


def device_tables(counts_u64, order):
    import torch
    d_counts = torch.from_numpy(np.ascontiguousarray(counts_u64).view(np.int64)).cuda()
    book = mh.Codebook()
    book.build_device(d_counts.data_ptr(), order)
    return book.download()


def host_tables(counts_u64, order):
    provider = mh.CodingProvider.from_counts_array(np.ascontiguousarray(counts_u64), order)
    return mh.Codebook(provider).download(), provider


def check(counts_u64, order):
    d_enc, d_ctx, d_meta = device_tables(counts_u64, order)
    try:
        (h_enc, h_ctx, h_meta), provider = host_tables(counts_u64, order)
    except mh.MhError as e:                          # a codeword longer than 56 bits: both sides refuse the table
        assert e.status == mh.MH_ERR_CODE_TOO_LONG and np.int32(d_meta[1]) == mh.MH_ERR_CODE_TOO_LONG
        return None
    n = 65536 if order else 256
    assert np.array_equal(d_enc[:n], h_enc[:n])
    assert int(d_meta[1]) == 0
    assert int(d_meta[2]) == provider.max_code_bits()
    if h_meta[0]:                                    # the host made a context-row table: same rows, same entries
        assert int(d_meta[0]) == int(h_meta[0])
        assert np.array_equal(d_ctx, h_ctx)
    # the sums the encoder sizes itself by: sum of count x codeword length and sum of the counts, modulo 2^64
    lens = provider.code_lengths()
    c = np.ascontiguousarray(counts_u64)[:n].astype(np.uint64)
    with np.errstate(over="ignore"):
        assert int(d_meta[4]) | (int(d_meta[5]) << 32) == int(np.sum(c * lens, dtype=np.uint64))
        assert int(d_meta[6]) | (int(d_meta[7]) << 32) == int(np.sum(np.where(_ctx_live(c, order), c, 0), dtype=np.uint64))
    return provider


def _ctx_live(c, order):
    """Counts that belong to a context with a tree (a context whose counts are all 0 in the reference's 32-bit view has none)."""
    if not order:
        return np.ones(c.shape, dtype=bool)
    rows = c.reshape(256, 256)
    live_rows = ((rows & np.uint64(0xFFFFFFFF)) != 0).any(axis=1)
    return np.repeat(live_rows, 256)


@pytest.mark.parametrize("name,counts", _count_vectors(), ids=[n for n, _ in _count_vectors()])
def test_one_tree_from_count_vectors(name, counts):
    c = (np.asarray(counts).astype(np.int64) & 0xFFFFFFFF).astype(np.uint64)
    if not c.any():
        pytest.skip("empty vector")
    check(c, 0)
    # the same vector as context rows 3, 77 and 255 of a Markov table, other rows from a second vector
    m = np.zeros((256, 256), dtype=np.uint64)
    m[3] = c; m[77] = c[::-1]; m[255] = c
    check(m.reshape(-1), 1)


@pytest.mark.parametrize("name", ["input_a.txt", "input_ipsum.txt", "input_wiki_cpp.txt", "input_wiki_cpp.html"])
@pytest.mark.parametrize("order", [1, 0])
def test_corpus_histograms(name, order):
    data = golden_input(name)
    counts = o.histogram(data, bool(order)).astype(np.int64).astype(np.uint64)
    check(counts, order)


def test_all_256_contexts_live():
    """K = 256 data: every context has a tree (rank 255 must not be mistaken for 'no tree')."""
    rng = np.random.default_rng(11)
    c = rng.integers(1, 50, 65536).astype(np.uint64)
    check(c, 1)


def test_wrapped_count_is_reported_and_the_encoder_refuses():
    import torch
    c = np.zeros(65536, dtype=np.uint64)
    c[0x20 * 256 + 65] = 1 << 32                     # a live count whose int counter shows 0
    c[0x20 * 256 + 66] = 5
    _, _, meta = device_tables(c, 1)
    assert np.int32(meta[1]) == mh.MH_ERR_COUNT_WRAPPED


@pytest.mark.parametrize("order", [1, 0])
def test_encoder_runs_from_device_built_tables(order):
    """histogram -> tables on the device -> encoder, no host round trip: the stream equals the oracle's. A table that does
    not fit the optimistic launch (188 live contexts) makes the encoder answer d_result[3] != 0 instead of writing."""
    import torch
    ipsum = o.histogram(golden_input("input_ipsum.txt"), True).astype(np.uint32)
    for data, fits in ((o.synth_markov(ipsum, 3, 4096, 0, (1 << 20) + 77), True), (golden_input("input_wiki_cpp.html"), order == 0)):
        want = o.compress_from_input(data, bool(order))[0]
        n = len(data)
        d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        d_counts = torch.zeros(65536, dtype=torch.int64, device="cuda")
        d_res = torch.zeros(4, dtype=torch.int64, device="cuda")
        cap = n + n // 8 + 4096
        d_pay = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
        ws = mh.Workspace(n, cap)
        book = mh.Codebook()
        mh.gpu_histogram(d_in.data_ptr(), n, 0x20, order, d_counts.data_ptr(), ws)
        book.build_device(d_counts.data_ptr(), order)
        mh.gpu_encode(d_in.data_ptr(), n, 0x20, book, 0, d_pay.data_ptr(), cap, d_res.data_ptr(), ws)
        res = d_res.cpu().numpy()
        if fits:
            assert res[3] == 0 and res[2] == 0
            bits = int(res[0])
            got = bytes([0x30 | ((~order & 1) << 3) | ((8 - bits % 8) % 8)]) + d_pay[:(bits + 7) // 8].cpu().numpy().tobytes()
            assert got == want
        else:
            assert res[3] != 0
