"""The multi-GPU path without 8 GPUs.

* CPU, world_size 2 over gloo: every rank holds a byte range, the local histograms are all-gathered, every rank
  builds the tables with the product's host code and derives its own global bit offset; the shard payloads
  (the oracle stands in for the encode kernel here) are merged and must equal the unsharded stream byte for byte.
* GPU (`-m gpu`): the same with k logical shards on one device through mh_gpu_histogram / mh_gpu_encode /
  mh_gpu_decode with prev0 and bit_base.
"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

import oracle_py as o
from conftest import golden_input
from mhlib import load

mh = load()
sharding = importlib.import_module("markov-huffman-coding_b200.sharding")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, order, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = golden_input("input_wiki_cpp.html")
        cut = [0, 150001, len(data)] if world == 2 else [len(data) * i // world for i in range(world + 1)]
        mine = data[cut[rank]:cut[rank + 1]]
        prev0 = 0x20 if rank == 0 else data[cut[rank] - 1]
        bins = 65536 if order else 256
        local = o.histogram(mine, bool(order), prev0).astype(np.int64)          # stand-in for mh_gpu_histogram
        gathered = [torch.zeros(bins, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(local))
        all_counts = np.stack([g.numpy() for g in gathered]).astype(np.uint64)
        provider = mh.CodingProvider.from_counts_array(sharding.global_counts(all_counts), order)   # product host code
        base, bits = sharding.shard_bit_bases(all_counts, provider.code_lengths())
        table = o.Table.from_bytes(provider.write_coding_tree())
        payload, nbits = table.encode_shard(mine, prev0, int(base[rank]))                            # stand-in for mh_gpu_encode
        assert nbits == int(bits[rank])
        parts = [None] * world
        dist.all_gather_object(parts, (payload, int(base[rank]), nbits))
        if rank == 0:
            merged = sharding.merge_payload_shards(parts)
            total = int(base[-1] + bits[-1])
            stream = bytes([sharding.stream_header(order, total)]) + merged
            want_stream, want_table = o.compress_from_input(data, bool(order))
            q.put((stream == want_stream, provider.write_coding_tree() == want_table))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("order", [0, 1])
def test_two_rank_gloo_sharded_compress_equals_unsharded(order):
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, order, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == (True, True)


def test_merge_handles_empty_and_byte_aligned_shards():
    t = o.Table.from_counts(o.histogram(b"abracadabra" * 9, True), True)
    data = b"abracadabra" * 9
    cuts = [0, 0, 11, 11, 40, len(data)]
    parts, base = [], 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        prev0 = 0x20 if a == 0 else data[a - 1]
        payload, nbits = t.encode_shard(data[a:b], prev0, base)
        parts.append((payload, base, nbits))
        base += nbits
    assert bytes([sharding.stream_header(1, base)]) + sharding.merge_payload_shards(parts) == t.compress(data)


@pytest.mark.gpu
@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("k", [2, 5])
def test_logical_shards_on_one_gpu(order, k):
    import torch
    ipsum = o.histogram(golden_input("input_ipsum.txt"), True).astype(np.uint32)
    data = o.synth_markov(ipsum, 31, 4096, 0, (3 << 20) + 1234)
    n = len(data)
    cuts = [n * i // k + (7 * i if 0 < i < k else 0) for i in range(k + 1)]
    dev = torch.device("cuda", 0)
    d_all = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    ws = mh.Workspace(n, n + n // 8 + 4096)
    bins = 65536 if order else 256
    all_counts = np.zeros((k, bins), dtype=np.uint64)
    d_counts = torch.zeros(bins, dtype=torch.int64, device=dev)
    for g in range(k):
        prev0 = 0x20 if g == 0 else data[cuts[g] - 1]
        shard = d_all[cuts[g]:cuts[g + 1]]                       # unaligned device pointers on purpose
        mh.gpu_histogram(shard.data_ptr(), shard.numel(), prev0, order, d_counts.data_ptr(), ws)
        all_counts[g] = d_counts.cpu().numpy().view(np.uint64)
    provider = mh.CodingProvider.from_counts_array(sharding.global_counts(all_counts), order)
    want_stream, want_table = o.compress_from_input(data, bool(order))
    assert provider.write_coding_tree() == want_table
    base, bits = sharding.shard_bit_bases(all_counts, provider.code_lengths())
    book, dectab = mh.Codebook(provider), mh.DecodeTable(provider)
    d_res = torch.zeros(4, dtype=torch.int64, device=dev)
    parts, outs = [], []
    for g in range(k):
        prev0 = 0x20 if g == 0 else data[cuts[g] - 1]
        shard = d_all[cuts[g]:cuts[g + 1]]
        cap = shard.numel() + shard.numel() // 8 + 4096
        d_pay = torch.zeros(cap, dtype=torch.uint8, device=dev)
        mh.gpu_encode(shard.data_ptr(), shard.numel(), prev0, book, int(base[g]), d_pay.data_ptr(), cap, d_res.data_ptr(), ws)
        res = d_res.cpu().numpy()
        assert int(res[0]) == int(bits[g]) and int(res[2]) == 0
        parts.append((d_pay.cpu().numpy().tobytes(), int(base[g]), int(bits[g])))
        d_out = torch.zeros(shard.numel(), dtype=torch.uint8, device=dev)
        mh.gpu_decode(d_pay.data_ptr(), int(base[g]), int(bits[g]), prev0, dectab, d_out.data_ptr(), shard.numel(), d_res.data_ptr(), ws)
        res = d_res.cpu().numpy()
        assert int(res[0]) == shard.numel() and int(res[1]) == 0 and int(res[2]) == 0
        outs.append(d_out.cpu().numpy().tobytes())
    total = int(base[-1] + bits[-1])
    assert bytes([sharding.stream_header(order, total)]) + sharding.merge_payload_shards(parts) == want_stream
    assert b"".join(outs) == data


@pytest.mark.gpu
@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("k", [2, 3, 8])
def test_bit_range_sharded_decode_with_seam_handshake(order, k):
    """Decode ONE stream as k bit-range shards: shard 0 knows its state, the others start a warm-up before their
    range, guess, and must arrive at the state their predecessor ended in (mh_gpu_decode_shard)."""
    import torch
    ipsum = o.histogram(golden_input("input_ipsum.txt"), True).astype(np.uint32)
    data = o.synth_markov(ipsum, 77, 4096, 0, (2 << 20) + 4321)
    stream, table = o.compress_from_input(data, bool(order))
    total_bits = (len(stream) - 1) * 8 - (stream[0] & 7)
    dev = torch.device("cuda", 0)
    pad = 64                                                        # readable bytes past the payload (zero = pop_rest padding)
    d_pay = torch.zeros(len(stream) - 1 + pad, dtype=torch.uint8, device=dev)
    d_pay[: len(stream) - 1] = torch.frombuffer(bytearray(stream[1:]), dtype=torch.uint8).to(dev)
    provider = mh.CodingProvider.from_table_file(table)
    dectab = mh.DecodeTable(provider)
    ws = mh.Workspace(len(data), len(stream) + pad)
    warm = mh.DECODE_WARM_UNIT
    bounds = [total_bits * g // k + (13 * g if 0 < g < k else 0) for g in range(k + 1)]   # arbitrary, mid-codeword boundaries
    d_res = torch.zeros(4, dtype=torch.int64, device=dev)
    outs, seams = [], []
    for g in range(k):
        origin = 0 if g == 0 else bounds[g] - warm
        assert origin >= 0
        ptr_off = (origin // 32) * 4
        d_out = torch.zeros(len(data), dtype=torch.uint8, device=dev)
        mh.gpu_decode_shard(d_pay.data_ptr() + ptr_off, origin % 32, bounds[g + 1] - origin, d_pay.numel() - ptr_off,
                            g == 0, 0x20, 0 if g == 0 else warm, g == k - 1, dectab, d_out.data_ptr(), d_out.numel(), d_res.data_ptr(), ws)
        res = d_res.cpu().numpy().view(np.uint64)
        assert int(res[1]) == 0 and int(res[2]) == 0, res
        outs.append(d_out[: int(res[0])].cpu().numpy().tobytes())
        seams.append((int(res[3]) >> 32, int(res[3]) & 0xFFFFFFFF))
    for g in range(1, k):
        assert seams[g][0] == seams[g - 1][1], "shard %d did not converge onto its predecessor's end state" % g
    assert b"".join(outs) == data


@pytest.mark.gpu
def test_shard_restart_from_exact_seam_state():
    """The fallback of the handshake: a shard decoded again from its predecessor's recorded end state."""
    import torch
    data = o.synth_fibonacci(40, 48, 5, 0, 1 << 20)
    stream, table = o.compress_from_input(data, True)
    total_bits = (len(stream) - 1) * 8 - (stream[0] & 7)
    dev = torch.device("cuda", 0)
    d_pay = torch.zeros(len(stream) - 1 + 64, dtype=torch.uint8, device=dev)
    d_pay[: len(stream) - 1] = torch.frombuffer(bytearray(stream[1:]), dtype=torch.uint8).to(dev)
    dectab = mh.DecodeTable(mh.CodingProvider.from_table_file(table))
    ws = mh.Workspace(len(data), len(stream) + 64)
    d_res = torch.zeros(4, dtype=torch.int64, device=dev)
    cut = total_bits // 2 + 5
    d_a = torch.zeros(len(data), dtype=torch.uint8, device=dev)
    mh.gpu_decode_shard(d_pay.data_ptr(), 0, cut, d_pay.numel(), True, 0x20, 0, False, dectab, d_a.data_ptr(), d_a.numel(), d_res.data_ptr(), ws)
    res = d_res.cpu().numpy().view(np.uint64)
    n_a, end_state = int(res[0]), int(res[3]) & 0xFFFFFFFF
    start = cut + (end_state >> 8)                                  # first codeword boundary at or after the cut
    ptr_off = (start // 32) * 4
    d_b = torch.zeros(len(data), dtype=torch.uint8, device=dev)
    mh.gpu_decode_shard(d_pay.data_ptr() + ptr_off, start % 32, total_bits - start, d_pay.numel() - ptr_off, True, end_state & 255, 0, True,
                        dectab, d_b.data_ptr(), d_b.numel(), d_res.data_ptr(), ws)
    res = d_res.cpu().numpy().view(np.uint64)
    assert int(res[1]) == 0 and int(res[2]) == 0
    assert d_a[:n_a].cpu().numpy().tobytes() + d_b[: int(res[0])].cpu().numpy().tobytes() == data


def test_seam_pair_fixup_and_shard_sizes_from_histograms():
    """Shards count their first byte as following ' '; fix_seam_pairs moves that one count to the real pair, after
    which the summed histograms equal the unsharded one and sum(count x length) gives every shard's payload size."""
    d = golden_input("input_wiki_cpp.html")
    cuts = [0, 100000, 200001, len(d)]
    ac = np.stack([o.histogram(d[a:b], True, 0x20).astype(np.int64).astype(np.uint64) for a, b in zip(cuts[:-1], cuts[1:])])
    raw_total = ac.sum(axis=0, dtype=np.uint64)        # what a reduction on the GPU delivers before the seams are known
    sharding.fix_seam_pairs(ac, [d[a] for a in cuts[:-1]], [d[b - 1] for b in cuts[1:]])
    assert np.array_equal(sharding.global_counts(ac), o.histogram(d, True).astype(np.int64).astype(np.uint64))
    assert np.array_equal(sharding.fix_seam_total(raw_total, [d[a] for a in cuts[:-1]], [d[b - 1] for b in cuts[1:]]), sharding.global_counts(ac))
    assert np.array_equal(sharding.shard_bits(ac, mh.CodingProvider.from_counts_array(sharding.global_counts(ac), 1).code_lengths(np.uint8)),
                          sharding.shard_bits(ac, mh.CodingProvider.from_counts_array(sharding.global_counts(ac), 1).code_lengths()))
    provider = mh.CodingProvider.from_counts_array(sharding.global_counts(ac), 1)
    base, bits = sharding.shard_bit_bases(ac, provider.code_lengths())
    table = o.Table.from_counts(o.histogram(d, True), True)
    for g, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        assert int(bits[g]) == table.encode_shard(d[a:b], 0x20 if a == 0 else d[a - 1], int(base[g]))[1]
