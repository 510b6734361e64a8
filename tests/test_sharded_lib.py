"""The multi-GPU path inside the library (mh_comm_* / mh_sharded_*, csrc/mh_shard.cu).

Every rank is driven by its own host thread, as a multi-GPU host driver would. On a box with one GPU the ranks share the
device through the library's in-process transport; with >= 2 GPUs the same test also runs over NCCL, one rank per
device. The merged shards must equal the stream the (unsharded) oracle writes, byte for byte, and every rank must get
its own byte range back — through the exact layout and through the speculative bit-range decode with halo exchange and
seam handshake.
"""
import importlib
import threading

import numpy as np
import pytest

import oracle_py as o
from conftest import golden_input
from mhlib import load

mh = load()
sharding = importlib.import_module("markov-huffman-coding_b200.sharding")
pytestmark = pytest.mark.gpu


def run_ranks(comms, fn):
    """fn(rank, comm) on one thread per rank; re-raises the first failure."""
    errors, results = [], [None] * len(comms)

    def body(r):
        try:
            results[r] = fn(r, comms[r])
        except BaseException as e:  # noqa: BLE001 - surfaced below
            errors.append((r, e))

    threads = [threading.Thread(target=body, args=(r,)) for r in range(len(comms))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    if errors:
        raise AssertionError("rank %d: %r" % errors[0])
    return results


def sharded_round_trip(comms, data, cuts, order, speculative):
    import torch
    world = len(comms)
    want_stream, want_table = o.compress_from_input(data, bool(order))
    off = mh.shard_payload_offset()

    def rank_fn(r, comm):
        dev = torch.device("cuda", comm.device)
        with torch.cuda.device(dev):
            stream = torch.cuda.Stream(device=dev)
            mine = data[cuts[r]:cuts[r + 1]]
            d_in = torch.frombuffer(bytearray(mine) or bytearray(1), dtype=torch.uint8).to(dev)
            cap = mh.shard_local_bytes(len(mine) + len(mine) // 8 + 4096)
            d_local = torch.zeros(cap, dtype=torch.uint8, device=dev)
            d_out = torch.zeros(max(1, len(mine)) + 64, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(dev)
            layout, provider = comm.compress(d_in.data_ptr(), len(mine), order, d_local.data_ptr(), cap, prepare_decode=(r % 2 == 0),
                                             stream=stream.cuda_stream)
            assert provider.write_coding_tree() == want_table
            base, bits = int(layout.bit_base[r]), int(layout.n_bits[r])
            nbytes = ((base & 7) + bits + 7) // 8
            payload = d_local[off:off + nbytes].cpu().numpy().tobytes()
            n_out, out_off = comm.decompress(provider, d_local.data_ptr(), cap, layout, d_out.data_ptr(), len(mine), speculative=speculative,
                                             stream=stream.cuda_stream)
            stream.synchronize()
            assert (n_out, out_off) == (len(mine), cuts[r])
            assert d_out[:n_out].cpu().numpy().tobytes() == mine
            assert int(layout.prev0[r]) == (0x20 if cuts[r] == 0 else data[cuts[r] - 1])
            return payload, base, bits, int(layout.total_bits)

    parts = run_ranks(comms, rank_fn)
    assert all(p[3] == parts[0][3] for p in parts)
    merged = sharding.merge_payload_shards([(p[0], p[1], p[2]) for p in parts])
    assert bytes([sharding.stream_header(order, parts[0][3])]) + merged == want_stream
    assert world == comms[0].world


@pytest.fixture(scope="module")
def text():
    ipsum = o.histogram(golden_input("input_ipsum.txt"), True).astype(np.uint32)
    return o.synth_markov(ipsum, 77, 4096, 0, (2 << 20) + 4321)


@pytest.mark.parametrize("order", [1, 0], ids=["markov", "huffman"])
@pytest.mark.parametrize("world", [2, 3, 5])
def test_ranks_sharing_one_gpu_equal_the_unsharded_stream(text, world, order):
    """In-process transport, `world` ranks on device 0: cuts in the middle of codewords and bytes."""
    n = len(text)
    cuts = [n * i // world + (13 * i if 0 < i < world else 0) for i in range(world + 1)]
    comms = mh.Comm.create_local(world, devices=[0] * world, use_nccl=0)
    try:
        assert comms[1].transport == "in-process"
        for speculative in (True, False):
            sharded_round_trip(comms, text, cuts, order, speculative)
        st = comms[1].stats()
        assert st["calls"] == 4 and st["gather_us"] > 0 and st["halo_us"] > 0 and st["seam_us"] > 0
    finally:
        for c in comms:
            c.close()


def test_binary_data_and_an_empty_shard():
    """K = 256 data (no pair table: the 8-bit LUT path) and a rank without any bytes (exact layout)."""
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, 700_001, dtype=np.uint8).tobytes()
    comms = mh.Comm.create_local(3, devices=[0, 0, 0], use_nccl=0)
    try:
        sharded_round_trip(comms, data, [0, 300_000, 300_000, len(data)], 1, False)
        sharded_round_trip(comms, data, [0, 250_007, 500_000, len(data)], 1, True)
    finally:
        for c in comms:
            c.close()


def test_speculative_mode_refuses_shards_shorter_than_the_warm_up(text):
    comms = mh.Comm.create_local(2, devices=[0, 0], use_nccl=0)
    try:
        with pytest.raises(AssertionError, match="invalid argument"):
            sharded_round_trip(comms, text[:3000], [0, 2500, 3000], 1, True)
    finally:
        for c in comms:
            c.close()


@pytest.mark.parametrize("order", [1, 0], ids=["markov", "huffman"])
@pytest.mark.parametrize("world", [2, 3])
def test_host_buffer_calls_of_a_multi_gpu_driver(text, world, order):
    """mh_sharded_compress_host / _decompress_host: all ranks of one process work on the SAME host buffers (what the CLI
    does with several GPUs). The stream equals the oracle's; extraction cuts the payload at arbitrary bits."""
    data = np.frombuffer(text, dtype=np.uint8)
    want_stream, want_table = o.compress_from_input(text, bool(order))
    out = np.zeros(len(text) + len(text) // 8 + 4200, dtype=np.uint8)
    back = np.zeros(len(text) + 16, dtype=np.uint8)
    comms = mh.Comm.create_local(world, devices=[0] * world, use_nccl=0)
    try:
        res = run_ranks(comms, lambda r, c: c.compress_host(data, order, out))
        n = res[0][0]
        assert all(x[0] == n for x in res) and res[0][1] is not None and all(x[1] is None for x in res[1:])
        assert out[:n].tobytes() == want_stream
        provider = res[0][1]
        assert provider.write_coding_tree() == want_table
        stream = out[:n].copy()
        got = run_ranks(comms, lambda r, c: c.decompress_host(provider, stream, back))
        assert all(x == len(text) for x in got) and back[:len(text)].tobytes() == text
        small = np.frombuffer(want_stream[:2000], dtype=np.uint8)
        with pytest.raises(AssertionError, match="invalid argument"):
            run_ranks(comms, lambda r, c: c.decompress_host(provider, small, back))
    finally:
        for c in comms:
            c.close()


def test_nccl_ranks_one_per_gpu_equal_the_unsharded_stream(text):
    """>= 2 GPUs: the same through NCCL (ncclCommInitAll, one host thread per GPU)."""
    import torch
    if torch.cuda.device_count() < 2 or not mh.comm_available():
        pytest.skip("needs two GPUs and NCCL")
    world = min(torch.cuda.device_count(), 8)
    n = len(text)
    cuts = [n * i // world + (5 * i if 0 < i < world else 0) for i in range(world + 1)]
    comms = mh.Comm.create_local(world, use_nccl=2)
    try:
        assert comms[0].transport == "nccl"
        for order in (1, 0):
            for speculative in (True, False):
                sharded_round_trip(comms, text, cuts, order, speculative)
    finally:
        for c in comms:
            c.close()
