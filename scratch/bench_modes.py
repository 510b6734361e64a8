#!/usr/bin/env python3
"""Kernel times of BASELINE.json's other single-GPU configurations (parity-test cases, not bench lines):

    python scratch/bench_modes.py text 1 1073741824       # config 1: 1 GiB Markov text, Markov mode (the bench workload)
    python scratch/bench_modes.py text 0 1073741824       # config 2: the same text with -h (one tree)
    python scratch/bench_modes.py fib  1 268435456        # config 3: 256 MiB Fibonacci-skewed stream (codewords > 8 bits)

Prints one JSON line: per-kernel ms (CUDA events inside the library), GB/s of uncompressed data per phase, and checks the
round trip. Inputs are resident in HBM; the tables are built on the host from the GPU histogram like in bench.py."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import numpy as np
    import torch
    kind, order, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    mh = importlib.import_module("markov-huffman-coding_b200")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    if kind == "text":
        data = np.frombuffer(open(os.path.join(ROOT, "tests/golden/inputs/input_ipsum.txt"), "rb").read(), dtype=np.uint8)
        prev = np.concatenate([np.array([0x20], dtype=np.uint8), data[:-1]])
        tc = np.zeros(65536, dtype=np.uint32)
        np.add.at(tc, prev.astype(np.int64) * 256 + data, 1)
        mh.synth_markov(tc, 20261018, 65536, 0, d_in.data_ptr(), n, stream)
    else:
        mh.synth_fibonacci(40, 48, 4321, 0, d_in.data_ptr(), n, stream)
    cap = n + n // 8 + 4096
    d_payload = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    bins = 65536 if order else 256
    d_counts = torch.zeros(65536, dtype=torch.int64, device=dev)
    d_res = torch.zeros(8, dtype=torch.int64, device=dev)
    ws = mh.Workspace(n, cap)
    book = dectab = None

    def step():
        nonlocal book, dectab
        mh.gpu_histogram(d_in.data_ptr(), n, 0x20, order, d_counts.data_ptr(), ws, stream)
        counts = d_counts[:bins].cpu().numpy().view(np.uint64)
        provider = mh.CodingProvider.from_counts_array(np.ascontiguousarray(counts), order)
        if book is None:
            book, dectab = mh.Codebook(provider), mh.DecodeTable(provider)
        book.update(provider, stream)
        mh.gpu_encode(d_in.data_ptr(), n, 0x20, book, 0, d_payload.data_ptr(), cap, d_res.data_ptr(), ws, stream)
        dectab.update(provider, stream)
        bits = int(d_res[0].item())
        mh.gpu_decode(d_payload.data_ptr(), 0, bits, 0x20, dectab, d_out.data_ptr(), n, d_res[4:].data_ptr(), ws, stream)
        r = d_res.cpu().tolist()
        assert r[4] == n and r[5] == 0 and r[6] == 0, r
        return bits, provider

    for _ in range(2):
        bits, provider = step()
    assert torch.equal(d_in, d_out), "round trip mismatch"
    mh.profile_enable(True)
    for _ in range(steps):
        step()
    prof = mh.profile_report()
    mh.profile_enable(False)
    kern = {k: v["ms"] / v["launches"] for k, v in prof.items()}
    enc = sum(v for k, v in kern.items() if k.startswith("encode"))
    hist = sum(v for k, v in kern.items() if k.startswith("hist"))
    dec = sum(v for k, v in kern.items() if k.startswith("dec_"))
    print(json.dumps({"kind": kind, "order": order, "bytes": n, "compressed_ratio": bits / 8 / n, "max_code_bits": provider.max_code_bits(),
                      "pair_table": provider.pair_lut() is not None, "kernels_ms": {k: round(v, 4) for k, v in kern.items()},
                      "histogram_gbs": n / hist / 1e6, "encode_gbs": n / enc / 1e6, "decode_gbs": n / dec / 1e6}))


if __name__ == "__main__":
    main()
