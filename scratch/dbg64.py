"""Debug: the SPT = 64 encoder launch against the SPT = 32 one (tunable enc_spt) through the device API."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
from mhlib import load
from conftest import golden_input
import oracle_py as o
mh = load()
dev = torch.device("cuda:0")

def enc(data, order, spt):
    mh.tunable_set("enc_spt", spt)
    n = len(data)
    d_in = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).to(dev)
    cap = n + n // 8 + 4096
    d_out = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
    d_counts = torch.zeros(65536, dtype=torch.int64, device=dev)
    d_res = torch.zeros(8, dtype=torch.int64, device=dev)
    ws = mh.Workspace(max(n, 1 << 20), cap)
    st = torch.cuda.current_stream().cuda_stream
    mh.gpu_histogram(d_in.data_ptr(), n, 0x20, order, d_counts.data_ptr(), ws, st)
    book = mh.Codebook(); book.build_device(d_counts.data_ptr(), order, st)
    mh.gpu_encode(d_in.data_ptr(), n, 0x20, book, 0, d_out.data_ptr(), cap, d_res.data_ptr(), ws, st)
    torch.cuda.synchronize()
    res = d_res.cpu().numpy().tolist()
    meta = book.download()[2]
    bits = res[0]
    return res[:4], d_out[: (bits + 7) // 8].cpu().numpy(), meta

def cmp(name, data, order):
    r32, s32, meta = enc(data, order, 32)
    r64, s64, _ = enc(data, order, -1)
    m = min(len(s32), len(s64))
    diff = np.nonzero(s32[:m] != s64[:m])[0]
    print(name, "order", order, "n", len(data), "res32", r32, "res64", r64, "meta", meta.tolist(),
          "first diff byte", (int(diff[0]), len(diff)) if len(diff) else None, flush=True)

cmp("fib24", golden_input("edge_fib24_ties.bin"), 0)
cmp("fib24", golden_input("edge_fib24_ties.bin"), 1)
cmp("fib24[:30720]", golden_input("edge_fib24_ties.bin")[:30720], 0)
cmp("fib24[:20000]", golden_input("edge_fib24_ties.bin")[:20000], 0)
cmp("ipsum", golden_input("input_ipsum.txt"), 1)
cmp("ipsum", golden_input("input_ipsum.txt"), 0)
big = golden_input("input_ipsum.txt") * 40
cmp("ipsum x40", big, 1)
cmp("ipsum x40", big, 0)
big = golden_input("input_ipsum.txt") * 400
cmp("ipsum x400", big, 1)
rng = np.random.default_rng(5)
skew = bytes(rng.choice(256, size=3_000_000, p=np.r_[np.full(8, 0.1), np.full(248, 0.2 / 248)]).astype(np.uint8))
cmp("skewed binary (mean > 5.5 bits in order 0?)", skew, 0)
cmp("skewed binary", skew, 1)
cmp("fib40", golden_input("edge_fib40_256k.bin"), 0)
cmp("fib40", golden_input("edge_fib40_256k.bin"), 1)
