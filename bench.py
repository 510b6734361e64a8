#!/usr/bin/env python3
"""bench.py — encode/decode throughput of the B200 Markov-Huffman codec (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one compress (order-1 histogram -> host tree build -> table upload -> encode) plus one extract (decode) of
the rank's resident input, i.e. 2 x N uncompressed bytes go through the hot path per step and rank.
  value  whole-job GB/s of uncompressed data over the K timed steps, inputs already in HBM (CUDA events, max over ranks)
  e2e    the same step through the host-buffer C-ABI session (mh_session_compress / mh_session_decompress) with pinned
         HOST buffers: H2D of the input, D2H of the compressed stream, H2D of the stream, D2H of the decoded bytes
  roofline  the dominant kernel of the step, its algorithmic bytes / its device time (CUDA events on the launching
         stream, recorded inside the library around every launch) against MEASURED_PEAKS.json's copy bandwidth
  cpu_baseline  the reference's own single-threaded CPU build (oracle/_ref) timed on this box on a bounded sample

Workload at N = 1: BASELINE.json configs[1] — 1 GiB of synthetic order-1 Markov text with input_ipsum.txt statistics,
Markov mode. At N > 1 each rank holds its own 1 GiB byte range of one logical N GiB stream (weak scaling): shard g
seeds its histogram and encoder with the last byte of shard g-1, the per-GPU histograms are all-gathered over NCCL
and summed so every rank builds identical tables, and the per-GPU bit totals (sum of local counts x code lengths) are
exclusive-scanned so every shard is encoded at its global bit offset. Decode treats the shards as ONE stream cut by
bit ranges: neighbours exchange a ~1 KiB halo, rank 0 starts exactly, every other rank starts a warm-up before its range
from a guessed state, and the ranks all-gather their seam states until each agrees with its predecessor's end state
(markov-huffman-coding_b200/sharding.py, mh_gpu_decode_shard).

`--impl reference` runs the unmodified reference (built from /root/reference/src into oracle/_ref by oracle/Makefile)
through its own CLI on this box's host cores — it has no threads, so one core — on a bounded sample per step.
"""
import argparse
import ctypes
import importlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GIB = 1 << 30
SEED = 20261018
SEG_BYTES = 65536
WORKLOAD = "1 GiB synthetic order-1 Markov text (input_ipsum.txt statistics, seed %d, 64 KiB segments), Markov mode" % SEED
REF_SAMPLE_BYTES = 64 << 20


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, n_bytes):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture
    (profiles/*_traffic.json, taken at 1 GiB), or None when the workload size differs / no capture exists."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            cap = json.load(open(path))
            if cap.get("input_bytes") != n_bytes:
                continue
            for name, rec in cap["kernels"].items():
                if kernel.split("<")[0] in name:
                    return rec["traffic_bytes"]
        except Exception:
            continue
    return None


def ipsum_transition_counts():
    import numpy as np
    data = np.frombuffer(open(os.path.join(ROOT, "tests/golden/inputs/input_ipsum.txt"), "rb").read(), dtype=np.uint8)
    prev = np.concatenate([np.array([0x20], dtype=np.uint8), data[:-1]])
    tc = np.zeros(65536, dtype=np.uint32)
    np.add.at(tc, prev.astype(np.int64) * 256 + data, 1)
    return tc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        if shutil.which("nvidia-smi") is None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation, through its CLI, on a bounded sample
# ---------------------------------------------------------------------------------------------------------
def reference_sample(n_bytes):
    import oracle_py as o
    return o.synth_markov(ipsum_transition_counts(), SEED, SEG_BYTES, 0, n_bytes)


def time_reference_once(sample_path, workdir):
    """One `-d` compress + one `-x` extract of the sample with the reference binaries. Returns (t_enc, t_dec, ok)."""
    import oracle_py as o
    comp, tab, dec = (os.path.join(workdir, x) for x in ("s.cm", "s.e", "s.dm"))
    t0 = time.perf_counter()
    subprocess.run([o.REF_STOCK, sample_path, "-o", comp, "-", "-d", tab], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t1 = time.perf_counter()
    subprocess.run([o.REF_PATCHED, comp, "-o", dec, "-x", "-e", tab], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t2 = time.perf_counter()
    ok = os.path.getsize(dec) == os.path.getsize(sample_path)
    return t1 - t0, t2 - t1, ok


def time_port_once(sample):
    import oracle_py as o
    t0 = time.perf_counter()
    stream, table = o.compress_from_input(sample, True)
    t1 = time.perf_counter()
    out = o.Table.from_bytes(table).decompress(stream, cap=len(sample) + 16)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, out == sample


def cpu_reference(steps, warmup, sample_bytes):
    """Returns (GB/s of uncompressed data over `steps` steps, details)."""
    import oracle_py as o
    o.build() if not os.path.exists(os.path.join(o.ORACLE_DIR, "libmh_oracle.so")) else None
    sample = reference_sample(sample_bytes)
    have_ref = os.path.exists(o.REF_STOCK) and os.path.exists(o.REF_PATCHED)
    base = "/dev/shm" if os.access("/dev/shm", os.W_OK) else None
    workdir = tempfile.mkdtemp(prefix="mhbench_", dir=base)
    try:
        path = os.path.join(workdir, "sample.bin")
        with open(path, "wb") as fh:
            fh.write(sample)
        t_enc = t_dec = 0.0
        for i in range(warmup + steps):
            a, b, ok = time_reference_once(path, workdir) if have_ref else time_port_once(sample)
            assert ok, "reference round trip failed"
            if i >= warmup:
                t_enc += a; t_dec += b
    finally:
        shutil.rmtree(workdir, ignore_errors=True)
    gbs = 2.0 * sample_bytes * steps / (t_enc + t_dec) / 1e9
    cpu_model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                cpu_model = line.split(":", 1)[1].strip(); break
    except Exception:
        pass
    detail = {
        "value": gbs, "unit": "GB/s", "cores": 1,
        "kind": "reference" if have_ref else "port",
        "sample": "%d MiB prefix of the workload, compress (-d) + extract (-x) through the %s, files in tmpfs; host %s, %d cores, 1 used (the reference has no threads)"
                  % (sample_bytes >> 20, "reference CLI built from its own sources (oracle/_ref)" if have_ref else "oracle port (oracle/mh_oracle.c)", cpu_model, os.cpu_count() or 0),
        "encode_gbs": sample_bytes * steps / t_enc / 1e9, "decode_gbs": sample_bytes * steps / t_dec / 1e9,
    }
    return gbs, (t_enc + t_dec) / steps * 1e3, detail


def run_reference_arm(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    gbs, ms, detail = cpu_reference(steps, warmup, REF_SAMPLE_BYTES)
    line = {
        "impl": "reference", "metric": "encode/decode GB/s (uncompressed)", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": {"workload": WORKLOAD, "step": "compress + extract of a bounded sample on the host CPU"},
        "cpu_baseline": detail,
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bytes", type=int, default=int(os.environ.get("MH_BENCH_BYTES", GIB)), help="input bytes per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    args.warmup = max(3, args.warmup)

    import numpy as np
    import torch
    import torch.distributed as dist
    mh = importlib.import_module("markov-huffman-coding_b200")   # raises if libmh_gpu.so is missing: no fallback
    sharding = importlib.import_module("markov-huffman-coding_b200.sharding")

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.bytes
    stream = torch.cuda.current_stream().cuda_stream

    # ---- resident synthetic input: rank r holds bytes [r*n, (r+1)*n) of the logical stream ----
    tc = ipsum_transition_counts()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    mh.synth_markov(tc, SEED, SEG_BYTES, rank * (n // SEG_BYTES), d_in.data_ptr(), n, stream)
    payload_cap = n + n // 8 + 4096
    # the payload sits LOCAL_PAD bytes into a local buffer so that a neighbour's warm-up halo can be spliced in front
    d_local = torch.zeros(sharding.LOCAL_PAD + payload_cap + 256, dtype=torch.uint8, device=dev)
    d_payload = d_local[sharding.LOCAL_PAD:]
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    MSG = 65536 + 2   # the histogram, then this shard's first and last byte: one message per rank and step
    d_msg = torch.zeros(MSG, dtype=torch.int64, device=dev)
    d_counts = d_msg[:65536]
    d_res_enc = torch.zeros(4, dtype=torch.int64, device=dev)
    d_res_dec = torch.zeros(4, dtype=torch.int64, device=dev)
    h_counts = torch.empty(MSG * world, dtype=torch.int64, pin_memory=True)
    h_total = torch.empty(65536, dtype=torch.int64, pin_memory=True)     # the summed histogram (world > 1: reduced on the GPU)
    h_res = torch.empty(8, dtype=torch.int64, pin_memory=True)
    ws = mh.Workspace(n, payload_cap)
    book = dectab = None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2 (126 MB); inputs are also >> L2
    if world > 1:
        edge_idx = torch.tensor([0, n - 1], device=dev)
        gathered = torch.zeros(MSG * world, dtype=torch.int64, device=dev)
        shard_decoder = sharding.ShardedDecoder(mh, dist, torch, rank, world, 1, dev)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    state = {}
    host_us = {"trees": 0.0, "codebook": 0.0, "dectable": 0.0}   # host work on the step's critical path (timed steps)

    def step(timed):
        """compress then extract of the resident shard; returns (ms_encode_phase, ms_decode_phase)."""
        nonlocal book, dectab
        ev[0].record()
        prev0, bit_base = 0x20, 0
        # Every shard counts its first byte as following ' '; the one seam pair per shard is corrected on the host from
        # the first / last bytes that travel with the histograms, so the whole exchange is a single all-gather.
        mh.gpu_histogram(d_in.data_ptr(), n, 0x20, 1, d_counts.data_ptr(), ws, stream)
        if world > 1:
            d_msg[65536:] = d_in[edge_idx]
            dist.all_gather_into_tensor(gathered, d_msg)
            h_counts.copy_(gathered, non_blocking=True)
            h_total.copy_(gathered.view(world, MSG)[:, :65536].sum(0), non_blocking=True)   # the host only fixes the seam pairs
        else:
            h_counts[:MSG].copy_(d_msg, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        msgs = h_counts.numpy().view(np.uint64).reshape(world, MSG)
        all_counts = msgs[:, :65536]
        if world > 1:
            sharding.fix_seam_pairs(all_counts, msgs[:, 65536], msgs[:, 65537])
            if rank > 0:
                prev0 = int(msgs[rank - 1, 65537])
        th0 = time.perf_counter()
        if world > 1:
            total_counts = h_total.numpy().view(np.uint64)
            sharding.fix_seam_total(total_counts, msgs[:, 65536], msgs[:, 65537])
        else:
            total_counts = sharding.global_counts(all_counts)
        provider = mh.CodingProvider.from_counts_array(total_counts, 1)     # identical on every rank
        th1 = time.perf_counter()
        if book is None:
            book, dectab = mh.Codebook(provider), mh.DecodeTable(provider)
        book.update(provider, stream)
        th2 = time.perf_counter()
        expect_bits = None
        if world > 1:   # every rank derives every shard's global bit offset from the gathered histograms
            base, shard_bits = sharding.shard_bit_bases(all_counts, provider.code_lengths(np.uint8), total_counts)
            bit_base, expect_bits = int(base[rank]), int(shard_bits[rank])
        mh.gpu_encode(d_in.data_ptr(), n, prev0, book, bit_base, d_payload.data_ptr(), payload_cap, d_res_enc.data_ptr(), ws, stream)
        h_res[:4].copy_(d_res_enc, non_blocking=True)
        ev[1].record()
        th3 = time.perf_counter()
        dectab.update(provider, stream)     # the decoder's tables are flattened on the host while the encoder runs
        th4 = time.perf_counter()
        torch.cuda.current_stream().synchronize()
        bits = int(h_res[0])
        assert int(h_res[2]) == 0, "encode: capacity"
        assert expect_bits is None or bits == expect_bits, "shard payload size differs from sum(count x length)"
        if timed:
            host_us["trees"] += (th1 - th0) * 1e6; host_us["codebook"] += (th2 - th1) * 1e6; host_us["dectable"] += (th4 - th3) * 1e6
        if world > 1:
            # one stream, decoded by bit ranges: halo exchange, speculative start + warm-up, seam handshake over NCCL
            got = shard_decoder.decode(d_local, base, shard_bits, dectab, d_out, n, d_res_dec, ws, stream)
            assert got == n, "sharded decode returned %d symbols" % got
            state["seam_rounds"] = shard_decoder.rounds
        else:
            mh.gpu_decode(d_payload.data_ptr(), bit_base, bits, prev0, dectab, d_out.data_ptr(), n, d_res_dec.data_ptr(), ws, stream)
        h_res[4:].copy_(d_res_dec, non_blocking=True)
        ev[2].record()
        torch.cuda.current_stream().synchronize()
        assert int(h_res[4]) == n and int(h_res[5]) == 0 and int(h_res[6]) == 0, "decode: %s" % h_res[4:].tolist()
        state["bits"], state["provider"] = bits, provider
        return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    assert torch.equal(d_out, d_in), "round trip mismatch"       # size-independent parity property at the full size

    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    mh.profile_enable(True)
    launches0 = mh.kernel_launches()
    t_enc = t_dec = 0.0
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record()
    for _ in range(args.steps):
        a, b = step(True)
        t_enc += a; t_dec += b
    t_end.record()
    barrier()
    total_ms = t_begin.elapsed_time(t_end)
    launches = mh.kernel_launches() - launches0
    prof = mh.profile_report()
    mh.profile_enable(False)
    clock_info = clocks.stop() if rank == 0 else None
    assert torch.equal(d_out, d_in), "round trip mismatch after the timed steps"

    times = torch.tensor([total_ms, t_enc, t_dec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, t_enc, t_dec = times.tolist()
    value = 2.0 * n * world * args.steps / (total_ms * 1e-3) / 1e9

    # ---- end to end through the host-buffer session API (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        del flush
        session = mh.Session(n, device=local_rank)
        h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(d_in); torch.cuda.synchronize()
        h_stream = torch.empty(payload_cap + 1, dtype=torch.uint8, pin_memory=True)
        h_back = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        np_in, np_stream, np_back = h_in.numpy(), h_stream.numpy(), h_back.numpy()
        lib = mh._lib
        out_len, dec_len = ctypes.c_uint64(0), ctypes.c_uint64(0)
        e2e_steps = max(2, min(args.steps, 5))
        t_e2e = 0.0
        stream_bytes = 0
        for i in range(1 + e2e_steps):
            barrier()
            t0 = time.perf_counter()
            table = ctypes.c_void_p()
            rc = lib.mh_session_compress(session._h, np_in.ctypes.data, n, 1, np_stream.ctypes.data, np_stream.size, ctypes.byref(out_len), ctypes.byref(table))
            assert rc == 0, rc
            rc = lib.mh_session_decompress(session._h, table, np_stream.ctypes.data, out_len.value, np_back.ctypes.data, n, ctypes.byref(dec_len))
            assert rc == 0 and dec_len.value == n, (rc, dec_len.value)
            dt = time.perf_counter() - t0
            lib.mh_table_destroy(table)
            if i >= 1:
                t_e2e += dt
            stream_bytes = out_len.value
        assert bytes(np_back[:4096]) == bytes(np_in[:4096]) and np.array_equal(np_back[-4096:], np_in[-4096:])
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": 2.0 * n * world * e2e_steps / t.item() / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(n + stream_bytes + 512 * 1024 + 640 * 1024), "d2h_bytes_per_step": int(stream_bytes + n + 65536 * 8),
               "steps": e2e_steps, "api": "mh_session_compress + mh_session_decompress, pinned host buffers, per rank"}
        session.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_src = peaks()
    c_bytes = (state["bits"] + 7) // 8
    algo = {   # algorithmic bytes per launch; D1 only reads the payload, D4 reads it and writes N
        "hist_kernel<1>": n, "hist_lane_kernel": n, "hist0_lane_kernel": n, "encode_kernel": n + c_bytes, "dec_sync_kernel": c_bytes, "dec_write_kernel": c_bytes + n,
    }
    kern = {k: v["ms"] / max(1, v["launches"]) for k, v in prof.items()}
    dominant = max(kern, key=kern.get)
    dom_ms = kern[dominant]
    achieved = algo.get(dominant, 0) / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    phase = {
        "histogram": {"algorithmic_bytes": n, "ms": sum(v for k, v in kern.items() if k.startswith("hist"))},
        "encode": {"algorithmic_bytes": n + c_bytes, "ms": kern.get("encode_kernel", 0)},
        "decode": {"algorithmic_bytes": c_bytes + n, "ms": sum(v for k, v in kern.items() if k.startswith("dec_"))},
    }
    for p in phase.values():
        p["gbs"] = p["algorithmic_bytes"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] else None
        p["frac"] = p["gbs"] / peak if p["gbs"] else None
    line = {
        "metric": "encode/decode GB/s (uncompressed)", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD if n == GIB else WORKLOAD.replace("1 GiB", "%d MiB" % (n >> 20)), "bytes_per_gpu": n,
                   "step": "compress (histogram + host trees + encode) then extract (decode); 2 x bytes_per_gpu uncompressed bytes per step and GPU",
                   "l2": "inputs (>= 1 GiB) exceed the 126 MB L2; no flush needed", "compressed_ratio": c_bytes / n,
                   "sharding": ("encode: byte-range shards, NCCL all-gather of histograms, exclusive scan of per-GPU bit totals; decode: bit-range shards, "
                                "NCCL halo exchange, speculative start with warm-up, seam handshake (%d extra rounds)" % state.get("seam_rounds", 0)) if world > 1 else "single GPU"},
        "encode_gbs": n * world * args.steps / (t_enc * 1e-3) / 1e9, "decode_gbs": n * world * args.steps / (t_dec * 1e-3) / 1e9,
        "gpu_launches": int(launches), "kernels_ms_per_launch": kern, "phases": phase, "host_us_per_step": {k: round(v / args.steps, 1) for k, v in host_us.items()},
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(dominant, n),
                     "peak_source": peak_src, "algorithmic_bytes": algo.get(dominant, 0), "ms_per_launch": dom_ms},
        "clocks": clock_info,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        _, _, detail = cpu_reference(1, 0, REF_SAMPLE_BYTES)
        line["cpu_baseline"] = detail
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
