#!/usr/bin/env python3
"""Regenerate tests/golden/ from the reference itself (run in the build container, where /root/reference exists).

Inputs  : the reference's own test corpus /root/reference/test/input/* (copied to tests/golden/inputs/ so the
          GPU box, which has no /root/reference, can read them) plus seeded synthetic edge cases.
Outputs : tests/golden/golden.json — for every (input, mode): size + SHA-256 of the STOCK reference build's
          `-d` outputs (compressed stream, table file), the full bytes of every table file and of the streams
          of the tiny cases, and whether the patched build's `-x` round trip restored the input.
The reference is built by oracle/Makefile (`make -C oracle ref`) from the sources where they lie.
"""
import base64, ctypes, hashlib, json, os, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_INPUTS = "/root/reference/test/input"
STOCK = os.path.join(ROOT, "oracle/_ref/markovhuffman_stock")
PATCHED = os.path.join(ROOT, "oracle/_ref/markovhuffman_patched")
INLINE_LIMIT = 8192  # streams up to this size are stored in full


def synth_cases():
    """Seeded edge inputs the reference's corpus lacks (its own test compresses the compiler's output)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as o
    import numpy as np
    rng = np.random.default_rng(20261018)
    cases = {
        "edge_empty.bin": b"",
        "edge_single_Z.bin": b"Z",
        "edge_run_q.bin": b"q" * 1000,                       # single-symbol contexts only (F4)
        "edge_two_syms.bin": bytes(rng.integers(0, 2, 4096, dtype=np.uint8) + 65),
        "edge_random_64k.bin": bytes(rng.integers(0, 256, 65536, dtype=np.uint8)),   # K = 256 binary data
        "edge_fib40_256k.bin": o.synth_fibonacci(40, 48, 1234, 0, 262144),           # codewords > 8 bits
        "edge_fib24_ties.bin": o.synth_fibonacci(24, 97, 7, 0, 50000),
    }
    return cases


def run(exe, args):
    p = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if p.returncode != 0:
        raise RuntimeError("%s %s failed: %s" % (exe, args, p.stderr.decode()))


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "liboracle"])
    inputs_dir = os.path.join(HERE, "inputs")
    os.makedirs(inputs_dir, exist_ok=True)
    for f in sorted(os.listdir(REF_INPUTS)):
        shutil.copyfile(os.path.join(REF_INPUTS, f), os.path.join(inputs_dir, f))
    for name, data in synth_cases().items():
        with open(os.path.join(inputs_dir, name), "wb") as fh:
            fh.write(data)

    golden = {"generator": "tests/golden/make_golden.py", "reference_build": "oracle/Makefile (stock -d, patched -x)", "cases": []}
    with tempfile.TemporaryDirectory() as tmp:
        for f in sorted(os.listdir(inputs_dir)):
            src = os.path.join(inputs_dir, f)
            data = open(src, "rb").read()
            for mode in ("markov", "huffman"):
                s, t, d = (os.path.join(tmp, x) for x in ("s", "t", "d"))
                run(STOCK, [src, "-o", s, "-h" if mode == "huffman" else "-", "-d", t])
                stream, table = open(s, "rb").read(), open(t, "rb").read()
                roundtrip = None
                if len(table) > 0:
                    run(PATCHED, [s, "-o", d, "-xh" if mode == "huffman" else "-x", "-e", t])
                    roundtrip = open(d, "rb").read() == data
                case = {
                    "input": f, "mode": mode, "input_bytes": len(data),
                    "stream_bytes": len(stream), "stream_sha256": hashlib.sha256(stream).hexdigest(),
                    "table_bytes": len(table), "table_sha256": hashlib.sha256(table).hexdigest(),
                    "header": "%02x" % stream[0],
                    "table_b64": base64.b64encode(table).decode(),
                    "patched_roundtrip_ok": roundtrip,
                }
                if len(stream) <= INLINE_LIMIT:
                    case["stream_b64"] = base64.b64encode(stream).decode()
                golden["cases"].append(case)
                print("%-24s %-8s in=%8d stream=%8d table=%5d rt=%s" % (f, mode, len(data), len(stream), len(table), roundtrip))
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(golden, fh, indent=1)


if __name__ == "__main__":
    main()
