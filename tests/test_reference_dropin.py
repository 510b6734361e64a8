"""The drop-in boundary, demonstrated: the reference's OWN command line (its main.cpp, histogram loop, tree builder,
table loader / writer, all compiled from /root/reference/src) with its two hot loops — i_coding_provider::compress and
::decompress (src/coding.cpp:61-160) — replaced by oracle/gpu_coding_provider.cpp, the binding INTEGRATION.md §2 shows,
linked against libmh_gpu.so (oracle/_ref/markovhuffman_gpu, built by `make -C oracle ref_gpu`).

The GPU test is the reference's own test flow (test/main.py:68-105): for every input, encode with `-h -d` and with
`- -d`, decode with `-xh -e` / `-x -e`, compare with the input — plus what the reference never pinned: the compressed
files and the table files equal the stock reference build's (tests/golden/golden.json)."""
import filecmp
import hashlib
import os
import subprocess

import pytest

from conftest import GOLDEN_DIR, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "markovhuffman_gpu")
INPUTS = ["input_a.txt", "input_b.txt", "input_ipsum.txt", "input_wiki_cpp.txt", "input_wiki_cpp.html"]

needs_bin = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/markovhuffman_gpu is not built (needs /root/reference: make -C oracle ref_gpu)")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@needs_bin
def test_without_a_device_the_bound_reference_fails_loudly(tmp_path):
    """CPU box: the two loops have no CPU path behind the boundary; the reference's CLI reports the error and exits 1."""
    import importlib
    mh = importlib.import_module("markov-huffman-coding_b200")
    if mh.device_count() > 0:
        pytest.skip("a CUDA device is present")
    src = os.path.join(GOLDEN_DIR, "inputs", "input_b.txt")
    r = subprocess.run([BIN, src, "-o", str(tmp_path / "o"), "-", "-d", str(tmp_path / "t")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@needs_bin
@pytest.mark.gpu
@pytest.mark.parametrize("name", INPUTS)
def test_reference_test_flow_through_the_gpu_binding(name, tmp_path):
    golden = {(c["input"], c["mode"]): c for c in load_golden()}
    src = os.path.join(GOLDEN_DIR, "inputs", name)
    for mode, enc_flag, dec_flag in (("huffman", "-h", "-xh"), ("markov", "-", "-x")):   # test/main.py:17-50
        comp, table, back = (str(tmp_path / (name + "." + mode + ext)) for ext in (".cm", ".e", ".dm"))
        r = subprocess.run([BIN, src, "-o", comp, enc_flag, "-d", table], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([BIN, comp, "-o", back, dec_flag, "-e", table], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert filecmp.cmp(src, back, shallow=False)                        # test/main.py:77
        want = golden[(name, mode)]
        assert sha(comp) == want["stream_sha256"] and sha(table) == want["table_sha256"]
        # -e compress with the table just written gives the same file (the loader + the GPU encoder)
        again = str(tmp_path / (name + "." + mode + ".again"))
        r = subprocess.run([BIN, src, "-o", again] + ([enc_flag] if mode == "huffman" else []) + ["-e", table], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert filecmp.cmp(comp, again, shallow=False)


@needs_bin
@pytest.mark.gpu
def test_the_binding_reports_the_reference_decode_errors(tmp_path):
    src = os.path.join(GOLDEN_DIR, "inputs", "input_ipsum.txt")
    comp, table, table_h = str(tmp_path / "c"), str(tmp_path / "t"), str(tmp_path / "th")
    assert subprocess.run([BIN, src, "-o", comp, "-", "-d", table], capture_output=True).returncode == 0
    assert subprocess.run([BIN, src, "-o", str(tmp_path / "ch"), "-h", "-d", table_h], capture_output=True).returncode == 0
    bad = str(tmp_path / "bad")
    open(bad, "wb").write(b"\x77" + open(comp, "rb").read()[1:])
    r = subprocess.run([BIN, bad, "-o", str(tmp_path / "o1"), "-x", "-e", table], capture_output=True, text=True)
    assert r.returncode == 1 and "Input appears corrupt" in r.stderr                      # src/coding.cpp:103-106
    r = subprocess.run([BIN, comp, "-o", str(tmp_path / "o2"), "-xh", "-e", table_h], capture_output=True, text=True)
    assert r.returncode == 1 and "does not match provided encoding table" in r.stderr      # src/coding.cpp:107-110
