"""The C++ host driver (bin/markovhuffman) against the reference's own CLI behaviour: flag grammar and validation
(CPU, no device needed because the errors come first) and, on a GPU, byte-identical files versus the reference
binaries built from the reference's sources (oracle/_ref)."""
import os
import subprocess

import pytest

import oracle_py as o
from conftest import GOLDEN_DIR
from mhlib import load

mh = load()
CLI = mh.CLI_PATH
INPUTS = os.path.join(GOLDEN_DIR, "inputs")


def run(args, exe=CLI, stdin=None):
    return subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, input=stdin)


def test_help_and_validation_errors_match_reference():
    have_ref = os.path.exists(o.REF_STOCK)
    cases = [
        [],                                                   # usage, exit 1
        ["-x"],                                               # no input
        ["-o"],                                               # flag without value, then no input
        ["in.bin", "-e", "a", "-d", "b"],                     # both -e and -d
        ["in.bin", "-x"],                                     # extract without table
        ["/nonexistent/input.bin", "-o", "/tmp/mh_cli_o"],    # cannot open input
        ["-q", "-z"],                                         # unknown options, then no input
    ]
    for args in cases:
        mine = run(args)
        assert mine.returncode == 1, args
        if have_ref:
            ref = run(args, exe=o.REF_STOCK)
            assert ref.returncode == mine.returncode, args
            assert ref.stderr == mine.stderr, (args, ref.stderr, mine.stderr)
            assert ref.stdout == mine.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["input_b.txt", "input_wiki_cpp.html", "edge_fib40_256k.bin"])
@pytest.mark.parametrize("simple", [False, True])
def test_cli_files_identical_to_reference(tmp_path, name, simple):
    if not os.path.exists(o.REF_STOCK):
        pytest.skip("oracle/_ref not present")
    src = os.path.join(INPUTS, name)
    mode = ["-h"] if simple else ["-"]           # the reference's test script passes a bare "-" for Markov mode
    mine, ref = {}, {}
    for tag, exe, store in (("mine", CLI, mine), ("ref", o.REF_STOCK, ref)):
        out, tab = str(tmp_path / (tag + ".c")), str(tmp_path / (tag + ".e"))
        p = run([src, "-o", out] + mode + ["-d", tab], exe=exe)
        assert p.returncode == 0, p.stderr
        store["stream"], store["table"], store["stderr"] = open(out, "rb").read(), open(tab, "rb").read(), p.stderr
    assert mine["stream"] == ref["stream"]
    assert mine["table"] == ref["table"]
    # same progress lines, modulo the file names in them
    assert mine["stderr"].replace(b"mine.", b"X.") == ref["stderr"].replace(b"ref.", b"X.")
    # extract with our CLI and with the patched reference: both restore the input
    dec = str(tmp_path / "mine.d")
    p = run([str(tmp_path / "mine.c"), "-o", dec, "-xh" if simple else "-x", "-e", str(tmp_path / "mine.e")])
    assert p.returncode == 0, p.stderr
    assert open(dec, "rb").read() == open(src, "rb").read()
    # -e compress equals -d compress
    again = str(tmp_path / "again.c")
    p = run([src, "-o", again] + (["-h"] if simple else []) + ["-e", str(tmp_path / "mine.e")])
    assert p.returncode == 0, p.stderr
    assert open(again, "rb").read() == mine["stream"]


@pytest.mark.gpu
def test_cli_debug_dump_stdout_and_pipe_header_quirk(tmp_path):
    if not os.path.exists(o.REF_STOCK):
        pytest.skip("oracle/_ref not present")
    src = os.path.join(INPUTS, "input_b.txt")
    for mode in ([], ["-h"]):
        mine = run([src, "-o", str(tmp_path / "o1"), "-g"] + mode)
        ref = run([src, "-o", str(tmp_path / "o2"), "-g"] + mode, exe=o.REF_STOCK)
        assert mine.returncode == 0 and mine.stdout == ref.stdout
        # no -o: the stream goes to stdout; on a pipe the reference cannot seek back, so the 0x80 placeholder stays
        # first and the header byte lands at the end
        mine = run([src] + mode)
        ref = run([src] + mode, exe=o.REF_STOCK)
        assert mine.stdout == ref.stdout and mine.stdout[:1] == b"\x80"


@pytest.mark.gpu
def test_cli_decode_error_messages(tmp_path):
    src = os.path.join(INPUTS, "input_b.txt")
    c_m, t_m, c_h, t_h = (str(tmp_path / x) for x in ("m.c", "m.e", "h.c", "h.e"))
    assert run([src, "-o", c_m, "-d", t_m]).returncode == 0
    assert run([src, "-o", c_h, "-h", "-d", t_h]).returncode == 0
    p = run([c_h, "-o", str(tmp_path / "x"), "-x", "-e", t_m])           # Huffman stream, Markov table
    assert p.returncode == 1 and b"File encoding method does not match provided encoding table." in p.stderr
    p = run([c_m, "-o", str(tmp_path / "x"), "-xh", "-e", t_m])          # -h with a Markov table file
    assert p.returncode == 1 and b"Incorrect encoding table provided for current operation; expected simple Huffman, found Markov-Huffman" in p.stderr
    bad = str(tmp_path / "bad.c")
    open(bad, "wb").write(b"\x80" + open(c_m, "rb").read()[1:])
    p = run([bad, "-o", str(tmp_path / "x"), "-x", "-e", t_m])
    assert p.returncode == 1 and b"Input appears corrupt." in p.stderr


@pytest.mark.gpu
def test_cli_extract_grows_its_output_buffer(tmp_path):
    """1 bit per symbol: the decoded file is 8x the stream, far beyond the extractor's first guess."""
    src = str(tmp_path / "run.bin")
    open(src, "wb").write(b"q" * 3_000_001)
    c, t, d = (str(tmp_path / x) for x in ("c", "t", "d"))
    assert run([src, "-o", c, "-d", t]).returncode == 0
    assert os.path.getsize(c) == 1 + (3_000_001 + 7) // 8
    assert run([c, "-o", d, "-x", "-e", t]).returncode == 0
    assert open(d, "rb").read() == open(src, "rb").read()


@pytest.mark.gpu
@pytest.mark.parametrize("simple", [False, True], ids=["markov", "huffman"])
def test_cli_streams_files_larger_than_its_device_buffers(tmp_path, simple, monkeypatch):
    """with the device buffers capped far below the file size (MH_CLI_MAX_BYTES; by default a fifth of the free
    device memory) the CLI streams the file through in chunks and still writes the reference's bytes."""
    if not os.path.exists(o.REF_STOCK):
        pytest.skip("oracle/_ref not present")
    src = os.path.join(INPUTS, "input_wiki_cpp.html")      # 343 KB
    mode = ["-h"] if simple else []
    ref_c, ref_e = str(tmp_path / "ref.c"), str(tmp_path / "ref.e")
    assert run([src, "-o", ref_c] + mode + ["-d", ref_e], exe=o.REF_STOCK).returncode == 0
    monkeypatch.setenv("MH_CLI_MAX_BYTES", "40000")
    out_c, out_e, out_d = str(tmp_path / "mine.c"), str(tmp_path / "mine.e"), str(tmp_path / "mine.d")
    p = subprocess.run([CLI, src, "-o", out_c] + mode + ["-d", out_e], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr
    assert open(out_c, "rb").read() == open(ref_c, "rb").read()
    assert open(out_e, "rb").read() == open(ref_e, "rb").read()
    p = subprocess.run([CLI, out_c, "-o", out_d, "-xh" if simple else "-x", "-e", out_e], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr
    assert open(out_d, "rb").read() == open(src, "rb").read()


@pytest.mark.gpu
@pytest.mark.parametrize("simple", [False, True], ids=["markov", "huffman"])
@pytest.mark.parametrize("ranks", ["2", "3"])
def test_cli_cuts_a_file_over_several_ranks(tmp_path, simple, ranks, monkeypatch):
    """SURVEY §8e through the command line: MH_CLI_GPUS ranks (one host thread each; on a box with fewer GPUs they share
    devices through the in-process transport) compress ONE file by byte ranges and extract it by bit ranges cut at
    arbitrary bits. Same files as the reference binary writes."""
    if not os.path.exists(o.REF_STOCK):
        pytest.skip("oracle/_ref not present")
    src = os.path.join(INPUTS, "input_wiki_cpp.html")      # 343 KB
    mode = ["-h"] if simple else []
    ref_c, ref_e = str(tmp_path / "ref.c"), str(tmp_path / "ref.e")
    assert run([src, "-o", ref_c] + mode + ["-d", ref_e], exe=o.REF_STOCK).returncode == 0
    monkeypatch.setenv("MH_CLI_GPUS", ranks)
    out_c, out_e, out_d = str(tmp_path / "mine.c"), str(tmp_path / "mine.e"), str(tmp_path / "mine.d")
    p = subprocess.run([CLI, src, "-o", out_c] + mode + ["-d", out_e], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr
    assert open(out_c, "rb").read() == open(ref_c, "rb").read()
    assert open(out_e, "rb").read() == open(ref_e, "rb").read()
    p = subprocess.run([CLI, out_c, "-o", out_d, "-xh" if simple else "-x", "-e", out_e], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr
    assert open(out_d, "rb").read() == open(src, "rb").read()
    # a file too short to cut falls back to one GPU without a word
    tiny = os.path.join(INPUTS, "input_b.txt")
    assert run([tiny, "-o", out_c, "-d", out_e]).returncode == 0
    assert run([out_c, "-o", out_d, "-x", "-e", out_e]).returncode == 0
    assert open(out_d, "rb").read() == open(tiny, "rb").read()
