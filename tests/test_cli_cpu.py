"""CLI argument surface (main.cpp:134-151) checked without a GPU: exit codes and messages of the reference's
CLI11 parser, format sniffing, and the loud failure when no B200 is present."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

EXE = os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "kaarme")
FA = os.path.join(GOLDEN, "g1_multiline.fasta")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(EXE):
        import __graft_entry__
        __graft_entry__.build()


def run(*args):
    return subprocess.run([EXE] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


# (args, exit code, message) measured on the reference binary (oracle/_ref/kaarme) in the build container
@pytest.mark.parametrize("args,rc,msg", [
    ([], 106, "INPUT is required"),
    ([FA], 106, "KLEN is required"),
    ([FA, 21], 106, "Exactly 1 option from [-s,--hash-tab-size,-u,--unq-kmers] is required"),
    ([FA, 21, "-s", 100, "-u", 100, "-b"], 106, "is required and 2 were given"),
    ([FA, 21, "-u", 100], 107, "--unq-kmers requires --use-bfilter"),
    ([FA, 21, "-s", 100, "-b"], 107, "--use-bfilter requires --unq-kmers"),
    ([FA, 21, "-s", 100, "-f", 0.5], 107, "--bfilter-fpr requires --use-bfilter"),
    (["/nonexistent.fa", 21, "-s", 100], 105, "INPUT: File does not exist: /nonexistent.fa"),
    ([FA, 0, "-s", 100], 105, "KLEN: Value 0 not in range"),
    ([FA, 21, "-s", 100, "-m", 3], 105, "--hash-table-type: Value 3 not in range 0 to 2"),
    ([FA, 21, "-s", 100, "-t", 2], 105, "--threads: Value 2 not in range 3 to 64"),
    ([FA, 21, "-s", 100, "--bogus"], 109, "The following argument was not expected: --bogus"),
    ([FA, 21, "-u", 5, "-b", "-f", 2], 105, "--bfilter-fpr: Value 2 not in range 0.001000 to 0.999000"),
])
def test_cli_errors_match_reference(args, rc, msg):
    p = run(*args)
    assert p.returncode == rc
    assert msg in p.stderr


def test_help():
    p = run("--help")
    assert p.returncode == 0
    for flag in ("-m,--hash-table-type", "-a,--min-k-abu", "-t,--threads", "-o,--output-file", "-b,--use-bfilter",
                 "-f,--bfilter-fpr", "-s,--hash-tab-size", "-u,--unq-kmers"):
        assert flag in p.stdout


def test_ill_formed_input(tmp_path):
    bad = tmp_path / "bad.fasta"
    bad.write_text("ACGT\n")
    p = run(bad, 21, "-s", 1000)
    assert p.returncode == 1 and "is ill-formed" in p.stderr     # main.cpp:168-171


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = run(FA, 21, "-s", 1000)
    assert p.returncode == 2 and "no CPU fallback" in p.stderr


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name,k", [("g1_multiline.fasta", 21), ("g5_long.fasta", 51), ("g2_reads.fa", 31), ("g3_plain.txt", 21)])
def test_cli_sharding_matches_python_and_oracle(oracle, name, k, world):
    """--gpus N host logic of the C++ CLI (make_slice): same slices as kaarme_gpu.slice_context, and counting every
    rank's slice (context fed but not counted) with the oracle reproduces the whole-file count"""
    import importlib
    import numpy as np
    K = importlib.import_module("canonical-k-mer-hash-table_b200").kaarme_gpu
    path = os.path.join(GOLDEN, name)
    data = open(path, "rb").read()
    mode = oracle.PLAIN if name.endswith(".txt") else oracle.FASTA
    p = run(path, k, "-s", 1000, "--gpus", world, "--print-slices")
    assert p.returncode == 0, p.stderr
    rows = [list(map(int, ln.split()[1:])) for ln in p.stdout.splitlines() if ln.startswith("slice ")]
    assert len(rows) == world
    total = {}
    windows = 0
    for r, (rank, ctx_lo, lo, hi, hdr) in enumerate(rows):
        assert rank == r and (lo, hi) == K.shard_ranges(len(data), world)[r]
        assert (ctx_lo, bool(hdr)) == K.slice_context(data, lo, k, K.INPUT_PLAIN if mode == oracle.PLAIN else K.INPUT_FASTA)
        a = oracle.count(data[ctx_lo:hi], k, mode, bool(hdr))
        b = oracle.count(data[ctx_lo:lo], k, mode, bool(hdr))
        windows += a.total_windows - b.total_windows
        for c, sign in ((a, 1), (b, -1)):
            for kk, cc in zip(map(bytes, c.keys.view(np.uint8).reshape(c.n, c.W * 8)), c.counts.tolist()):
                total[kk] = total.get(kk, 0) + sign * cc
    whole = oracle.count(data, k, mode)
    assert windows == whole.total_windows
    want = dict(zip(map(bytes, whole.keys.view(np.uint8).reshape(whole.n, whole.W * 8)), whole.counts.tolist()))
    assert {kk: cc for kk, cc in total.items() if cc} == want


def test_fastq_is_detected_and_declined_like_the_reference(tmp_path):
    """main.cpp:27-68 detects FASTQ by extension + '@'; parallel_parser.hpp:797-800 then prints "Not implemented yet" and
    returns normally -- before any device is touched, so this runs without a GPU"""
    fq = tmp_path / "reads.fastq"
    fq.write_text("@r1\nACGTACGTACGT\n+\nIIIIIIIIIIII\n")
    p = run(fq, 5, "-s", 1000, "-o", tmp_path / "o.txt")
    assert p.returncode == 0 and "input format:             FASTQ" in p.stdout and "Not implemented yet" in p.stdout
    assert not (tmp_path / "o.txt").exists()
    bad = tmp_path / "bad.fq"
    bad.write_text(">r1\nACGT\n")
    p = run(bad, 5, "-s", 1000)
    assert p.returncode == 1 and "is ill-formed" in p.stderr


def test_mode_1_is_declined_like_an_unknown_mode(tmp_path):
    p = run(FA, 21, "-s", 1000, "-m", 1, "-o", tmp_path / "o.txt")
    assert p.returncode == 0 and "Chosen mode not recognized" in p.stdout and not (tmp_path / "o.txt").exists()
