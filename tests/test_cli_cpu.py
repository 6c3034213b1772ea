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
