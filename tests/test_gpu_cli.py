"""The drop-in executable on a B200: same command line, same output format, parity with the golden outputs."""
import hashlib
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "kaarme")
CASES = json.load(open(os.path.join(GOLDEN, "golden.json")))


def run_cli(args, cwd=None):
    return subprocess.run([EXE] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=cwd, text=True)


def sorted_sha(path):
    with open(path, "rb") as f:
        lines = sorted(f.read().splitlines(keepends=True))
    return len(lines), hashlib.sha256(b"".join(lines)).hexdigest()


@pytest.mark.parametrize("case", [c for c in CASES if c["mode"] == 0 and c["k"] in (21, 51, 255) and c["a"] == 2
                                  and c["input"] != "g4_polya.fasta"],
                         ids=lambda c: f"{c['input']}-k{c['k']}-{'b' if c['unique'] else 's'}")
def test_cli_matches_reference_output(case, tmp_path):
    out = tmp_path / "out.txt"
    args = [os.path.join(GOLDEN, case["input"]), case["k"], "-m", 0, "-a", case["a"], "-t", 4, "-o", out]
    args += ["-b", "-u", case["unique"], "-f", case["fpr"]] if case["unique"] else ["-s", case["slots"]]
    p = run_cli(args)
    assert p.returncode == 0, p.stderr
    assert sorted_sha(out) == (case["n_lines"], case["sha256"])
    assert "Running settings:" in p.stdout and "Time used to build hash table:" in p.stdout
    if not case["unique"]:
        assert f"Hash table size is: {case['table_slots']}" in p.stdout


def test_cli_default_output_name_and_count_wrap(tmp_path):
    p = run_cli([os.path.join(GOLDEN, "g4_polya.fasta"), 21, "-m", 0, "-s", 200000], cwd=tmp_path)
    assert p.returncode == 0, p.stderr
    case = [c for c in CASES if c["input"] == "g4_polya.fasta" and c["mode"] == 0][0]
    assert sorted_sha(tmp_path / "g4_polya.kaarme_counts") == (case["n_lines"], case["sha256"])


def test_cli_plain_text_input_and_ill_formed(tmp_path):
    case = [c for c in CASES if c["input"] == "g3_plain.txt" and c["k"] == 21 and c["mode"] == 0 and c["a"] == 1][0]
    out = tmp_path / "o"
    p = run_cli([os.path.join(GOLDEN, "g3_plain.txt"), 21, "-m", 0, "-a", 1, "-s", 200000, "-o", out])
    assert p.returncode == 0 and "ONE-STR-PER-LINE" in p.stdout
    assert sorted_sha(out) == (case["n_lines"], case["sha256"])
    bad = tmp_path / "bad.fasta"
    bad.write_text("ACGT\n")
    p = run_cli([bad, 21, "-s", 1000])
    assert p.returncode == 1 and "is ill-formed" in p.stderr


def test_cli_table_full_exits_nonzero(tmp_path):
    p = run_cli([os.path.join(GOLDEN, "g5_long.fasta"), 31, "-m", 0, "-s", 1000, "-o", tmp_path / "o"])
    assert p.returncode == 1 and "Hash table is full" in p.stdout


@pytest.mark.parametrize("case", [c for c in CASES if c["mode"] == 2 and c["k"] in (21, 51, 127) and c["a"] == 2
                                  and c["input"] in ("g1_multiline.fasta", "g5_long.fasta", "g4_polya.fasta")],
                         ids=lambda c: f"{c['input']}-k{c['k']}-{'b' if c['unique'] else 's'}")
def test_cli_kaarme_mode(case, tmp_path):
    """-m 2 (the CLI default): counted, compacted to 8-byte slots, decoded on export; same output as the reference"""
    out, js = tmp_path / "out.txt", tmp_path / "stats.json"
    args = [os.path.join(GOLDEN, case["input"]), case["k"], "-a", case["a"], "-t", 4, "-o", out, "--stats-json", js]
    args += ["-b", "-u", case["unique"], "-f", case["fpr"]] if case["unique"] else ["-s", case["slots"]]
    p = run_cli(args)
    assert p.returncode == 0, p.stderr
    assert sorted_sha(out) == (case["n_lines"], case["sha256"])
    assert "Starting atomic variable pointer hash table" in p.stdout and "Written k-mers:" in p.stdout
    st = json.load(open(js))
    assert st["kaarme"]["kmers"] == st["count"]["distinct"]
    assert st["kaarme"]["bytes"] == 8 * st["kaarme"]["kmers"] + 8 * ((case["k"] + 31) // 32) * st["kaarme"]["roots"]
