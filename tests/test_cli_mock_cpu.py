"""The drop-in executable's HOST code, end to end on a CPU.

host/kaarme_main.cpp is host logic around the C ABI (argument handling, input sharding for --gpus N, the reader ring,
context feeds, record / text sinks, the parallel writer, the Kaarme file format).  tests/native/mock_abi.cpp implements
that ABI on the CPU with the oracle as the counter (a test double: it lives under tests/, the product never loads it),
and the CLI source is linked a second time against it (tests/native/_build/kaarme_mock).  Everything the host side can
get wrong -- a byte fed twice or not at all at a chunk / rank boundary, a context byte counted, a torn or dropped
output buffer, a bad file header -- changes the output, which is compared with the reference binary's golden outputs
(tests/golden/golden.json) or with the oracle.  The kernels themselves are covered by the -m gpu tests."""
import hashlib
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

NATIVE = os.path.join(ROOT, "tests", "native")
BUILD = os.path.join(NATIVE, "_build")
PKG = os.path.join(ROOT, "canonical-k-mer-hash-table_b200")
CASES = json.load(open(os.path.join(GOLDEN, "golden.json")))


@pytest.fixture(scope="module")
def exe(oracle):
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "libkaarme_gpu_mock.so")
    out = os.path.join(BUILD, "kaarme_mock")
    odir = os.path.join(ROOT, "oracle")
    srcs = [os.path.join(NATIVE, "mock_abi.cpp"), os.path.join(PKG, "host", "kaarme_main.cpp"), os.path.join(PKG, "host", "kg_reader.hpp"),
            os.path.join(PKG, "host", "kg_writer.hpp"), os.path.join(PKG, "csrc", "kg_text.cuh"), os.path.join(ROOT, "include", "kaarme_gpu.h")]
    if not (os.path.exists(lib) and os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared", "-o", lib, srcs[0], "-L" + odir, "-loracle",
                        "-Wl,-rpath," + odir], check=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-I" + os.path.join(ROOT, "include"), "-o", out, srcs[1],
                        "-L" + BUILD, "-lkaarme_gpu_mock", "-lz", "-Wl,-rpath,$ORIGIN"], check=True)
    return out


def run(exe, args, cwd=None, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=cwd, text=True, env=e)


def sorted_sha(path):
    with open(path, "rb") as f:
        lines = sorted(f.read().splitlines(keepends=True))
    return len(lines), hashlib.sha256(b"".join(lines)).hexdigest()


def sha_of(text):
    lines = sorted(text.splitlines(keepends=True))
    return len(lines), hashlib.sha256(b"".join(lines)).hexdigest()


def size_args(case):
    return ["-b", "-u", case["unique"], "-f", case["fpr"]] if case["unique"] else ["-s", case["slots"]]


SEL = [c for c in CASES if c["k"] in (21, 51, 255) and (c["a"] == 2 or c["unique"] is None)]


@pytest.mark.parametrize("case", SEL, ids=lambda c: f"{c['input']}-k{c['k']}-m{c['mode']}-a{c['a']}-{'b' if c['unique'] else 's'}")
def test_cli_host_path_reproduces_reference_outputs(exe, case, tmp_path):
    """file -> reader ring -> feeds -> text sink -> parallel writer -> file == the reference binary's output"""
    out = tmp_path / "out.txt"
    p = run(exe, [os.path.join(GOLDEN, case["input"]), case["k"], "-m", case["mode"], "-a", case["a"], "-t", 4, "-o", out] + size_args(case))
    assert p.returncode == 0, p.stderr
    assert sorted_sha(out) == (case["n_lines"], case["sha256"])
    if not case["unique"]:
        assert f"Hash table size is: {case['table_slots']}" in p.stdout


@pytest.mark.parametrize("gpus", [2, 3, 5, 8])
@pytest.mark.parametrize("name,k,mode", [("g1_multiline.fasta", 21, 0), ("g5_long.fasta", 51, 0), ("g2_reads.fa", 255, 0),
                                         ("g3_plain.txt", 21, 0), ("g2_reads.fa", 21, 2)])
def test_cli_sharded_input(exe, name, k, mode, gpus, tmp_path):
    """--gpus N: every rank reads its byte range plus the k-1 bases of context before it; no k-mer is lost or counted
    twice at a rank boundary, whatever falls there (mid-line, mid-header, a newline)"""
    case = [c for c in CASES if c["input"] == name and c["k"] == k and c["mode"] == mode and c["a"] == 2][0]
    out = tmp_path / "out.txt"
    p = run(exe, [os.path.join(GOLDEN, name), k, "-m", mode, "-a", 2, "-t", 6, "-o", out, "--gpus", gpus] + size_args(case))
    assert p.returncode == 0, p.stderr
    assert sorted_sha(out) == (case["n_lines"], case["sha256"])
    assert f"GPU x{gpus}" in p.stdout


@pytest.fixture(scope="module")
def big_fasta(tmp_path_factory):
    """~5.3 MB: several 1 MiB batches, so the 3-buffer ring wraps and chunk boundaries fall inside lines and headers"""
    rng = np.random.default_rng(21)
    g = rng.integers(0, 4, 400_000)
    recs = []
    for i in range(520):
        p = int(rng.integers(0, 400_000 - 10_000))
        r = np.frombuffer(b"ACGT", np.uint8)[g[p:p + 10_000]].tobytes().decode()
        if i % 7 == 0:
            r = r[:5000] + "N" + r[5001:]
        recs.append(f">read_{i} some description text\n" + "\n".join(r[j:j + 70] for j in range(0, len(r), 70)) + "\n")
    path = tmp_path_factory.mktemp("big") / "big.fasta"
    path.write_text("".join(recs))
    return str(path)


@pytest.mark.parametrize("extra", [[], ["--host-format"], ["--gpus", 3], ["--gpus", 4, "--host-format"]],
                         ids=["text", "hostfmt", "gpus3", "gpus4-hostfmt"])
def test_cli_many_batches(exe, oracle, big_fasta, extra, tmp_path):
    data = open(big_fasta, "rb").read()
    assert len(data) > 5 << 20
    want = sha_of(oracle.count(data, 31).text(2, oracle.TABLE_PLAIN))
    out = tmp_path / "out.txt"
    p = run(exe, [big_fasta, 31, "-m", 0, "-a", 2, "-t", 8, "-s", 2_000_000, "-o", out, "--batch-mb", 1] + extra)
    assert p.returncode == 0, p.stderr
    assert sorted_sha(out) == want


def test_cli_kaarme_file_roundtrip_and_rejects(exe, tmp_path):
    """--dump-kaarme writes header + slots + roots; --from-kaarme reads them back (size and header checked first)"""
    case = [c for c in CASES if c["input"] == "g2_reads.fa" and c["k"] == 51 and c["mode"] == 2 and c["a"] == 2 and c["unique"] is None][0]
    out, dump, out2 = tmp_path / "out.txt", tmp_path / "g2.kaarme", tmp_path / "out2.txt"
    p = run(exe, [os.path.join(GOLDEN, "g2_reads.fa"), 51, "-a", 2, "-t", 4, "-s", case["slots"], "-o", out, "--dump-kaarme", dump])
    assert p.returncode == 0, p.stderr
    assert sorted_sha(out) == (case["n_lines"], case["sha256"])
    blob = dump.read_bytes()
    magic, ver, k, W, flags, n_kmers, n_roots = struct.unpack_from("<8sIIIIQQ", blob, 0)
    assert (magic, ver, k, W, flags) == (b"KAARMEG1", 1, 51, 2, 0) and len(blob) == 64 + 8 * n_kmers + 8 * W * n_roots
    for extra in ([], ["--host-format"]):
        p = run(exe, [dump, 51, "--from-kaarme", "-a", 2, "-o", out2] + extra)
        assert p.returncode == 0, p.stderr + p.stdout
        assert sorted_sha(out2) == (case["n_lines"], case["sha256"])
        assert f"Written k-mers: {case['n_lines']}" in p.stdout and f"Skipped k-mers: {n_kmers - case['n_lines']}" in p.stdout
    # -a above every count: an empty file, not an error
    p = run(exe, [dump, 51, "--from-kaarme", "-a", 100000, "-o", out2])
    assert p.returncode == 0 and out2.read_bytes() == b""
    # rejects: wrong KLEN, truncated, trailing garbage, wrong magic, not -m 2, --gpus 2
    assert run(exe, [dump, 31, "--from-kaarme", "-o", out2]).returncode == 1
    for name, data in (("cut", blob[:-8]), ("long", blob + b"\0" * 8), ("magic", b"X" + blob[1:]), ("tiny", blob[:10])):
        f = tmp_path / f"{name}.kaarme"
        f.write_bytes(data)
        p = run(exe, [f, 51, "--from-kaarme", "-o", out2])
        assert p.returncode == 1 and "ill-formed" in p.stderr, name
    p = run(exe, [os.path.join(GOLDEN, "g2_reads.fa"), 51, "-m", 0, "-s", 1000, "--dump-kaarme", dump, "-o", out2])
    assert p.returncode == 1 and "--dump-kaarme needs -m 2" in p.stderr
    p = run(exe, [os.path.join(GOLDEN, "g2_reads.fa"), 51, "-s", 100000, "--gpus", 2, "--dump-kaarme", dump, "-o", out2])
    assert p.returncode == 1


def test_cli_output_conventions(exe, tmp_path):
    fa = os.path.join(GOLDEN, "g1_multiline.fasta")
    # default output name (main.cpp:189-191) in the working directory
    p = run(exe, [fa, 21, "-m", 0, "-s", 200000], cwd=tmp_path)
    assert p.returncode == 0 and (tmp_path / "g1_multiline.kaarme_counts").exists()
    # -a 0 writes nothing and creates nothing (parallel_parser.hpp:860-861)
    p = run(exe, [fa, 21, "-m", 0, "-s", 200000, "-a", 0, "-o", tmp_path / "none.txt"])
    assert p.returncode == 0 and not (tmp_path / "none.txt").exists()
    # table full: message + exit 1 (kmer_hash_table.cpp:2552-2556)
    p = run(exe, [fa, 21, "-m", 0, "-s", 100, "-o", tmp_path / "full.txt"])
    assert p.returncode == 1 and "Hash table is full" in p.stdout
    # unwritable output
    p = run(exe, [fa, 21, "-m", 0, "-s", 200000, "-o", tmp_path / "no_such_dir" / "o.txt"])
    assert p.returncode == 1 and "cannot open output file" in p.stderr
    # stats JSON
    js = tmp_path / "s.json"
    p = run(exe, [fa, 21, "-s", 200000, "-o", tmp_path / "o.txt", "--stats-json", js])
    st = json.load(open(js))
    assert p.returncode == 0 and st["gpus"] == 1 and st["count"]["input_kmers"] > 0 and st["written"] == sorted_sha(tmp_path / "o.txt")[0]


def test_cli_without_devices_creates_nothing(exe, tmp_path):
    """the device check comes before anything is created on disk"""
    out = tmp_path / "o.txt"
    p = run(exe, [os.path.join(GOLDEN, "g1_multiline.fasta"), 21, "-s", 1000, "-o", out], env={"KG_MOCK_DEVICES": "0"})
    assert p.returncode == 2 and "no CPU fallback" in p.stderr and not out.exists()
    p = run(exe, [os.path.join(GOLDEN, "g1_multiline.fasta"), 21, "-s", 1000, "-o", out, "--gpus", 4], env={"KG_MOCK_DEVICES": "2"})
    assert p.returncode == 2 and not out.exists()


def test_cli_output_to_a_pipe(exe, tmp_path):
    """-o /dev/stdout: not seekable, the writer falls back to sequential writes; log lines and k-mer lines share the
    stream, so only the k-mer lines are compared"""
    case = [c for c in CASES if c["input"] == "g1_multiline.fasta" and c["k"] == 21 and c["mode"] == 0 and c["a"] == 2][0]
    p = subprocess.run([exe, os.path.join(GOLDEN, "g1_multiline.fasta"), "21", "-m", "0", "-a", "2", "-s", str(case["slots"]),
                        "-o", "/dev/stdout"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0
    kmer_lines = [l for l in p.stdout.splitlines(keepends=True) if len(l) > 22 and l[21:22] == b" " and set(l[:21]) <= set(b"ACGT")]
    assert sha_of(b"".join(kmer_lines)) == (case["n_lines"], case["sha256"])


def test_cli_reference_bloom_flag_reaches_the_config(exe, tmp_path):
    """--reference-bloom sets KG_CFG_REFERENCE_BLOOM in kg_config.reserved (the test double reports the flags it saw)"""
    case = [c for c in CASES if c["input"] == "g5_long.fasta" and c["k"] == 51 and c["mode"] == 0 and c["a"] == 2 and c["unique"]][0]
    out = tmp_path / "o.txt"
    for extra, want in (([], "0"), (["--reference-bloom"], "1")):
        p = run(exe, [os.path.join(GOLDEN, "g5_long.fasta"), 51, "-m", 0, "-a", 2, "-o", out] + size_args(case) + extra,
                env={"KG_MOCK_SHOW_FLAGS": "1"})
        assert p.returncode == 0 and f"mock: config flags {want}" in p.stderr
        assert sorted_sha(out) == (case["n_lines"], case["sha256"])


@pytest.mark.parametrize("gpus", [2, 4])
def test_cli_kaarme_mode_on_several_shards(exe, gpus, tmp_path):
    """--gpus N -m 2: every shard compacts its own structure and the output is decoded from them (the reference's
    default table mode, main.cpp:137, on the sharded path); per-shard bytes per k-mer are reported"""
    case = [c for c in CASES if c["input"] == "g2_reads.fa" and c["k"] == 51 and c["mode"] == 2 and c["a"] == 2 and not c["unique"]][0]
    out = tmp_path / "out.txt"
    p = run(exe, [os.path.join(GOLDEN, "g2_reads.fa"), 51, "-m", 2, "-a", 2, "-t", 6, "-o", out, "--gpus", gpus] + size_args(case))
    assert p.returncode == 0, p.stderr
    assert sorted_sha(out) == (case["n_lines"], case["sha256"])
    assert p.stdout.count("  shard ") == gpus and "Kaarme bytes:" in p.stdout


def test_cli_gzip_input(exe, oracle, big_fasta, tmp_path):
    """.gz input (the reference's zlib path does not work; here the stream is inflated in order into the same ring):
    format taken from the name under the .gz, one member or several, many batches; truncated streams and --gpus > 1
    are refused"""
    import gzip
    case = [c for c in CASES if c["input"] == "g2_reads.fa" and c["k"] == 21 and c["mode"] == 0 and c["a"] == 2 and c["unique"] is None][0]
    raw = open(os.path.join(GOLDEN, "g2_reads.fa"), "rb").read()
    gz = tmp_path / "g2_reads.fa.gz"
    gz.write_bytes(gzip.compress(raw))
    out = tmp_path / "out.txt"
    p = run(exe, [gz, 21, "-m", 0, "-a", 2, "-s", case["slots"], "-o", out])
    assert p.returncode == 0, p.stderr
    assert "gzip compressed:          yes" in p.stdout and "input format:             FASTA" in p.stdout
    assert sorted_sha(out) == (case["n_lines"], case["sha256"])
    # two members back to back = the concatenated text; cut in the middle of a record
    cut = raw.index(b"\n", len(raw) // 2) + 10
    gz2 = tmp_path / "two.fa.gz"
    gz2.write_bytes(gzip.compress(raw[:cut]) + gzip.compress(raw[cut:]))
    p = run(exe, [gz2, 21, "-m", 0, "-a", 2, "-s", case["slots"], "-o", out])
    assert p.returncode == 0 and sorted_sha(out) == (case["n_lines"], case["sha256"])
    # Bloom mode reads the stream twice
    bcase = [c for c in CASES if c["input"] == "g5_long.fasta" and c["k"] == 51 and c["mode"] == 0 and c["a"] == 2 and c["unique"]][0]
    gz3 = tmp_path / "g5.fasta.gz"
    gz3.write_bytes(gzip.compress(open(os.path.join(GOLDEN, "g5_long.fasta"), "rb").read()))
    p = run(exe, [gz3, 51, "-m", 0, "-a", 2, "-o", out] + size_args(bcase))
    assert p.returncode == 0 and sorted_sha(out) == (bcase["n_lines"], bcase["sha256"])
    # many 1 MiB batches out of one stream
    data = open(big_fasta, "rb").read()
    gz4 = tmp_path / "big.fasta.gz"
    gz4.write_bytes(gzip.compress(data, 1))
    p = run(exe, [gz4, 31, "-m", 0, "-a", 2, "-s", 2_000_000, "-o", out, "--batch-mb", 1])
    assert p.returncode == 0 and sorted_sha(out) == sha_of(oracle.count(data, 31).text(2, oracle.TABLE_PLAIN))
    # default output name strips one extension, like main.cpp:189-191
    p = run(exe, [gz, 21, "-m", 0, "-s", case["slots"]], cwd=tmp_path)
    assert p.returncode == 0 and (tmp_path / "g2_reads.fa.kaarme_counts").exists()
    # refused: truncated stream, sharding, content that does not match the name
    bad = tmp_path / "cut.fa.gz"
    bad.write_bytes(gz.read_bytes()[:-200])
    p = run(exe, [bad, 21, "-m", 0, "-s", case["slots"], "-o", out])
    assert p.returncode == 1 and "gzip error" in p.stderr
    p = run(exe, [gz, 21, "-m", 0, "-s", case["slots"], "-o", out, "--gpus", 2])
    assert p.returncode == 1 and "--gpus 1" in p.stderr
    plain_named_fa = tmp_path / "plain.fa.gz"
    plain_named_fa.write_bytes(gzip.compress(b"ACGTACGT\n"))
    p = run(exe, [plain_named_fa, 5, "-s", 100])
    assert p.returncode == 1 and "ill-formed" in p.stderr
