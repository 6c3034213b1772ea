"""Pins oracle/ (the CPU restatement) against the reference: golden vectors minted by the reference binary
(tests/golden/make_golden.py) and, when oracle/_ref/kaarme exists (build container), the binary itself."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

CASES = json.load(open(os.path.join(GOLDEN, "golden.json")))
KATS = json.load(open(os.path.join(GOLDEN, "kats.json")))


def _mode(name):
    return 2 if name.endswith(".txt") else 0


def _read(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def test_xxh64_kats(oracle):
    # SURVEY.md section 8c vectors + tests/golden/kats.json (vendored xxhash.c v0.8.2)
    L = oracle.lib()
    assert L.ko_xxh64_u64(0, 2411) == 0x557C61D855EB7E2E
    assert L.ko_xxh64_u64(1, 3253) == 0xF485676EA5E6C371
    assert L.ko_xxh64_u64(0x0123456789ABCDEF, 0) == 0xEA3C52081E9843EC
    for v, s, h in KATS["xxh64"]:
        assert L.ko_xxh64_u64(v, s) == h


def test_math_kats(oracle):
    L = oracle.lib()
    for n, p in KATS["prime"]:
        assert L.ko_next_prime3mod4(n) == p
    for a, m, x in KATS["inv"]:
        assert L.ko_modinv(a, m) == x
    assert L.ko_modinv(5, 1 << 54) == 3602879701896397
    assert L.ko_next_prime3mod4(8000000) == 8000023


def test_rolling_hash_kats(oracle):
    for s, q, tbm, hf, hb, fc in KATS["roll"]:
        codes = ["ACGT".index(c) for c in s]
        assert oracle.rolling_hashes(codes, len(s), q, bool(tbm)) == (hf, hb)
        c = oracle.count(b">x\n" + s.encode() + b"\n", len(s))
        rc = s[::-1].translate(str.maketrans("ACGT", "TGCA"))
        assert oracle.key_strings(c.keys, len(s)) == [s if fc else rc]


def test_bloom_params(oracle):
    # -u 4e6, fpr .01 => m = 2^26, 7/6 hashes
    assert oracle.bloom_params(4000000, 0.01) == (1 << 26, 7, 6)
    assert oracle.bloom_params(80000000, 0.01)[0] == 1 << 30


@pytest.mark.parametrize("case", [c for c in CASES if c["unique"] is None],
                         ids=lambda c: f"{c['input']}-k{c['k']}-m{c['mode']}-a{c['a']}")
def test_count_golden(oracle, case):
    c = oracle.count(_read(case["input"]), case["k"], _mode(case["input"]))
    txt = c.text(case["a"], case["mode"])
    assert txt.count(b"\n") == case["n_lines"]
    assert hashlib.sha256(txt).hexdigest() == case["sha256"]
    assert oracle.lib().ko_next_prime3mod4(case["slots"]) == case["table_slots"]


@pytest.mark.parametrize("case", [c for c in CASES if c["unique"] is not None],
                         ids=lambda c: f"{c['input']}-k{c['k']}-m{c['mode']}-a{c['a']}-u{c['unique']}-f{c['fpr']}")
def test_bloom_golden(oracle, case):
    data = _read(case["input"])
    c, st = oracle.count_bloom(data, case["k"], case["unique"], case["fpr"], _mode(case["input"]))
    assert (st.new_in_first, st.new_in_second, st.table_slots) == (
        case["new_in_first"], case["new_in_second"], case["table_slots"])
    txt = c.text(case["a"], case["mode"])
    assert txt.count(b"\n") == case["n_lines"]
    assert hashlib.sha256(txt).hexdigest() == case["sha256"]
    # the north-star rule: at -a >= 2 the Bloom output equals the no-Bloom ground truth
    if case["a"] >= 2:
        truth = oracle.count(data, case["k"], _mode(case["input"])).text(case["a"], case["mode"])
        assert txt == truth


def test_edge_inputs(oracle):
    assert oracle.count(b"", 21).n == 0
    assert oracle.count(b">only a header", 21).n == 0
    assert oracle.count(b">h\nACGT\n", 5).n == 0                       # record shorter than k
    c = oracle.count(b">h\nACGT\nACGT\n", 5)                              # newline inside a record is skipped
    assert c.total_windows == 4
    c = oracle.count(b"ACGT\nACGT\n", 5, oracle.PLAIN)                     # PLAIN: every line is its own string
    assert c.total_windows == 0
    c = oracle.count(b"tail of a header\nACGTA\n", 5, oracle.FASTA, starts_in_header=True)
    assert c.total_windows == 1
    c = oracle.count(b">p\nACGT\n", 4)                                    # palindrome keeps the forward string
    assert oracle.key_strings(c.keys, 4) == ["ACGT"]


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "kaarme")),
                    reason="oracle/_ref/kaarme not built (only in the build container)")
@pytest.mark.parametrize("k,mode", [(21, 0), (51, 0), (51, 2), (127, 0), (255, 2)])
def test_live_reference(oracle, tmp_path, k, mode):
    rng = np.random.default_rng(k * 10 + mode)
    g = rng.integers(0, 4, 30000)
    recs = []
    for i in range(40):
        p = int(rng.integers(0, 30000 - 3000))
        r = "".join("ACGT"[x] for x in g[p:p + 3000])
        recs.append(f">r{i}\n" + "\n".join(r[j:j + 60] for j in range(0, 3000, 60)) + "\n")
    path = tmp_path / "live.fasta"
    path.write_text("".join(recs))
    ref, _ = oracle.run_ref(str(path), k, mode=mode, slots=300000, min_abundance=2)
    assert oracle.count(path.read_bytes(), k).text(2, mode) == ref


def test_binding_text_formatter_matches_oracle_writer(oracle):
    """kaarme_gpu.keys_to_text (used by tests and the file-level mirrors) == the oracle's restatement of the
    reference writer, for every key width"""
    import importlib
    kg = importlib.import_module("canonical-k-mer-hash-table_b200")
    data = _read("g1_multiline.fasta")
    for k in (5, 21, 32, 33, 51, 64, 127, 255):
        c = oracle.count(data, k)
        assert kg.keys_to_text(c.keys, c.counts, k) == c.text(1)


def test_kaarme_file_fixture_decodes_to_the_reference_output(oracle):
    """tests/golden/g2_reads_k21.kaarme was written on a B200 by `kaarme g2_reads.fa 21 -s 200000 -a 2 --dump-kaarme`
    (profiles/io_rows_probe.sh).  The oracle's restatement of reconstruct_kmer_in_slot (kmer_hash_table.cpp:3848-4058)
    walks every chain of that file on the CPU; the lines it yields are the reference binary's own -m 2 output
    (golden.json).  Pins the file layout (header + kmer.hpp:108-123 slot words + roots) across rounds."""
    import hashlib
    import struct
    blob = _read("g2_reads_k21.kaarme")
    magic, ver, k, W, flags, n_kmers, n_roots = struct.unpack_from("<8sIIIIQQ", blob, 0)
    assert (magic, ver, k, W, flags) == (b"KAARMEG1", 1, 21, 1, 0)
    assert len(blob) == 64 + 8 * n_kmers + 8 * W * n_roots
    slots = np.frombuffer(blob, np.uint64, n_kmers, 64)
    roots = np.frombuffer(blob, np.uint64, n_roots * W, 64 + 8 * n_kmers)
    lines = []
    for i in range(n_kmers):
        d = int(slots[i])
        assert d & 1                                         # dense: every slot is occupied
        count = (d >> 12) & 16383
        hops, codes = oracle.kaarme_decode(slots, roots, k, i)
        assert hops >= 0
        kmer = "".join("ACGT"[x] for x in codes)
        assert (int(codes[0]), int(codes[-1])) == ((d >> 10) & 3, (d >> 8) & 3)      # stored end characters
        if count >= 2:
            lines.append(f"{kmer} {count}\n".encode())
    case = [c for c in json.load(open(os.path.join(GOLDEN, "golden.json")))
            if c["input"] == "g2_reads.fa" and c["k"] == 21 and c["mode"] == 2 and c["a"] == 2 and c["unique"] is None][0]
    assert len(lines) == case["n_lines"]
    assert hashlib.sha256(b"".join(sorted(lines))).hexdigest() == case["sha256"]
    # every k-mer of the input is in the structure exactly once
    want = oracle.count(_read("g2_reads.fa"), k)
    assert n_kmers == want.n
