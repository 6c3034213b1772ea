"""GPU tests of the rows SURVEY.md section 8f adds after the count path:
  f1  GPU-side text dump (kg_export_text / kg_format_text)  == the reference writer's lines, bit for bit
  f3  Kaarme structure as an interchange format (kg_kaarme_download -> file -> kg_kaarme_upload -> GPU decode)
Everything goes through the C ABI or the CLI; comparisons are exact."""
import hashlib
import importlib
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

kg = importlib.import_module("canonical-k-mer-hash-table_b200")
K = kg.kaarme_gpu
EXE = os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "kaarme")
CASES = json.load(open(os.path.join(GOLDEN, "golden.json")))


def _read(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def _sorted_lines(text):
    return sorted(text.splitlines(keepends=True))


def _counted(data, k, table_mode=K.TABLE_PLAIN, slots=400000, input_mode=K.INPUT_FASTA):
    c = kg.Counter(k=k, table_mode=table_mode, input_mode=input_mode, min_slots=slots)
    c.run_pass(K.PASS_COUNT, data)
    if table_mode == K.TABLE_KAARME:
        c.compact()
    return c


@pytest.mark.parametrize("k", [1, 5, 16, 31, 32, 33, 51, 64, 65, 96, 127, 128, 255, 256])
def test_text_dump_equals_record_export_all_widths(k, oracle):
    """every key width: GPU-formatted lines == lines formatted from the exported records == the oracle's writer"""
    data = _read("g1_multiline.fasta")
    with _counted(data, k) as c:
        keys, counts = c.export(1, K.COUNT_EXACT)
        text, n = c.export_text(1, K.COUNT_EXACT)
    assert n == len(counts)
    assert len(text) == sum(k + 2 + len(str(int(x))) for x in counts)
    assert _sorted_lines(text) == _sorted_lines(kg.keys_to_text(keys, counts, k))
    assert b"".join(_sorted_lines(text)) == oracle.count(data, k).text(1)


@pytest.mark.parametrize("table_mode,count_mode", [(K.TABLE_PLAIN, K.COUNT_EXACT), (K.TABLE_PLAIN, K.COUNT_REFERENCE),
                                                   (K.TABLE_KAARME, K.COUNT_REFERENCE)])
def test_text_dump_large_counts_and_thresholds(table_mode, count_mode, oracle):
    """poly-A input: counts with many digits, the reference's 16-bit wrap / 14-bit saturation, -a thresholds"""
    data = _read("g4_polya.fasta")
    want = oracle.count(data, 21)
    omode = oracle.TABLE_EXACT if count_mode == K.COUNT_EXACT else (oracle.TABLE_PLAIN if table_mode == K.TABLE_PLAIN else oracle.TABLE_KAARME)
    with _counted(data, 21, table_mode, slots=200000) as c:
        for a in (0, 1, 2, 3, 1000):
            text, n = c.export_text(a, count_mode)
            assert b"".join(_sorted_lines(text)) == want.text(a, omode)
            assert n == text.count(b"\n")


def test_text_dump_many_blocks_and_chunks():
    """> 1 M records: many blocks reserving byte ranges with one atomic each, several text buffers, ragged 16-byte
    heads and tails everywhere -- nothing lost, nothing duplicated, nothing torn"""
    rng = np.random.default_rng(5)
    g = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 500_000)].tobytes()
    data = b">g\n" + g + b"\n>again\n" + g[:100_000] + b"\n"
    k = 51
    with _counted(data, k, slots=1_500_000) as c:
        keys, counts = c.export(1, K.COUNT_EXACT)
        text, n = c.export_text(1, K.COUNT_EXACT)
        text2, n2 = c.export_text(2, K.COUNT_EXACT)
    assert n == len(counts) > 490_000
    lines = text.splitlines(keepends=True)
    assert len(lines) == n and all(len(l) == k + 3 for l in lines)
    want = kg.keys_to_text(keys, counts, k)
    assert hashlib.sha256(b"".join(sorted(lines))).hexdigest() == hashlib.sha256(b"".join(_sorted_lines(want))).hexdigest()
    sel = counts >= 2
    assert n2 == int(sel.sum()) and _sorted_lines(text2) == _sorted_lines(kg.keys_to_text(keys[sel], counts[sel], k))


def test_text_dump_empty_table():
    with _counted(b">x\nACG\n", 21) as c:
        assert c.export_text(1) == (b"", 0)


@pytest.mark.parametrize("k", [21, 51, 64, 127, 255])
def test_kaarme_upload_roundtrip(k, oracle):
    """download the compact structure, load it into a FRESH context, decode there: same k-mers, same counts; and the
    oracle's restatement of reconstruct_kmer_in_slot decodes the very same words"""
    data = _read("g1_multiline.fasta")
    with _counted(data, k, K.TABLE_KAARME, slots=20000) as c:
        slots, roots = c.kaarme_download()
        keys, counts = c.export(1, K.COUNT_REFERENCE)
    with kg.Counter(k=k, table_mode=K.TABLE_KAARME, min_slots=1, partitions=1, batch_bytes=1 << 20) as d:
        d.kaarme_upload(slots, roots)
        keys2, counts2 = d.export(1, K.COUNT_EXACT)
        text, n = d.export_text(2, K.COUNT_EXACT)
    assert (keys == keys2).all() and (counts == counts2).all()
    assert b"".join(_sorted_lines(text)) == oracle.count(data, k).text(2, oracle.TABLE_KAARME)
    known = set(oracle.key_strings(keys, k))
    for s in np.random.default_rng(k).integers(0, len(slots), 40):
        hops, codes = oracle.kaarme_decode(slots, roots, k, int(s))
        assert hops >= 0 and "".join("ACGT"[x] for x in codes) in known


def test_kaarme_upload_rejects_malformed_structure():
    """a pointer out of range / a cycle must surface as an error from the export, never as a wild read"""
    k = 21
    bad = np.array([(5 << 26) | (3 << 12) | 0b11, (0 << 26) | (3 << 12) | 0b11], dtype=np.uint64)   # slot 0 -> slot 5 (missing)
    cyc = np.array([(1 << 26) | (3 << 12) | 0b11, (0 << 26) | (3 << 12) | 0b11], dtype=np.uint64)   # 0 -> 1 -> 0
    for slots in (bad, cyc):
        with kg.Counter(k=k, table_mode=K.TABLE_KAARME, min_slots=1, partitions=1, batch_bytes=1 << 20) as d:
            d.kaarme_upload(slots, np.zeros(1, np.uint64))
            with pytest.raises(kg.KaarmeError):
                d.export(1)
    with kg.Counter(k=k, table_mode=K.TABLE_PLAIN, min_slots=1000) as p:
        with pytest.raises(kg.KaarmeError):
            p.kaarme_upload(bad, np.zeros(1, np.uint64))


def run_cli(args, cwd=None):
    return subprocess.run([EXE] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=cwd, text=True)


def sorted_sha(path):
    with open(path, "rb") as f:
        lines = sorted(f.read().splitlines(keepends=True))
    return len(lines), hashlib.sha256(b"".join(lines)).hexdigest()


@pytest.mark.parametrize("case", [c for c in CASES if c["k"] in (21, 127) and c["a"] == 2 and c["unique"] is None
                                  and c["input"] in ("g1_multiline.fasta", "g2_reads.fa")],
                         ids=lambda c: f"{c['input']}-k{c['k']}-m{c['mode']}")
def test_cli_host_format_equals_gpu_format(case, tmp_path):
    """--host-format (records + host threads) and the default (GPU text dump) write the same lines = the reference's"""
    outs = []
    for extra in ([], ["--host-format"]):
        out = tmp_path / f"out{len(outs)}.txt"
        p = run_cli([os.path.join(GOLDEN, case["input"]), case["k"], "-m", case["mode"], "-a", 2, "-t", 4, "-s", case["slots"], "-o", out] + extra)
        assert p.returncode == 0, p.stderr
        outs.append(sorted_sha(out))
    assert outs[0] == outs[1] == (case["n_lines"], case["sha256"])


@pytest.mark.parametrize("k", [21, 51, 127])
def test_cli_dump_and_decode_kaarme_file(k, tmp_path, oracle):
    """-m 2 --dump-kaarme writes the structure; `kaarme --from-kaarme` decodes the file on the GPU to the same output;
    the oracle decodes the file on the CPU to the same k-mers"""
    case = [c for c in CASES if c["input"] == "g2_reads.fa" and c["k"] == k and c["mode"] == 2 and c["a"] == 2 and c["unique"] is None][0]
    out, dump, out2 = tmp_path / "out.txt", tmp_path / "g2.kaarme", tmp_path / "out2.txt"
    p = run_cli([os.path.join(GOLDEN, "g2_reads.fa"), k, "-a", 2, "-t", 4, "-s", case["slots"], "-o", out, "--dump-kaarme", dump])
    assert p.returncode == 0, p.stderr
    assert sorted_sha(out) == (case["n_lines"], case["sha256"])
    p = run_cli([dump, k, "--from-kaarme", "-a", 2, "-o", out2])
    assert p.returncode == 0, p.stderr + p.stdout
    assert sorted_sha(out2) == (case["n_lines"], case["sha256"])
    assert f"Written k-mers: {case['n_lines']}" in p.stdout
    # wrong KLEN, truncated file
    assert run_cli([dump, k + 1, "--from-kaarme", "-o", out2]).returncode == 1
    blob = dump.read_bytes()
    (tmp_path / "cut.kaarme").write_bytes(blob[:-8])
    p = run_cli([tmp_path / "cut.kaarme", k, "--from-kaarme", "-o", out2])
    assert p.returncode == 1 and "ill-formed" in p.stderr
    # the oracle reads the same file
    magic, ver, kk, W, _flags, n_kmers, n_roots = struct.unpack_from("<8sIIIIQQ", blob, 0)
    assert (magic, ver, kk, W) == (b"KAARMEG1", 1, k, (k + 31) // 32)
    slots = np.frombuffer(blob, np.uint64, n_kmers, 64)
    roots = np.frombuffer(blob, np.uint64, n_roots * W, 64 + 8 * n_kmers)
    want = set(l.split(b" ")[0].decode() for l in out.read_bytes().splitlines())
    sel = [i for i in range(n_kmers) if ((int(slots[i]) >> 12) & 16383) >= 2]
    assert len(sel) == case["n_lines"]
    for i in sel[:200]:
        hops, codes = oracle.kaarme_decode(slots, roots, k, i)
        assert hops >= 0
        assert "".join("ACGT"[x] for x in codes) in want
