"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C ABI vs the oracle / golden vectors.
Bit-exact: same canonical k-mers, same counts, same -a threshold (integer work, no tolerance)."""
import hashlib
import importlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

kg = importlib.import_module("canonical-k-mer-hash-table_b200")
K = kg.kaarme_gpu

CASES = json.load(open(os.path.join(GOLDEN, "golden.json")))
COMP = str.maketrans("ACGT", "TGCA")


def _read(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def _imode(name):
    return K.INPUT_PLAIN if name.endswith(".txt") else K.INPUT_FASTA


def gpu_count(data, k, input_mode=K.INPUT_FASTA, table_mode=K.TABLE_PLAIN, slots=400000, a=1,
              count_mode=K.COUNT_EXACT, batch_bytes=0, feeder=None, starts_in_header=False, partitions=0):
    with kg.Counter(k=k, table_mode=table_mode, input_mode=input_mode, min_slots=slots, batch_bytes=batch_bytes,
                    partitions=partitions) as c:
        c.pass_begin(K.PASS_COUNT)
        c.stream_begin(starts_in_header)
        if feeder:
            feeder(c, data)
        else:
            c.feed(data)
        st = c.pass_end()
        keys, counts = c.export(a, count_mode)
    return keys, counts, st


def assert_same(keys, counts, want):
    assert keys.shape == want.keys.shape, (keys.shape, want.keys.shape)
    assert (keys == want.keys).all()
    assert (counts.astype(np.uint64) == want.counts).all()


def make_fasta(rng, genome_len, nreads, read_len, wrap=0, err=0.0, n_rate=0.0):
    g = rng.integers(0, 4, genome_len).astype(np.uint8)
    out = []
    for i in range(nreads):
        p = int(rng.integers(0, genome_len - read_len + 1))
        r = g[p:p + read_len].copy()
        if err:
            e = rng.random(read_len) < err
            r[e] = (r[e] + rng.integers(1, 4, int(e.sum()))) & 3
        s = np.frombuffer(b"ACGT", np.uint8)[r].tobytes().decode()
        if rng.random() < 0.5:
            s = s[::-1].translate(COMP)
        if n_rate:
            s = "".join("N" if rng.random() < n_rate else ch for ch in s)
        if rng.random() < 0.3:
            s = s.lower()
        body = "\n".join(s[j:j + wrap] for j in range(0, len(s), wrap)) if wrap else s
        out.append(f">read{i} len={read_len}\n{body}\n")
    return "".join(out).encode()


@pytest.mark.parametrize("case", [c for c in CASES if c["unique"] is None and c["mode"] == 0],
                         ids=lambda c: f"{c['input']}-k{c['k']}-a{c['a']}")
def test_golden_plain(case):
    """-m 0 against the reference binary's sorted output (tests/golden/golden.json)."""
    data = _read(case["input"])
    keys, counts, st = gpu_count(data, case["k"], _imode(case["input"]), K.TABLE_PLAIN, case["slots"], case["a"],
                                 K.COUNT_REFERENCE)
    txt = kg.keys_to_text(keys, counts, case["k"])
    assert txt.count(b"\n") == case["n_lines"]
    assert hashlib.sha256(txt).hexdigest() == case["sha256"]
    assert st["table_slots"] == case["table_slots"]


@pytest.mark.parametrize("k", [1, 5, 21, 31, 32, 33, 51, 63, 64, 65, 96, 97, 127, 128, 129, 160, 200, 255, 256])
def test_k_sweep_vs_oracle(oracle, k):
    rng = np.random.default_rng(1000 + k)
    data = make_fasta(rng, 20000, 60, 900, wrap=70, err=0.01, n_rate=0.001)
    want = oracle.count(data, k)
    keys, counts, st = gpu_count(data, k)
    assert st["input_kmers"] == want.total_windows
    assert st["distinct"] == want.n
    assert_same(keys, counts, want)


@pytest.mark.parametrize("batch", [4096, 8192, 65536])
@pytest.mark.parametrize("k", [31, 51, 127])
def test_batch_boundaries_invisible(oracle, k, batch):
    """windows, headers and multi-line records crossing device batch boundaries (carry of k-1 bases)"""
    rng = np.random.default_rng(k + batch)
    data = make_fasta(rng, 50000, 120, 1500, wrap=80, err=0.005, n_rate=0.0005)
    want = oracle.count(data, k)
    keys, counts, st = gpu_count(data, k, batch_bytes=batch)
    assert st["input_kmers"] == want.total_windows
    assert_same(keys, counts, want)


def test_ragged_feeds(oracle):
    """the byte stream may be cut anywhere: 1-byte feeds, cuts inside headers, inside windows"""
    rng = np.random.default_rng(77)
    data = make_fasta(rng, 5000, 30, 400, wrap=60, n_rate=0.002)
    want = oracle.count(data, 31)

    def feeder(c, d):
        pos = 0
        while pos < len(d):
            n = int(rng.integers(1, 700))
            c.feed(d[pos:pos + n])
            pos += n

    keys, counts, st = gpu_count(data, 31, feeder=feeder, batch_bytes=4096)
    assert st["input_kmers"] == want.total_windows
    assert_same(keys, counts, want)


def test_context_feed(oracle):
    """KG_FEED_CONTEXT bytes warm the window but their k-mers are not counted (slice overlap of text_reader.h)"""
    rng = np.random.default_rng(5)
    data = make_fasta(rng, 8000, 1, 6000, wrap=70)           # one multi-line record
    cut = 3001
    k = 51
    whole = oracle.count(data, k)
    first = oracle.count(data[:cut], k)

    def feeder(c, d):
        c.feed(d[:cut], K.FEED_CONTEXT)
        c.feed(d[cut:])

    keys, counts, st = gpu_count(data, k, feeder=feeder)
    assert st["input_kmers"] == whole.total_windows - first.total_windows
    # second-half counts = whole - first
    tot = dict(zip(map(bytes, whole.keys.view(np.uint8).reshape(whole.n, -1)), whole.counts))
    for kk, cc in zip(map(bytes, first.keys.view(np.uint8).reshape(first.n, -1)), first.counts):
        tot[kk] -= cc
    want = {kk: cc for kk, cc in tot.items() if cc}
    got = dict(zip(map(bytes, keys.view(np.uint8).reshape(len(counts), -1)), counts.astype(np.uint64)))
    assert got == want


def test_starts_in_header(oracle):
    data = b"tail of a broken header ACGTACGTACGTACGTACGTACGTACGT\nACGTTGCAAGGCTTAACCGGTACGT\n>r2\nGGGGGCCCCCAAAAATTTTTACGTAC\n"
    want = oracle.count(data, 21, oracle.FASTA, starts_in_header=True)
    keys, counts, st = gpu_count(data, 21, starts_in_header=True)
    assert st["input_kmers"] == want.total_windows
    assert_same(keys, counts, want)


def test_plain_mode(oracle):
    rng = np.random.default_rng(9)
    g = "".join("ACGT"[x] for x in rng.integers(0, 4, 4000))
    lines = [g[i:i + int(rng.integers(10, 300))] for i in range(0, 3600, 150)]
    lines[3] = lines[3][:40] + "N" + lines[3][40:]
    lines[5] = lines[5].lower()
    data = ("\n".join(lines) + "\n").encode()
    for k in (21, 51):
        want = oracle.count(data, k, oracle.PLAIN)
        keys, counts, st = gpu_count(data, k, K.INPUT_PLAIN)
        assert st["input_kmers"] == want.total_windows
        assert_same(keys, counts, want)


@pytest.mark.parametrize("data", [b"", b">only a header", b">h\nACGT\n", b">h\n\n\n\n", b">a\n>b\n>c\n", b"\n\n\n",
                                  b">h\n" + b"N" * 100 + b"\n"])
def test_empty_and_degenerate(oracle, data):
    want = oracle.count(data, 21)
    keys, counts, st = gpu_count(data, 21)
    assert st["input_kmers"] == want.total_windows == 0
    assert len(counts) == 0


def test_count_width_emulation(oracle):
    """uint16 wrap (-m 0) and 14-bit saturation (-m 2) of the reported counts; golden from the reference"""
    data = _read("g4_polya.fasta")
    want = oracle.count(data, 21)
    keys, counts, _ = gpu_count(data, 21, a=2, count_mode=K.COUNT_REFERENCE)
    assert_same(keys, counts, want.filtered(2, oracle.TABLE_PLAIN))
    keys, counts, _ = gpu_count(data, 21, a=2, count_mode=K.COUNT_EXACT)
    assert_same(keys, counts, want.filtered(2, oracle.TABLE_EXACT))
    assert int(counts.max()) == 89960


def test_table_full_is_an_error():
    rng = np.random.default_rng(3)
    data = make_fasta(rng, 20000, 20, 1000)
    with kg.Counter(k=31, min_slots=1000) as c:
        c.pass_begin(K.PASS_COUNT)
        c.stream_begin()
        c.feed(data)
        with pytest.raises(kg.TableFull):
            c.pass_end()


def test_host_pinned_and_device_feeds_agree(oracle):
    import torch
    rng = np.random.default_rng(11)
    data = make_fasta(rng, 30000, 100, 1000, wrap=80, err=0.01)
    want = oracle.count(data, 51)
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8)
    for src in (data, t.pin_memory(), t.cuda()):
        keys, counts, st = gpu_count(src, 51, batch_bytes=16384)
        assert st["input_kmers"] == want.total_windows
        assert_same(keys, counts, want)


def test_min_abundance_zero_writes_nothing():
    data = b">r\nACGTACGTACGTACGTACGTACGTACGTACGT\n"
    keys, counts, _ = gpu_count(data, 21, a=0)
    assert len(counts) == 0


def test_medium_config3_shape(oracle):
    """BASELINE config 3 shape at 1/50 scale (5 Mbp x 1x coverage of 150 bp reads, 1 % error, k=31): exact parity,
    plus the size-independent checks used at full size: sum(counts) == input k-mers, sorted unique keys."""
    rng = np.random.default_rng(42)
    data = make_fasta(rng, 5_000_000, 33_000, 150, err=0.01)
    want = oracle.count(data, 31)
    keys, counts, st = gpu_count(data, 31, slots=16_000_000)
    assert st["input_kmers"] == want.total_windows
    assert int(counts.astype(np.uint64).sum()) == st["input_kmers"]
    assert_same(keys, counts, want)


# ---- double Bloom filter (-b): the north-star rule ------------------------------------------------------------
def gpu_bloom(data, k, unique, fpr, input_mode=K.INPUT_FASTA, table_mode=K.TABLE_PLAIN, a=1, batch_bytes=0):
    with kg.Counter(k=k, table_mode=table_mode, input_mode=input_mode, use_bloom=True, expected_unique=unique,
                    fpr=fpr, batch_bytes=batch_bytes) as c:
        b = c.run_pass(K.PASS_BLOOM, data)
        st = c.run_pass(K.PASS_COUNT, data)
        keys, counts = c.export(a, K.COUNT_EXACT)
    return keys, counts, b, st


@pytest.mark.parametrize("case", [c for c in CASES if c["unique"] is not None and c["mode"] == 0 and c["a"] == 2],
                         ids=lambda c: f"{c['input']}-k{c['k']}-u{c['unique']}-f{c['fpr']}")
def test_bloom_golden(oracle, case):
    """-m 0 -b: at -a >= 2 the output equals the reference's (== the no-Bloom ground truth): no false negatives.
    False positives (singletons admitted) are reported, not hidden."""
    data = _read(case["input"])
    keys, counts, b, st = gpu_bloom(data, case["k"], case["unique"], case["fpr"], _imode(case["input"]), a=2)
    txt = kg.keys_to_text(keys, counts, case["k"])
    assert txt.count(b"\n") == case["n_lines"]
    assert hashlib.sha256(txt).hexdigest() == case["sha256"]
    m, nh, _ = oracle.bloom_params(case["unique"], case["fpr"])
    assert (b["bloom_bits"], b["bloom_hashes"]) == (m, nh)
    assert st["table_slots"] == oracle.lib().ko_next_prime3mod4(2 * b["new_in_second"])   # main.cpp:454
    # the counters track the reference's (different hash functions => not identical)
    # (racing occurrences are counted once per racer: a small over-count, never the reference's under-count)
    assert -0.01 * case["new_in_second"] - 5 <= b["new_in_second"] - case["new_in_second"] <= 0.06 * case["new_in_second"] + 5
    assert abs(b["new_in_first"] - case["new_in_first"]) <= 0.06 * case["new_in_first"] + 5


@pytest.mark.parametrize("k,fpr", [(21, 0.01), (51, 0.01), (31, 0.2), (127, 0.05)])
def test_bloom_fp_fn_vs_truth(oracle, k, fpr):
    rng = np.random.default_rng(k)
    data = make_fasta(rng, 60000, 400, 600, wrap=70, err=0.02)
    truth = oracle.count(data, k)
    n_single = int((truth.counts == 1).sum())
    keys, counts, b, st = gpu_bloom(data, k, unique=truth.n, fpr=fpr, a=1, batch_bytes=65536)
    got = dict(zip(map(bytes, keys.view(np.uint8).reshape(len(counts), -1)), counts.astype(np.uint64)))
    want = dict(zip(map(bytes, truth.keys.view(np.uint8).reshape(truth.n, -1)), truth.counts))
    fn = [kk for kk, cc in want.items() if cc >= 2 and kk not in got]
    assert not fn, f"{len(fn)} false negatives"
    for kk, cc in got.items():
        assert want[kk] == cc                      # admitted k-mers carry their exact count
    fp = sum(1 for kk, cc in got.items() if cc == 1)
    assert fp <= max(10, 3.0 * fpr * n_single), (fp, n_single)   # blocked filter: allow 3x the nominal rate
    assert st["input_kmers"] == truth.total_windows
    print(f"k={k} fpr={fpr}: singletons={n_single} false_positives={fp} ({fp / max(1, n_single):.4f}) false_negatives=0")


def test_bloom_kmer_repeated_only_inside_one_warp(oracle):
    """two occurrences processed concurrently by neighbouring threads must still reach filter 2 (race branch,
    double_bloomfilter.hpp:401-411)"""
    rng = np.random.default_rng(2)
    unit = "".join("ACGT"[x] for x in rng.integers(0, 4, 64))
    data = (">r\n" + unit * 2 + "\n").encode()      # every 31-mer of the unit occurs exactly twice, 64 bases apart
    truth = oracle.count(data, 31).filtered(2)
    keys, counts, b, st = gpu_bloom(data, 31, unique=1000, fpr=0.01, a=2)
    assert_same(keys, counts, truth)


# ---- bucketed path on one GPU (hist -> scan -> scatter -> insert), the same kernels the multi-GPU exchange uses --
@pytest.mark.parametrize("partitions", [2, 8, 32, 256, 1000])
@pytest.mark.parametrize("k", [31, 51, 255])
def test_partitioned_insert(oracle, k, partitions):
    rng = np.random.default_rng(k * 100 + partitions)
    data = make_fasta(rng, 40000, 150, 1200, wrap=80, err=0.01, n_rate=0.0005)
    want = oracle.count(data, k)
    keys, counts, st = gpu_count(data, k, batch_bytes=32768, partitions=partitions)
    assert st["input_kmers"] == want.total_windows
    assert st["inserted_kmers"] == want.total_windows
    assert_same(keys, counts, want)


def test_partitioned_bloom(oracle):
    rng = np.random.default_rng(8)
    data = make_fasta(rng, 60000, 400, 600, wrap=70, err=0.02)
    truth = oracle.count(data, 51)
    with kg.Counter(k=51, use_bloom=True, expected_unique=truth.n, fpr=0.01, partitions=16, batch_bytes=65536) as c:
        c.run_pass(K.PASS_BLOOM, data)
        c.run_pass(K.PASS_COUNT, data)
        keys, counts = c.export(2, K.COUNT_EXACT)
    assert_same(keys, counts, truth.filtered(2))


# ---- Kaarme (-m 2): compaction of the counted table into 8-byte slots + roots, decoded on export -----------------
def gpu_kaarme(data, k, input_mode=K.INPUT_FASTA, slots=400000, a=1, bloom=None, batch_bytes=0, download=False):
    kw = dict(use_bloom=True, expected_unique=bloom[0], fpr=bloom[1]) if bloom else dict(min_slots=slots)
    with kg.Counter(k=k, table_mode=K.TABLE_KAARME, input_mode=input_mode, batch_bytes=batch_bytes, **kw) as c:
        if bloom:
            c.run_pass(K.PASS_BLOOM, data)
        st = c.run_pass(K.PASS_COUNT, data)
        cs = c.compact()
        arrays = c.kaarme_download() if download else None
        keys, counts = c.export(a, K.COUNT_REFERENCE)
    return keys, counts, st, cs, arrays


@pytest.mark.parametrize("case", [c for c in CASES if c["mode"] == 2 and c["k"] in (21, 33, 51, 127, 255)],
                         ids=lambda c: f"{c['input']}-k{c['k']}-a{c['a']}-{'b' if c['unique'] else 's'}")
def test_golden_kaarme(case):
    """-m 2 against the reference binary's sorted output; with -b only -a 2 is hash-independent"""
    if case["unique"] and case["a"] < 2:
        pytest.skip("-a 1 output of the Bloom mode depends on the filter's false positives")
    data = _read(case["input"])
    keys, counts, st, cs, _ = gpu_kaarme(data, case["k"], _imode(case["input"]), case["slots"] or 0, case["a"],
                                         bloom=(case["unique"], case["fpr"]) if case["unique"] else None)
    txt = kg.keys_to_text(keys, counts, case["k"])
    assert txt.count(b"\n") == case["n_lines"]
    assert hashlib.sha256(txt).hexdigest() == case["sha256"]
    assert cs["kmers"] == st["distinct"] and cs["bytes"] == 8 * cs["kmers"] + 8 * ((case["k"] + 31) // 32) * cs["roots"]


@pytest.mark.parametrize("k", [5, 21, 31, 32, 33, 51, 64, 65, 96, 127, 128, 200, 255, 256])
def test_kaarme_vs_oracle(oracle, k):
    """every k, including k = 0 mod 32 where the reference's own -m 2 is broken"""
    rng = np.random.default_rng(7000 + k)
    data = make_fasta(rng, 30000, 80, 1500, wrap=70, err=0.01, n_rate=0.0005)
    want = oracle.count(data, k).filtered(1, oracle.TABLE_KAARME)
    keys, counts, st, cs, _ = gpu_kaarme(data, k, batch_bytes=16384)
    assert_same(keys, counts, want)
    assert cs["kmers"] == want.n and 0 < cs["roots"] < want.n
    print(f"k={k}: {cs['kmers']} k-mers, {cs['roots']} roots, {cs['bytes'] / cs['kmers']:.2f} B/k-mer, max chain {cs['max_chain']}")


def test_kaarme_structure_decodes_with_reference_algorithm(oracle):
    """the downloaded slots/roots, decoded by the oracle's restatement of reconstruct_kmer_in_slot
    (kmer_hash_table.cpp:3848-4058), give exactly the exported k-mers: the structure is bit-compatible"""
    rng = np.random.default_rng(99)
    k = 51
    data = make_fasta(rng, 20000, 60, 1200, wrap=80, err=0.01)
    keys, counts, st, cs, (slots, roots) = gpu_kaarme(data, k, download=True)
    assert len(slots) == cs["kmers"] and roots.shape[0] == cs["roots"]
    got = set()
    for i in range(len(slots)):
        hops, codes = oracle.kaarme_decode(slots, roots, k, i)
        assert hops >= 0
        got.add(("".join("ACGT"[c] for c in codes), int((slots[i] >> np.uint64(12)) & np.uint64(16383))))
    want = set(zip(oracle.key_strings(keys, k), (int(c) for c in counts)))
    assert got == want
    assert cs["bytes"] < 10 * cs["kmers"]          # ~8 B per k-mer + a few roots


def test_kaarme_after_bloom(oracle):
    rng = np.random.default_rng(31)
    data = make_fasta(rng, 60000, 300, 800, wrap=70, err=0.02)
    truth = oracle.count(data, 31)
    keys, counts, st, cs, _ = gpu_kaarme(data, 31, a=2, bloom=(truth.n, 0.01))
    assert_same(keys, counts, truth.filtered(2, oracle.TABLE_KAARME))


def test_kaarme_count_saturation(oracle):
    data = _read("g4_polya.fasta")
    keys, counts, st, cs, _ = gpu_kaarme(data, 21, a=2)
    assert_same(keys, counts, oracle.count(data, 21).filtered(2, oracle.TABLE_KAARME))
    assert int(counts.max()) == 16383


def test_partition_overflow_falls_back_to_direct_insert(oracle):
    """pathological skew: one k-mer repeated 10^5 times overflows its fixed-capacity bucket region of the one-pass
    scatter; the overflowing keys are inserted directly and nothing is lost"""
    data = _read("g4_polya.fasta")
    want = oracle.count(data, 21)
    for parts in (8, 64):
        keys, counts, st = gpu_count(data, 21, batch_bytes=65536, partitions=parts)
        assert st["input_kmers"] == want.total_windows == st["inserted_kmers"]
        assert_same(keys, counts, want)
    want51 = oracle.count(data, 51)
    keys, counts, st = gpu_count(data, 51, batch_bytes=65536, partitions=16)
    assert_same(keys, counts, want51)


# ---- file-level mirrors of the reference functors (parallel_parser.hpp:229,1181,1577,2255) ---------------------------
@pytest.mark.parametrize("k,mode", [(127, K.TABLE_PLAIN), (200, K.TABLE_PLAIN), (51, K.TABLE_KAARME), (97, K.TABLE_KAARME)])
@pytest.mark.parametrize("partitions", [1, 8])
def test_contended_generic_slots(oracle, k, mode, partitions):
    """a handful of distinct multi-word k-mers hammered by every block at once (the same short read, 60 000 times): the
    generic slot protocol (CAS-claim, key words, release-publish; acquire on the reader's side; parked probes on the
    bucketed path) must neither split a k-mer over two slots nor lose an occurrence -- distinct and counts == oracle"""
    rng = np.random.default_rng(k)
    read = "".join("ACGT"[x] for x in rng.integers(0, 4, k + 40))
    data = b"".join(f">r{i}\n{read}\n".encode() for i in range(4)) * 15000
    want = oracle.count(data, k)
    assert want.n <= 41 and want.counts.max() == 60000
    keys, counts, st = gpu_count(data, k, table_mode=mode, slots=5000, partitions=partitions, batch_bytes=1 << 20)
    assert st["distinct"] == want.n
    assert_same(keys, counts, want)


@pytest.mark.parametrize("fn,mode,bloom", [("parse_input_atomic_flag", 0, False), ("parse_input_pointer_atomic_variable", 2, False),
                                           ("parse_input_atomic_flag_BF", 0, True), ("parse_input_pointer_atomic_variable_BF", 2, True)])
def test_functor_mirrors(tmp_path, fn, mode, bloom):
    # (the reference's sorted -m 0 and -m 2 outputs are identical while counts stay below 16 384, SURVEY section 8;
    #  the golden set holds the no-Bloom g5 run for -m 0 only)
    case = [c for c in CASES if c["input"] == "g5_long.fasta" and c["k"] == 51 and c["a"] == 2
            and bool(c["unique"]) == bloom and (c["mode"] == mode or not bloom)][0]
    out = tmp_path / "out.txt"
    f = getattr(kg, fn)
    src = os.path.join(GOLDEN, "g5_long.fasta")
    if bloom:
        f(src, str(out), 51, case["unique"], case["fpr"], 2)
    else:
        f(src, str(out), 51, case["slots"], 2)
    lines = sorted(out.read_bytes().splitlines(keepends=True))
    assert len(lines) == case["n_lines"]
    assert hashlib.sha256(b"".join(lines)).hexdigest() == case["sha256"]
