"""CPU checks of host-testable native code on the product path:
  * csrc/kg_text.cuh   kg_format_line / kg_ndigits -- the per-record formatter the GPU text dump runs per thread
                       (__host__ __device__, so the same code runs here) vs the oracle's restatement of the reference
                       writer (kmer_hash_table.cpp:2022-2043) and the binding's keys_to_text;
  * host/kg_writer.hpp ParallelWriter -- the CLI's output side: whole text buffers appended with concurrent pwrite(2)s
                       (replaces the ofstream loop of kmer_hash_table.cpp:2013-2050) vs the bytes handed to it;
  * host/kg_reader.hpp SliceReader -- the CLI's threaded reader ring (replaces text_reader.h:91-226 + io_worker,
                       parallel_parser.hpp:275-338) vs the bytes of the file.
The harnesses live in tests/native/ and are compiled here (nvcc host compilation / g++); no GPU is touched."""
import importlib
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

NATIVE = os.path.join(ROOT, "tests", "native")
BUILD = os.path.join(NATIVE, "_build")


def _build(name, cmd):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    srcs = [a for a in cmd if a.endswith((".cu", ".cpp"))]
    deps = srcs + [os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "csrc", "kg_text.cuh"),
                   os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "csrc", "kg_refhash.cuh"),
                   os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "csrc", "kg_skm.cuh"),
                   os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "csrc", "kg_count.cuh"),
                   os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "csrc", "kg_device.cuh"),
                   os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "host", "kg_reader.hpp"),
                   os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "host", "kg_writer.hpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(cmd + ["-o", exe], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return exe


@pytest.fixture(scope="module")
def text_exe():
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
    return _build("text_format_host", [nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets",
                                       os.path.join(NATIVE, "text_format_host.cu")])


@pytest.fixture(scope="module")
def reader_exe():
    return _build("reader_host", ["g++", "-O2", "-std=c++17", "-Wall", "-pthread", os.path.join(NATIVE, "reader_host.cpp")])


def _format(exe, keys, counts, k):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    payload = struct.pack("<II", k, len(counts)) + keys.tobytes() + counts.tobytes()
    p = subprocess.run([exe], input=payload, stdout=subprocess.PIPE, check=True)
    return p.stdout


@pytest.mark.parametrize("k", [1, 2, 5, 16, 21, 31, 32, 33, 51, 63, 64, 65, 96, 127, 128, 129, 255, 256])
def test_format_line_random_keys_all_widths(text_exe, k):
    """every key width and every count-digit length (1..10 digits), random keys"""
    kg = importlib.import_module("canonical-k-mer-hash-table_b200")
    rng = np.random.default_rng(k)
    W = (k + 31) // 32
    n = 400
    keys = rng.integers(0, 1 << 63, size=(n, W), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, W), dtype=np.uint64)
    top = 2 * k - 64 * (W - 1)
    if top < 64:
        keys[:, 0] &= np.uint64((1 << top) - 1)
    counts = np.concatenate([
        np.array([0, 1, 9, 10, 99, 100, 999, 1000, 9999, 10000, 16383, 65535, 65536, 99999, 100000, 999999, 1000000,
                  9999999, 10000000, 99999999, 100000000, 999999999, 1000000000, 4294967295], dtype=np.uint64),
        10 ** rng.integers(0, 10, n - 24).astype(np.uint64) + rng.integers(0, 1000, n - 24).astype(np.uint64)])
    counts = np.minimum(counts, 4294967295).astype(np.uint32)
    got = _format(text_exe, keys, counts, k)
    assert got == kg.keys_to_text(keys, counts, k)
    for line in got.splitlines():
        kmer, cnt = line.split(b" ")
        assert len(kmer) == k and set(kmer) <= set(b"ACGT") and cnt == str(int(cnt)).encode()


@pytest.mark.parametrize("k", [5, 21, 32, 33, 51, 64, 127, 255])
def test_format_line_matches_oracle_writer(text_exe, oracle, k):
    """same bytes as the oracle's restatement of the reference writer on real counted data, with the reference's
    count widths applied (-m 0 wrap / -m 2 saturation)"""
    with open(os.path.join(GOLDEN, "g4_polya.fasta"), "rb") as f:
        data = f.read()
    c = oracle.count(data, k)
    for mode in (oracle.TABLE_EXACT, oracle.TABLE_PLAIN, oracle.TABLE_KAARME):
        f = c.filtered(2, mode)
        assert _format(text_exe, f.keys, f.counts.astype(np.uint32), k) == c.text(2, mode)


def _read_slice(exe, path, ctx_lo, lo, hi, buf_bytes, nbufs, io_threads):
    p = subprocess.run([exe, path] + [str(x) for x in (ctx_lo, lo, hi, buf_bytes, nbufs, io_threads)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, (p.returncode, p.stderr)
    nctx, nbody, chunks = struct.unpack("<QQQ", p.stdout[:24])
    assert len(p.stdout) == 24 + nctx + nbody
    return p.stdout[24:24 + nctx], p.stdout[24 + nctx:], chunks


def test_reader_delivers_the_slice_in_order(reader_exe, tmp_path):
    rng = np.random.default_rng(7)
    blob = rng.integers(0, 256, 9_000_037, dtype=np.uint8).tobytes()
    path = tmp_path / "blob.bin"
    path.write_bytes(blob)
    n = len(blob)
    cases = [
        (0, 0, n, 1 << 20, 3, 4),              # whole file, several chunks
        (0, 0, n, 16 << 20, 2, 8),             # one chunk larger than the file, split over 8 preads
        (1000, 1050, n - 3, 4096, 3, 1),       # context + tiny buffers, single reader thread
        (123457, 123457, 5_000_001, 777_777, 4, 3),   # no context, odd sizes
        (n - 10, n - 5, n, 1 << 20, 2, 2),     # tail of the file
        (500, 500, 500, 4096, 2, 2),           # empty range
        (0, 40, 40, 4096, 2, 2),               # context only
    ]
    for ctx_lo, lo, hi, buf_bytes, nbufs, io in cases:
        ctx, body, chunks = _read_slice(reader_exe, str(path), ctx_lo, lo, hi, buf_bytes, nbufs, io)
        assert ctx == blob[ctx_lo:lo], (ctx_lo, lo, hi)
        assert body == blob[lo:hi], (ctx_lo, lo, hi)
        want_chunks = -(-(lo - ctx_lo) // buf_bytes) + -(-(hi - lo) // buf_bytes)
        assert chunks == want_chunks


def test_reader_past_end_of_file_and_missing_file(reader_exe, tmp_path):
    path = tmp_path / "short.bin"
    path.write_bytes(b"ACGT" * 1000)
    ctx, body, _ = _read_slice(reader_exe, str(path), 0, 0, 10_000, 1 << 12, 2, 2)   # range longer than the file
    assert ctx == b"" and body == b"ACGT" * 1000
    p = subprocess.run([reader_exe, str(tmp_path / "nope"), "0", "0", "10", "4096", "2", "1"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE)
    assert p.returncode == 5 and b"cannot open" in p.stderr


def test_reader_matches_cli_slices(reader_exe):
    """the slices the CLI computes for --gpus N (make_slice) read back exactly the context and range they name"""
    exe = os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "kaarme")
    fa = os.path.join(GOLDEN, "g5_long.fasta")
    blob = open(fa, "rb").read()
    out = subprocess.run([exe, fa, "51", "-s", "1000", "--gpus", "3", "--print-slices"], stdout=subprocess.PIPE, text=True).stdout
    rows = [list(map(int, l.split()[1:])) for l in out.splitlines() if l.startswith("slice ")]
    assert len(rows) == 3
    for _, ctx_lo, lo, hi, _hdr in rows:
        ctx, body, _ = _read_slice(reader_exe, fa, ctx_lo, lo, hi, 8192, 3, 2)
        assert ctx == blob[ctx_lo:lo] and body == blob[lo:hi]


@pytest.fixture(scope="module")
def writer_exe():
    return _build("writer_host", ["g++", "-O2", "-std=c++17", "-Wall", "-pthread", os.path.join(NATIVE, "writer_host.cpp")])


def test_writer_appends_buffers_back_to_back(writer_exe, tmp_path):
    rng = np.random.default_rng(3)
    blob = rng.integers(0, 256, 70_000_003, dtype=np.uint8).tobytes()
    src = tmp_path / "src.bin"
    src.write_bytes(blob)
    for threads, chunk in [(1, 64 << 20), (8, 64 << 20), (4, 9_999_999), (16, 1 << 20), (3, 70_000_003), (8, 4096)]:
        out = tmp_path / f"out_{threads}_{chunk}.bin"
        p = subprocess.run([writer_exe, str(src), str(out), str(threads), str(chunk)])
        assert p.returncode == 0
        assert out.read_bytes() == blob, (threads, chunk)
        out.unlink()
    # continues at the current offset of the descriptor
    out = tmp_path / "prefixed.bin"
    assert subprocess.run([writer_exe, str(src), str(out), "8", str(32 << 20), "HEADER\n"]).returncode == 0
    assert out.read_bytes() == b"HEADER\n" + blob
    # ... and leaves it at the end: a plain write(2) through the same descriptor lands after the data
    assert subprocess.run([writer_exe, str(src), str(out), "8", str(32 << 20), "HEADER\n", "TRAILER\n"]).returncode == 0
    assert out.read_bytes() == b"HEADER\n" + blob + b"TRAILER\n"


def test_writer_on_a_pipe_falls_back_to_sequential_writes(writer_exe, tmp_path):
    blob = np.random.default_rng(4).integers(0, 256, 20_000_001, dtype=np.uint8).tobytes()
    src = tmp_path / "src.bin"
    src.write_bytes(blob)
    p = subprocess.run([writer_exe, str(src), "-", "8", str(16 << 20)], stdout=subprocess.PIPE)
    assert p.returncode == 0 and p.stdout == blob


def test_writer_reports_errors(writer_exe, tmp_path):
    src = tmp_path / "src.bin"
    src.write_bytes(b"x" * 100)
    with open("/dev/full", "wb") as full:
        p = subprocess.run([writer_exe, str(src), "-", "2", "50"], stdout=full)
    assert p.returncode == 5


@pytest.fixture(scope="module")
def skm_exe():
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
    return _build("skm_host", [nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", os.path.join(NATIVE, "skm_host.cu")])


def _skm_input(k, nb, pl, carry, runs):
    blob = struct.pack("<IIIII", k, nb, pl, carry, len(runs))
    for r in runs:
        blob += struct.pack("<I", len(r)) + bytes(r)
    return blob


@pytest.mark.parametrize("k", [1, 5, 11, 12, 21, 26, 27, 31, 32, 33, 37, 51, 64, 65, 96, 127, 128, 200, 255, 256])
def test_skm_scatter_logic_matches_definition(skm_exe, k):
    """csrc/kg_skm.cuh on the CPU: the scatter's two phases (m-mer hashes per packed word with halo, sliding minimum,
    run segmentation) driven block by block exactly as kg_skm_scatter drives them.  Every valid window must be covered by
    exactly one descriptor, never across a packed word, with the bucket the DEFINITION gives (minimum over the window's
    canonical m-mer hashes, brute force), the right has-predecessor flag, and kg_window_at / kg_key_bucket must agree
    with keys packed straight from the bases (forward and reverse complement give the same bucket)."""
    rng = np.random.default_rng(1000 + k)
    runs = []
    for i in range(40):
        n = int(rng.integers(1, 3 * k + 700)) if i % 3 else int(rng.integers(1, k + 2))
        r = rng.integers(0, 4, n).astype(np.uint8)
        if i % 7 == 0:
            r[:] = r[0]                                            # homopolymer: every window shares one minimizer
        if i % 11 == 0 and n > 8:
            r[:] = np.resize(r[:2], n)                             # dinucleotide repeat
        runs.append(r)
    runs.append(rng.integers(0, 4, 6000).astype(np.uint8))         # crosses a 128-word block boundary
    for nb, pl, carry in ((1, 1, 0), (256, 256, 0), (1024, 128, k + 13), (64, 8, 0), (7, 7, 3)):
        p = subprocess.run([skm_exe], input=_skm_input(k, nb, pl, carry, runs), stdout=subprocess.PIPE, check=True)
        out = p.stdout.decode().strip()
        assert out.startswith("OK"), (k, nb, pl, carry, out)


def test_skm_buckets_balance_on_random_sequence(skm_exe):
    """a random genome spreads evenly over the minimizer buckets (no bucket above 2x its share at 64 buckets) and the
    descriptors are few: ~11 windows each at k = 51"""
    rng = np.random.default_rng(7)
    runs = [rng.integers(0, 4, 10000).astype(np.uint8) for _ in range(60)]
    p = subprocess.run([skm_exe], input=_skm_input(51, 64, 8, 0, runs), stdout=subprocess.PIPE, check=True)
    out = p.stdout.decode().strip()
    assert out.startswith("OK"), out
    f = dict(x.split("=") for x in out.split()[1:])
    assert float(f["max_bucket_share"]) < 2.0 / 64
    assert int(f["windows"]) / int(f["descriptors"]) > 8


@pytest.fixture(scope="module")
def refhash_exe():
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
    return _build("refhash_host", [nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", os.path.join(NATIVE, "refhash_host.cu")])


def _refhash_windows(exe, codes, k):
    payload = struct.pack("<II", k, len(codes)) + bytes(codes)
    out = subprocess.run([exe], input=payload, stdout=subprocess.PIPE, check=True).stdout.decode().splitlines()
    return [list(map(int, l.split())) for l in out]


def test_refhash_known_answers_from_the_reference(refhash_exe):
    """csrc/kg_refhash.cuh (the reference's own hashes, for the bit-exact Bloom emulation of SURVEY 8f-4) against the
    known answers minted from the reference's objects: XXH64 of 8 bytes, base-5 hashes mod 2^54 of whole k-mers"""
    import json
    kats = json.load(open(os.path.join(GOLDEN, "kats.json")))
    payload = struct.pack("<II", 0, len(kats["xxh64"])) + b"".join(struct.pack("<QQ", v, s) for v, s, _ in kats["xxh64"])
    got = subprocess.run([refhash_exe], input=payload, stdout=subprocess.PIPE, check=True).stdout.decode().split()
    assert list(map(int, got)) == [h for _, _, h in kats["xxh64"]]
    n = 0
    for seq, q, _tbm, hf, hb, _fwd in kats["roll"]:
        if q != 1 << 54:
            continue
        rows = _refhash_windows(refhash_exe, ["ACGT".index(c) for c in seq], len(seq))
        assert len(rows) == 1 and rows[0][:3] == [hf, hb, min(hf, hb)]
        n += 1
    assert n >= 8


@pytest.mark.parametrize("k", [1, 5, 21, 31, 32, 33, 51, 64, 65, 127, 128, 255, 256])
def test_refhash_rolling_equals_definition(refhash_exe, oracle, k):
    """every window of a random run: Horner start + O(1) rolling updates == the polynomial definition (big integers),
    a few windows == the oracle's restatement of RollingHasherDual, and the 16 XXH64 values == the oracle's XXH64"""
    rng = np.random.default_rng(k)
    n = k + 150
    codes = [int(x) for x in rng.integers(0, 4, n)]
    rows = _refhash_windows(refhash_exe, codes, k)
    assert len(rows) == n - k + 1
    M54 = (1 << 54) - 1
    L = oracle.lib()
    for j, row in enumerate(rows):
        w = codes[j:j + k]
        hf = sum(w[i] * 5 ** (k - 1 - i) for i in range(k)) & M54
        hb = sum((3 - w[i]) * 5 ** i for i in range(k)) & M54
        assert row[:3] == [hf, hb, min(hf, hb)], j
        if j % 40 == 0:
            assert oracle.rolling_hashes(w, k, 1 << 54, True) == (hf, hb)
            assert row[3:] == [L.ko_xxh64_u64(min(hf, hb), L.ko_bloom_seed(i)) for i in range(16)]
