"""BASELINE.json full-size configurations through size-independent properties (the oracle takes minutes there):
 - sum of exported counts == input k-mers == n_reads * (L - k + 1)        (every window counted exactly once)
 - the direct insert and the L2-blocked partitioned insert export the same (k-mer, count) multiset (checksums)
 - C3 (1 % substitution errors): the measured number of k-mers with count >= 2 from SURVEY.md section 6."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

kg = importlib.import_module("canonical-k-mer-hash-table_b200")
K = kg.kaarme_gpu


def checksums(c, a):
    n, csum, ksum = 0, 0, np.uint64(0)
    W = c.W

    def sink(user, keys, counts, m):
        nonlocal n, csum, ksum
        kk = np.ctypeslib.as_array(keys, shape=(m * W,)).reshape(m, W)
        cc = np.ctypeslib.as_array(counts, shape=(m,)).astype(np.uint64)
        n += m
        csum += int(cc.sum())
        with np.errstate(over="ignore"):
            mix = (kk * np.arange(1, W + 1, dtype=np.uint64)[None, :] * np.uint64(0x9E3779B97F4A7C15)).sum(axis=1, dtype=np.uint64)
            ksum = ksum + (mix * (cc + np.uint64(1))).sum(dtype=np.uint64)
        return 0

    cb = K.SINK_FN(sink)
    c._check(K.lib().kg_export(c._h, a, K.COUNT_EXACT, cb, None), "kg_export")
    return n, csum, int(ksum)


def run(fasta, k, slots, partitions):
    with kg.Counter(k=k, min_slots=slots, partitions=partitions, batch_bytes=256 << 20) as c:
        c.pass_begin(K.PASS_COUNT)
        c.stream_begin(False)
        c.feed_device(fasta.data_ptr(), fasta.numel())
        st = c.pass_end()
        return st, checksums(c, 1), checksums(c, 2)


def test_c3_full_size():
    import torch
    import bench_data
    fasta, meta = bench_data.make_config("C3", torch.device("cuda", 0))
    assert meta["input_kmers"] == 199_999_920
    st1, all1, two1 = run(fasta, 31, meta["slots"], partitions=1)
    st2, all2, two2 = run(fasta, 31, meta["slots"], partitions=0)
    assert st1["partitions"] == 1 and st2["partitions"] > 1
    for st, al in ((st1, all1), (st2, all2)):
        assert st["input_kmers"] == meta["input_kmers"] == al[1]   # sum of counts == windows
        assert st["distinct"] == al[0]
        assert st["table_slots"] == 160_000_003                    # next_prime3mod4(160e6), functions_math.cpp:53-96
    assert all1 == all2 and two1 == two2                           # same multiset either way
    assert 6_500_000 < two1[0] < 8_000_000                         # ~7.13 M k-mers with count >= 2


def test_c4_k127_multiword_full_batches():
    """C4 shape at 1/8 size, k = 127 (4-word keys): direct == partitioned, counts sum to the windows"""
    import torch
    import bench_data
    fasta, meta = bench_data.make_config("C4", torch.device("cuda", 0), scale=0.125)
    n_in = meta["n_reads"] * (meta["L"] - 127 + 1)
    st1, all1, two1 = run(fasta, 127, meta["slots"], partitions=1)
    st2, all2, two2 = run(fasta, 127, meta["slots"], partitions=64)
    assert st1["input_kmers"] == st2["input_kmers"] == n_in == all1[1] == all2[1]
    assert all1 == all2 and two1 == two2
