"""N > 1 host logic on CPU: two gloo ranks each take a byte range of one FASTA, repair the cut with
slice_context (k-1 bases of context + header state), count their range with the oracle, and the merged result
must equal the whole-file count.  (The device exchange itself needs >= 2 GPUs: tests/multigpu_check.py.)"""
import importlib
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT

kg = importlib.import_module("canonical-k-mer-hash-table_b200")
K = kg.kaarme_gpu

WORKER = textwrap.dedent("""
    import importlib, os, sys, zlib
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
    import oracle_py as o
    K = importlib.import_module("canonical-k-mer-hash-table_b200").kaarme_gpu
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    data = open({path!r}, "rb").read()
    k = {k}
    lo, hi = K.shard_ranges(len(data), world)[rank]
    ctx_lo, hdr = K.slice_context(data, lo, k)
    a = o.count(data[ctx_lo:hi], k, o.FASTA, hdr)
    b = o.count(data[ctx_lo:lo], k, o.FASTA, hdr)
    def as_dict(c):
        return dict(zip(map(bytes, c.keys.view(np.uint8).reshape(c.n, c.W * 8)), c.counts.tolist()))
    mine = as_dict(a)
    for kk, cc in as_dict(b).items():
        mine[kk] -= cc
    # hash-shard ownership: send each k-mer to its owner rank, like the device exchange does
    outbox = [dict() for _ in range(world)]
    for kk, cc in mine.items():
        if cc:
            outbox[zlib.crc32(kk) % world][kk] = cc
    gathered = [None] * world
    dist.all_gather_object(gathered, outbox)
    shard = dict()
    for sender in gathered:
        for kk, cc in sender[rank].items():
            shard[kk] = shard.get(kk, 0) + cc
    allshards = [None] * world
    dist.all_gather_object(allshards, shard)
    if rank == 0:
        whole = o.count(data, k)
        want = as_dict(whole)
        merged = dict()
        for s in allshards:
            assert not (set(s) & set(merged)), "shards must own disjoint k-mers"
            merged.update(s)
        assert merged == want, (len(merged), len(want))
        print("SHARDING_OK", len(want))
    dist.destroy_process_group()
""")


@pytest.mark.parametrize("k", [21, 51])
def test_two_rank_sharding(tmp_path, k):
    rng = np.random.default_rng(k)
    g = "".join("ACGT"[x] for x in rng.integers(0, 4, 30000))
    recs = []
    for i in range(25):
        p = int(rng.integers(0, 30000 - 2500))
        r = g[p:p + 2500]
        recs.append(f">read{i} some long header text so that a cut can land inside it {i}\n" +
                    "\n".join(r[j:j + 60] for j in range(0, 2500, 60)) + "\n")
    path = tmp_path / "in.fasta"
    path.write_text("".join(recs))
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, path=str(path), k=k))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29600 + k), str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0 and "SHARDING_OK" in p.stdout, p.stdout[-3000:]


@pytest.mark.parametrize("cut_in", ["header", "line", "newline", "start", "end"])
def test_slice_context_cuts(oracle, cut_in):
    """every kind of cut position: inside a header, inside a sequence line, on a newline, at 0 and at EOF"""
    data = (b">r1 header one\nACGTTGCAAGGCTTAACCGGTACGTTGCAAGG\nCTTAACCGGTACGATCGATCGGATCGATTTAG\n"
            b">r2 second header\nGGGATTTACCCAGGATTTACGGATTACAGGAT\nTTACCAGGGATTTTACCCGGGAAATTTCCCGG\n")
    k = 21
    pos = {"header": data.index(b"second"), "line": data.index(b"GGATCGATTTAG"), "newline": data.index(b"\nCTTAACC"),
           "start": 0, "end": len(data)}[cut_in]
    ctx_lo, hdr = K.slice_context(data, pos, k)
    whole = oracle.count(data, k)
    left = oracle.count(data[:pos], k)
    a = oracle.count(data[ctx_lo:], k, oracle.FASTA, hdr)
    b = oracle.count(data[ctx_lo:pos], k, oracle.FASTA, hdr)
    assert a.total_windows - b.total_windows == whole.total_windows - left.total_windows
