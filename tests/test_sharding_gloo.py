"""N > 1 host logic on CPU: two gloo ranks each take a byte range of one FASTA, repair the cut with
slice_context (k-1 bases of context + header state), count their range with the oracle, and the merged result
must equal the whole-file count; and the round protocol of the device exchange (csrc/kaarme_gpu.cu skm_round /
kg_pass_end) modelled with gloo all-reduces.  (The device exchange itself needs >= 2 GPUs: tests/test_gpu_multigpu.py.)"""
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


ROUNDS = textwrap.dedent("""
    # Model of the round protocol of the sharded path (csrc/kaarme_gpu.cu: skm_round, kg_pass_end): every round is one
    # all-reduce of "I have a batch"; a rank that has fed its last batch keeps taking part (contributing 0) until a round
    # whose sum is 0.  Slot b = round & 1 of a rank may be rewritten for round j + 2 only after the all-reduce of round
    # j + 1, which every rank enters after its insert of round j (the peers read the slot in place during that insert).
    import sys
    import torch, torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    batches = {batches}[rank]
    rounds, inserted_upto, log = 0, -1, []
    while True:
        have = 1 if rounds < batches else 0
        # (scatter of round `rounds` into slot rounds & 1 happens here: the slot's previous user was round - 2)
        t = torch.tensor([have, inserted_upto], dtype=torch.int64)
        s = t.clone(); dist.all_reduce(s, op=dist.ReduceOp.SUM)
        m = t.clone(); dist.all_reduce(m, op=dist.ReduceOp.MIN)
        # entering this all-reduce, every rank had finished the insert of round - 1: so slot (rounds - 1) & 1 is free
        assert int(m[1]) == rounds - 1, (rounds, int(m[1]))
        inserted_upto = rounds                      # insert of this round (reads the peers' slot `rounds & 1`)
        log.append(int(s[0]))
        rounds += 1
        if int(s[0]) == 0:
            break
    want = max({batches}) + 1
    assert rounds == want, (rank, rounds, want)
    assert log[:-1] == [sum(1 for b in {batches} if r < b) for r in range(want - 1)]
    print("ROUNDS_OK", rank, rounds)
    dist.destroy_process_group()
""")


@pytest.mark.parametrize("batches", [[3, 3], [5, 2], [0, 4], [0, 0], [1, 7]])
def test_round_protocol_terminates_together(tmp_path, batches):
    """ranks that feed different numbers of batches leave the pass after the same number of rounds (the first round in
    which nobody had a batch), and a slot is never rewritten before every rank has finished reading it"""
    script = tmp_path / "rounds.py"
    script.write_text(ROUNDS.format(batches=batches))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29650 + sum(batches)), str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.count("ROUNDS_OK") == 2, p.stdout[-3000:]
