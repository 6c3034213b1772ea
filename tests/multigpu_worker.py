"""Multi-GPU parity worker: run under torchrun on a box with >= 2 B200s, normally BY tests/test_gpu_multigpu.py
(a `-m gpu` pytest module that skips on fewer than two GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_worker.py
Every rank feeds its own slice of one FASTA (byte-range split + k-1 bases of context), canonical k-mers are sharded by
the hash of their minimizer, and the union of the shard exports must equal the oracle's count of the whole file, bit
for bit -- plain tables, uneven slices (ranks feed different numbers of batches), multi-word keys, the double Bloom
filter, and the Kaarme structure built per shard and decoded again.  Rank 0 prints one line per case and a final
`MULTIGPU_RESULT {json}` line; the exit code is the number of failed cases."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

# (k, Bloom, uneven slices, partitions per shard, table mode)
CASES = ((51, False, False, 1, 0), (31, False, True, 4, 0), (127, False, False, 16, 0), (51, True, False, 8, 0),
         (51, False, True, 0, 0), (21, True, True, 32, 0), (31, True, False, 0, 0), (255, False, True, 2, 0),
         (51, False, False, 4, 2), (31, False, True, 1, 2), (127, True, False, 2, 2), (9, False, False, 2, 0))


def main():
    kg = importlib.import_module("canonical-k-mer-hash-table_b200")
    K = kg.kaarme_gpu
    import oracle_py as oracle
    from test_gpu_parity import make_fasta
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures, results = 0, []
    cases = CASES
    if os.environ.get("KAARME_MULTIGPU_CASES"):                       # e.g. "0,3,8": a subset (box time on 8 GPUs is dear)
        cases = [CASES[int(i)] for i in os.environ["KAARME_MULTIGPU_CASES"].split(",")]
    for k, use_bloom, uneven, npart, mode in cases:
        rng = np.random.default_rng(1234 + k)
        data = make_fasta(rng, 200000, 600, 2000, wrap=80, err=0.01, n_rate=0.0003)
        truth = oracle.count(data, k) if rank == 0 else None          # the checker runs on rank 0 only
        n_distinct = [truth.n if rank == 0 else None]
        dist.broadcast_object_list(n_distinct, src=0)
        # uneven: rank 0 gets 3/4 of the file, so ranks feed different numbers of batches
        if uneven:
            cuts = [0] + [len(data) * 3 // 4 + (len(data) // 4) * i // (world - 1) for i in range(world)]
        else:
            cuts = [len(data) * i // world for i in range(world + 1)]
        cuts[-1] = len(data)
        lo, hi = cuts[rank], cuts[rank + 1]
        ctx_lo, in_hdr = K.slice_context(data, lo, k)
        c = kg.Counter(k=k, table_mode=mode, min_slots=2_000_000, use_bloom=use_bloom, expected_unique=n_distinct[0], fpr=0.01,
                       device=local, rank=rank, world=world, batch_bytes=1 << 20, partitions=npart)
        uid = [kg.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        c.comm_init(uid[0], rank, world)
        stats = []
        for which in ([K.PASS_BLOOM] if use_bloom else []) + [K.PASS_COUNT]:
            c.pass_begin(which)
            c.stream_begin(in_hdr)
            c.feed(data[ctx_lo:lo], K.FEED_CONTEXT)
            c.feed(data[lo:hi])
            stats.append(c.pass_end())
        cs = c.compact() if mode == K.TABLE_KAARME else None
        keys, counts = c.export(2 if use_bloom else 1, K.COUNT_EXACT if mode == 0 else K.COUNT_REFERENCE)
        parts = [None] * world
        dist.all_gather_object(parts, (keys, counts, stats[-1]["input_kmers"], stats[-1]["inserted_kmers"], cs))
        if rank == 0:
            allk = np.concatenate([p[0] for p in parts])
            allc = np.concatenate([p[1] for p in parts])
            order = np.lexsort([allk[:, j] for j in range(allk.shape[1] - 1, -1, -1)])
            allk, allc = allk[order], allc[order]
            want = truth.filtered(2 if use_bloom else 1, oracle.TABLE_KAARME if mode == 2 else oracle.TABLE_EXACT)
            ok = (allk.shape == want.keys.shape and (allk == want.keys).all() and (allc.astype(np.uint64) == want.counts).all()
                  and sum(p[2] for p in parts) == truth.total_windows)
            if not use_bloom:
                ok = ok and sum(p[3] for p in parts) == truth.total_windows
            extra = ""
            if mode == K.TABLE_KAARME:
                kb = [round(p[4]["bytes"] / max(1, p[4]["kmers"]), 2) for p in parts]
                roots = sum(p[4]["roots"] for p in parts)
                extra = f"; Kaarme B/k-mer per shard {kb}, roots {roots} of {sum(p[4]['kmers'] for p in parts)}"
            line = (f"multigpu k={k} bloom={use_bloom} uneven={uneven} partitions={npart} mode={mode} world={world}: "
                    f"{'OK' if ok else 'MISMATCH'} (distinct {len(allc)} vs {want.n}; input {sum(p[2] for p in parts)} vs "
                    f"{truth.total_windows}; per-shard {[len(p[1]) for p in parts]}{extra})")
            print(line, flush=True)
            results.append({"k": k, "bloom": use_bloom, "uneven": uneven, "partitions": npart, "mode": mode, "ok": bool(ok)})
            failures += 0 if ok else 1
        c.close()
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_RESULT " + json.dumps({"world": world, "cases": results, "failures": failures}), flush=True)
    dist.destroy_process_group()
    return failures


if __name__ == "__main__":
    sys.exit(main())
