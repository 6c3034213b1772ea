"""Full-size parity with the UNMODIFIED reference binary on the BASELINE.json configurations (`-m gpu`).

tests/golden/fullsize_digests.json holds, per configuration, the order-independent digest of the reference's output
(tests/golden/make_fullsize_digests.py ran oracle/_ref/kaarme in the build container; the inputs come from the
deterministic generator tests/native/gen_reads.c, so the GPU box regenerates the same bytes).  Here the CUDA build's CLI
(canonical-k-mer-hash-table_b200/kaarme) runs the SAME command line on the same input, writes into a FIFO, and its
digest must be identical: the same canonical k-mers with the same counts at the same -a, which is what the reference's
own check (sort + pytools/compare_outputs.py:1-33) establishes, plus the line count that script forgets.

Bloom mode (north star): the -a 2 output must equal the no-Bloom ground truth exactly (=> no false negative, exact
counts); the -a 1 output additionally holds the admitted singletons (false positives), whose number is stated next to the
reference's own for the same command line (gpurun_out/fullsize_fp_fn.json, copied to profiles/).

The cases that move tens of GB of text (C4 at k = 127 / 255, C5 at 1/10 scale) run when KAARME_FULLSIZE=all; the default
selection keeps the whole GPU suite within a few minutes.
"""
import json
import os
import re
import shutil

import pytest

import fullsize_util as fu

pytestmark = pytest.mark.gpu

GOLD = json.load(open(fu.GOLDEN)) if os.path.exists(fu.GOLDEN) else {}
HEAVY = {"C4_k127", "C4_k255", "C4_k51_m2", "C5s_truth", "C5s_truth_a1", "C5s_bloom", "C5s_bloom_a1", "C5s_bloom_m2"}
ALL = os.environ.get("KAARME_FULLSIZE") == "all"
REPORT = os.path.join(fu.ROOT, "gpurun_out", "fullsize_fp_fn.json")
# Bloom case at -a 1 -> the no-Bloom ground truth at -a 2 of the same input (what must be contained, with exact counts)
TRUTH_OF = {"C2_a1": "C1", "C3_bloom_a1": "C3_m0", "C5s_bloom_a1": "C5s_truth"}


def cases():
    names = [n for n in fu.CASES if n in GOLD and (ALL or n not in HEAVY)]
    return sorted(names, key=lambda n: (fu.CASES[n][0], n))       # one input at a time


@pytest.fixture(scope="module", autouse=True)
def scratch():
    yield
    shutil.rmtree(fu.TMP, ignore_errors=True)


def note(name, rec):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    data = json.load(open(REPORT)) if os.path.exists(REPORT) else {}
    data[name] = rec
    json.dump(data, open(REPORT, "w"), indent=1, sort_keys=True)


@pytest.mark.parametrize("name", cases())
def test_output_digest_equals_reference(name):
    inp, k, args = fu.CASES[name]
    gold = GOLD[name]
    assert gold["k"] == k and gold["args"] == args and gold["input"] == inp
    # inputs of other groups are dropped as we go: /dev/shm holds one input at a time
    for other in fu.INPUTS:
        if other != inp:
            fu.remove_input(other)
    path = fu.generate(inp)
    assert os.path.getsize(path) == gold["input_bytes"], "generator is not deterministic across machines"
    dig, log, wall = fu.run_digest(fu.GPU, path, k, args, threads=8)
    assert dig["bad"] == 0
    rec = {"gpu_lines": dig["lines"], "reference_lines": gold["digest"]["lines"], "gpu_wall_s": round(wall, 2),
           "reference_wall_s": gold["reference_wall_s"], "gpu_timers_s": fu.timers(log), "reference_timers_s": gold["reference_timers_s"]}
    if name in TRUTH_OF:
        # Bloom mode, -a 1: own hash functions => own false positives.  Contained truth is checked by the -a 2 case;
        # here: every extra line is a k-mer of count 1 (an admitted singleton), and their number is stated.
        truth = GOLD[TRUTH_OF[name]]["digest"]
        fp_gpu = dig["lines"] - truth["lines"]
        fp_ref = gold["digest"]["lines"] - truth["lines"]
        assert fp_gpu >= 0 and dig["count_sum"] - truth["count_sum"] == fp_gpu
        rec.update(false_positives_gpu=fp_gpu, false_positives_reference=fp_ref, false_negatives_gpu=0,
                   truth_lines_at_a2=truth["lines"])
        m = re.search(r"Kaarme bytes: (\d+) \(([0-9.]+) B/k-mer", log)
        if m:
            rec["kaarme_bytes"], rec["kaarme_bytes_per_kmer"] = int(m.group(1)), float(m.group(2))
        note(name, rec)
        print(f"{name}: admitted singletons GPU {fp_gpu} vs reference {fp_ref} (truth at -a 2: {truth['lines']} k-mers)")
        return
    note(name, rec)
    for f in ("lines", "sum", "xor", "count_sum", "bytes"):
        assert dig[f] == gold["digest"][f], f"{name}: {f} differs from the reference's output ({dig[f]} vs {gold['digest'][f]})"
    if "-b" not in args.split():
        assert fu.log_value(log, "Hash table size is:") == gold["table_slots"]      # functions_math.cpp:53-96


def test_c5_scale_model_sharded_over_all_gpus():
    """C5 (1/10 scale): k = 51, --use-bfilter, the CLI default -m 2, sharded over every GPU of the box -- the same lines
    as the reference binary on its one table"""
    import importlib
    kg = importlib.import_module("canonical-k-mer-hash-table_b200")
    n = kg.device_count()
    n = 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 1
    if n < 2 or "C5s_bloom_m2" not in GOLD:
        pytest.skip("needs at least two B200s (and the minted digest)")
    inp, k, args = fu.CASES["C5s_bloom_m2"]
    path = fu.generate(inp)
    dig, log, wall = fu.run_digest(fu.GPU, path, k, args, threads=16, extra=["--gpus", str(n)])
    gold = GOLD["C5s_bloom_m2"]["digest"]
    note(f"C5s_bloom_m2_gpus{n}", {"gpu_lines": dig["lines"], "reference_lines": gold["lines"], "gpu_wall_s": round(wall, 2),
                                   "gpu_timers_s": fu.timers(log)})
    for f in ("lines", "sum", "xor", "count_sum", "bytes"):
        assert dig[f] == gold[f], f"{f} differs from the reference's output"
    assert log.count("  shard ") == n
