#!/usr/bin/env python3
"""Mint the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; `make -C oracle ref` builds oracle/_ref/*):
    python tests/golden/make_golden.py
Writes
    g1_multiline.fasta g2_reads.fa g3_plain.txt g4_polya.fasta g5_long.fasta   seeded inputs
    golden.json    for every (input, k, mode, -a, bloom) the reference's sorted-output sha256, line count,
                   first lines, and the log numbers that pin table sizing / Bloom counters
    kats.json      XXH64 (vendored xxhash.c), rolling hash / prime / inverse (reference objects)
The reference has no tests or golden files of its own, so these are the pins.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as o  # noqa: E402

COMP = str.maketrans("ACGT", "TGCA")


def rand_dna(rng, n):
    return "".join("ACGT"[x] for x in rng.integers(0, 4, n))


def make_inputs():
    rng = np.random.default_rng(20261018)
    g = rand_dna(rng, 6000)
    # g1: multi-line records, lowercase, N, CR, '>' in the middle of a line, record shorter than k,
    # empty record, empty line inside a record, no trailing newline
    lines = [">rec1 multi-line 70 columns"]
    lines += [g[i:i + 70] for i in range(0, 2800, 70)]
    lines += [">rec2 lowercase and N"]
    lines += [g[2000:2300].lower(), g[2300:2400] + "N" + g[2400:2600], "", g[2600:2700]]
    lines += [">rec3 too short", "ACGTACGTAC"]
    lines += [">rec4 empty", ">rec5 CRLF inside"]
    lines += [g[3000:3300] + "\r", g[3300:3600]]
    lines += [">rec6 header symbol inside a sequence line"]
    lines += [g[100:400] + ">this tail is a header", g[400:900]]
    lines += [">rec7 reverse complement of rec1 start", g[0:700][::-1].translate(COMP)]
    lines += [">rec8 IUPAC", g[4000:4200] + "RYKM" + g[4200:4500]]
    lines += [">rec9 last record without newline", g[5000:5600]]
    with open(os.path.join(HERE, "g1_multiline.fasta"), "w", newline="") as f:
        f.write("\n".join(lines))
    # g2: short reads with substitution errors from a small genome (duplicates, counts up to ~30)
    genome = rand_dna(rng, 1500)
    recs = []
    for i in range(160):
        p = int(rng.integers(0, 1500 - 150))
        r = list(genome[p:p + 150])
        for j in range(150):
            if rng.random() < 0.01:
                r[j] = "ACGT"[("ACGT".index(r[j]) + int(rng.integers(1, 4))) & 3]
        r = "".join(r)
        if rng.random() < 0.5:
            r = r[::-1].translate(COMP)
        recs.append(f">r{i}\n{r}\n")
    with open(os.path.join(HERE, "g2_reads.fa"), "w") as f:
        f.write("".join(recs))
    # g3: PLAIN (one string per line)
    pl = [g[0:400], g[300:500].lower(), "ACGT", g[1000:1100] + "N" + g[1100:1400], "", g[350:800],
          g[0:400][::-1].translate(COMP)]
    with open(os.path.join(HERE, "g3_plain.txt"), "w") as f:
        f.write("\n".join(pl) + "\n")
    # g4: count-width artefacts: poly-A run of 70 000 (count 69 980 at k=21 > 65 535 > 16 383)
    with open(os.path.join(HERE, "g4_polya.fasta"), "w") as f:
        f.write(">polyA\n" + "A" * 70000 + "\n>tail\n" + g[0:300] + "\n" + "T" * 20000 + "\n")
    # g5: big enough that the Bloom-sized plain table (2 x new_in_second slots x 2 B counters) is a fresh
    # mmap: BasicAtomicFlagHashTableLong never zeroes counts[] (kmer_hash_table.cpp:2000-2001), so with
    # a smaller table the reference's own -m 0 -b output contains heap garbage.
    genome = rand_dna(rng, 45000)
    with open(os.path.join(HERE, "g5_long.fasta"), "w") as f:
        for i in range(3):
            p = 0 if i == 0 else int(rng.integers(0, 5000))
            r = genome[p:p + 42000]
            if i == 1:
                r = r[::-1].translate(COMP)
            f.write(f">long{i}\n" + "\n".join(r[j:j + 80] for j in range(0, len(r), 80)) + "\n")


def ref_case(name, k, mode, a, slots=None, unique=None, fpr=None):
    path = os.path.join(HERE, name)
    out, log = o.run_ref(path, k, mode=mode, slots=slots, unique=unique, fpr=fpr, min_abundance=a,
                         threads=3, out=f"/tmp/golden_{os.getpid()}.out")
    nums = {}
    for line in log.splitlines():
        for tag, key in (("Hash table size is:", "table_slots"),
                         ("New k-mers in first bloom filter", "new_in_first"),
                         ("New k-mers in second bloom filter", "new_in_second")):
            if line.startswith(tag):
                nums[key] = int(line.split()[-1])
    lines = out.splitlines()
    return {"input": name, "k": k, "mode": mode, "a": a, "slots": slots, "unique": unique, "fpr": fpr,
            "n_lines": len(lines), "sha256": hashlib.sha256(out).hexdigest(),
            "head": [ln.decode() for ln in lines[:2]], **nums}


def main():
    if not o.have_ref():
        o.build(ref=True)
    make_inputs()
    cases = []
    for name in ("g1_multiline.fasta", "g2_reads.fa", "g3_plain.txt"):
        for k in (21, 31, 32, 33, 51, 64, 127, 255):
            for mode in (0, 2):
                if mode == 2 and k % 32 == 0:
                    continue  # reference bug: -m 2 garbage for k = 0 mod 32
                for a in (1, 2):
                    cases.append(ref_case(name, k, mode, a, slots=200000))
    for mode in (0, 2):
        cases.append(ref_case("g4_polya.fasta", 21, mode, 2, slots=200000))
    # single worker (-t 3) => the Bloom counters and the -a 1 output are deterministic
    for name, k, u in (("g1_multiline.fasta", 21, 8000), ("g1_multiline.fasta", 51, 6000),
                       ("g2_reads.fa", 31, 4000), ("g3_plain.txt", 21, 2000)):
        for a in (1, 2):  # -m 2 only: see g5 for why -m 0 -b needs a big table
            cases.append(ref_case(name, k, 2, a, unique=u, fpr=0.01))
    for k in (31, 51):
        for mode in (0, 2):
            for a in (1, 2):
                cases.append(ref_case("g5_long.fasta", k, mode, a, unique=100000, fpr=0.01))
    cases.append(ref_case("g5_long.fasta", 31, 0, 2, unique=50000, fpr=0.2))
    cases.append(ref_case("g5_long.fasta", 51, 0, 2, slots=200000))
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(cases, f, indent=0)

    # KATs
    kat = {"xxh64": [], "roll": [], "prime": [], "inv": []}
    rng = np.random.default_rng(7)
    vals = [0, 1, 0x0123456789ABCDEF, (1 << 54) - 1] + [int(x) for x in rng.integers(0, 1 << 54, 12)]
    args = []
    for i, v in enumerate(vals):
        for s in (o.lib().ko_bloom_seed(i % 7), 0):
            args += [f"{v:x}", str(s)]
    for line in subprocess.run([o.REF_XXH] + args, check=True, stdout=subprocess.PIPE).stdout.decode().split("\n"):
        if line:
            v, s, h = line.split()
            kat["xxh64"].append([int(v, 16), int(s), int(h, 16)])
    kat_bin = os.path.join(os.path.dirname(o.REF_BIN), "ref_kat")
    strs = ["ACGTTGCAAGGCTTAACCGGT", "CGTTGCAAGGCTTAACCGGTA", rand_dna(rng, 31), rand_dna(rng, 51),
            rand_dna(rng, 127), "A" * 21, "T" * 33, "ACGT" * 8 + "A"]
    for s in strs:
        for q, tbm in ((8000023, 0), (160000003, 0), (1 << 54, 0), (1 << 54, 1)):
            r = subprocess.run([kat_bin, "roll", str(q), str(tbm), s], check=True, stdout=subprocess.PIPE)
            hf, hb, fc = r.stdout.decode().split()
            kat["roll"].append([s, q, tbm, int(hf), int(hb), int(fc)])
    for n in (0, 2, 3, 100, 4000000, 8000000, 14260556, 160000000):
        r = subprocess.run([kat_bin, "prime", str(n)], check=True, stdout=subprocess.PIPE)
        kat["prime"].append([n, int(r.stdout)])
    for a, m in ((5, 1 << 54), (5, 8000023), (5, 160000003), (5, 4000039)):
        r = subprocess.run([kat_bin, "inv", str(a), str(m)], check=True, stdout=subprocess.PIPE)
        kat["inv"].append([a, m, int(r.stdout)])
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(kat, f)
    print(f"{len(cases)} golden cases, {sum(len(v) for v in kat.values())} KATs")


if __name__ == "__main__":
    main()
