#!/usr/bin/env python3
"""Full-size parity of the drop-in CLI against the UNMODIFIED reference binary (oracle/_ref/kaarme), on a B200 box.

Not collected by pytest (takes minutes).  For the BASELINE.json configurations that fit a quick run it generates the
input file (seeded), runs `oracle/_ref/kaarme` on the host cores and `canonical-k-mer-hash-table_b200/kaarme` on the
GPU with the SAME command line, sorts both outputs and requires them to be byte-identical (the check
pytools/compare_outputs.py does, plus the line-count check it forgets).  Prints one line per configuration with the
reference's and the GPU build's own "Time used to build hash table" timers.

    python tests/fullsize_reference_parity.py [--quick] [--extras]

C1/C2 use a declared STAND-IN for example/ecoli1x.fasta (absent from the reference checkout, .MISSING_LARGE_BLOBS):
a seeded 4 641 652 bp random genome, one record, 70 columns, with 3 % of it re-inserted as repeats so that the
`-a 2` output is not empty.
"""
import hashlib
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "kaarme")
GPU = os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "kaarme")
TMP = "/dev/shm/kaarme_parity"


def ecoli_standin(path):
    rng = np.random.default_rng(1)
    G = 4_641_652
    g = rng.integers(0, 4, G).astype(np.uint8)
    for _ in range(140):                                # ~3 % repeats of 1 kbp
        a, b = int(rng.integers(0, G - 1000)), int(rng.integers(0, G - 1000))
        g[b:b + 1000] = g[a:a + 1000]
    s = np.frombuffer(b"ACGT", np.uint8)[g].tobytes()
    with open(path, "wb") as f:
        f.write(b">ecoli1x_standin seed=1 G=4641652\n")
        for i in range(0, G, 70):
            f.write(s[i:i + 70] + b"\n")


def c3_reads(path, scale):
    import torch
    import bench_data
    dev = torch.device("cuda", 0)
    fasta, meta = bench_data.make_config("C3", dev, scale=scale)
    with open(path, "wb") as f:
        f.write(fasta.cpu().numpy().tobytes())
    return meta


def timer(log, tag="Time used to build hash table:"):
    for line in log.splitlines():
        if line.startswith(tag):
            return int(line.split()[-2]) * 1e-6
    return float("nan")


def run(exe, args, out):
    if os.path.exists(out):
        os.remove(out)
    t0 = time.perf_counter()
    p = subprocess.run([exe] + [str(a) for a in args] + ["-o", out], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{exe} rc={p.returncode}\n{p.stdout[-1500:]}")
    return p.stdout, wall


def sorted_digest(path):
    srt = path + ".sorted"
    subprocess.run(f"LC_ALL=C sort -S 2G --parallel=8 {path} > {srt}", shell=True, check=True)
    h, n = hashlib.sha256(), 0
    with open(srt, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
            n += b.count(b"\n")
    os.remove(srt)
    return n, h.hexdigest()


def case(name, path, k, args_common, gpu_extra=(), ref_threads=None):
    threads = max(3, min(64, os.cpu_count() or 3))
    ref_log, ref_wall = run(REF, [path, k] + args_common + ["-t", ref_threads or threads], f"{TMP}/ref.out")
    gpu_log, gpu_wall = run(GPU, [path, k] + args_common + ["-t", threads] + list(gpu_extra), f"{TMP}/gpu.out")
    rn, rh = sorted_digest(f"{TMP}/ref.out")
    gn, gh = sorted_digest(f"{TMP}/gpu.out")
    ok = (rn, rh) == (gn, gh)
    extra = ""
    for line in gpu_log.splitlines():
        if line.startswith("Kaarme bytes:") or line.startswith("Max secondary"):
            extra += " | " + line.strip()
    for line in ref_log.splitlines():
        if line.startswith("Main array slots used") or line.startswith("Max secondary"):
            extra += " | ref: " + line.strip()
    print(f"{name}: {'IDENTICAL' if ok else 'DIFFERENT'}  lines ref {rn} gpu {gn}  sha {gh[:16]}  "
          f"build-table timer ref {timer(ref_log):.2f} s ({threads - 2} workers) vs gpu {timer(gpu_log):.3f} s  "
          f"wall ref {ref_wall:.1f} s gpu {gpu_wall:.1f} s{extra}", flush=True)
    return ok


def main():
    quick = "--quick" in sys.argv
    os.makedirs(TMP, exist_ok=True)
    ok = True
    eco = f"{TMP}/ecoli1x_standin.fasta"
    ecoli_standin(eco)
    ok &= case("C1 ecoli-standin k=51 -m 0 -s 8000000 -a 2", eco, 51, ["-m", 0, "-s", 8000000, "-a", 2])
    ok &= case("C1' ecoli-standin k=51 -m 0 -s 8000000 -a 1", eco, 51, ["-m", 0, "-s", 8000000, "-a", 1])
    ok &= case("C2 ecoli-standin k=51 -m 2 -u 4000000 -b -a 2", eco, 51, ["-m", 2, "-u", 4000000, "-b", "-a", 2])
    ok &= case("C2' ecoli-standin k=51 -m 2 -s 8000000 -a 1", eco, 51, ["-m", 2, "-s", 8000000, "-a", 1])
    c3 = f"{TMP}/c3.fasta"
    meta = c3_reads(c3, 0.1 if quick else 1.0)
    s = int(160_000_000 * (0.1 if quick else 1.0))
    ok &= case(f"C3 {meta['n_reads']} reads k=31 -m 0 -s {s} -a 2", c3, 31, ["-m", 0, "-s", s, "-a", 2])
    ok &= case(f"C3 {meta['n_reads']} reads k=31 -m 2 -s {s} -a 2", c3, 31, ["-m", 2, "-s", s, "-a", 2])
    ok &= case(f"C3 {meta['n_reads']} reads k=31 -m 0 -b -u {s // 2} -a 2", c3, 31, ["-m", 0, "-b", "-u", s // 2, "-a", 2])
    if "--extras" in sys.argv:
        # rows added after the count path, at full size:
        #  host formatter instead of the GPU text dump; saved Kaarme structure decoded again; and the bit-exact Bloom
        #  emulation against the reference run with ONE worker (-t 3), where its -a 1 output (false positives included)
        #  is deterministic
        ok &= case(f"C3 k=31 -m 0 -a 2 --host-format", c3, 31, ["-m", 0, "-s", s, "-a", 2], gpu_extra=["--host-format"])
        run(GPU, [c3, 31, "-m", 2, "-s", s, "-a", 2, "--dump-kaarme", f"{TMP}/c3.kaarme"], f"{TMP}/gpu_dump.out")
        run(GPU, [f"{TMP}/c3.kaarme", 31, "--from-kaarme", "-a", 2], f"{TMP}/gpu_decoded.out")
        same = sorted_digest(f"{TMP}/gpu_dump.out") == sorted_digest(f"{TMP}/gpu_decoded.out")
        print(f"C3 k=31 -m 2 --dump-kaarme -> --from-kaarme: {'IDENTICAL' if same else 'DIFFERENT'} "
              f"({os.path.getsize(f'{TMP}/c3.kaarme')} bytes on disk)", flush=True)
        ok &= same
        ok &= case("C2' ecoli-standin k=51 -m 0 -u 4000000 -b -a 1, reference with one worker vs --reference-bloom", eco, 51,
                   ["-m", 0, "-u", 4000000, "-b", "-a", 1], gpu_extra=["--reference-bloom"], ref_threads=3)
    for f in os.listdir(TMP):
        os.remove(os.path.join(TMP, f))
    print("ALL IDENTICAL" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
