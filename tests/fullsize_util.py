"""Shared plumbing of the full-size parity cases (TEST INFRASTRUCTURE; nothing here is on the product path).

Inputs come from tests/native/gen_reads.c (deterministic: the same bytes in the build container and on the GPU box),
outputs are reduced to an order-independent digest by tests/native/linedigest.c, read from a FIFO the counter writes
to with `-o` -- no 5-26 GB text file, no sort.  tests/golden/make_fullsize_digests.py runs the UNMODIFIED reference
binary (oracle/_ref/kaarme) through this and commits the digests; tests/test_gpu_fullsize_reference.py runs the CUDA
build (canonical-k-mer-hash-table_b200/kaarme) through the same plumbing on the GPU box and compares.
The check itself is the reference's own: same set of canonical k-mers with identical counts at the same -a
(pytools/compare_outputs.py:1-33 after sort), plus the line count that script forgets.
"""
import json
import os
import subprocess
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")
BUILD = os.path.join(NATIVE, "_build")
REF = os.path.join(ROOT, "oracle", "_ref", "kaarme")
GPU = os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "kaarme")
TMP = os.environ.get("KAARME_FULLSIZE_TMP", "/dev/shm/kaarme_fullsize")
GOLDEN = os.path.join(ROOT, "tests", "golden", "fullsize_digests.json")

# BASELINE.json configurations.  "ecoli" is the declared STAND-IN for example/ecoli1x.fasta
# (absent from the reference checkout): a seeded 4 641 652 bp random genome, one record, 70 columns, 140 repeats of
# 1 kbp so that the -a 2 output is not empty.  "c5s" is C5 at the 1/10 scale SURVEY.md section 8d allows.
INPUTS = {
    "ecoli": ["genome", 4_641_652, 1, 70, 140],
    "c3": ["reads", 5_000_000, 1_666_666, 150, 0.01, 42, 0],
    "c4": ["reads", 100_000_000, 200_000, 10_000, 0, 43, 80],
    "c5s": ["reads", 100_000_000, 20_000_000, 150, 0, 44, 0],
}

# name -> (input, k, command-line arguments shared by both binaries)
CASES = {
    "C1": ("ecoli", 51, "-m 0 -s 8000000 -a 2"),
    "C1_a1": ("ecoli", 51, "-m 0 -s 8000000 -a 1"),
    "C2": ("ecoli", 51, "-m 2 -u 4000000 -b -a 2"),
    "C2_a1": ("ecoli", 51, "-m 2 -u 4000000 -b -a 1"),
    "C3_m0": ("c3", 31, "-m 0 -s 160000000 -a 2"),
    "C3_m2": ("c3", 31, "-m 2 -s 160000000 -a 2"),
    "C3_m0_a1": ("c3", 31, "-m 0 -s 160000000 -a 1"),
    "C3_bloom": ("c3", 31, "-m 0 -b -u 80000000 -a 2"),
    "C3_bloom_a1": ("c3", 31, "-m 0 -b -u 80000000 -a 1"),
    "C3_bloom_m2": ("c3", 31, "-m 2 -b -u 80000000 -a 2"),
    "C4_k21": ("c4", 21, "-m 0 -s 250000000 -a 2"),
    "C4_k51": ("c4", 51, "-m 0 -s 250000000 -a 2"),
    "C4_k127": ("c4", 127, "-m 0 -s 250000000 -a 2"),
    "C4_k255": ("c4", 255, "-m 0 -s 250000000 -a 2"),
    "C4_k51_m2": ("c4", 51, "-m 2 -s 250000000 -a 2"),
    "C5s_truth": ("c5s", 51, "-m 0 -s 250000000 -a 2"),
    "C5s_truth_a1": ("c5s", 51, "-m 0 -s 250000000 -a 1"),
    "C5s_bloom": ("c5s", 51, "-m 0 -b -u 100000000 -a 2"),
    "C5s_bloom_a1": ("c5s", 51, "-m 0 -b -u 100000000 -a 1"),
    "C5s_bloom_m2": ("c5s", 51, "-m 2 -b -u 100000000 -a 2"),
}


def build_tools():
    os.makedirs(BUILD, exist_ok=True)
    for name in ("gen_reads", "linedigest"):
        src, exe = os.path.join(NATIVE, name + ".c"), os.path.join(BUILD, name)
        if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
            subprocess.run(["gcc", "-O2", "-o", exe, src], check=True)


def input_path(name):
    return os.path.join(TMP, name + ".fasta")


def generate(name):
    """Writes the input once per TMP directory; returns its path."""
    build_tools()
    os.makedirs(TMP, exist_ok=True)
    path = input_path(name)
    if not os.path.exists(path):
        spec = INPUTS[name]
        subprocess.run([os.path.join(BUILD, "gen_reads"), path + ".tmp", spec[0]] + [str(x) for x in spec[1:]], check=True)
        os.replace(path + ".tmp", path)
    return path


def remove_input(name):
    try:
        os.remove(input_path(name))
    except FileNotFoundError:
        pass


def run_digest(exe, path, k, args, threads, extra=(), timeout=7200):
    """Runs `exe INPUT k args -t threads extra -o FIFO`, digesting the FIFO; returns (digest dict, stdout log, wall s)."""
    build_tools()
    os.makedirs(TMP, exist_ok=True)
    fifo = os.path.join(TMP, f"out_{os.getpid()}_{threading.get_ident()}.fifo")
    if os.path.exists(fifo):
        os.remove(fifo)
    os.mkfifo(fifo)
    try:
        dig = subprocess.Popen(f"exec {os.path.join(BUILD, 'linedigest')} < {fifo}", shell=True, stdout=subprocess.PIPE, text=True)
        t0 = time.perf_counter()
        cmd = [exe, path, str(k)] + args.split() + ["-t", str(threads)] + list(extra) + ["-o", fifo]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
        wall = time.perf_counter() - t0
        # the counter may never have opened the FIFO (an error before the writer, -a 0): unblock the reader's open()
        try:
            os.close(os.open(fifo, os.O_WRONLY | os.O_NONBLOCK))
        except OSError:
            pass
        if p.returncode != 0:
            dig.kill()
            raise RuntimeError(f"{' '.join(cmd)} -> rc {p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
        out, _ = dig.communicate(timeout=600)
        return json.loads(out), p.stdout, wall
    finally:
        os.remove(fifo)


def log_value(log, prefix, cast=int, which=-1):
    """The number that follows `prefix` on a stdout line of either binary (e.g. 'Hash table size is:')."""
    for line in log.splitlines():
        if line.strip().startswith(prefix):
            rest = line.strip()[len(prefix):].replace("microseconds", " ").split()
            try:
                return cast(rest[which if which >= 0 else 0])
            except (ValueError, IndexError):
                continue
    return None


def timers(log):
    """The reference's own timers (parallel_parser.hpp:865-868, :2971), in seconds."""
    out = {}
    for tag, key in (("Time used to build hash table:", "build_s"), ("Time used to bloom filter k-mers:", "bloom_s"),
                     ("Time used to write k-mers in a file:", "write_s")):
        for line in log.splitlines():
            if line.startswith(tag):
                out[key] = int(line.split()[-2]) * 1e-6
    return out
