#!/usr/bin/env python3
"""bench.py -- the driver's measurement contract for the k-mer counting hot path.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

metric   input k-mers/s (BASELINE.json), k=51, config C4 (100 Mbp genome, 20x of 10 kbp reads, -m 0
         -s 250000000): the configuration the metric is quoted on; fits one GPU (2.03 GB FASTA + 8 GB table).
step     one full counting pass over the whole synthetic read set into a freshly cleared table
         (kg_pass_begin .. kg_pass_end through the C ABI).
value    device-timed (CUDA events inside the library, compute stream), input already resident in HBM.
e2e      same pass fed from PINNED HOST memory through kg_feed (H2D inside the timed region) + the D2H read of
         the pass statistics; wall clock bracketed by synchronize.
N > 1    weak scaling: every rank draws its own 20x read set from a genome N times larger, k-mers are
         hash-sharded across ranks and exchanged with NCCL.
Only the cpu_baseline / --impl reference legs execute anything under oracle/ (the unmodified reference
binary oracle/_ref/kaarme when it was built, else the C port) -- as the thing being compared against.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "input_kmers_per_sec_k51"
UNIT = "k-mers/s"
WORKLOAD = "C4"
K = 51


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--scale", type=float, default=float(os.environ.get("KAARME_BENCH_SCALE", "1.0")),
                   help="fraction of the named workload (1.0 = the BASELINE configuration; <1 only for debugging)")
    p.add_argument("--k", type=int, default=K)
    p.add_argument("--workload", default=WORKLOAD)
    p.add_argument("--partitions", type=int, default=int(os.environ.get("KAARME_PARTITIONS", "0")),
                   help="table regions each batch is bucketed into before inserting (0 = library default, 1 = direct insert)")
    p.add_argument("--batch-mb", type=int, default=int(os.environ.get("KAARME_BATCH_MB", "256")))
    p.add_argument("--bloom", action="store_true", help="two-pass double-Bloom-filter mode (-b -u <genome size>)")
    p.add_argument("--fpr", type=float, default=0.01)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


# ---- clocks ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---- CPU reference leg -----------------------------------------------------------------------------------------
# The reference's own counter (the unmodified binary oracle/_ref/kaarme) on this box's host cores, on a sample of the
# SAME workload: REF_FRACTION of the genome at the same coverage (so every k-mer is seen as often as in the full job and
# the table is hit/missed in the same proportion), table -s scaled alike.  The timed quantity is its own "Time used to
# build hash table" (parallel_parser.hpp:865-867: read + count).  Runs pass -a 65535 so that the untimed single-threaded
# writer emits nothing; the file_to_file leg runs once more with -a 2 and takes the wall clock.
REF_FRACTION = 0.1


def workload_string(c, k, scale, world, bloom=False, fpr=0.01):
    G = int(c["G"] * scale * world)
    tail = f"-b -u {G} -f {fpr}" if bloom else f"-s {int(c['slots'] * scale) * world}"
    return f"{c['name']}: {G} bp genome, {c['cov']}x of {c['L']} bp reads, k={k}, -m 0 {tail}"


def cpu_reference_run(sample_path, k, slots, threads, write=False, timeout=1500):
    """-> dict(build_s, write_s, wall_s, out_bytes) of one run of oracle/_ref/kaarme on a file"""
    import oracle.oracle_py as o  # the one place bench.py executes oracle/: as the baseline being measured
    out = sample_path + ".out"
    cmd = [o.REF_BIN, sample_path, str(k), "-m", "0", "-s", str(slots), "-a", "2" if write else "65535", "-t", str(threads), "-o", out]
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=timeout, text=True)
    wall = time.perf_counter() - t0
    nbytes = os.path.getsize(out) if os.path.exists(out) else 0
    if os.path.exists(out):
        os.remove(out)
    r = {"wall_s": wall, "out_bytes": nbytes}
    for line in p.stdout.splitlines():
        if line.startswith("Time used to build hash table:"):
            r["build_s"] = int(line.split()[-2]) * 1e-6
        if line.startswith("Time used to write k-mers in a file:"):
            r["write_s"] = int(line.split()[-2]) * 1e-6
    if "build_s" not in r:
        raise RuntimeError("reference run failed: " + p.stdout[-500:])
    return r


def cpu_port_run(sample_bytes, k):
    import oracle.oracle_py as o
    t0 = time.perf_counter()
    c = o.count(sample_bytes, k)
    return time.perf_counter() - t0, c.total_windows


def reference_fraction(workload, scale, bench_data, threads):
    """about REF_FRACTION of the workload, rounded so that the sample is a whole number of 10 MiB chunks per worker: the
    reference's parallelism is one chunk per worker at a time (parallel_parser.hpp:834-843), and a sample that leaves
    workers idle in its last wave would understate it"""
    c = bench_data.CONFIGS[workload]
    full_bytes = c["G"] * scale * c["cov"] * (1.0 + 1.0 / c["wrap"] if c["wrap"] else 1.0 + 12.0 / c["L"])
    workers = max(1, threads - 2)
    waves = max(1, round(REF_FRACTION * full_bytes / (10 << 20) / workers))
    return min(1.0, waves * workers * (10 << 20) * 0.999 / full_bytes)


def reference_sample(workload, k, scale, torch, bench_data, dev, threads=None):
    """the sample every reference step counts: ~REF_FRACTION of the workload, same coverage -> (path, meta)"""
    threads = threads or max(3, min(64, os.cpu_count() or 3))
    frac = reference_fraction(workload, scale, bench_data, threads)
    fasta, meta = bench_data.make_config(workload, dev, scale=scale * frac)
    meta["fraction"] = frac
    data = fasta.cpu().numpy().tobytes()
    path = None
    for d in ("/dev/shm", "/tmp"):                     # RAM-backed if there is room, else local disk
        try:
            path = os.path.join(d, f"kaarme_bench_ref_{os.getpid()}.fasta")
            with open(path, "wb") as f:
                f.write(data)
            break
        except OSError:
            if path and os.path.exists(path):
                os.remove(path)
            path = None
    if path is None:
        raise RuntimeError("no room for the reference sample in /dev/shm or /tmp")
    del data
    meta["input_kmers"] = meta["n_reads"] * (meta["L"] - k + 1)
    meta["path_bytes"] = int(fasta.numel())
    del fasta
    return path, meta


def sample_string(meta, k):
    return (f"{meta['fraction']:.3f} of the workload at the same coverage (whole 10 MiB chunks per worker): {meta['G']} bp genome, {meta['cov']}x of {meta['L']} bp reads "
            f"({meta['n_reads']} reads, {meta['path_bytes']} bytes, {meta['input_kmers']} input k-mers), -m 0 -s {meta['slots']}")


def cpu_baseline(workload, k, scale, torch, bench_data, dev):
    import oracle.oracle_py as o
    cores = os.cpu_count() or 1
    if o.have_ref():
        threads = max(3, min(64, cores))
        path, meta = reference_sample(workload, k, scale, torch, bench_data, dev)
        try:
            r = cpu_reference_run(path, k, meta["slots"], threads)
            w = cpu_reference_run(path, k, meta["slots"], threads, write=True)      # file -> output file, wall clock
        finally:
            os.remove(path)
        return {"value": meta["input_kmers"] / r["build_s"], "unit": UNIT, "cores": threads - 2, "kind": "reference", "host_cores": cores,
                "sample": sample_string(meta, k) + f"; oracle/_ref/kaarme -t {threads}, its own build-table timer", "seconds": r["build_s"],
                "file_to_file": {"wall_s": w["wall_s"], "kmers_per_s": meta["input_kmers"] / w["wall_s"], "output_bytes": w["out_bytes"],
                                 "build_s": w["build_s"], "write_s": w.get("write_s")}}
    fasta, meta = bench_data.make_config(workload, dev, scale=scale * 0.02)
    data = fasta.cpu().numpy().tobytes()
    secs, tw = cpu_port_run(data, k)
    return {"value": tw / secs, "unit": UNIT, "cores": 1, "kind": "port", "host_cores": cores,
            "sample": f"0.02 of the workload ({len(data)} bytes), oracle/liboracle.so ko_count", "seconds": secs}


def split_u64(vals):
    out = []
    for v in vals:
        out += [v & 0xFFFFFFFF, v >> 32]
    return out


def join_u64(parts):
    return [(int(parts[2 * i]) + (int(parts[2 * i + 1]) << 32)) & 0xFFFFFFFFFFFFFFFF for i in range(len(parts) // 2)]


# ---- main ------------------------------------------------------------------------------------------------------
def trace(msg):
    if os.environ.get("KAARME_BENCH_TRACE"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def main():
    args = parse_args()
    if os.environ.get("KAARME_BENCH_TRACE"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["KAARME_BENCH_TRACE"]), exit=False)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    import torch
    import torch.distributed as dist
    import bench_data

    k = args.k
    if args.impl == "reference":
        return run_reference(args, torch, bench_data)

    if not torch.cuda.is_available():
        print(json.dumps({"metric": METRIC, "error": "no CUDA device: the product path has no CPU fallback"}))
        return 2
    kg = importlib.import_module("canonical-k-mer-hash-table_b200")
    KG = kg.kaarme_gpu
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    fasta, meta = bench_data.make_config(args.workload, dev, scale=args.scale, rank=rank, world=world)
    meta["name"] = args.workload
    meta["k"] = k
    meta["input_kmers"] = meta["n_reads"] * (meta["L"] - k + 1)
    torch.cuda.synchronize()
    total_slots = meta["slots"] * world
    ctr = kg.Counter(k=k, table_mode=KG.TABLE_PLAIN, input_mode=KG.INPUT_FASTA, min_slots=total_slots,
                     use_bloom=args.bloom, expected_unique=meta["G"] if args.bloom else 0, fpr=args.fpr,
                     device=local_rank, rank=rank, world=world, batch_bytes=args.batch_mb << 20,
                     partitions=args.partitions)
    if world > 1:
        uid = [kg.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctr.comm_init(uid[0], rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    bloom_info = {}

    def one_pass(which, feed):
        ctr.pass_begin(which)
        ctr.stream_begin(False)
        feed()
        return ctr.pass_end()

    def step(feed):
        """one step = the whole job: (Bloom pass +) count pass; device_ms etc. are summed over the passes"""
        if not args.bloom:
            return one_pass(KG.PASS_COUNT, feed)
        b = one_pass(KG.PASS_BLOOM, feed)
        st = one_pass(KG.PASS_COUNT, feed)
        bloom_info.update(new_in_first=b["new_in_first"], new_in_second=b["new_in_second"], bloom_bits=b["bloom_bits"],
                          bloom_hashes=b["bloom_hashes"], bloom_pass_ms=b["device_ms"], count_pass_ms=st["device_ms"],
                          bloom_partitions=b["partitions"], table_slots=st["table_slots"])
        for key in ("device_ms", "count_ms", "parse_ms"):
            st[key] += b[key]
        return st

    def step_device():
        return step(lambda: ctr.feed_device(fasta.data_ptr(), fasta.numel()))

    # ---- device-resident metric -------------------------------------------------------------------------------
    trace("context ready, warm-up")
    for _ in range(args.warmup):
        st = step_device()
        trace("warm-up step done")
    assert st["input_kmers"] == meta["input_kmers"], (st["input_kmers"], meta["input_kmers"])
    launches0 = ctr.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, count_ms, parse_ms, insert_ms, insert_launches = 0.0, 0.0, 0.0, 0.0, 0
    for _ in range(args.steps):
        st = step_device()
        dev_ms += st["device_ms"]
        count_ms += st["count_ms"]
        parse_ms += st["parse_ms"]
        insert_ms += st["insert_ms"]
        insert_launches += st["insert_launches"]
    barrier()
    trace("timed steps done")
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctr.launch_count() - launches0
    distinct = st["distinct"]
    inserted_rank = st["inserted_kmers"]
    t = torch.tensor([dev_ms, wall * 1e3, count_ms, parse_ms, insert_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, count_ms, parse_ms, insert_ms = t.tolist()
    total_kmers = meta["input_kmers"] * world
    value = total_kmers * args.steps / (dev_ms * 1e-3)

    # ---- end to end through the C ABI from pinned host memory --------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty(fasta.numel(), dtype=torch.uint8, pin_memory=True)
        host.copy_(fasta)
        torch.cuda.synchronize()

        def step_host():
            return step(lambda: ctr.feed(host))

        for _ in range(min(args.warmup, 2)):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st2 = step_host()
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        assert st2["input_kmers"] == meta["input_kmers"]
        trace("e2e done")
        import ctypes
        e2e = {"value": total_kmers * args.steps / te.item(), "unit": UNIT,
               "h2d_bytes_per_step": int(fasta.numel()) * world * (2 if args.bloom else 1),
               "d2h_bytes_per_step": ctypes.sizeof(KG.PassStats) * world,
               "ms_per_step": te.item() * 1e3 / args.steps}
        del host

    # ---- correctness of what was just timed (untimed) ----------------------------------------------------------------
    # every window counted exactly once: sum over shards of (sum of counts) == sum over ranks of input k-mers; and the
    # sharded / bucketed result equals an independent count of the same reads by ONE table on rank 0 through the direct
    # path (kg_count_kernel, no bucketing, no exchange): order-independent checksums (kg_checksum), additive over shards.
    verify = None
    if not args.bloom:
        mine = ctr.checksum(1, KG.COUNT_EXACT)
        tsum = torch.tensor(split_u64(mine) + [meta["input_kmers"]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tsum)
        parts = tsum.tolist()
        sharded = join_u64(parts[:8])
        if rank == 0:
            single = kg.Counter(k=k, table_mode=KG.TABLE_PLAIN, input_mode=KG.INPUT_FASTA, min_slots=total_slots, device=local_rank,
                                batch_bytes=args.batch_mb << 20, partitions=1)
            single.pass_begin(KG.PASS_COUNT)
            for r in range(world):
                fr = fasta if r == rank else bench_data.make_config(args.workload, dev, scale=args.scale, rank=r, world=world)[0]
                torch.cuda.synchronize()          # the library reads on its own (non-blocking) streams
                single.stream_begin(False)
                single.feed_device(fr.data_ptr(), fr.numel())
                torch.cuda.synchronize()
                if r != rank:
                    del fr
            single.pass_end()
            ref = list(single.checksum(1, KG.COUNT_EXACT))
            single.close()
            verify = {"kmers": sharded[0], "sum_of_counts": sharded[1], "input_kmers_all_ranks": parts[8],
                      "checksum": [hex(x) for x in sharded[2:]], "single_table_direct_path": {"kmers": ref[0], "sum_of_counts": ref[1],
                                                                                             "checksum": [hex(x) for x in ref[2:]]},
                      "match": sharded == ref and sharded[1] == parts[8],
                      "how": "kg_checksum of every shard, added; against ONE table filled from all ranks' reads by kg_count_kernel on rank 0"}
            if not verify["match"]:
                print("bench.py: VERIFY MISMATCH " + json.dumps(verify), file=sys.stderr, flush=True)
    trace("verify done")
    if world > 1:
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- file -> output file through the CLI (wall clock; N = 1) ---------------------------------------------------------
    file_to_file = None
    if world == 1 and not args.no_e2e:
        try:
            path = f"/dev/shm/kaarme_bench_{os.getpid()}.fasta"
            with open(path, "wb") as f:
                f.write(fasta.cpu().numpy().tobytes())
            exe = os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "kaarme")
            cmd = [exe, path, str(k), "-m", "0", "-s", str(total_slots), "-a", "2", "-t", str(max(3, min(64, os.cpu_count() or 3))), "-o", path + ".out"]
            t0 = time.perf_counter()
            p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
            wall_cli = time.perf_counter() - t0
            file_to_file = {"wall_s": wall_cli, "kmers_per_s": meta["input_kmers"] / wall_cli, "rc": p.returncode,
                            "output_bytes": os.path.getsize(path + ".out") if os.path.exists(path + ".out") else 0,
                            "input_bytes": int(fasta.numel()),
                            "timers": {ln.split(":")[0]: ln.split(":")[1].strip() for ln in p.stdout.splitlines() if ln.startswith("Time used")},
                            "cmd": "kaarme INPUT 51 -m 0 -s 250000000 -a 2 (FASTA and output on /dev/shm)"}
        except Exception as e:  # noqa: BLE001
            file_to_file = {"error": str(e)}
        finally:
            for q in (path, path + ".out"):
                if os.path.exists(q):
                    os.remove(q)

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------
    # direct path: kg_count_kernel<W,TABLE>; bucketed path (partitions > 1 or N > 1): kg_skm_insert<W,TABLE>.
    # Its own CUDA-event time comes from the library (events around every launch, on the launching stream).
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    W = (k + 31) // 32
    S = -(-(8 * W + 4) // 32)
    rec = meta["record_bytes"]
    bytes_per_kmer = rec / (meta["L"] - k + 1) + 2 * 32 * S      # fused parse->insert figure
    bucketed = st["partitions"] > 1 or world > 1
    kernel = f"kg_skm_insert<{W},TABLE>" if bucketed else f"kg_count_kernel<{W},TABLE>"
    kmers_in_kernel = inserted_rank * args.steps                   # k-mers this rank's kernel launches processed
    achieved = kmers_in_kernel * bytes_per_kmer / (insert_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp)).get(kernel.split("<")[0], {})
        if tj.get("dram_bytes_per_kmer") is not None and tj.get("W") == W:   # the capture was taken at k=51 (W=2)
            traffic = tj["dram_bytes_per_kmer"] * kmers_in_kernel / max(1, insert_launches)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel, "peak_source": peak_src,
                "bytes_per_kmer": bytes_per_kmer, "launches": insert_launches,
                "avg_launch_ms": insert_ms / max(1, insert_launches),
                "algorithmic_bytes_per_launch": kmers_in_kernel * bytes_per_kmer / max(1, insert_launches),
                "kernel_share_of_step": insert_ms / dev_ms,
                # the same algorithmic bytes over the WHOLE step (parse + bucketing + insert of all batches of this rank)
                "step": {"achieved": meta["input_kmers"] * args.steps * bytes_per_kmer / (dev_ms * 1e-3) / 1e9,
                         "frac": meta["input_kmers"] * args.steps * bytes_per_kmer / (dev_ms * 1e-3) / 1e9 / peak},
                "note": "algorithmic bytes = SURVEY 8d figure (ASCII input once + one 32 B sector read and written back per "
                        "k-mer), charged to the dominant kernel (frac) and to the whole step (step.frac); the L2-blocked insert "
                        "keeps the live table region in L2, so its DRAM traffic is BELOW the algorithmic bytes"}
    try:
        kmers_per_s_kernel = kmers_in_kernel / (insert_ms * 1e-3)
        ceil_dram = kg.atomic_ceiling(local_rank, region_bytes=8 << 30, n_ops=1 << 29, reps=2)
        ceil_l2 = kg.atomic_ceiling(local_rank, region_bytes=32 << 20, n_ops=1 << 29, reps=2)
        roofline["atomic"] = {"achieved_sectors_per_s": kmers_per_s_kernel * S,
                              "ceiling_sectors_per_s": ceil_dram, "frac": kmers_per_s_kernel * S / ceil_dram,
                              "ceiling_l2_resident_sectors_per_s": ceil_l2, "frac_of_l2_resident": kmers_per_s_kernel * S / ceil_l2,
                              "how": "uniform-random RED.ADD on 32-byte sectors (kg_atomic_ceiling): over 8 GiB (DRAM-random) "
                                     "and over 32 MiB (L2-resident)"}
    except Exception as e:  # noqa: BLE001
        roofline["atomic"] = {"error": str(e)}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dev_ms / args.steps, "wall_ms_per_step": wall_ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": {"workload": workload_string(dict(bench_data.CONFIGS[args.workload], name=args.workload), k, args.scale, world, args.bloom, args.fpr),
                      "scale": args.scale,
                      "bloom": bloom_info or None, "fasta_bytes_per_gpu": int(fasta.numel()),
                      "table_bytes_per_gpu": ctr.table_info()["slots"] * ctr.table_info()["slot_bytes"],
                      "input_kmers_per_gpu": meta["input_kmers"], "distinct_rank0": distinct,
                      "l2": "inputs (2 GB/GPU) and table (4 GB/GPU) are far larger than the 126 MB L2; no flush between steps",
                      "parallelism": (f"minimizer-sharded x{world}: 8-byte run descriptors read in place over NVLink, packed reads "
                                      f"pulled by the copy engines, one-word NCCL all-reduce per round") if world > 1 else "single GPU",
                      "partitions": st["partitions"], "batch_mb": args.batch_mb},
           "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "e2e": e2e, "verify": verify, "file_to_file": file_to_file,
           "stage_ms_per_step": {"parse": parse_ms / args.steps, "bucket+insert": count_ms / args.steps,
                                 "insert_kernel": insert_ms / args.steps}}
    if world == 1 and not args.no_cpu_baseline:
        try:
            del fasta
            out["cpu_baseline"] = cpu_baseline(args.workload, k, args.scale, torch, bench_data, dev)
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"error": str(e)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_reference(args, torch, bench_data):
    """--impl reference: the reference's own CPU counter on this box's host cores; every step counts the same bounded
    sample of the workload (REF_FRACTION of the genome at the same coverage)."""
    import oracle.oracle_py as o
    k = args.k
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    cores = os.cpu_count() or 1
    threads = max(3, min(64, cores))
    c = dict(bench_data.CONFIGS[args.workload], name=args.workload)
    times, f2f = [], None
    if o.have_ref():
        path, meta = reference_sample(args.workload, k, args.scale, torch, bench_data, dev)
        kmers = meta["input_kmers"]
        try:
            for i in range(args.warmup + args.steps):
                r = cpu_reference_run(path, k, meta["slots"], threads)
                if i >= args.warmup:
                    times.append(r["build_s"])
            w = cpu_reference_run(path, k, meta["slots"], threads, write=True)
            f2f = {"wall_s": w["wall_s"], "kmers_per_s": kmers / w["wall_s"], "output_bytes": w["out_bytes"], "build_s": w["build_s"],
                   "write_s": w.get("write_s")}
        finally:
            os.remove(path)
        kind, used, sample = "reference", threads - 2, sample_string(meta, k) + f"; oracle/_ref/kaarme -t {threads}"
    else:
        fasta, meta = bench_data.make_config(args.workload, dev, scale=args.scale * 0.02)
        data = fasta.cpu().numpy().tobytes()
        for i in range(args.warmup + args.steps):
            secs, kmers = cpu_port_run(data, k)
            if i >= args.warmup:
                times.append(secs)
        kind, used, sample = "port", 1, f"0.02 of the workload ({len(data)} bytes), oracle/liboracle.so ko_count"
    value = kmers * len(times) / sum(times)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": {"workload": workload_string(c, k, args.scale, max(1, args.gpus), args.bloom, args.fpr), "scale": args.scale},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample + " per step", "host_cores": cores},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "file_to_file": f2f}
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
