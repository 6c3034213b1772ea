#!/usr/bin/env python3
"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/:
   r01_launches_summary.txt, r01_insert_segs_ncu.txt, r01_count_direct_ncu.txt, traffic.json"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launches():
    rows = [r for r in csv.reader(open(os.path.join(G, "r01_launches.csv"))) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        n = r[ki].split("(")[0][:60]
        agg[n][0] += 1
        agg[n][1] += v
    ours = {n: v for n, v in agg.items() if "kg_" in n and "ceiling" not in n}
    tot = sum(v[1] for v in ours.values())
    with open(os.path.join(P, "r01_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 : python bench.py --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e\n")
        f.write("# C4 shape at 1/4 (497.5 M k-mers per step, 2 batches of 256 MiB), warm-up step + timed step; kg_* kernels only\n")
        f.write("# (torch data-generation kernels and the roofline probe kg_atomic_ceiling_kernel excluded). Cold-cache, serialised: compare SHARES.\n")
        for n, (c, v) in sorted(ours.items(), key=lambda x: -x[1][1]):
            f.write(f"{v:12.1f} us {100 * v / tot:6.2f}%  x{c:4d}  {n}\n")
        f.write(f"{tot:12.1f} us total\n")
    return ours, tot


def raw(rep, out, header, kmers):
    txt = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H = rows[0]
    keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "lts__throughput",
            "l1tex__throughput", "sm__throughput", "sm__warps_active", "launch__registers_per_thread", "launch__grid_size", "lts__t_sector_hit_rate",
            "per_issue_active", "smsp__inst_executed.sum", "lts__t_requests_srcunit_tex_op", "launch__occupancy_limit", "sm__maximum_warps"]
    res = {}
    with open(os.path.join(P, out), "w") as f:
        f.write(header)
        for i, h in enumerate(H):
            if any(s in h for s in keep):
                f.write(h + " | " + " | ".join(r[i] for r in rows[1:]) + "\n")
            if h in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
                res[h] = [(r[i], rows[1][i]) for r in rows[2:]]

    def to_bytes(v, unit):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    rd = sum(to_bytes(v, u) for v, u in res["dram__bytes_read.sum"]) / len(res["dram__bytes_read.sum"])
    wr = sum(to_bytes(v, u) for v, u in res["dram__bytes_write.sum"]) / len(res["dram__bytes_write.sum"])
    return {"W": 2, "k": 51, "dram_bytes_per_launch": rd + wr, "dram_bytes_per_kmer": (rd + wr) / kmers, "kmers_per_launch": kmers,
            "launch_ms": [float(v) for v, _ in res["gpu__time_duration.sum"]]}


if __name__ == "__main__":
    launches()
    kpl = 497_500_000 / 2  # k-mers per launch at --scale 0.25 with 256 MiB batches (2 batches)
    t = {}
    t["kg_insert_segs_kernel"] = raw("r01_insert_segs.ncu-rep", "r01_insert_segs_ncu.txt",
                                     "# ncu --set full --clock-control none --import-source on -k regex:kg_insert_segs_kernel -s 2 -c 2 ; python bench.py --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e\n"
                                     "# L2-blocked insert through the segment table (partitions auto = 64, table 1 GB of packed 16-byte slots), 248.75 M k-mers per launch\n", kpl)
    t["kg_count_kernel"] = raw("r01_count_direct.ncu-rep", "r01_count_direct_ncu.txt",
                               "# ncu --set full ... -k regex:kg_count_kernel -s 2 -c 2 ; python bench.py --scale 0.25 ... --partitions 1   (direct, DRAM-random insert)\n", kpl)
    json.dump(t, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print(json.dumps(t, indent=1))
