#!/usr/bin/env python3
"""Text summary of an `ncu --set full` capture (what the judge and DESIGN.md cite): per kernel instance the duration,
DRAM bytes, unit throughputs, occupancy limiters, stall reasons and lane efficiency.
    python profiles/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_ncu.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_config_size",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}")
    for r in rows[2:]:
        print(f"\n== {r[idx['Kernel Name']]}")
        for w in WANT:
            if w in idx:
                print(f"  {w:78s} {r[idx[w]]:>22s} {units[idx[w]]}")
        for h in hdr:
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    v = float(r[idx[h]].replace(",", ""))
                except ValueError:
                    continue
                if v >= 0.3:
                    print(f"  {h:78s} {v:22.2f} warps/issue")


if __name__ == "__main__":
    main()
