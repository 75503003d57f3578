#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (raw page) into the handful of counters the roofline
report uses.  Usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [row]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum",
        "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size"]


def main():
    rep = sys.argv[1]
    row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + row]
    d = dict(zip(hdr, zip(units, vals)))
    print("kernel:", d.get("Kernel Name", ("", "?"))[1])
    for k in KEYS:
        if k in d:
            print(f"{k:75s} {d[k][1]:>16s} {d[k][0]}")
    stalls = []
    for h, (u, v) in d.items():
        if "issue_stalled" in h and h.endswith("per_warp_active.pct"):
            try:
                stalls.append((float(v), h.split("issue_stalled_")[1].replace("_per_warp_active.pct", "")))
            except ValueError:
                pass
    print("top stalls (% of warp-active cycles):", ", ".join(f"{n} {v:.1f}" for v, n in sorted(stalls, reverse=True)[:6]))


if __name__ == "__main__":
    main()
