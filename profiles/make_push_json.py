#!/usr/bin/env python
"""Extract the numbers bench.py quotes from an `ncu --set full` capture of the push kernel into
profiles/push_kernel_ncu.json.  Usage: python profiles/make_push_json.py REP SUMMARY_TXT_NAME [WORKLOAD]"""
import csv
import json
import os
import subprocess
import sys

rep, src = sys.argv[1], sys.argv[2]
workload = sys.argv[3] if len(sys.argv) > 3 else "thermal_2048x256_m2_ppc64"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
d = dict(zip(rows[0], zip(rows[1], rows[2])))


def val(k):
    u, v = d[k]
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


j = {
    "workload": workload,
    "kernel": "k_push_v2<2,true> (strip CTAs + DMMA deposit + fused particle_bcs)",
    "source": f"profiles/{src} (ncu --set full --clock-control none, one launch)",
    "dram_bytes_read": val("dram__bytes_read.sum"),
    "dram_bytes_write": val("dram__bytes_write.sum"),
    "lsu_data_pipe_pct": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "shared_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "fp64_pipe_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "dmma_pipe_pct": val("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": val("launch__registers_per_thread"),
    "shared_atomics": val("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum"),
    "global_red_instructions": val("smsp__inst_executed_op_global_red.sum"),
    "warp_instructions": val("smsp__inst_executed.sum"),
    "gpu_time_ms_under_ncu": val("gpu__time_duration.sum") * (1e-6 if d["gpu__time_duration.sum"][0] == "ns" else
                                                              1e-3 if d["gpu__time_duration.sum"][0] == "us" else 1.0),
}
j["dram_bytes_per_launch"] = j["dram_bytes_read"] + j["dram_bytes_write"]
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "push_kernel_ncu.json")
json.dump(j, open(path, "w"), indent=1)
print(json.dumps(j, indent=1))
