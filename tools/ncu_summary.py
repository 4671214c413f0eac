#!/usr/bin/env python
"""Headline numbers + stall breakdown of the first kernel in an ncu report: python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, r = rows[0], rows[2 + idx]
d = dict(zip(h, r))
print(d.get("Kernel Name"))
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for k in keys:
    if k in d:
        print(f"  {k:72s} {d[k]}")
st = {k: float(v.replace(",", "")) for k, v in d.items() if "smsp__pcsamp_warps_issue_stalled" in k and v}
tot = sum(st.values()) or 1
print("  stall samples:", ", ".join(f"{k[33:]} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
