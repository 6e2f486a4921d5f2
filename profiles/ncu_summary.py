#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the few numbers we track: usage  ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:100])
    for k in KEYS:
        if k in hdr:
            print(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = []
    for k, v in zip(hdr, r):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
            try: st.append((float(v.replace(",", "")), k.split("issue_stalled_")[1].split("_per_issue")[0]))
            except ValueError: pass
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
