#!/usr/bin/env python
"""Executed instructions and stall samples of one kernel, split at its BAR.SYNC instructions:
usage ncu_phases.py file.ncu-rep [kernel-index | kernel-name-substring]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; kernels.append(cur)
    elif r and r[0] == "Address": cur["hdr"] = r
    elif cur is not None and len(r) > 5: cur["rows"].append(r)
sel = sys.argv[2] if len(sys.argv) > 2 else "0"
k = kernels[int(sel)] if sel.isdigit() else next(x for x in kernels if sel in x["name"])
h, R = k["hdr"], k["rows"]
ie, si = h.index("Instructions Executed"), h.index("# Samples")
tot, stot = sum(int(r[ie]) for r in R), sum(int(r[si]) for r in R)
print(k["name"][:80], "warp-instructions", tot, "samples", stot)
acc = sacc = start = 0
for i, r in enumerate(R):
    acc += int(r[ie]); sacc += int(r[si])
    if "BAR.SYNC" in r[1] or i == len(R) - 1:
        print(f"[{start:4d}-{i:4d}] exec {acc/1e9:6.3f} G ({100*acc/tot:4.1f}%)  samples {100*sacc/stot:4.1f}%")
        acc = sacc = 0; start = i + 1
