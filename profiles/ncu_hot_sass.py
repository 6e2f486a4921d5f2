#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples: usage ncu_hot_sass.py file.ncu-rep [kernel-index] [topN]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; kernels.append(cur)
    elif r and r[0] == "Address": cur["hdr"] = r
    elif cur is not None and len(r) > 5: cur["rows"].append(r)
k = kernels[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
h = k["hdr"]; si = h.index("# Samples"); ie = h.index("Instructions Executed")
tot = sum(int(r[si]) for r in k["rows"])
print(k["name"][:90], "total samples", tot, "instructions", len(k["rows"]))
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
for idx, r in sorted(enumerate(k["rows"]), key=lambda x: -int(x[1][si]))[:top]:
    st = sorted(((int(r[i]), h[i][6:]) for i in stall_cols if r[i].isdigit()), reverse=True)[:2]
    print(f"{idx:5d} {100*int(r[si])/tot:5.1f}%  exec={r[ie]:>9s}  {r[1].strip()[:70]:70s} {st}")
