#!/usr/bin/env python
"""Turn the raw captures of a round (gpurun_out/) into the committed summaries under profiles/:

    python profiles/make_round_profiles.py r02 gpurun_out/<full>.ncu-rep gpurun_out/<launches>.csv

* <round>_ncu_full_1024cube.txt   ncu --set full of passes A (x2), B, C, D at full size: durations, DRAM bytes, pipe utilisation, stall mix,
                                  instruction phases of pass C between its block barriers
* <round>_launch_shares.txt       the launch list (ncu --metrics gpu__time_duration.sum) of the last timed step reduced to per-kernel shares
* traffic.json                    DRAM bytes per kernel and step (read by bench.py for roofline.traffic)
* <round>_sass_tma.txt            the TMA / mbarrier instructions in the SASS of the hot kernels of libmetro_b200.so
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def main():
    rnd, rep, launches = sys.argv[1], sys.argv[2], sys.argv[3]
    # ---- full capture
    out = [f"ncu --set full --clock-control none --import-source on, bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline ({os.path.basename(rep)})\n"]
    out.append(run([sys.executable, os.path.join(HERE, "ncu_summary.py"), rep]))
    raw = list(csv.reader(run(["ncu", "-i", rep, "--page", "raw", "--csv"]).splitlines()))
    hdr = raw[0]
    names = [r[hdr.index("Kernel Name")] for r in raw[2:]]
    traffic = {"note": f"dram__bytes_read.sum + dram__bytes_write.sum per kernel and step, ncu --set full, 1024^3 pair; profiles/{rnd}_ncu_full_1024cube.txt"}
    acc = {"pass_a": 0.0, "pass_b": 0.0, "pass_c": 0.0, "pass_d": 0.0}
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for r, n in zip(raw[2:], names):
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[hdr.index(k)].replace(",", "")) * unit[raw[1][hdr.index(k)]]
        key = "pass_a" if "part_kernel<512, 1>" in n or "part_kernel<1024, 1>" in n or "part_kernel<256, 1>" in n else \
            "pass_b" if "part_kernel" in n else "pass_c" if "bucket_join" in n else "pass_d" if "hit_reduce" in n else None
        if key:
            acc[key] += b
    for k, v in acc.items():
        traffic[k + "_dram_bytes_per_step"] = v
    old = json.load(open(os.path.join(HERE, "traffic.json")))
    traffic["onesweep_dram_bytes_per_key"] = old.get("onesweep_dram_bytes_per_key")
    json.dump(traffic, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    for i, n in enumerate(names):
        if "bucket_join" in n or "hit_reduce" in n:
            out.append("\n-- instructions and stall samples between block barriers --\n")
            out.append(run([sys.executable, os.path.join(HERE, "ncu_phases.py"), rep, str(i)]))
            out.append("\n-- hottest SASS instructions --\n")
            out.append(run([sys.executable, os.path.join(HERE, "ncu_hot_sass.py"), rep, str(i), "16"]))
    open(os.path.join(HERE, f"{rnd}_ncu_full_1024cube.txt"), "w").write("".join(out))
    # ---- launch shares of the last timed step
    rows = list(csv.reader(open(launches)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, R = rows[hi], rows[hi + 1:]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    ends = [i for i, r in enumerate(R) if "tree_off_end_kernel" in r[kn] or "set_i64_kernel" in r[kn]]
    step = R[ends[-2] + 1: ends[-1] + 1]
    agg, tot = OrderedDict(), 0.0
    for r in step:
        n = r[kn].split("(")[0][:52]
        v = float(r[mv].replace(",", ""))
        agg.setdefault(n, [0.0, 0])
        agg[n][0] += v; agg[n][1] += 1; tot += v
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline; last timed step "
             f"({len(step)} launches, {tot / 1e6:.3f} ms serialised)\n"]
    for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"{n:54s} {v / 1e6:8.3f} ms  x{c:<3d} {100 * v / tot:5.1f}%\n")
    open(os.path.join(HERE, f"{rnd}_launch_shares.txt"), "w").write("".join(lines))
    # ---- TMA / mbarrier evidence in the SASS of the product library
    so = os.path.join(ROOT, "metrocpp_b200", "libmetro_b200.so")
    sass = run(["cuobjdump", "-sass", so])
    cur, found = None, OrderedDict()
    for ln in sass.splitlines():
        if "Function :" in ln:
            cur = ln.split("Function :")[1].strip()
        elif cur and any(m in ln for m in ("UBLKCP", "UBLKPF", "SYNCS", "UTMA")):
            ins = ln.split("*/")[1].strip().split(";")[0] if "*/" in ln else ln.strip()
            found.setdefault(cur, OrderedDict())
            key = ins.split()[0]
            found[cur][key] = found[cur].get(key, 0) + 1
    with open(os.path.join(HERE, f"{rnd}_sass_tma.txt"), "w") as f:
        f.write("cuobjdump -sass metrocpp_b200/libmetro_b200.so: TMA (UBLKCP = cp.async.bulk, UBLKPF = cp.async.bulk.prefetch.L2) and mbarrier (SYNCS.*) "
                "instructions per kernel\n")
        for fn, d in found.items():
            dem = run(["c++filt", fn]).strip()[:110]
            f.write(f"{dem}\n    " + ", ".join(f"{k} x{v}" for k, v in d.items()) + "\n")
    print("written", rnd)


if __name__ == "__main__":
    main()
