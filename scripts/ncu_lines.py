#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by source line.

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K > mix.csv
    python scripts/ncu_lines.py mix.csv [top]
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
agg = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    if r[0] != "" and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0
        agg.append((cur_file, int(r[0]), r[1].strip()[:110], num("# Samples"), num("Instructions Executed"),
                    {k: num(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k},
                    num("L1 Wavefronts Shared Excessive")))
tot = sum(a[3] for a in agg) or 1.0
tot_inst = sum(a[4] for a in agg) or 1.0
print(f"total samples {tot:.0f}, total warp instructions {tot_inst:.0f}")
agg.sort(key=lambda a: -a[3])
for f, ln, src, smp, inst, st, exc in agg[:top]:
    reasons = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    rs = " ".join(f"{k[6:]}={v:.0f}" for k, v in reasons if v > 0)
    print(f"{100*smp/tot:5.1f}% smp {100*inst/tot_inst:5.1f}% inst  {f}:{ln:<4d} {src}\n        [{rs}] excess_smem_wf={exc:.0f}")
