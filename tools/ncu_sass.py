#!/usr/bin/env python
"""Hot SASS instructions of kernel #idx in an ncu report: python tools/ncu_sass.py report.ncu-rep idx [min_pct]"""
import csv
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
k = -1
hdr = None
out = []
name = ""
for r in rows:
    if r and r[0] == "Kernel Name":
        k += 1
        if k == idx:
            name = r[1]
        continue
    if k != idx:
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if not hdr or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        n = int(d["Instructions Executed"])
    except ValueError:
        continue
    out.append((n, d["Source"].strip(), d["# Samples"]))
tot = sum(o[0] for o in out) or 1
print(name[:100], "total warp instructions", tot)
for n, sx, sm in out:
    if n > tot * minp / 100:
        print(f"{100 * n / tot:5.2f}% {sm:>5} {sx[:100]}")
