#!/usr/bin/env python
"""Per-source-line opcode histogram of the kernel in an ncu report (needs --import-source on):
python tools/ncu_ops.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
cur = None
per = {}
ops = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    if r[2] == "-":
        cur = (r[0], r[1].strip()[:80])
        continue
    try:
        n = int(d["Instructions Executed"])
    except ValueError:
        continue
    toks = d.get("Source", "").split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    per.setdefault(cur, {})
    per[cur][op] = per[cur].get(op, 0) + n
    o = op.split(".")[0]
    ops[o] = ops.get(o, 0) + n
tot = sum(ops.values()) or 1
print("total warp instructions", tot)
print("opcodes:", ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in sorted(ops.items(), key=lambda x: -x[1])[:18]))
for k, v in sorted(per.items(), key=lambda x: -sum(x[1].values()))[:top]:
    s = sum(v.values())
    print(f"{k[0]:>5} {100 * s / tot:5.1f}% {k[1]}")
    print("       ", ", ".join(f"{o}:{100 * c / tot:.1f}" for o, c in sorted(v.items(), key=lambda x: -x[1])[:8]))
