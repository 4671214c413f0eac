#!/usr/bin/env python
"""Opcode histogram of every kernel in libeaof_orb.so (cuobjdump -sass), with the Blackwell tells called out:
UTCIMMA (tcgen05.mma kind::i8), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UTMALDG (cp.async.bulk.tensor), UBLKCP
(cp.async.bulk), SYNCS (mbarrier), VIMNMX3 / VABSDIFF4 / IDP (DPX and byte-SIMD).
python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "eao-fusion_b200", "lib", "libeaof_orb.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
        depth, cut = 0, len(dem)
        for i, ch in enumerate(dem):  # argument list = first "(" outside template brackets
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        cur = dem[:cut].replace("void ", "")
        per[cur] = per.get(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        per[cur][m.group(1).split(".")[0]] += 1
tells = ["UTCIMMA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "SYNCS", "VIMNMX3", "VABSDIFF4", "IDP", "POPC", "MATCH", "SHFL", "VOTE"]
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} (sm_100a): static instruction counts per kernel")
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{t:>9s}" for t in tells))
for k, c in per.items():
    print(f"{k[:58]:58s} {sum(c.values()):6d} " + " ".join(f"{c.get(t, 0):9d}" for t in tells))
print()
for k, c in per.items():
    print(k)
    print("   " + ", ".join(f"{o} {n}" for o, n in c.most_common(14)))
