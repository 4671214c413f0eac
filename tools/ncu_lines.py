#!/usr/bin/env python
"""Per-source-line summary of an ncu report: python tools/ncu_lines.py report.ncu-rep [min_pct]
Reads `ncu -i REP --page source --csv --print-source cuda,sass` and prints, per CUDA source line, the share of
warp instructions executed and of stall samples (lines whose share is below min_pct are folded)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = None
    per = {}
    fname = ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or len(r) < len(hdr) or r[2] != "-":
            continue  # only the per-CUDA-line aggregate rows (SASS rows carry an address)
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr, r))
        key = (fname, ln)
        e = per.setdefault(key, [0, 0, r[1]])
        e[0] += int(d["Instructions Executed"] or 0)
        e[1] += int(d["# Samples"] or 0)
    ti = sum(v[0] for v in per.values()) or 1
    ts = sum(v[1] for v in per.values()) or 1
    print(f"total warp instructions {ti}, samples {ts}")
    for (f, ln), (ni, ns, src) in sorted(per.items()):
        if 100 * ni / ti >= min_pct or 100 * ns / ts >= min_pct:
            print(f"{f}:{ln:5d} inst {100 * ni / ti:5.1f}%  samples {100 * ns / ts:5.1f}%  {src.strip()[:100]}")


if __name__ == "__main__":
    main()
