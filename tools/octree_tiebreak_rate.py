#!/usr/bin/env python
"""SURVEY.md Appendix C-1, measured: how often the stock reference (glibc malloc: equal-size quadtree nodes ordered by heap
address) differs from the canonical creation-order rule the CUDA path, the restatement and oracle/_ref's arena implement.
Runs the unmodified reference ORBextractor.cc both ways over frames of the configs[1] sequence (CPU only; needs oracle/_ref).
python tools/octree_tiebreak_rate.py [n_frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "eao-fusion_b200")):
    sys.path.insert(0, p)
import numpy as np

from eaof import workload
from oracle import pyoracle as po

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
frames = workload.Sequence(640, 480).frames(0, n)
canon, stock = po.RefExtractor(canonical=True), po.RefExtractor(canonical=False)
frames_diff = kp_total = kp_diff = 0
for f in frames:
    ck, cd = canon.extract(f)
    sk, sd = stock.extract(f)
    kp_total += len(ck)
    if len(ck) != len(sk) or not np.array_equal(ck, sk) or not np.array_equal(cd, sd):
        frames_diff += 1
        a = {(float(k["x"]), float(k["y"]), int(k["octave"])) for k in ck}
        b = {(float(k["x"]), float(k["y"]), int(k["octave"])) for k in sk}
        kp_diff += len(a ^ b)
print(f"stock glibc vs canonical tie-break over {n} frames of configs[1]: {frames_diff} frames differ "
      f"({100.0 * frames_diff / n:.1f} %), {kp_diff} keypoints in the symmetric difference of {kp_total} "
      f"({100.0 * kp_diff / max(kp_total, 1):.3f} %)")
