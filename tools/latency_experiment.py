#!/usr/bin/env python
"""Single-frame latency of ORBextractor::operator() through the host-buffer C ABI (what the drop-in class calls), and of the
Tracking-shaped loop: extract one frame, then SearchByProjection against the previous one (src/Tracking.cc:1717-1763)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "eao-fusion_b200")):
    sys.path.insert(0, p)
import numpy as np

import eaof
from eaof import workload


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(q * len(v)))] * 1e6


def main():
    W, H, NF = 640, 480, 1000
    fr = workload.Sequence(W, H).frames(0, 64)
    ex = eaof.ORBextractor(NF, 1.2, 8, 20, 7, width=W, height=H, max_batch=2)
    for i in range(10):
        ex(fr[i])
    lat = []
    for i in range(300):
        t0 = time.perf_counter()
        k, d = ex(fr[i % 64])
        lat.append(time.perf_counter() - t0)
    print(f"graph={'off' if os.environ.get('EAOF_NO_GRAPH') else 'on'} operator(): p50 {pct(lat, .5):.1f} us  p5 {pct(lat, .05):.1f}  p95 {pct(lat, .95):.1f}  "
          f"keypoints {len(k)} launches {ex.last_launch_count()}", flush=True)
    # Tracking-shaped loop (src/Tracking.cc:1717-1763): extract frame t, then SearchByProjection(Cur = t, Last = t-1) through the
    # host-buffer single-pair call the drop-in ORBmatcher uses (uploads both frames' arrays, 3 kernels, downloads the matches)
    mt = eaof.ORBmatcher(0.9, True, max_features=4096)
    sf = ex2_tables = None
    ex = eaof.ORBextractor(NF, 1.2, 8, 20, 7, width=W, height=H, max_batch=2)
    sf = ex.GetScaleFactors()
    bounds = (0.0, float(W), 0.0, float(H))
    ginv = (np.float32(64) / np.float32(W), np.float32(48) / np.float32(H))
    prev = ex(fr[0])
    tm, tt = [], []
    for i in range(1, 301):
        t0 = time.perf_counter()
        k, d = ex(fr[i % 64])
        t1 = time.perf_counter()
        lk, ld = prev
        cur = dict(x=k["x"], y=k["y"], octave=k["octave"], angle=k["angle"], desc=d)
        last = dict(u=lk["x"] - np.float32(2), v=lk["y"] - np.float32(1), octave=lk["octave"], angle=lk["angle"], desc=ld)
        n, m, dd = mt.SearchByProjection(cur, last, 15.0, bounds=bounds, grid_inv=ginv, scale_factors=sf)
        t2 = time.perf_counter()
        prev = (k, d)
        if i > 10:
            tm.append(t2 - t1)
            tt.append(t2 - t0)
    print(f"tracking loop: SearchByProjection(Cur,Last) p50 {pct(tm, .5):.1f} us p95 {pct(tm, .95):.1f}; extract+match p50 {pct(tt, .5):.1f} us "
          f"p95 {pct(tt, .95):.1f}; matches {n}", flush=True)
    mt.close()
    ex.close()


if __name__ == "__main__":
    main()
