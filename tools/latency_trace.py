#!/usr/bin/env python
"""Where the time of one ORBextractor::operator() call goes (host timestamps inside the library, eaof_debug_latency_trace)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "eao-fusion_b200")):
    sys.path.insert(0, p)
import numpy as np

import eaof
from eaof import workload


def main():
    W, H, NF = 640, 480, 1000
    fr = workload.Sequence(W, H).frames(0, 64)
    ex = eaof.ORBextractor(NF, 1.2, 8, 20, 7, width=W, height=H, max_batch=2)
    L = eaof.lib()
    L.eaof_debug_latency_trace.restype = C.c_int
    L.eaof_debug_latency_trace.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    for i in range(20):
        ex(fr[i])
    rows, tot = [], []
    buf = (C.c_double * 6)()
    for i in range(300):
        t0 = time.perf_counter()
        ex(fr[i % 64])
        tot.append(time.perf_counter() - t0)
        L.eaof_debug_latency_trace(ex.h, buf)
        rows.append([buf[j + 1] - buf[j] for j in range(5)])
    r = np.median(np.array(rows), axis=0) * 1e6
    print(f"call p50 {np.median(tot) * 1e6:.1f} us | upload queued {r[0]:.1f} | graph launch {r[1]:.1f} | python between async and wait {r[2]:.1f} | "
          f"stream sync {r[3]:.1f} | copy out {r[4]:.1f} | inside library {r.sum():.1f}", flush=True)
    ex.close()


if __name__ == "__main__":
    main()
