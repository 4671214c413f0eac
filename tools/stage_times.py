#!/usr/bin/env python
"""Per-stage times of one extraction batch (CUDA events inside the library) + output digest:
python tools/stage_times.py [config] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "eao-fusion_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import eaof
from eaof import workload


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "configs[1]"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    reps = 8
    cfg = workload.CONFIGS[name]
    W, H, nf = cfg["width"], cfg["height"], cfg["nfeatures"]
    fr = workload.Sequence(W, H).frames(0, B)
    d = torch.from_numpy(fr).cuda()
    ex = eaof.ORBextractor(nf, 1.2, 8, 20, 7, width=W, height=H, max_batch=B)
    ex.extract_batch_device(d.data_ptr(), B); ex.sync()
    res = ex.fetch(B)
    dig = workload.combine([workload.frame_digest(*r) for r in res])
    ex.set_profiling(True)
    acc = {}
    for _ in range(reps):
        ex.extract_batch_device(d.data_ptr(), B); ex.sync()
        for k, v in ex.stage_times().items():
            acc.setdefault(k, []).append(v)
    ex.set_profiling(False)
    st = torch.cuda.ExternalStream(ex.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        ex.extract_batch_device(d.data_ptr(), B)
    e1.record(st); ex.sync()
    tot = e0.elapsed_time(e1) / reps
    print(f"{name} B={B}: " + " ".join(f"{k}={np.median(v):.3f}" for k, v in acc.items()) + f" | batch={tot:.3f} ms | digest {dig[:16]}", flush=True)


if __name__ == "__main__":
    main()
