#!/usr/bin/env python
"""Stage times of one build variant (EAOF_LIB_PATH) on the configs[1] batch, k_fast staged by LDG (EAOF_FAST_TMA=0) against
the TMA-staged persistent kernel, with a digest comparison of the two paths' outputs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "eao-fusion_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import eaof
from eaof import workload


def run(tma, d, B, W, H, nf, reps=6):
    os.environ["EAOF_FAST_TMA"] = str(int(tma))
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
    # whole batch, unprofiled (blur on the side stream)
    st = torch.cuda.ExternalStream(ex.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        ex.extract_batch_device(d.data_ptr(), B)
    e1.record(st); ex.sync()
    tot = e0.elapsed_time(e1) / reps
    del st, e0, e1
    ex.close()
    return dig, {k: float(np.median(v)) for k, v in acc.items()}, tot


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "configs[1]"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    cfg = workload.CONFIGS[name]
    W, H, nf = cfg["width"], cfg["height"], cfg["nfeatures"]
    fr = workload.Sequence(W, H).frames(0, B)
    d = torch.from_numpy(fr).cuda()
    d0, t0, w0 = run(0, d, B, W, H, nf)
    d1, t1, w1 = run(1, d, B, W, H, nf)
    d2, t2, w2 = run(2, d, B, W, H, nf)
    if os.environ.get("EAOF_EXPERIMENT") == "pyramid":
        os.environ["EAOF_PYR_FUSED"] = "0"
        dp, tp, wp = run(0, d, B, W, H, nf)
        os.environ["EAOF_PYR_FUSED"] = "2"
        d0, t0, w0 = run(0, d, B, W, H, nf)
        os.environ["EAOF_PYR_FUSED"] = "1"
        print(f"{name} B={B}: per-level pyramid={tp['pyramid']:.3f} batch={wp:.3f} | fused pyramid={t0['pyramid']:.3f} batch={w0:.3f} ms | "
              f"same output: {dp == d0}", flush=True)
        return
    print(f"{os.path.basename(os.environ.get('EAOF_LIB_PATH', 'default'))} {name} B={B}: ldg fast={t0['fast']:.3f} batch={w0:.3f} | "
          f"tma-persistent fast={t1['fast']:.3f} batch={w1:.3f} | tma-oneshot fast={t2['fast']:.3f} batch={w2:.3f} ms | "
          f"same output: {d0 == d1 == d2}", flush=True)


if __name__ == "__main__":
    main()
