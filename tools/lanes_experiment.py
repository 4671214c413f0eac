#!/usr/bin/env python
"""Does running two half-batches on two handles (own streams) beat one full batch?  Device-resident frames."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "eao-fusion_b200")]
import numpy as np, torch, eaof
from eaof import synth
W, H, B = 640, 480, 250
frames = synth.make_frames(500, W, H, tex=synth.base_texture(W, H, seed=1235))
d = torch.from_numpy(frames).cuda()
def run(n_lanes, steps=20):
    per = B // n_lanes
    exs = [eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=W, height=H, max_batch=per) for _ in range(n_lanes)]
    def step(i):
        for l, ex in enumerate(exs):
            ex.extract_batch_device(d.data_ptr() + ((i % 2) * B + l * per) * W * H, per)
    for i in range(3): step(i)
    for ex in exs: ex.sync()
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(steps): step(i)
    for ex in exs: ex.sync()
    dt = time.perf_counter() - t
    for ex in exs: ex.close()
    return per * n_lanes * steps / dt
for n in (1, 2, 5):
    print(f"lanes {n}: {run(n):9.0f} frames/s (extraction only)")
