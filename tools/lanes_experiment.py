#!/usr/bin/env python
"""Scheduling experiments on device-resident frames (extraction only):
  lanes n     : one batch split over n handles that run concurrently (own streams)
  alternate n : full batches issued round-robin to n handles (software pipelining across batches)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "eao-fusion_b200")]
import numpy as np, torch, eaof
from eaof import synth
W, H, B = 640, 480, 250
frames = synth.make_frames(500, W, H, tex=synth.base_texture(W, H, seed=1235))
d = torch.from_numpy(frames).cuda()
def mk(n, per): return [eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=W, height=H, max_batch=per) for _ in range(n)]
def timed(exs, step, steps=24):
    for i in range(4): step(i)
    for ex in exs: ex.sync()
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(steps): step(i)
    for ex in exs: ex.sync()
    dt = time.perf_counter() - t
    for ex in exs: ex.close()
    return B * steps / dt
def lanes(n):
    per = B // n
    exs = mk(n, per)
    def step(i):
        for l, ex in enumerate(exs):
            ex.extract_batch_device(d.data_ptr() + ((i % 2) * B + l * per) * W * H, per)
    return timed(exs, step)
def alternate(n):
    exs = mk(n, B)
    def step(i):
        exs[i % n].extract_batch_device(d.data_ptr() + (i % 2) * B * W * H, B)
    return timed(exs, step)
for n in (1, 2):
    print(f"lanes {n}: {lanes(n):9.0f} frames/s")
for n in (1, 2, 3):
    print(f"alternate {n}: {alternate(n):9.0f} frames/s")
