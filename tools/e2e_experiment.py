#!/usr/bin/env python
"""Where does the end-to-end (host buffers) pipeline lose time against the device-resident path?  Extraction only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "eao-fusion_b200")]
import numpy as np, torch, eaof
from eaof import synth
W, H, B = 640, 480, 250
frames = synth.make_frames(1000, W, H, tex=synth.base_texture(W, H, seed=1235))
h_all = torch.from_numpy(frames).pin_memory()
d_all = torch.from_numpy(frames).cuda()
def mk(chunk):
    ex = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=W, height=H, max_batch=B)
    ex.set_pipeline_chunk(chunk)
    return ex
def bufs(ex):
    return torch.empty((B, ex.cap, 6), dtype=torch.float32).pin_memory(), torch.empty((B, ex.cap, 32), dtype=torch.uint8).pin_memory()
def run(n_handles, chunk, steps=24, download=True):
    exs = [mk(chunk) for _ in range(n_handles)]
    bb = [bufs(ex) for ex in exs]
    def issue(k):
        s = k % n_handles
        exs[s].extract_batch_async(h_all.data_ptr() + (k % 4) * B * W * H, B, bb[s][0].data_ptr() if download else 0,
                                   bb[s][1].data_ptr() if download else 0)
    def finish(k): exs[k % n_handles].extract_batch_wait()
    for k in range(n_handles): issue(k)
    for k in range(n_handles): finish(k)
    torch.cuda.synchronize(); t = time.perf_counter()
    for k in range(min(n_handles, steps)): issue(k)
    for k in range(steps):
        finish(k)
        if k + n_handles < steps: issue(k + n_handles)
    dt = time.perf_counter() - t
    for ex in exs: ex.close()
    return B * steps / dt
def device_only(steps=24):
    ex = mk(B)
    for i in range(3): ex.extract_batch_device(d_all.data_ptr() + (i % 4) * B * W * H, B)
    ex.sync(); t = time.perf_counter()
    for i in range(steps): ex.extract_batch_device(d_all.data_ptr() + (i % 4) * B * W * H, B)
    ex.sync(); dt = time.perf_counter() - t
    ex.close()
    return B * steps / dt
print(f"device-resident, 1 handle            : {device_only():9.0f} frames/s")
for nh, ch, dl in ((1, B, True), (1, 55, True), (2, B, True), (2, 55, True), (3, B, True), (2, B, False), (2, 125, True)):
    print(f"host buffers, {nh} handle(s), chunk {ch:3d}, download={dl}: {run(nh, ch, download=dl):9.0f} frames/s")
