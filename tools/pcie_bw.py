#!/usr/bin/env python
"""Pinned host<->device copy bandwidth of this box (the ceiling of bench.py's e2e leg): python tools/pcie_bw.py"""
import torch

def bw(nbytes, h2d, reps=10):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        (d.copy_(h, non_blocking=True) if h2d else h.copy_(d, non_blocking=True))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        (d.copy_(h, non_blocking=True) if h2d else h.copy_(d, non_blocking=True))
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

for mb in (4, 17, 77, 256):
    n = mb << 20
    print(f"{mb:4d} MiB  H2D {bw(n, True):6.1f} GB/s   D2H {bw(n, False):6.1f} GB/s")
# both directions at once
n = 77 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print(f"bidirectional 77 MiB each way: {2 * n * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9:6.1f} GB/s total")
