#!/usr/bin/env python
"""Aggregate pinned host->device / device->host bandwidth with ALL ranks copying at the same time (the ceiling of bench.py's e2e
leg at N GPUs: every rank uploads 77 MB of frames and downloads 15 MB of results per 250-frame batch).  Run under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29561 tools/pcie_bw_concurrent.py

Prints one line with per-rank and aggregate GB/s for H2D alone, D2H alone, and both directions in bench's 5 : 1 byte ratio."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    up, down = 77 << 20, 15 << 20
    hu = torch.empty(up, dtype=torch.uint8).pin_memory(); du = torch.empty(up, dtype=torch.uint8, device="cuda")
    hd = torch.empty(down, dtype=torch.uint8).pin_memory(); dd = torch.empty(down, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(do_up, do_down, reps=20):
        for _ in range(2):
            if do_up: du.copy_(hu, non_blocking=True)
            if do_down: hd.copy_(dd, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if do_up:
                with torch.cuda.stream(s1): du.copy_(hu, non_blocking=True)
            if do_down:
                with torch.cuda.stream(s2): hd.copy_(dd, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), reps

    out = {"world": world, "bytes_up_per_copy": up, "bytes_down_per_copy": down}
    dt, reps = run(True, False)
    out["h2d_gbs_per_rank"] = up * reps / dt / 1e9
    out["h2d_gbs_aggregate"] = world * up * reps / dt / 1e9
    dt, reps = run(False, True)
    out["d2h_gbs_aggregate"] = world * down * reps / dt / 1e9
    dt, reps = run(True, True)
    out["both_gbs_aggregate"] = world * (up + down) * reps / dt / 1e9
    out["e2e_frames_per_s_ceiling"] = world * 250 * reps / dt  # one (77 MB up, 15 MB down) round = one 250-frame batch
    if rank == 0:
        print("PCIE_CONCURRENT " + json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
