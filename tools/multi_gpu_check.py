#!/usr/bin/env python
"""Multi-GPU correctness of the sharded paths; run under torchrun on N GPUs of one node:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
      tools/multi_gpu_check.py

Each rank extracts its contiguous frame block (no collective), the descriptor blocks are all-gathered over NCCL and
the cross-frame brute-force pair list is partitioned round-robin (eaof/sweep.py).  Rank 0 repeats everything alone and
requires identical keypoints, descriptors and matches; a few pairs are also checked against the CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "eao-fusion_b200")):
    sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

import eaof
from eaof import shard, sweep, synth


def blocks_of(ex, n):
    """(desc [n,cap,32], angle [n,cap], counts [n]) torch views/copies of an extractor's device results."""
    k, d, c, cap = ex.device_results()
    kps = torch.as_tensor(sweep.DevicePtr(k, (ex.max_batch, cap, 6), "<f4"), device="cuda")[:n]
    desc = torch.as_tensor(sweep.DevicePtr(d, (ex.max_batch, cap, 32), "|u1"), device="cuda")[:n]
    cnt = torch.as_tensor(sweep.DevicePtr(c, (ex.max_batch,), "<i4"), device="cuda")[:n]
    return desc.clone(), kps[:, :, 3].contiguous(), cnt.clone()


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, NF, n_frames = 640, 480, 1000, 13
    frames = synth.make_frames(n_frames, W, H, tex=synth.base_texture(W, H, seed=99))
    b, e = shard.frame_block(n_frames, rank, world)
    ex = eaof.ORBextractor(NF, 1.2, 8, 20, 7, width=W, height=H, max_batch=n_frames, device=local)
    d_fr = torch.from_numpy(frames[b:e]).cuda()
    ex.extract_batch_device(d_fr.data_ptr(), e - b)
    ex.sync()
    desc, ang, cnt = blocks_of(ex, e - b)
    pairs = np.array([(i, j) for i in range(n_frames) for j in range(n_frames) if i != j and (i + 2 * j) % 3 == 0], np.int32)
    mt = eaof.ORBmatcher(0.9, True, max_features=ex.cap, max_pairs=64, device=local)
    sel, m, dd, nm = sweep.sweep(mt, eaof.BOW_KF_FRAME, desc, ang, cnt, pairs, n_frames, rank, world, dist if world > 1 else None)
    # collect everything on rank 0
    stride = ex.cap
    full_m = torch.full((len(pairs), stride), -2, dtype=torch.int32, device="cuda")
    full_n = torch.full((len(pairs),), -2, dtype=torch.int32, device="cuda")
    full_m[torch.from_numpy(sel).cuda()] = m
    full_n[torch.from_numpy(sel).cuda()] = nm
    if world > 1:
        dist.all_reduce(full_m, op=dist.ReduceOp.MAX)
        dist.all_reduce(full_n, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        ex1 = eaof.ORBextractor(NF, 1.2, 8, 20, 7, width=W, height=H, max_batch=n_frames, device=local)
        d_all = torch.from_numpy(frames).cuda()
        ex1.extract_batch_device(d_all.data_ptr(), n_frames)
        ex1.sync()
        desc1, ang1, cnt1 = blocks_of(ex1, n_frames)
        sel1, m1, dd1, nm1 = sweep.sweep(mt, eaof.BOW_KF_FRAME, desc1, ang1, cnt1, pairs, n_frames)
        assert (sel1 == np.arange(len(pairs))).all()
        ok = bool(torch.equal(full_m, m1)) and bool(torch.equal(full_n, nm1))
        # oracle spot checks
        from oracle import pyoracle as po
        hd, ha, hc = desc1.cpu().numpy(), ang1.cpu().numpy(), cnt1.cpu().numpy()
        for pi in (0, len(pairs) // 2, len(pairs) - 1):
            q, t = pairs[pi]
            nq, nt = int(hc[q]), int(hc[t])
            nodes_q, nodes_t = eaof.csr_from_nodes(np.zeros(nq, int)), eaof.csr_from_nodes(np.zeros(nt, int))
            on, om, _ = po.o_search_by_bow(0, 0.9, True, hd[q, :nq], ha[q, :nq], None, nodes_q, hd[t, :nt], ha[t, :nt], None, nodes_t)
            ok = ok and on == int(full_n[pi]) and np.array_equal(om, full_m[pi, :nt].cpu().numpy())
        print(f"MULTI_GPU_CHECK world={world} pairs={len(pairs)} matches/pair={float(full_n.float().mean()):.1f} "
              f"{'OK' if ok else 'MISMATCH'}", flush=True)
        ex1.close()
    mt.close()
    ex.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
