#!/usr/bin/env python
"""Multi-GPU correctness of the sharded paths; run under torchrun on N GPUs of one node:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
      tools/multi_gpu_check.py

Each rank extracts its contiguous frame block plus one halo frame (no collective) and matches the consecutive pairs it
owns; the descriptor blocks are all-gathered by ncclAllGather inside libeaof_orb.so (include/eaof_sweep.h; the host only
carries the 128-byte NCCL id) and the cross-frame brute-force pair list is partitioned round-robin (eaof/sweep.py).
Rank 0 repeats everything alone and requires identical keypoints, descriptors and matches; a few pairs are also
checked against the CPU oracle."""
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
    hb, he = shard.halo_block(b, e)
    ex = eaof.ORBextractor(NF, 1.2, 8, 20, 7, width=W, height=H, max_batch=n_frames, device=local)
    mt = eaof.ORBmatcher(0.9, True, max_features=ex.cap, max_pairs=64, device=local)
    # consecutive-frame SearchByProjection over the pairs this rank owns: block + halo frame, no collective
    d_fr = torch.from_numpy(frames[hb:he]).cuda()
    ex.extract_batch_device(d_fr.data_ptr(), he - hb)
    own = shard.consecutive_pairs(b, e, n_frames)
    proj = torch.full((n_frames, ex.cap), -3, dtype=torch.int32, device="cuda")  # row f = matches of pair (f-1, f)
    if own:
        pl = np.array([a - hb for a, _ in own], np.int32)
        pc = np.array([c - hb for _, c in own], np.int32)
        pm = torch.full((len(own), ex.cap), -1, dtype=torch.int32, device="cuda")  # rows are written up to the frame's count
        pd = torch.full((len(own), ex.cap), -1, dtype=torch.int32, device="cuda")
        pn = torch.zeros(len(own), dtype=torch.int32, device="cuda")
        mt.projection_batch_device(ex, pl, pc, np.full(len(own), -2, np.float32), np.full(len(own), -1, np.float32), 15.0,
                                   pm.data_ptr(), pd.data_ptr(), pn.data_ptr())
        mt.sync()
        proj[torch.tensor([c for _, c in own], device="cuda")] = pm
    ex.sync()
    desc, ang, cnt = blocks_of(ex, he - hb)
    desc, ang, cnt = desc[b - hb:], ang[b - hb:], cnt[b - hb:]  # the halo frame belongs to the previous rank
    pairs = np.array([(i, j) for i in range(n_frames) for j in range(n_frames) if i != j and (i + 2 * j) % 3 == 0], np.int32)
    ids = [sweep.Sweep.unique_id() if (rank == 0 and world > 1) else None]
    if world > 1:
        dist.broadcast_object_list(ids, src=0)
    sw = sweep.Sweep(rank, world, local, ids[0])
    sel, m, dd, nm = sweep.sweep_cabi(sw, mt, eaof.BOW_KF_FRAME, desc, ang, cnt, pairs, n_frames)
    ag_ms, ag_bytes = sw.last_allgather()
    # collect everything on rank 0
    stride = ex.cap
    full_m = torch.full((len(pairs), stride), -2, dtype=torch.int32, device="cuda")
    full_n = torch.full((len(pairs),), -2, dtype=torch.int32, device="cuda")
    full_m[torch.from_numpy(sel).cuda()] = m
    full_n[torch.from_numpy(sel).cuda()] = nm
    if world > 1:
        dist.all_reduce(full_m, op=dist.ReduceOp.MAX)
        dist.all_reduce(full_n, op=dist.ReduceOp.MAX)
        dist.all_reduce(proj, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        ex1 = eaof.ORBextractor(NF, 1.2, 8, 20, 7, width=W, height=H, max_batch=n_frames, device=local)
        d_all = torch.from_numpy(frames).cuda()
        ex1.extract_batch_device(d_all.data_ptr(), n_frames)
        ex1.sync()
        desc1, ang1, cnt1 = blocks_of(ex1, n_frames)
        sw1 = sweep.Sweep(0, 1, local)
        sel1, m1, dd1, nm1 = sweep.sweep_cabi(sw1, mt, eaof.BOW_KF_FRAME, desc1, ang1, cnt1, pairs, n_frames)
        sw1.close()
        assert (sel1 == np.arange(len(pairs))).all()
        ok = bool(torch.equal(full_m, m1)) and bool(torch.equal(full_n, nm1))
        # consecutive-frame matching, single GPU over the whole sequence
        allp = shard.consecutive_pairs(0, n_frames, n_frames)
        pm1 = torch.full((len(allp), ex1.cap), -1, dtype=torch.int32, device="cuda")
        pd1 = torch.full((len(allp), ex1.cap), -1, dtype=torch.int32, device="cuda")
        pn1 = torch.zeros(len(allp), dtype=torch.int32, device="cuda")
        mt1 = eaof.ORBmatcher(0.9, True, max_features=ex1.cap, max_pairs=len(allp), device=local)
        mt1.projection_batch_device(ex1, np.array([a for a, _ in allp], np.int32), np.array([c for _, c in allp], np.int32),
                                    np.full(len(allp), -2, np.float32), np.full(len(allp), -1, np.float32), 15.0,
                                    pm1.data_ptr(), pd1.data_ptr(), pn1.data_ptr())
        mt1.sync()
        ok_proj = bool(torch.equal(proj[1:], pm1))
        mt1.close()
        ok = ok and ok_proj
        # oracle spot checks
        from oracle import pyoracle as po
        hd, ha, hc = desc1.cpu().numpy(), ang1.cpu().numpy(), cnt1.cpu().numpy()
        for pi in (0, len(pairs) // 2, len(pairs) - 1):
            q, t = pairs[pi]
            nq, nt = int(hc[q]), int(hc[t])
            nodes_q, nodes_t = eaof.csr_from_nodes(np.zeros(nq, int)), eaof.csr_from_nodes(np.zeros(nt, int))
            on, om, _ = po.o_search_by_bow(0, 0.9, True, hd[q, :nq], ha[q, :nq], None, nodes_q, hd[t, :nt], ha[t, :nt], None, nodes_t)
            ok = ok and on == int(full_n[pi]) and np.array_equal(om, full_m[pi, :nt].cpu().numpy())
        print(f"MULTI_GPU_CHECK world={world} nccl={sweep.Sweep.nccl_version() if world > 1 else 0} sweep_pairs={len(pairs)} "
              f"matches/pair={float(full_n.float().mean()):.1f} allgather_ms={ag_ms:.3f} allgather_bytes={ag_bytes} "
              f"consecutive_pairs={n_frames - 1} halo_matching={'OK' if ok_proj else 'MISMATCH'} "
              f"{'OK' if ok else 'MISMATCH'}", flush=True)
        ex1.close()
    sw.close()
    mt.close()
    ex.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
