"""Cross-shard brute-force Hamming sweep (BASELINE.json configs[4]).

Every rank holds the descriptor blocks of the frames it extracted (one block = one frame: `stride` rows of 32 bytes,
angles, a feature count).  The only exchange step of the whole ORB path happens here: the blocks are all-gathered once
over NCCL (NVLink/NVSwitch) so that every rank can match any (query frame, target frame) pair; the global pair list is
then partitioned round-robin and each rank runs `eaof_match_bruteforce_batch_device` on its share.  There is no other
collective on the data path.  `torch.distributed` is plumbing only: device buffers in, device buffers out.
"""
from __future__ import annotations

import numpy as np

from . import shard


class DevicePtr:
    """Wraps a raw device pointer so that torch can view it (no copy): torch.as_tensor(DevicePtr(...), device='cuda')."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def padded_blocks(n_frames: int, world: int) -> int:
    """Blocks every rank contributes to the all-gather (the last ranks pad with empty blocks)."""
    return -(-n_frames // world)


def global_block(frame: int, n_frames: int, world: int) -> int:
    """Index of `frame`'s block inside the gathered array: owner rank r holds frames frame_block(n, r, world) and
    contributes padded_blocks() slots."""
    per = padded_blocks(n_frames, world)
    for r in range(world):
        b, e = shard.frame_block(n_frames, r, world)
        if b <= frame < e:
            return r * per + (frame - b)
    raise IndexError(frame)


def gather_blocks(desc, angle, counts, n_frames: int, dist_module=None):
    """desc [n_local, stride, 32] u8, angle [n_local, stride] f32, counts [n_local] i32 (torch tensors on this rank's
    device; n_local = size of this rank's frame_block).  Returns the gathered (desc, angle, counts) with
    world*padded_blocks() blocks each.  One all-gather per array; with world == 1 nothing is exchanged."""
    import torch
    dist = dist_module
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    per = padded_blocks(n_frames, world)
    n_local, stride = desc.shape[0], desc.shape[1]

    def pad(t, fill=0):
        if t.shape[0] == per:
            return t.contiguous()
        out = torch.full((per,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
        out[:n_local] = t
        return out
    d, a, c = pad(desc), pad(angle), pad(counts)
    if world == 1:
        return d, a, c
    gd = torch.empty((world * per, stride, 32), dtype=d.dtype, device=d.device)
    ga = torch.empty((world * per, stride), dtype=a.dtype, device=a.device)
    gc = torch.empty((world * per,), dtype=c.dtype, device=c.device)
    dist.all_gather_into_tensor(gd, d)
    dist.all_gather_into_tensor(ga, a)
    dist.all_gather_into_tensor(gc, c)
    return gd, ga, gc


def my_pairs(pairs: np.ndarray, n_frames: int, rank: int, world: int):
    """This rank's share of the global (query frame, target frame) list, as gathered-block indices, plus the
    positions of those pairs in the global list."""
    sel = np.asarray(list(shard.sweep_pairs(len(pairs), rank, world)), np.int64)
    pq = np.array([global_block(int(pairs[i, 0]), n_frames, world) for i in sel], np.int32)
    pt = np.array([global_block(int(pairs[i, 1]), n_frames, world) for i in sel], np.int32)
    return sel, pq, pt


def sweep(matcher, mode, desc, angle, counts, pairs, n_frames, rank=0, world=1, dist_module=None):
    """All-gather the blocks, match this rank's pairs.  Returns (positions in the global pair list, match [p, stride],
    dist [p, stride], nmatches [p]) as torch tensors on the device (positions: numpy)."""
    import torch
    gd, ga, gc = gather_blocks(desc, angle, counts, n_frames, dist_module)
    sel, pq, pt = my_pairs(np.asarray(pairs), n_frames, rank, world)
    stride = gd.shape[1]
    n = max(len(sel), 1)
    d_match = torch.full((n, stride), -1, dtype=torch.int32, device=gd.device)
    d_dist = torch.full((n, stride), -1, dtype=torch.int32, device=gd.device)
    d_nm = torch.zeros(n, dtype=torch.int32, device=gd.device)
    torch.cuda.current_stream().synchronize()  # the gathered blocks must be complete before the matcher's own stream reads them
    for s0 in range(0, len(sel), matcher.max_pairs):
        s1 = min(len(sel), s0 + matcher.max_pairs)
        matcher.bruteforce_batch_device(mode, pq[s0:s1], pt[s0:s1], gd.data_ptr(), ga.data_ptr(), gc.data_ptr(), stride,
                                        d_match[s0:].data_ptr(), d_dist[s0:].data_ptr(), d_nm[s0:].data_ptr())
    matcher.sync()
    return sel, d_match[:len(sel)], d_dist[:len(sel)], d_nm[:len(sel)]
