"""Cross-shard brute-force Hamming sweep (BASELINE.json configs[4]).

Every rank holds the descriptor blocks of the frames it extracted (one block = one frame: `stride` rows of 32 bytes,
angles, a feature count).  The only exchange step of the whole ORB path happens here: the blocks are all-gathered once
over NCCL (NVLink/NVSwitch) so that every rank can match any (query frame, target frame) pair; the global pair list is
then partitioned round-robin and each rank matches its share.

On GPUs the exchange and the matching go through the C ABI of include/eaof_sweep.h (`Sweep`: ncclCommInitRank /
ncclAllGather inside libeaof_orb.so); the host program only carries the 128-byte NCCL id to the other ranks.
`gather_blocks` is the same block layout over a `torch.distributed` process group — it exists for the world_size-2
gloo tests of the layout on CPU boxes (tests/test_host_logic.py) and is not used by bench.py.
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


class Sweep:
    """include/eaof_sweep.h over ctypes: one NCCL communicator owned by libeaof_orb.so.  `id_bytes`: the 128-byte id
    made by `Sweep.unique_id()` on one rank and carried to the others by the caller (None for world == 1)."""

    def __init__(self, rank=0, world=1, device=0, id_bytes=None):
        import ctypes as C
        from . import _ck, lib
        self.L = lib()
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        self.L.eaof_sweep_create.argtypes = [vp, ci, ci, ci, C.POINTER(vp)]
        self.L.eaof_sweep_destroy.argtypes = [vp]
        self.L.eaof_sweep_allgather_blocks.argtypes = [vp, ci, ci] + [vp] * 7
        self.L.eaof_sweep_match.argtypes = [vp, vp, ci, cf, ci, ci, ci] + [vp] * 6 + [ci] + [vp] * 5
        self.L.eaof_sweep_last_allgather.argtypes = [vp, C.POINTER(cf), C.POINTER(C.c_longlong)]
        h = vp()
        buf = None
        if world > 1:
            assert id_bytes is not None and len(id_bytes) == 128
            buf = (C.c_uint8 * 128).from_buffer_copy(bytes(id_bytes))
        _ck(self.L.eaof_sweep_create(buf, rank, world, device, C.byref(h)))
        self.h, self.rank, self.world = h, rank, world

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from . import _ck, lib
        buf = (C.c_uint8 * 128)()
        _ck(lib().eaof_sweep_unique_id(buf))
        return bytes(buf)

    @staticmethod
    def nccl_version() -> int:
        import ctypes as C
        from . import _ck, lib
        v = C.c_int()
        _ck(lib().eaof_sweep_nccl_version(C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "h", None):
            self.L.eaof_sweep_destroy(self.h)
            self.h = None

    __del__ = close

    def allgather_blocks(self, per, stride, d_desc, d_angle, d_count, g_desc, g_angle, g_count, stream_ptr=None):
        from . import _ck
        _ck(self.L.eaof_sweep_allgather_blocks(self.h, per, stride, d_desc, d_angle, d_count, g_desc, g_angle, g_count, stream_ptr))

    def match(self, matcher, mode, per, stride, d_desc, d_angle, d_count, g_desc, g_angle, g_count, pq, pt, d_match, d_dist, d_nm):
        """All-gather + this rank's pairs (pq / pt: int32 numpy arrays of gathered-block indices), asynchronous on the
        matcher's stream."""
        from . import _ck
        pq = np.ascontiguousarray(pq, np.int32)
        pt = np.ascontiguousarray(pt, np.int32)
        _ck(self.L.eaof_sweep_match(self.h, matcher.h, mode, matcher.mfNNratio, int(matcher.mbCheckOrientation), per, stride,
                                    d_desc, d_angle, d_count, g_desc, g_angle, g_count, len(pq),
                                    pq.ctypes.data if len(pq) else None, pt.ctypes.data if len(pt) else None, d_match, d_dist, d_nm))

    def last_allgather(self):
        import ctypes as C
        from . import _ck
        ms, b = C.c_float(), C.c_longlong()
        _ck(self.L.eaof_sweep_last_allgather(self.h, C.byref(ms), C.byref(b)))
        return ms.value, b.value


def sweep_cabi(sw: "Sweep", matcher, mode, desc, angle, counts, pairs, n_frames):
    """The product path: desc [n_local, stride, 32] u8, angle [n_local, stride] f32, counts [n_local] i32 (torch tensors
    on this rank's device) -> (positions of this rank's pairs in the global list, match, dist, nmatches) with the
    exchange done by ncclAllGather inside libeaof_orb.so."""
    import torch
    world, rank = sw.world, sw.rank
    per = padded_blocks(n_frames, world)
    n_local, stride = desc.shape[0], desc.shape[1]

    def pad(t):
        if t.shape[0] == per:
            return t.contiguous()
        out = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        out[:n_local] = t
        return out
    d, a, c = pad(desc), pad(angle), pad(counts)
    gd = torch.empty((world * per, stride, 32), dtype=torch.uint8, device=d.device)
    ga = torch.empty((world * per, stride), dtype=torch.float32, device=d.device)
    gc = torch.empty((world * per,), dtype=torch.int32, device=d.device)
    sel, pq, pt = my_pairs(np.asarray(pairs), n_frames, rank, world)
    n = max(len(sel), 1)
    d_match = torch.full((n, stride), -1, dtype=torch.int32, device=d.device)
    d_dist = torch.full((n, stride), -1, dtype=torch.int32, device=d.device)
    d_nm = torch.zeros(n, dtype=torch.int32, device=d.device)
    torch.cuda.current_stream().synchronize()  # inputs / outputs were made on torch's stream, the sweep runs on the matcher's
    sw.match(matcher, mode, per, stride, d.data_ptr(), a.data_ptr(), c.data_ptr(), gd.data_ptr(), ga.data_ptr(), gc.data_ptr(),
             pq, pt, d_match.data_ptr(), d_dist.data_ptr(), d_nm.data_ptr())
    matcher.sync()
    return sel, d_match[:len(sel)], d_dist[:len(sel)], d_nm[:len(sel)]
