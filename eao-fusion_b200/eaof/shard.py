"""Sharding plans for the multi-GPU paths: one process per GPU, no collective on the extraction
path; the cross-frame Hamming sweep all-gathers descriptor blocks once and then partitions the pair list."""
from __future__ import annotations


def frame_block(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [begin, end) of frames owned by `rank` (blocks differ by at most one frame)."""
    base, rem = divmod(n_frames, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def consecutive_pairs(begin: int, end: int, n_frames: int) -> list[tuple[int, int]]:
    """(last, cur) pairs whose CURRENT frame lies in [begin, end): the pair across a block boundary belongs to the
    rank that owns `cur`, which therefore also extracts frame begin-1 as a halo (cheaper than a transfer)."""
    return [(f - 1, f) for f in range(max(begin, 1), min(end, n_frames))]


def halo_block(begin: int, end: int) -> tuple[int, int]:
    """Frames a rank must extract to match every pair it owns: its block plus one halo frame in front."""
    return max(begin - 1, 0), end


def sweep_pairs(n_pairs: int, rank: int, world: int) -> range:
    """Round-robin partition of a global pair list for the brute-force sweep."""
    return range(rank, n_pairs, world)
