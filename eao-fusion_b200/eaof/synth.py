"""Deterministic synthetic textured frames.

All arithmetic is integer (numpy int64), so the frames are reproducible bit-for-bit
without cv2: a base texture of three octaves of box-smoothed uniform noise, plus
rotated filled rectangles and thin line segments; frame *t* of a sequence is a crop
that drifts by (2, 1) px per frame, which mimics slow camera motion so that
consecutive frames can be matched with a known inter-frame translation.
"""
from __future__ import annotations

import numpy as np

_DRIFT_PERIOD = 900
_MARGIN = 64


def _box_blur(a: np.ndarray, r: int) -> np.ndarray:
    """(2r+1)^2 box mean with edge replication, integer floor division."""
    k = 2 * r + 1
    p = np.pad(a, r, mode="edge").astype(np.int64)
    c = np.cumsum(p, axis=0)
    c = np.concatenate([np.zeros((1, c.shape[1]), np.int64), c], axis=0)
    v = c[k:, :] - c[:-k, :]
    c = np.cumsum(v, axis=1)
    c = np.concatenate([np.zeros((c.shape[0], 1), np.int64), c], axis=1)
    h = c[:, k:] - c[:, :-k]
    return h // (k * k)


def _octave(rng: np.random.Generator, h: int, w: int, r: int) -> np.ndarray:
    n = rng.integers(0, 256, size=(h, w), dtype=np.int64) * 64
    for _ in range(3):  # three box passes ~ Gaussian
        n = _box_blur(n, r)
    lo, hi = int(n.min()), int(n.max())
    return (n - lo) * 255 // max(hi - lo, 1)


def base_texture(width: int, height: int, seed: int = 1234) -> np.ndarray:
    """u8 texture of size (height+1024) x (width+1024)."""
    H, W = height + 2 * _MARGIN + _DRIFT_PERIOD, width + 2 * _MARGIN + _DRIFT_PERIOD
    rng = np.random.Generator(np.random.PCG64(seed))
    tex = (25 * _octave(rng, H, W, 1) + 35 * _octave(rng, H, W, 3) + 40 * _octave(rng, H, W, 8)) // 100
    tex = tex.astype(np.int64)
    n_rect = (H * W) // 4900
    n_line = (H * W) // 13000
    ys, xs = np.mgrid[0:H, 0:W]
    for i in range(n_rect + n_line):
        is_line = i >= n_rect
        cx, cy = int(rng.integers(0, W)), int(rng.integers(0, H))
        ux, uy = 0, 0
        while ux == 0 and uy == 0:
            ux, uy = int(rng.integers(-16, 17)), int(rng.integers(-16, 17))
        a = int(rng.integers(4, 40))           # half length along u
        b2 = int(rng.integers(1, 3)) if is_line else 2 * int(rng.integers(4, 40))  # full width across u
        g = int(rng.integers(0, 256))
        rad = a + b2 + 2
        x0, x1, y0, y1 = max(cx - rad, 0), min(cx + rad + 1, W), max(cy - rad, 0), min(cy + rad + 1, H)
        if x0 >= x1 or y0 >= y1:
            continue
        dx = xs[y0:y1, x0:x1] - cx
        dy = ys[y0:y1, x0:x1] - cy
        n2 = ux * ux + uy * uy
        along = dx * ux + dy * uy
        across = -dx * uy + dy * ux
        m = (along * along <= a * a * n2) & (4 * across * across <= b2 * b2 * n2)
        tex[y0:y1, x0:x1][m] = g
    return tex.astype(np.uint8)


def frame_offset(t: int) -> tuple[int, int]:
    return _MARGIN + (2 * t) % _DRIFT_PERIOD, _MARGIN + t % _DRIFT_PERIOD


def make_frames(n: int, width: int = 640, height: int = 480, seed: int = 1234, start: int = 0,
                tex: np.ndarray | None = None) -> np.ndarray:
    """(n, height, width) u8, C-contiguous."""
    if tex is None:
        tex = base_texture(width, height, seed)
    out = np.empty((n, height, width), np.uint8)
    for i in range(n):
        ox, oy = frame_offset(start + i)
        out[i] = tex[oy:oy + height, ox:ox + width]
    return out


def adversarial_frames(width: int, height: int) -> dict[str, np.ndarray]:
    """Edge-case frames the parity tests run (SURVEY.md §8(c) 'fixtures to create')."""
    rng = np.random.Generator(np.random.PCG64(99))
    yy, xx = np.mgrid[0:height, 0:width]
    f = {
        "flat0": np.zeros((height, width), np.uint8),
        "flat128": np.full((height, width), 128, np.uint8),
        "flat255": np.full((height, width), 255, np.uint8),
        "checker1": (((xx + yy) & 1) * 255).astype(np.uint8),
        "checker8": ((((xx >> 3) + (yy >> 3)) & 1) * 255).astype(np.uint8),
        "noise": rng.integers(0, 256, size=(height, width), dtype=np.uint8),
        "lowcontrast": (120 + rng.integers(0, 12, size=(height, width))).astype(np.uint8),
    }
    one = np.full((height, width), 20, np.uint8)
    one[height // 2, width // 2] = 255
    f["single_bright"] = one
    return f


def vocabulary(k: int = 10, L: int = 6, seed: int = 7, weighting: int = 0, scoring: int = 0) -> dict:
    """A complete k-ary vocabulary tree of depth L in the array form eaof_voc_create takes (the shape of ORBvoc, which
    is k=10, L=6: 1,111,111 nodes, 10^6 words).  Nodes are numbered level by level, children of a node are consecutive;
    a child descriptor is its parent's with a share of the bits flipped that shrinks with depth; leaf weights are
    random positive idf-like values.  Vectorised (seconds for the ORBvoc shape)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    levels = [np.zeros((1, 32), np.uint8)]
    for lvl in range(L):
        parent = np.repeat(levels[-1], k, axis=0)
        p = 0.25 / (1.6 ** lvl)
        flips = np.packbits(rng.random((len(parent), 256), dtype=np.float32) < p, axis=1)
        levels.append(parent ^ flips if lvl else rng.integers(0, 256, (k, 32), dtype=np.uint8))
    desc = np.concatenate(levels)
    n = len(desc)
    n_inner = n - k ** L
    child_start = np.minimum(np.arange(n + 1, dtype=np.int64) * k, n - 1).astype(np.int32)  # node i -> children 1+i*k ..
    child_idx = np.arange(1, n, dtype=np.int32)
    weight = np.zeros(n, np.float64)
    weight[n_inner:] = rng.uniform(0.5, 12.0, k ** L)
    word_id = np.full(n, -1, np.int32)
    word_id[n_inner:] = np.arange(k ** L, dtype=np.int32)
    return dict(L=L, k=k, child_start=child_start, child_idx=child_idx, desc=desc, weight=weight, word_id=word_id,
                weighting=weighting, scoring=scoring)
