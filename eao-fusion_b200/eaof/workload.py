"""The BASELINE.json workloads as data: frame sizes / extractor parameters of every config, the synthetic sequences
(global frame index -> frame, so that any rank can build exactly its shard), and the result digests used for the
parity-on-workload and determinism-across-world-sizes checks.  No oracle code here (tests/ and bench.py bring that)."""
from __future__ import annotations

import hashlib

import numpy as np

from . import synth

# SURVEY.md §8(d) / BASELINE.json configs; all with 8 levels, 1.2, iniThFAST 20, minThFAST 7
CONFIGS = {
    "configs[0]": dict(width=640, height=480, nfeatures=1000, frames=1,
                       what="single synthetic 640x480 frame, nfeatures=1000 (TUM RGB-D yaml) through operator()"),
    "configs[1]": dict(width=640, height=480, nfeatures=1000, frames=1000,
                       what="1000-frame synthetic 640x480 sequence, nfeatures=1000, nlevels=8, scaleFactor=1.2, iniThFAST=20, "
                            "minThFAST=7; batched extraction + consecutive-frame SearchByProjection-style matching"),
    "configs[2]": dict(width=848, height=480, nfeatures=1200, frames=10000,
                       what="RealSense D435i settings: 848x480, nfeatures=1200, 8 levels, batched over 10k frames"),
    "configs[3]": dict(width=1920, height=1080, nfeatures=4000, frames=10000,
                       what="high-res stress: 1920x1080, nfeatures=4000, nlevels=8, scaleFactor=1.2, 10k frames"),
}
NLEVELS, SCALE, INI_TH, MIN_TH = 8, 1.2, 20, 7
SEQ_LEN = 1000           # frames of one scene; a longer sequence is a series of scenes (one texture each)
MATCH_TH = 15.0          # TrackWithMotionModel's window, src/Tracking.cc:1749-1753
SHIFT = (-2.0, -1.0)     # the synthetic sequence drifts by (2,1) px per frame (synth.frame_offset)


class Sequence:
    """Global frame t of a width x height sequence: scene t // SEQ_LEN (its own texture, seed 1235 + scene), frame
    t % SEQ_LEN of that scene.  Scene 0 is the configs[1] sequence of round 1."""

    def __init__(self, width: int, height: int, seed: int = 1235):
        self.w, self.h, self.seed = width, height, seed
        self._tex = {}

    def texture(self, scene: int) -> np.ndarray:
        if scene not in self._tex:
            self._tex[scene] = synth.base_texture(self.w, self.h, seed=self.seed + scene)
        return self._tex[scene]

    def frames(self, begin: int, end: int) -> np.ndarray:
        """(end-begin, h, w) u8, global frames [begin, end)."""
        out = np.empty((max(end - begin, 0), self.h, self.w), np.uint8)
        t = begin
        while t < end:
            scene = t // SEQ_LEN
            stop = min(end, (scene + 1) * SEQ_LEN)
            out[t - begin:stop - begin] = synth.make_frames(stop - t, self.w, self.h, start=t % SEQ_LEN, tex=self.texture(scene))
            t = stop
        return out


def frame_digest(kps: np.ndarray, desc: np.ndarray) -> bytes:
    """sha256 over the keypoint records (x, y, size, angle, response, octave — 24 bytes each) and the descriptors."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(kps).tobytes())
    h.update(np.ascontiguousarray(desc).tobytes())
    return h.digest()


def pair_digest(match: np.ndarray, dist: np.ndarray | None = None) -> bytes:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(match, np.int32).tobytes())
    if dist is not None:
        h.update(np.ascontiguousarray(dist, np.int32).tobytes())
    return h.digest()


def combine(digests) -> str:
    """One hex digest of a list of per-item digests in item order."""
    h = hashlib.sha256()
    for d in digests:
        h.update(d)
    return h.hexdigest()
