"""Parity on the bench workloads themselves: the whole configs[1] sequence — 1000 frames and all 999
consecutive pairs — and long runs of the other BASELINE sizes, CUDA path against the unmodified reference ORBextractor.cc
(oracle/_ref/liborb_ref.so, quadtree ties in creation order) and the matcher oracle, compared through sha256 digests per
frame / per pair so that a mismatch names the frame."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mods():
    import torch
    import eaof
    import bench
    return torch, eaof, bench


def test_configs1_full_sequence_1000_frames_999_pairs():
    torch, eaof, bench = _mods()
    from eaof import workload
    seq = workload.Sequence(640, 480)
    out = bench.determinism_and_parity(eaof, torch, None, 0, 1, 0, seq, os.cpu_count() or 1, check_cpu=True)
    par = out["parity_vs_cpu_reference"]
    assert "failed" not in par, par
    assert par["frames_checked"] == 1000 and par["pairs_checked"] == 999
    assert par["identical"], par
    assert out["hash"] == par["hash_cpu"]


def test_sharded_digests_equal_the_single_run():
    """The sharding plan of eaof/shard.py run on ONE GPU, shard after shard: block + halo frame per shard, digests in frame
    order must equal the unsharded run (what bench.py's determinism leg does across real ranks)."""
    torch, eaof, bench = _mods()
    from eaof import shard, workload
    seq = workload.Sequence(640, 480)
    n = 120
    whole = seq.frames(0, n)
    f1, p1 = bench.digests_of(eaof, torch, 0, whole, first_is_halo=False, B=50)
    assert len(f1) == n and len(p1) == n - 1
    for world in (2, 3):
        fs, ps = [], []
        for r in range(world):
            b, e = shard.frame_block(n, r, world)
            hb, he = shard.halo_block(b, e)
            f, p = bench.digests_of(eaof, torch, 0, whole[hb:he], first_is_halo=hb < b, B=50)
            fs += f
            ps += p
        assert fs == f1 and ps == p1


@pytest.mark.parametrize("name,n", [("configs[2]", 100), ("configs[3]", 32)])
def test_other_baseline_sizes_long_runs(name, n):
    torch, eaof, bench = _mods()
    from eaof import workload
    from oracle import pyoracle as po
    cfg = workload.CONFIGS[name]
    w, h, nf = cfg["width"], cfg["height"], cfg["nfeatures"]
    frames = workload.Sequence(w, h, seed=4321).frames(0, n)
    B = 25 if w < 1000 else 8
    ex = eaof.ORBextractor(nf, workload.SCALE, workload.NLEVELS, workload.INI_TH, workload.MIN_TH, width=w, height=h, max_batch=B)
    got = []
    for s in range(0, n, B):
        got += ex.extract_batch(frames[s:s + B])
    ex.close()
    _, _, ref = po.ref_extract_many(frames, nf, workload.SCALE, workload.NLEVELS, workload.INI_TH, workload.MIN_TH,
                                    threads=os.cpu_count() or 1, canonical=True)
    bad = [i for i in range(n) if workload.frame_digest(*got[i]) != workload.frame_digest(*ref[i])]
    assert not bad, f"{name}: frames {bad[:8]} differ from the reference"
