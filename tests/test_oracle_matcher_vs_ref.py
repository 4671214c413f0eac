"""Pins oracle/match_oracle.cc (the restatement the CUDA matcher is diffed against) to the reference's own code:
oracle/_ref/libmatch_ref.so is the UNMODIFIED /root/reference/src/ORBmatcher.cc compiled against oracle/matchshim and
driven with the same arrays.  Match indices and match counts must agree exactly."""
import os

import numpy as np
import pytest

from matchdata import planted_pair, random_nodes, tri_inputs as _tri_inputs
from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not (os.path.exists(po.MATCH_REF_SO) or os.path.exists("/root/reference/src/ORBmatcher.cc")),
                                reason="reference matcher binary not built and /root/reference absent")


def _csr(node):
    import eaof
    return eaof.csr_from_nodes(node)


def test_constants_and_distance():
    L = po.match_ref_lib()
    assert (L.mref_th_low(), L.mref_th_high(), L.mref_histo_length()) == (50, 100, 30)
    rng = np.random.Generator(np.random.PCG64(3))
    a = rng.integers(0, 256, size=(500, 32), dtype=np.uint8)
    b = rng.integers(0, 256, size=(500, 32), dtype=np.uint8)
    a[:3] = 0; b[:3] = 255; b[3:6] = a[3:6]
    r = po.r_hamming(a, b)
    assert np.array_equal(r, po.o_hamming(a, b))
    assert np.array_equal(r, np.unpackbits(a ^ b, axis=1).sum(1))


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("ratio", [0.6, 0.75, 0.9])
@pytest.mark.parametrize("ori", [True, False])
def test_search_by_bow(mode, ratio, ori):
    total = 0
    for seed, (nq, nt, nn) in enumerate([(400, 500, 1), (600, 600, 12), (300, 7, 2), (1, 1, 1), (64, 129, 5), (500, 500, 60)]):
        q, aq, t, at = planted_pair(nq, nt, 10 + seed, dup=5 if nt > 50 else 0)
        nodes_q = _csr(random_nodes(nq, nn, 20 + seed) if nn > 1 else np.zeros(nq, int))
        nodes_t = _csr(random_nodes(nt, nn, 30 + seed) if nn > 1 else np.zeros(nt, int))
        rng = np.random.Generator(np.random.PCG64(100 + seed))
        vq = (rng.random(nq) > 0.1).astype(np.uint8)
        vt = (rng.random(nt) > 0.1).astype(np.uint8)
        on, omatch, _ = po.o_search_by_bow(mode, ratio, ori, q, aq, vq, nodes_q, t, at, vt, nodes_t)
        rn, rmatch = po.r_search_by_bow(mode, ratio, ori, q, aq, vq, nodes_q, t, at, vt, nodes_t)
        assert rn == on, (seed, rn, on)
        assert np.array_equal(rmatch, omatch), seed
        total += on
    assert total > 0


@pytest.mark.parametrize("only_stereo", [False, True])
@pytest.mark.parametrize("ori", [True, False])
def test_search_for_triangulation(only_stereo, ori):
    total = 0
    for seed in range(4):
        k1, k2, F12, sf, ls = _tri_inputs(50 + seed)
        for epipole in ((320.0, 240.0), (-1e4, -1e4)):  # inside the image (the epipole-distance test bites) / far away
            on, om, _ = po.o_search_for_triangulation(k1, k2, F12, epipole, sf, ls, only_stereo, ori)
            rn, rm = po.r_search_for_triangulation(k1, k2, F12, epipole, sf, ls, only_stereo, ori)
            assert rn == on and np.array_equal(rm, om), (seed, rn, on)
            total += int((om >= 0).sum())
    assert total > 100


def _proj_inputs(seed, n=900, stereo=False, flags=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    q, aq, t, at = planted_pair(n, n, seed, flip=0.06, frac=0.8, dup=8)
    x = rng.uniform(-5, 645, n).astype(np.float32)
    y = rng.uniform(-5, 485, n).astype(np.float32)
    octv = rng.integers(0, 8, n).astype(np.int32)
    cur = dict(x=x, y=y, octave=octv, angle=at, desc=t)
    # Last feature i projects near Cur feature perm[i]
    perm = rng.permutation(n)
    last = dict(u=(x[perm] + rng.normal(0, 3, n)).astype(np.float32), v=(y[perm] + rng.normal(0, 3, n)).astype(np.float32),
                octave=np.clip(octv[perm] + rng.integers(-1, 2, n), 0, 7).astype(np.int32), angle=aq, desc=q)
    # make descriptors of the planted partner line up with the geometry for half of the rows
    half = perm[: n // 2]
    cur["desc"][half] = q[: n // 2] ^ np.packbits((rng.random((n // 2, 256)) < 0.05).astype(np.uint8), axis=1)
    if stereo:
        cur["uright"] = np.where(rng.random(n) > 0.3, x - 20 + rng.normal(0, 4, n), -1).astype(np.float32)
        last["invz"] = rng.choice(np.array([1, 0.5, 0.25, 2, -1], np.float32), n, p=[0.5, 0.2, 0.1, 0.15, 0.05])
    if flags:
        cur["taken"] = (rng.random(n) < 0.1).astype(np.uint8)
        last["valid"] = (rng.random(n) > 0.1).astype(np.uint8)
        last["obs"] = (rng.random(n) > 0.2).astype(np.uint8)
    return cur, last


@pytest.mark.parametrize("th", [7.0, 15.0, 30.0])
@pytest.mark.parametrize("case", ["plain", "flags", "stereo", "stereo_fwd", "stereo_bwd", "no_ori"])
def test_search_by_projection_last(th, case):
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    bounds = (0.0, 640.0, 0.0, 480.0)
    ginv = (np.float32(64) / np.float32(640), np.float32(48) / np.float32(480))
    total = 0
    for seed in range(3):
        cur, last = _proj_inputs(200 + seed, stereo=case.startswith("stereo"), flags=case in ("flags", "stereo"))
        mode = {"stereo_fwd": 1, "stereo_bwd": 2}.get(case, 0)
        kw = dict(bounds=bounds, grid_inv=ginv, scale_factors=sf, mbf=40.0 if case.startswith("stereo") else 0.0, search_mode=mode)
        on, om, _ = po.o_search_by_projection(cur, last, th, case != "no_ori", **kw)
        rn, rm = po.r_search_by_projection(cur, last, th, case != "no_ori", **kw)
        assert rn == on, (seed, rn, on)
        assert np.array_equal(rm, om), seed
        total += on
    assert total > 0


_SF = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
_BOUNDS = (0.0, 640.0, 0.0, 480.0)
_GINV = (np.float32(64) / np.float32(640), np.float32(48) / np.float32(480))


@pytest.mark.parametrize("th", [1.0, 3.0, 5.0])
@pytest.mark.parametrize("ratio", [0.8, 0.6])
@pytest.mark.parametrize("case", ["plain", "flags", "stereo"])
def test_search_by_projection_mappoints(th, ratio, case):
    from matchdata import window_scene
    total = rejected = 0
    for seed in range(3):
        F, mp = window_scene(300 + seed, stereo=case == "stereo", flags=case != "plain")
        kw = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF)
        on, om, _ = po.o_search_by_projection_mappoints(F, mp, th, ratio, **kw)
        rn, rm = po.r_search_by_projection_mappoints(F, mp, th, ratio, **kw)
        assert rn == on, (seed, rn, on)
        assert np.array_equal(rm, om), seed
        total += on
        rejected += po.o_search_by_projection_mappoints(F, mp, th, 1e9, **kw)[0] - on
    assert total > 100 and (rejected > 0 or th == 1.0)  # the same-level ratio test must have bitten somewhere


@pytest.mark.parametrize("th", [7.0, 15.0])
@pytest.mark.parametrize("orb_dist", [50, 100])
@pytest.mark.parametrize("case", ["plain", "flags", "no_ori"])
def test_search_by_projection_kf(th, orb_dist, case):
    from matchdata import kf_scene
    lsf = float(np.log(np.float32(1.2)))
    total = 0
    for seed in range(3):
        F, kf = kf_scene(400 + seed, flags=case == "flags")
        valid, u, v, lvl = po.kf_projection(kf, lsf, 8)
        kq = dict(valid=valid, u=u, v=v, level=lvl, angle=kf["angle"], desc=kf["desc"])
        kw = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF)
        on, om, _ = po.o_search_by_projection_kf(F, kq, th, orb_dist, case != "no_ori", **kw)
        rn, rm = po.r_search_by_projection_kf(F, kf, th, orb_dist, case != "no_ori", log_scale_factor=lsf, **kw)
        assert rn == on, (seed, rn, on)
        assert np.array_equal(rm, om), seed
        total += on
    assert total > 30


@pytest.mark.parametrize("window", [10, 100])
@pytest.mark.parametrize("ratio", [0.9, 0.6])
@pytest.mark.parametrize("ori", [True, False])
def test_search_for_initialization(window, ratio, ori):
    from matchdata import init_scene
    total = 0
    for seed in range(3):
        F1, F2, prev = init_scene(500 + seed)
        on, om, op = po.o_search_for_initialization(F1, F2, prev, window, ratio, ori, bounds=_BOUNDS, grid_inv=_GINV)
        rn, rm, rp = po.r_search_for_initialization(F1, F2, prev, window, ratio, ori, bounds=_BOUNDS, grid_inv=_GINV)
        assert rn == on and np.array_equal(rm, om) and np.array_equal(rp, op), (seed, rn, on)
        total += on
    assert total > 300
