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


# ---- map-side matchers: SearchByProjection(KF,Scw), Fuse x2, SearchBySim3 (src/ORBmatcher.cc:290-403, :825-1326) ----------
_LSF = float(np.log(np.float32(1.2)))
_INV_SIGMA2 = (np.float32(1.0) / (_SF * _SF)).astype(np.float32)
_KW = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF)


def sim3kf_case(seed, flags):
    """(KF with matched / matched_q, pts, q) for SearchByProjection(KF,Scw)."""
    from matchdata import map_scene
    KF, pts = map_scene(seed, flags=flags)
    n = len(KF["x"])
    rng = np.random.Generator(np.random.PCG64(seed + 17))
    mq = np.full(n, -1, np.int32)
    if flags:
        k = rng.permutation(n)
        mq[k[:40]] = -2
        mq[k[40:80]] = rng.permutation(n)[:40]
    KF["matched_q"], KF["matched"] = mq, (mq != -1).astype(np.uint8)
    valid, u, v, lvl = po.map_projection(pts, _LSF, 8, bounds=_BOUNDS)
    valid &= (1 - pts["bad"])
    valid[mq[mq >= 0]] = 0
    return KF, pts, dict(valid=valid, u=u, v=v, level=lvl, desc=pts["desc"])


@pytest.mark.parametrize("th", [5, 10])
@pytest.mark.parametrize("flags", [False, True])
def test_search_by_projection_sim3kf(th, flags):
    total = 0
    for seed in range(3):
        KF, pts, q = sim3kf_case(600 + seed, flags)
        on, om, _ = po.o_search_by_projection_sim3kf(KF, q, th, **_KW)
        rn, rm = po.r_search_by_projection_sim3kf(KF, pts, th, log_scale_factor=_LSF, scw_scale=2.0 if seed else 1.0, **_KW)
        assert rn == on, (seed, rn, on)
        assert np.array_equal(rm, np.where(om >= 0, om, KF["matched_q"])), seed
        total += on
    assert total > 100


def fuse_case(seed, stereo, mbf=20.0):
    from matchdata import fuse_scene
    KF, pts = fuse_scene(seed, stereo=stereo)
    valid, u, v, lvl = po.map_projection(pts, _LSF, 8, bounds=_BOUNDS)
    valid &= (pts["state"] == 3)
    q = dict(valid=valid, u=u, v=v, level=lvl, desc=pts["desc"], ur=(u - np.float32(mbf)).astype(np.float32))
    return KF, pts, q


@pytest.mark.parametrize("th", [3.0, 6.0])
@pytest.mark.parametrize("stereo", [False, True])
def test_fuse(th, stereo):
    total = replaced = gated = 0
    for seed in range(3):
        KF, pts, q = fuse_case(700 + seed, stereo)
        _, mq, _ = po.o_window_best(1, KF, q, th, 50, inv_level_sigma2=_INV_SIGMA2, **_KW)
        gated += int(np.sum(po.o_window_best(0, KF, q, th, 50, **_KW)[1] != mq))
        o = po.o_fuse_apply(mq, pts["state"], pts["qid"], pts["obs"], KF["slot_state"], KF["slot_obs"])
        r = po.r_fuse(KF, pts, th, inv_level_sigma2=_INV_SIGMA2, log_scale_factor=_LSF, mbf=20.0, **_KW)
        assert r[0] == o[0], (seed, r[0], o[0])
        for a, b, name in zip(r[1:], o[1:], ("addedAt", "replacedBy", "slotReplacedBy", "slotHolder")):
            assert np.array_equal(a, b), (seed, name)
        total += o[0]
        replaced += int(np.sum(o[2] >= 0) + np.sum(o[3] >= 0))
    assert total > 100 and replaced > 10 and gated > 0  # the chi-square gate and both Replace directions were exercised


def fuse_sim3_case(seed):
    from matchdata import map_scene
    KF, pts = map_scene(seed, flags=True, noise=1.0)
    n = len(KF["x"])
    rng = np.random.Generator(np.random.PCG64(seed + 19))
    KF["slot_state"] = rng.choice(np.array([0, 1, 2], np.uint8), n, p=[0.5, 0.42, 0.08])
    sq = np.full(n, -1, np.int32)
    k = rng.permutation(n)[:40]
    sq[k] = rng.permutation(n)[:40]
    sq[k[pts["bad"][sq[k]] > 0]] = -1  # KeyFrame::GetMapPoints leaves bad points out; keep the case unambiguous
    KF["slot_query"] = sq
    valid, u, v, lvl = po.map_projection(pts, _LSF, 8, bounds=_BOUNDS)
    valid &= (1 - pts["bad"])
    valid[sq[sq >= 0]] = 0
    return KF, pts, dict(valid=valid, u=u, v=v, level=lvl, desc=pts["desc"])


@pytest.mark.parametrize("th", [4.0, 10.0])
def test_fuse_sim3(th):
    total = repl = 0
    for seed in range(3):
        KF, pts, q = fuse_sim3_case(800 + seed)
        _, mq, _ = po.o_window_best(0, KF, q, th, 50, **_KW)
        o = po.o_fuse_sim3_apply(mq, KF["slot_state"], KF["slot_query"])
        r = po.r_fuse_sim3(KF, pts, th, log_scale_factor=_LSF, scw_scale=4.0 if seed else 1.0, **_KW)
        assert r[0] == o[0], (seed, r[0], o[0])
        for a, b, name in zip(r[1:], o[1:], ("addedAt", "replacePoint", "slotHolder")):
            assert np.array_equal(a, b), (seed, name)
        total += o[0]
        repl += int(np.sum(o[2] >= 0))
    assert total > 100 and repl > 10


def sim3_queries(K1, K2, pre12):
    """Post-projection query arrays of both directions of SearchBySim3 (s12 = 1, R12 = I, t12 = 0)."""
    def direction(K, done):
        valid, u, v, lvl = po.map_projection(K, _LSF, 8, bounds=_BOUNDS, check_normal=False)
        valid &= (K["state"] == 3) & ~done
        return dict(valid=valid, u=u, v=v, level=lvl, desc=K["pdesc"])
    done1 = pre12 >= 0
    done2 = np.zeros(len(K2["x"]), bool)
    done2[pre12[done1]] = True
    return direction(K1, done1), direction(K2, done2)


@pytest.mark.parametrize("th", [7.5, 3.0])
@pytest.mark.parametrize("flags", [False, True])
def test_search_by_sim3(th, flags):
    from matchdata import sim3_scene
    total = onesided = 0
    for seed in range(3):
        K1, K2, pre12 = sim3_scene(900 + seed, flags=flags)
        q1, q2 = sim3_queries(K1, K2, pre12)
        _, m1, _ = po.o_window_best(0, K2, q1, th, 100, **_KW)
        _, m2, _ = po.o_window_best(0, K1, q2, th, 100, **_KW)
        on, om = po.o_sim3_agreement(m1, m2)
        rn, rm = po.r_search_by_sim3(K1, K2, pre12, th, log_scale_factor=_LSF, **_KW)
        assert rn == on, (seed, rn, on)
        assert np.array_equal(rm, np.where(om >= 0, om, pre12)), seed
        total += on
        onesided += int(np.sum(m1 >= 0)) - on
    assert total > 300 and (onesided > 0 or not flags)


def test_distinctive_descriptor_restatement():
    """eaoo_distinctive_descriptor (src/MapPoint.cc:273-301) against a numpy restatement: N x N distances, row medians at
    index int(0.5*(N-1)) of the sorted row, first minimum."""
    rng = np.random.Generator(np.random.PCG64(77))
    for n in [1, 2, 3, 4, 7, 20, 33, 64, 150]:
        base = rng.integers(0, 256, size=(1, 32), dtype=np.uint8)
        d = base ^ np.packbits((rng.random((n, 256)) < rng.uniform(0.02, 0.3, (n, 1))).astype(np.uint8), axis=1)
        if n > 3:
            d[n // 2] = d[0]  # ties between rows
        D = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(2)
        med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        best, m = po.o_distinctive_descriptor(d)
        assert best == int(np.argmin(med)) and m == int(med.min()), n
    assert po.o_distinctive_descriptor(np.zeros((0, 32), np.uint8))[0] == -1


def test_distinctive_descriptor_vs_reference_mappoint():
    """The restatement against the UNMODIFIED reference MapPoint.cc (oracle/_ref/libmappoint_ref.so): same chosen
    observation, with ties between rows, duplicate descriptors and bad keyframes (skipped at src/MapPoint.cc:265)."""
    rng = np.random.Generator(np.random.PCG64(78))
    for trial, n in enumerate([1, 2, 3, 4, 5, 6, 9, 16, 31, 32, 33, 64, 101, 200] * 3):
        base = rng.integers(0, 256, size=(1, 32), dtype=np.uint8)
        d = base ^ np.packbits((rng.random((n, 256)) < rng.uniform(0.0, 0.3, (n, 1))).astype(np.uint8), axis=1)
        if n > 3:
            d[n - 1] = d[1]
        bad = (rng.random(n) < 0.2).astype(np.uint8) if trial % 3 == 2 else np.zeros(n, np.uint8)
        ri, rdesc = po.r_distinctive_descriptor(d, bad)
        good = np.nonzero(bad == 0)[0]
        ob, _ = po.o_distinctive_descriptor(d[good])
        if ob < 0:
            assert ri == -1 and np.all(rdesc == 0xA5), (trial, n)  # no usable observation: descriptor untouched
        else:
            assert ri == good[ob] and np.array_equal(rdesc, d[good[ob]]), (trial, n, ri, good[ob])


def test_predict_scale_of_the_real_mappoint_equals_the_standin():
    """matchshim's MapPoint::PredictScale (used by libmatch_ref.so) against the reference's own MapPoint.cc."""
    L, M = po.match_ref_lib(), po.mappoint_ref_lib()
    rng = np.random.Generator(np.random.PCG64(79))
    lsf = float(np.log(np.float32(1.2)))
    for _ in range(2000):
        d0, sc, cur = (float(np.float32(v)) for v in (rng.uniform(0.5, 20), rng.uniform(1, 3.6), rng.uniform(0.3, 40)))
        max_dist = float(np.float32(d0) * np.float32(sc))  # MapPoint(Pos, pMap, pFrame, idxF): mfMaxDistance = dist*levelScaleFactor
        assert M.mpref_predict_scale(d0, sc, cur, lsf) == L.mref_predict_scale(max_dist, cur, lsf)
