"""GPU matcher parity: libeaof_orb's Hamming search (C ABI, include/eaof_match.h) against the oracle restatement of
src/ORBmatcher.cc — match indices, distances and match counts bit-exact, including greedy exclusion, ratio test,
TH_LOW/TH_HIGH and rotation-histogram pruning."""
import numpy as np
import pytest

from matchdata import frame_features, planted_pair, projected_last, random_nodes, tri_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def matcher_factory():
    import eaof
    made = []

    def make(ratio, ori=True, max_features=4096, max_pairs=1):
        m = eaof.ORBmatcher(ratio, ori, max_features=max_features, max_pairs=max_pairs)
        made.append(m)
        return m
    yield make
    for m in made:
        m.close()


def test_descriptor_distance(matcher_factory):
    from oracle import pyoracle as po
    rng = np.random.Generator(np.random.PCG64(1))
    a = rng.integers(0, 256, size=(5000, 32), dtype=np.uint8)
    b = rng.integers(0, 256, size=(5000, 32), dtype=np.uint8)
    a[:10] = 0; b[:5] = 255; b[5:10] = 0
    m = matcher_factory(0.6)
    d = m.DescriptorDistance(a, b)
    assert np.array_equal(d, np.unpackbits(a ^ b, axis=1).sum(1))
    assert np.array_equal(d[:200], po.o_hamming(a[:200], b[:200]))


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("ratio", [0.6, 0.75, 0.9])
def test_bow_single_node_bruteforce(matcher_factory, mode, ratio):
    import eaof
    from oracle import pyoracle as po
    m = matcher_factory(ratio)
    for seed, (nq, nt) in enumerate([(2000, 2000), (1000, 1500), (300, 7), (1, 1), (64, 129)]):
        q, aq, t, at = planted_pair(nq, nt, seed, dup=5 if nt > 50 else 0)
        nodes_q = eaof.csr_from_nodes(np.zeros(nq, int))
        nodes_t = eaof.csr_from_nodes(np.zeros(nt, int))
        rng = np.random.Generator(np.random.PCG64(100 + seed))
        vq = (rng.random(nq) > 0.1).astype(np.uint8)
        vt = (rng.random(nt) > 0.1).astype(np.uint8)
        n, match, dist = m.SearchByBoW(mode, q, aq, vq, nodes_q, t, at, vt, nodes_t)
        on, omatch, odist = po.o_search_by_bow(mode, ratio, True, q, aq, vq, nodes_q, t, at, vt, nodes_t)
        assert n == on, (seed, n, on)
        assert np.array_equal(match, omatch) and np.array_equal(dist, odist)
    assert on > 0


@pytest.mark.parametrize("mode", [0, 1])
def test_bow_multi_node_and_no_orientation(matcher_factory, mode):
    import eaof
    from oracle import pyoracle as po
    for ori in (True, False):
        m = matcher_factory(0.7, ori)
        for seed in range(4):
            q, aq, t, at = planted_pair(1200, 1100, 40 + seed, flip=0.05)
            # give planted partners a good chance of sharing a node: nodes from a hash of the first descriptor byte pair
            nq = (q[:, 0].astype(int) // 32) * 5 + 3
            ntd = (t[:, 0].astype(int) // 32) * 5 + 3
            nq[::17] = -1
            ntd[::13] = 99 + seed  # a node only T has
            nodes_q, nodes_t = eaof.csr_from_nodes(nq), eaof.csr_from_nodes(ntd)
            n, match, dist = m.SearchByBoW(mode, q, aq, None, nodes_q, t, at, None, nodes_t)
            on, omatch, odist = po.o_search_by_bow(mode, 0.7, ori, q, aq, None, nodes_q, t, at, None, nodes_t)
            assert n == on and np.array_equal(match, omatch) and np.array_equal(dist, odist)
        assert on > 20


def test_bow_heavy_ties_force_the_rescan_path(matcher_factory):
    """All-equal / all-zero / all-one descriptor sets: every query has thousands of candidates below the near-list
    threshold, so phase 2 must take the exact re-scan path; first-wins tie-breaking and greedy exclusion decide."""
    import eaof
    from oracle import pyoracle as po
    m = matcher_factory(0.9)
    rng = np.random.Generator(np.random.PCG64(9))
    cases = []
    z = np.zeros((300, 32), np.uint8)
    cases.append((z, z.copy()))
    cases.append((np.full((200, 32), 255, np.uint8), np.full((260, 32), 255, np.uint8)))
    base = rng.integers(0, 256, size=(1, 32), dtype=np.uint8)
    near = np.repeat(base, 400, axis=0)
    near[np.arange(400), rng.integers(0, 32, 400)] ^= (1 << rng.integers(0, 8, 400)).astype(np.uint8)
    cases.append((near[:180], near[150:]))
    for q, t in cases:
        aq = rng.uniform(0, 360, len(q)).astype(np.float32)
        at = rng.uniform(0, 360, len(t)).astype(np.float32)
        for mode in (0, 1):
            for ratio_m in (m, ):
                nodes_q = eaof.csr_from_nodes(np.zeros(len(q), int))
                nodes_t = eaof.csr_from_nodes(np.zeros(len(t), int))
                n, match, dist = ratio_m.SearchByBoW(mode, q, aq, None, nodes_q, t, at, None, nodes_t)
                on, omatch, odist = po.o_search_by_bow(mode, 0.9, True, q, aq, None, nodes_q, t, at, None, nodes_t)
                assert n == on and np.array_equal(match, omatch) and np.array_equal(dist, odist)


def test_bow_empty_inputs(matcher_factory):
    import eaof
    m = matcher_factory(0.6)
    e = np.zeros((0, 32), np.uint8)
    ea = np.zeros(0, np.float32)
    q, aq, t, at = planted_pair(10, 10, 1)
    empty_nodes = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    n, match, _ = m.SearchByBoW(0, e, ea, None, empty_nodes, t, at, None, eaof.csr_from_nodes(np.zeros(10, int)))
    assert n == 0 and np.all(match == -1) and len(match) == 10
    n, match, _ = m.SearchByBoW(0, q, aq, None, eaof.csr_from_nodes(np.zeros(10, int)), e, ea, None, empty_nodes)
    assert n == 0 and len(match) == 0


@pytest.fixture(scope="module")
def extracted(frames640):
    import eaof
    ex = eaof.ORBextractor(width=640, height=480, max_batch=8)
    res = ex.extract_batch(frames640)
    sf = ex.GetScaleFactors()
    yield ex, res, sf
    ex.close()


def _bounds():
    return (0.0, 640.0, 0.0, 480.0), (np.float32(64) / np.float32(640), np.float32(48) / np.float32(480))


@pytest.mark.parametrize("th", [15.0, 7.0, 30.0])
def test_projection_consecutive_frames(matcher_factory, extracted, th):
    """BASELINE configs[1]: consecutive frames of the synthetic sequence drift by (2,1) px, so a Last keypoint at
    (x,y) projects to (x-2, y-1) in Cur."""
    from oracle import pyoracle as po
    ex, res, sf = extracted
    bounds, ginv = _bounds()
    m = matcher_factory(0.9)
    tot = 0
    for f in range(1, len(res)):
        cur = frame_features(*res[f])
        last = projected_last(*res[f - 1], -2.0, -1.0)
        n, match, dist = m.SearchByProjection(cur, last, th, bounds=bounds, grid_inv=ginv, scale_factors=sf)
        on, omatch, odist = po.o_search_by_projection(cur, last, th, True, bounds=bounds, grid_inv=ginv, scale_factors=sf)
        assert n == on
        assert np.array_equal(match, omatch) and np.array_equal(dist, odist)
        tot += n
    assert tot > 500


def test_projection_flags_stereo_and_modes(matcher_factory, extracted):
    from oracle import pyoracle as po
    ex, res, sf = extracted
    bounds, ginv = _bounds()
    rng = np.random.Generator(np.random.PCG64(5))
    for ori in (True, False):
        m = matcher_factory(0.9, ori)
        for mode in (0, 1, 2):
            cur = frame_features(*res[2])
            last = projected_last(*res[1], -2.0, -1.0)
            nc, nl = len(cur["x"]), len(last["u"])
            cur["taken"] = (rng.random(nc) < 0.2).astype(np.uint8)
            cur["uright"] = np.where(rng.random(nc) < 0.5, cur["x"] - 30 + rng.normal(0, 8, nc), -1).astype(np.float32)
            last["valid"] = (rng.random(nl) < 0.9).astype(np.uint8)
            last["obs"] = (rng.random(nl) < 0.7).astype(np.uint8)
            last["invz"] = np.where(rng.random(nl) < 0.05, -0.5, 0.75).astype(np.float32)
            last["u"][::50] += 700  # outside the image
            kw = dict(bounds=bounds, grid_inv=ginv, scale_factors=sf, mbf=40.0, search_mode=mode)
            n, match, dist = m.SearchByProjection(cur, last, 15.0, **kw)
            on, omatch, odist = po.o_search_by_projection(cur, last, 15.0, ori, **kw)
            assert n == on and np.array_equal(match, omatch) and np.array_equal(dist, odist)


def test_projection_crowded_window_takes_rescan_path(matcher_factory):
    """Many identical descriptors inside one search window: the four kept candidates get taken by earlier queries."""
    from oracle import pyoracle as po
    rng = np.random.Generator(np.random.PCG64(3))
    n = 300
    cur = dict(x=(320 + rng.uniform(-6, 6, n)).astype(np.float32), y=(240 + rng.uniform(-6, 6, n)).astype(np.float32),
               octave=np.zeros(n, np.int32), angle=rng.uniform(0, 360, n).astype(np.float32),
               desc=np.repeat(rng.integers(0, 256, size=(1, 32), dtype=np.uint8), n, axis=0))
    last = dict(u=(320 + rng.uniform(-3, 3, n)).astype(np.float32), v=(240 + rng.uniform(-3, 3, n)).astype(np.float32),
                octave=np.zeros(n, np.int32), angle=rng.uniform(0, 360, n).astype(np.float32), desc=cur["desc"].copy())
    bounds, ginv = _bounds()
    sf = np.array([1.0, 1.2], np.float32)
    m = matcher_factory(0.9)
    r = m.SearchByProjection(cur, last, 15.0, bounds=bounds, grid_inv=ginv, scale_factors=sf)
    o = po.o_search_by_projection(cur, last, 15.0, True, bounds=bounds, grid_inv=ginv, scale_factors=sf)
    assert r[0] == o[0] and np.array_equal(r[1], o[1]) and np.array_equal(r[2], o[2])
    assert r[0] > 4


def test_batched_device_paths_equal_single_pair_api(matcher_factory, extracted, frames640):
    """eaof_match_projection_batch_device / eaof_match_bruteforce_batch_device on extractor results resident on the GPU."""
    import torch
    import eaof
    from oracle import pyoracle as po
    ex, res, sf = extracted
    nfr = len(res)
    ex.extract_batch(frames640)  # leaves the results of all frames on the device
    cap = ex.cap
    m = matcher_factory(0.9, True, max_features=cap, max_pairs=16)
    npairs = nfr - 1
    d_match = torch.full((npairs, cap), -7, dtype=torch.int32, device="cuda")
    d_dist = torch.full((npairs, cap), -7, dtype=torch.int32, device="cuda")
    d_n = torch.zeros(npairs, dtype=torch.int32, device="cuda")
    lastf, curf = np.arange(0, nfr - 1), np.arange(1, nfr)
    m.projection_batch_device(ex, lastf, curf, np.full(npairs, -2.0), np.full(npairs, -1.0), 15.0, d_match.data_ptr(),
                              d_dist.data_ptr(), d_n.data_ptr())
    m.sync()
    bounds, ginv = _bounds()
    hm, hd, hn = d_match.cpu().numpy(), d_dist.cpu().numpy(), d_n.cpu().numpy()
    for p in range(npairs):
        cur = frame_features(*res[p + 1])
        last = projected_last(*res[p], -2.0, -1.0)
        on, omatch, odist = po.o_search_by_projection(cur, last, 15.0, True, bounds=bounds, grid_inv=ginv, scale_factors=sf)
        nc = len(cur["x"])
        assert hn[p] == on
        assert np.array_equal(hm[p, :nc], omatch) and np.array_equal(hd[p, :nc], odist)
    # brute force over the same device-resident descriptors: all ordered pairs (i, j), i != j
    kps_ptr, desc_ptr, cnt_ptr, cap2 = ex.device_results()
    assert cap2 == cap
    ang = torch.zeros((nfr, cap), dtype=torch.float32, device="cuda")
    for f in range(nfr):
        ang[f, :len(res[f][0])] = torch.from_numpy(res[f][0]["angle"].copy()).cuda()
    pq = [i for i in range(nfr) for j in range(nfr) if i != j][:16]
    pt = [j for i in range(nfr) for j in range(nfr) if i != j][:16]
    bm = torch.full((len(pq), cap), -7, dtype=torch.int32, device="cuda")
    bd = torch.full((len(pq), cap), -7, dtype=torch.int32, device="cuda")
    bn = torch.zeros(len(pq), dtype=torch.int32, device="cuda")
    for mode in (0, 1):
        m.bruteforce_batch_device(mode, pq, pt, desc_ptr, ang.data_ptr(), cnt_ptr, cap, bm.data_ptr(), bd.data_ptr(), bn.data_ptr())
        m.sync()
        hm, hd, hn = bm.cpu().numpy(), bd.cpu().numpy(), bn.cpu().numpy()
        for k, (i, j) in enumerate(zip(pq, pt)):
            (kq, dq), (kt, dt) = res[i], res[j]
            nodes_q = eaof.csr_from_nodes(np.zeros(len(kq), int))
            nodes_t = eaof.csr_from_nodes(np.zeros(len(kt), int))
            on, omatch, odist = po.o_search_by_bow(mode, 0.9, True, dq, kq["angle"], None, nodes_q, dt, kt["angle"], None, nodes_t)
            nout = len(kt) if mode == 0 else len(kq)
            assert hn[k] == on
            assert np.array_equal(hm[k, :nout], omatch) and np.array_equal(hd[k, :nout], odist)


@pytest.mark.parametrize("only_stereo", [False, True])
@pytest.mark.parametrize("ori", [True, False])
def test_triangulation(matcher_factory, only_stereo, ori):
    """SearchForTriangulation (src/ORBmatcher.cc:657-823): later-candidate-wins ties, epipole and epipolar-line gates."""
    from oracle import pyoracle as po
    m = matcher_factory(0.6, ori)
    total = 0
    for seed, (n1, n2, nn) in enumerate([(500, 600, 8), (2000, 2000, 40), (300, 900, 1), (5, 3, 1)]):
        k1, k2, F12, sf, ls = tri_inputs(70 + seed, n1, n2, nn)
        for epipole in ((320.0, 240.0), (-1e4, -1e4)):
            n, match, dist = m.SearchForTriangulation(k1, k2, F12, epipole, sf, ls, only_stereo)
            on, om, od = po.o_search_for_triangulation(k1, k2, F12, epipole, sf, ls, only_stereo, ori)
            assert n == on, (seed, n, on)
            assert np.array_equal(match, om) and np.array_equal(dist, od)
            total += on
    assert total > 100
    # empty sides
    k1, k2, F12, sf, ls = tri_inputs(1, 10, 10, 1)
    e = dict(k1, desc=k1["desc"][:0], x=k1["x"][:0], y=k1["y"][:0], angle=k1["angle"][:0], free=k1["free"][:0],
             stereo=k1["stereo"][:0], nodes=(np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32)))
    n, match, _ = m.SearchForTriangulation(e, k2, F12, (0.0, 0.0), sf, ls)
    assert n == 0 and len(match) == 0


_SF = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
_BOUNDS = (0.0, 640.0, 0.0, 480.0)
_GINV = (np.float32(64) / np.float32(640), np.float32(48) / np.float32(480))


@pytest.mark.parametrize("th", [1.0, 3.0, 5.0])
@pytest.mark.parametrize("ratio", [0.8, 0.6])
@pytest.mark.parametrize("case", ["plain", "flags", "stereo"])
def test_projection_mappoints(matcher_factory, th, ratio, case):
    """SearchByProjection(Frame&, vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129) through eaof_match_windows:
    same-level ratio rule, stereo gate, greedy exclusion, match count."""
    from matchdata import window_scene
    from oracle import pyoracle as po
    m = matcher_factory(ratio)
    total = 0
    for seed in range(3):
        F, mp = window_scene(300 + seed, stereo=case == "stereo", flags=case != "plain")
        kw = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF)
        n, match, dist = m.SearchByProjectionMapPoints(F, mp, th, **kw)
        on, om, od = po.o_search_by_projection_mappoints(F, mp, th, ratio, **kw)
        assert n == on, (seed, n, on)
        assert np.array_equal(match, om) and np.array_equal(dist, od)
        total += on
    assert total > 100


def test_projection_mappoints_crowded_windows_rescan(matcher_factory):
    """Dozens of near-identical features inside every window and most of them taken: the kept list runs dry and the
    exact re-scan decides."""
    from oracle import pyoracle as po
    rng = np.random.Generator(np.random.PCG64(9))
    n_f, n_q = 600, 120
    base = rng.integers(0, 256, size=(1, 32), dtype=np.uint8)
    desc = np.repeat(base, n_f, 0) ^ np.packbits((rng.random((n_f, 256)) < 0.02).astype(np.uint8), axis=1)
    F = dict(x=rng.uniform(300, 340, n_f).astype(np.float32), y=rng.uniform(220, 260, n_f).astype(np.float32),
             octave=rng.integers(0, 2, n_f).astype(np.int32), desc=desc, taken=(rng.random(n_f) < 0.5).astype(np.uint8))
    mp = dict(x=rng.uniform(300, 340, n_q).astype(np.float32), y=rng.uniform(220, 260, n_q).astype(np.float32),
              level=np.ones(n_q, np.int32), desc=np.repeat(base, n_q, 0), cos=np.full(n_q, 0.9, np.float32))
    kw = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF)
    for ratio in (0.8, 1.0):
        m = matcher_factory(ratio)
        n, match, dist = m.SearchByProjectionMapPoints(F, mp, 5.0, **kw)
        on, om, od = po.o_search_by_projection_mappoints(F, mp, 5.0, ratio, **kw)
        assert n == on and np.array_equal(match, om) and np.array_equal(dist, od), (ratio, n, on)
    assert on > 20


@pytest.mark.parametrize("th", [7.0, 15.0])
@pytest.mark.parametrize("orb_dist", [50, 100])
@pytest.mark.parametrize("case", ["plain", "flags", "no_ori"])
def test_projection_keyframe(matcher_factory, th, orb_dist, case):
    """SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1474-1601)."""
    from matchdata import kf_scene
    from oracle import pyoracle as po
    m = matcher_factory(0.9, case != "no_ori")
    lsf = float(np.log(np.float32(1.2)))
    total = 0
    for seed in range(3):
        F, kf = kf_scene(400 + seed, flags=case == "flags")
        valid, u, v, lvl = po.kf_projection(kf, lsf, 8)
        kq = dict(valid=valid, u=u, v=v, level=lvl, angle=kf["angle"], desc=kf["desc"])
        kw = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF)
        n, match, dist = m.SearchByProjectionKF(F, kq, th, orb_dist, **kw)
        on, om, od = po.o_search_by_projection_kf(F, kq, th, orb_dist, case != "no_ori", **kw)
        assert n == on, (seed, n, on)
        assert np.array_equal(match, om) and np.array_equal(dist, od)
        total += on
    assert total > 30


@pytest.mark.parametrize("window", [10, 100])
@pytest.mark.parametrize("ratio", [0.9, 0.6])
@pytest.mark.parametrize("ori", [True, False])
def test_initialization(matcher_factory, window, ratio, ori):
    """SearchForInitialization (src/ORBmatcher.cc:405-520) incl. match stealing and the vbPrevMatched update."""
    from matchdata import init_scene
    from oracle import pyoracle as po
    m = matcher_factory(ratio, ori)
    total = 0
    for seed in range(3):
        F1, F2, prev = init_scene(500 + seed)
        n, m12, pm = m.SearchForInitialization(F1, F2, prev, window, bounds=_BOUNDS, grid_inv=_GINV)
        on, om, op = po.o_search_for_initialization(F1, F2, prev, window, ratio, ori, bounds=_BOUNDS, grid_inv=_GINV)
        assert n == on, (seed, n, on)
        assert np.array_equal(m12, om) and np.array_equal(pm, op)
        total += on
    assert total > 300


# ---- map-side matchers (src/ORBmatcher.cc:290-403, :825-1326) and ComputeDistinctiveDescriptors ---------------------------
def _win_query(q, th, sf):
    lvl = q["level"].astype(np.int32)
    out = dict(u=q["u"], v=q["v"], radius=(np.float32(th) * sf[lvl]).astype(np.float32), min_level=lvl - 1, max_level=lvl,
               desc=q["desc"], valid=q["valid"])
    if "ur" in q:
        out["ur"] = q["ur"]
    return out


@pytest.mark.parametrize("th", [5, 10])
@pytest.mark.parametrize("flags", [False, True])
def test_projection_sim3_keyframe(matcher_factory, th, flags):
    """SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) = greedy window search, TH_LOW, no rotation check."""
    from oracle import pyoracle as po
    from test_oracle_matcher_vs_ref import _KW, _SF, sim3kf_case
    m = matcher_factory(0.75)
    total = 0
    for seed in range(3):
        KF, pts, q = sim3kf_case(600 + seed, flags)
        on, om, od = po.o_search_by_projection_sim3kf(KF, q, th, **_KW)
        F = dict(x=KF["x"], y=KF["y"], octave=KF["octave"], desc=KF["desc"], taken=KF["matched"])
        n, match, dist = m.SearchWindows(0, F, _win_query(q, th, _SF), 50, 0, False, bounds=_KW["bounds"], grid_inv=_KW["grid_inv"])
        assert n == on, (seed, n, on)
        assert np.array_equal(match, om) and np.array_equal(dist, od), seed
        total += on
    assert total > 100


@pytest.mark.parametrize("th", [3.0, 6.0])
@pytest.mark.parametrize("stereo", [False, True])
def test_fuse_search(matcher_factory, th, stereo):
    """The search step of Fuse(KeyFrame*, vpMapPoints, th) with the reprojection chi-square gate."""
    from oracle import pyoracle as po
    from test_oracle_matcher_vs_ref import _INV_SIGMA2, _KW, _SF, fuse_case
    m = matcher_factory(0.6)
    total = 0
    for seed in range(3):
        KF, pts, q = fuse_case(700 + seed, stereo)
        on, om, od = po.o_window_best(1, KF, q, th, 50, inv_level_sigma2=_INV_SIGMA2, **_KW)
        n, match, dist = m.SearchWindowsIndependent(1, KF, _win_query(q, th, _SF), 50, bounds=_KW["bounds"],
                                                    grid_inv=_KW["grid_inv"], inv_level_sigma2=_INV_SIGMA2)
        assert n == on, (seed, n, on)
        assert np.array_equal(match, om) and np.array_equal(dist, od), seed
        total += on
    assert total > 100


@pytest.mark.parametrize("th,accept", [(4.0, 50), (10.0, 50), (7.5, 100)])
def test_independent_windows_no_gate(matcher_factory, th, accept):
    """The search step of Fuse(KeyFrame*, Scw, ...) (TH_LOW) and of one SearchBySim3 direction (TH_HIGH)."""
    from oracle import pyoracle as po
    from test_oracle_matcher_vs_ref import _KW, _SF, fuse_sim3_case
    m = matcher_factory(0.6)
    total = 0
    for seed in range(3):
        KF, pts, q = fuse_sim3_case(800 + seed)
        on, om, od = po.o_window_best(0, KF, q, th, accept, **_KW)
        n, match, dist = m.SearchWindowsIndependent(0, KF, _win_query(q, th, _SF), accept, bounds=_KW["bounds"], grid_inv=_KW["grid_inv"])
        assert n == on and np.array_equal(match, om) and np.array_equal(dist, od), (seed, n, on)
        total += on
    assert total > 100


def test_search_by_sim3_both_directions(matcher_factory):
    from matchdata import sim3_scene
    from oracle import pyoracle as po
    from test_oracle_matcher_vs_ref import _KW, _SF, sim3_queries
    m = matcher_factory(0.75)
    for seed in range(2):
        K1, K2, pre12 = sim3_scene(900 + seed, flags=True)
        q1, q2 = sim3_queries(K1, K2, pre12)
        g1 = m.SearchWindowsIndependent(0, K2, _win_query(q1, 7.5, _SF), 100, bounds=_KW["bounds"], grid_inv=_KW["grid_inv"])[1]
        g2 = m.SearchWindowsIndependent(0, K1, _win_query(q2, 7.5, _SF), 100, bounds=_KW["bounds"], grid_inv=_KW["grid_inv"])[1]
        o1 = po.o_window_best(0, K2, q1, 7.5, 100, **_KW)[1]
        o2 = po.o_window_best(0, K1, q2, 7.5, 100, **_KW)[1]
        assert np.array_equal(g1, o1) and np.array_equal(g2, o2), seed
        assert po.o_sim3_agreement(g1, g2)[0] > 100


def test_independent_windows_edge_cases(matcher_factory):
    m = matcher_factory(0.6)
    e = np.zeros(0, np.float32)
    KF = dict(x=e, y=e, octave=np.zeros(0, np.int32), desc=np.zeros((0, 32), np.uint8))
    q = dict(u=np.float32([10]), v=np.float32([10]), radius=np.float32([5]), min_level=np.int32([0]), max_level=np.int32([1]),
             desc=np.zeros((1, 32), np.uint8))
    n, match, dist = m.SearchWindowsIndependent(0, KF, q, 50, bounds=(0, 640, 0, 480), grid_inv=(0.1, 0.1))
    assert n == 0 and match.tolist() == [-1]
    # two identical targets in one cell: the first in grid order wins; a target exactly at the radius is outside
    KF = dict(x=np.float32([10, 10, 15]), y=np.float32([10, 10, 10]), octave=np.int32([0, 0, 0]), desc=np.zeros((3, 32), np.uint8))
    KF["desc"][2] = 0
    n, match, dist = m.SearchWindowsIndependent(0, KF, q, 50, bounds=(0, 640, 0, 480), grid_inv=(0.1, 0.1))
    assert n == 1 and match.tolist() == [0] and dist.tolist() == [0]
    q["u"] = np.float32([20])  # only target 2 at |dx| = 5 = r: not < r
    n, match, _ = m.SearchWindowsIndependent(0, KF, q, 50, bounds=(0, 640, 0, 480), grid_inv=(0.1, 0.1))
    assert n == 0 and match.tolist() == [-1]


def test_distinctive_descriptors(matcher_factory):
    from oracle import pyoracle as po
    rng = np.random.Generator(np.random.PCG64(5))
    sizes = [0, 1, 2, 3, 4, 5, 8, 31, 32, 33, 64, 100, 257, 600] + rng.integers(1, 40, 400).tolist()
    descs = []
    for n in sizes:
        base = rng.integers(0, 256, size=(1, 32), dtype=np.uint8)
        d = base ^ np.packbits((rng.random((n, 256)) < rng.uniform(0.0, 0.3, (n, 1))).astype(np.uint8), axis=1)
        if n > 3:
            d[n - 1] = d[0]
        descs.append(d)
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    allrows = np.concatenate(descs)
    for max_features in (4096, 400):  # the small handle forces the chunked path
        m = matcher_factory(0.6, max_features=max_features)
        best, med = m.DistinctiveDescriptors(starts, allrows)
        for p, d in enumerate(descs):
            ob, om = po.o_distinctive_descriptor(d)
            assert best[p] == ob and (ob < 0 or med[p] == om), (p, len(d), best[p], ob)


def test_tensor_core_hamming_equals_popc_kernel(matcher_factory, monkeypatch):
    """Phase 1 of the brute-force matcher has two kernels: k_bow_dense (XOR + POPC) and k_bow_dense_umma (tcgen05.mma kind::i8 on
    +-1-expanded descriptors, near lists scanned out of TMEM).  Ragged blocks (2000 / 1999 / 257 / 130 / 128 / 1 / 0 descriptors),
    planted near-duplicates with few flipped bits (lists longer than NEAR_K -> exact re-scan in phase 2) and plain noise: both
    kernels must give identical matches, distances and counts, and equal the oracle's SearchByBoW on single pairs."""
    import torch
    import eaof
    from oracle import pyoracle as po
    rng = np.random.default_rng(123)
    stride = 2000
    counts = np.array([2000, 1999, 257, 130, 128, 1, 0, 2000, 777, 2000], np.int32)
    nb = len(counts)
    desc = rng.integers(0, 256, (nb, stride, 32), dtype=np.uint8)
    for b in range(1, nb):  # most descriptors of block b are noisy copies of descriptors of block b-1
        if counts[b - 1] == 0:
            continue
        for i in range(counts[b]):
            if i % 4 == 3:
                continue
            j = (i * 7) % counts[b - 1] if i % 4 else (i // 8) % counts[b - 1]  # every 4th: many queries share a target
            d = desc[b - 1, j].copy()
            flips = rng.integers(0, 256, rng.integers(0, 60))
            for f in flips:
                d[f >> 3] ^= np.uint8(1 << (f & 7))
            desc[b, i] = d
    angle = rng.uniform(0, 360, (nb, stride)).astype(np.float32)
    pq = [b for b in range(1, nb)] + [0, 7, 9, 1, 8, 2]
    pt = [b - 1 for b in range(1, nb)] + [7, 9, 0, 8, 1, 9]
    d_desc, d_angle, d_cnt = torch.from_numpy(desc).cuda(), torch.from_numpy(angle).cuda(), torch.from_numpy(counts).cuda()
    out = {}
    for knob in ("1", "0"):
        monkeypatch.setenv("EAOF_BOW_UMMA", knob)
        m = matcher_factory(0.8, True, max_features=stride, max_pairs=32)
        res = []
        for mode in (0, 1):
            bm = torch.full((len(pq), stride), -7, dtype=torch.int32, device="cuda")
            bd = torch.full((len(pq), stride), -7, dtype=torch.int32, device="cuda")
            bn = torch.zeros(len(pq), dtype=torch.int32, device="cuda")
            m.bruteforce_batch_device(mode, pq, pt, d_desc.data_ptr(), d_angle.data_ptr(), d_cnt.data_ptr(), stride, bm.data_ptr(),
                                      bd.data_ptr(), bn.data_ptr())
            m.sync()
            res.append((bm.cpu().numpy(), bd.cpu().numpy(), bn.cpu().numpy()))
        out[knob] = res
    for mode in (0, 1):
        for a, b in zip(out["1"][mode], out["0"][mode]):
            assert np.array_equal(a, b), f"tensor-core and POPC kernels differ (mode {mode})"
        hm, hd, hn = out["1"][mode]
        assert hn.sum() > 300
        for k in (0, 1, 2, 9, 11):
            i, j = pq[k], pt[k]
            nq, nt = int(counts[i]), int(counts[j])
            nodes_q, nodes_t = eaof.csr_from_nodes(np.zeros(nq, int)), eaof.csr_from_nodes(np.zeros(nt, int))
            on, omatch, odist = po.o_search_by_bow(mode, 0.8, True, desc[i, :nq], angle[i, :nq], None, nodes_q, desc[j, :nt], angle[j, :nt],
                                                   None, nodes_t)
            nout = nt if mode == 0 else nq
            assert hn[k] == on
            assert np.array_equal(hm[k, :nout], omatch) and np.array_equal(hd[k, :nout], odist)


def test_bow_resolve_without_the_proposal_owners(matcher_factory):
    """A matcher whose max_features exceeds what phase 2's shared-memory owner array covers (11000) resolves every batch
    sequentially (BowArgs::spec == 0): same answers as the oracle, as for the 32-at-a-time resolution."""
    import eaof
    from oracle import pyoracle as po
    m = matcher_factory(0.8, True, max_features=12000)
    total = 0
    for mode in (0, 1):
        for seed, (nq, nt) in enumerate([(2000, 2000), (300, 900)]):
            q, aq, t, at = planted_pair(nq, nt, 60 + seed, dup=5)
            nodes_q, nodes_t = eaof.csr_from_nodes(np.zeros(nq, int)), eaof.csr_from_nodes(np.zeros(nt, int))
            n, match, dist = m.SearchByBoW(mode, q, aq, None, nodes_q, t, at, None, nodes_t)
            on, omatch, odist = po.o_search_by_bow(mode, 0.8, True, q, aq, None, nodes_q, t, at, None, nodes_t)
            assert n == on and np.array_equal(match, omatch) and np.array_equal(dist, odist)
            total += on
    assert total > 100
