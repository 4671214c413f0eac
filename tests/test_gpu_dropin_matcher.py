"""The drop-in ORB_SLAM2::ORBmatcher methods (eao-fusion_b200/dropin/ORBmatcher.cc, compiled against the reference's own
ORBmatcher.h) behind the same C harness and the same array-backed Frame/KeyFrame/MapPoint stand-ins that drive the
UNMODIFIED reference ORBmatcher.cc: called through the class interface, the GPU path must leave exactly the map-point
assignments, match vectors and counts the reference leaves."""
import os

import numpy as np
import pytest

from matchdata import init_scene, kf_scene, planted_pair, random_nodes, tri_inputs, window_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN_SO = os.path.join(ROOT, "tests", "cpp", "_build", "libmatch_dropin.so")

_SF = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
_BOUNDS = (0.0, 640.0, 0.0, 480.0)
_GINV = (np.float32(64) / np.float32(640), np.float32(48) / np.float32(480))


@pytest.fixture(scope="module")
def libs():
    from oracle import pyoracle as po
    if not os.path.exists(DROPIN_SO):
        pytest.fail(f"{DROPIN_SO} missing: run `make` where /root/reference is present")
    return po, po.match_harness_lib(DROPIN_SO)


def _csr(node):
    import eaof
    return eaof.csr_from_nodes(node)


def test_constants_and_descriptor_distance(libs):
    po, D = libs
    assert (D.mref_th_low(), D.mref_th_high(), D.mref_histo_length()) == (50, 100, 30)
    rng = np.random.Generator(np.random.PCG64(3))
    a = rng.integers(0, 256, size=(300, 32), dtype=np.uint8)
    b = rng.integers(0, 256, size=(300, 32), dtype=np.uint8)
    assert np.array_equal(po.r_hamming(a, b, L=D), po.r_hamming(a, b))


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("ratio", [0.6, 0.9])
def test_search_by_bow(libs, mode, ratio):
    po, D = libs
    for seed, (nq, nt, nn) in enumerate([(400, 500, 1), (600, 600, 12), (300, 7, 2), (1, 1, 1), (2000, 2000, 60)]):
        q, aq, t, at = planted_pair(nq, nt, 10 + seed, dup=5 if nt > 50 else 0)
        nodes_q = _csr(random_nodes(nq, nn, 20 + seed) if nn > 1 else np.zeros(nq, int))
        nodes_t = _csr(random_nodes(nt, nn, 30 + seed) if nn > 1 else np.zeros(nt, int))
        rng = np.random.Generator(np.random.PCG64(100 + seed))
        vq = (rng.random(nq) > 0.1).astype(np.uint8)
        vt = (rng.random(nt) > 0.1).astype(np.uint8)
        args = (mode, ratio, True, q, aq, vq, nodes_q, t, at, vt, nodes_t)
        rn, rm = po.r_search_by_bow(*args)
        dn, dm = po.r_search_by_bow(*args, L=D)
        assert dn == rn and np.array_equal(dm, rm), (seed, dn, rn)


@pytest.mark.parametrize("only_stereo", [False, True])
def test_search_for_triangulation(libs, only_stereo):
    po, D = libs
    for seed in range(3):
        k1, k2, F12, sf, ls = tri_inputs(50 + seed)
        for epipole in ((320.0, 240.0), (-1e4, -1e4)):
            rn, rm = po.r_search_for_triangulation(k1, k2, F12, epipole, sf, ls, only_stereo, True)
            dn, dm = po.r_search_for_triangulation(k1, k2, F12, epipole, sf, ls, only_stereo, True, L=D)
            assert dn == rn and np.array_equal(dm, rm), (seed, dn, rn)
    assert rn > 0


@pytest.mark.parametrize("case", ["plain", "flags", "stereo", "stereo_fwd", "stereo_bwd", "no_ori"])
def test_search_by_projection_last(libs, case):
    from test_oracle_matcher_vs_ref import _proj_inputs
    po, D = libs
    for seed in range(3):
        cur, last = _proj_inputs(200 + seed, stereo=case.startswith("stereo"), flags=case in ("flags", "stereo"))
        mode = {"stereo_fwd": 1, "stereo_bwd": 2}.get(case, 0)
        kw = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF, mbf=40.0 if case.startswith("stereo") else 0.0, search_mode=mode)
        rn, rm = po.r_search_by_projection(cur, last, 15.0, case != "no_ori", **kw)
        dn, dm = po.r_search_by_projection(cur, last, 15.0, case != "no_ori", L=D, **kw)
        assert dn == rn and np.array_equal(dm, rm), (seed, dn, rn)
    assert rn > 0


@pytest.mark.parametrize("case", ["plain", "flags", "stereo"])
def test_search_by_projection_mappoints(libs, case):
    po, D = libs
    for seed in range(3):
        F, mp = window_scene(300 + seed, stereo=case == "stereo", flags=case != "plain")
        kw = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF)
        rn, rm = po.r_search_by_projection_mappoints(F, mp, 3.0, 0.8, **kw)
        dn, dm = po.r_search_by_projection_mappoints(F, mp, 3.0, 0.8, L=D, **kw)
        assert dn == rn and np.array_equal(dm, rm), (seed, dn, rn)
    assert rn > 0


@pytest.mark.parametrize("case", ["plain", "flags", "no_ori"])
def test_search_by_projection_kf(libs, case):
    po, D = libs
    lsf = float(np.log(np.float32(1.2)))
    for seed in range(3):
        F, kf = kf_scene(400 + seed, flags=case == "flags")
        kw = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF, log_scale_factor=lsf)
        rn, rm = po.r_search_by_projection_kf(F, kf, 15.0, 100, case != "no_ori", **kw)
        dn, dm = po.r_search_by_projection_kf(F, kf, 15.0, 100, case != "no_ori", L=D, **kw)
        assert dn == rn and np.array_equal(dm, rm), (seed, dn, rn)
    assert rn > 0


@pytest.mark.parametrize("window", [10, 100])
def test_search_for_initialization(libs, window):
    po, D = libs
    for seed in range(3):
        F1, F2, prev = init_scene(500 + seed)
        rn, rm, rp = po.r_search_for_initialization(F1, F2, prev, window, 0.9, True, bounds=_BOUNDS, grid_inv=_GINV)
        dn, dm, dp = po.r_search_for_initialization(F1, F2, prev, window, 0.9, True, bounds=_BOUNDS, grid_inv=_GINV, L=D)
        assert dn == rn and np.array_equal(dm, rm) and np.array_equal(dp, rp), (seed, dn, rn)
    assert rn > 0


# ---- map-side matchers through the class interface: Fuse x2, SearchBySim3, SearchByProjection(KF,Scw) ---------------------
@pytest.mark.parametrize("flags", [False, True])
def test_search_by_projection_sim3kf(libs, flags):
    from test_oracle_matcher_vs_ref import _KW, _LSF, sim3kf_case
    po, D = libs
    for seed in range(3):
        KF, pts, _ = sim3kf_case(600 + seed, flags)
        kw = dict(log_scale_factor=_LSF, scw_scale=2.0 if seed else 1.0, **_KW)
        rn, rm = po.r_search_by_projection_sim3kf(KF, pts, 10, **kw)
        dn, dm = po.r_search_by_projection_sim3kf(KF, pts, 10, L=D, **kw)
        assert dn == rn and np.array_equal(dm, rm), (seed, dn, rn)
    assert rn > 0


@pytest.mark.parametrize("stereo", [False, True])
def test_fuse(libs, stereo):
    from test_oracle_matcher_vs_ref import _INV_SIGMA2, _KW, _LSF, fuse_case
    po, D = libs
    for seed in range(3):
        KF, pts, _ = fuse_case(700 + seed, stereo)
        kw = dict(inv_level_sigma2=_INV_SIGMA2, log_scale_factor=_LSF, mbf=20.0, **_KW)
        r = po.r_fuse(KF, pts, 3.0, **kw)
        d = po.r_fuse(KF, pts, 3.0, L=D, **kw)
        assert d[0] == r[0] and all(np.array_equal(a, b) for a, b in zip(d[1:], r[1:])), (seed, d[0], r[0])
    assert r[0] > 0


def test_fuse_sim3(libs):
    from test_oracle_matcher_vs_ref import _KW, _LSF, fuse_sim3_case
    po, D = libs
    for seed in range(3):
        KF, pts, _ = fuse_sim3_case(800 + seed)
        kw = dict(log_scale_factor=_LSF, scw_scale=4.0 if seed else 1.0, **_KW)
        r = po.r_fuse_sim3(KF, pts, 4.0, **kw)
        d = po.r_fuse_sim3(KF, pts, 4.0, L=D, **kw)
        assert d[0] == r[0] and all(np.array_equal(a, b) for a, b in zip(d[1:], r[1:])), (seed, d[0], r[0])
    assert r[0] > 0


@pytest.mark.parametrize("flags", [False, True])
def test_search_by_sim3(libs, flags):
    from matchdata import sim3_scene
    from test_oracle_matcher_vs_ref import _KW, _LSF
    po, D = libs
    for seed in range(3):
        K1, K2, pre12 = sim3_scene(900 + seed, flags=flags)
        rn, rm = po.r_search_by_sim3(K1, K2, pre12, 7.5, log_scale_factor=_LSF, **_KW)
        dn, dm = po.r_search_by_sim3(K1, K2, pre12, 7.5, log_scale_factor=_LSF, L=D, **_KW)
        assert dn == rn and np.array_equal(dm, rm), (seed, dn, rn)
    assert rn > 0
