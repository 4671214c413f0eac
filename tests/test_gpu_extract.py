"""GPU parity tests proper: the CUDA path, through the C ABI, against the oracle — bit-exact at every stage.

Stage order follows ORBextractor::operator() (src/ORBextractor.cc:1043-1105): pyramid -> per-cell FAST candidates ->
DistributeOctTree -> IC_Angle -> GaussianBlur -> rBRIEF.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sorted_rows(a):
    a = np.asarray(a)
    return a[np.lexsort((a[:, 2], a[:, 0], a[:, 1]))] if len(a) else a


@pytest.fixture(scope="module")
def ex640():
    import eaof
    e = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=8)
    yield e
    e.close()


def test_tables_match_reference(ex640):
    from oracle import pyoracle as po
    t = po.RefExtractor().tables()
    assert np.array_equal(ex640.GetScaleFactors(), t["scale"])
    assert np.array_equal(ex640.GetInverseScaleFactors(), t["inv_scale"])
    assert np.array_equal(ex640.GetScaleSigmaSquares(), t["sigma2"])
    assert np.array_equal(ex640.GetInverseScaleSigmaSquares(), t["inv_sigma2"])
    assert np.array_equal(ex640.features_per_level(), t["quotas"])


def test_stages_match_oracle(ex640, frames640):
    from oracle import pyoracle as po
    for fi in range(2):
        img = frames640[fi]
        kps, desc = ex640(img)
        okps, odesc, pl, bl, cl = po.o_extract(img, dumps=True)
        for l in range(8):
            assert np.array_equal(ex640.pyramid_level(l, with_border=True), pl[l]), f"pyramid level {l}"
            assert np.array_equal(ex640.blurred_level(l), bl[l]), f"blur level {l}"
            assert np.array_equal(_sorted_rows(ex640.candidates(l)), _sorted_rows(cl[l])), f"FAST candidates level {l}"
        assert len(kps) == len(okps)
        for field in ("x", "y", "octave", "response", "size", "angle"):
            assert np.array_equal(kps[field], okps[field]), field
        assert np.array_equal(desc, odesc)


def test_end_to_end_matches_reference_binary(ex640, frames640):
    """oracle/_ref = the reference's own ORBextractor.cc compiled unmodified."""
    from oracle import pyoracle as po
    ref = po.RefExtractor()
    res = ex640.extract_batch(frames640)
    for fi in range(len(frames640)):
        rk, rd = ref.extract(frames640[fi])
        k, d = res[fi]
        assert len(k) == len(rk)
        assert np.array_equal(k, rk)
        assert np.array_equal(d, rd)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_blur_modes(frames640, mode):
    import eaof
    from oracle import pyoracle as po
    e = eaof.ORBextractor(width=640, height=480, blur_mode=mode)
    k, d = e(frames640[0])
    ok, od = po.o_extract(frames640[0], blur_mode=mode)
    assert np.array_equal(k, ok) and np.array_equal(d, od)
    e.close()


def test_adversarial_frames():
    import eaof
    from eaof import synth
    from oracle import pyoracle as po
    e = eaof.ORBextractor(width=320, height=240, max_batch=4)
    ref = po.RefExtractor()
    for name, img in synth.adversarial_frames(320, 240).items():
        k, d = e(img)
        rk, rd = ref.extract(img)
        assert len(k) == len(rk), name
        assert np.array_equal(k, rk), name
        assert np.array_equal(d, rd), name
    e.close()


def test_shapes_the_reference_cannot_handle_are_rejected():
    """100x80 with 8 levels gives a 0-row detection window at level 5: the reference divides by zero there
    (src/ORBextractor.cc:543), so the C ABI refuses the shape instead of inventing a result."""
    import eaof
    with pytest.raises(eaof.EaofError, match="-3"):
        eaof.ORBextractor(300, 1.2, 8, 20, 7, width=100, height=80)


@pytest.mark.parametrize("shape,nf", [((848, 480), 1200), ((1920, 1080), 4000), ((160, 120), 300), ((200, 150), 50),
                                      ((752, 480), 2000)])
def test_other_config_sizes(shape, nf):
    import eaof
    from eaof import synth
    from oracle import pyoracle as po
    w, h = shape
    tex = synth.base_texture(w, h, seed=1234 + w)
    fr = synth.make_frames(2, w, h, tex=tex)
    e = eaof.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=2)
    ref = po.RefExtractor(nf, 1.2, 8, 20, 7)
    res = e.extract_batch(fr)
    for fi in range(2):
        rk, rd = ref.extract(fr[fi])
        k, d = res[fi]
        assert len(k) == len(rk)
        assert np.array_equal(k, rk) and np.array_equal(d, rd)
    e.close()


def test_other_pyramid_parameters():
    import eaof
    from eaof import synth
    from oracle import pyoracle as po
    tex = synth.base_texture(320, 240, seed=77)
    fr = synth.make_frames(1, 320, 240, tex=tex)
    for nf, sf, nl, ini, mn in ((500, 1.5, 4, 20, 7), (800, 2.0, 3, 30, 10), (300, 1.1, 6, 12, 5), (40, 1.2, 8, 20, 7)):
        e = eaof.ORBextractor(nf, sf, nl, ini, mn, width=320, height=240)
        ref = po.RefExtractor(nf, sf, nl, ini, mn)
        k, d = e(fr[0])
        rk, rd = ref.extract(fr[0])
        assert len(k) == len(rk), (nf, sf, nl)
        assert np.array_equal(k, rk) and np.array_equal(d, rd), (nf, sf, nl)
        e.close()


def test_batch_equals_single(ex640, frames640):
    res = ex640.extract_batch(frames640)
    for fi in (0, 3, 5):
        k, d = ex640(frames640[fi])
        assert np.array_equal(k, res[fi][0]) and np.array_equal(d, res[fi][1])


def test_chunked_host_pipeline_and_staged_outputs(frames640, monkeypatch):
    """eaof_orb_extract_batch overlaps upload / kernels / download of consecutive chunks (EAOF_CHUNK frames each);
    results must not depend on the chunking, nor on whether the caller's cap equals the device capacity (direct
    download) or is larger (pinned staging + host compaction)."""
    import ctypes as C
    import eaof
    monkeypatch.setenv("EAOF_CHUNK", "2")
    ex = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=6)
    monkeypatch.delenv("EAOF_CHUNK")
    one = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=6)
    a = ex.extract_batch(frames640)          # 3 chunks of 2 frames
    b = one.extract_batch(frames640)         # single chunk
    for f in range(len(frames640)):
        assert np.array_equal(a[f][0], b[f][0]) and np.array_equal(a[f][1], b[f][1])
    a2 = ex.extract_batch(frames640[:5])     # ragged last chunk, workspace reused
    for f in range(5):
        assert np.array_equal(a2[f][0], b[f][0]) and np.array_equal(a2[f][1], b[f][1])
    cap = ex.cap + 7
    n = len(frames640)
    kps = np.zeros((n, cap), eaof.KP_DTYPE)
    desc = np.zeros((n, cap, 32), np.uint8)
    cnt = np.zeros(n, np.int32)
    fr = np.ascontiguousarray(frames640)
    rc = ex.L.eaof_orb_extract_batch(ex.h, fr.ctypes.data, n, 640, 480, 640, 640 * 480, kps.ctypes.data, desc.ctypes.data,
                                     cap, cnt.ctypes.data)
    assert rc == 0, ex.L.eaof_last_error()
    for f in range(n):
        assert cnt[f] == len(b[f][0])
        assert np.array_equal(kps[f, :cnt[f]], b[f][0]) and np.array_equal(desc[f, :cnt[f]], b[f][1])
    # a cap that cannot hold the result is an argument error, not a truncation
    small = 16
    rc = ex.L.eaof_orb_extract_batch(ex.h, fr.ctypes.data, n, 640, 480, 640, 640 * 480, kps.ctypes.data, desc.ctypes.data,
                                     small, cnt.ctypes.data)
    assert rc == -1
    ex.close()
    one.close()


def test_device_sincosf_sweep():
    """The device restatement of glibc sinf/cosf against the host's libm, all floats in [0, 6.4] (strided)."""
    import eaof
    import ctypes as C
    L = eaof.lib()
    L.eaof_debug_sincosf_mismatches.restype = C.c_long
    L.eaof_debug_sincosf_mismatches.argtypes = [C.c_float, C.c_uint32]
    assert L.eaof_debug_sincosf_mismatches(6.4, 7) == 0


def _color_frames(frames, color, seed=9):
    """Interleaved colour frames whose channels are different textures (so the conversion matters)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n, h, w = frames.shape
    ch = 4 if color >= 2 else 3
    out = np.empty((n, h, w, ch), np.uint8)
    out[..., 0] = frames
    out[..., 1] = np.roll(frames, 7, axis=2)
    out[..., 2] = 255 - np.roll(frames, 5, axis=1)
    if ch == 4:
        out[..., 3] = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    return out


@pytest.mark.parametrize("color", [0, 1, 2, 3])
@pytest.mark.parametrize("mode", [0, 1])
def test_color_ingest(frames640, color, mode):
    """cvtColor + extraction on the device == the oracle's gray conversion followed by the reference extraction."""
    import eaof
    from oracle import pyoracle as po
    col = _color_frames(frames640[:3], color)
    ex = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=3)
    res = ex.extract_batch_color(col, color, mode)
    for f in range(3):
        gray = po.o_cvt_gray(col[f], color, mode)
        assert np.array_equal(ex.pyramid_level(0, frame=f), gray), f
        ok, od = po.o_extract(gray)
        k, d = res[f]
        assert len(k) == len(ok) and np.array_equal(k, ok) and np.array_equal(d, od), f
    ex.close()


def test_color_ingest_odd_width_takes_the_unaligned_path():
    import eaof
    from eaof import synth
    from oracle import pyoracle as po
    fr = synth.make_frames(2, 322, 241, tex=synth.base_texture(322, 241, seed=3))
    ex = eaof.ORBextractor(500, 1.2, 8, 20, 7, width=322, height=241, max_batch=2)
    for color in (0, 3):
        col = _color_frames(fr, color)
        res = ex.extract_batch_color(col, color, 0)
        for f in range(2):
            gray = po.o_cvt_gray(col[f], color, 0)
            assert np.array_equal(ex.pyramid_level(0, frame=f), gray)
            ok, od = po.o_extract(gray, nfeatures=500)
            assert np.array_equal(res[f][0], ok) and np.array_equal(res[f][1], od)
    ex.close()


@pytest.mark.parametrize("kind", ["f32", "u16"])
def test_stereo_from_rgbd(ex640, frames640, kind):
    from oracle import pyoracle as po
    rng = np.random.Generator(np.random.PCG64(12))
    n = 3
    res = ex640.extract_batch(frames640[:n])
    if kind == "u16":
        depth = rng.integers(0, 40000, (n, 480, 640)).astype(np.uint16)
        depth[rng.random(depth.shape) < 0.2] = 0          # holes
        scale = 1.0 / 5000.0                               # TUM DepthMapFactor
    else:
        depth = rng.uniform(-0.5, 8.0, (n, 480, 640)).astype(np.float32)
        depth[rng.random(depth.shape) < 0.1] = 0
        scale = 1.0
    ur, dd = ex640.stereo_from_rgbd(depth, 40.0, scale)
    tot = 0
    for f in range(n):
        k = res[f][0]
        our, odd = po.o_stereo_from_rgbd(k, depth[f], 40.0, np.float32(scale))
        assert np.array_equal(ur[f, :len(k)], our) and np.array_equal(dd[f, :len(k)], odd), f
        tot += int(np.sum(odd > 0))
    assert tot > 1000


@pytest.mark.parametrize("disparity,half,mb,mbf", [(14, True, 0.1, 40.0), (3, False, 0.1, 40.0), (60, True, 0.5, 20.0), (0, False, 0.1, 40.0)])
def test_stereo_matches(disparity, half, mb, mbf):
    """Two extractor handles (left / right camera), ComputeStereoMatches on the device against the oracle fed with the
    same keypoints and the handles' own pyramids (which are the oracle's, test_stages_match_oracle)."""
    import eaof
    from matchdata import stereo_pair
    from oracle import pyoracle as po
    n = 3
    pairs = [stereo_pair(seed, disparity=disparity, half_pixel=half) for seed in range(n)]
    exL = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=n)
    exR = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=n)
    resL = exL.extract_batch(np.stack([p[0] for p in pairs]))
    resR = exR.extract_batch(np.stack([p[1] for p in pairs]))
    ur, dd = exL.stereo_matches(exR, n, mb, mbf)
    sf, isf = exL.GetScaleFactors(), exL.GetInverseScaleFactors()
    total = 0
    for f in range(n):
        pL = [exL.pyramid_level(l, frame=f, with_border=True) for l in range(8)]
        pR = [exR.pyramid_level(l, frame=f, with_border=True) for l in range(8)]
        kL, dL = resL[f]
        kR, dR = resR[f]
        ou, od, _ = po.o_stereo_matches(kL, dL, kR, dR, pL, pR, sf, isf, mb, mbf)
        assert np.array_equal(ur[f, :len(kL)], ou) and np.array_equal(dd[f, :len(kL)], od), (f, int(np.sum(ur[f, :len(kL)] != ou)))
        total += int(np.sum(ou >= 0))
    assert total > 600 or disparity >= mbf / mb - 2
    # the next batches of both handles (writers of what the stereo kernels read) still produce the right results
    again = exR.extract_batch(np.stack([p[0] for p in pairs]))
    assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(again, resL))
    exL.close()
    exR.close()


def test_stereo_matches_rejects_mismatched_handles():
    import eaof
    a = eaof.ORBextractor(500, 1.2, 8, 20, 7, width=320, height=240, max_batch=1)
    b = eaof.ORBextractor(500, 1.2, 8, 20, 7, width=322, height=240, max_batch=1)
    with pytest.raises(eaof.EaofError):
        a.stereo_matches(b, 1, 0.1, 40.0)
    with pytest.raises(eaof.EaofError):
        a.stereo_matches(a, 1, 0.1, 40.0)
    a.close()
    b.close()


@pytest.mark.parametrize("dist", [[0.2624, -0.9531, -0.0054, 0.0026, 1.1633], [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05],
                                  [0.0, 0.1, 0.0, 0.0], [], [5.0, -30.0, 0.3, -0.2, 80.0]])
def test_undistort_keypoints(ex640, frames640, dist):
    from oracle import pyoracle as po
    n = 3
    res = ex640.extract_batch(frames640[:n])
    K = (np.float32(517.306408), np.float32(516.469215), np.float32(318.643040), np.float32(255.313989))
    for mode in (0, 1):
        gx, gy = ex640.undistort_keypoints(n, K, dist, mode)
        for f in range(n):
            k = res[f][0]
            if len(dist) == 0 or dist[0] == 0.0:
                ox, oy = k["x"], k["y"]  # mvKeysUn = mvKeys
            else:
                ox, oy = po.o_undistort_points(k["x"], k["y"], K, dist, guard=mode)
            assert np.array_equal(gx[f, :len(k)], ox) and np.array_equal(gy[f, :len(k)], oy), (mode, f)


@pytest.mark.parametrize("shape,nf", [((848, 480), 1200), ((1920, 1080), 4000)])
def test_frame_helpers_at_the_other_config_sizes(shape, nf):
    """BASELINE.json configs[2] / [3] geometry through the rows either side of the extractor: colour ingest, stereo matches
    between two handles, RGB-D depth lookup, vocabulary transform + SearchByBoW on the device."""
    import torch
    import eaof
    from eaof import synth
    from matchdata import stereo_pair
    from oracle import pyoracle as po
    from vocdata import make_vocabulary, tree_from
    w, h = shape
    tex = synth.base_texture(w, h, seed=77 + w)
    left, right = stereo_pair(1, width=w, height=h, disparity=21, tex=tex)
    exL = eaof.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=2)
    exR = eaof.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=2)
    # colour ingest: BGR whose conversion is the identity on equal channels
    col = np.repeat(np.stack([left, right])[..., None], 3, axis=3)
    resC = exL.extract_batch_color(col, 0, 0)
    resL = exL.extract_batch(np.stack([left, right]))
    assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(resC, resL))
    resR = exR.extract_batch(np.stack([right, left]))
    # stereo: frame 0 = (left, right), frame 1 = (right, left) (negative disparities: nothing may match)
    ur, dd = exL.stereo_matches(exR, 2, 0.1, 60.0)
    sf, isf = exL.GetScaleFactors(), exL.GetInverseScaleFactors()
    for f in range(2):
        pL = [exL.pyramid_level(l, frame=f, with_border=True) for l in range(8)]
        pR = [exR.pyramid_level(l, frame=f, with_border=True) for l in range(8)]
        ou, od, _ = po.o_stereo_matches(resL[f][0], resL[f][1], resR[f][0], resR[f][1], pL, pR, sf, isf, 0.1, 60.0)
        n = len(resL[f][0])
        assert np.array_equal(ur[f, :n], ou) and np.array_equal(dd[f, :n], od), f
    assert np.sum(ur[0] >= 0) > nf // 5
    # RGB-D
    rng = np.random.Generator(np.random.PCG64(3))
    depth = rng.uniform(0.2, 9.0, (2, h, w)).astype(np.float32)
    gu, gd = exL.stereo_from_rgbd(depth, 40.0)
    for f in range(2):
        ou, od = po.o_stereo_from_rgbd(resL[f][0], depth[f], 40.0)
        n = len(resL[f][0])
        assert np.array_equal(gu[f, :n], ou) and np.array_equal(gd[f, :n], od)
    # vocabulary transform + SearchByBoW, device-resident
    voc = make_vocabulary(10, 3, 0, 0, seed=9)
    tree = tree_from(voc)
    cap = exL.cap
    v = eaof.ORBVocabulary(tree, max_features=cap, max_sets=2)
    mt = eaof.ORBmatcher(0.75, True, max_features=cap, max_pairs=2)
    dev = torch.device("cuda:0")
    nw, nn = torch.zeros(2, dtype=torch.int32, device=dev), torch.zeros(2, dtype=torch.int32, device=dev)
    wi, ni, fi = (torch.zeros(2 * cap, dtype=torch.int32, device=dev) for _ in range(3))
    wv = torch.zeros(2 * cap, dtype=torch.float64, device=dev)
    ns = torch.zeros(2 * (cap + 1), dtype=torch.int32, device=dev)
    v.transform_orb_device(exL, 2, 2, nw.data_ptr(), wi.data_ptr(), wv.data_ptr(), nn.data_ptr(), ni.data_ptr(), ns.data_ptr(),
                           fi.data_ptr())
    v.sync()
    d_match = torch.zeros((1, cap), dtype=torch.int32, device=dev)
    d_dist = torch.zeros((1, cap), dtype=torch.int32, device=dev)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev)
    mt.bow_orb_device(exL, 2, 0, [0], [1], nn.data_ptr(), ni.data_ptr(), ns.data_ptr(), fi.data_ptr(), d_match.data_ptr(),
                      d_dist.data_ptr(), d_n.data_ptr())
    mt.sync()
    fv = [po.o_voc_transform(tree, resL[f][1], 2) for f in range(2)]
    for f in range(2):
        a = int(nw[f])
        assert a == len(fv[f][0]) and np.array_equal(wv[f * cap:f * cap + a].cpu().numpy(), fv[f][1])
    nodes = [(x[2].astype(np.int32), x[3], x[4].astype(np.int32)) for x in fv]
    on, om, od = po.o_search_by_bow(0, 0.75, True, resL[0][1], resL[0][0]["angle"], None, nodes[0], resL[1][1],
                                    resL[1][0]["angle"], None, nodes[1])
    nt = len(resL[1][0])
    assert int(d_n[0]) == on and np.array_equal(d_match[0, :nt].cpu().numpy(), om) and np.array_equal(d_dist[0, :nt].cpu().numpy(), od)
    assert on > 50
    mt.close()
    v.close()
    exL.close()
    exR.close()


@pytest.mark.parametrize("size", [(640, 480, 1000), (848, 480, 1200), (333, 251, 600), (1920, 1080, 4000)])
def test_tma_staged_fast_variants_match(size, monkeypatch):
    """The two TMA-staged forms of the FAST kernel (k_fast_tma: persistent warps, double-buffered boxes; k_fast_tma1: one
    box per one-warp CTA) against the LDG-staged default: identical keypoints and descriptors ($EAOF_FAST_TMA selects)."""
    import eaof
    from eaof import synth, workload
    w, h, nf = size
    n = 5 if w < 1000 else 2
    frames = synth.make_frames(n, w, h, tex=synth.base_texture(w, h, seed=77))
    dig = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("EAOF_FAST_TMA", mode)
        ex = eaof.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=n)
        res = ex.extract_batch(frames)
        ex.close()
        dig[mode] = [workload.frame_digest(k, d) for k, d in res]
    assert dig["1"] == dig["0"] and dig["2"] == dig["0"]


@pytest.mark.parametrize("size", [(640, 480, 1000), (848, 480, 1200), (333, 251, 600), (1920, 1080, 4000)])
@pytest.mark.parametrize("knob", ["EAOF_FAST_GENERIC=1", "EAOF_PYR_BULK=0", "EAOF_DESC_TMA=0", "EAOF_PYR_FUSED=0", "EAOF_PYR_FUSED=2",
                                  "EAOF_OCT_WIDTH=0", "EAOF_OCT_WIDTH=1", "EAOF_OCT_WIDTH=2", "EAOF_OCT_KEYS_GLOBAL=1"])
def test_kernel_variants_match_the_default(size, knob, monkeypatch):
    """Every stage has more than one kernel behind it (lane = row FAST vs the task-per-word k_fast_generic, bulk-copy pyramid vs
    per-thread loads vs the fused one-launch pyramid, TMA-staged descriptor patches vs gathers, three quadtree CTA widths with
    keys in shared or global memory); the library picks by geometry, batch size and alignment.  Whichever it picks, the output
    must be the same bytes: each knob forces one alternative, compared with the default choice (itself pinned to the oracle by
    the tests above)."""
    import eaof
    from eaof import synth, workload
    w, h, nf = size
    n = 5 if w < 1000 else 2
    frames = synth.make_frames(n, w, h, tex=synth.base_texture(w, h, seed=91))

    def run():
        ex = eaof.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=n)
        res = ex.extract_batch(frames)
        ex.close()
        return [workload.frame_digest(k, d) for k, d in res]

    base = run()
    name, val = knob.split("=")
    monkeypatch.setenv(name, val)
    assert run() == base


def test_fast_thresholds_outside_the_row_kernel(monkeypatch):
    """iniThFAST >= 128 (the carry trick of the byte-lane pre-test changes form) and minThFAST >= iniThFAST run through
    k_fast_generic by the library's own choice: against the oracle."""
    import eaof
    from eaof import synth
    from oracle import pyoracle as po
    frames = synth.make_frames(2, 320, 240, tex=synth.base_texture(320, 240, seed=5))
    for ini, mn in ((130, 40), (20, 20), (12, 30)):
        ex = eaof.ORBextractor(300, 1.2, 8, ini, mn, width=320, height=240, max_batch=2)
        res = ex.extract_batch(frames)
        ex.close()
        for f in range(2):
            ok, od = po.o_extract(frames[f], nfeatures=300, ini_th=ini, min_th=mn)
            k, d = res[f]
            assert len(k) == len(ok) and np.array_equal(k, ok) and np.array_equal(d, od), (ini, mn, f)
