"""Pins the oracle's restated OpenCV primitives (oracle/cv_primitives.h) against the cv2 wheel in this image.

The reference calls OpenCV (not vendored, README.md:44 pins 3.3.1 in prose) for resize / copyMakeBorder / FAST /
GaussianBlur / fastAtan2 (src/ORBextractor.cc:1120,1122,809,1086,103).  cv2 4.13 shares the 8-bit algorithms for
all of them except the Gaussian taps (4.x: sum-256 taps, checked here as BLUR_CV4; 3.3.1: sum-257 taps).
"""
import numpy as np
import pytest

from oracle import pyoracle as po

cv2 = pytest.importorskip("cv2")


def _imgs(tex640):
    rng = np.random.Generator(np.random.PCG64(7))
    yield tex640[100:580, 200:840]
    yield rng.integers(0, 256, size=(480, 640), dtype=np.uint8)
    yield rng.integers(0, 256, size=(173, 211), dtype=np.uint8)
    sat = rng.integers(0, 2, size=(120, 160), dtype=np.uint8) * 255
    yield sat


def test_resize_matches_cv2(tex640):
    for img in _imgs(tex640):
        h, w = img.shape
        for s in (1.2, 1.44, 2.0736, 3.5831816):
            dw, dh = int(round(w / s)), int(round(h / s))
            ours = po.o_resize(img, dw, dh)
            ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(ours, ref)


def test_pyramid_chain_matches_cv2(tex640):
    img = tex640[64:544, 64:704]
    sizes = po.level_sizes(640, 480)
    assert sizes == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
    cur_o = cur_c = img
    for (w, h) in sizes[1:]:
        cur_o = po.o_resize(cur_o, w, h)
        cur_c = cv2.resize(cur_c, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(cur_o, cur_c)


def test_border_matches_cv2(tex640):
    for img in _imgs(tex640):
        ours = po.o_border(img, 19)
        ref = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
        assert np.array_equal(ours, ref)


def test_fast_matches_cv2_values_and_order(tex640):
    det = {t: cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True,
                                             type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16) for t in (20, 7)}
    rng = np.random.Generator(np.random.PCG64(3))
    cells = [tex640[y:y + 44, x:x + 43] for y, x in rng.integers(0, 900, size=(40, 2))]
    cells += [rng.integers(0, 256, size=(37, 38), dtype=np.uint8) for _ in range(10)]
    cells += [tex640[0:300, 0:400], np.zeros((20, 20), np.uint8), rng.integers(0, 256, size=(7, 7), dtype=np.uint8),
              rng.integers(0, 256, size=(6, 40), dtype=np.uint8)]
    total = 0
    for c in cells:
        c = np.ascontiguousarray(c)
        for t in (20, 7):
            ours = po.o_fast(c, t)
            kp = det[t].detect(c)
            ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kp], np.int32).reshape(-1, 3)
            assert np.array_equal(ours, ref)
            total += len(ref)
    assert total > 500


def test_blur_cv4_matches_cv2(tex640):
    for img in _imgs(tex640):
        ours = po.o_blur(img, po.BLUR_CV4)
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(ours, ref)


def test_blur_cv331_taps_are_rint_256g():
    g = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()
    assert [int(v) for v in np.rint(g * np.float32(256))] == [18, 34, 49, 55, 49, 34, 18]
    # explicit integer formula on a small image
    rng = np.random.Generator(np.random.PCG64(5))
    img = rng.integers(0, 256, size=(40, 50), dtype=np.uint8)
    k = np.array([18, 34, 49, 55, 49, 34, 18], np.int64)
    p = np.pad(img.astype(np.int64), 3, mode="reflect")
    rows = sum(k[i] * p[:, i:i + 50] for i in range(7))
    acc = sum(k[j] * rows[j:j + 40, :] for j in range(7))
    expect = np.clip((acc + 32768) >> 16, 0, 255).astype(np.uint8)
    assert np.array_equal(po.o_blur(img, po.BLUR_CV331), expect)
    sse = po.o_blur(img, po.BLUR_CV331_SSE2)
    assert np.abs(sse.astype(int) - expect.astype(int)).max() <= 1


def test_fast_atan2_matches_cv2():
    rng = np.random.Generator(np.random.PCG64(11))
    y = rng.integers(-2 ** 23, 2 ** 23, size=200000).astype(np.float32)
    x = rng.integers(-2 ** 23, 2 ** 23, size=200000).astype(np.float32)
    y[:100] = 0
    x[50:150] = 0
    ours = po.o_atan2(y, x)
    # the scalar cv::fastAtan2(float, float) is what src/ORBextractor.cc:103 calls (cv2.phase's SIMD path contracts FMAs)
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(ours, ref)


def test_sincosf_restatement_matches_libm():
    # strided sweep of every 97th float in [0, 6.4] (the exhaustive sweep, 1.09e9 floats, was run once: 0 mismatches)
    assert po.oracle_lib().eaoo_sincosf_sweep(6.4, 97) == 0
    x = np.linspace(0, 2 * np.pi, 100001).astype(np.float32)
    s, c = po.o_sincosf(x)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 1e-6


def test_ctor_tables_match_reference():
    for nf, sf, nl in ((1000, 1.2, 8), (1200, 1.2, 8), (2000, 1.2, 8), (4000, 1.2, 8), (500, 1.5, 5), (300, 2.0, 3)):
        a = po.RefExtractor(nf, sf, nl).tables()
        b = po.o_tables(nf, sf, nl)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    t = po.o_tables(1000, 1.2, 8)
    assert list(t["quotas"]) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(t["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


@pytest.mark.parametrize("code,color", [("COLOR_BGR2GRAY", 0), ("COLOR_RGB2GRAY", 1), ("COLOR_BGRA2GRAY", 2), ("COLOR_RGBA2GRAY", 3)])
def test_cvt_gray_cv4_formula_equals_cv2(code, color):
    """The 15-bit RGB2Gray<uchar> formula (EAOF_GRAY_CV4) is what cv2 4.x computes; the 14-bit one (EAOF_GRAY_CV331, the
    reference's pinned OpenCV 3.3.1) differs from it on a small fraction of pixels and cannot be checked here."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.Generator(np.random.PCG64(4))
    ch = 4 if color >= 2 else 3
    img = rng.integers(0, 256, (240, 321, ch), dtype=np.uint8)
    img[:16, :16] = 255
    img[16:32, :16] = 0
    want = cv2.cvtColor(img, getattr(cv2, code))
    assert np.array_equal(po.o_cvt_gray(img, color, 1), want)
    diff = po.o_cvt_gray(img, color, 0).astype(int) - want
    assert 0 < np.count_nonzero(diff) < 0.02 * diff.size and np.abs(diff).max() == 1


@pytest.mark.parametrize("dist", [[0.2624, -0.9531, -0.0054, 0.0026, 1.1633],          # TUM1.yaml
                                  [0.2312, -0.7849, -0.0033, -0.0001, 0.9172],        # TUM2.yaml
                                  [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05],  # EuRoC, 4 coefficients
                                  [5.0, -30.0, 0.3, -0.2, 80.0]])                      # absurd: drives icdist negative
def test_undistort_points_equals_cv2(dist):
    """cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) restated in double arithmetic == cv2 4.13 (guard = 1);
    the OpenCV 3.3.1 form (guard = 0) differs only where icdist < 0, which real calibrations never reach."""
    rng = np.random.Generator(np.random.PCG64(11))
    n = 5000
    x = rng.uniform(-20, 660, n).astype(np.float32)
    y = rng.uniform(-20, 500, n).astype(np.float32)
    K = np.array([[517.306408, 0, 318.643040], [0, 516.469215, 255.313989], [0, 0, 1]], np.float32)
    D = np.asarray(dist, np.float32)
    want = cv2.undistortPoints(np.stack([x, y], 1).reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    gx, gy = po.o_undistort_points(x, y, (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), D, guard=1)
    assert np.array_equal(gx, want[:, 0]) and np.array_equal(gy, want[:, 1])
    ox, oy = po.o_undistort_points(x, y, (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), D, guard=0)
    if abs(dist[0]) < 1:
        assert np.array_equal(ox, gx) and np.array_equal(oy, gy)
