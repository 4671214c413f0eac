"""Frame::ComputeStereoMatches / ComputeStereoFromRGBD: the restatements against the reference's own function text
(cut out of src/Frame.cc at build time and compiled unmodified into oracle/_ref/libstereo_ref.so): bit-identical
mvuRight / mvDepth."""
import numpy as np
import pytest

from matchdata import stereo_pair
from oracle import pyoracle as po


def extract_with_pyramid(img, nfeatures=1000):
    ex = po.RefExtractor(nfeatures)
    k, d = ex.extract(img, keep_pyramid=True)
    pyr = [ex.level(l, with_border=True) for l in range(8)]
    t = ex.tables()
    return k, d, pyr, t["scale"], t["inv_scale"]


@pytest.mark.parametrize("disparity,half", [(14, True), (3, False), (60, True), (0, False)])
@pytest.mark.parametrize("mb,mbf", [(0.1, 40.0), (0.5, 20.0)])
def test_stereo_matches_restatement_equals_reference(disparity, half, mb, mbf):
    total = 0
    for seed in range(2):
        left, right = stereo_pair(seed, disparity=disparity, half_pixel=half)
        kL, dL, pL, sf, isf = extract_with_pyramid(left)
        kR, dR, pR, _, _ = extract_with_pyramid(right)
        ou, od, sad = po.o_stereo_matches(kL, dL, kR, dR, pL, pR, sf, isf, mb, mbf)
        ru, rd = po.r_stereo_matches(kL, dL, kR, dR, pL, pR, sf, isf, mb, mbf)
        assert np.array_equal(ou, ru) and np.array_equal(od, rd), (seed, int(np.sum(ou != ru)))
        total += int(np.sum(ou >= 0))
        if disparity < mbf / mb - 2:
            ok = ou >= 0
            assert ok.sum() > 200 and np.median(kL["x"][ok] - ou[ok]) == pytest.approx(disparity + (0.5 if half else 0), abs=0.6)
            assert np.sum((sad >= 0) & ~ok) > 0  # the median filter removed something
    assert total > 0 or disparity >= mbf / mb


def test_stereo_matches_degenerate_inputs():
    left, right = stereo_pair(5)
    kL, dL, pL, sf, isf = extract_with_pyramid(left)
    flat = np.full_like(right, 90)
    kR, dR, pR, _, _ = extract_with_pyramid(right)
    # flat right pyramid: every SAD row is constant -> 0/0 in the parabola -> NaN disparity is never accepted
    pFlat = [np.full_like(p, 90) for p in pR]
    ou, od, _ = po.o_stereo_matches(kL, dL, kR, dR, pL, pFlat, sf, isf, 0.1, 40.0)
    ru, rd = po.r_stereo_matches(kL, dL, kR, dR, pL, pFlat, sf, isf, 0.1, 40.0)
    assert np.array_equal(ou, ru) and np.array_equal(od, rd)
    # no right keypoints at all
    e = kR[:0]
    ou, od, _ = po.o_stereo_matches(kL, dL, e, dR[:0], pL, pR, sf, isf, 0.1, 40.0)
    assert np.all(ou == -1) and np.all(od == -1)
    del flat


def test_stereo_from_rgbd_restatement_equals_reference():
    rng = np.random.Generator(np.random.PCG64(8))
    left, _ = stereo_pair(2)
    k, _, _, _, _ = extract_with_pyramid(left)
    depth = rng.uniform(-0.5, 8.0, (480, 640)).astype(np.float32)
    depth[rng.random(depth.shape) < 0.1] = 0
    ou, od = po.o_stereo_from_rgbd(k, depth, 40.0)
    ru, rd = po.r_stereo_from_rgbd(k, depth, 40.0)
    assert np.array_equal(ou, ru) and np.array_equal(od, rd) and np.sum(od > 0) > 500
    raw = rng.integers(0, 40000, (480, 640)).astype(np.uint16)
    conv = (raw.astype(np.float32) * np.float32(1.0 / 5000.0)).astype(np.float32)
    ou, od = po.o_stereo_from_rgbd(k, raw, 40.0, np.float32(1.0 / 5000.0))
    ru, rd = po.r_stereo_from_rgbd(k, conv, 40.0)
    assert np.array_equal(ou, ru) and np.array_equal(od, rd)
