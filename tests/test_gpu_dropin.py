"""The drop-in C++ class ORB_SLAM2::ORBextractor (eao-fusion_b200/dropin) against the reference class
(oracle/_ref = /root/reference/src/ORBextractor.cc compiled unmodified), both called through
operator()(image, mask, keypoints, descriptors) as Frame::ExtractORB does (src/Frame.cc:616-622)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("x", "y", "size", "angle", "response", "octave")


def test_getters_before_first_frame():
    from dropin import DropinExtractor
    from oracle import pyoracle as po
    d = DropinExtractor()
    t, r = d.tables(), po.RefExtractor().tables()
    for k in ("scale", "inv_scale", "sigma2", "inv_sigma2"):
        assert np.array_equal(t[k], r[k]), k
    assert t["levels"] == 8 and t["scale_factor"] == np.float32(1.2)
    d.close()


def test_operator_call_matches_reference_class(frames640):
    from dropin import DropinExtractor
    from oracle import pyoracle as po
    d, ref = DropinExtractor(), po.RefExtractor()
    for fi in range(3):
        n, rows, k, desc = d.extract(frames640[fi])
        rk, rd = ref.extract(frames640[fi], keep_pyramid=True)
        assert n == rows == len(rk)
        for f in FIELDS:
            assert np.array_equal(k[f], rk[f]), f
        assert (k["class_id"] == -1).all()
        assert np.array_equal(desc, rd)
        for l in range(8):  # mvImagePyramid: same ROI content and same 19-px border around it
            assert np.array_equal(d.level(l), ref.level(l)), l
            assert np.array_equal(d.level(l, with_border=True), ref.level(l, with_border=True)), l
    d.close()


def test_frame_size_may_change_between_calls():
    """The reference re-derives all geometry per frame; the drop-in rebuilds its workspace when the size changes."""
    from dropin import DropinExtractor
    from eaof import synth
    from oracle import pyoracle as po
    d, ref = DropinExtractor(500), po.RefExtractor(500)
    for (w, h) in ((320, 240), (400, 300), (320, 240)):
        img = synth.make_frames(1, w, h, tex=synth.base_texture(w, h, seed=w))[0]
        n, rows, k, desc = d.extract(img)
        rk, rd = ref.extract(img)
        assert n == len(rk) and np.array_equal(desc, rd)
        for f in FIELDS:
            assert np.array_equal(k[f], rk[f]), f
    d.close()


def test_row_stride_larger_than_width(frames640):
    """cv::Mat ROI input (step > cols), as produced by image(cv::Rect(...))."""
    from dropin import DropinExtractor
    from oracle import pyoracle as po
    big = np.zeros((300, 512), np.uint8)
    big[:240, :320] = frames640[0][:240, :320]
    roi = big[:240, :320]
    assert roi.strides[0] == 512
    d, ref = DropinExtractor(500), po.RefExtractor(500)
    lib = d.L
    import ctypes as C
    from dropin import KP
    kps = np.zeros(d.cap, KP)
    desc = np.zeros((d.cap, 32), np.uint8)
    rows = C.c_int()
    n = lib.dropin_extract(d.h, roi.ctypes.data, 320, 240, 512, kps.ctypes.data, desc.ctypes.data, d.cap, 0, C.byref(rows))
    rk, rd = ref.extract(np.ascontiguousarray(roi))
    assert n == len(rk) and np.array_equal(desc[:n], rd)
    assert np.array_equal(kps["x"][:n], rk["x"]) and np.array_equal(kps["angle"][:n], rk["angle"])
    d.close()


def test_empty_image_leaves_outputs_untouched():
    """src/ORBextractor.cc:1046-1047: silent return before anything is cleared."""
    from dropin import DropinExtractor
    d = DropinExtractor()
    n, rows, _, _ = d.extract(None, pre_n=5)
    assert n == 5 and rows == 5
    d.close()


def test_zero_keypoints_releases_descriptors():
    """src/ORBextractor.cc:1064-1065,1072: a flat image gives no keypoints -> descriptors.release(), keypoints.clear()."""
    from dropin import DropinExtractor
    d = DropinExtractor()
    n, rows, _, _ = d.extract(np.full((240, 320), 128, np.uint8), pre_n=5)
    assert n == 0 and rows == 0
    d.close()
