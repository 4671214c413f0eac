"""ctypes view of tests/cpp/_build/libdropin_harness.so: the drop-in C++ class ORB_SLAM2::ORBextractor
(eao-fusion_b200/dropin) driven the way Frame::ExtractORB drives the reference class (src/Frame.cc:616-622)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "cpp", "_build", "libdropin_harness.so")
KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
               ("class_id", "<i4")])
_L = None


def lib():
    global _L
    if _L is None:
        L = C.CDLL(SO)
        vp, ci = C.c_void_p, C.c_int
        L.dropin_create.restype = vp
        L.dropin_create.argtypes = [ci, C.c_float, ci, ci, ci]
        L.dropin_destroy.argtypes = [vp]
        L.dropin_set_blur_mode.argtypes = [vp, ci]
        L.dropin_set_pyramid.argtypes = [vp, ci]
        L.dropin_tables.argtypes = [vp, vp, vp, vp, vp, C.POINTER(ci), C.POINTER(C.c_float)]
        L.dropin_extract.argtypes = [vp, vp, ci, ci, C.c_size_t, vp, vp, ci, ci, C.POINTER(ci)]
        L.dropin_level.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci), vp, C.c_size_t, ci]
        L.dropin_time_calls.argtypes = [vp, vp, ci, ci, ci, C.c_size_t, ci, vp]
        _L = L
    return _L


class DropinExtractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.h = self.L.dropin_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        assert self.h
        self.nlevels = nlevels
        self.cap = nfeatures + 4 * nlevels + 64

    def close(self):
        if getattr(self, "h", None):
            self.L.dropin_destroy(self.h)
            self.h = None

    __del__ = close

    def tables(self):
        n = self.nlevels
        a, b, c, d = (np.zeros(n, np.float32) for _ in range(4))
        lv, sc = C.c_int(), C.c_float()
        self.L.dropin_tables(self.h, a.ctypes.data, b.ctypes.data, c.ctypes.data, d.ctypes.data, C.byref(lv), C.byref(sc))
        return dict(scale=a, inv_scale=b, sigma2=c, inv_sigma2=d, levels=lv.value, scale_factor=sc.value)

    def extract(self, img, pre_n=0):
        """Returns (n_keypoints, descriptor_rows, keypoints, descriptors)."""
        kps = np.zeros(self.cap, KP)
        desc = np.zeros((self.cap, 32), np.uint8)
        rows = C.c_int(-1)
        if img is None or img.size == 0:
            n = self.L.dropin_extract(self.h, None, 0, 0, 0, kps.ctypes.data, desc.ctypes.data, self.cap, pre_n, C.byref(rows))
        else:
            img = np.ascontiguousarray(img, np.uint8)
            n = self.L.dropin_extract(self.h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], kps.ctypes.data,
                                      desc.ctypes.data, self.cap, pre_n, C.byref(rows))
        assert n != -2, "drop-in threw (no CUDA device?)"
        return n, rows.value, kps[:n].copy(), desc[:max(rows.value, 0)].copy()

    def time_calls(self, frames, calls, download_pyramid=False):
        """Microseconds of `calls` consecutive operator() calls measured on the C++ side; frames: (n, h, w) u8."""
        frames = np.ascontiguousarray(frames, np.uint8)
        out = np.zeros(calls, np.float64)
        self.L.dropin_set_pyramid(self.h, 1 if download_pyramid else 0)
        n = self.L.dropin_time_calls(self.h, frames.ctypes.data, frames.shape[0], frames.shape[2], frames.shape[1], frames.strides[1],
                                     calls, out.ctypes.data)
        assert n != -2, "drop-in threw (no CUDA device?)"
        return out

    def level(self, l, with_border=False):
        w, h = C.c_int(), C.c_int()
        if self.L.dropin_level(self.h, l, C.byref(w), C.byref(h), None, 0, 0) != 0:
            return None
        W, H = (w.value + 38, h.value + 38) if with_border else (w.value, h.value)
        out = np.zeros((H, W), np.uint8)
        assert self.L.dropin_level(self.h, l, C.byref(w), C.byref(h), out.ctypes.data, W, 1 if with_border else 0) == 0
        return out
