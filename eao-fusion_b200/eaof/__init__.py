"""Thin ctypes loader for libeaof_orb.so (the C ABI declared in include/eaof_orb.h).

The product is the CUDA library plus the C++ drop-in classes in ../dropin; this Python layer exists only so that
tests/, bench.py and __graft_entry__.py can drive the C ABI.  It never falls back to a CPU implementation: a missing
library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("EAOF_LIB_PATH") or os.path.join(PKG_DIR, "lib", "libeaof_orb.so")  # override: build-variant experiments

BLUR_CV331, BLUR_CV4, BLUR_CV331_SSE2 = 0, 1, 2
ERR_BUSY = -6  # EAOF_ERR_BUSY, include/eaof_orb.h
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4")])


class EaofError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int), ("ini_th_fast", C.c_int),
                ("min_th_fast", C.c_int), ("blur_mode", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("max_batch", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EaofError(f"{LIB_PATH} is missing: run `make` (python -c 'import __graft_entry__ as g; g.build()')")
        L = C.CDLL(LIB_PATH)
        vp, ci, sz = C.c_void_p, C.c_int, C.c_size_t
        L.eaof_last_error.restype = C.c_char_p
        L.eaof_orb_create.argtypes = [C.POINTER(Params), ci, C.POINTER(vp)]
        L.eaof_orb_destroy.argtypes = [vp]
        L.eaof_orb_max_keypoints.argtypes = [vp]
        L.eaof_orb_scale_tables.argtypes = [vp, vp, vp, vp, vp, vp]
        L.eaof_orb_level_size.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci)]
        L.eaof_orb_extract.argtypes = [vp, vp, ci, ci, sz, vp, vp, ci, C.POINTER(ci)]
        L.eaof_orb_extract_batch.argtypes = [vp, vp, ci, ci, ci, sz, sz, vp, vp, ci, vp]
        L.eaof_orb_extract_batch_async.argtypes = [vp, vp, ci, ci, ci, sz, sz, vp, vp, ci]
        L.eaof_orb_extract_batch_wait.argtypes = [vp, vp]
        L.eaof_orb_set_pipeline_chunk.argtypes = [vp, ci]
        L.eaof_orb_extract_batch_device.argtypes = [vp, vp, ci, ci, ci, sz, sz]
        L.eaof_orb_sync.argtypes = [vp]
        L.eaof_orb_extract_batch_device_color.argtypes = [vp, vp, ci, ci, ci, sz, sz, ci, ci]
        L.eaof_orb_extract_batch_color.argtypes = [vp, vp, ci, ci, ci, sz, sz, ci, ci, vp, vp, ci, vp]
        L.eaof_orb_stereo_from_rgbd_device.argtypes = [vp, ci, vp, ci, C.c_float, sz, sz, vp, C.c_float, vp, vp]
        L.eaof_orb_stereo_from_rgbd.argtypes = [vp, ci, vp, ci, C.c_float, sz, sz, C.c_float, vp, vp, ci]
        L.eaof_orb_undistort_keypoints_device.argtypes = [vp, ci] + [C.c_float] * 4 + [vp, ci, ci, vp, vp]
        L.eaof_orb_undistort_keypoints.argtypes = [vp, ci] + [C.c_float] * 4 + [vp, ci, ci, vp, vp, ci]
        L.eaof_stereo_matches_device.argtypes = [vp, vp, ci, C.c_float, C.c_float, vp, vp]
        L.eaof_stereo_matches.argtypes = [vp, vp, ci, C.c_float, C.c_float, vp, vp, ci]
        L.eaof_orb_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(ci)]
        L.eaof_orb_fetch_results.argtypes = [vp, ci, vp, vp, ci, vp]
        L.eaof_orb_pyramid_level.argtypes = [vp, ci, ci, ci, vp, sz]
        L.eaof_orb_debug_blurred_level.argtypes = [vp, ci, ci, vp, sz]
        L.eaof_orb_debug_candidates.argtypes = [vp, ci, ci, vp, ci, C.POINTER(ci)]
        L.eaof_orb_set_profiling.argtypes = [vp, ci]
        L.eaof_orb_stage_times.argtypes = [vp, vp]
        L.eaof_orb_last_launch_count.argtypes = [vp]
        L.eaof_ring_create.argtypes = [ci, ci, ci, ci, C.POINTER(vp)]
        L.eaof_ring_destroy.argtypes = [vp]
        L.eaof_ring_acquire.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
        L.eaof_ring_commit.argtypes = [vp, C.c_double]
        L.eaof_ring_pending.argtypes = [vp]
        L.eaof_ring_peek.argtypes = [vp, ci, C.POINTER(vp), C.POINTER(C.c_double)]
        L.eaof_ring_release.argtypes = [vp, ci]
        L.eaof_orb_extract_ring.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci, vp, vp, C.POINTER(ci)]
        L.eaof_orb_stream.restype = vp
        L.eaof_orb_stream.argtypes = [vp]
        _lib = L
    return _lib


def _ck(rc: int):
    if rc != 0:
        raise EaofError(f"eaof error {rc}: {lib().eaof_last_error().decode()}")


class ORBextractor:
    """Mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:45-111) over the C ABI.

    Constructor arguments keep the reference's order and meaning; width/height/max_batch size the device workspace.
    """

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, *, width=640,
                 height=480, max_batch=1, blur_mode=BLUR_CV331, device=0):
        self.L = lib()
        self.params = Params(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, blur_mode, width, height, max_batch)
        h = C.c_void_p()
        _ck(self.L.eaof_orb_create(C.byref(self.params), device, C.byref(h)))
        self.h = h
        self.nlevels = nlevels
        self.width, self.height, self.max_batch = width, height, max_batch
        self.cap = self.L.eaof_orb_max_keypoints(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.eaof_orb_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- getters (include/ORBextractor.h:60-80)
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return float(self.params.scale_factor)

    def _tables(self):
        n = self.nlevels
        a, b, c, d = (np.zeros(n, np.float32) for _ in range(4))
        q = np.zeros(n, np.int32)
        _ck(self.L.eaof_orb_scale_tables(self.h, a.ctypes.data, b.ctypes.data, c.ctypes.data, d.ctypes.data, q.ctypes.data))
        return a, b, c, d, q

    def GetScaleFactors(self):
        return self._tables()[0]

    def GetInverseScaleFactors(self):
        return self._tables()[1]

    def GetScaleSigmaSquares(self):
        return self._tables()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3]

    def features_per_level(self):
        return self._tables()[4]

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        _ck(self.L.eaof_orb_level_size(self.h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    # ---- operator()
    def __call__(self, image: np.ndarray):
        """One frame, host buffers.  Returns (keypoints[KP_DTYPE], descriptors[n,32]) or None for an empty image."""
        if image is None or image.size == 0:
            return None  # the reference returns silently, src/ORBextractor.cc:1046-1047
        assert image.dtype == np.uint8 and image.ndim == 2  # assert(image.type() == CV_8UC1), :1050
        image = np.ascontiguousarray(image)
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int()
        _ck(self.L.eaof_orb_extract(self.h, image.ctypes.data, image.shape[1], image.shape[0], image.strides[0],
                                    kps.ctypes.data, desc.ctypes.data, self.cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, frames: np.ndarray):
        """frames (n,h,w) u8 on the host -> list of (keypoints, descriptors)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        kps = np.zeros((n, self.cap), KP_DTYPE)
        desc = np.zeros((n, self.cap, 32), np.uint8)
        cnt = np.zeros(n, np.int32)
        _ck(self.L.eaof_orb_extract_batch(self.h, frames.ctypes.data, n, w, h, w, w * h, kps.ctypes.data,
                                          desc.ctypes.data, self.cap, cnt.ctypes.data))
        return [(kps[f, :cnt[f]].copy(), desc[f, :cnt[f]].copy()) for f in range(n)]

    def extract_batch_color(self, frames: np.ndarray, color=0, gray_mode=0):
        """frames (n,h,w,3|4) u8 interleaved colour on the host; color = COLOR_BGR/RGB/BGRA/RGBA, gray_mode = GRAY_CV331 /
        GRAY_CV4: cvtColor (src/Tracking.cc:324-337) + extraction.  Returns a list of (keypoints, descriptors)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w, ch = frames.shape
        assert ch == (4 if color >= 2 else 3)
        kps = np.zeros((n, self.cap), KP_DTYPE)
        desc = np.zeros((n, self.cap, 32), np.uint8)
        cnt = np.zeros(n, np.int32)
        _ck(self.L.eaof_orb_extract_batch_color(self.h, frames.ctypes.data, n, w, h, w * ch, w * h * ch, int(color),
                                                int(gray_mode), kps.ctypes.data, desc.ctypes.data, self.cap, cnt.ctypes.data))
        return [(kps[f, :cnt[f]].copy(), desc[f, :cnt[f]].copy()) for f in range(n)]

    def stereo_from_rgbd(self, depth: np.ndarray, mbf: float, depth_scale=1.0):
        """Frame::ComputeStereoFromRGBD (src/Frame.cc:1016-1037) for the keypoints of the last batch.  depth (n,h,w)
        float32 (the CV_32F map) or uint16 (raw map, d = raw*depth_scale).  Returns (mvuRight, mvDepth) as (n, cap)."""
        depth = np.ascontiguousarray(depth)
        assert depth.dtype in (np.float32, np.uint16)
        n, h, w = depth.shape
        px = depth.dtype.itemsize
        ur = np.full((n, self.cap), -1, np.float32)
        dd = np.full((n, self.cap), -1, np.float32)
        _ck(self.L.eaof_orb_stereo_from_rgbd(self.h, n, depth.ctypes.data, 1 if depth.dtype == np.uint16 else 0,
                                             float(depth_scale), w * px, w * h * px, float(mbf), ur.ctypes.data,
                                             dd.ctypes.data, self.cap))
        return ur, dd

    def undistort_keypoints(self, n: int, K, dist, mode=0):
        """Frame::UndistortKeyPoints (src/Frame.cc:773-803) for the last batch.  K = (fx, fy, cx, cy).  Returns (x, y) as (n, cap)."""
        d = np.ascontiguousarray(dist, np.float32)
        xo, yo = np.zeros((n, self.cap), np.float32), np.zeros((n, self.cap), np.float32)
        _ck(self.L.eaof_orb_undistort_keypoints(self.h, n, float(K[0]), float(K[1]), float(K[2]), float(K[3]),
                                                d.ctypes.data if len(d) else None, len(d), int(mode), xo.ctypes.data,
                                                yo.ctypes.data, self.cap))
        return xo, yo

    def stereo_matches(self, right: "ORBextractor", n: int, mb: float, mbf: float):
        """Frame::ComputeStereoMatches (src/Frame.cc:841-1013): self = left camera's extractor, right = the right one,
        over the first n frames of their last batches.  Returns (mvuRight, mvDepth) as (n, cap)."""
        ur = np.full((n, self.cap), -1, np.float32)
        dd = np.full((n, self.cap), -1, np.float32)
        _ck(self.L.eaof_stereo_matches(self.h, right.h, n, float(mb), float(mbf), ur.ctypes.data, dd.ctypes.data, self.cap))
        return ur, dd

    def extract_batch_async(self, frames_ptr: int, n: int, kps_ptr: int, desc_ptr: int):
        """Enqueue upload + kernels + download of n packed frames at host address frames_ptr (pinned); outputs laid
        out [n][cap] at kps_ptr / desc_ptr.  Collect with extract_batch_wait()."""
        _ck(self.L.eaof_orb_extract_batch_async(self.h, frames_ptr, n, self.width, self.height, self.width,
                                                self.width * self.height, kps_ptr, desc_ptr, self.cap))
        self._pending_n = n

    def extract_batch_wait(self):
        cnt = np.zeros(self._pending_n, np.int32)
        _ck(self.L.eaof_orb_extract_batch_wait(self.h, cnt.ctypes.data))
        return cnt

    def set_pipeline_chunk(self, frames: int):
        _ck(self.L.eaof_orb_set_pipeline_chunk(self.h, frames))

    def extract_batch_device(self, d_ptr: int, n: int, stride=None, frame_pitch=None):
        stride = stride or self.width
        frame_pitch = frame_pitch or self.width * self.height
        _ck(self.L.eaof_orb_extract_batch_device(self.h, d_ptr, n, self.width, self.height, stride, frame_pitch))

    def sync(self):
        _ck(self.L.eaof_orb_sync(self.h))

    def fetch(self, n: int):
        kps = np.zeros((n, self.cap), KP_DTYPE)
        desc = np.zeros((n, self.cap, 32), np.uint8)
        cnt = np.zeros(n, np.int32)
        _ck(self.L.eaof_orb_fetch_results(self.h, n, kps.ctypes.data, desc.ctypes.data, self.cap, cnt.ctypes.data))
        return [(kps[f, :cnt[f]].copy(), desc[f, :cnt[f]].copy()) for f in range(n)]

    def fetch_counts(self, n: int):
        cnt = np.zeros(n, np.int32)
        _ck(self.L.eaof_orb_fetch_results(self.h, n, None, None, 0, cnt.ctypes.data))
        return cnt

    def device_results(self):
        k, d, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        cap = C.c_int()
        _ck(self.L.eaof_orb_device_results(self.h, C.byref(k), C.byref(d), C.byref(c), C.byref(cap)))
        return k.value, d.value, c.value, cap.value

    # ---- mvImagePyramid (include/ORBextractor.h:85) and stage dumps
    def pyramid_level(self, level, frame=0, with_border=False):
        w, h = self.level_size(level)
        W, H = (w + 38, h + 38) if with_border else (w, h)
        out = np.zeros((H, W), np.uint8)
        _ck(self.L.eaof_orb_pyramid_level(self.h, frame, level, 1 if with_border else 0, out.ctypes.data, W))
        return out

    def blurred_level(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        _ck(self.L.eaof_orb_debug_blurred_level(self.h, frame, level, out.ctypes.data, w))
        return out

    def candidates(self, level, frame=0):
        w, h = self.level_size(level)
        cap = (w // 2 + 2) * (h // 2 + 2)
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int()
        _ck(self.L.eaof_orb_debug_candidates(self.h, frame, level, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    def set_profiling(self, on=True):
        _ck(self.L.eaof_orb_set_profiling(self.h, 1 if on else 0))

    def stage_times(self):
        ms = np.zeros(6, np.float32)
        _ck(self.L.eaof_orb_stage_times(self.h, ms.ctypes.data))
        return dict(zip(("pyramid", "fast", "octree", "blur", "angle_desc", "total"), (float(v) for v in ms)))

    def stream_ptr(self):
        return self.L.eaof_orb_stream(self.h)

    def last_launch_count(self):
        return self.L.eaof_orb_last_launch_count(self.h)


# ------------------------------------------------------------------------------------------------------------
COLOR_BGR, COLOR_RGB, COLOR_BGRA, COLOR_RGBA = 0, 1, 2, 3
GRAY_CV331, GRAY_CV4 = 0, 1

# Matcher (include/eaof_match.h)
TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30  # ORBmatcher::TH_HIGH/TH_LOW/HISTO_LENGTH, src/ORBmatcher.cc:37-39
BOW_KF_FRAME, BOW_KF_KF = 0, 1
_mlib_ready = False


class RingFull(EaofError):
    """eaof_ring_acquire found no free slot (EAOF_ERR_BUSY): the consumer is behind."""


class FrameRing:
    """Pinned frame ring (include/eaof_orb.h): the camera callback (ros_test/src/message_flow.cc:250-254)
    writes frames into page-locked slots, the tracker takes every pending frame in one batched extraction.  One producer
    thread (push), one consumer thread (extract / release)."""

    def __init__(self, slots, width, height, channels=1):
        self.L = lib()
        h = C.c_void_p()
        _ck(self.L.eaof_ring_create(slots, width, height, channels, C.byref(h)))
        self.h = h
        self.slots, self.width, self.height, self.channels = slots, width, height, channels
        self.slot_bytes = (width * height * channels + 4095) // 4096 * 4096  # slot pitch inside the ring (include/eaof_orb.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.eaof_ring_destroy(self.h)
            self.h = None

    __del__ = close

    def _view(self, ptr):
        shape = (self.height, self.width) if self.channels == 1 else (self.height, self.width, self.channels)
        n = self.height * self.width * self.channels
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr)).reshape(shape)

    def push(self, image: np.ndarray, timestamp=0.0):
        """Producer: copy one frame into the next free slot and publish it.  Raises RingFull when no slot is free."""
        slot, stride = C.c_void_p(), C.c_size_t()
        rc = self.L.eaof_ring_acquire(self.h, C.byref(slot), C.byref(stride))
        if rc == ERR_BUSY:
            raise RingFull(self.L.eaof_last_error().decode())
        _ck(rc)
        self._view(slot.value)[...] = image
        _ck(self.L.eaof_ring_commit(self.h, float(timestamp)))

    def pending(self):
        return self.L.eaof_ring_pending(self.h)

    def peek(self, k=0):
        """Consumer: (view of the k-th oldest pending frame inside the ring, its time stamp)."""
        slot, ts = C.c_void_p(), C.c_double()
        _ck(self.L.eaof_ring_peek(self.h, k, C.byref(slot), C.byref(ts)))
        return self._view(slot.value), ts.value

    def release(self, n):
        _ck(self.L.eaof_ring_release(self.h, n))

    def extract(self, ex: "ORBextractor", max_frames=None, color=0, gray_mode=0):
        """Consumer: extract the oldest pending frames (eaof_orb_extract_ring).  Returns (results, timestamps) with
        results = list of (keypoints, descriptors); the frames stay pending until release()."""
        m = ex.max_batch if max_frames is None else max_frames
        kps = np.zeros((max(m, 1), ex.cap), KP_DTYPE)
        desc = np.zeros((max(m, 1), ex.cap, 32), np.uint8)
        cnt = np.zeros(max(m, 1), np.int32)
        ts = np.zeros(max(m, 1), np.float64)
        n = C.c_int()
        _ck(self.L.eaof_orb_extract_ring(ex.h, self.h, m, int(color), int(gray_mode), kps.ctypes.data, desc.ctypes.data,
                                         ex.cap, cnt.ctypes.data, ts.ctypes.data, C.byref(n)))
        return [(kps[f, :cnt[f]].copy(), desc[f, :cnt[f]].copy()) for f in range(n.value)], ts[:n.value].copy()


def _mlib():
    global _mlib_ready
    L = lib()
    if not _mlib_ready:
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.eaof_matcher_create.argtypes = [ci, ci, ci, C.POINTER(vp)]
        L.eaof_matcher_destroy.argtypes = [vp]
        L.eaof_matcher_stream.restype = vp
        L.eaof_matcher_stream.argtypes = [vp]
        L.eaof_matcher_sync.argtypes = [vp]
        L.eaof_hamming_distances.argtypes = [vp, vp, vp, ci, vp]
        L.eaof_match_bow.argtypes = [vp, ci, cf, ci, ci, vp, vp, vp, ci, vp, vp, vp, ci, vp, vp, vp, ci, vp, vp, vp, vp, vp,
                                     C.POINTER(ci)]
        L.eaof_match_projection.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, vp, cf, cf, cf, cf, cf, cf, ci, vp, vp, vp, vp,
                                            vp, vp, vp, vp, vp, ci, cf, cf, ci, ci, vp, vp, C.POINTER(ci)]
        L.eaof_match_projection_batch_device.argtypes = [vp, vp, ci, vp, vp, vp, vp, cf, vp, vp, vp]
        L.eaof_match_bruteforce_batch_device.argtypes = [vp, ci, cf, ci, ci, vp, vp, vp, vp, vp, ci, vp, vp, vp]
        L.eaof_match_triangulation.argtypes = ([vp, ci, ci] + [ci] + [vp] * 6 + [ci] + [vp] * 7 + [ci] + [vp] * 3 + [ci] +
                                               [vp] * 3 + [vp, cf, cf, vp, vp, ci, vp, vp, C.POINTER(ci)])
        L.eaof_match_windows.argtypes = ([vp, ci, ci] + [vp] * 7 + [cf] * 6 + [ci] + [vp] * 10 + [ci, cf, ci, ci, vp, vp,
                                                                                               C.POINTER(ci)])
        L.eaof_match_initialization.argtypes = [vp, cf, ci, ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, cf, cf, cf, cf, cf, cf, ci,
                                                vp, C.POINTER(ci)]
        L.eaof_match_windows_independent.argtypes = ([vp, ci, ci] + [vp] * 5 + [cf] * 4 + [vp, ci, ci] + [vp] * 8 +
                                                     [ci, vp, vp, C.POINTER(ci)])
        L.eaof_distinctive_descriptors.argtypes = [vp, ci, vp, vp, vp, vp]
        L.eaof_match_bow_orb_device.argtypes = [vp, vp, ci, ci, cf, ci, ci] + [vp] * 9
        L.eaof_matcher_last_distance_count.restype = C.c_longlong
        L.eaof_matcher_last_distance_count.argtypes = [vp]
        _mlib_ready = True
    return L


def _p(a):
    return None if a is None else a.ctypes.data


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


def csr_from_nodes(node_of_feature):
    """DBoW2::FeatureVector for tests: feature i belongs to node node_of_feature[i] (-1 = in no node).
    Returns (node_ids ascending, starts, idx) with features of a node in ascending index order, as
    FeatureVector::addFeature produces them."""
    node_of_feature = np.asarray(node_of_feature)
    ids = np.unique(node_of_feature[node_of_feature >= 0]).astype(np.int32)
    starts = [0]
    idx = []
    for nid in ids:
        f = np.nonzero(node_of_feature == nid)[0]
        idx.extend(f.tolist())
        starts.append(len(idx))
    return ids, np.asarray(starts, np.int32), np.asarray(idx, np.int32)


class ORBmatcher:
    """Mirror of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:37-102) over the C ABI, on plain arrays."""

    TH_HIGH, TH_LOW, HISTO_LENGTH = TH_HIGH, TH_LOW, HISTO_LENGTH

    def __init__(self, nnratio=0.6, checkOri=True, *, max_features=4096, max_pairs=1, device=0):
        self.L = _mlib()
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        h = C.c_void_p()
        _ck(self.L.eaof_matcher_create(device, max_pairs, max_features, C.byref(h)))
        self.h = h
        self.max_features, self.max_pairs = max_features, max_pairs

    def close(self):
        if getattr(self, "h", None):
            self.L.eaof_matcher_destroy(self.h)
            self.h = None

    __del__ = close

    def DescriptorDistance(self, a, b):
        a = _arr(a, np.uint8).reshape(-1, 32)
        b = _arr(b, np.uint8).reshape(-1, 32)
        out = np.zeros(len(a), np.int32)
        _ck(self.L.eaof_hamming_distances(self.h, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data))
        return out

    def SearchByBoW(self, mode, desc_q, angle_q, valid_q, nodes_q, desc_t, angle_t, valid_t, nodes_t):
        """nodes_* = (ids, starts, idx) CSR.  Returns (nmatches, match, dist)."""
        dq, dt = _arr(desc_q, np.uint8), _arr(desc_t, np.uint8)
        aq, at = _arr(angle_q, np.float32), _arr(angle_t, np.float32)
        vq, vt = _arr(valid_q, np.uint8), _arr(valid_t, np.uint8)
        iq, sq, xq = (_arr(v, np.int32) for v in nodes_q)
        it, st, xt = (_arr(v, np.int32) for v in nodes_t)
        nq, nt = len(dq), len(dt)
        nout = nt if mode == BOW_KF_FRAME else nq
        match = np.full(max(nout, 1), -1, np.int32)
        dist = np.full(max(nout, 1), -1, np.int32)
        n = C.c_int()
        _ck(self.L.eaof_match_bow(self.h, mode, self.mfNNratio, int(self.mbCheckOrientation), nq, _p(dq), _p(aq), _p(vq),
                                  nt, _p(dt), _p(at), _p(vt), len(iq), _p(iq), _p(sq), _p(xq), len(it), _p(it), _p(st),
                                  _p(xt), match.ctypes.data, dist.ctypes.data, C.byref(n)))
        return n.value, match[:nout], dist[:nout]

    def SearchByProjection(self, cur, last, th, *, bounds, grid_inv, scale_factors, mbf=0.0, search_mode=0):
        """cur: dict(x,y,octave,angle,desc[,uright,taken]); last: dict(u,v,octave,angle,desc[,valid,invz,obs]).
        bounds=(minX,maxX,minY,maxY).  Returns (nmatches, match_cur, dist_cur)."""
        cx, cy = _arr(cur["x"], np.float32), _arr(cur["y"], np.float32)
        co, ca, cd = _arr(cur["octave"], np.int32), _arr(cur["angle"], np.float32), _arr(cur["desc"], np.uint8)
        cur_r, ctk = _arr(cur.get("uright"), np.float32), _arr(cur.get("taken"), np.uint8)
        lu, lv = _arr(last["u"], np.float32), _arr(last["v"], np.float32)
        lo, la, ld = _arr(last["octave"], np.int32), _arr(last["angle"], np.float32), _arr(last["desc"], np.uint8)
        lval, linv, lobs = _arr(last.get("valid"), np.uint8), _arr(last.get("invz"), np.float32), _arr(last.get("obs"), np.uint8)
        sf = _arr(scale_factors, np.float32)
        nc, nl = len(cx), len(lu)
        match = np.full(max(nc, 1), -1, np.int32)
        dist = np.full(max(nc, 1), -1, np.int32)
        n = C.c_int()
        _ck(self.L.eaof_match_projection(self.h, nc, _p(cx), _p(cy), _p(co), _p(ca), _p(cd), _p(cur_r), _p(ctk),
                                         bounds[0], bounds[1], bounds[2], bounds[3], grid_inv[0], grid_inv[1], nl, _p(lval),
                                         _p(lu), _p(lv), _p(linv), _p(lo), _p(la), _p(ld), _p(lobs), _p(sf), len(sf),
                                         float(th), float(mbf), search_mode, int(self.mbCheckOrientation),
                                         match.ctypes.data, dist.ctypes.data, C.byref(n)))
        return n.value, match[:nc], dist[:nc]

    def SearchForInitialization(self, F1, F2, prev_matched, window_size=10, *, bounds, grid_inv):
        """src/ORBmatcher.cc:405-520.  F1: dict(octave,angle,desc); F2: dict(x,y,octave,angle,desc); prev_matched (n1,2).
        Returns (nmatches, vnMatches12, updated vbPrevMatched)."""
        o1, a1, d1 = _arr(F1["octave"], np.int32), _arr(F1["angle"], np.float32), _arr(F1["desc"], np.uint8)
        pm = np.ascontiguousarray(prev_matched, np.float32).copy()
        x2, y2, o2 = _arr(F2["x"], np.float32), _arr(F2["y"], np.float32), _arr(F2["octave"], np.int32)
        a2, d2 = _arr(F2["angle"], np.float32), _arr(F2["desc"], np.uint8)
        m12 = np.full(max(len(o1), 1), -1, np.int32)
        n = C.c_int()
        _ck(self.L.eaof_match_initialization(self.h, self.mfNNratio, int(self.mbCheckOrientation), len(o1), _p(o1), _p(a1),
                                             _p(d1), pm.ctypes.data, len(x2), _p(x2), _p(y2), _p(o2), _p(a2), _p(d2),
                                             bounds[0], bounds[1], bounds[2], bounds[3], grid_inv[0], grid_inv[1],
                                             int(window_size), m12.ctypes.data, C.byref(n)))
        return n.value, m12[:len(o1)], pm

    def SearchWindows(self, rule, F, q, th_accept, hist_mode, check_bounds, *, bounds, grid_inv):
        """eaof_match_windows.  F: dict(x,y,octave,desc[,angle,uright,taken]); q: dict(u,v,radius,min_level,max_level,desc
        [,valid,ur,angle,obs]).  Returns (nmatches, match_t, dist_t)."""
        tx, ty, to = _arr(F["x"], np.float32), _arr(F["y"], np.float32), _arr(F["octave"], np.int32)
        ta, td = _arr(F.get("angle"), np.float32), _arr(F["desc"], np.uint8)
        tr, tt = _arr(F.get("uright"), np.float32), _arr(F.get("taken"), np.uint8)
        qu, qv, qr = _arr(q["u"], np.float32), _arr(q["v"], np.float32), _arr(q["radius"], np.float32)
        ql0, ql1 = _arr(q["min_level"], np.int32), _arr(q["max_level"], np.int32)
        qval, qur, qa = _arr(q.get("valid"), np.uint8), _arr(q.get("ur"), np.float32), _arr(q.get("angle"), np.float32)
        qd, qo = _arr(q["desc"], np.uint8), _arr(q.get("obs"), np.uint8)
        nt, nq = len(tx), len(qu)
        match = np.full(max(nt, 1), -1, np.int32)
        dist = np.full(max(nt, 1), -1, np.int32)
        n = C.c_int()
        _ck(self.L.eaof_match_windows(self.h, rule, nt, _p(tx), _p(ty), _p(to), _p(ta), _p(td), _p(tr), _p(tt), bounds[0],
                                      bounds[1], bounds[2], bounds[3], grid_inv[0], grid_inv[1], nq, _p(qval), _p(qu), _p(qv),
                                      _p(qr), _p(ql0), _p(ql1), _p(qur), _p(qa), _p(qd), _p(qo), int(th_accept),
                                      self.mfNNratio, hist_mode, int(check_bounds), match.ctypes.data, dist.ctypes.data,
                                      C.byref(n)))
        return n.value, match[:nt], dist[:nt]

    def SearchByProjectionMapPoints(self, F, mp, th, *, bounds, grid_inv, scale_factors):
        """SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:45-129.  mp: dict(x,y,level,desc
        [,in_view,bad,xr,cos,obs]).  The window radius (RadiusByViewingCos * th * scale[level]) is the caller's."""
        sf = _arr(scale_factors, np.float32)
        lvl = _arr(mp["level"], np.int32)
        cos = _arr(mp.get("cos"), np.float32)
        r = np.where((cos if cos is not None else np.ones(len(lvl), np.float32)).astype(np.float64) > 0.998,
                     np.float32(2.5), np.float32(4.0)).astype(np.float32)      # RadiusByViewingCos :131-137
        if th != 1.0:
            r = (r * np.float32(th)).astype(np.float32)
        valid = np.ones(len(lvl), np.uint8)
        if mp.get("in_view") is not None:
            valid &= _arr(mp["in_view"], np.uint8)
        if mp.get("bad") is not None:
            valid &= (1 - _arr(mp["bad"], np.uint8))
        q = dict(u=mp["x"], v=mp["y"], radius=(r * sf[lvl]).astype(np.float32), min_level=lvl - 1, max_level=lvl,
                 desc=mp["desc"], valid=valid, ur=mp.get("xr"), obs=mp.get("obs"))
        if F.get("uright") is not None and q["ur"] is None:
            q["ur"] = np.zeros(len(lvl), np.float32)
        return self.SearchWindows(1, F, q, TH_HIGH, 0, False, bounds=bounds, grid_inv=grid_inv)

    def SearchByProjectionKF(self, cur, kq, th, orb_dist, *, bounds, grid_inv, scale_factors):
        """SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:1474-1601, after the
        caller's projection.  kq: dict(valid,u,v,level,angle,desc)."""
        sf = _arr(scale_factors, np.float32)
        lvl = _arr(kq["level"], np.int32)
        q = dict(u=kq["u"], v=kq["v"], radius=(np.float32(th) * sf[lvl]).astype(np.float32), min_level=lvl - 1,
                 max_level=lvl + 1, desc=kq["desc"], valid=kq["valid"], angle=kq["angle"])
        F = dict(cur)
        F.pop("uright", None)
        return self.SearchWindows(0, F, q, orb_dist, 1 if self.mbCheckOrientation else 0, True, bounds=bounds,
                                  grid_inv=grid_inv)

    def SearchWindowsIndependent(self, gate, KF, q, th_accept, *, bounds, grid_inv, inv_level_sigma2=None):
        """eaof_match_windows_independent: the search step of Fuse x2 / SearchBySim3 (src/ORBmatcher.cc:825-1326).
        KF: dict(x,y,octave,desc[,uright]); q: dict(u,v,radius,min_level,max_level,desc[,valid,ur]).
        Returns (naccepted, match_q, dist_q)."""
        tx, ty, to = _arr(KF["x"], np.float32), _arr(KF["y"], np.float32), _arr(KF["octave"], np.int32)
        td, tr = _arr(KF["desc"], np.uint8), _arr(KF.get("uright"), np.float32)
        qu, qv, qr = _arr(q["u"], np.float32), _arr(q["v"], np.float32), _arr(q["radius"], np.float32)
        ql0, ql1 = _arr(q["min_level"], np.int32), _arr(q["max_level"], np.int32)
        qval, qur, qd = _arr(q.get("valid"), np.uint8), _arr(q.get("ur"), np.float32), _arr(q["desc"], np.uint8)
        inv = _arr(inv_level_sigma2, np.float32)
        nt, nq = len(tx), len(qu)
        match = np.full(max(nq, 1), -1, np.int32)
        dist = np.full(max(nq, 1), -1, np.int32)
        n = C.c_int()
        _ck(self.L.eaof_match_windows_independent(self.h, int(gate), nt, _p(tx), _p(ty), _p(to), _p(td), _p(tr), bounds[0],
                                                  bounds[2], grid_inv[0], grid_inv[1], _p(inv), 0 if inv is None else len(inv),
                                                  nq, _p(qval), _p(qu), _p(qv), _p(qr), _p(ql0), _p(ql1), _p(qur), _p(qd),
                                                  int(th_accept), match.ctypes.data, dist.ctypes.data, C.byref(n)))
        return n.value, match[:nq], dist[:nq]

    def DistinctiveDescriptors(self, starts, desc):
        """Batched MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-307): point p owns rows
        starts[p]:starts[p+1] of desc.  Returns (best row per point relative to its start, its median distance)."""
        st, d = _arr(starts, np.int32), _arr(desc, np.uint8)
        npts = len(st) - 1
        best = np.full(max(npts, 1), -1, np.int32)
        med = np.full(max(npts, 1), -1, np.int32)
        _ck(self.L.eaof_distinctive_descriptors(self.h, npts, _p(st), _p(d), best.ctypes.data, med.ctypes.data))
        return best[:npts], med[:npts]

    def SearchForTriangulation(self, k1, k2, F12, epipole, scale_factors, level_sigma2, only_stereo=False):
        """k1: dict(desc,x,y,angle,free[,stereo],nodes); k2: the same plus octave.  free = feature has no map point.
        Returns (nmatches, match12, dist12); vMatchedPairs = [(i, match12[i]) for match12[i] >= 0]."""
        d1, d2 = _arr(k1["desc"], np.uint8), _arr(k2["desc"], np.uint8)
        x1, y1, a1 = (_arr(k1[k], np.float32) for k in ("x", "y", "angle"))
        x2, y2, a2 = (_arr(k2[k], np.float32) for k in ("x", "y", "angle"))
        o2 = _arr(k2["octave"], np.int32)
        f1, f2 = _arr(k1["free"], np.uint8), _arr(k2["free"], np.uint8)
        s1, s2 = _arr(k1.get("stereo"), np.uint8), _arr(k2.get("stereo"), np.uint8)
        i1, st1, ix1 = (_arr(v, np.int32) for v in k1["nodes"])
        i2, st2, ix2 = (_arr(v, np.int32) for v in k2["nodes"])
        F = _arr(F12, np.float32).reshape(9)
        sf, ls = _arr(scale_factors, np.float32), _arr(level_sigma2, np.float32)
        n1, n2 = len(d1), len(d2)
        match = np.full(max(n1, 1), -1, np.int32)
        dist = np.full(max(n1, 1), -1, np.int32)
        n = C.c_int()
        _ck(self.L.eaof_match_triangulation(self.h, int(self.mbCheckOrientation), int(only_stereo), n1, _p(d1), _p(x1),
                                            _p(y1), _p(a1), _p(f1), _p(s1), n2, _p(d2), _p(x2), _p(y2), _p(o2), _p(a2),
                                            _p(f2), _p(s2), len(i1), _p(i1), _p(st1), _p(ix1), len(i2), _p(i2), _p(st2),
                                            _p(ix2), _p(F), float(epipole[0]), float(epipole[1]), _p(sf), _p(ls), len(sf),
                                            match.ctypes.data, dist.ctypes.data, C.byref(n)))
        return n.value, match[:n1], dist[:n1]

    def stream_ptr(self):
        return self.L.eaof_matcher_stream(self.h)

    def sync(self):
        _ck(self.L.eaof_matcher_sync(self.h))

    def projection_batch_device(self, ex: "ORBextractor", last_frames, cur_frames, shift_x, shift_y, th, d_match, d_dist,
                                d_n):
        lf, cf_ = _arr(last_frames, np.int32), _arr(cur_frames, np.int32)
        sx, sy = _arr(shift_x, np.float32), _arr(shift_y, np.float32)
        _ck(self.L.eaof_match_projection_batch_device(self.h, ex.h, len(lf), _p(lf), _p(cf_), _p(sx), _p(sy), float(th),
                                                      d_match, d_dist, d_n))

    def bow_orb_device(self, ex: "ORBextractor", n_frames, mode, pair_q, pair_t, d_nn, d_ni, d_ns, d_fi, d_match, d_dist, d_n):
        """eaof_match_bow_orb_device: SearchByBoW between frames of an extractor batch with device-resident FeatureVectors."""
        pq, pt = _arr(pair_q, np.int32), _arr(pair_t, np.int32)
        _ck(self.L.eaof_match_bow_orb_device(self.h, ex.h, int(n_frames), int(mode), self.mfNNratio, int(self.mbCheckOrientation),
                                             len(pq), _p(pq), _p(pt), d_nn, d_ni, d_ns, d_fi, d_match, d_dist, d_n))

    def bruteforce_batch_device(self, mode, pair_q, pair_t, d_desc, d_angle, d_counts, block_stride, d_match, d_dist, d_n):
        pq, pt = _arr(pair_q, np.int32), _arr(pair_t, np.int32)
        _ck(self.L.eaof_match_bruteforce_batch_device(self.h, mode, self.mfNNratio, int(self.mbCheckOrientation), len(pq),
                                                      _p(pq), _p(pt), d_desc, d_angle, d_counts, block_stride, d_match,
                                                      d_dist, d_n))


# ---------------------------------------------------------------------------------------------------------------
# Vocabulary (include/eaof_voc.h)
_vlib_ready = False


def _vlib():
    global _vlib_ready
    L = lib()
    if not _vlib_ready:
        vp, ci = C.c_void_p, C.c_int
        L.eaof_voc_create.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp, ci, ci, ci, ci, C.POINTER(vp)]
        L.eaof_voc_destroy.argtypes = [vp]
        L.eaof_voc_transform.argtypes = [vp, ci, vp, vp, ci] + [vp] * 7
        L.eaof_voc_transform_orb_device.argtypes = [vp, vp, ci, ci] + [vp] * 7
        L.eaof_voc_sync.argtypes = [vp]
        L.eaof_voc_stream.restype = vp
        L.eaof_voc_stream.argtypes = [vp]
        _vlib_ready = True
    return L


class ORBVocabulary:
    """ORBVocabulary::transform(features, BowVector, FeatureVector, levelsup) over the C ABI.  tree: dict(L, child_start,
    child_idx, desc[n,32], weight[n] f64, word_id[n], weighting, scoring) — the arrays a loaded DBoW2 vocabulary holds."""

    def __init__(self, tree, *, max_features=4096, max_sets=1, device=0):
        self.L = _vlib()
        self.h = C.c_void_p()
        cs, ci_, d = _arr(tree["child_start"], np.int32), _arr(tree["child_idx"], np.int32), _arr(tree["desc"], np.uint8)
        w, wid = _arr(tree["weight"], np.float64), _arr(tree["word_id"], np.int32)
        _ck(self.L.eaof_voc_create(device, int(tree["L"]), len(w), _p(cs), _p(ci_), _p(d), _p(w), _p(wid), int(tree["weighting"]),
                                   int(tree["scoring"]), int(max_features), int(max_sets), C.byref(self.h)))
        self.max_features, self.max_sets = max_features, max_sets

    def close(self):
        if self.h:
            self.L.eaof_voc_destroy(self.h)
            self.h = C.c_void_p()

    def transform_sets(self, sets, levelsup=4):
        """sets: list of (n_i, 32) u8 arrays.  Returns one (word_ids, word_vals, node_ids, node_starts, feat_idx) per set."""
        cnt = [len(s) for s in sets]
        start = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
        tot = int(start[-1])
        desc = np.concatenate([np.asarray(s, np.uint8).reshape(-1, 32) for s in sets]) if tot else np.zeros((0, 32), np.uint8)
        desc = np.ascontiguousarray(desc)
        ns = len(sets)
        nw, nn = np.zeros(ns, np.int32), np.zeros(ns, np.int32)
        wi, wv = np.zeros(tot + 1, np.uint32), np.zeros(tot + 1, np.float64)
        ni, nst, fi = np.zeros(tot + 1, np.uint32), np.zeros(tot + ns + 1, np.int32), np.zeros(tot + 1, np.uint32)
        _ck(self.L.eaof_voc_transform(self.h, ns, start.ctypes.data, desc.ctypes.data, int(levelsup), nw.ctypes.data,
                                      wi.ctypes.data, wv.ctypes.data, nn.ctypes.data, ni.ctypes.data, nst.ctypes.data,
                                      fi.ctypes.data))
        out = []
        for s in range(ns):
            o = int(start[s])
            st = nst[o + s:o + s + nn[s] + 1].copy()
            out.append((wi[o:o + nw[s]].copy(), wv[o:o + nw[s]].copy(), ni[o:o + nn[s]].copy(), st, fi[o:o + st[-1]].copy()))
        return out

    def transform(self, feats, levelsup=4):
        return self.transform_sets([feats], levelsup)[0]

    def transform_orb_device(self, ex: "ORBextractor", n_frames, levelsup, d_nw, d_wi, d_wv, d_nn, d_ni, d_ns, d_fi):
        _ck(self.L.eaof_voc_transform_orb_device(self.h, ex.h, n_frames, int(levelsup), d_nw, d_wi, d_wv, d_nn, d_ni, d_ns, d_fi))

    def sync(self):
        _ck(self.L.eaof_voc_sync(self.h))

    def stream_ptr(self):
        return self.L.eaof_voc_stream(self.h)
