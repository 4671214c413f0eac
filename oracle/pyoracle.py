"""ORACLE — test infrastructure only.

ctypes bindings for the two CPU checkers:
  * ``RefExtractor``  -> oracle/_ref/liborb_ref.so  (the UNMODIFIED reference ORBextractor.cc + cv shim)
  * ``oracle_lib()``  -> oracle/liborb_oracle.so    (our independent restatement, stage by stage)
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arms may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "liborb_ref.so")
ORACLE_SO = os.environ.get("EAOF_ORACLE_SO") or os.path.join(HERE, "liborb_oracle.so")  # override: the ASan/UBSan build (make -C oracle asan)

BLUR_CV331, BLUR_CV4, BLUR_CV331_SSE2 = 0, 1, 2

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4")])

_u8p = C.POINTER(C.c_uint8)


def _ptr(a: np.ndarray, t=_u8p):
    return a.ctypes.data_as(t)


def build(ref: bool = True) -> None:
    subprocess.check_call(["make", "-s", "-C", HERE])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            build(ref=True)
        L = C.CDLL(REF_SO)
        L.orbref_create.restype = C.c_void_p
        L.orbref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbref_destroy.argtypes = [C.c_void_p]
        L.orbref_set_canonical.argtypes = [C.c_void_p, C.c_int]
        L.orbref_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orbref_pattern.argtypes = [C.c_void_p, C.c_void_p]
        L.orbref_extract.restype = C.c_int
        L.orbref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int]
        L.orbref_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orbref_copy_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
        L.orbref_bench.restype = C.c_double
        L.orbref_bench.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long)]
        L.orbref_arena_allocs.restype = C.c_long
        L.orbref_extract_many.restype = C.c_double
        L.orbref_extract_many.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_void_p]
        _ref = L
    return _ref


class RefExtractor:
    """The reference's ORBextractor (src/ORBextractor.cc) behind its own constructor signature."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, blur_mode=BLUR_CV331,
                 canonical=True):
        self.L = ref_lib()
        self.nlevels = nlevels
        self.blur_mode = blur_mode
        self.h = self.L.orbref_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.L.orbref_set_canonical(self.h, 1 if canonical else 0)
        self.cap = nfeatures + 4 * nlevels + 64

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orbref_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sf, isf, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        q = np.zeros(n, np.int32)
        um = np.zeros(16, np.int32)
        self.L.orbref_tables(self.h, sf.ctypes.data, isf.ctypes.data, s2.ctypes.data, is2.ctypes.data,
                             q.ctypes.data, um.ctypes.data)
        return dict(scale=sf, inv_scale=isf, sigma2=s2, inv_sigma2=is2, quotas=q, umax=um)

    def pattern(self):
        p = np.zeros(1024, np.int32)
        self.L.orbref_pattern(self.h, p.ctypes.data)
        return p

    def extract(self, img: np.ndarray, keep_pyramid=False):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        self.L.orbref_set_blur_mode(self.blur_mode)
        n = self.L.orbref_extract(self.h, img.ctypes.data, w, h, w, kps.ctypes.data, desc.ctypes.data, self.cap,
                                  1 if keep_pyramid else 0)
        if n < 0:
            return None
        assert n <= self.cap
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l: int, with_border=False) -> np.ndarray:
        w, h = C.c_int(), C.c_int()
        assert self.L.orbref_level_size(self.h, l, C.byref(w), C.byref(h)) == 0
        W, H = (w.value + 38, h.value + 38) if with_border else (w.value, h.value)
        out = np.zeros((H, W), np.uint8)
        assert self.L.orbref_copy_level(self.h, l, out.ctypes.data, W, 1 if with_border else 0) == 0
        return out


def ref_bench(frames: np.ndarray, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, threads=1,
              canonical=False, repeat=1, blur_mode=BLUR_CV331):
    """Wall seconds for the reference extractor over frames (n,h,w) on `threads` host threads."""
    L = ref_lib()
    L.orbref_set_blur_mode(blur_mode)
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    tot = C.c_long(0)
    secs = L.orbref_bench(nfeatures, scale_factor, nlevels, ini_th, min_th, frames.ctypes.data, n, w, h, threads,
                          1 if canonical else 0, repeat, C.byref(tot))
    return secs, tot.value


def ref_extract_many(frames: np.ndarray, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, threads=1,
                     canonical=True, blur_mode=BLUR_CV331, want_results=True):
    """The reference extractor over frames (n,h,w), frame-parallel on `threads` host threads, results kept.
    Returns (wall seconds, per-frame seconds [n], list of (keypoints, descriptors) or None)."""
    L = ref_lib()
    L.orbref_set_blur_mode(blur_mode)
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    cap = nfeatures + 4 * nlevels + 64
    cnt = np.zeros(n, np.int32)
    per = np.zeros(n, np.float64)
    kps = np.zeros((n, cap), KP_DTYPE) if want_results else None
    desc = np.zeros((n, cap, 32), np.uint8) if want_results else None
    secs = L.orbref_extract_many(nfeatures, scale_factor, nlevels, ini_th, min_th, frames.ctypes.data, n, w, h, threads,
                                 1 if canonical else 0, kps.ctypes.data if want_results else None,
                                 desc.ctypes.data if want_results else None, cap, cnt.ctypes.data, per.ctypes.data)
    res = None
    if want_results:
        assert int(cnt.max(initial=0)) <= cap
        res = [(kps[i, :cnt[i]], desc[i, :cnt[i]]) for i in range(n)]
    return secs, per, res


# ------------------------------------------------------------------------------------------------
# Independent restatement (oracle/orb_oracle.cc, oracle/match_oracle.cc)
_orc = None


def oracle_lib():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        vp, ci, cf, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.eaoo_resize.argtypes = [vp, ci, ci, sz, vp, ci, ci, sz]
        L.eaoo_border.argtypes = [vp, ci, ci, sz, vp, sz, ci]
        L.eaoo_fast.restype = ci
        L.eaoo_fast.argtypes = [vp, ci, ci, sz, ci, vp, ci]
        L.eaoo_blur.argtypes = [vp, ci, ci, sz, vp, sz, ci]
        L.eaoo_atan2.argtypes = [vp, vp, vp, ci]
        L.eaoo_sincosf.argtypes = [vp, vp, vp, C.c_long]
        L.eaoo_sincosf_sweep.restype = C.c_long
        L.eaoo_sincosf_sweep.argtypes = [cf, C.c_uint32]
        L.eaoo_tables.argtypes = [ci, cf, ci, vp, vp, vp, vp, vp, vp]
        L.eaoo_fast_cells.restype = ci
        L.eaoo_fast_cells.argtypes = [vp, ci, ci, ci, ci, vp, ci]
        L.eaoo_octree.restype = ci
        L.eaoo_octree.argtypes = [vp, ci, ci, ci, ci, vp, ci]
        L.eaoo_extract.restype = ci
        L.eaoo_extract.argtypes = [vp, ci, ci, sz, ci, cf, ci, ci, ci, ci, vp, vp, ci, vp, vp, vp, vp, ci]
        _orc = L
    return _orc


def o_resize(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    oracle_lib().eaoo_resize(src.ctypes.data, src.shape[1], src.shape[0], src.shape[1], dst.ctypes.data, dw, dh, dw)
    return dst


def o_border(src: np.ndarray, b: int = 19) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
    oracle_lib().eaoo_border(src.ctypes.data, w, h, w, dst.ctypes.data, w + 2 * b, b)
    return dst


def o_fast(img: np.ndarray, th: int) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = (w // 2 + 1) * (h // 2 + 1)
    out = np.zeros((cap, 3), np.int32)
    n = oracle_lib().eaoo_fast(img.ctypes.data, w, h, w, th, out.ctypes.data, cap)
    return out[:n].copy()


def o_blur(img: np.ndarray, mode: int) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    dst = np.zeros((h, w), np.uint8)
    oracle_lib().eaoo_blur(img.ctypes.data, w, h, w, dst.ctypes.data, w, mode)
    return dst


def o_atan2(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(len(y), np.float32)
    oracle_lib().eaoo_atan2(y.ctypes.data, x.ctypes.data, out.ctypes.data, len(y))
    return out


def o_sincosf(x: np.ndarray):
    x = np.ascontiguousarray(x, np.float32)
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    oracle_lib().eaoo_sincosf(x.ctypes.data, s.ctypes.data, c.ctypes.data, len(x))
    return s, c


def o_tables(nfeatures=1000, scale_factor=1.2, nlevels=8):
    n = nlevels
    sf, isf, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
    q = np.zeros(n, np.int32)
    um = np.zeros(16, np.int32)
    oracle_lib().eaoo_tables(nfeatures, scale_factor, nlevels, sf.ctypes.data, isf.ctypes.data, s2.ctypes.data,
                             is2.ctypes.data, q.ctypes.data, um.ctypes.data)
    return dict(scale=sf, inv_scale=isf, sigma2=s2, inv_sigma2=is2, quotas=q, umax=um)


def level_sizes(width, height, scale_factor=1.2, nlevels=8):
    t = o_tables(1000, scale_factor, nlevels)
    return [(int(np.rint(np.float32(width) * t["inv_scale"][l])), int(np.rint(np.float32(height) * t["inv_scale"][l])))
            for l in range(nlevels)]


def o_fast_cells(bordered: np.ndarray, ini_th=20, min_th=7) -> np.ndarray:
    bordered = np.ascontiguousarray(bordered, np.uint8)
    h, w = bordered.shape[0] - 38, bordered.shape[1] - 38
    cap = (w // 2 + 2) * (h // 2 + 2)
    out = np.zeros((cap, 3), np.int32)
    n = oracle_lib().eaoo_fast_cells(bordered.ctypes.data, w, h, ini_th, min_th, out.ctypes.data, cap)
    return out[:n].copy()


def o_octree(cands: np.ndarray, W: int, H: int, N: int) -> np.ndarray:
    cands = np.ascontiguousarray(cands, np.int32)
    cap = max(len(cands), 1)
    sel = np.zeros(cap, np.int32)
    n = oracle_lib().eaoo_octree(cands.ctypes.data, len(cands), W, H, N, sel.ctypes.data, cap)
    return sel[:n].copy()


def o_extract(img: np.ndarray, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7,
              blur_mode=BLUR_CV331, dumps=False):
    """Full restated operator().  Returns (kps, desc) or, with dumps, (kps, desc, pyr_levels, blur_levels, cands)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = nfeatures + 4 * nlevels + 64
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    L = oracle_lib()
    if not dumps:
        n = L.eaoo_extract(img.ctypes.data, w, h, w, nfeatures, scale_factor, nlevels, ini_th, min_th, blur_mode,
                           kps.ctypes.data, desc.ctypes.data, cap, None, None, None, None, 0)
        if n < 0:
            return None
        return kps[:n].copy(), desc[:n].copy()
    sizes = level_sizes(w, h, scale_factor, nlevels)
    pyr = np.zeros(sum((a + 38) * (b + 38) for a, b in sizes), np.uint8)
    blur = np.zeros(sum(a * b for a, b in sizes), np.uint8)
    ccap = sum((a // 2 + 2) * (b // 2 + 2) for a, b in sizes)
    cand = np.zeros((ccap, 3), np.int32)
    ccnt = np.zeros(nlevels, np.int32)
    n = L.eaoo_extract(img.ctypes.data, w, h, w, nfeatures, scale_factor, nlevels, ini_th, min_th, blur_mode,
                       kps.ctypes.data, desc.ctypes.data, cap, pyr.ctypes.data, blur.ctypes.data, cand.ctypes.data,
                       ccnt.ctypes.data, ccap)
    pl, bl, cl = [], [], []
    po = bo = co = 0
    for l, (a, b) in enumerate(sizes):
        pl.append(pyr[po:po + (a + 38) * (b + 38)].reshape(b + 38, a + 38)); po += (a + 38) * (b + 38)
        bl.append(blur[bo:bo + a * b].reshape(b, a)); bo += a * b
        cl.append(cand[co:co + ccnt[l]].copy()); co += ccnt[l]
    return kps[:n].copy(), desc[:n].copy(), pl, bl, cl


# ------------------------------------------------------------------------------------------------
# Matcher restatement (oracle/match_oracle.cc)
_m_ready = False


def _mo():
    global _m_ready
    L = oracle_lib()
    if not _m_ready:
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.eaoo_hamming.restype = ci
        L.eaoo_hamming.argtypes = [vp, vp]
        L.eaoo_three_maxima.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
        L.eaoo_search_by_bow.restype = ci
        L.eaoo_search_by_bow.argtypes = [ci, ci, vp, vp, vp, ci, vp, vp, vp, ci, vp, vp, vp, ci, vp, vp, vp, cf, ci, vp, vp]
        L.eaoo_search_for_triangulation.restype = ci
        L.eaoo_build_grid.argtypes = [ci, vp, vp, cf, cf, cf, cf, vp, vp]
        L.eaoo_search_by_projection_last.restype = ci
        L.eaoo_search_by_projection_last.argtypes = [ci, vp, vp, vp, vp, vp, vp, vp, cf, cf, cf, cf, cf, cf, ci, vp, vp, vp,
                                                     vp, vp, vp, vp, vp, vp, cf, cf, ci, ci, vp, vp]
        _m_ready = True
    return L


def _a(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


def _pp(a):
    return None if a is None else a.ctypes.data


def o_hamming(a, b):
    a = _a(a, np.uint8).reshape(-1, 32)
    b = _a(b, np.uint8).reshape(-1, 32)
    L = _mo()
    return np.array([L.eaoo_hamming(a[i].ctypes.data, b[i].ctypes.data) for i in range(len(a))], np.int32)


def o_three_maxima(sizes):
    s = _a(sizes, np.int32)
    i1, i2, i3 = C.c_int(), C.c_int(), C.c_int()
    _mo().eaoo_three_maxima(s.ctypes.data, len(s), C.byref(i1), C.byref(i2), C.byref(i3))
    return i1.value, i2.value, i3.value


def o_search_by_bow(mode, nnratio, check_ori, desc_q, angle_q, valid_q, nodes_q, desc_t, angle_t, valid_t, nodes_t):
    dq, dt = _a(desc_q, np.uint8), _a(desc_t, np.uint8)
    aq, at = _a(angle_q, np.float32), _a(angle_t, np.float32)
    vq, vt = _a(valid_q, np.uint8), _a(valid_t, np.uint8)
    iq, sq, xq = (_a(v, np.int32) for v in nodes_q)
    it, st, xt = (_a(v, np.int32) for v in nodes_t)
    nq, nt = len(dq), len(dt)
    nout = nt if mode == 0 else nq
    match = np.full(max(nout, 1), -1, np.int32)
    dist = np.full(max(nout, 1), -1, np.int32)
    n = _mo().eaoo_search_by_bow(mode, nq, _pp(dq), _pp(aq), _pp(vq), nt, _pp(dt), _pp(at), _pp(vt), len(iq), _pp(iq), _pp(sq),
                                 _pp(xq), len(it), _pp(it), _pp(st), _pp(xt), nnratio, int(check_ori), match.ctypes.data,
                                 dist.ctypes.data)
    return n, match[:nout], dist[:nout]


def o_build_grid(x, y, min_x, min_y, inv_w, inv_h):
    x, y = _a(x, np.float32), _a(y, np.float32)
    cs = np.zeros(64 * 48 + 1, np.int32)
    ci = np.zeros(max(len(x), 1), np.int32)
    _mo().eaoo_build_grid(len(x), x.ctypes.data, y.ctypes.data, min_x, min_y, inv_w, inv_h, cs.ctypes.data, ci.ctypes.data)
    return cs, ci[:cs[-1]]


def o_search_by_projection(cur, last, th, check_ori, *, bounds, grid_inv, scale_factors, mbf=0.0, search_mode=0):
    cx, cy = _a(cur["x"], np.float32), _a(cur["y"], np.float32)
    co, ca, cd = _a(cur["octave"], np.int32), _a(cur["angle"], np.float32), _a(cur["desc"], np.uint8)
    cur_r, ctk = _a(cur.get("uright"), np.float32), _a(cur.get("taken"), np.uint8)
    lu, lv = _a(last["u"], np.float32), _a(last["v"], np.float32)
    lo, la, ld = _a(last["octave"], np.int32), _a(last["angle"], np.float32), _a(last["desc"], np.uint8)
    lval, linv, lobs = _a(last.get("valid"), np.uint8), _a(last.get("invz"), np.float32), _a(last.get("obs"), np.uint8)
    sf = _a(scale_factors, np.float32)
    nc, nl = len(cx), len(lu)
    match = np.full(max(nc, 1), -1, np.int32)
    dist = np.full(max(nc, 1), -1, np.int32)
    n = _mo().eaoo_search_by_projection_last(nc, _pp(cx), _pp(cy), _pp(co), _pp(ca), _pp(cd), _pp(cur_r), _pp(ctk), bounds[0],
                                             bounds[1], bounds[2], bounds[3], grid_inv[0], grid_inv[1], nl, _pp(lval), _pp(lu),
                                             _pp(lv), _pp(linv), _pp(lo), _pp(la), _pp(ld), _pp(lobs), _pp(sf), float(th),
                                             float(mbf), search_mode, int(check_ori), match.ctypes.data, dist.ctypes.data)
    return n, match[:nc], dist[:nc]


def o_search_for_triangulation(k1, k2, F12, epipole, scale_factors, level_sigma2, only_stereo, check_ori):
    """k1: dict(desc,x,y,angle,free[,stereo],nodes); k2: dict(desc,x,y,octave,angle,free[,stereo],nodes).
    Returns (nmatches, match12, dist12).  src/ORBmatcher.cc:657-823."""
    d1, d2 = _a(k1["desc"], np.uint8), _a(k2["desc"], np.uint8)
    x1, y1, a1 = (_a(k1[k], np.float32) for k in ("x", "y", "angle"))
    x2, y2, a2 = (_a(k2[k], np.float32) for k in ("x", "y", "angle"))
    o2 = _a(k2["octave"], np.int32)
    f1, f2 = _a(k1["free"], np.uint8), _a(k2["free"], np.uint8)
    s1, s2 = _a(k1.get("stereo"), np.uint8), _a(k2.get("stereo"), np.uint8)
    i1, st1, ix1 = (_a(v, np.int32) for v in k1["nodes"])
    i2, st2, ix2 = (_a(v, np.int32) for v in k2["nodes"])
    F = _a(F12, np.float32).reshape(9)
    sf, ls = _a(scale_factors, np.float32), _a(level_sigma2, np.float32)
    n1, n2 = len(d1), len(d2)
    match = np.full(max(n1, 1), -1, np.int32)
    dist = np.full(max(n1, 1), -1, np.int32)
    L = _mo()
    L.eaoo_search_for_triangulation.argtypes = ([C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 7 +
                                                [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 +
                                                [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.c_int, C.c_void_p, C.c_void_p])
    n = L.eaoo_search_for_triangulation(n1, _pp(d1), _pp(x1), _pp(y1), _pp(a1), _pp(f1), _pp(s1), n2, _pp(d2), _pp(x2),
                                        _pp(y2), _pp(o2), _pp(a2), _pp(f2), _pp(s2), len(i1), _pp(i1), _pp(st1), _pp(ix1),
                                        len(i2), _pp(i2), _pp(st2), _pp(ix2), _pp(F), float(epipole[0]), float(epipole[1]),
                                        _pp(sf), _pp(ls), int(only_stereo), int(check_ori), match.ctypes.data,
                                        dist.ctypes.data)
    return n, match[:n1], dist[:n1]


# ------------------------------------------------------------------------------------------------------------
# The UNMODIFIED reference ORBmatcher.cc (oracle/_ref/libmatch_ref.so, built by `make -C oracle matchref`), driven
# through oracle/match_ref_harness.cc with the same arrays as the o_* restatements above.
MATCH_REF_SO = os.path.join(HERE, "_ref", "libmatch_ref.so")
_mref = None


def _setup_mref(L):
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    L.mref_hamming.argtypes = [vp, vp]
    L.mref_predict_scale.argtypes = [cf, cf, cf]
    L.mref_norm3.restype = cf
    L.mref_norm3.argtypes = [cf, cf, cf]
    L.mref_search_by_bow.argtypes = [ci, ci, vp, vp, vp, ci, vp, vp, vp, ci, vp, vp, vp, ci, vp, vp, vp, cf, ci, vp]
    L.mref_search_for_triangulation.argtypes = ([ci] + [vp] * 6 + [ci] + [vp] * 7 + [ci] + [vp] * 3 + [ci] + [vp] * 3 +
                                                [vp, cf, cf, vp, vp, ci, ci, ci, cf, vp])
    L.mref_search_by_projection_last.argtypes = [ci, vp, vp, vp, vp, vp, vp, vp, cf, cf, cf, cf, cf, cf, ci, vp, vp, vp,
                                                 vp, vp, vp, vp, vp, vp, ci, cf, cf, ci, ci, cf, vp]
    L.mref_search_by_projection_mappoints.argtypes = [ci, vp, vp, vp, vp, vp, vp, cf, cf, cf, cf, cf, cf, ci, vp, vp, vp,
                                                      vp, vp, vp, vp, vp, vp, vp, ci, cf, cf, vp]
    L.mref_search_for_initialization.argtypes = [ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, cf, cf, cf, cf, cf, cf, ci,
                                                 cf, ci, vp]
    L.mref_search_by_projection_kf.argtypes = [ci, vp, vp, vp, vp, vp, vp, cf, cf, cf, cf, cf, cf, ci, vp, vp, vp, vp,
                                               vp, vp, vp, vp, vp, ci, cf, cf, ci, ci, cf, vp]
    L.mref_dot3.restype = C.c_double
    L.mref_dot3.argtypes = [cf] * 6
    L.mref_search_by_projection_sim3kf.argtypes = [ci] + [vp] * 5 + [ci] * 4 + [cf, cf, ci] + [vp] * 9 + [ci, cf, cf, ci, vp]
    L.mref_fuse.argtypes = ([ci] + [vp] * 7 + [ci] * 4 + [cf, cf, vp, vp, ci, cf, cf, ci] + [vp] * 10 + [cf] + [vp] * 4)
    L.mref_fuse_sim3.argtypes = [ci] + [vp] * 6 + [ci] * 4 + [cf, cf, vp, ci, cf, cf, ci] + [vp] * 8 + [cf] + [vp] * 3
    L.mref_search_by_sim3.argtypes = ([ci] + [vp] * 11 + [ci] + [vp] * 11 + [vp] + [ci] * 4 + [cf, cf, vp, ci, cf, cf, cf, vp])
    return L


def match_ref_lib():
    global _mref
    if _mref is None:
        if not os.path.exists(MATCH_REF_SO):
            subprocess.check_call(["make", "-s", "-C", HERE, "matchref"])
        _mref = _setup_mref(C.CDLL(MATCH_REF_SO))
    return _mref


def match_harness_lib(path):
    """The same C harness (oracle/match_ref_harness.cc) linked against another implementation of ORB_SLAM2::ORBmatcher,
    e.g. the drop-in class over the C ABI (tests/cpp/_build/libmatch_dropin.so).  Use as r_*(..., L=lib)."""
    return _setup_mref(C.CDLL(path))


def r_hamming(a, b, L=None):
    a = _a(a, np.uint8).reshape(-1, 32)
    b = _a(b, np.uint8).reshape(-1, 32)
    L = L or match_ref_lib()
    return np.array([L.mref_hamming(a[i].ctypes.data, b[i].ctypes.data) for i in range(len(a))], np.int32)


def r_search_by_bow(mode, nnratio, check_ori, desc_q, angle_q, valid_q, nodes_q, desc_t, angle_t, valid_t, nodes_t, L=None):
    dq, dt = _a(desc_q, np.uint8), _a(desc_t, np.uint8)
    aq, at = _a(angle_q, np.float32), _a(angle_t, np.float32)
    vq, vt = _a(valid_q, np.uint8), _a(valid_t, np.uint8)
    iq, sq, xq = (_a(v, np.int32) for v in nodes_q)
    it, st, xt = (_a(v, np.int32) for v in nodes_t)
    nq, nt = len(dq), len(dt)
    nout = nt if mode == 0 else nq
    match = np.full(max(nout, 1), -1, np.int32)
    n = (L or match_ref_lib()).mref_search_by_bow(mode, nq, _pp(dq), _pp(aq), _pp(vq), nt, _pp(dt), _pp(at), _pp(vt), len(iq),
                                           _pp(iq), _pp(sq), _pp(xq), len(it), _pp(it), _pp(st), _pp(xt), nnratio,
                                           int(check_ori), match.ctypes.data)
    return n, match[:nout]


def r_search_for_triangulation(k1, k2, F12, epipole, scale_factors, level_sigma2, only_stereo, check_ori, L=None):
    d1, d2 = _a(k1["desc"], np.uint8), _a(k2["desc"], np.uint8)
    x1, y1, a1 = (_a(k1[k], np.float32) for k in ("x", "y", "angle"))
    x2, y2, a2 = (_a(k2[k], np.float32) for k in ("x", "y", "angle"))
    o2 = _a(k2["octave"], np.int32)
    f1, f2 = _a(k1["free"], np.uint8), _a(k2["free"], np.uint8)
    s1, s2 = _a(k1.get("stereo"), np.uint8), _a(k2.get("stereo"), np.uint8)
    i1, st1, ix1 = (_a(v, np.int32) for v in k1["nodes"])
    i2, st2, ix2 = (_a(v, np.int32) for v in k2["nodes"])
    F = _a(F12, np.float32).reshape(9)
    sf, ls = _a(scale_factors, np.float32), _a(level_sigma2, np.float32)
    n1, n2 = len(d1), len(d2)
    match = np.full(max(n1, 1), -1, np.int32)
    n = (L or match_ref_lib()).mref_search_for_triangulation(n1, _pp(d1), _pp(x1), _pp(y1), _pp(a1), _pp(f1), _pp(s1), n2, _pp(d2),
                                                      _pp(x2), _pp(y2), _pp(o2), _pp(a2), _pp(f2), _pp(s2), len(i1), _pp(i1),
                                                      _pp(st1), _pp(ix1), len(i2), _pp(i2), _pp(st2), _pp(ix2), _pp(F),
                                                      float(epipole[0]), float(epipole[1]), _pp(sf), _pp(ls), len(sf),
                                                      int(only_stereo), int(check_ori), 0.6, match.ctypes.data)
    return n, match[:n1]


def r_search_by_projection(cur, last, th, check_ori, *, bounds, grid_inv, scale_factors, mbf=0.0, search_mode=0, L=None):
    cx, cy = _a(cur["x"], np.float32), _a(cur["y"], np.float32)
    co, ca, cd = _a(cur["octave"], np.int32), _a(cur["angle"], np.float32), _a(cur["desc"], np.uint8)
    cur_r, ctk = _a(cur.get("uright"), np.float32), _a(cur.get("taken"), np.uint8)
    lu, lv = _a(last["u"], np.float32), _a(last["v"], np.float32)
    lo, la, ld = _a(last["octave"], np.int32), _a(last["angle"], np.float32), _a(last["desc"], np.uint8)
    lval, linv, lobs = _a(last.get("valid"), np.uint8), _a(last.get("invz"), np.float32), _a(last.get("obs"), np.uint8)
    sf = _a(scale_factors, np.float32)
    nc, nl = len(cx), len(lu)
    match = np.full(max(nc, 1), -1, np.int32)
    n = (L or match_ref_lib()).mref_search_by_projection_last(nc, _pp(cx), _pp(cy), _pp(co), _pp(ca), _pp(cd), _pp(cur_r), _pp(ctk),
                                                       bounds[0], bounds[1], bounds[2], bounds[3], grid_inv[0], grid_inv[1],
                                                       nl, _pp(lval), _pp(lu), _pp(lv), _pp(linv), _pp(lo), _pp(la), _pp(ld),
                                                       _pp(lobs), _pp(sf), len(sf), float(th), float(mbf), search_mode,
                                                       int(check_ori), 0.9, match.ctypes.data)
    return n, match[:nc]


def _frame_arrays(F):
    return (_a(F["x"], np.float32), _a(F["y"], np.float32), _a(F["octave"], np.int32), _a(F.get("angle"), np.float32),
            _a(F["desc"], np.uint8), _a(F.get("uright"), np.float32), _a(F.get("taken"), np.uint8))


def o_search_by_projection_mappoints(F, mp, th, nnratio, *, bounds, grid_inv, scale_factors):
    """F: dict(x,y,octave,desc[,uright,taken]); mp: dict(x,y,level,desc[,in_view,bad,xr,cos,obs]).
    Returns (nmatches, matchF, distF).  src/ORBmatcher.cc:45-129."""
    fx, fy, fo, _, fd, fr, ft = _frame_arrays(F)
    px, py, pl, pd = _a(mp["x"], np.float32), _a(mp["y"], np.float32), _a(mp["level"], np.int32), _a(mp["desc"], np.uint8)
    iv, bad, pxr = _a(mp.get("in_view"), np.uint8), _a(mp.get("bad"), np.uint8), _a(mp.get("xr"), np.float32)
    pc, po_ = _a(mp.get("cos"), np.float32), _a(mp.get("obs"), np.uint8)
    sf = _a(scale_factors, np.float32)
    nF, nM = len(fx), len(px)
    match, dist = np.full(max(nF, 1), -1, np.int32), np.full(max(nF, 1), -1, np.int32)
    L = _mo()
    ci, cf, vp = C.c_int, C.c_float, C.c_void_p
    L.eaoo_search_by_projection_mappoints.argtypes = [ci, vp, vp, vp, vp, vp, vp, cf, cf, cf, cf, ci, vp, vp, vp, vp, vp, vp,
                                                      vp, vp, vp, vp, cf, cf, vp, vp]
    n = L.eaoo_search_by_projection_mappoints(nF, _pp(fx), _pp(fy), _pp(fo), _pp(fd), _pp(fr), _pp(ft), bounds[0], bounds[2],
                                              grid_inv[0], grid_inv[1], nM, _pp(iv), _pp(bad), _pp(px), _pp(py), _pp(pxr),
                                              _pp(pl), _pp(pc), _pp(pd), _pp(po_), _pp(sf), float(th), float(nnratio),
                                              match.ctypes.data, dist.ctypes.data)
    return n, match[:nF], dist[:nF]


def r_search_by_projection_mappoints(F, mp, th, nnratio, *, bounds, grid_inv, scale_factors, L=None):
    fx, fy, fo, _, fd, fr, ft = _frame_arrays(F)
    px, py, pl, pd = _a(mp["x"], np.float32), _a(mp["y"], np.float32), _a(mp["level"], np.int32), _a(mp["desc"], np.uint8)
    iv, bad, pxr = _a(mp.get("in_view"), np.uint8), _a(mp.get("bad"), np.uint8), _a(mp.get("xr"), np.float32)
    pc, po_ = _a(mp.get("cos"), np.float32), _a(mp.get("obs"), np.uint8)
    sf = _a(scale_factors, np.float32)
    nF, nM = len(fx), len(px)
    match = np.full(max(nF, 1), -1, np.int32)
    n = (L or match_ref_lib()).mref_search_by_projection_mappoints(nF, _pp(fx), _pp(fy), _pp(fo), _pp(fd), _pp(fr), _pp(ft),
                                                            bounds[0], bounds[1], bounds[2], bounds[3], grid_inv[0],
                                                            grid_inv[1], nM, _pp(iv), _pp(bad), _pp(px), _pp(py), _pp(pxr),
                                                            _pp(pl), _pp(pc), _pp(pd), _pp(po_), _pp(sf), len(sf), float(th),
                                                            float(nnratio), match.ctypes.data)
    return n, match[:nF]


def kf_projection(kf, log_scale_factor, n_levels):
    """What the caller of SearchByProjection(Cur,KF) derives per KF map point before the search (:1497-1525), computed
    with the reference binary's own helpers: pixel (u, v) for an identity pose with fx=fy=1, cx=cy=0 and wz=1, the 3-D
    distance gate and MapPoint::PredictScale.  Returns (valid, u, v, level)."""
    L = match_ref_lib()
    st, wx, wy, wz = _a(kf["state"], np.uint8), _a(kf["wx"], np.float32), _a(kf["wy"], np.float32), _a(kf["wz"], np.float32)
    mx, mn = _a(kf["max_dist"], np.float32), _a(kf["min_dist"], np.float32)
    n = len(st)
    valid, u, v, lvl = np.zeros(n, np.uint8), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
    for i in range(n):
        assert wz[i] == 1.0
        u[i], v[i] = wx[i], wy[i]
        d = np.float32(L.mref_norm3(float(wx[i]), float(wy[i]), float(wz[i])))
        inside = not (d < np.float32(0.8) * mn[i] or d > np.float32(1.2) * mx[i])
        valid[i] = st[i] == 3 and inside
        lvl[i] = L.mref_predict_scale(float(mx[i]), float(d), float(log_scale_factor)) if inside else 0
        if valid[i]:
            assert 0 <= lvl[i] < n_levels, "test data must keep the predicted level inside the pyramid"
    return valid, u, v, lvl


def o_search_by_projection_kf(cur, kq, th, orb_dist, check_ori, *, bounds, grid_inv, scale_factors):
    """cur: dict(x,y,octave,angle,desc[,taken]); kq: dict(valid,u,v,level,angle,desc) (see kf_projection).
    Returns (nmatches, matchCur, distCur).  src/ORBmatcher.cc:1474-1601."""
    cx, cy, co, ca, cd, _, ct = _frame_arrays(cur)
    kv, ku, kvv, kl = _a(kq["valid"], np.uint8), _a(kq["u"], np.float32), _a(kq["v"], np.float32), _a(kq["level"], np.int32)
    ka, kd = _a(kq["angle"], np.float32), _a(kq["desc"], np.uint8)
    sf = _a(scale_factors, np.float32)
    nC, nK = len(cx), len(ku)
    match, dist = np.full(max(nC, 1), -1, np.int32), np.full(max(nC, 1), -1, np.int32)
    L = _mo()
    ci, cf, vp = C.c_int, C.c_float, C.c_void_p
    L.eaoo_search_by_projection_kf.argtypes = [ci, vp, vp, vp, vp, vp, vp, cf, cf, cf, cf, cf, cf, ci, vp, vp, vp, vp, vp, vp,
                                               vp, cf, ci, ci, vp, vp]
    n = L.eaoo_search_by_projection_kf(nC, _pp(cx), _pp(cy), _pp(co), _pp(ca), _pp(cd), _pp(ct), bounds[0], bounds[1],
                                       bounds[2], bounds[3], grid_inv[0], grid_inv[1], nK, _pp(kv), _pp(ku), _pp(kvv),
                                       _pp(kl), _pp(ka), _pp(kd), _pp(sf), float(th), int(orb_dist), int(check_ori),
                                       match.ctypes.data, dist.ctypes.data)
    return n, match[:nC], dist[:nC]


def r_search_by_projection_kf(cur, kf, th, orb_dist, check_ori, *, bounds, grid_inv, scale_factors, log_scale_factor, L=None):
    """kf: dict(state,wx,wy,wz,max_dist,min_dist,angle,desc); state 0 none, 1 bad, 2 already found, 3 usable."""
    cx, cy, co, ca, cd, _, ct = _frame_arrays(cur)
    st, wx, wy, wz = _a(kf["state"], np.uint8), _a(kf["wx"], np.float32), _a(kf["wy"], np.float32), _a(kf["wz"], np.float32)
    mx, mn = _a(kf["max_dist"], np.float32), _a(kf["min_dist"], np.float32)
    ka, kd = _a(kf["angle"], np.float32), _a(kf["desc"], np.uint8)
    sf = _a(scale_factors, np.float32)
    nC, nK = len(cx), len(st)
    match = np.full(max(nC, 1), -1, np.int32)
    n = (L or match_ref_lib()).mref_search_by_projection_kf(nC, _pp(cx), _pp(cy), _pp(co), _pp(ca), _pp(cd), _pp(ct), bounds[0],
                                                     bounds[1], bounds[2], bounds[3], grid_inv[0], grid_inv[1], nK, _pp(st),
                                                     _pp(wx), _pp(wy), _pp(wz), _pp(mx), _pp(mn), _pp(ka), _pp(kd), _pp(sf),
                                                     len(sf), float(log_scale_factor), float(th), int(orb_dist),
                                                     int(check_ori), 0.9, match.ctypes.data)
    return n, match[:nC]


def _init_args(F1, F2, prev):
    return (_a(F1["octave"], np.int32), _a(F1["angle"], np.float32), _a(F1["desc"], np.uint8),
            np.ascontiguousarray(prev, np.float32).copy(), _a(F2["x"], np.float32), _a(F2["y"], np.float32),
            _a(F2["octave"], np.int32), _a(F2["angle"], np.float32), _a(F2["desc"], np.uint8))


def o_search_for_initialization(F1, F2, prev, window, nnratio, check_ori, *, bounds, grid_inv):
    """Returns (nmatches, matches12, updated prev).  src/ORBmatcher.cc:405-520."""
    o1, a1, d1, pm, x2, y2, o2, a2, d2 = _init_args(F1, F2, prev)
    m = np.full(max(len(o1), 1), -1, np.int32)
    L = _mo()
    ci, cf, vp = C.c_int, C.c_float, C.c_void_p
    L.eaoo_search_for_initialization.argtypes = [ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, cf, cf, cf, cf, ci, cf, ci, vp]
    n = L.eaoo_search_for_initialization(len(o1), _pp(o1), _pp(a1), _pp(d1), _pp(pm), len(x2), _pp(x2), _pp(y2), _pp(o2),
                                         _pp(a2), _pp(d2), bounds[0], bounds[2], grid_inv[0], grid_inv[1], int(window),
                                         float(nnratio), int(check_ori), m.ctypes.data)
    return n, m[:len(o1)], pm


def r_search_for_initialization(F1, F2, prev, window, nnratio, check_ori, *, bounds, grid_inv, L=None):
    o1, a1, d1, pm, x2, y2, o2, a2, d2 = _init_args(F1, F2, prev)
    m = np.full(max(len(o1), 1), -1, np.int32)
    n = (L or match_ref_lib()).mref_search_for_initialization(len(o1), _pp(o1), _pp(a1), _pp(d1), _pp(pm), len(x2), _pp(x2), _pp(y2),
                                                       _pp(o2), _pp(a2), _pp(d2), bounds[0], bounds[1], bounds[2], bounds[3],
                                                       grid_inv[0], grid_inv[1], int(window), float(nnratio), int(check_ori),
                                                       m.ctypes.data)
    return n, m[:len(o1)], pm


# ---- map-side matchers (LocalMapping / LoopClosing): SearchByProjection(KF,Scw), Fuse x2, SearchBySim3 -------------------

def map_projection(pts, log_scale_factor, n_levels, *, bounds, check_normal=True, scale=1.0, L=None):
    """What the map-side loops derive per map point before the descriptor search (src/ORBmatcher.cc:318-356, :852-885,
    :1160-1190), for the harness geometry (identity pose, fx=fy=1, cx=cy=0, wz=+-1, camera-frame point = scale * world
    point), computed with the reference binary's own helpers (cv::norm, Mat::dot, MapPoint::PredictScale as compiled
    into libmatch_ref.so).  pts: dict(wx,wy,wz,max_dist,min_dist[,normal]).  Returns (valid, u, v, level): valid = passes
    depth, image bounds, distance range and viewing angle."""
    L = L or match_ref_lib()
    wx, wy, wz = _a(pts["wx"], np.float32), _a(pts["wy"], np.float32), _a(pts["wz"], np.float32)
    mx, mn = _a(pts["max_dist"], np.float32), _a(pts["min_dist"], np.float32)
    nrm = _a(pts.get("normal"), np.float32)
    n = len(wx)
    sc = np.float32(scale)
    valid, u, v, lvl = np.zeros(n, np.uint8), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
    for i in range(n):
        assert abs(wz[i]) == 1.0
        if wz[i] < 0:
            continue
        u[i], v[i] = wx[i], wy[i]
        if not (u[i] >= bounds[0] and u[i] < bounds[1] and v[i] >= bounds[2] and v[i] < bounds[3]):  # KeyFrame::IsInImage
            continue
        d = np.float32(L.mref_norm3(float(sc * wx[i]), float(sc * wy[i]), float(sc * wz[i])))
        if d < np.float32(0.8) * mn[i] or d > np.float32(1.2) * mx[i]:
            continue
        if check_normal:
            nx, ny, nz = (nrm[i] if nrm is not None else (wx[i], wy[i], wz[i]))
            if L.mref_dot3(float(wx[i]), float(wy[i]), float(wz[i]), float(nx), float(ny), float(nz)) < 0.5 * float(d):
                continue
        lvl[i] = L.mref_predict_scale(float(mx[i]), float(d), float(log_scale_factor))
        assert 0 <= lvl[i] < n_levels, "test data must keep the predicted level inside the pyramid"
        valid[i] = 1
    return valid, u, v, lvl


def _kf_arrays(KF):
    return (_a(KF["x"], np.float32), _a(KF["y"], np.float32), _a(KF["octave"], np.int32), _a(KF["desc"], np.uint8),
            _a(KF.get("uright"), np.float32))


def _pt_arrays(pts):
    return (_a(pts["wx"], np.float32), _a(pts["wy"], np.float32), _a(pts["wz"], np.float32), _a(pts["max_dist"], np.float32),
            _a(pts["min_dist"], np.float32), _a(pts.get("normal"), np.float32), _a(pts["desc"], np.uint8))


def o_search_by_projection_sim3kf(KF, q, th, *, bounds, grid_inv, scale_factors):
    """KF: dict(x,y,octave,desc, matched = flags vpMatched[k] != NULL); q: dict(valid,u,v,level,desc).
    Returns (nmatches, matchT, distT).  src/ORBmatcher.cc:290-403."""
    tx, ty, to, td, _ = _kf_arrays(KF)
    tm = _a(KF.get("matched"), np.uint8)
    qv, qu, qvv, ql, qd = (_a(q["valid"], np.uint8), _a(q["u"], np.float32), _a(q["v"], np.float32), _a(q["level"], np.int32),
                           _a(q["desc"], np.uint8))
    sf = _a(scale_factors, np.float32)
    nT, nQ = len(tx), len(qu)
    match, dist = np.full(max(nT, 1), -1, np.int32), np.full(max(nT, 1), -1, np.int32)
    L = _mo()
    ci, cf, vp = C.c_int, C.c_float, C.c_void_p
    L.eaoo_search_by_projection_sim3kf.argtypes = [ci] + [vp] * 5 + [cf] * 4 + [ci] + [vp] * 6 + [ci, vp, vp]
    n = L.eaoo_search_by_projection_sim3kf(nT, _pp(tx), _pp(ty), _pp(to), _pp(td), _pp(tm), bounds[0], bounds[2], grid_inv[0],
                                           grid_inv[1], nQ, _pp(qv), _pp(qu), _pp(qvv), _pp(ql), _pp(qd), _pp(sf), int(th),
                                           match.ctypes.data, dist.ctypes.data)
    return n, match[:nT], dist[:nT]


def r_search_by_projection_sim3kf(KF, pts, th, *, bounds, grid_inv, scale_factors, log_scale_factor, scw_scale=1.0, L=None):
    """KF as above with matched_q (int per feature: -1 none, -2 foreign point, i = vpPoints[i]); pts: dict(wx,wy,wz,
    max_dist,min_dist,normal,desc,bad).  Returns (nmatches, tag of vpMatched[k] afterwards)."""
    tx, ty, to, td, _ = _kf_arrays(KF)
    tm = _a(KF["matched_q"], np.int32)
    wx, wy, wz, mx, mn, nrm, qd = _pt_arrays(pts)
    bad = _a(pts.get("bad"), np.uint8)
    sf = _a(scale_factors, np.float32)
    nT, nQ = len(tx), len(wx)
    match = np.full(max(nT, 1), -1, np.int32)
    n = (L or match_ref_lib()).mref_search_by_projection_sim3kf(
        nT, _pp(tx), _pp(ty), _pp(to), _pp(td), _pp(tm), int(bounds[0]), int(bounds[1]), int(bounds[2]), int(bounds[3]),
        grid_inv[0], grid_inv[1], nQ, _pp(bad), _pp(wx), _pp(wy), _pp(wz), _pp(mx), _pp(mn), _pp(nrm), _pp(qd), _pp(sf), len(sf),
        float(log_scale_factor), float(scw_scale), int(th), match.ctypes.data)
    return n, match[:nT]


def o_window_best(gate, KF, q, th, th_accept, *, bounds, grid_inv, scale_factors, inv_level_sigma2=None):
    """The search step of Fuse x2 / SearchBySim3.  q: dict(valid,u,v,level,desc[,ur]).  Returns (n, matchQ, distQ)."""
    tx, ty, to, td, tr = _kf_arrays(KF)
    qv, qu, qvv, ql, qd = (_a(q["valid"], np.uint8), _a(q["u"], np.float32), _a(q["v"], np.float32), _a(q["level"], np.int32),
                           _a(q["desc"], np.uint8))
    qur, inv, sf = _a(q.get("ur"), np.float32), _a(inv_level_sigma2, np.float32), _a(scale_factors, np.float32)
    nT, nQ = len(tx), len(qu)
    match, dist = np.full(max(nQ, 1), -1, np.int32), np.full(max(nQ, 1), -1, np.int32)
    L = _mo()
    ci, cf, vp = C.c_int, C.c_float, C.c_void_p
    L.eaoo_window_best.argtypes = [ci, ci] + [vp] * 6 + [cf] * 4 + [ci] + [vp] * 7 + [cf, ci, vp, vp]
    n = L.eaoo_window_best(int(gate), nT, _pp(tx), _pp(ty), _pp(to), _pp(td), _pp(tr), _pp(inv), bounds[0], bounds[2],
                           grid_inv[0], grid_inv[1], nQ, _pp(qv), _pp(qu), _pp(qvv), _pp(qur), _pp(ql), _pp(qd), _pp(sf),
                           float(th), int(th_accept), match.ctypes.data, dist.ctypes.data)
    return n, match[:nQ], dist[:nQ]


def o_fuse_apply(match_q, qstate, qid, qobs, slot_state, slot_obs):
    """Bookkeeping of Fuse(KeyFrame*, vpMapPoints, th) on array state (eaoo_fuse_apply).  Returns (nFused, addedAt,
    replacedBy, slotReplacedBy, slotHolder) in the tag convention of mref_fuse (candidate -> qid, entry point of feature
    k -> 100000+k)."""
    nQ, nT = len(match_q), len(slot_state)
    TAG = 100000
    qstate, slot_state, qid = np.asarray(qstate), np.asarray(slot_state), _a(qid, np.int32)
    slot = np.where(slot_state > 0, nQ + np.arange(nT), -1).astype(np.int32)
    bad = np.zeros(nQ + nT, np.uint8); obs = np.zeros(nQ + nT, np.int32); inkf = np.zeros(nQ + nT, np.uint8)
    bad[:nQ] = qstate == 1; inkf[:nQ] = qstate == 2; obs[:nQ] = qobs
    bad[nQ:] = slot_state == 2; obs[nQ:] = slot_obs; inkf[nQ:] = slot_state > 0
    repl = np.full(nQ + nT, -1, np.int32)
    added = np.full(max(nQ, 1), -1, np.int32)
    qnull = (qstate == 0).astype(np.uint8)
    L = _mo()
    L.eaoo_fuse_apply.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 9
    n = L.eaoo_fuse_apply(nT, nQ, _pp(_a(match_q, np.int32)), _pp(qnull), _pp(qid), _pp(slot), _pp(bad), _pp(obs), _pp(inkf),
                          _pp(repl), _pp(added))
    tag = lambda p: -1 if p < 0 else (int(p) if p < nQ else TAG + int(p) - nQ)
    replaced_by = np.array([-1 if qnull[i] else tag(repl[qid[i]]) for i in range(nQ)], np.int32)
    added_at = np.full(nQ, -1, np.int32)   # the AddObservation index is a property of the point: every alias reports it
    for i in range(nQ):
        if added[i] >= 0:
            added_at[(qid == qid[i]) & (qnull == 0)] = added[i]
    slot_replaced_by = np.array([tag(repl[nQ + k]) if slot_state[k] else -1 for k in range(nT)], np.int32)
    slot_holder = np.array([tag(p) for p in slot], np.int32)
    return n, added_at, replaced_by, slot_replaced_by, slot_holder


def r_fuse(KF, pts, th, *, bounds, grid_inv, scale_factors, inv_level_sigma2, log_scale_factor, mbf, L=None):
    """KF: dict(x,y,octave,desc,uright,slot_state,slot_obs); pts: dict(wx,..,desc,state,qid,obs).
    Returns (nFused, addedAt, replacedBy, slotReplacedBy, slotHolder)."""
    tx, ty, to, td, tr = _kf_arrays(KF)
    ss, so = _a(KF["slot_state"], np.uint8), _a(KF["slot_obs"], np.int32)
    wx, wy, wz, mx, mn, nrm, qd = _pt_arrays(pts)
    qs, qid, qo = _a(pts["state"], np.uint8), _a(pts["qid"], np.int32), _a(pts["obs"], np.int32)
    sf, inv = _a(scale_factors, np.float32), _a(inv_level_sigma2, np.float32)
    nT, nQ = len(tx), len(wx)
    added, repl = np.full(max(nQ, 1), -1, np.int32), np.full(max(nQ, 1), -1, np.int32)
    srep, shold = np.full(max(nT, 1), -1, np.int32), np.full(max(nT, 1), -1, np.int32)
    n = (L or match_ref_lib()).mref_fuse(
        nT, _pp(tx), _pp(ty), _pp(to), _pp(td), _pp(tr), _pp(ss), _pp(so), int(bounds[0]), int(bounds[1]), int(bounds[2]),
        int(bounds[3]), grid_inv[0], grid_inv[1], _pp(inv), _pp(sf), len(sf), float(log_scale_factor), float(mbf), nQ, _pp(qs),
        _pp(qid), _pp(wx), _pp(wy), _pp(wz), _pp(mx), _pp(mn), _pp(nrm), _pp(qo), _pp(qd), float(th), added.ctypes.data,
        repl.ctypes.data, srep.ctypes.data, shold.ctypes.data)
    return n, added[:nQ], repl[:nQ], srep[:nT], shold[:nT]


def o_fuse_sim3_apply(match_q, slot_state, slot_query):
    """Result step of Fuse(KeyFrame*, Scw, ...) (eaoo_fuse_sim3_apply).  Returns (nFused, addedAt, replacePoint, slotHolder)."""
    nQ, nT = len(match_q), len(slot_state)
    TAG = 100000
    slot = np.where(np.asarray(slot_query) >= 0, slot_query, np.where(np.asarray(slot_state) > 0, nQ + np.arange(nT), -1)).astype(np.int32)
    bad = np.zeros(nQ + nT, np.uint8)
    bad[nQ:] = np.asarray(slot_state) == 2
    repl, added = np.full(max(nQ, 1), -1, np.int32), np.full(max(nQ, 1), -1, np.int32)
    L = _mo()
    ci, vp = C.c_int, C.c_void_p
    L.eaoo_fuse_sim3_apply.argtypes = [ci, ci] + [vp] * 5
    n = L.eaoo_fuse_sim3_apply(nT, nQ, _pp(_a(match_q, np.int32)), _pp(slot), _pp(bad), _pp(repl), _pp(added))
    tag = lambda p: -1 if p < 0 else (int(p) if p < nQ else TAG + int(p) - nQ)
    return n, added[:nQ], np.array([tag(p) for p in repl[:nQ]], np.int32), np.array([tag(p) for p in slot], np.int32)


def r_fuse_sim3(KF, pts, th, *, bounds, grid_inv, scale_factors, log_scale_factor, scw_scale=1.0, L=None):
    """KF: dict(x,y,octave,desc,slot_state,slot_query); pts: dict(wx,..,desc,bad).
    Returns (nFused, addedAt, replacePoint, slotHolder)."""
    tx, ty, to, td, _ = _kf_arrays(KF)
    ss, sq = _a(KF["slot_state"], np.uint8), _a(KF["slot_query"], np.int32)
    wx, wy, wz, mx, mn, nrm, qd = _pt_arrays(pts)
    bad = _a(pts.get("bad"), np.uint8)
    sf = _a(scale_factors, np.float32)
    nT, nQ = len(tx), len(wx)
    added, repl = np.full(max(nQ, 1), -1, np.int32), np.full(max(nQ, 1), -1, np.int32)
    shold = np.full(max(nT, 1), -1, np.int32)
    n = (L or match_ref_lib()).mref_fuse_sim3(
        nT, _pp(tx), _pp(ty), _pp(to), _pp(td), _pp(ss), _pp(sq), int(bounds[0]), int(bounds[1]), int(bounds[2]), int(bounds[3]),
        grid_inv[0], grid_inv[1], _pp(sf), len(sf), float(log_scale_factor), float(scw_scale), nQ, _pp(bad), _pp(wx), _pp(wy),
        _pp(wz), _pp(mx), _pp(mn), _pp(nrm), _pp(qd), float(th), added.ctypes.data, repl.ctypes.data, shold.ctypes.data)
    return n, added[:nQ], repl[:nQ], shold[:nT]


def o_sim3_agreement(m1, m2):
    m1, m2 = _a(m1, np.int32), _a(m2, np.int32)
    out = np.full(max(len(m1), 1), -1, np.int32)
    L = _mo()
    L.eaoo_sim3_agreement.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    n = L.eaoo_sim3_agreement(len(m1), _pp(m1), len(m2), _pp(m2), out.ctypes.data)
    return n, out[:len(m1)]


def r_search_by_sim3(K1, K2, pre12, th, *, bounds, grid_inv, scale_factors, log_scale_factor, s12=1.0, L=None):
    """Ka: dict(x,y,octave,desc, state, wx,wy,wz,max_dist,min_dist, pdesc).  Returns (nFound, match12)."""
    def arrs(K):
        return (_a(K["x"], np.float32), _a(K["y"], np.float32), _a(K["octave"], np.int32), _a(K["desc"], np.uint8),
                _a(K["state"], np.uint8), _a(K["wx"], np.float32), _a(K["wy"], np.float32), _a(K["wz"], np.float32),
                _a(K["max_dist"], np.float32), _a(K["min_dist"], np.float32), _a(K["pdesc"], np.uint8))
    a1, a2 = arrs(K1), arrs(K2)
    pre = _a(pre12, np.int32)
    sf = _a(scale_factors, np.float32)
    n1, n2 = len(a1[0]), len(a2[0])
    m12 = np.full(max(n1, 1), -1, np.int32)
    n = (L or match_ref_lib()).mref_search_by_sim3(
        n1, *[_pp(a) for a in a1], n2, *[_pp(a) for a in a2], _pp(pre), int(bounds[0]), int(bounds[1]), int(bounds[2]),
        int(bounds[3]), grid_inv[0], grid_inv[1], _pp(sf), len(sf), float(log_scale_factor), float(s12), float(th), m12.ctypes.data)
    return n, m12[:n1]


def o_distinctive_descriptor(desc):
    d = _a(desc, np.uint8).reshape(-1, 32)
    L = _mo()
    L.eaoo_distinctive_descriptor.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    med = C.c_int(-1)
    return L.eaoo_distinctive_descriptor(len(d), _pp(d), C.byref(med)), med.value


MAPPOINT_REF_SO = os.path.join(HERE, "_ref", "libmappoint_ref.so")
_mpref = None


def mappoint_ref_lib():
    """The unmodified reference MapPoint.cc (+ ORBmatcher.cc) behind oracle/mappoint_ref_harness.cc."""
    global _mpref
    if _mpref is None:
        if not os.path.exists(MAPPOINT_REF_SO):
            subprocess.check_call(["make", "-s", "-C", HERE, "mappointref"])
        _mpref = C.CDLL(MAPPOINT_REF_SO)
        _mpref.mpref_distinctive_descriptor.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _mpref.mpref_predict_scale.argtypes = [C.c_float] * 4
    return _mpref


def r_distinctive_descriptor(desc, kf_bad=None):
    """MapPoint::ComputeDistinctiveDescriptors of the reference on n observations.  Returns (index of the chosen
    observation or -1, the 32 descriptor bytes the map point holds afterwards)."""
    d = _a(desc, np.uint8).reshape(-1, 32)
    bad = _a(kf_bad, np.uint8)
    out = np.zeros(32, np.uint8)
    i = mappoint_ref_lib().mpref_distinctive_descriptor(len(d), _pp(d), _pp(bad), out.ctypes.data)
    return i, out


# ---- frame ingest either side of the extractor ------------------------------------------------

def o_cvt_gray(img, color=0, mode=0):
    """cv::cvtColor(..., CV_BGR2GRAY | CV_RGB2GRAY | CV_BGRA2GRAY | CV_RGBA2GRAY) on 8U as Tracking calls it
    (src/Tracking.cc:324-337): OpenCV's RGB2Gray<uchar> integer formula.  color 0 BGR, 1 RGB, 2 BGRA, 3 RGBA; mode 0 =
    OpenCV 3.3.1 (coefficients 1868/9617/4899, shift 14), 1 = OpenCV 4.x (3735/19235/9798, shift 15; checked against cv2
    in tests/test_oracle_primitives.py)."""
    img = np.asarray(img, np.uint8)
    kb, kg, kr, sh = (1868, 9617, 4899, 14) if mode == 0 else (3735, 19235, 9798, 15)
    c0, c1, c2 = (img[..., i].astype(np.int64) for i in range(3))
    b, r = (c2, c0) if color in (1, 3) else (c0, c2)
    return ((b * kb + c1 * kg + r * kr + (1 << (sh - 1))) >> sh).astype(np.uint8)


def o_stereo_from_rgbd(kps, depth, mbf, depth_scale=1.0):
    """Frame::ComputeStereoFromRGBD (src/Frame.cc:1016-1037) with mvKeysUn = mvKeys: returns (mvuRight, mvDepth).
    depth float32, or uint16 converted like imDepth.convertTo(CV_32F, mDepthMapFactor) (src/Tracking.cc:340-341)."""
    if depth.dtype == np.uint16:
        depth = (depth.astype(np.float32) * np.float32(depth_scale)).astype(np.float32)
    n = len(kps)
    ur, dd = np.full(n, -1, np.float32), np.full(n, -1, np.float32)
    for i in range(n):
        u, v = int(kps["x"][i]), int(kps["y"][i])  # cv::Mat::at<float>(float v, float u): truncation
        d = depth[v, u]
        if d > 0:
            dd[i] = d
            ur[i] = np.float32(kps["x"][i]) - np.float32(mbf) / d
    return ur, dd


# ---- DBoW2 vocabulary: ORBVocabulary::transform ------------------------------------------------------
VOC_REF_SO = os.path.join(HERE, "_ref", "libvoc_ref.so")
_vref = None


def _setup_vref(L):
    vp, ci = C.c_void_p, C.c_int
    L.vref_load.restype = vp
    L.vref_load.argtypes = [C.c_char_p]
    L.vref_free.argtypes = [vp]
    L.vref_info.argtypes = [vp, vp]
    L.vref_tree.argtypes = [vp] * 7
    L.vref_transform.argtypes = [vp, ci, vp, ci] + [vp] * 6
    L.vref_score.restype = C.c_double
    L.vref_score.argtypes = [vp, ci, vp, vp, ci, vp, vp]
    return L


def voc_ref_lib():
    global _vref
    if _vref is None:
        if not os.path.exists(VOC_REF_SO):
            subprocess.check_call(["make", "-s", "-C", HERE, "vocref"])
        _vref = _setup_vref(C.CDLL(VOC_REF_SO))
    return _vref


def voc_harness_lib(path):
    """The same harness (oracle/voc_ref_harness.cc) built around the drop-in vocabulary class."""
    return _setup_vref(C.CDLL(path))


class RefVocabulary:
    """A vocabulary loaded by the reference's own TemplatedVocabulary::loadFromTextFile."""

    def __init__(self, path, L=None):
        self.L = L or voc_ref_lib()
        self.h = self.L.vref_load(path.encode())
        if not self.h:
            raise RuntimeError(f"vocabulary {path} did not load")
        info = np.zeros(6, np.int32)
        self.L.vref_info(self.h, info.ctypes.data)
        self.k, self.depth, self.n_nodes, self.n_words, self.scoring, self.weighting = (int(v) for v in info)

    def close(self):
        if self.h:
            self.L.vref_free(self.h)
            self.h = None

    def tree(self):
        """dict(parent, desc, weight, word_id, child_start, child_idx) as the reference holds the tree."""
        n = self.n_nodes
        t = dict(parent=np.zeros(n, np.int32), desc=np.zeros((n, 32), np.uint8), weight=np.zeros(n, np.float64),
                 word_id=np.zeros(n, np.int32), child_start=np.zeros(n + 1, np.int32), child_idx=np.zeros(max(n - 1, 1), np.int32))
        self.L.vref_tree(self.h, *[t[k].ctypes.data for k in ("parent", "desc", "weight", "word_id", "child_start", "child_idx")])
        t.update(L=self.depth, k=self.k, scoring=self.scoring, weighting=self.weighting)
        return t

    def transform(self, feats, levelsup=4):
        """Returns (word_ids, word_values, node_ids, node_starts, feat_idx)."""
        f = _a(feats, np.uint8).reshape(-1, 32)
        n = len(f)
        cnt = np.zeros(2, np.int32)
        wi, wv = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float64)
        ni, ns, fi = np.zeros(max(n, 1), np.uint32), np.zeros(n + 1, np.int32), np.zeros(max(n, 1), np.uint32)
        self.L.vref_transform(self.h, n, _pp(f), int(levelsup), cnt.ctypes.data, wi.ctypes.data, wv.ctypes.data, ni.ctypes.data,
                              ns.ctypes.data, fi.ctypes.data)
        return wi[:cnt[0]], wv[:cnt[0]], ni[:cnt[1]], ns[:cnt[1] + 1], fi[:ns[cnt[1]]]

    def score(self, a, b):
        ia, va, ib, vb = _a(a[0], np.uint32), _a(a[1], np.float64), _a(b[0], np.uint32), _a(b[1], np.float64)
        return self.L.vref_score(self.h, len(ia), _pp(ia), _pp(va), len(ib), _pp(ib), _pp(vb))


def o_voc_transform(tree, feats, levelsup=4):
    """oracle/voc_oracle.cc on the tree arrays of RefVocabulary.tree().  Same return layout as RefVocabulary.transform."""
    f = _a(feats, np.uint8).reshape(-1, 32)
    n = len(f)
    L = oracle_lib()
    ci, vp = C.c_int, C.c_void_p
    L.eaoo_voc_transform.restype = None
    L.eaoo_voc_transform.argtypes = [ci, ci, vp, vp, vp, vp, vp, ci, ci, ci, vp, ci] + [vp] * 6
    cnt = np.zeros(2, np.int32)
    wi, wv = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float64)
    ni, ns, fi = np.zeros(max(n, 1), np.uint32), np.zeros(n + 1, np.int32), np.zeros(max(n, 1), np.uint32)
    cs, cx, nd = _a(tree["child_start"], np.int32), _a(tree["child_idx"], np.int32), _a(tree["desc"], np.uint8)
    wt, wd = _a(tree["weight"], np.float64), _a(tree["word_id"], np.int32)
    L.eaoo_voc_transform(int(tree["L"]), len(wt), _pp(cs), _pp(cx), _pp(nd), _pp(wt), _pp(wd), int(tree["weighting"]),
                         int(tree["scoring"]), n, _pp(f),
                         int(levelsup), cnt.ctypes.data, wi.ctypes.data, wv.ctypes.data, ni.ctypes.data, ns.ctypes.data,
                         fi.ctypes.data)
    return wi[:cnt[0]], wv[:cnt[0]], ni[:cnt[1]], ns[:cnt[1] + 1], fi[:ns[cnt[1]]]


# ---- Frame::ComputeStereoMatches / ComputeStereoFromRGBD -------------------------------------
STEREO_REF_SO = os.path.join(HERE, "_ref", "libstereo_ref.so")
_sref = None


def stereo_ref_lib():
    """The text of Frame::ComputeStereoMatches / ComputeStereoFromRGBD cut out of the reference's src/Frame.cc at build
    time and compiled unmodified (oracle/stereo_ref_harness.cc)."""
    global _sref
    if _sref is None:
        if not os.path.exists(STEREO_REF_SO):
            subprocess.check_call(["make", "-s", "-C", HERE, "stereoref"])
        _sref = C.CDLL(STEREO_REF_SO)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        _sref.sref_stereo_matches.restype = None
        _sref.sref_stereo_matches.argtypes = [ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, cf, cf, vp, vp]
        _sref.sref_stereo_from_rgbd.restype = None
        _sref.sref_stereo_from_rgbd.argtypes = [ci, vp, vp, vp, vp, ci, ci, cf, vp, vp]
    return _sref


def pack_pyramid(levels):
    """Bordered level images [(h+38, w+38) u8] -> (buffer, off[], w[], h[]) as the stereo oracles take them."""
    off, ws, hs, o = [], [], [], 0
    for im in levels:
        off.append(o); hs.append(im.shape[0] - 38); ws.append(im.shape[1] - 38)
        o += im.size
    buf = np.concatenate([np.ascontiguousarray(im, np.uint8).ravel() for im in levels])
    return buf, np.asarray(off, np.int32), np.asarray(ws, np.int32), np.asarray(hs, np.int32)


def _stereo_args(kL, dL, kR, dR, pyrL, pyrR, scale, inv_scale):
    xL, yL, oL = _a(kL["x"], np.float32), _a(kL["y"], np.float32), _a(kL["octave"], np.int32)
    xR, yR, oR = _a(kR["x"], np.float32), _a(kR["y"], np.float32), _a(kR["octave"], np.int32)
    bL, off, w, h = pack_pyramid(pyrL)
    bR, off2, _, _ = pack_pyramid(pyrR)
    assert np.array_equal(off, off2)
    sf, isf = _a(scale, np.float32), _a(inv_scale, np.float32)
    keep = (xL, yL, oL, _a(dL, np.uint8), xR, yR, oR, _a(dR, np.uint8), sf, isf, bL, bR, off, w, h)
    return keep


def o_stereo_matches(kL, dL, kR, dR, pyrL, pyrR, scale, inv_scale, mb, mbf):
    """oracle/match_oracle.cc eaoo_stereo_matches.  kL/kR: keypoint records (x, y, octave); pyrL/pyrR: lists of bordered
    level images.  Returns (mvuRight, mvDepth, sad)."""
    xL, yL, oL, dL, xR, yR, oR, dR, sf, isf, bL, bR, off, w, h = _stereo_args(kL, dL, kR, dR, pyrL, pyrR, scale, inv_scale)
    n = len(xL)
    ur, dp, sad = np.full(max(n, 1), -1, np.float32), np.full(max(n, 1), -1, np.float32), np.full(max(n, 1), -1, np.int32)
    L = oracle_lib()
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    L.eaoo_stereo_matches.restype = None
    L.eaoo_stereo_matches.argtypes = [ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, cf, cf, vp, vp, vp]
    L.eaoo_stereo_matches(n, _pp(xL), _pp(yL), _pp(oL), _pp(dL), len(xR), _pp(xR), _pp(yR), _pp(oR), _pp(dR), len(sf), _pp(sf),
                          _pp(isf), _pp(bL), _pp(bR), _pp(off), _pp(w), _pp(h), float(mb), float(mbf), ur.ctypes.data,
                          dp.ctypes.data, sad.ctypes.data)
    return ur[:n], dp[:n], sad[:n]


def r_stereo_matches(kL, dL, kR, dR, pyrL, pyrR, scale, inv_scale, mb, mbf):
    xL, yL, oL, dL, xR, yR, oR, dR, sf, isf, bL, bR, off, w, h = _stereo_args(kL, dL, kR, dR, pyrL, pyrR, scale, inv_scale)
    n = len(xL)
    ur, dp = np.full(max(n, 1), -1, np.float32), np.full(max(n, 1), -1, np.float32)
    stereo_ref_lib().sref_stereo_matches(n, _pp(xL), _pp(yL), _pp(oL), _pp(dL), len(xR), _pp(xR), _pp(yR), _pp(oR), _pp(dR),
                                         len(sf), _pp(sf), _pp(isf), _pp(bL), _pp(bR), _pp(off), _pp(w), _pp(h), float(mb),
                                         float(mbf), ur.ctypes.data, dp.ctypes.data)
    return ur[:n], dp[:n]


def r_stereo_from_rgbd(kps, depth, mbf, x_un=None):
    x, y = _a(kps["x"], np.float32), _a(kps["y"], np.float32)
    xu = _a(x_un, np.float32)
    d = np.ascontiguousarray(depth, np.float32)
    n = len(x)
    ur, dp = np.full(max(n, 1), -1, np.float32), np.full(max(n, 1), -1, np.float32)
    stereo_ref_lib().sref_stereo_from_rgbd(n, _pp(x), _pp(y), _pp(xu), d.ctypes.data, d.shape[1], d.shape[0], float(mbf),
                                           ur.ctypes.data, dp.ctypes.data)
    return ur[:n], dp[:n]


def o_undistort_points(x, y, K, dist, guard=0):
    """Frame::UndistortKeyPoints' cv::undistortPoints (src/Frame.cc:773-803).  K = (fx, fy, cx, cy); dist = k1 k2 p1 p2 [k3 ...].
    guard=1: OpenCV 4.x semantics (checked against cv2), guard=0: OpenCV 3.3.1."""
    x, y, d = _a(x, np.float32), _a(y, np.float32), _a(dist, np.float32)
    n = len(x)
    xo, yo = np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.float32)
    L = oracle_lib()
    cf, ci, vp = C.c_float, C.c_int, C.c_void_p
    L.eaoo_undistort_points.restype = None
    L.eaoo_undistort_points.argtypes = [ci, vp, vp, cf, cf, cf, cf, vp, ci, ci, vp, vp]
    L.eaoo_undistort_points(n, _pp(x), _pp(y), float(K[0]), float(K[1]), float(K[2]), float(K[3]), _pp(d), len(d), int(guard),
                            xo.ctypes.data, yo.ctypes.data)
    return xo[:n], yo[:n]
