// ORACLE — test infrastructure only. Nothing under oracle/ is on the product path.
//
// Plain-array restatements of the five OpenCV primitives the reference's ORB
// front-end calls (OpenCV is NOT vendored in /root/reference; README.md:44 pins
// 3.3.1 in prose).  Call sites in the reference:
//   resize          src/ORBextractor.cc:1120
//   copyMakeBorder  src/ORBextractor.cc:1122-1123,1127-1128
//   FAST            src/ORBextractor.cc:809-810,814-815
//   GaussianBlur    src/ORBextractor.cc:1086
//   fastAtan2       src/ORBextractor.cc:103
//   cvRound         src/ORBextractor.cc:81,115,119-120,442,460,1112
// The algorithms restated here are OpenCV's published ones (imgproc/resize.cpp
// fixed-point INTER_LINEAR, core/copy.cpp borderInterpolate, features2d/fast.cpp
// FAST_t<16> + cornerScore<16>, imgproc/filter.cpp 8U symmetric separable
// filter, core/mathfuncs_core atan polynomial).  tests/test_oracle_primitives.py
// pins resize / copyMakeBorder / FAST / fastAtan2 and the CV4 blur taps against
// the cv2 4.13 wheel in this image.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace cvprim {

// cvRound: round-half-to-even (SSE cvtss2si / cvtsd2si under the default MXCSR).
static inline int round_f(float v) { return (int)lrintf(v); }
static inline int round_d(double v) { return (int)lrint(v); }
static inline int floor_d(double v) { int i = (int)v; return i - (i > v); }
static inline int ceil_d(double v) { int i = (int)v; return i + (i < v); }
static inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }
static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

static inline int reflect101(int p, int len) {
    // borderInterpolate(BORDER_REFLECT_101); len==1 -> 0
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

// dst is (sw+l+r) x (sh+t+b); dst may alias a buffer that already holds src at (l,t).
static inline void copy_make_border_reflect101(const uint8_t* src, int sw, int sh, size_t sstep,
                                               uint8_t* dst, size_t dstep, int top, int bottom,
                                               int left, int right) {
    const int dw = sw + left + right;
    std::vector<int> xmap(dw);
    for (int x = 0; x < dw; ++x) xmap[x] = reflect101(x - left, sw);
    std::vector<uint8_t> row(dw);
    // inner rows first (safe for the in-place case because each row is staged)
    for (int y = 0; y < sh; ++y) {
        const uint8_t* s = src + (size_t)y * sstep;
        for (int x = 0; x < dw; ++x) row[x] = s[xmap[x]];
        memcpy(dst + (size_t)(y + top) * dstep, row.data(), dw);
    }
    for (int y = 0; y < top; ++y)
        memcpy(dst + (size_t)y * dstep, dst + (size_t)(reflect101(y - top, sh) + top) * dstep, dw);
    for (int y = 0; y < bottom; ++y)
        memcpy(dst + (size_t)(top + sh + y) * dstep,
               dst + (size_t)(reflect101(sh + y, sh) + top) * dstep, dw);
}

// 8UC1 INTER_LINEAR, 11-bit fixed-point coefficients (INTER_RESIZE_COEF_BITS).
static inline void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst,
                                    int dw, int dh, size_t dstep) {
    const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ia(2 * dw), ib(2 * dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = floor_d(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ia[2 * dx] = sat_short(round_f((1.f - fx) * 2048.f));
        ia[2 * dx + 1] = sat_short(round_f(fx * 2048.f));
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = floor_d(fy);
        fy -= sy;
        yofs[dy] = sy;
        ib[2 * dy] = sat_short(round_f((1.f - fy) * 2048.f));
        ib[2 * dy + 1] = sat_short(round_f(fy * 2048.f));
    }
    std::vector<int> r0(dw), r1(dw);
    for (int dy = 0; dy < dh; ++dy) {
        int sy0 = yofs[dy], sy1 = yofs[dy] + 1;
        sy0 = sy0 < 0 ? 0 : sy0 > sh - 1 ? sh - 1 : sy0;
        sy1 = sy1 < 0 ? 0 : sy1 > sh - 1 ? sh - 1 : sy1;
        const uint8_t* s0 = src + (size_t)sy0 * sstep;
        const uint8_t* s1 = src + (size_t)sy1 * sstep;
        for (int dx = 0; dx < dw; ++dx) {
            const int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            const int a0 = ia[2 * dx], a1 = ia[2 * dx + 1];
            r0[dx] = s0[sx] * a0 + s0[sx1] * a1;
            r1[dx] = s1[sx] * a0 + s1[sx1] * a1;
        }
        const int b0 = ib[2 * dy], b1 = ib[2 * dy + 1];
        uint8_t* d = dst + (size_t)dy * dstep;
        for (int dx = 0; dx < dw; ++dx)
            d[dx] = (uint8_t)((((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---- FAST-9/16 ------------------------------------------------------------------------------
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// Arc score: max over the 16 contiguous 9-arcs of min(ring-v) (bright) / min(v-ring) (dark).
// A pixel is a FAST-9 corner at threshold t iff arc_best > t, and OpenCV's cornerScore<16>
// returns arc_best-1 for such a pixel.
static inline int fast_arc_best(const uint8_t* p, size_t step) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = (int)p[(ptrdiff_t)kRingDy[k] * (ptrdiff_t)step + kRingDx[k]] - v;
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) {
            mn = d[k + j] < mn ? d[k + j] : mn;
            mx = d[k + j] > mx ? d[k + j] : mx;
        }
        if (mn > best) best = mn;    // brighter arc
        if (-mx > best) best = -mx;  // darker arc
    }
    return best;
}

struct FastPt { int x, y, score; };

// The same arc score with sliding minima / maxima (2-, 4-, 8-windows, then 9 = 8 + 1): ~100 operations instead of the
// 16 x 9 scan above; fast_arc_best stays as the plain statement the tests compare against.
static inline int fast_arc_best_quick(const uint8_t* p, size_t step) {
    int d[16], mn2[16], mx2[16], mn4[16], mx4[16];
    const int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = (int)p[(ptrdiff_t)kRingDy[k] * (ptrdiff_t)step + kRingDx[k]] - v;
    for (int k = 0; k < 16; ++k) {
        const int e = d[(k + 1) & 15];
        mn2[k] = d[k] < e ? d[k] : e;
        mx2[k] = d[k] > e ? d[k] : e;
    }
    for (int k = 0; k < 16; ++k) {
        const int a = mn2[(k + 2) & 15], c = mx2[(k + 2) & 15];
        mn4[k] = mn2[k] < a ? mn2[k] : a;
        mx4[k] = mx2[k] > c ? mx2[k] : c;
    }
    int best = -256;
    for (int k = 0; k < 16; ++k) {
        int mn = mn4[k] < mn4[(k + 4) & 15] ? mn4[k] : mn4[(k + 4) & 15];
        int mx = mx4[k] > mx4[(k + 4) & 15] ? mx4[k] : mx4[(k + 4) & 15];
        const int e = d[(k + 8) & 15];
        mn = mn < e ? mn : e;
        mx = mx > e ? mx : e;
        if (mn > best) best = mn;
        if (-mx > best) best = -mx;
    }
    return best;
}

}  // namespace cvprim
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
namespace cvprim {

// cv::FAST(img, kps, threshold, nonmaxSuppression=true, TYPE_9_16): raster order output.
// Organised like OpenCV's FAST_t<16>: a three-row ring of score rows, the early-out on 16 pixels at a time (SSE2:
// every 9-arc contains one pixel of each opposite pair (k, k+8), so a corner is darker-or-brighter than the threshold
// band on one pixel of all 8 pairs), the exact arc score only for the survivors, NMS one row behind.
static inline void fast9_nms(const uint8_t* img, int w, int h, size_t step, int threshold,
                             std::vector<FastPt>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    // per-call buffers (no storage may outlive the call: the reference harness recycles its allocation arena per frame);
    // cells are at most 66 px wide, so the common case lives on the stack
    const int rw = w + 2;
    int stackBuf[3 * 70 + 3 * 69];
    std::vector<int> heapBuf;
    int* sbuf = stackBuf;
    if (w > 68) { heapBuf.resize((size_t)3 * rw + (size_t)3 * (w + 1)); sbuf = heapBuf.data(); }
    int* cbuf = sbuf + (size_t)3 * rw;  // corner columns of rows y-2, y-1, y; [0] = count
    memset(sbuf, 0, sizeof(int) * (size_t)3 * rw);  // scores of those rows (column x at index x+1), 0 = not a corner
    ptrdiff_t off[16];
    for (int k = 0; k < 16; ++k) off[k] = (ptrdiff_t)kRingDy[k] * (ptrdiff_t)step + kRingDx[k];
    const int t = threshold < 0 ? 0 : threshold > 255 ? 255 : threshold;
    for (int y = 3; y < h - 2; ++y) {
        int* cur = &sbuf[(size_t)((y - 3) % 3) * rw];
        int* ccur = &cbuf[(size_t)((y - 3) % 3) * (w + 1)];
        memset(cur, 0, sizeof(int) * rw);
        int nc = 0;
        if (y < h - 3) {
            const uint8_t* row = img + (size_t)y * step;
            int x = 3;
#if defined(__SSE2__)
            const __m128i vt = _mm_set1_epi8((char)t), zero = _mm_setzero_si128();
            for (; x + 16 <= w - 3 || (x < w - 3 && w - 3 - 16 >= 3); ) {
                int x0 = x;
                if (x0 + 16 > w - 3) x0 = w - 3 - 16;  // last block overlaps the previous one
                const uint8_t* p = row + x0;
                const __m128i v = _mm_loadu_si128((const __m128i*)p);
                const __m128i lo = _mm_subs_epu8(v, vt), hi = _mm_adds_epu8(v, vt);
                __m128i dark = _mm_set1_epi8((char)0xff), bright = dark;
                for (int k = 0; k < 8; ++k) {
                    const __m128i a = _mm_loadu_si128((const __m128i*)(p + off[k]));
                    const __m128i b2 = _mm_loadu_si128((const __m128i*)(p + off[k + 8]));
                    // a < lo  <=>  subs(lo, a) != 0 ;  a > hi  <=>  subs(a, hi) != 0
                    const __m128i da = _mm_or_si128(_mm_subs_epu8(lo, a), _mm_subs_epu8(lo, b2));
                    const __m128i ba = _mm_or_si128(_mm_subs_epu8(a, hi), _mm_subs_epu8(b2, hi));
                    dark = _mm_andnot_si128(_mm_cmpeq_epi8(da, zero), dark);
                    bright = _mm_andnot_si128(_mm_cmpeq_epi8(ba, zero), bright);
                    if (k == 1 || k == 3) {
                        if (_mm_movemask_epi8(_mm_or_si128(dark, bright)) == 0) break;
                    }
                }
                unsigned m = (unsigned)_mm_movemask_epi8(_mm_or_si128(dark, bright));
                if (x0 < x) m &= ~0u << (x - x0);  // pixels the previous block already handled
                while (m) {
                    const int j = __builtin_ctz(m);
                    m &= m - 1;
                    const int b = fast_arc_best_quick(p + j, step);
                    if (b > t) { cur[x0 + j + 1] = b - 1; ccur[1 + nc++] = x0 + j; }
                }
                x = x0 + 16;
                if (x >= w - 3) break;
            }
#endif
            for (; x < w - 3; ++x) {
                const uint8_t* p = row + x;
                const int v = p[0], lo = v - t, hi = v + t;
                int d = 3;
                for (int k = 0; k < 8 && d; ++k) {
                    const int a = p[off[k]], b2 = p[off[k + 8]];
                    d &= ((a < lo) | ((a > hi) << 1)) | ((b2 < lo) | ((b2 > hi) << 1));
                }
                if (!d) continue;
                const int b = fast_arc_best_quick(p, step);
                if (b > t) { cur[x + 1] = b - 1; ccur[1 + nc++] = x; }
            }
        }
        ccur[0] = nc;
        if (y == 3) continue;
        // NMS of row y-1 against rows y-2, y-1, y (rows outside [3, h-3) hold zeros)
        const int* prev = &sbuf[(size_t)((y - 4 + 3) % 3) * rw];
        const int* pprev = &sbuf[(size_t)((y - 5 + 6) % 3) * rw];
        const int* cprev = &cbuf[(size_t)((y - 4 + 3) % 3) * (w + 1)];
        const bool havePP = y - 2 >= 3;
        for (int i = 0; i < cprev[0]; ++i) {
            const int x = cprev[1 + i], xi = x + 1;
            const int sc = prev[xi];
            if (sc > prev[xi - 1] && sc > prev[xi + 1] && sc > cur[xi - 1] && sc > cur[xi] && sc > cur[xi + 1] &&
                (!havePP || (sc > pprev[xi - 1] && sc > pprev[xi] && sc > pprev[xi + 1])))
                out.push_back({x, y - 1, sc});
        }
    }
}

// ---- Gaussian blur 7x7, sigma 2, 8U, reflect-101 -------------------------------------------
enum BlurMode {
    BLUR_CV331 = 0,       // OpenCV 3.3.1 8U path: taps rint(256*g) = 18,34,49,55,.. (sum 257), (S+2^15)>>16
    BLUR_CV4 = 1,         // OpenCV >=3.4.1 / 4.x bit-exact path: taps 18,34,48,56,.. (sum 256), (S+2^15)>>16
    BLUR_CV331_SSE2 = 2,  // 3.3.1 taps; x < 4*floor(w/4) rounded half-to-even (SymmColumnVec_32s8u float
                          // path, cvtps2dq), scalar tail rounded half-up
};

static inline const int* blur_taps(int mode) {
    static const int t331[7] = {18, 34, 49, 55, 49, 34, 18};
    static const int t4[7] = {18, 34, 48, 56, 48, 34, 18};
    return mode == BLUR_CV4 ? t4 : t331;
}

static inline void gaussian_blur7(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst,
                                  size_t dstep, int mode) {
    const int* k = blur_taps(mode);
    // horizontal pass on a reflect-padded copy of the row (plain loops the compiler vectorises); sums fit 16 bits
    std::vector<uint16_t> rows((size_t)w * h);
    std::vector<uint8_t> pad((size_t)w + 6);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * sstep;
        for (int i = 0; i < 3; ++i) { pad[i] = s[reflect101(i - 3, w)]; pad[w + 3 + i] = s[reflect101(w + i, w)]; }
        memcpy(pad.data() + 3, s, w);
        const uint8_t* q = pad.data();
        uint16_t* r = &rows[(size_t)y * w];
        const int k0 = k[0], k1 = k[1], k2 = k[2], k3 = k[3];
        for (int x = 0; x < w; ++x)
            r[x] = (uint16_t)(k0 * (q[x] + q[x + 6]) + k1 * (q[x + 1] + q[x + 5]) + k2 * (q[x + 2] + q[x + 4]) + k3 * q[x + 3]);
    }
    const int simd_w = (mode == BLUR_CV331_SSE2) ? (w & ~3) : 0;
    std::vector<int> acc(w);
    for (int y = 0; y < h; ++y) {
        const uint16_t* r[7];
        for (int j = 0; j < 7; ++j) r[j] = &rows[(size_t)reflect101(y + j - 3, h) * w];
        const int k0 = k[0], k1 = k[1], k2 = k[2], k3 = k[3];
        for (int x = 0; x < w; ++x)
            acc[x] = k0 * ((int)r[0][x] + r[6][x]) + k1 * ((int)r[1][x] + r[5][x]) + k2 * ((int)r[2][x] + r[4][x]) + k3 * (int)r[3][x];
        uint8_t* d = dst + (size_t)y * dstep;
        for (int x = 0; x < simd_w; ++x) {
            int q = acc[x] >> 16;
            const int rem = acc[x] & 0xFFFF;
            q += (rem > 32768) || (rem == 32768 && (q & 1));
            d[x] = sat_u8(q);
        }
        for (int x = simd_w; x < w; ++x) d[x] = sat_u8((acc[x] + 32768) >> 16);
    }
}

// ---- fastAtan2 (degrees, [0,360)) -----------------------------------------------------------
static inline float fast_atan2(float y, float x) {
    static const float p1 = 0.9997878412794807f * (float)(180 / 3.1415926535897932384626433832795);
    static const float p3 = -0.3258083974640975f * (float)(180 / 3.1415926535897932384626433832795);
    static const float p5 = 0.1555786518463281f * (float)(180 / 3.1415926535897932384626433832795);
    static const float p7 = -0.04432655554792128f * (float)(180 / 3.1415926535897932384626433832795);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

}  // namespace cvprim
