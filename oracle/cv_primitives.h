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

// cv::FAST(img, kps, threshold, nonmaxSuppression=true, TYPE_9_16): raster order output.
static inline void fast9_nms(const uint8_t* img, int w, int h, size_t step, int threshold,
                             std::vector<FastPt>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    std::vector<int> score((size_t)w * h, 0);
    std::vector<uint8_t> corner((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            // OpenCV's early-out: every 9-arc contains one pixel of each opposite pair (k, k+8)
            const uint8_t* p = img + (size_t)y * step + x;
            const int v = p[0], lo = v - threshold, hi = v + threshold;
            int d = 3;
            for (int k = 0; k < 8 && d; ++k) {
                const int a = p[(ptrdiff_t)kRingDy[k] * (ptrdiff_t)step + kRingDx[k]];
                const int b = p[(ptrdiff_t)kRingDy[k + 8] * (ptrdiff_t)step + kRingDx[k + 8]];
                d &= ((a < lo) | ((a > hi) << 1)) | ((b < lo) | ((b > hi) << 1));
            }
            if (!d) continue;
            const int b = fast_arc_best(p, step);
            if (b > threshold) {
                score[(size_t)y * w + x] = b - 1;
                corner[(size_t)y * w + x] = 1;
            }
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            if (!corner[(size_t)y * w + x]) continue;
            const int* r = &score[(size_t)y * w + x];
            const int s = r[0];
            if (s > r[-1] && s > r[1] && s > r[-w - 1] && s > r[-w] && s > r[-w + 1] && s > r[w - 1] &&
                s > r[w] && s > r[w + 1])
                out.push_back({x, y, s});
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
    std::vector<int> rows((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * sstep;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            for (int i = 0; i < 7; ++i) acc += k[i] * s[reflect101(x + i - 3, w)];
            rows[(size_t)y * w + x] = acc;
        }
    }
    const int simd_w = (mode == BLUR_CV331_SSE2) ? (w & ~3) : 0;
    for (int y = 0; y < h; ++y) {
        uint8_t* d = dst + (size_t)y * dstep;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            for (int j = 0; j < 7; ++j) acc += k[j] * rows[(size_t)reflect101(y + j - 3, h) * w + x];
            int q;
            if (x < simd_w) {
                q = acc >> 16;
                const int rem = acc & 0xFFFF;
                q += (rem > 32768) || (rem == 32768 && (q & 1));
            } else {
                q = (acc + 32768) >> 16;
            }
            d[x] = sat_u8(q);
        }
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
