// ORACLE — test infrastructure only.
//
// Restatement of glibc 2.39's sinf/cosf (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h — the
// ARM "optimized routines" double-precision polynomial) for |x| < 120.
//
// Why it matters: the reference computes `(float)cos(angle)` / `(float)sin(angle)` with a FLOAT argument
// under `using namespace std` (src/ORBextractor.cc:64,113), so overload resolution picks std::cos(float)
// = cosf — NOT double cos — and g++ emits one sincosf call (checked: `nm -D oracle/_ref/liborb_ref.so`).
// cosf's polynomial has ~2^-28 relative error, so it differs from a correctly-rounded cosine on a few
// percent of inputs; bit-exact descriptors need this exact function.  The constants below were read out of
// this image's libm.so.6 (__sincosf_table) and the restatement was checked against libm's sinf/cosf for ALL
// 1,087,163,597 floats in [0, 6.4]: 0 mismatches, with and without FMA contraction
// (tests/test_oracle_primitives.py re-runs a strided version of that sweep).
#pragma once
#include <cstdint>
#include <cstring>

namespace glibcf {

static const double kHpiInv = 0x1.45F306DC9C883p+23;  // 2/pi * 2^24
static const double kHpi = 0x1.921FB54442D18p0;       // pi/2
static const double kC0 = 0x1p0, kC1 = -0x1.ffffffd0c621cp-2, kC2 = 0x1.55553e1068f19p-5,
                    kC3 = -0x1.6c087e89a359dp-10, kC4 = 0x1.99343027bf8c3p-16;
static const double kS1 = -0x1.555545995a603p-3, kS2 = 0x1.1107605230bc4p-7, kS3 = -0x1.994eb3774cf24p-13;

static inline uint32_t abstop12(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u >> 20) & 0x7ff;
}

// n even -> sine polynomial, n odd -> cosine polynomial; neg selects the negated cosine table.
static inline float poly(double x, double x2, bool neg, int n) {
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double s1 = kS2 + x2 * kS3;
        const double x7 = x3 * x2;
        const double s = x + x3 * kS1;
        return (float)(s + x7 * s1);
    }
    const double sg = neg ? -1.0 : 1.0;
    const double x4 = x2 * x2;
    const double c2 = sg * kC3 + x2 * (sg * kC4);
    const double c1 = sg * kC0 + x2 * (sg * kC1);
    const double x6 = x4 * x2;
    const double c = c1 + x4 * (sg * kC2);
    return (float)(c + x6 * c2);
}

// valid for 0 <= |y| < 120 (the reference only passes [0, 2*pi])
static inline void sincosf_restated(float y, float* sp, float* cp) {
    double x = y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {  // |y| < pi/4
        const double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) {
            *sp = y;
            *cp = 1.0f;
            return;
        }
        *sp = poly(x, x2, false, 0);
        *cp = poly(x, x2, false, 1);
        return;
    }
    const double r = x * kHpiInv;
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = x - n * kHpi;
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const bool neg = (n & 2) != 0;
    *sp = poly(x * s, x * x, neg, n);
    *cp = poly(x * s, x * x, neg, n ^ 1);
}

}  // namespace glibcf
