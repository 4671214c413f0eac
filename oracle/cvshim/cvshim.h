// ORACLE — test infrastructure only (never linked into the product library).
//
// Minimal stand-in for the slice of the OpenCV C++ API that
// /root/reference/src/ORBextractor.cc and include/ORBextractor.h use, so that
// the reference translation unit compiles UNMODIFIED, in place, in a container
// that has no OpenCV C++ headers.  Only the types and the five
// image primitives the reference calls exist here; the arithmetic lives in
// ../cv_primitives.h.  The same shim is used to compile the drop-in
// ORB_SLAM2::ORBextractor wrapper (eao-fusion_b200/dropin) for interface tests.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>  // OpenCV's core headers pull <cmath> in; overload resolution of cos/sin depends on it
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../cv_primitives.h"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t)&7)

static inline int cvRound(double v) { return cvprim::round_d(v); }
static inline int cvRound(float v) { return cvprim::round_f(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return cvprim::floor_d(v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { return cvprim::ceil_d(v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

int eaof_shim_blur_mode();        // set by the harness: cvprim::BlurMode
void eaof_shim_set_blur_mode(int);

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, float b) {
    a.x = (T)(a.x * b);
    a.y = (T)(a.y * b);
    return a;
}

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
             int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

// Only referenced by the reference's dead ComputeKeyPointsOld (src/ORBextractor.cc:1006,1024, never called).
struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        if (n < 0 || (size_t)n >= k.size()) return;
        std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        k.resize(n);
    }
};

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

struct MatStep {
    size_t v;
    MatStep(size_t s = 0) : v(s) {}
    operator size_t() const { return v; }
};

struct MatExpr {  // only Mat::zeros is needed
    int rows, cols, type;
};

class Mat {
public:
    int rows, cols;
    uchar* data;
    MatStep step;

    Mat() : rows(0), cols(0), data(nullptr), step(0) {}
    Mat(Size sz, int type) : rows(0), cols(0), data(nullptr), step(0) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), step(0) { create(r, c, type); }
    // external (non-owning) 8-bit buffer
    Mat(int r, int c, int /*type*/, void* ext, size_t stp) : rows(r), cols(c), data((uchar*)ext), step(stp) {}
    Mat(const MatExpr& e) : rows(0), cols(0), data(nullptr), step(0) { *this = e; }

    void create(int r, int c, int /*type*/) {
        if (data && r == rows && c == cols) return;  // OpenCV keeps the buffer when shape/type agree
        buf_.reset(new uchar[(size_t)r * c + 1], std::default_delete<uchar[]>());
        data = buf_.get();
        rows = r;
        cols = c;
        step = (size_t)c;
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    Mat& operator=(const MatExpr& e) {  // Mat::zeros assignment writes into an existing same-shaped buffer
        create(e.rows, e.cols, e.type);
        for (int y = 0; y < rows; ++y) memset(data + (size_t)y * step, 0, cols);
        return *this;
    }
    static MatExpr zeros(int r, int c, int type) { return MatExpr{r, c, type}; }

    int type() const { return CV_8UC1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t step1() const { return step; }
    bool isContinuous() const { return (size_t)cols == (size_t)step; }

    Mat operator()(const Rect& r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step + r.x;
        m.rows = r.height;
        m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; ++y) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
        return m;
    }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const {
        return *(const T*)(data + (size_t)y * step + x * sizeof(T));
    }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }

private:
    std::shared_ptr<uchar> buf_;
};

class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }

protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) { m_ = &m; }
    void create(int r, int c, int type) const { if (m_) m_->create(r, c, type); }
    void create(Size sz, int type) const { if (m_) m_->create(sz, type); }
    void release() const { if (m_) m_->release(); }
    bool needed() const { return m_ != nullptr; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
static inline InputArray noArray() { static _InputArray a; return a; }

static inline float fastAtan2(float y, float x) { return cvprim::fast_atan2(y, x); }

static inline void resize(InputArray _src, OutputArray _dst, Size dsize, double /*fx*/ = 0, double /*fy*/ = 0,
                          int interpolation = INTER_LINEAR) {
    assert(interpolation == INTER_LINEAR);
    (void)interpolation;
    Mat src = _src.getMat();
    _dst.create(dsize, src.type());
    Mat dst = _dst.getMat();
    cvprim::resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

// BORDER_ISOLATED is honoured trivially: the shim never looks outside the ROI.  (The reference's only
// non-isolated call, src/ORBextractor.cc:1127, passes the caller's whole image.)
static inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right,
                                  int borderType) {
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    (void)borderType;
    Mat src = _src.getMat();
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    cvprim::copy_make_border_reflect101(src.data, src.cols, src.rows, src.step, dst.data, dst.step, top, bottom,
                                        left, right);
}

static inline void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
    assert(nonmaxSuppression);
    (void)nonmaxSuppression;
    Mat img = _img.getMat();
    std::vector<cvprim::FastPt> pts;
    cvprim::fast9_nms(img.data, img.cols, img.rows, img.step, threshold, pts);
    keypoints.clear();
    keypoints.reserve(pts.size());
    for (const auto& p : pts) keypoints.push_back(KeyPoint((float)p.x, (float)p.y, 7.f, -1, (float)p.score));
}

static inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY = 0,
                                int borderType = BORDER_REFLECT_101) {
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sigmaX; (void)sigmaY; (void)borderType;
    Mat src = _src.getMat().clone();  // in-place calls are legal
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    cvprim::gaussian_blur7(src.data, src.cols, src.rows, src.step, dst.data, dst.step, eaof_shim_blur_mode());
}

}  // namespace cv
