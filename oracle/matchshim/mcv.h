// ORACLE — test infrastructure only (never linked into the product library).
//
// Stand-in for the slice of OpenCV that /root/reference/src/ORBmatcher.cc uses, so that the reference translation
// unit compiles UNMODIFIED, in place (see ../Makefile, target `matchref`).  Unlike ../cvshim (8-bit images for the
// extractor) this cv::Mat carries a depth: CV_8U descriptor rows and small CV_32F matrices (poses, points, F12).
// Float products accumulate in double and round once, like cv::gemm's GEMMSingleMul<float,double>; the parity tests
// use identity rotations / zero translations on the paths they pin, so no result depends on that choice.
#pragma once
#include <algorithm>  // <opencv2/core/core.hpp> brings these in; src/MapPoint.cc relies on it (sort, INT_MAX)
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 0

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;

    Mat() : rows(0), cols(0), data(nullptr), step(0), depth_(CV_8U) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), step(0), depth_(type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t stp = 0)
        : rows(r), cols(c), data((uchar*)ext), step(stp ? stp : (size_t)c * esz(type)), depth_(type) {}

    static size_t esz(int type) { return type == CV_32F ? 4 : 1; }
    size_t elemSize() const { return esz(depth_); }
    int type() const { return depth_; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }

    void create(int r, int c, int type) {
        depth_ = type;
        buf_.reset(new uchar[(size_t)r * c * esz(type) + 4](), std::default_delete<uchar[]>());
        data = buf_.get();
        rows = r;
        cols = c;
        step = (size_t)c * esz(type);
    }
    Mat sub(int y, int x, int h, int w) const {
        Mat m(*this);
        m.data = data + (size_t)y * step + (size_t)x * elemSize();
        m.rows = h;
        m.cols = w;
        return m;
    }
    Mat rowRange(int a, int b) const { return sub(a, 0, b - a, cols); }
    Mat colRange(int a, int b) const { return sub(0, a, rows, b - a); }
    Mat row(int y) const { return sub(y, 0, 1, cols); }
    Mat col(int x) const { return sub(0, x, rows, 1); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, depth_);
        for (int y = 0; y < rows; ++y) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
        return m;
    }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    // single index: element i of a row or column vector
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }

    void copyTo(Mat& dst) const { dst = clone(); }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat ones(int r, int c, int type) {
        Mat m(r, c, type);
        for (int y = 0; y < r; ++y)
            for (int x = 0; x < c; ++x) {
                if (type == CV_32F) m.at<float>(y, x) = 1.f; else m.at<uchar>(y, x) = 1;
            }
        return m;
    }
    // 8U -> 32F / 32F -> 32F without scaling (Frame::ComputeStereoMatches converts its 11x11 patches, src/Frame.cc:939,957)
    void convertTo(Mat& dst, int type) const {
        assert(type == CV_32F);
        Mat m(rows, cols, CV_32F);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) m.at<float>(y, x) = depth_ == CV_32F ? at<float>(y, x) : (float)at<uchar>(y, x);
        dst = m;
    }
    Mat t() const {
        assert(depth_ == CV_32F);
        Mat m(cols, rows, CV_32F);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) m.at<float>(x, y) = at<float>(y, x);
        return m;
    }
    double dot(const Mat& o) const {
        assert(depth_ == CV_32F && rows * cols == o.rows * o.cols);
        double s = 0;
        const int n = rows * cols;
        for (int i = 0; i < n; ++i) s += (double)lin(i) * (double)o.lin(i);
        return s;
    }
    float lin(int i) const { return at<float>(i / cols, i % cols); }

private:
    int depth_;
    std::shared_ptr<uchar> buf_;
};

// cv::FileStorage / cv::FileNode: only named by TemplatedVocabulary::save/load(cv::FileStorage&) (Thirdparty/DBoW2/DBoW2/
// TemplatedVocabulary.h:1534-1740), which the harnesses never call — the vocabulary is read with loadFromTextFile.  The
// stand-ins make those members compile; reaching them aborts.
struct FileNode {
    FileNode operator[](const char*) const { std::abort(); }
    FileNode operator[](const std::string&) const { std::abort(); }
    FileNode operator[](int) const { std::abort(); }
    size_t size() const { std::abort(); }
    operator int() const { std::abort(); }
    operator double() const { std::abort(); }
    operator std::string() const { std::abort(); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const std::string&) const { std::abort(); }
    FileNode operator[](const char*) const { std::abort(); }
};
template <typename T> static inline FileStorage& operator<<(FileStorage& f, const T&) { std::abort(); return f; }

static inline Mat operator*(const Mat& a, const Mat& b) {
    assert(a.cols == b.rows);
    Mat m(a.rows, b.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < b.cols; ++x) {
            double s = 0;
            for (int k = 0; k < a.cols; ++k) s += (double)a.at<float>(y, k) * (double)b.at<float>(k, x);
            m.at<float>(y, x) = (float)s;
        }
    return m;
}
static inline Mat scaled(const Mat& a, double s) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) m.at<float>(y, x) = (float)((double)a.at<float>(y, x) * s);
    return m;
}
static inline Mat operator*(double s, const Mat& a) { return scaled(a, s); }
static inline Mat operator*(const Mat& a, double s) { return scaled(a, s); }
static inline Mat operator/(const Mat& a, double s) { return scaled(a, 1.0 / s); }
static inline Mat operator-(const Mat& a) { return scaled(a, -1.0); }
static inline Mat addw(const Mat& a, const Mat& b, float sb) {
    assert(a.rows == b.rows && a.cols == b.cols);
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) m.at<float>(y, x) = a.at<float>(y, x) + sb * b.at<float>(y, x);
    return m;
}
static inline Mat operator+(const Mat& a, const Mat& b) { return addw(a, b, 1.f); }
static inline Mat operator-(const Mat& a, const Mat& b) { return addw(a, b, -1.f); }
static inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }
enum { NORM_L1 = 2 };
static inline double norm(const Mat& a, const Mat& b, int normType) {  // normDiffL1_32f: double accumulator
    assert(normType == NORM_L1 && a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    double s = 0;
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) s += std::fabs((double)a.at<float>(y, x) - (double)b.at<float>(y, x));
    return s;
}

}  // namespace cv
