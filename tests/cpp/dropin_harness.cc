// TEST INFRASTRUCTURE — C harness around the drop-in ORB_SLAM2::ORBextractor (eao-fusion_b200/dropin), compiled
// against oracle/cvshim in place of OpenCV (this container has no OpenCV C++), so that tests/test_gpu_dropin.py can
// call the class exactly the way Frame::ExtractORB does (src/Frame.cc:616-622) and compare it with the reference
// class behind oracle/_ref/liborb_ref.so, which exposes the same C calls (oracle/ref_harness.cc).
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "ORBextractor.h"  // the drop-in header

namespace cv {
static int g_shim_blur_mode = 0;  // which of the three blur arithmetics the stand-in cv::GaussianBlur computes
int eaof_shim_blur_mode() { return g_shim_blur_mode; }
void eaof_shim_set_blur_mode(int m) { g_shim_blur_mode = m; }
}  // namespace cv

extern "C" {

struct dropin_kp { float x, y, size, angle, response; int octave; int class_id; };

void* dropin_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    try {
        return new ORB_SLAM2::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
    } catch (...) { return nullptr; }
}
void dropin_destroy(void* h) { delete (ORB_SLAM2::ORBextractor*)h; }
void dropin_set_blur_mode(void* h, int m) { ((ORB_SLAM2::ORBextractor*)h)->SetBlurMode(m); }
int dropin_blur_mode(void* h) { return ((ORB_SLAM2::ORBextractor*)h)->BlurMode(); }
void dropin_shim_set_blur_mode(int m) { cv::eaof_shim_set_blur_mode(m); }  // the "OpenCV build" the next constructor probes
void dropin_set_pyramid(void* h, int on) { ((ORB_SLAM2::ORBextractor*)h)->SetPyramidDownload(on != 0); }

void dropin_tables(void* h, float* sf, float* isf, float* s2, float* is2, int* levels, float* scale) {
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(),
                       d = e->GetInverseScaleSigmaSquares();
    *levels = e->GetLevels();
    *scale = e->GetScaleFactor();
    for (size_t i = 0; i < a.size(); ++i) { sf[i] = a[i]; isf[i] = b[i]; s2[i] = c[i]; is2[i] = d[i]; }
}

// Calls operator()(image, cv::Mat(), keypoints, descriptors).  pre_n > 0 pre-fills the outputs with pre_n dummy
// entries first, to observe the "untouched on empty image" and "released on zero keypoints" behaviours.
// Returns keypoints.size(); *desc_rows receives descriptors.rows; -2 = exception (no CUDA device etc.).
int dropin_extract(void* h, const uint8_t* img, int w, int hgt, size_t stride, dropin_kp* kps, uint8_t* desc, int cap,
                   int pre_n, int* desc_rows) {
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    std::vector<cv::KeyPoint> k(pre_n > 0 ? pre_n : 0);
    cv::Mat d;
    if (pre_n > 0) d.create(pre_n, 32, CV_8U);
    try {
        if (img && w > 0 && hgt > 0) {
            cv::Mat image(hgt, w, CV_8UC1, (void*)img, stride);
            (*e)(image, cv::Mat(), k, d);
        } else {
            (*e)(cv::Mat(), cv::Mat(), k, d);
        }
    } catch (...) { return -2; }
    const int n = (int)k.size();
    for (int i = 0; i < n && i < cap; ++i) {
        if (kps) kps[i] = {k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave, k[i].class_id};
        if (desc && i < d.rows) memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
    }
    if (desc_rows) *desc_rows = d.rows;
    return n;
}

// Times `calls` consecutive operator() calls on the C++ side (steady_clock around the call, nothing else inside): the latency
// a Frame constructor sees (src/Frame.cc:193,229-231 print exactly this).  Frames cycle through imgs[0..n_imgs).
int dropin_time_calls(void* h, const uint8_t* imgs, int n_imgs, int w, int hgt, size_t stride, int calls, double* out_us) {
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    std::vector<cv::KeyPoint> k;
    cv::Mat d;
    int n = 0;
    try {
        for (int i = 0; i < calls; ++i) {
            cv::Mat image(hgt, w, CV_8UC1, (void*)(imgs + (size_t)(i % n_imgs) * stride * hgt), stride);
            const auto t0 = std::chrono::steady_clock::now();
            (*e)(image, cv::Mat(), k, d);
            out_us[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            n = (int)k.size();
        }
    } catch (...) { return -2; }
    return n;
}

// mvImagePyramid[level]: size, and a copy of the ROI (with_border=0) or of the 19-px bordered parent buffer.
int dropin_level(void* h, int level, int* w, int* hgt, uint8_t* dst, size_t dstride, int with_border) {
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    if (level < 0 || level >= (int)e->mvImagePyramid.size()) return -1;
    const cv::Mat& m = e->mvImagePyramid[level];
    if (m.empty()) return -1;
    *w = m.cols;
    *hgt = m.rows;
    if (!dst) return 0;
    if (with_border) {
        const uint8_t* base = m.data - 19 * (size_t)m.step - 19;
        for (int y = 0; y < m.rows + 38; ++y) memcpy(dst + (size_t)y * dstride, base + (size_t)y * m.step, m.cols + 38);
    } else {
        for (int y = 0; y < m.rows; ++y) memcpy(dst + (size_t)y * dstride, m.ptr(y), m.cols);
    }
    return 0;
}

}  // extern "C"
