// ORACLE — test infrastructure only (never linked into the product library).
//
// Frame::ComputeStereoMatches (src/Frame.cc:841-1013) and Frame::ComputeStereoFromRGBD (:1016-1037) of the reference.
// src/Frame.cc as a whole cannot be compiled here (include/Frame.h pulls in PCL, Eigen, g2o, PEAC), so
// `make -C oracle stereoref` cuts the text of exactly these two member functions out of /root/reference/src/Frame.cc
// into oracle/_ref/frame_stereo_body.inc (git-ignored build output, never committed) and this file compiles it,
// unmodified, as members of the stand-in Frame of matchshim/slam_types.h, next to the unmodified src/ORBmatcher.cc
// (ORBmatcher::TH_HIGH, DescriptorDistance) -> oracle/_ref/libstereo_ref.so.  It pins eaoo_stereo_matches and
// o_stereo_from_rgbd to the reference's code as run here.
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>

#include "ORBmatcher.h"  // the reference's own header

namespace ORB_SLAM2 {
#include "_ref/frame_stereo_body.inc"
}

using namespace ORB_SLAM2;

namespace {
void fill(std::vector<cv::KeyPoint>& k, int n, const float* x, const float* y, const int* oct) {
    k.resize(n);
    for (int i = 0; i < n; ++i) k[i] = cv::KeyPoint(x[i], y[i], 31.f, 0.f, 0.f, oct[i]);
}
cv::Mat rows32(const uint8_t* d, int n) {
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) memcpy(m.data, d, 32 * (size_t)n);
    return m;
}
// level images as ROI views at (19,19) of bordered buffers, the layout ComputePyramid leaves (src/ORBextractor.cc:1114-1116)
void pyramid(ORBextractor& e, int nLevels, const uint8_t* buf, const int* off, const int* w, const int* h) {
    e.mvImagePyramid.resize(nLevels);
    for (int l = 0; l < nLevels; ++l) {
        cv::Mat full(h[l] + 38, w[l] + 38, CV_8U);
        memcpy(full.data, buf + off[l], (size_t)(w[l] + 38) * (h[l] + 38));
        e.mvImagePyramid[l] = full.rowRange(19, 19 + h[l]).colRange(19, 19 + w[l]);
    }
}
}  // namespace

extern "C" {

// Pyramids: bordered level buffers back to back (level l at off[l], (w[l]+38) x (h[l]+38) bytes).
void sref_stereo_matches(int nL, const float* xL, const float* yL, const int* octL, const uint8_t* descL, int nR,
                         const float* xR, const float* yR, const int* octR, const uint8_t* descR, int nLevels,
                         const float* scale, const float* invScale, const uint8_t* pyrL, const uint8_t* pyrR, const int* off,
                         const int* w, const int* h, float mb, float mbf, float* uRight, float* depth) {
    Frame F;
    ORBextractor eL, eR;
    F.N = nL;
    fill(F.mvKeys, nL, xL, yL, octL);
    fill(F.mvKeysRight, nR, xR, yR, octR);
    F.mDescriptors = rows32(descL, nL);
    F.mDescriptorsRight = rows32(descR, nR);
    F.mvScaleFactors.assign(scale, scale + nLevels);
    F.mvInvScaleFactors.assign(invScale, invScale + nLevels);
    F.mb = mb;
    F.mbf = mbf;
    pyramid(eL, nLevels, pyrL, off, w, h);
    pyramid(eR, nLevels, pyrR, off, w, h);
    F.mpORBextractorLeft = &eL;
    F.mpORBextractorRight = &eR;
    F.ComputeStereoMatches();
    for (int i = 0; i < nL; ++i) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

void sref_stereo_from_rgbd(int n, const float* x, const float* y, const float* xUn, const float* depthMap, int w, int h,
                           float mbf, float* uRight, float* depth) {
    Frame F;
    F.N = n;
    F.mvKeys.resize(n);
    F.mvKeysUn.resize(n);
    for (int i = 0; i < n; ++i) {
        F.mvKeys[i] = cv::KeyPoint(x[i], y[i], 31.f);
        F.mvKeysUn[i] = cv::KeyPoint(xUn ? xUn[i] : x[i], y[i], 31.f);
    }
    F.mbf = mbf;
    cv::Mat im(h, w, CV_32F);
    memcpy(im.data, depthMap, sizeof(float) * (size_t)w * h);
    F.ComputeStereoFromRGBD(im);
    for (int i = 0; i < n; ++i) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

}  // extern "C"
