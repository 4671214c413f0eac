// ORACLE — test infrastructure only (never linked into the product library).
//
// C harness around the UNMODIFIED reference translation units /root/reference/src/MapPoint.cc and src/ORBmatcher.cc,
// compiled in place by `make -C oracle mappointref` into oracle/_ref/libmappoint_ref.so with
// -DEAOF_REAL_MAPPOINT -include matchshim/slam_types.h: KeyFrame / Frame / Map are the array-backed stand-ins, MapPoint
// is the reference's own class.  It pins the restatement of MapPoint::ComputeDistinctiveDescriptors
// (src/MapPoint.cc:242-307, oracle/match_oracle.cc eaoo_distinctive_descriptor) to the reference's code as run here.
#include <cstdint>
#include <cstring>
#include <deque>
#include <vector>

#include "ORBmatcher.h"

using namespace ORB_SLAM2;

extern "C" {

// One map point observed by n keyframes, observation i = row i of `desc` (32 bytes each); kfBad[i]: KeyFrame::isBad().
// Observations are inserted in the order the reference will iterate them (std::map<KeyFrame*, size_t> orders by
// pointer: the keyframes live in one array, so address order = index order).  Writes the 32-byte descriptor the
// reference selects into `out`; returns the index of the (first) observation row equal to it among the good
// keyframes, -1 when the point kept its initial descriptor (no usable observation).
int mpref_distinctive_descriptor(int n, const uint8_t* desc, const uint8_t* kfBad, uint8_t* out) {
    Map map;
    std::vector<KeyFrame> kfs(n > 0 ? n : 1);
    cv::Mat pos(3, 1, CV_32F);
    pos.at<float>(2) = 1.f;
    KeyFrame ref;  // reference keyframe of the point, not an observer
    ref.N = 1;
    ref.mDescriptors = cv::Mat(1, 32, CV_8U);
    memset(ref.mDescriptors.data, 0xA5, 32);
    MapPoint mp(pos, &ref, &map);
    for (int i = 0; i < n; ++i) {
        KeyFrame& kf = kfs[i];
        kf.N = 1;
        kf.mnId = i;
        kf.mbBad = kfBad && kfBad[i];
        kf.mDescriptors = cv::Mat(1, 32, CV_8U);
        memcpy(kf.mDescriptors.data, desc + 32 * (size_t)i, 32);
        kf.mvuRight.assign(1, -1.f);
        kf.mvpMapPoints.assign(1, &mp);
        mp.AddObservation(&kf, 0);
    }
    mp.ComputeDistinctiveDescriptors();
    cv::Mat d = mp.GetDescriptor();
    if (d.empty()) return -1;
    memcpy(out, d.data, 32);
    for (int i = 0; i < n; ++i)
        if (!(kfBad && kfBad[i]) && memcmp(out, desc + 32 * (size_t)i, 32) == 0) return i;
    return -1;
}

// MapPoint::PredictScale of the real class (src/MapPoint.cc:385-394) for a point whose mfMaxDistance is set through
// UpdateNormalAndDepth is not reachable without a full keyframe; the formula is checked through the public method on a
// point built from a Frame (constructor :46-70 sets mfMaxDistance = dist * scaleFactor[level]).
int mpref_predict_scale(float dist0, float levelScale, float currentDist, float logScaleFactor) {
    Map map;
    Frame F;
    F.N = 1;
    F.mOw = cv::Mat(3, 1, CV_32F);  // camera at the origin
    F.mvKeysUn.assign(1, cv::KeyPoint(0.f, 0.f, 31.f, 0.f, 0.f, 0));
    F.mvScaleFactors.assign(8, levelScale);
    F.mnScaleLevels = 8;
    F.mDescriptors = cv::Mat(1, 32, CV_8U);
    cv::Mat pos(3, 1, CV_32F);
    pos.at<float>(2) = dist0;
    MapPoint mp(pos, &map, &F, 0);
    return mp.PredictScale(currentDist, logScaleFactor);
}

}  // extern "C"
