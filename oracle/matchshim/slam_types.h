// ORACLE — test infrastructure only (never linked into the product library).
//
// Force-included (-include) in front of /root/reference/src/ORBmatcher.cc.  It claims the include guards of the
// reference's MapPoint.h / KeyFrame.h / Frame.h (which drag in PCL, Eigen, g2o, PEAC, DBoW2 vocabulary — none
// installed here) and supplies array-backed stand-ins with exactly the members ORBmatcher.cc touches.  The reference's
// own include/ORBmatcher.h and src/ORBmatcher.cc are then compiled unmodified.  Member semantics are restated from:
//   Frame::GetFeaturesInArea / PosInGrid / AssignFeaturesToGrid   src/Frame.cc:599-614,696-761
//   KeyFrame::GetFeaturesInArea / IsInImage                        src/KeyFrame.cc:608-652
//   MapPoint::PredictScale / Get{Min,Max}DistanceInvariance       src/MapPoint.cc:373-394
//
// Second mode, -DEAOF_REAL_MAPPOINT (oracle/Makefile target `mappointref`): MapPoint is NOT replaced — the reference's
// own include/MapPoint.h + src/MapPoint.cc are compiled unmodified on top of the KeyFrame / Frame stand-ins below plus a
// Map stand-in, so that MapPoint::ComputeDistinctiveDescriptors, Replace, AddObservation, PredictScale ... are the
// reference's code (oracle/mappoint_ref_harness.cc).
#pragma once
#ifndef EAOF_REAL_MAPPOINT
#define MAPPOINT_H
#else
#define MAP_H
#endif
#define KEYFRAME_H
#define FRAME_H

#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include "FeatureVector.h"  // the reference's own Thirdparty/DBoW2/DBoW2/FeatureVector.h
#include "mcv.h"

#define FRAME_GRID_ROWS 48  // include/Frame.h:89-90
#define FRAME_GRID_COLS 64

namespace ORB_SLAM2 {
using namespace std;  // the reference headers do the same (include/Frame.h), and ORBmatcher.cc relies on it

class KeyFrame;
class Frame;

#ifdef EAOF_REAL_MAPPOINT
class MapPoint;
class Map {  // include/Map.h: what src/MapPoint.cc touches
public:
    std::mutex mMutexPointCreation;
    void EraseMapPoint(MapPoint* pMP) { erased.push_back(pMP); }
    std::vector<MapPoint*> erased;
};
#else
class MapPoint {
public:
    MapPoint() : mbTrackInView(false), mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), mnTrackScaleLevel(0),
                 mTrackViewCos(1), bad(false), nObs(1), minDist(0), maxDist(1e30f), mfMaxDistance(1e30f) {}
    bool isBad() { return bad; }
    cv::Mat GetDescriptor() { return desc.clone(); }
    cv::Mat GetWorldPos() { return pos.clone(); }
    cv::Mat GetNormal() { return normal.clone(); }
    int Observations() { return nObs; }
    float GetMinDistanceInvariance() { return 0.8f * minDist; }   // src/MapPoint.cc:373-377
    float GetMaxDistanceInvariance() { return 1.2f * maxDist; }   // src/MapPoint.cc:379-383
    int PredictScale(const float& currentDist, const float& logScaleFactor) {  // src/MapPoint.cc:385-394
        float ratio = mfMaxDistance / currentDist;
        return ceil(log(ratio) / logScaleFactor);
    }
    bool IsInKeyFrame(KeyFrame* pKF) { return obs.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return obs.count(pKF) ? (int)obs[pKF] : -1; }
    void AddObservation(KeyFrame* pKF, size_t idx) { if (!obs.count(pKF)) { obs[pKF] = idx; ++nObs; } }
    void Replace(MapPoint* pMP) { bad = true; replaced = pMP; }

    bool mbTrackInView;
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    int mnTrackScaleLevel;
    float mTrackViewCos;

    // harness-side state
    bool bad;
    int nObs;
    float minDist, maxDist, mfMaxDistance;
    cv::Mat desc, pos, normal;
    std::map<KeyFrame*, size_t> obs;
    MapPoint* replaced = nullptr;
};
#endif

// include/ORBextractor.h:85 — the one member Frame::ComputeStereoMatches reads
class ORBextractor {
public:
    std::vector<cv::Mat> mvImagePyramid;
};

struct GridOwner {
    std::vector<size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
};

class Frame : public GridOwner {
public:
    Frame() : N(0), mb(0), mbf(0), mfLogScaleFactor(0), mnScaleLevels(0) {}
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY) {
        posX = round((kp.pt.x - mnMinX) * mfGridElementWidthInv);
        posY = round((kp.pt.y - mnMinY) * mfGridElementHeightInv);
        if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) return false;
        return true;
    }
    void AssignFeaturesToGrid() {
        for (int i = 0; i < N; i++) {
            const cv::KeyPoint& kp = mvKeysUn[i];
            int nGridPosX, nGridPosY;
            if (PosInGrid(kp, nGridPosX, nGridPosY)) mGrid[nGridPosX][nGridPosY].push_back(i);
        }
    }
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                     const int maxLevel = -1) const {
        vector<size_t> vIndices;
        vIndices.reserve(N);
        const int nMinCellX = max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
        if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
        const int nMaxCellX = min((int)FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
        if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
        const int nMaxCellY = min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
        if (nMaxCellY < 0) return vIndices;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const vector<size_t> vCell = mGrid[ix][iy];
                if (vCell.empty()) continue;
                for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                    const cv::KeyPoint& kpUn = mvKeysUn[vCell[j]];
                    if (bCheckLevels) {
                        if (kpUn.octave < minLevel) continue;
                        if (maxLevel >= 0)
                            if (kpUn.octave > maxLevel) continue;
                    }
                    const float distx = kpUn.pt.x - x;
                    const float disty = kpUn.pt.y - y;
                    if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
                }
            }
        }
        return vIndices;
    }

    cv::Mat GetCameraCenter() { return mOw.clone(); }
    long unsigned int mnId = 0;
    cv::Mat mOw;
    // stereo / RGB-D members; the two methods are the reference's own text (src/Frame.cc:841-1037), cut out at build time
    // by `make -C oracle stereoref` and compiled in oracle/stereo_ref_harness.cc
    void ComputeStereoMatches();
    void ComputeStereoFromRGBD(const cv::Mat& imDepth);
    std::vector<cv::KeyPoint> mvKeysRight;
    cv::Mat mDescriptorsRight;
    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;

    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mTcw;
    float mb, mbf;
    float mfLogScaleFactor;
    int mnScaleLevels;
    std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    // statics in the reference (include/Frame.h); per-object here so tests can vary them
    float fx = 1, fy = 1, cx = 0, cy = 0;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
};

class KeyFrame : public GridOwner {
public:
    KeyFrame() : N(0), mb(0), mbf(0), mfLogScaleFactor(0), mnScaleLevels(0), mnGridCols(FRAME_GRID_COLS), mnGridRows(FRAME_GRID_ROWS) {}
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const {
        vector<size_t> vIndices;
        vIndices.reserve(N);
        const int nMinCellX = max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
        if (nMinCellX >= mnGridCols) return vIndices;
        const int nMaxCellX = min((int)mnGridCols - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
        if (nMinCellY >= mnGridRows) return vIndices;
        const int nMaxCellY = min((int)mnGridRows - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
        if (nMaxCellY < 0) return vIndices;
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const vector<size_t> vCell = mGrid[ix][iy];
                for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                    const cv::KeyPoint& kpUn = mvKeysUn[vCell[j]];
                    const float distx = kpUn.pt.x - x;
                    const float disty = kpUn.pt.y - y;
                    if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
                }
            }
        return vIndices;
    }
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    std::set<MapPoint*> GetMapPoints();  // src/KeyFrame.cc:252-265, defined below (needs the complete MapPoint)
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    void EraseMapPointMatch(const size_t& idx) { mvpMapPoints[idx] = static_cast<MapPoint*>(NULL); }
    void ReplaceMapPointMatch(const size_t& idx, MapPoint* pMP) { mvpMapPoints[idx] = pMP; }
    bool isBad() { return mbBad; }
    long unsigned int mnId = 0, mnFrameId = 0;
    bool mbBad = false;
    cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }

    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat Tcw, Ow;
    float mb, mbf;
    float mfLogScaleFactor;
    int mnScaleLevels;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    float fx = 1, fy = 1, cx = 0, cy = 0;
    int mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;  // ints in KeyFrame (include/KeyFrame.h), floats in Frame
    int mnGridCols, mnGridRows;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
};

}  // namespace ORB_SLAM2

#ifdef EAOF_REAL_MAPPOINT
#include <mutex>
#include "MapPoint.h"  // the reference's own include/MapPoint.h; its KeyFrame.h / Frame.h / Map.h includes are claimed above
#endif

namespace ORB_SLAM2 {
inline std::set<MapPoint*> KeyFrame::GetMapPoints() {
    std::set<MapPoint*> s;
    for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p);
    return s;
}
}  // namespace ORB_SLAM2
