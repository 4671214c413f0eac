// ORACLE — test infrastructure only (never linked into the product library).
//
// C harness around the UNMODIFIED reference translation unit /root/reference/src/ORBmatcher.cc, which
// oracle/Makefile (target `matchref`) compiles in place against oracle/matchshim/ into oracle/_ref/libmatch_ref.so.
// Every entry point takes the same plain arrays as the restatement in match_oracle.cc (eaoo_*), builds the
// array-backed Frame / KeyFrame / MapPoint stand-ins of matchshim/slam_types.h from them, calls the reference's
// own ORBmatcher method and translates the MapPoint* results back to indices.  tests/test_oracle_matcher_vs_ref.py
// compares the two, which pins the restatement (and therefore the CUDA path) to the reference's code as run here.
//
// Geometry trick: the reference projects map points itself (x3Dc = Rcw*x3Dw+tcw, u = fx*xc*invzc+cx).  The harness
// uses identity rotations, fx=fy=1, cx=cy=0 and world points (u/invz, v/invz, 1/invz) with invz a power of two, so the
// projection reproduces the (u, v, 1/z) arrays of the test exactly in float arithmetic.
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "ORBmatcher.h"  // the reference's own header, /root/reference/include

using namespace ORB_SLAM2;

namespace {

cv::Mat eye4() {
    cv::Mat m(4, 4, CV_32F);
    for (int i = 0; i < 4; ++i) m.at<float>(i, i) = 1.f;
    return m;
}
cv::Mat vec3(float x, float y, float z) {
    cv::Mat m(3, 1, CV_32F);
    m.at<float>(0) = x; m.at<float>(1) = y; m.at<float>(2) = z;
    return m;
}
cv::Mat desc_rows(const uint8_t* d, int n) {
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) memcpy(m.data, d, 32 * (size_t)n);
    return m;
}
cv::Mat desc_row(const uint8_t* d) {
    cv::Mat m(1, 32, CV_8U);
    memcpy(m.data, d, 32);
    return m;
}
void fill_featvec(DBoW2::FeatureVector& fv, int nNodes, const int* id, const int* start, const int* idx) {
    for (int a = 0; a < nNodes; ++a) {
        std::vector<unsigned int>& v = fv[(DBoW2::NodeId)id[a]];
        for (int i = start[a]; i < start[a + 1]; ++i) v.push_back((unsigned)idx[i]);
    }
}
void fill_keys(std::vector<cv::KeyPoint>& k, int n, const float* x, const float* y, const int* oct, const float* angle) {
    k.resize(n);
    for (int i = 0; i < n; ++i) k[i] = cv::KeyPoint(x ? x[i] : 0.f, y ? y[i] : 0.f, 31.f, angle ? angle[i] : 0.f, 0.f, oct ? oct[i] : 0);
}
struct Pool {
    std::deque<MapPoint> mp;
    std::map<MapPoint*, int> tag;
    MapPoint* make(int t) {
        mp.emplace_back();
        tag[&mp.back()] = t;
        return &mp.back();
    }
    int of(MapPoint* p) const {
        if (!p) return -1;
        auto it = tag.find(p);
        return it == tag.end() ? -1 : it->second;
    }
};
// "map point absent or bad": alternate between the two encodings the reference treats alike
MapPoint* invalid_mp(Pool& pool, int i) {
    if (i & 1) return nullptr;
    MapPoint* p = pool.make(-1);
    p->bad = true;
    return p;
}
void set_bounds(Frame& F, float minX, float maxX, float minY, float maxY, float invW, float invH) {
    F.mnMinX = minX; F.mnMaxX = maxX; F.mnMinY = minY; F.mnMaxY = maxY;
    F.mfGridElementWidthInv = invW; F.mfGridElementHeightInv = invH;
}

}  // namespace

extern "C" {

int mref_th_low() { return ORBmatcher::TH_LOW; }
int mref_th_high() { return ORBmatcher::TH_HIGH; }
int mref_histo_length() { return ORBmatcher::HISTO_LENGTH; }

// ORBmatcher::DescriptorDistance  src/ORBmatcher.cc:1649-1665
int mref_hamming(const uint8_t* a, const uint8_t* b) {
    return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b));
}

// MapPoint::PredictScale as the shim restates it (src/MapPoint.cc:385-394), for tests that need the level the
// reference will derive from a 3-D distance.
int mref_predict_scale(float maxDistance, float dist, float logScaleFactor) {
    MapPoint p;
    p.mfMaxDistance = maxDistance;
    return p.PredictScale(dist, logScaleFactor);
}
float mref_norm3(float x, float y, float z) { return (float)cv::norm(vec3(x, y, z)); }

// SearchByBoW, both overloads (src/ORBmatcher.cc:159-288, 522-655); arguments as eaoo_search_by_bow.
int mref_search_by_bow(int mode, int nQ, const uint8_t* descQ, const float* angleQ, const uint8_t* validQ, int nT,
                       const uint8_t* descT, const float* angleT, const uint8_t* validT, int nNodesQ, const int* nodeIdQ,
                       const int* nodeStartQ, const int* nodeIdxQ, int nNodesT, const int* nodeIdT, const int* nodeStartT,
                       const int* nodeIdxT, float nnratio, int checkOri, int* matchOut) {
    Pool pool;
    ORBmatcher matcher(nnratio, checkOri != 0);
    KeyFrame kf;
    kf.N = nQ;
    fill_keys(kf.mvKeysUn, nQ, nullptr, nullptr, nullptr, angleQ);
    kf.mvKeys = kf.mvKeysUn;
    kf.mDescriptors = desc_rows(descQ, nQ);
    kf.mvpMapPoints.resize(nQ);
    for (int i = 0; i < nQ; ++i) kf.mvpMapPoints[i] = (!validQ || validQ[i]) ? pool.make(i) : invalid_mp(pool, i);
    fill_featvec(kf.mFeatVec, nNodesQ, nodeIdQ, nodeStartQ, nodeIdxQ);
    if (mode == 0) {
        Frame F;
        F.N = nT;
        fill_keys(F.mvKeysUn, nT, nullptr, nullptr, nullptr, angleT);
        F.mvKeys = F.mvKeysUn;
        F.mDescriptors = desc_rows(descT, nT);
        fill_featvec(F.mFeatVec, nNodesT, nodeIdT, nodeStartT, nodeIdxT);
        std::vector<MapPoint*> out;
        const int n = matcher.SearchByBoW(&kf, F, out);
        for (int t = 0; t < nT; ++t) matchOut[t] = pool.of(out[t]);
        return n;
    }
    KeyFrame kf2;
    kf2.N = nT;
    fill_keys(kf2.mvKeysUn, nT, nullptr, nullptr, nullptr, angleT);
    kf2.mvKeys = kf2.mvKeysUn;
    kf2.mDescriptors = desc_rows(descT, nT);
    kf2.mvpMapPoints.resize(nT);
    for (int i = 0; i < nT; ++i) kf2.mvpMapPoints[i] = (!validT || validT[i]) ? pool.make(100000 + i) : invalid_mp(pool, i);
    fill_featvec(kf2.mFeatVec, nNodesT, nodeIdT, nodeStartT, nodeIdxT);
    std::vector<MapPoint*> out;
    const int n = matcher.SearchByBoW(&kf, &kf2, out);
    for (int q = 0; q < nQ; ++q) {
        const int t = pool.of(out[q]);
        matchOut[q] = t >= 100000 ? t - 100000 : -1;
    }
    return n;
}

// SearchForTriangulation (src/ORBmatcher.cc:657-823); arguments as eaoo_search_for_triangulation.
int mref_search_for_triangulation(int n1, const uint8_t* desc1, const float* x1, const float* y1, const float* angle1,
                                  const uint8_t* free1, const uint8_t* stereo1, int n2, const uint8_t* desc2,
                                  const float* x2, const float* y2, const int* oct2, const float* angle2,
                                  const uint8_t* free2, const uint8_t* stereo2, int nNodes1, const int* nodeId1,
                                  const int* nodeStart1, const int* nodeIdx1, int nNodes2, const int* nodeId2,
                                  const int* nodeStart2, const int* nodeIdx2, const float* F12, float ex, float ey,
                                  const float* scaleFactors2, const float* levelSigma2_2, int nLevels, int onlyStereo,
                                  int checkOri, float nnratio, int* match12) {
    Pool pool;
    ORBmatcher matcher(nnratio, checkOri != 0);
    KeyFrame k1, k2;
    k1.N = n1;
    fill_keys(k1.mvKeysUn, n1, x1, y1, nullptr, angle1);
    k1.mDescriptors = desc_rows(desc1, n1);
    k1.mvpMapPoints.resize(n1);
    k1.mvuRight.resize(n1);
    for (int i = 0; i < n1; ++i) {
        k1.mvpMapPoints[i] = free1[i] ? nullptr : pool.make(i);
        k1.mvuRight[i] = (stereo1 && stereo1[i]) ? 1.f : -1.f;
    }
    fill_featvec(k1.mFeatVec, nNodes1, nodeId1, nodeStart1, nodeIdx1);
    k1.Ow = vec3(ex, ey, 1.f);  // with R2w = I, t2w = 0, fx = fy = 1, cx = cy = 0 the epipole is exactly (ex, ey)
    k2.N = n2;
    fill_keys(k2.mvKeysUn, n2, x2, y2, oct2, angle2);
    k2.mDescriptors = desc_rows(desc2, n2);
    k2.mvpMapPoints.resize(n2);
    k2.mvuRight.resize(n2);
    for (int i = 0; i < n2; ++i) {
        k2.mvpMapPoints[i] = free2[i] ? nullptr : pool.make(i);
        k2.mvuRight[i] = (stereo2 && stereo2[i]) ? 1.f : -1.f;
    }
    fill_featvec(k2.mFeatVec, nNodes2, nodeId2, nodeStart2, nodeIdx2);
    k2.Tcw = eye4();
    k2.mvScaleFactors.assign(scaleFactors2, scaleFactors2 + nLevels);
    k2.mvLevelSigma2.assign(levelSigma2_2, levelSigma2_2 + nLevels);
    cv::Mat F(3, 3, CV_32F);
    for (int i = 0; i < 9; ++i) F.at<float>(i / 3, i % 3) = F12[i];
    std::vector<std::pair<size_t, size_t>> pairs;
    const int n = matcher.SearchForTriangulation(&k1, &k2, F, pairs, onlyStereo != 0);
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (auto& p : pairs) match12[p.first] = (int)p.second;
    return n;
}

// SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (src/ORBmatcher.cc:1328-1472); arguments as
// eaoo_search_by_projection_last.  linvz entries must be powers of two (see the geometry trick above).
int mref_search_by_projection_last(int nC, const float* cx, const float* cy, const int* coct, const float* cangle,
                                   const uint8_t* cdesc, const float* curight, const uint8_t* ctaken, float minX,
                                   float maxX, float minY, float maxY, float invW, float invH, int nL,
                                   const uint8_t* lvalid, const float* lu, const float* lv, const float* linvz,
                                   const int* loct, const float* langle, const uint8_t* ldesc, const uint8_t* lobs,
                                   const float* scaleFactors, int nLevels, float th, float mbf, int searchMode,
                                   int checkOri, float nnratio, int* matchCur) {
    Pool pool;
    ORBmatcher matcher(nnratio, checkOri != 0);
    Frame cur, last;
    cur.N = nC;
    fill_keys(cur.mvKeysUn, nC, cx, cy, coct, cangle);
    cur.mvKeys = cur.mvKeysUn;
    cur.mDescriptors = desc_rows(cdesc, nC);
    cur.mvuRight.resize(nC);
    cur.mvpMapPoints.resize(nC);
    for (int i = 0; i < nC; ++i) {
        cur.mvuRight[i] = curight ? curight[i] : -1.f;
        cur.mvpMapPoints[i] = (ctaken && ctaken[i]) ? pool.make(-1) : nullptr;  // Observations() == 1
    }
    set_bounds(cur, minX, maxX, minY, maxY, invW, invH);
    cur.AssignFeaturesToGrid();
    cur.mvScaleFactors.assign(scaleFactors, scaleFactors + nLevels);
    cur.mTcw = eye4();
    cur.mb = 1.f;
    cur.mbf = mbf;
    last.N = nL;
    fill_keys(last.mvKeysUn, nL, nullptr, nullptr, loct, langle);
    last.mvKeys = last.mvKeysUn;
    last.mvpMapPoints.resize(nL);
    last.mvbOutlier.assign(nL, false);
    for (int i = 0; i < nL; ++i) {
        if (lvalid && !lvalid[i]) {
            if (i & 1) last.mvpMapPoints[i] = nullptr;
            else { last.mvpMapPoints[i] = pool.make(-1); last.mvbOutlier[i] = true; }
            continue;
        }
        MapPoint* p = pool.make(i);
        const float invz = linvz ? linvz[i] : 1.f, z = 1.f / invz;
        p->pos = vec3(lu[i] * z, lv[i] * z, z);
        p->desc = desc_row(ldesc + 32 * (size_t)i);
        p->nObs = lobs ? (lobs[i] ? 1 : 0) : 1;
        last.mvpMapPoints[i] = p;
    }
    last.mTcw = eye4();
    // bForward / bBackward (:1348-1351) come from the relative pose: tlc = tlw when the current pose is identity
    const bool bMono = curight == nullptr;
    if (searchMode == 1) last.mTcw.at<float>(2, 3) = 10.f;
    if (searchMode == 2) last.mTcw.at<float>(2, 3) = -10.f;
    const int n = matcher.SearchByProjection(cur, last, th, bMono);
    for (int k = 0; k < nC; ++k) matchCur[k] = pool.of(cur.mvpMapPoints[k]);
    return n;
}

// SearchByProjection(Frame& F, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129).  Per map point: in_view
// (mbTrackInView), bad, projection (x, y, xr), predicted level, view cosine, descriptor, obs (Observations()>0).
int mref_search_by_projection_mappoints(int nF, const float* fx_, const float* fy_, const int* foct, const uint8_t* fdesc,
                                        const float* furight, const uint8_t* ftaken, float minX, float maxX, float minY,
                                        float maxY, float invW, float invH, int nMP, const uint8_t* inView,
                                        const uint8_t* bad, const float* px, const float* py, const float* pxr,
                                        const int* plevel, const float* pcos, const uint8_t* pdesc, const uint8_t* pobs,
                                        const float* scaleFactors, int nLevels, float th, float nnratio, int* matchF) {
    Pool pool;
    ORBmatcher matcher(nnratio, true);
    Frame F;
    F.N = nF;
    fill_keys(F.mvKeysUn, nF, fx_, fy_, foct, nullptr);
    F.mvKeys = F.mvKeysUn;
    F.mDescriptors = desc_rows(fdesc, nF);
    F.mvuRight.resize(nF);
    F.mvpMapPoints.resize(nF);
    for (int i = 0; i < nF; ++i) {
        F.mvuRight[i] = furight ? furight[i] : -1.f;
        F.mvpMapPoints[i] = (ftaken && ftaken[i]) ? pool.make(-1) : nullptr;
    }
    set_bounds(F, minX, maxX, minY, maxY, invW, invH);
    F.AssignFeaturesToGrid();
    F.mvScaleFactors.assign(scaleFactors, scaleFactors + nLevels);
    std::vector<MapPoint*> mps(nMP);
    for (int i = 0; i < nMP; ++i) {
        MapPoint* p = pool.make(i);
        p->mbTrackInView = !inView || inView[i];
        p->bad = bad && bad[i];
        p->mTrackProjX = px[i]; p->mTrackProjY = py[i]; p->mTrackProjXR = pxr ? pxr[i] : 0.f;
        p->mnTrackScaleLevel = plevel[i];
        p->mTrackViewCos = pcos ? pcos[i] : 1.f;
        p->desc = desc_row(pdesc + 32 * (size_t)i);
        p->nObs = pobs ? (pobs[i] ? 1 : 0) : 1;
        mps[i] = p;
    }
    const int n = matcher.SearchByProjection(F, mps, th);
    for (int k = 0; k < nF; ++k) matchF[k] = pool.of(F.mvpMapPoints[k]);
    return n;
}

// SearchForInitialization (src/ORBmatcher.cc:405-520).  prev: in/out vbPrevMatched (x,y pairs, n1 entries).
int mref_search_for_initialization(int n1, const int* oct1, const float* angle1, const uint8_t* desc1, float* prev,
                                   int n2, const float* x2, const float* y2, const int* oct2, const float* angle2,
                                   const uint8_t* desc2, float minX, float maxX, float minY, float maxY, float invW,
                                   float invH, int windowSize, float nnratio, int checkOri, int* matches12) {
    ORBmatcher matcher(nnratio, checkOri != 0);
    Frame F1, F2;
    F1.N = n1;
    fill_keys(F1.mvKeysUn, n1, nullptr, nullptr, oct1, angle1);
    F1.mDescriptors = desc_rows(desc1, n1);
    F2.N = n2;
    fill_keys(F2.mvKeysUn, n2, x2, y2, oct2, angle2);
    F2.mDescriptors = desc_rows(desc2, n2);
    set_bounds(F2, minX, maxX, minY, maxY, invW, invH);
    F2.AssignFeaturesToGrid();
    std::vector<cv::Point2f> pm(n1);
    for (int i = 0; i < n1; ++i) pm[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
    std::vector<int> m12;
    const int n = matcher.SearchForInitialization(F1, F2, pm, m12, windowSize);
    for (int i = 0; i < n1; ++i) { matches12[i] = m12[i]; prev[2 * i] = pm[i].x; prev[2 * i + 1] = pm[i].y; }
    return n;
}

// SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1474-1601).  Per KF feature:
// state 0 = no map point, 1 = bad, 2 = already found, 3 = usable; world point (wx, wy, wz) seen from an identity
// pose with fx=fy=1, cx=cy=0; maxDist/minDist are mfMaxDistance/mfMinDistance of the map point.
int mref_search_by_projection_kf(int nC, const float* cx, const float* cy, const int* coct, const float* cangle,
                                 const uint8_t* cdesc, const uint8_t* ctaken, float minX, float maxX, float minY,
                                 float maxY, float invW, float invH, int nK, const uint8_t* kstate, const float* wx,
                                 const float* wy, const float* wz, const float* kmaxDist, const float* kminDist,
                                 const float* kangle, const uint8_t* kdesc, const float* scaleFactors, int nLevels,
                                 float logScaleFactor, float th, int orbDist, int checkOri, float nnratio, int* matchCur) {
    Pool pool;
    ORBmatcher matcher(nnratio, checkOri != 0);
    Frame cur;
    cur.N = nC;
    fill_keys(cur.mvKeysUn, nC, cx, cy, coct, cangle);
    cur.mvKeys = cur.mvKeysUn;
    cur.mDescriptors = desc_rows(cdesc, nC);
    cur.mvpMapPoints.resize(nC);
    for (int i = 0; i < nC; ++i) cur.mvpMapPoints[i] = (ctaken && ctaken[i]) ? pool.make(-1) : nullptr;
    set_bounds(cur, minX, maxX, minY, maxY, invW, invH);
    cur.AssignFeaturesToGrid();
    cur.mvScaleFactors.assign(scaleFactors, scaleFactors + nLevels);
    cur.mfLogScaleFactor = logScaleFactor;
    cur.mnScaleLevels = nLevels;
    cur.mTcw = eye4();
    KeyFrame kf;
    kf.N = nK;
    fill_keys(kf.mvKeysUn, nK, nullptr, nullptr, nullptr, kangle);
    kf.mvpMapPoints.resize(nK);
    std::set<MapPoint*> found;
    for (int i = 0; i < nK; ++i) {
        if (kstate[i] == 0) { kf.mvpMapPoints[i] = nullptr; continue; }
        MapPoint* p = pool.make(i);
        p->bad = kstate[i] == 1;
        if (kstate[i] == 2) found.insert(p);
        p->pos = vec3(wx[i], wy[i], wz[i]);
        p->mfMaxDistance = kmaxDist[i];
        p->maxDist = kmaxDist[i];
        p->minDist = kminDist[i];
        p->desc = desc_row(kdesc + 32 * (size_t)i);
        kf.mvpMapPoints[i] = p;
    }
    const int n = matcher.SearchByProjection(cur, &kf, found, th, orbDist);
    for (int k = 0; k < nC; ++k) matchCur[k] = pool.of(cur.mvpMapPoints[k]);
    return n;
}

// ---- map-side matchers of the LocalMapping / LoopClosing threads ------------------------------------------------------
// Geometry as above: keyframe pose identity, fx=fy=1, cx=cy=0, world points (wx, wy, wz) with wz = +-1, so that the
// projection is (wx, wy) exactly; the distance gate, viewing-angle gate and MapPoint::PredictScale run inside the
// reference on (maxDist, minDist, normal).

double mref_dot3(float ax, float ay, float az, float bx, float by, float bz) { return vec3(ax, ay, az).dot(vec3(bx, by, bz)); }

}  // extern "C"

namespace {
void assign_grid(KeyFrame& kf) {  // KeyFrame copies the Frame's grid (src/KeyFrame.cc:52-58); Frame::AssignFeaturesToGrid
    for (int i = 0; i < kf.N; i++) {
        const cv::KeyPoint& kp = kf.mvKeysUn[i];
        const int posX = round((kp.pt.x - kf.mnMinX) * kf.mfGridElementWidthInv);
        const int posY = round((kp.pt.y - kf.mnMinY) * kf.mfGridElementHeightInv);
        if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) continue;
        kf.mGrid[posX][posY].push_back(i);
    }
}
void setup_kf(KeyFrame& kf, int n, const float* x, const float* y, const int* oct, const uint8_t* desc, const float* uright,
              int minX, int maxX, int minY, int maxY, float invW, float invH, const float* scaleFactors,
              const float* invSigma2, int nLevels, float logScaleFactor) {
    kf.N = n;
    fill_keys(kf.mvKeysUn, n, x, y, oct, nullptr);
    kf.mvKeys = kf.mvKeysUn;
    kf.mDescriptors = desc_rows(desc, n);
    kf.mvuRight.resize(n);
    for (int i = 0; i < n; ++i) kf.mvuRight[i] = uright ? uright[i] : -1.f;
    kf.mvpMapPoints.assign(n, nullptr);
    kf.mnMinX = minX; kf.mnMaxX = maxX; kf.mnMinY = minY; kf.mnMaxY = maxY;
    kf.mfGridElementWidthInv = invW; kf.mfGridElementHeightInv = invH;
    assign_grid(kf);
    kf.mvScaleFactors.assign(scaleFactors, scaleFactors + nLevels);
    if (invSigma2) kf.mvInvLevelSigma2.assign(invSigma2, invSigma2 + nLevels);
    kf.mfLogScaleFactor = logScaleFactor;
    kf.mnScaleLevels = nLevels;
    kf.Tcw = eye4();
    kf.Ow = vec3(0.f, 0.f, 0.f);
}
void setup_point(MapPoint* p, int i, const float* wx, const float* wy, const float* wz, const float* maxDist,
                 const float* minDist, const float* normal, const uint8_t* desc) {
    p->pos = vec3(wx[i], wy[i], wz[i]);
    p->mfMaxDistance = maxDist[i];
    p->maxDist = maxDist[i];
    p->minDist = minDist[i];
    p->normal = normal ? vec3(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]) : vec3(wx[i], wy[i], wz[i]);
    p->desc = desc_row(desc + 32 * (size_t)i);
}
const int SLOT_TAG = 100000;  // tag of the map point a keyframe feature holds on entry = SLOT_TAG + feature index
}  // namespace

extern "C" {

// SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)  src/ORBmatcher.cc:290-403.  Scw = scwScale * identity.
// tmatched[k]: -1 = vpMatched[k] NULL, -2 = a map point that is not among vpPoints, i >= 0 = vpPoints[i] (so that point
// counts as already found).  qbad[i]: isBad().  matchT[k] = index of the map point in vpMatched[k] afterwards.
int mref_search_by_projection_sim3kf(int nT, const float* tx, const float* ty, const int* toct, const uint8_t* tdesc,
                                     const int* tmatched, int minX, int maxX, int minY, int maxY, float invW, float invH,
                                     int nQ, const uint8_t* qbad, const float* wx, const float* wy, const float* wz,
                                     const float* maxDist, const float* minDist, const float* normal, const uint8_t* qdesc,
                                     const float* scaleFactors, int nLevels, float logScaleFactor, float scwScale, int th,
                                     int* matchT) {
    Pool pool;
    ORBmatcher matcher(0.75f, true);
    KeyFrame kf;
    setup_kf(kf, nT, tx, ty, toct, tdesc, nullptr, minX, maxX, minY, maxY, invW, invH, scaleFactors, nullptr, nLevels, logScaleFactor);
    std::vector<MapPoint*> pts(nQ);
    for (int i = 0; i < nQ; ++i) {
        pts[i] = pool.make(i);
        pts[i]->bad = qbad && qbad[i];
        setup_point(pts[i], i, wx, wy, wz, maxDist, minDist, normal, qdesc);
    }
    std::vector<MapPoint*> matched(nT, nullptr);
    for (int k = 0; k < nT; ++k) {
        if (tmatched[k] >= 0) matched[k] = pts[tmatched[k]];
        else if (tmatched[k] == -2) matched[k] = pool.make(-2);
    }
    cv::Mat Scw = eye4();
    for (int i = 0; i < 3; ++i) Scw.at<float>(i, i) = scwScale;
    const int n = matcher.SearchByProjection(&kf, Scw, pts, matched, th);
    for (int k = 0; k < nT; ++k) matchT[k] = pool.of(matched[k]);
    return n;
}

// Fuse(KeyFrame*, const vector<MapPoint*>&, th)  src/ORBmatcher.cc:825-961.  Keyframe features: slotState 0 = no map
// point, 1 = a good one, 2 = a bad one (Observations() = slotObs).  Candidates: qstate 0 = NULL, 1 = bad, 2 = already in
// the keyframe, 3 = usable; qid[i] = canonical index when the same MapPoint* appears more than once (qid[i] <= i).
// Outputs (stand-in MapPoint semantics of matchshim): per candidate addedAt (AddObservation index or -1), replacedBy (tag
// of the point it was replaced by, -1 none); per keyframe feature slotReplacedBy (same for the point held on entry) and
// slotHolder (tag of the point held afterwards).  Tags: candidate i -> qid, entry point of feature k -> 100000+k.
int mref_fuse(int nT, const float* tx, const float* ty, const int* toct, const uint8_t* tdesc, const float* turight,
              const uint8_t* slotState, const int* slotObs, int minX, int maxX, int minY, int maxY, float invW, float invH,
              const float* invSigma2, const float* scaleFactors, int nLevels, float logScaleFactor, float mbf, int nQ,
              const uint8_t* qstate, const int* qid, const float* wx, const float* wy, const float* wz, const float* maxDist,
              const float* minDist, const float* normal, const int* qobs, const uint8_t* qdesc, float th, int* addedAt,
              int* replacedBy, int* slotReplacedBy, int* slotHolder) {
    Pool pool;
    ORBmatcher matcher(0.6f, true);
    KeyFrame kf, other;
    setup_kf(kf, nT, tx, ty, toct, tdesc, turight, minX, maxX, minY, maxY, invW, invH, scaleFactors, invSigma2, nLevels, logScaleFactor);
    kf.mbf = mbf;
    std::vector<MapPoint*> entry(nT, nullptr);
    for (int k = 0; k < nT; ++k) {
        if (!slotState[k]) continue;
        entry[k] = pool.make(SLOT_TAG + k);
        entry[k]->bad = slotState[k] == 2;
        entry[k]->nObs = slotObs[k];
        entry[k]->obs[&kf] = k;
        kf.mvpMapPoints[k] = entry[k];
    }
    std::vector<MapPoint*> pts(nQ, nullptr);
    for (int i = 0; i < nQ; ++i) {
        if (qstate[i] == 0) continue;
        if (qid && qid[i] != i) { pts[i] = pts[qid[i]]; continue; }
        pts[i] = pool.make(i);
        pts[i]->bad = qstate[i] == 1;
        pts[i]->nObs = qobs[i];
        if (qstate[i] == 2) pts[i]->obs[&kf] = 0;
        else pts[i]->obs[&other] = 0;
        setup_point(pts[i], i, wx, wy, wz, maxDist, minDist, normal, qdesc);
    }
    const int n = matcher.Fuse(&kf, pts, th);
    for (int i = 0; i < nQ; ++i) {
        addedAt[i] = -1;
        replacedBy[i] = -1;
        if (!pts[i]) continue;
        if (qstate[i] != 2 && pts[i]->obs.count(&kf)) addedAt[i] = (int)pts[i]->obs[&kf];
        replacedBy[i] = pool.of(pts[i]->replaced);
    }
    for (int k = 0; k < nT; ++k) {
        slotReplacedBy[k] = entry[k] ? pool.of(entry[k]->replaced) : -1;
        slotHolder[k] = pool.of(kf.mvpMapPoints[k]);
    }
    return n;
}

// Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)  src/ORBmatcher.cc:963-1100.  slotState as mref_fuse, except
// slotQuery[k] >= 0: the feature holds vpPoints[slotQuery[k]] (so that candidate is "already found").  replacePoint[i] =
// tag of vpReplacePoint[i] afterwards.
int mref_fuse_sim3(int nT, const float* tx, const float* ty, const int* toct, const uint8_t* tdesc, const uint8_t* slotState,
                   const int* slotQuery, int minX, int maxX, int minY, int maxY, float invW, float invH,
                   const float* scaleFactors, int nLevels, float logScaleFactor, float scwScale, int nQ, const uint8_t* qbad,
                   const float* wx, const float* wy, const float* wz, const float* maxDist, const float* minDist,
                   const float* normal, const uint8_t* qdesc, float th, int* addedAt, int* replacePoint, int* slotHolder) {
    Pool pool;
    ORBmatcher matcher(0.8f, true);
    KeyFrame kf;
    setup_kf(kf, nT, tx, ty, toct, tdesc, nullptr, minX, maxX, minY, maxY, invW, invH, scaleFactors, nullptr, nLevels, logScaleFactor);
    std::vector<MapPoint*> pts(nQ);
    for (int i = 0; i < nQ; ++i) {
        pts[i] = pool.make(i);
        pts[i]->bad = qbad && qbad[i];
        setup_point(pts[i], i, wx, wy, wz, maxDist, minDist, normal, qdesc);
    }
    for (int k = 0; k < nT; ++k) {
        if (slotQuery && slotQuery[k] >= 0) { kf.mvpMapPoints[k] = pts[slotQuery[k]]; continue; }
        if (!slotState[k]) continue;
        kf.mvpMapPoints[k] = pool.make(SLOT_TAG + k);
        kf.mvpMapPoints[k]->bad = slotState[k] == 2;
    }
    cv::Mat Scw = eye4();
    for (int i = 0; i < 3; ++i) Scw.at<float>(i, i) = scwScale;
    std::vector<MapPoint*> repl(nQ, nullptr);
    const int n = matcher.Fuse(&kf, Scw, pts, th, repl);
    for (int i = 0; i < nQ; ++i) {
        replacePoint[i] = pool.of(repl[i]);
        addedAt[i] = pts[i]->obs.count(&kf) ? (int)pts[i]->obs[&kf] : -1;
    }
    for (int k = 0; k < nT; ++k) slotHolder[k] = pool.of(kf.mvpMapPoints[k]);
    return n;
}

// SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)  src/ORBmatcher.cc:1102-1326 with R12 = I, t12 = 0 and both
// keyframe poses identity.  Per keyframe a (a = 1, 2): features (x, y, octave, descriptor) and per feature the map point
// it holds: state 0 = none, 1 = bad, 3 = good, with world point / distance range / descriptor.  pre12[i1] = index of the
// KF2 feature whose map point vpMatches12[i1] holds on entry (-1 none).  match12[i1] = KF2 feature index of
// vpMatches12[i1] afterwards.
int mref_search_by_sim3(int n1, const float* x1, const float* y1, const int* oct1, const uint8_t* kdesc1,
                        const uint8_t* state1, const float* wx1, const float* wy1, const float* wz1, const float* maxDist1,
                        const float* minDist1, const uint8_t* pdesc1, int n2, const float* x2, const float* y2,
                        const int* oct2, const uint8_t* kdesc2, const uint8_t* state2, const float* wx2, const float* wy2,
                        const float* wz2, const float* maxDist2, const float* minDist2, const uint8_t* pdesc2,
                        const int* pre12, int minX, int maxX, int minY, int maxY, float invW, float invH,
                        const float* scaleFactors, int nLevels, float logScaleFactor, float s12, float th, int* match12) {
    Pool pool;
    ORBmatcher matcher(0.75f, true);
    KeyFrame k1, k2;
    setup_kf(k1, n1, x1, y1, oct1, kdesc1, nullptr, minX, maxX, minY, maxY, invW, invH, scaleFactors, nullptr, nLevels, logScaleFactor);
    setup_kf(k2, n2, x2, y2, oct2, kdesc2, nullptr, minX, maxX, minY, maxY, invW, invH, scaleFactors, nullptr, nLevels, logScaleFactor);
    for (int i = 0; i < n1; ++i) {
        if (!state1[i]) continue;
        MapPoint* p = pool.make(i);
        p->bad = state1[i] == 1;
        setup_point(p, i, wx1, wy1, wz1, maxDist1, minDist1, nullptr, pdesc1);
        p->obs[&k1] = i;
        k1.mvpMapPoints[i] = p;
    }
    for (int i = 0; i < n2; ++i) {
        if (!state2[i]) continue;
        MapPoint* p = pool.make(SLOT_TAG + i);
        p->bad = state2[i] == 1;
        setup_point(p, i, wx2, wy2, wz2, maxDist2, minDist2, nullptr, pdesc2);
        p->obs[&k2] = i;
        k2.mvpMapPoints[i] = p;
    }
    std::vector<MapPoint*> m12(n1, nullptr);
    for (int i = 0; i < n1; ++i)
        if (pre12 && pre12[i] >= 0) m12[i] = k2.mvpMapPoints[pre12[i]];
    cv::Mat R12 = eye4().rowRange(0, 3).colRange(0, 3).clone();
    cv::Mat t12 = vec3(0.f, 0.f, 0.f);
    const int n = matcher.SearchBySim3(&k1, &k2, m12, s12, R12, t12, th);
    for (int i = 0; i < n1; ++i) {
        const int t = pool.of(m12[i]);
        match12[i] = t >= SLOT_TAG ? t - SLOT_TAG : -1;
    }
    return n;
}

}  // extern "C"
