/**
 * Drop-in implementation of the Hamming-search methods of ORB_SLAM2::ORBmatcher for EAO-Fusion, backed by
 * libeaof_orb.so (B200, sm_100a) through the C ABI of include/eaof_match.h.
 *
 * It is compiled against the reference's OWN header include/ORBmatcher.h (class declaration unchanged) and the
 * reference's Frame / KeyFrame / MapPoint, and replaces these definitions of src/ORBmatcher.cc:
 *   ORBmatcher::ORBmatcher, TH_LOW/TH_HIGH/HISTO_LENGTH (:37-42), DescriptorDistance (:1649-1665), RadiusByViewingCos (:131-137),
 *   SearchByProjection(Frame&, const vector<MapPoint*>&, th)            (:45-129)
 *   SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)                   (:159-288)
 *   SearchForInitialization(...)                                         (:405-520)
 *   SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)                (:522-655)
 *   SearchForTriangulation(...)                                          (:657-823)
 *   SearchByProjection(Frame&, const Frame&, th, bMono)                  (:1328-1472)
 *   SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, d)  (:1474-1601)
 *   SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)          (:290-403)
 *   Fuse(KeyFrame*, const vector<MapPoint*>&, th)                        (:825-961)
 *   Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)                   (:963-1100)
 *   SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)             (:1102-1326)
 *   CheckDistEpipolarLine (:140-157) and ComputeThreeMaxima (:1603-1644): the two protected helpers.  Their work happens
 *   inside the kernels (eaof_match_triangulation's epipolar gate, the rotation-histogram pruning of every matcher); the
 *   host definitions at the end of this file exist so that the class declared in include/ORBmatcher.h links completely.
 * i.e. every definition of the reference's ORBmatcher.cc: a maintainer replaces that file by this one in the build
 * (INTEGRATION.md §3).  Every method marshals the fields the reference loop reads into plain arrays, calls the
 * library, and applies the result to the host objects exactly where the reference does (the map mutations of Fuse run
 * on the host, in candidate order, with the skip tests re-read at each turn); projections
 * (Rcw*x3Dw+tcw, PredictScale, ...) stay on the host with the reference's own expressions.  There is no CPU search path:
 * a library failure throws.
 *
 * tests/cpp builds this file against oracle/matchshim (array-backed stand-ins with the same member names) and runs it
 * through the same harness as the unmodified reference translation unit; results must be identical.
 */
#include "ORBmatcher.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "FeatureVector.h"
#include "eaof_match.h"

using namespace std;

namespace ORB_SLAM2
{

const int ORBmatcher::TH_HIGH = EAOF_TH_HIGH;
const int ORBmatcher::TH_LOW = EAOF_TH_LOW;
const int ORBmatcher::HISTO_LENGTH = EAOF_HISTO_LENGTH;

namespace
{

void Throw(const char* what)
{
    std::string msg = std::string("ORBmatcher(eaof): ") + what + ": " + eaof_last_error();
    fprintf(stderr, "%s\n", msg.c_str());
    throw std::runtime_error(msg);  // no CPU fallback, by contract
}

// One matcher handle per calling thread (Tracking, LocalMapping and LoopClosing call concurrently), grown on demand.
struct TlsMatcher
{
    eaof_matcher* h;
    int cap;
    TlsMatcher() : h(NULL), cap(0) {}
    ~TlsMatcher() { if(h) eaof_matcher_destroy(h); }
};

// Map-side searches (local map in SearchByProjection(F, vpMapPoints), loop map points in Fuse / SearchByProjection(KF, Scw))
// can hold far more points than a matcher workspace row (65535): invalid queries are dropped before upload and the valid
// ones go through the library in chunks of this many, in order.
const size_t kQueryChunk = 32768;

eaof_matcher* Matcher(size_t nFeatures)
{
    static thread_local TlsMatcher tls;
    if(!tls.h || (int)nFeatures > tls.cap)
    {
        if(tls.h)
        {
            eaof_matcher_destroy(tls.h);
            tls.h = NULL;
        }
        if(nFeatures > 65535)  // callers chunk their query lists (kQueryChunk); frame features never get near this
            Throw("more than 65535 features in one frame or query chunk");
        int cap = 4096;
        while(cap < (int)nFeatures)
            cap *= 2;
        if(cap > 65535)
            cap = 65535;
        const char* dev = getenv("EAOF_DEVICE");
        if(eaof_matcher_create(dev && *dev ? atoi(dev) : 0, 1, cap, &tls.h) != EAOF_OK)
            Throw("eaof_matcher_create");
        tls.cap = cap;
    }
    return tls.h;
}

// DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>) -> CSR in ascending node order
struct Csr
{
    vector<int> id, start, idx;
    explicit Csr(const DBoW2::FeatureVector& fv)
    {
        start.push_back(0);
        for(DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        {
            id.push_back((int)it->first);
            for(size_t i = 0; i < it->second.size(); i++)
                idx.push_back((int)it->second[i]);
            start.push_back((int)idx.size());
        }
    }
    int n() const { return (int)id.size(); }
};

struct KeyArrays
{
    vector<float> x, y, angle;
    vector<int> octave;
    explicit KeyArrays(const vector<cv::KeyPoint>& k) : x(k.size()), y(k.size()), angle(k.size()), octave(k.size())
    {
        for(size_t i = 0; i < k.size(); i++)
        {
            x[i] = k[i].pt.x;
            y[i] = k[i].pt.y;
            angle[i] = k[i].angle;
            octave[i] = k[i].octave;
        }
    }
};

vector<unsigned char> Rows(const cv::Mat& desc, size_t n)
{
    vector<unsigned char> out(32 * (n ? n : 1));
    for(size_t i = 0; i < n; i++)
        memcpy(&out[32 * i], desc.ptr<unsigned char>((int)i), 32);
    return out;
}

template <typename T> const T* P(const vector<T>& v) { return v.empty() ? NULL : &v[0]; }
template <typename T> T* P(vector<T>& v) { return v.empty() ? NULL : &v[0]; }

}  // namespace

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri)
{
}

// One pair on the host is cheaper than a launch; the batched form is eaof_hamming_distances.
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b)
{
    const unsigned int* pa = a.ptr<unsigned int>();
    const unsigned int* pb = b.ptr<unsigned int>();
    int dist = 0;
    for(int i = 0; i < 8; i++)
        dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

float ORBmatcher::RadiusByViewingCos(const float& viewCos)
{
    return viewCos > 0.998 ? 2.5 : 4.0;
}

int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th)
{
    const size_t nAll = vpMapPoints.size(), nF = (size_t)F.N;
    if(nAll == 0 || nF == 0)
        return 0;
    const bool bFactor = th != 1.0;
    // only the points the reference's loop does not skip (:54-58) are uploaded, in their order
    vector<size_t> orig;
    orig.reserve(nAll);
    for(size_t i = 0; i < nAll; i++)
        if(vpMapPoints[i]->mbTrackInView && !vpMapPoints[i]->isBad())
            orig.push_back(i);
    const size_t nMP = orig.size();
    if(nMP == 0)
        return 0;
    vector<unsigned char> obs(nMP, 0), qdesc(32 * nMP);
    vector<float> u(nMP, 0.f), v(nMP, 0.f), radius(nMP, 0.f), ur(nMP, 0.f);
    vector<int> minL(nMP, 0), maxL(nMP, 0);
    for(size_t j = 0; j < nMP; j++)
    {
        MapPoint* pMP = vpMapPoints[orig[j]];
        const int nPredictedLevel = pMP->mnTrackScaleLevel;
        float r = RadiusByViewingCos(pMP->mTrackViewCos);
        if(bFactor)
            r *= th;
        u[j] = pMP->mTrackProjX;
        v[j] = pMP->mTrackProjY;
        ur[j] = pMP->mTrackProjXR;
        radius[j] = r * F.mvScaleFactors[nPredictedLevel];
        minL[j] = nPredictedLevel - 1;
        maxL[j] = nPredictedLevel;
        obs[j] = pMP->Observations() > 0;
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&qdesc[32 * j], d.ptr<unsigned char>(), 32);
    }
    KeyArrays keys(F.mvKeysUn);
    vector<unsigned char> tdesc = Rows(F.mDescriptors, nF), taken(nF, 0);
    vector<int> match(nF, -1);
    int nTotal = 0;
    // the greedy exclusion only looks at what earlier points left in F.mvpMapPoints (:78-80), so a chunk of later points
    // sees exactly the state the earlier chunks produced
    for(size_t c0 = 0; c0 < nMP; c0 += kQueryChunk)
    {
        const size_t nc = min(kQueryChunk, nMP - c0);
        for(size_t k = 0; k < nF; k++)
            taken[k] = F.mvpMapPoints[k] && F.mvpMapPoints[k]->Observations() > 0;
        int n = 0;
        if(eaof_match_windows(Matcher(max(nF, nc)), EAOF_WIN_RATIO_SAME_LEVEL, (int)nF, P(keys.x), P(keys.y), P(keys.octave), NULL,
                              P(tdesc), P(F.mvuRight), P(taken), F.mnMinX, F.mnMaxX, F.mnMinY, F.mnMaxY,
                              F.mfGridElementWidthInv, F.mfGridElementHeightInv, (int)nc, NULL, &u[c0], &v[c0], &radius[c0],
                              &minL[c0], &maxL[c0], &ur[c0], NULL, &qdesc[32 * c0], &obs[c0], TH_HIGH, mfNNratio, 0, 0, P(match),
                              NULL, &n) != EAOF_OK)
            Throw("eaof_match_windows");
        for(size_t k = 0; k < nF; k++)
            if(match[k] >= 0)
                F.mvpMapPoints[k] = vpMapPoints[orig[c0 + match[k]]];
        nTotal += n;
    }
    return nTotal;
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
{
    const vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    const size_t nQ = vpMapPointsKF.size(), nT = (size_t)F.N;
    if(nQ == 0 || nT == 0)
        return 0;
    vector<unsigned char> validQ(nQ, 0);
    for(size_t i = 0; i < nQ; i++)
        validQ[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();
    KeyArrays kq(pKF->mvKeysUn), kt(F.mvKeys);
    Csr fq(pKF->mFeatVec), ft(F.mFeatVec);
    vector<unsigned char> dq = Rows(pKF->mDescriptors, nQ), dt = Rows(F.mDescriptors, nT);
    vector<int> match(nT, -1);
    int n = 0;
    if(eaof_match_bow(Matcher(max(nQ, nT)), EAOF_BOW_KF_FRAME, mfNNratio, mbCheckOrientation, (int)nQ, P(dq), P(kq.angle),
                      P(validQ), (int)nT, P(dt), P(kt.angle), NULL, fq.n(), P(fq.id), P(fq.start), P(fq.idx), ft.n(),
                      P(ft.id), P(ft.start), P(ft.idx), P(match), NULL, &n) != EAOF_OK)
        Throw("eaof_match_bow");
    for(size_t t = 0; t < nT; t++)
        if(match[t] >= 0)
            vpMapPointMatches[t] = vpMapPointsKF[match[t]];
    return n;
}

int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12,
                                        int windowSize)
{
    const size_t n1 = F1.mvKeysUn.size(), n2 = F2.mvKeysUn.size();
    vnMatches12 = vector<int>(n1, -1);
    if(n1 == 0 || n2 == 0)
        return 0;
    KeyArrays k1(F1.mvKeysUn), k2(F2.mvKeysUn);
    vector<unsigned char> d1 = Rows(F1.mDescriptors, n1), d2 = Rows(F2.mDescriptors, n2);
    vector<float> prev(2 * n1);
    for(size_t i = 0; i < n1; i++)
    {
        prev[2 * i] = vbPrevMatched[i].x;
        prev[2 * i + 1] = vbPrevMatched[i].y;
    }
    int n = 0;
    if(eaof_match_initialization(Matcher(max(n1, n2)), mfNNratio, mbCheckOrientation, (int)n1, P(k1.octave), P(k1.angle),
                                 P(d1), P(prev), (int)n2, P(k2.x), P(k2.y), P(k2.octave), P(k2.angle), P(d2), F2.mnMinX,
                                 F2.mnMaxX, F2.mnMinY, F2.mnMaxY, F2.mfGridElementWidthInv, F2.mfGridElementHeightInv,
                                 windowSize, P(vnMatches12), &n) != EAOF_OK)
        Throw("eaof_match_initialization");
    for(size_t i = 0; i < n1; i++)
        if(vnMatches12[i] >= 0)
            vbPrevMatched[i] = F2.mvKeysUn[vnMatches12[i]].pt;
    return n;
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
{
    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    const size_t nQ = vpMapPoints1.size(), nT = vpMapPoints2.size();
    if(nQ == 0 || nT == 0)
        return 0;
    vector<unsigned char> validQ(nQ, 0), validT(nT, 0);
    for(size_t i = 0; i < nQ; i++)
        validQ[i] = vpMapPoints1[i] && !vpMapPoints1[i]->isBad();
    for(size_t i = 0; i < nT; i++)
        validT[i] = vpMapPoints2[i] && !vpMapPoints2[i]->isBad();
    KeyArrays kq(pKF1->mvKeysUn), kt(pKF2->mvKeysUn);
    Csr fq(pKF1->mFeatVec), ft(pKF2->mFeatVec);
    vector<unsigned char> dq = Rows(pKF1->mDescriptors, nQ), dt = Rows(pKF2->mDescriptors, nT);
    vector<int> match(nQ, -1);
    int n = 0;
    if(eaof_match_bow(Matcher(max(nQ, nT)), EAOF_BOW_KF_KF, mfNNratio, mbCheckOrientation, (int)nQ, P(dq), P(kq.angle),
                      P(validQ), (int)nT, P(dt), P(kt.angle), P(validT), fq.n(), P(fq.id), P(fq.start), P(fq.idx), ft.n(),
                      P(ft.id), P(ft.start), P(ft.idx), P(match), NULL, &n) != EAOF_OK)
        Throw("eaof_match_bow");
    for(size_t q = 0; q < nQ; q++)
        if(match[q] >= 0)
            vpMatches12[q] = vpMapPoints2[match[q]];
    return n;
}

int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
                                       vector<pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo)
{
    // epipole in the second image, :663-670
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w*Cw+t2w;
    const float invz = 1.0f/C2.at<float>(2);
    const float ex = pKF2->fx*C2.at<float>(0)*invz+pKF2->cx;
    const float ey = pKF2->fy*C2.at<float>(1)*invz+pKF2->cy;

    vMatchedPairs.clear();
    const size_t n1 = (size_t)pKF1->N, n2 = (size_t)pKF2->N;
    if(n1 == 0 || n2 == 0)
        return 0;
    vector<unsigned char> free1(n1), stereo1(n1), free2(n2), stereo2(n2);
    for(size_t i = 0; i < n1; i++)
    {
        free1[i] = pKF1->GetMapPoint(i) == NULL;
        stereo1[i] = pKF1->mvuRight[i] >= 0;
    }
    for(size_t i = 0; i < n2; i++)
    {
        free2[i] = pKF2->GetMapPoint(i) == NULL;
        stereo2[i] = pKF2->mvuRight[i] >= 0;
    }
    KeyArrays k1(pKF1->mvKeysUn), k2(pKF2->mvKeysUn);
    Csr f1(pKF1->mFeatVec), f2(pKF2->mFeatVec);
    vector<unsigned char> d1 = Rows(pKF1->mDescriptors, n1), d2 = Rows(pKF2->mDescriptors, n2);
    float F[9];
    for(int i = 0; i < 9; i++)
        F[i] = F12.at<float>(i / 3, i % 3);
    vector<int> match(n1, -1);
    int n = 0;
    if(eaof_match_triangulation(Matcher(max(n1, n2)), mbCheckOrientation, bOnlyStereo, (int)n1, P(d1), P(k1.x), P(k1.y),
                                P(k1.angle), P(free1), P(stereo1), (int)n2, P(d2), P(k2.x), P(k2.y), P(k2.octave),
                                P(k2.angle), P(free2), P(stereo2), f1.n(), P(f1.id), P(f1.start), P(f1.idx), f2.n(),
                                P(f2.id), P(f2.start), P(f2.idx), F, ex, ey, P(pKF2->mvScaleFactors),
                                P(pKF2->mvLevelSigma2), (int)pKF2->mvScaleFactors.size(), P(match), NULL, &n) != EAOF_OK)
        Throw("eaof_match_triangulation");
    vMatchedPairs.reserve(n > 0 ? n : 0);
    for(size_t i = 0; i < n1; i++)
        if(match[i] >= 0)
            vMatchedPairs.push_back(make_pair(i, (size_t)match[i]));
    return n;
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
{
    // relative pose, bForward / bBackward: :1340-1351
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0,3).colRange(0,3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0,3).col(3);
    const cv::Mat twc = -Rcw.t()*tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0,3).colRange(0,3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0,3).col(3);
    const cv::Mat tlc = Rlw*twc+tlw;
    const bool bForward = tlc.at<float>(2)>CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc.at<float>(2)>CurrentFrame.mb && !bMono;

    const size_t nL = (size_t)LastFrame.N, nC = (size_t)CurrentFrame.N;
    if(nL == 0 || nC == 0)
        return 0;
    vector<unsigned char> lvalid(nL, 0), lobs(nL, 0), ldesc(32 * nL);
    vector<float> lu(nL, 0.f), lv(nL, 0.f), linvz(nL, 1.f), langle(nL, 0.f);
    vector<int> loct(nL, 0);
    for(size_t i = 0; i < nL; i++)
    {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if(!pMP || LastFrame.mvbOutlier[i])
            continue;
        // project, :1362-1371
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw*x3Dw+tcw;
        const float xc = x3Dc.at<float>(0);
        const float yc = x3Dc.at<float>(1);
        const float invzc = 1.0/x3Dc.at<float>(2);
        lvalid[i] = 1;
        linvz[i] = invzc;
        lu[i] = CurrentFrame.fx*xc*invzc+CurrentFrame.cx;
        lv[i] = CurrentFrame.fy*yc*invzc+CurrentFrame.cy;
        loct[i] = LastFrame.mvKeys[i].octave;
        langle[i] = LastFrame.mvKeysUn[i].angle;
        lobs[i] = pMP->Observations() > 0;
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&ldesc[32 * i], d.ptr<unsigned char>(), 32);
    }
    KeyArrays kc(CurrentFrame.mvKeysUn);
    vector<unsigned char> cdesc = Rows(CurrentFrame.mDescriptors, nC), taken(nC, 0);
    for(size_t k = 0; k < nC; k++)
        taken[k] = CurrentFrame.mvpMapPoints[k] && CurrentFrame.mvpMapPoints[k]->Observations() > 0;
    vector<int> match(nC, -1), dist(nC, -1);
    int n = 0;
    if(eaof_match_projection(Matcher(max(nL, nC)), (int)nC, P(kc.x), P(kc.y), P(kc.octave), P(kc.angle), P(cdesc),
                             P(CurrentFrame.mvuRight), P(taken), CurrentFrame.mnMinX, CurrentFrame.mnMaxX,
                             CurrentFrame.mnMinY, CurrentFrame.mnMaxY, CurrentFrame.mfGridElementWidthInv,
                             CurrentFrame.mfGridElementHeightInv, (int)nL, P(lvalid), P(lu), P(lv), P(linvz), P(loct),
                             P(langle), P(ldesc), P(lobs), P(CurrentFrame.mvScaleFactors),
                             (int)CurrentFrame.mvScaleFactors.size(), th, CurrentFrame.mbf,
                             bForward ? 1 : (bBackward ? 2 : 0), mbCheckOrientation, P(match), P(dist), &n) != EAOF_OK)
        Throw("eaof_match_projection");
    for(size_t k = 0; k < nC; k++)
    {
        if(match[k] >= 0)
            CurrentFrame.mvpMapPoints[k] = LastFrame.mvpMapPoints[match[k]];
        else if(dist[k] == -2)  // matched, then removed by the rotation check (:1463)
            CurrentFrame.mvpMapPoints[k] = static_cast<MapPoint*>(NULL);
    }
    return n;
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, const float th,
                                   const int ORBdist)
{
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0,3).colRange(0,3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0,3).col(3);
    const cv::Mat Ow = -Rcw.t()*tcw;

    const vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const size_t nK = vpMPs.size(), nC = (size_t)CurrentFrame.N;
    if(nK == 0 || nC == 0)
        return 0;
    vector<unsigned char> valid(nK, 0), qdesc(32 * nK);
    vector<float> u(nK, 0.f), v(nK, 0.f), radius(nK, 0.f), angle(nK, 0.f);
    vector<int> minL(nK, 0), maxL(nK, 0);
    for(size_t i = 0; i < nK; i++)
    {
        MapPoint* pMP = vpMPs[i];
        if(!pMP || pMP->isBad() || sAlreadyFound.count(pMP))
            continue;
        // project and predict the scale level, :1497-1525
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw*x3Dw+tcw;
        const float xc = x3Dc.at<float>(0);
        const float yc = x3Dc.at<float>(1);
        const float invzc = 1.0/x3Dc.at<float>(2);
        u[i] = CurrentFrame.fx*xc*invzc+CurrentFrame.cx;
        v[i] = CurrentFrame.fy*yc*invzc+CurrentFrame.cy;
        if(u[i]<CurrentFrame.mnMinX || u[i]>CurrentFrame.mnMaxX)
            continue;
        if(v[i]<CurrentFrame.mnMinY || v[i]>CurrentFrame.mnMaxY)
            continue;
        cv::Mat PO = x3Dw-Ow;
        float dist3D = cv::norm(PO);
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        if(dist3D<minDistance || dist3D>maxDistance)
            continue;
        int nPredictedLevel = pMP->PredictScale(dist3D,CurrentFrame.mfLogScaleFactor);
        valid[i] = 1;
        radius[i] = th*CurrentFrame.mvScaleFactors[nPredictedLevel];
        minL[i] = nPredictedLevel-1;
        maxL[i] = nPredictedLevel+1;
        angle[i] = pKF->mvKeysUn[i].angle;
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&qdesc[32 * i], d.ptr<unsigned char>(), 32);
    }
    KeyArrays kc(CurrentFrame.mvKeysUn);
    vector<unsigned char> cdesc = Rows(CurrentFrame.mDescriptors, nC), taken(nC, 0);
    for(size_t k = 0; k < nC; k++)
        taken[k] = CurrentFrame.mvpMapPoints[k] != NULL;
    vector<int> match(nC, -1);
    int n = 0;
    if(eaof_match_windows(Matcher(max(nK, nC)), EAOF_WIN_BEST, (int)nC, P(kc.x), P(kc.y), P(kc.octave), P(kc.angle), P(cdesc),
                          NULL, P(taken), CurrentFrame.mnMinX, CurrentFrame.mnMaxX, CurrentFrame.mnMinY, CurrentFrame.mnMaxY,
                          CurrentFrame.mfGridElementWidthInv, CurrentFrame.mfGridElementHeightInv, (int)nK, P(valid), P(u),
                          P(v), P(radius), P(minL), P(maxL), NULL, P(angle), P(qdesc), NULL, ORBdist, mfNNratio,
                          mbCheckOrientation ? 1 : 0, 1, P(match), NULL, &n) != EAOF_OK)
        Throw("eaof_match_windows");
    for(size_t k = 0; k < nC; k++)
        if(match[k] >= 0)
            CurrentFrame.mvpMapPoints[k] = vpMPs[match[k]];
    return n;
}

namespace
{

// Candidate map points seen from a keyframe: what the map-side loops derive per point before the descriptor search
// (:318-360, :852-887, :1001-1047, :1160-1194): projected pixel, predicted level, search radius.
struct Projected
{
    vector<unsigned char> valid, desc;
    vector<float> u, v, ur, radius;
    vector<int> minL, maxL;
    explicit Projected(size_t n) : valid(n, 0), desc(32 * (n ? n : 1)), u(n, 0.f), v(n, 0.f), ur(n, 0.f), radius(n, 0.f), minL(n, 0), maxL(n, 0) {}
    void Set(size_t i, float u_, float v_, float ur_, float r, int level, MapPoint* pMP)
    {
        valid[i] = 1;
        u[i] = u_;
        v[i] = v_;
        ur[i] = ur_;
        radius[i] = r;
        minL[i] = level - 1;
        maxL[i] = level;
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&desc[32 * i], d.ptr<unsigned char>(), 32);
    }
    // the valid entries only, in order (orig[j] = index in this list): what goes to the library
    void Compact(Projected& out, vector<size_t>& orig) const
    {
        orig.clear();
        for(size_t i = 0; i < valid.size(); i++)
            if(valid[i])
                orig.push_back(i);
        const size_t n = orig.size();
        out = Projected(n);
        for(size_t j = 0; j < n; j++)
        {
            const size_t i = orig[j];
            out.valid[j] = 1;
            out.u[j] = u[i]; out.v[j] = v[i]; out.ur[j] = ur[i]; out.radius[j] = radius[i];
            out.minL[j] = minL[i]; out.maxL[j] = maxL[i];
            memcpy(&out.desc[32 * j], &desc[32 * i], 32);
        }
    }
};

// Points given in the world frame, keyframe pose (Rcw, tcw, Ow): shared by SearchByProjection(KF,Scw), Fuse and Fuse(Scw)
bool ProjectWorld(KeyFrame* pKF, MapPoint* pMP, const cv::Mat& Rcw, const cv::Mat& tcw, const cv::Mat& Ow, float th, float bf,
                  float& u, float& v, float& ur, float& radius, int& level)
{
    cv::Mat p3Dw = pMP->GetWorldPos();
    cv::Mat p3Dc = Rcw*p3Dw+tcw;
    if(p3Dc.at<float>(2)<0.0f)
        return false;
    const float invz = 1/p3Dc.at<float>(2);
    const float x = p3Dc.at<float>(0)*invz;
    const float y = p3Dc.at<float>(1)*invz;
    u = pKF->fx*x+pKF->cx;
    v = pKF->fy*y+pKF->cy;
    if(!pKF->IsInImage(u,v))
        return false;
    ur = u-bf*invz;
    const float maxDistance = pMP->GetMaxDistanceInvariance();
    const float minDistance = pMP->GetMinDistanceInvariance();
    cv::Mat PO = p3Dw-Ow;
    const float dist3D = cv::norm(PO);
    if(dist3D<minDistance || dist3D>maxDistance)
        return false;
    cv::Mat Pn = pMP->GetNormal();
    if(PO.dot(Pn)<0.5*dist3D)
        return false;
    level = pMP->PredictScale(dist3D,pKF->mfLogScaleFactor);
    radius = th*pKF->mvScaleFactors[level];
    return true;
}

struct KeyFrameArrays
{
    KeyArrays keys;
    vector<unsigned char> desc;
    explicit KeyFrameArrays(KeyFrame* pKF) : keys(pKF->mvKeysUn), desc(Rows(pKF->mDescriptors, pKF->mvKeysUn.size())) {}
};

// best keyframe feature per projected point, no exclusion between points: the valid points are compacted and go through
// the library in chunks (a loop-closure map can hold more points than one workspace row)
void SearchIndependent(KeyFrame* pKF, const Projected& qAll, int gate, int thAccept, vector<int>& match)
{
    const size_t nT = pKF->mvKeysUn.size();
    match.assign(qAll.valid.size(), -1);
    if(nT == 0 || qAll.valid.empty())
        return;
    Projected q(0);
    vector<size_t> orig;
    qAll.Compact(q, orig);
    const size_t nQ = orig.size();
    if(nQ == 0)
        return;
    KeyFrameArrays t(pKF);
    const bool chi2 = gate == EAOF_GATE_FUSE_CHI2;
    vector<int> part(min(nQ, kQueryChunk), -1);
    for(size_t c0 = 0; c0 < nQ; c0 += kQueryChunk)
    {
        const size_t nc = min(kQueryChunk, nQ - c0);
        int n = 0;
        if(eaof_match_windows_independent(Matcher(max(nT, nc)), gate, (int)nT, P(t.keys.x), P(t.keys.y), P(t.keys.octave), P(t.desc),
                                          chi2 ? P(pKF->mvuRight) : NULL, pKF->mnMinX, pKF->mnMinY, pKF->mfGridElementWidthInv,
                                          pKF->mfGridElementHeightInv, chi2 ? P(pKF->mvInvLevelSigma2) : NULL,
                                          (int)pKF->mvScaleFactors.size(), (int)nc, NULL, &q.u[c0], &q.v[c0], &q.radius[c0],
                                          &q.minL[c0], &q.maxL[c0], &q.ur[c0], &q.desc[32 * c0], thAccept, P(part), NULL,
                                          &n) != EAOF_OK)
            Throw("eaof_match_windows_independent");
        for(size_t j = 0; j < nc; j++)
            match[orig[c0 + j]] = part[j];
    }
}

void DecomposeSim3(const cv::Mat& Scw, cv::Mat& Rcw, cv::Mat& tcw, cv::Mat& Ow)
{
    cv::Mat sRcw = Scw.rowRange(0,3).colRange(0,3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    Rcw = sRcw/scw;
    tcw = Scw.rowRange(0,3).col(3)/scw;
    Ow = -Rcw.t()*tcw;
}

}  // namespace

int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th)
{
    cv::Mat Rcw, tcw, Ow;
    DecomposeSim3(Scw, Rcw, tcw, Ow);
    set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));

    const size_t nQ = vpPoints.size(), nT = pKF->mvKeysUn.size();
    if(nQ == 0 || nT == 0)
        return 0;
    Projected q(nQ);
    for(size_t i = 0; i < nQ; i++)
    {
        MapPoint* pMP = vpPoints[i];
        if(pMP->isBad() || spAlreadyFound.count(pMP))
            continue;
        float u, v, ur, radius;
        int level;
        if(ProjectWorld(pKF, pMP, Rcw, tcw, Ow, th, 0.f, u, v, ur, radius, level))
            q.Set(i, u, v, ur, radius, level, pMP);
    }
    Projected qc(0);
    vector<size_t> orig;
    q.Compact(qc, orig);
    const size_t nV = orig.size();
    if(nV == 0)
        return 0;
    KeyFrameArrays t(pKF);
    vector<unsigned char> taken(nT, 0);
    vector<int> match(nT, -1);
    int nTotal = 0;
    // a matched feature is closed for the later points (:375-376, :396): the greedy window search; later chunks see what
    // the earlier ones left in vpMatched
    for(size_t c0 = 0; c0 < nV; c0 += kQueryChunk)
    {
        const size_t nc = min(kQueryChunk, nV - c0);
        for(size_t k = 0; k < nT; k++)
            taken[k] = vpMatched[k] != NULL;
        int n = 0;
        if(eaof_match_windows(Matcher(max(nT, nc)), EAOF_WIN_BEST, (int)nT, P(t.keys.x), P(t.keys.y), P(t.keys.octave), NULL, P(t.desc),
                              NULL, P(taken), pKF->mnMinX, pKF->mnMaxX, pKF->mnMinY, pKF->mnMaxY, pKF->mfGridElementWidthInv,
                              pKF->mfGridElementHeightInv, (int)nc, NULL, &qc.u[c0], &qc.v[c0], &qc.radius[c0], &qc.minL[c0],
                              &qc.maxL[c0], NULL, NULL, &qc.desc[32 * c0], NULL, TH_LOW, mfNNratio, 0, 0, P(match), NULL,
                              &n) != EAOF_OK)
            Throw("eaof_match_windows");
        for(size_t k = 0; k < nT; k++)
            if(match[k] >= 0)
                vpMatched[k] = vpPoints[orig[c0 + match[k]]];
        nTotal += n;
    }
    return nTotal;
}

int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th)
{
    cv::Mat Rcw = pKF->GetRotation();
    cv::Mat tcw = pKF->GetTranslation();
    cv::Mat Ow = pKF->GetCameraCenter();

    const size_t nQ = vpMapPoints.size();
    Projected q(nQ);
    for(size_t i = 0; i < nQ; i++)
    {
        MapPoint* pMP = vpMapPoints[i];
        if(!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF))  // neither state is ever left again
            continue;
        float u, v, ur, radius;
        int level;
        if(ProjectWorld(pKF, pMP, Rcw, tcw, Ow, th, pKF->mbf, u, v, ur, radius, level))
            q.Set(i, u, v, ur, radius, level, pMP);
    }
    vector<int> match;
    SearchIndependent(pKF, q, EAOF_GATE_FUSE_CHI2, TH_LOW, match);

    // the reference's bookkeeping, in candidate order; isBad / IsInKeyFrame change while the loop runs (:846-849)
    int nFused = 0;
    for(size_t i = 0; i < nQ; i++)
    {
        MapPoint* pMP = vpMapPoints[i];
        if(!pMP || match[i] < 0)
            continue;
        if(pMP->isBad() || pMP->IsInKeyFrame(pKF))
            continue;
        const int bestIdx = match[i];
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
        if(!pMPinKF)
        {
            pMP->AddObservation(pKF,bestIdx);
            pKF->AddMapPoint(pMP,bestIdx);
        }
        else if(!pMPinKF->isBad())
        {
            if(pMPinKF->Observations()>pMP->Observations())
                pMP->Replace(pMPinKF);
            else
                pMPinKF->Replace(pMP);
        }
        nFused++;
    }
    return nFused;
}

int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint)
{
    cv::Mat Rcw, tcw, Ow;
    DecomposeSim3(Scw, Rcw, tcw, Ow);
    const set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();

    const size_t nQ = vpPoints.size();
    Projected q(nQ);
    for(size_t i = 0; i < nQ; i++)
    {
        MapPoint* pMP = vpPoints[i];
        if(pMP->isBad() || spAlreadyFound.count(pMP))
            continue;
        float u, v, ur, radius;
        int level;
        if(ProjectWorld(pKF, pMP, Rcw, tcw, Ow, th, 0.f, u, v, ur, radius, level))
            q.Set(i, u, v, ur, radius, level, pMP);
    }
    vector<int> match;
    SearchIndependent(pKF, q, EAOF_GATE_NONE, TH_LOW, match);

    int nFused = 0;
    for(size_t i = 0; i < nQ; i++)
    {
        if(match[i] < 0)
            continue;
        MapPoint* pMP = vpPoints[i];
        const int bestIdx = match[i];
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
        if(!pMPinKF)
        {
            pMP->AddObservation(pKF,bestIdx);
            pKF->AddMapPoint(pMP,bestIdx);
        }
        else if(!pMPinKF->isBad())
            vpReplacePoint[i] = pMPinKF;
        nFused++;
    }
    return nFused;
}

int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th)
{
    cv::Mat R1w = pKF1->GetRotation();
    cv::Mat t1w = pKF1->GetTranslation();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12*R12;
    cv::Mat sR21 = (1.0/s12)*R12.t();
    cv::Mat t21 = -sR21*t12;

    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = vpMapPoints1.size(), N2 = vpMapPoints2.size();
    vector<bool> vbAlreadyMatched1(N1,false), vbAlreadyMatched2(N2,false);
    for(int i = 0; i < N1; i++)
    {
        MapPoint* pMP = vpMatches12[i];
        if(!pMP)
            continue;
        vbAlreadyMatched1[i] = true;
        const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
        if(idx2>=0 && idx2<N2)
            vbAlreadyMatched2[idx2] = true;
    }

    // one direction: the points of keyframe A moved into keyframe B's camera and searched among B's features
    struct Dir
    {
        static void Run(const vector<MapPoint*>& vpA, const vector<bool>& vbDoneA, const cv::Mat& RAw, const cv::Mat& tAw,
                        const cv::Mat& sRBA, const cv::Mat& tBA, KeyFrame* pKFB, float fx, float fy, float cx, float cy, float th,
                        vector<int>& vnMatch)
        {
            const size_t n = vpA.size();
            Projected q(n);
            for(size_t i = 0; i < n; i++)
            {
                MapPoint* pMP = vpA[i];
                if(!pMP || vbDoneA[i] || pMP->isBad())
                    continue;
                cv::Mat p3Dw = pMP->GetWorldPos();
                cv::Mat p3DcA = RAw*p3Dw + tAw;
                cv::Mat p3DcB = sRBA*p3DcA + tBA;
                if(p3DcB.at<float>(2)<0.0)
                    continue;
                const float invz = 1.0/p3DcB.at<float>(2);
                const float x = p3DcB.at<float>(0)*invz;
                const float y = p3DcB.at<float>(1)*invz;
                const float u = fx*x+cx;
                const float v = fy*y+cy;
                if(!pKFB->IsInImage(u,v))
                    continue;
                const float maxDistance = pMP->GetMaxDistanceInvariance();
                const float minDistance = pMP->GetMinDistanceInvariance();
                const float dist3D = cv::norm(p3DcB);
                if(dist3D<minDistance || dist3D>maxDistance)
                    continue;
                const int level = pMP->PredictScale(dist3D,pKFB->mfLogScaleFactor);
                q.Set(i, u, v, 0.f, th*pKFB->mvScaleFactors[level], level, pMP);
            }
            SearchIndependent(pKFB, q, EAOF_GATE_NONE, TH_HIGH, vnMatch);
        }
    };
    // both directions project with pKF1's intrinsics, as the reference does (:1105-1108)
    vector<int> vnMatch1, vnMatch2;
    Dir::Run(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy, th, vnMatch1);
    Dir::Run(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, pKF1, pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy, th, vnMatch2);

    int nFound = 0;
    for(int i1 = 0; i1 < N1; i1++)
    {
        const int idx2 = vnMatch1[i1];
        if(idx2 >= 0 && vnMatch2[idx2] == i1)
        {
            vpMatches12[i1] = vpMapPoints2[idx2];
            nFound++;
        }
    }
    return nFound;
}

// The two protected helpers of the class.  No method of this file calls them (the kernels apply both rules on the device),
// they are defined for link completeness with the reference's header.
bool ORBmatcher::CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF2)
{
    // l2 = kp1^T * F12 (a, b, c); squared point-line distance of kp2 against the chi-square bound at kp2's level
    const float a = kp1.pt.x * F12.at<float>(0, 0) + kp1.pt.y * F12.at<float>(1, 0) + F12.at<float>(2, 0);
    const float b = kp1.pt.x * F12.at<float>(0, 1) + kp1.pt.y * F12.at<float>(1, 1) + F12.at<float>(2, 1);
    const float c = kp1.pt.x * F12.at<float>(0, 2) + kp1.pt.y * F12.at<float>(1, 2) + F12.at<float>(2, 2);
    const float num = a * kp2.pt.x + b * kp2.pt.y + c;
    const float den = a * a + b * b;
    if(den == 0)
        return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * pKF2->mvLevelSigma2[kp2.octave];
}

void ORBmatcher::ComputeThreeMaxima(vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3)
{
    int best[3] = {0, 0, 0};
    int where[3] = {-1, -1, -1};
    for(int i = 0; i < L; i++)
    {
        const int s = (int)histo[i].size();
        int slot = s > best[0] ? 0 : s > best[1] ? 1 : s > best[2] ? 2 : 3;  // strictly larger: earlier bins win ties
        for(int k = 2; k > slot; k--)
        {
            best[k] = best[k - 1];
            where[k] = where[k - 1];
        }
        if(slot < 3)
        {
            best[slot] = s;
            where[slot] = i;
        }
    }
    ind1 = where[0];
    ind2 = where[1];
    ind3 = where[2];
    if(best[1] < 0.1f * (float)best[0])
    {
        ind2 = -1;
        ind3 = -1;
    }
    else if(best[2] < 0.1f * (float)best[0])
        ind3 = -1;
}

} //namespace ORB_SLAM
