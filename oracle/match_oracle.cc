// ORACLE — test infrastructure only.  Nothing here is linked into, imported by or executed from the product path.
//
// Line-by-line restatement, on plain arrays, of the reference's Hamming search loops
// (/root/reference/src/ORBmatcher.cc) and of the Frame grid they query (src/Frame.cc).  MapPoint*/KeyFrame*/Frame
// objects become arrays: "has a good map point" is a validity flag, DBoW2::FeatureVector
// (std::map<NodeId, vector<unsigned>>) is a CSR triple (sorted node ids, starts, feature indices), and the 64x48
// feature grid is rebuilt from the undistorted keypoints exactly as Frame::AssignFeaturesToGrid does.
//
// PARITY STATUS: pinned.  The reference ships no tests or golden vectors, but its ORBmatcher.cc compiles here
// unmodified against oracle/matchshim (array-backed Frame/KeyFrame/MapPoint stand-ins) into oracle/_ref/libmatch_ref.so;
// tests/test_oracle_matcher_vs_ref.py drives both with the same arrays and requires identical match indices and
// counts for every loop restated below.  (Distances are not an output of the reference methods; the restatement's
// distances are checked against DescriptorDistance separately.)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;  // src/ORBmatcher.cc:37-39
const int GRID_COLS = 64, GRID_ROWS = 48;                 // include/Frame.h:89-90

// ORBmatcher::DescriptorDistance  src/ORBmatcher.cc:1649-1665 (the SWAR bit hack, kept literally)
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4);
        memcpy(&y, b + 4 * i, 4);
        unsigned int v = x ^ y;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// ORBmatcher::ComputeThreeMaxima  src/ORBmatcher.cc:1603-1644
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

int rot_bin(float a1, float a2, float factor) {  // e.g. src/ORBmatcher.cc:238-244
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

struct Csr { int n; const int* id; const int* start; const int* idx; };

}  // namespace

extern "C" {

int eaoo_hamming(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

void eaoo_three_maxima(const int* sizes, int L, int* i1, int* i2, int* i3) {
    std::vector<std::vector<int>> h(L);
    for (int i = 0; i < L; ++i) h[i].resize(sizes[i]);
    *i1 = *i2 = *i3 = -1;
    three_maxima(h.data(), L, *i1, *i2, *i3);
}

// mode 0: SearchByBoW(KeyFrame*, Frame&, ...)      src/ORBmatcher.cc:159-288   (accept best <= TH_LOW)
//         matchOut/distOut have nT entries: index of the matched query (KF feature) per target (F feature), -1 = none
// mode 1: SearchByBoW(KeyFrame*, KeyFrame*, ...)   src/ORBmatcher.cc:522-655   (accept best <  TH_LOW, targets need
//         a good map point) matchOut/distOut have nQ entries: index of the matched target per query
// Q = pKF / pKF1 side (outer loop), T = F / pKF2 side (inner loop).  validQ/validT: "map point present and not bad".
int eaoo_search_by_bow(int mode, int nQ, const uint8_t* descQ, const float* angleQ, const uint8_t* validQ, int nT,
                       const uint8_t* descT, const float* angleT, const uint8_t* validT, int nNodesQ,
                       const int* nodeIdQ, const int* nodeStartQ, const int* nodeIdxQ, int nNodesT, const int* nodeIdT,
                       const int* nodeStartT, const int* nodeIdxT, float nnratio, int checkOri, int* matchOut,
                       int* distOut) {
    const int nOut = mode == 0 ? nT : nQ;
    for (int i = 0; i < nOut; ++i) { matchOut[i] = -1; if (distOut) distOut[i] = -1; }
    std::vector<char> matchedT(nT, 0);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;  // :172 / :541 (upstream quirk)
    int nmatches = 0;
    int a = 0, b = 0;
    while (a < nNodesQ && b < nNodesT) {
        if (nodeIdQ[a] == nodeIdT[b]) {
            for (int iq = nodeStartQ[a]; iq < nodeStartQ[a + 1]; ++iq) {
                const int q = nodeIdxQ[iq];
                if (validQ && !validQ[q]) continue;
                int best1 = 256, bestIdx = -1, best2 = 256;
                for (int it = nodeStartT[b]; it < nodeStartT[b + 1]; ++it) {
                    const int t = nodeIdxT[it];
                    if (matchedT[t]) continue;
                    if (mode == 1 && validT && !validT[t]) continue;
                    const int dist = descriptor_distance(descQ + 32 * (size_t)q, descT + 32 * (size_t)t);
                    if (dist < best1) { best2 = best1; best1 = dist; bestIdx = t; }
                    else if (dist < best2) { best2 = dist; }
                }
                const bool thOk = mode == 0 ? best1 <= TH_LOW : best1 < TH_LOW;
                if (thOk && (float)best1 < nnratio * (float)best2) {
                    matchedT[bestIdx] = 1;
                    const int outIdx = mode == 0 ? bestIdx : q;
                    matchOut[outIdx] = mode == 0 ? q : bestIdx;
                    if (distOut) distOut[outIdx] = best1;
                    if (checkOri) rotHist[rot_bin(angleQ[q], angleT[bestIdx], factor)].push_back(outIdx);
                    ++nmatches;
                }
            }
            ++a; ++b;
        } else if (nodeIdQ[a] < nodeIdT[b]) {
            while (a < nNodesQ && nodeIdQ[a] < nodeIdT[b]) ++a;  // lower_bound on a sorted map
        } else {
            while (b < nNodesT && nodeIdT[b] < nodeIdQ[a]) ++b;
        }
    }
    if (checkOri) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int j : rotHist[i]) { matchOut[j] = -1; if (distOut) distOut[j] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchForTriangulation  src/ORBmatcher.cc:657-823 with CheckDistEpipolarLine :140-157.
// freeQ/freeT: feature has NO map point; stereoQ/stereoT: mvuRight >= 0.  F12 row-major 3x3.  match12 has n1 entries.
int eaoo_search_for_triangulation(int n1, const uint8_t* desc1, const float* x1, const float* y1, const float* angle1,
                                  const uint8_t* free1, const uint8_t* stereo1, int n2, const uint8_t* desc2,
                                  const float* x2, const float* y2, const int* oct2, const float* angle2,
                                  const uint8_t* free2, const uint8_t* stereo2, int nNodes1, const int* nodeId1,
                                  const int* nodeStart1, const int* nodeIdx1, int nNodes2, const int* nodeId2,
                                  const int* nodeStart2, const int* nodeIdx2, const float* F12, float ex, float ey,
                                  const float* scaleFactors2, const float* levelSigma2_2, int onlyStereo, int checkOri,
                                  int* match12, int* dist12) {
    for (int i = 0; i < n1; ++i) { match12[i] = -1; if (dist12) dist12[i] = -1; }
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;  // :684
    int nmatches = 0, a = 0, b = 0;
    while (a < nNodes1 && b < nNodes2) {
        if (nodeId1[a] == nodeId2[b]) {
            for (int i1 = nodeStart1[a]; i1 < nodeStart1[a + 1]; ++i1) {
                const int idx1 = nodeIdx1[i1];
                if (!free1[idx1]) continue;
                const bool bStereo1 = stereo1 && stereo1[idx1];
                if (onlyStereo && !bStereo1) continue;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int i2 = nodeStart2[b]; i2 < nodeStart2[b + 1]; ++i2) {
                    const int idx2 = nodeIdx2[i2];
                    if (!free2[idx2]) continue;  // vbMatched2 is never set inside the loop (upstream behaviour)
                    const bool bStereo2 = stereo2 && stereo2[idx2];
                    if (onlyStereo && !bStereo2) continue;
                    const int dist = descriptor_distance(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - x2[idx2], distey = ey - y2[idx2];
                        if (distex * distex + distey * distey < 100 * scaleFactors2[oct2[idx2]]) continue;
                    }
                    // CheckDistEpipolarLine
                    const float la = x1[idx1] * F12[0] + y1[idx1] * F12[3] + F12[6];
                    const float lb = x1[idx1] * F12[1] + y1[idx1] * F12[4] + F12[7];
                    const float lc = x1[idx1] * F12[2] + y1[idx1] * F12[5] + F12[8];
                    const float num = la * x2[idx2] + lb * y2[idx2] + lc;
                    const float den = la * la + lb * lb;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    if (dsqr < 3.84 * levelSigma2_2[oct2[idx2]]) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match12[idx1] = bestIdx2;
                    if (dist12) dist12[idx1] = bestDist;
                    ++nmatches;
                    if (checkOri) rotHist[rot_bin(angle1[idx1], angle2[bestIdx2], factor)].push_back(idx1);
                }
            }
            ++a; ++b;
        } else if (nodeId1[a] < nodeId2[b]) {
            while (a < nNodes1 && nodeId1[a] < nodeId2[b]) ++a;
        } else {
            while (b < nNodes2 && nodeId2[b] < nodeId1[a]) ++b;
        }
    }
    if (checkOri) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int j : rotHist[i]) { match12[j] = -1; if (dist12) dist12[j] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// Frame::AssignFeaturesToGrid + PosInGrid  src/Frame.cc:599-614,751-761.  cellStart has 64*48+1 entries, cells are
// numbered ix*48+iy (mGrid[ix][iy]); cellIdx holds feature indices in ascending order per cell.
void eaoo_build_grid(int n, const float* x, const float* y, float minX, float minY, float invW, float invH,
                     int* cellStart, int* cellIdx) {
    std::vector<std::vector<int>> g(GRID_COLS * GRID_ROWS);
    for (int i = 0; i < n; ++i) {
        const int px = (int)roundf((x[i] - minX) * invW), py = (int)roundf((y[i] - minY) * invH);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
        g[px * GRID_ROWS + py].push_back(i);
    }
    int o = 0;
    for (int c = 0; c < GRID_COLS * GRID_ROWS; ++c) {
        cellStart[c] = o;
        for (int i : g[c]) cellIdx[o++] = i;
    }
    cellStart[GRID_COLS * GRID_ROWS] = o;
}

static void features_in_area(const int* cellStart, const int* cellIdx, const float* kx, const float* ky,
                             const int* koct, float x, float y, float r, int minLevel, int maxLevel, float minX,
                             float minY, float invW, float invH, std::vector<int>& out) {
    // Frame::GetFeaturesInArea  src/Frame.cc:696-749
    out.clear();
    const int nMinCellX = std::max(0, (int)floorf((x - minX - r) * invW));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)ceilf((x - minX + r) * invW));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)floorf((y - minY - r) * invH));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)ceilf((y - minY + r) * invH));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
        for (int iy = nMinCellY; iy <= nMaxCellY; ++iy) {
            const int c = ix * GRID_ROWS + iy;
            for (int j = cellStart[c]; j < cellStart[c + 1]; ++j) {
                const int k = cellIdx[j];
                if (bCheckLevels) {
                    if (koct[k] < minLevel) continue;
                    if (maxLevel >= 0 && koct[k] > maxLevel) continue;
                }
                const float dx = kx[k] - x, dy = ky[k] - y;
                if (fabsf(dx) < r && fabsf(dy) < r) out.push_back(k);
            }
        }
}

// ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)  src/ORBmatcher.cc:1328-1472.
// The 3-D projection (:1364-1377) is the caller's: lu/lv are the projected pixel, linvz = 1/z (NULL: all positive).
// lvalid: map point present and not an outlier; lobs: pMP->Observations()>0 (NULL: all true); ctaken: Cur feature
// already holds a map point with observations (NULL: none).  searchMode 0: octave-1..octave+1, 1: forward
// (>= octave), 2: backward (0..octave).  matchCur/distCur: nC entries, index of the Last feature / its distance.
int eaoo_search_by_projection_last(int nC, const float* cx, const float* cy, const int* coct, const float* cangle,
                                   const uint8_t* cdesc, const float* curight, const uint8_t* ctaken, float minX,
                                   float maxX, float minY, float maxY, float invW, float invH, int nL,
                                   const uint8_t* lvalid, const float* lu, const float* lv, const float* linvz,
                                   const int* loct, const float* langle, const uint8_t* ldesc, const uint8_t* lobs,
                                   const float* scaleFactors, float th, float mbf, int searchMode, int checkOri,
                                   int* matchCur, int* distCur) {
    std::vector<int> cellStart(GRID_COLS * GRID_ROWS + 1), cellIdx(nC > 0 ? nC : 1);
    eaoo_build_grid(nC, cx, cy, minX, minY, invW, invH, cellStart.data(), cellIdx.data());
    std::vector<char> hasObs(nC, 0);
    for (int i = 0; i < nC; ++i) { matchCur[i] = -1; if (distCur) distCur[i] = -1; hasObs[i] = ctaken ? ctaken[i] : 0; }
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = HISTO_LENGTH / 360.0f;  // :1337
    int nmatches = 0;
    std::vector<int> cand;
    for (int i = 0; i < nL; ++i) {
        if (lvalid && !lvalid[i]) continue;
        const float invzc = linvz ? linvz[i] : 1.f;
        if (invzc < 0) continue;
        const float u = lu[i], v = lv[i];
        if (u < minX || u > maxX) continue;
        if (v < minY || v > maxY) continue;
        const int oct = loct[i];
        const float radius = th * scaleFactors[oct];
        if (searchMode == 1) features_in_area(cellStart.data(), cellIdx.data(), cx, cy, coct, u, v, radius, oct, -1, minX, minY, invW, invH, cand);
        else if (searchMode == 2) features_in_area(cellStart.data(), cellIdx.data(), cx, cy, coct, u, v, radius, 0, oct, minX, minY, invW, invH, cand);
        else features_in_area(cellStart.data(), cellIdx.data(), cx, cy, coct, u, v, radius, oct - 1, oct + 1, minX, minY, invW, invH, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if (hasObs[i2]) continue;
            if (curight && curight[i2] > 0) {
                const float ur = u - mbf * invzc;
                const float er = fabsf(ur - curight[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(ldesc + 32 * (size_t)i, cdesc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            matchCur[bestIdx2] = i;
            if (distCur) distCur[bestIdx2] = bestDist;
            hasObs[bestIdx2] = lobs ? lobs[i] : 1;
            ++nmatches;
            if (checkOri) rotHist[rot_bin(langle[i], cangle[bestIdx2], factor)].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i != i1 && i != i2 && i != i3)
                for (int j : rotHist[i]) { matchCur[j] = -1; if (distCur) distCur[j] = -2; --nmatches; }  // -2: matched, then pruned
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th)  src/ORBmatcher.cc:45-129.
// Per map point: inView (mbTrackInView), bad, projection (px, py, pxr), predicted level, view cosine, descriptor,
// pobs = Observations()>0.  Frame: undistorted keypoints, descriptors, mvuRight (NULL: all -1), ftaken = feature
// already holds a map point with observations.  matchF: nF entries, index of the map point assigned to the feature.
int eaoo_search_by_projection_mappoints(int nF, const float* fx, const float* fy, const int* foct, const uint8_t* fdesc,
                                        const float* furight, const uint8_t* ftaken, float minX, float minY, float invW,
                                        float invH, int nMP, const uint8_t* inView, const uint8_t* bad, const float* px,
                                        const float* py, const float* pxr, const int* plevel, const float* pcos,
                                        const uint8_t* pdesc, const uint8_t* pobs, const float* scaleFactors, float th,
                                        float nnratio, int* matchF, int* distF) {
    std::vector<int> cellStart(GRID_COLS * GRID_ROWS + 1), cellIdx(nF > 0 ? nF : 1);
    eaoo_build_grid(nF, fx, fy, minX, minY, invW, invH, cellStart.data(), cellIdx.data());
    std::vector<char> hasObs(nF, 0);
    for (int i = 0; i < nF; ++i) { matchF[i] = -1; if (distF) distF[i] = -1; hasObs[i] = ftaken ? ftaken[i] : 0; }
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> cand;
    for (int iMP = 0; iMP < nMP; ++iMP) {
        if (inView && !inView[iMP]) continue;
        if (bad && bad[iMP]) continue;
        const int nPredictedLevel = plevel[iMP];
        float r = (pcos ? pcos[iMP] : 1.f) > 0.998 ? 2.5 : 4.0;  // RadiusByViewingCos :131-137 (double compare)
        if (bFactor) r *= th;
        const float rr = r * scaleFactors[nPredictedLevel];
        features_in_area(cellStart.data(), cellIdx.data(), fx, fy, foct, px[iMP], py[iMP], rr, nPredictedLevel - 1,
                         nPredictedLevel, minX, minY, invW, invH, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : cand) {
            if (hasObs[idx]) continue;
            if (furight && furight[idx] > 0) {
                const float er = fabsf((pxr ? pxr[iMP] : 0.f) - furight[idx]);
                if (er > r * scaleFactors[nPredictedLevel]) continue;
            }
            const int dist = descriptor_distance(pdesc + 32 * (size_t)iMP, fdesc + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = foct[idx]; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = foct[idx]; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            matchF[bestIdx] = iMP;
            if (distF) distF[bestIdx] = bestDist;
            hasObs[bestIdx] = pobs ? pobs[iMP] : 1;
            ++nmatches;
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist)  src/ORBmatcher.cc:1474-1601, after
// the caller's projection (:1497-1525): per KF feature kvalid (map point present, not bad, not already found, distance
// inside the scale pyramid), projected pixel (ku, kv), predicted level, keypoint angle, map-point descriptor.
// ctaken: Cur feature already holds a map point.  matchCur: nC entries, index of the KF feature.
int eaoo_search_by_projection_kf(int nC, const float* cx, const float* cy, const int* coct, const float* cangle,
                                 const uint8_t* cdesc, const uint8_t* ctaken, float minX, float maxX, float minY,
                                 float maxY, float invW, float invH, int nK, const uint8_t* kvalid, const float* ku,
                                 const float* kv, const int* klevel, const float* kangle, const uint8_t* kdesc,
                                 const float* scaleFactors, float th, int orbDist, int checkOri, int* matchCur,
                                 int* distCur) {
    std::vector<int> cellStart(GRID_COLS * GRID_ROWS + 1), cellIdx(nC > 0 ? nC : 1);
    eaoo_build_grid(nC, cx, cy, minX, minY, invW, invH, cellStart.data(), cellIdx.data());
    std::vector<char> has(nC, 0);
    for (int i = 0; i < nC; ++i) { matchCur[i] = -1; if (distCur) distCur[i] = -1; has[i] = ctaken ? ctaken[i] : 0; }
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;  // :1486
    int nmatches = 0;
    std::vector<int> cand;
    for (int i = 0; i < nK; ++i) {
        if (!kvalid[i]) continue;
        const float u = ku[i], v = kv[i];
        if (u < minX || u > maxX) continue;
        if (v < minY || v > maxY) continue;
        const int nPredictedLevel = klevel[i];
        const float radius = th * scaleFactors[nPredictedLevel];
        features_in_area(cellStart.data(), cellIdx.data(), cx, cy, coct, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1,
                         minX, minY, invW, invH, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if (has[i2]) continue;
            const int dist = descriptor_distance(kdesc + 32 * (size_t)i, cdesc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= orbDist) {
            matchCur[bestIdx2] = i;
            if (distCur) distCur[bestIdx2] = bestDist;
            has[bestIdx2] = 1;
            ++nmatches;
            if (checkOri) rotHist[rot_bin(kangle[i], cangle[bestIdx2], factor)].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != i1 && i != i2 && i != i3)
                for (int j : rotHist[i]) { matchCur[j] = -1; if (distCur) distCur[j] = -2; --nmatches; }  // -2: matched, then pruned
    }
    return nmatches;
}

// ORBmatcher::SearchForInitialization  src/ORBmatcher.cc:405-520.  prev: vbPrevMatched as (x,y) pairs, in/out.
int eaoo_search_for_initialization(int n1, const int* oct1, const float* angle1, const uint8_t* desc1, float* prev, int n2,
                                   const float* x2, const float* y2, const int* oct2, const float* angle2,
                                   const uint8_t* desc2, float minX, float minY, float invW, float invH, int windowSize,
                                   float nnratio, int checkOri, int* matches12) {
    std::vector<int> cellStart(GRID_COLS * GRID_ROWS + 1), cellIdx(n2 > 0 ? n2 : 1);
    eaoo_build_grid(n2, x2, y2, minX, minY, invW, invH, cellStart.data(), cellIdx.data());
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;  // :413
    std::vector<int> vMatchedDistance(n2, 0x7fffffff), vnMatches21(n2, -1), cand;
    for (int i1 = 0; i1 < n1; ++i1) {
        const int level1 = oct1[i1];
        if (level1 > 0) continue;
        features_in_area(cellStart.data(), cellIdx.data(), x2, y2, oct2, prev[2 * i1], prev[2 * i1 + 1], (float)windowSize,
                         level1, level1, minX, minY, invW, invH, cand);
        if (cand.empty()) continue;
        int bestDist = 0x7fffffff, bestDist2 = 0x7fffffff, bestIdx2 = -1;
        for (int i2 : cand) {
            const int dist = descriptor_distance(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { matches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOri) rotHist[rot_bin(angle1[i1], angle2[bestIdx2], factor)].push_back(i1);
            }
        }
    }
    if (checkOri) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int idx1 : rotHist[i])
                if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; ++i1)
        if (matches12[i1] >= 0) { prev[2 * i1] = x2[matches12[i1]]; prev[2 * i1 + 1] = y2[matches12[i1]]; }
    return nmatches;
}

// ---- map-side window matchers: KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:608-647, the Frame walk without a level
// filter) and a best-distance loop with the level window [predicted-1, predicted] applied by the caller -------------

// ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)  src/ORBmatcher.cc:290-403, after the
// caller's projection (:318-360): per candidate map point qvalid (not bad, not already found, positive depth, inside
// the image, inside the scale-invariance range, viewing angle below 60 deg), pixel (qu, qv), predicted level, descriptor.
// tmatched: vpMatched[idx] != NULL on entry.  matchT/distT: nT entries, index of the map point newly assigned to the
// keyframe feature (-1: unchanged).
int eaoo_search_by_projection_sim3kf(int nT, const float* tx, const float* ty, const int* toct, const uint8_t* tdesc,
                                     const uint8_t* tmatched, float minX, float minY, float invW, float invH, int nQ,
                                     const uint8_t* qvalid, const float* qu, const float* qv, const int* qlevel,
                                     const uint8_t* qdesc, const float* scaleFactors, int th, int* matchT, int* distT) {
    std::vector<int> cellStart(GRID_COLS * GRID_ROWS + 1), cellIdx(nT > 0 ? nT : 1), cand;
    eaoo_build_grid(nT, tx, ty, minX, minY, invW, invH, cellStart.data(), cellIdx.data());
    std::vector<uint8_t> matched(nT, 0);
    for (int k = 0; k < nT; ++k) { matched[k] = tmatched && tmatched[k]; matchT[k] = -1; if (distT) distT[k] = -1; }
    int nmatches = 0;
    for (int iMP = 0; iMP < nQ; ++iMP) {
        if (qvalid && !qvalid[iMP]) continue;
        const int nPredictedLevel = qlevel[iMP];
        const float radius = th * scaleFactors[nPredictedLevel];  // int * float, :358
        features_in_area(cellStart.data(), cellIdx.data(), tx, ty, toct, qu[iMP], qv[iMP], radius, -1, -1, minX, minY, invW,
                         invH, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            if (matched[idx]) continue;                                                    // :375-376
            if (toct[idx] < nPredictedLevel - 1 || toct[idx] > nPredictedLevel) continue;  // :380-381
            const int dist = descriptor_distance(qdesc + 32 * (size_t)iMP, tdesc + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW) {  // :394-398
            matched[bestIdx] = 1;
            matchT[bestIdx] = iMP;
            if (distT) distT[bestIdx] = bestDist;
            ++nmatches;
        }
    }
    return nmatches;
}

// The search step shared by Fuse(KeyFrame*, vpMapPoints, th) (src/ORBmatcher.cc:887-936, gate 1: the reprojection
// chi-square :901-925), Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) (:1049-1075, gate 0) and both directions of
// SearchBySim3 (:1196-1223, :1276-1303, gate 0): best target of every query, independent of the other queries.
// Returns the number of queries with best <= thAccept.
int eaoo_window_best(int gate, int nT, const float* tx, const float* ty, const int* toct, const uint8_t* tdesc,
                     const float* turight, const float* invLevelSigma2, float minX, float minY, float invW, float invH,
                     int nQ, const uint8_t* qvalid, const float* qu, const float* qv, const float* qur, const int* qlevel,
                     const uint8_t* qdesc, const float* scaleFactors, float th, int thAccept, int* matchQ, int* distQ) {
    std::vector<int> cellStart(GRID_COLS * GRID_ROWS + 1), cellIdx(nT > 0 ? nT : 1), cand;
    eaoo_build_grid(nT, tx, ty, minX, minY, invW, invH, cellStart.data(), cellIdx.data());
    int n = 0;
    for (int i = 0; i < nQ; ++i) {
        matchQ[i] = -1;
        if (distQ) distQ[i] = -1;
        if (qvalid && !qvalid[i]) continue;
        const int nPredictedLevel = qlevel[i];
        const float u = qu[i], v = qv[i];
        const float radius = th * scaleFactors[nPredictedLevel];
        features_in_area(cellStart.data(), cellIdx.data(), tx, ty, toct, u, v, radius, -1, -1, minX, minY, invW, invH, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            const int kpLevel = toct[idx];
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            if (gate == 1) {
                const float kpx = tx[idx], kpy = ty[idx];
                if (turight && turight[idx] >= 0) {  // :901-914
                    const float kpr = turight[idx];
                    const float ex = u - kpx, ey = v - kpy, er = qur[i] - kpr;
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * invLevelSigma2[kpLevel] > 7.8) continue;
                } else {  // :915-925
                    const float ex = u - kpx, ey = v - kpy;
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * invLevelSigma2[kpLevel] > 5.99) continue;
                }
            }
            const int dist = descriptor_distance(qdesc + 32 * (size_t)i, tdesc + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= thAccept) {
            matchQ[i] = bestIdx;
            if (distQ) distQ[i] = bestDist;
            ++n;
        }
    }
    return n;
}

// What Fuse(KeyFrame*, vpMapPoints, th) does with the search result, in query order (src/ORBmatcher.cc:846-849 re-read
// at each turn, :938-957), on array state.  Map points are numbered: candidate i -> qid[i] (its first occurrence when
// the same MapPoint* is listed twice), the point a keyframe feature holds on entry -> nQ+idx.  Per point: bad, obs
// (Observations()), inKF (IsInKeyFrame(pKF)); slot[idx] = point held by feature idx (-1 none).  Replace(a by b) marks a
// bad and records replacedBy[a] = b (the stand-in MapPoint of oracle/matchshim).  matchQ = the search result of the
// candidates that pass the static skip tests (projection, distance, viewing angle); qnull = vpMapPoints[i] == NULL.
int eaoo_fuse_apply(int nT, int nQ, const int* matchQ, const uint8_t* qnull, const int* qid, int* slot, uint8_t* bad,
                    int* obs, uint8_t* inKF, int* replacedBy, int* addedAt) {
    int nFused = 0;
    for (int i = 0; i < nQ; ++i) {
        addedAt[i] = -1;
        if (qnull && qnull[i]) continue;
        const int p = qid ? qid[i] : i;
        if (bad[p] || inKF[p]) continue;
        const int bestIdx = matchQ[i];
        if (bestIdx < 0) continue;
        const int inKFpt = slot[bestIdx];
        if (inKFpt >= 0) {
            if (!bad[inKFpt]) {
                if (obs[inKFpt] > obs[p]) { bad[p] = 1; replacedBy[p] = inKFpt; }
                else { bad[inKFpt] = 1; replacedBy[inKFpt] = p; }
            }
        } else {
            inKF[p] = 1; ++obs[p]; addedAt[i] = bestIdx;  // AddObservation
            slot[bestIdx] = p;                             // AddMapPoint
        }
        ++nFused;
    }
    (void)nT;
    return nFused;
}

// Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint): the result step :1078-1093.  replacePoint[i] = point to be
// replaced (numbering as above, -1 none).
int eaoo_fuse_sim3_apply(int nT, int nQ, const int* matchQ, int* slot, const uint8_t* bad, int* replacePoint, int* addedAt) {
    int nFused = 0;
    for (int i = 0; i < nQ; ++i) {
        addedAt[i] = -1;
        replacePoint[i] = -1;
        const int bestIdx = matchQ[i];
        if (bestIdx < 0) continue;
        const int inKFpt = slot[bestIdx];
        if (inKFpt >= 0) {
            if (!bad[inKFpt]) replacePoint[i] = inKFpt;
        } else {
            addedAt[i] = bestIdx;
            slot[bestIdx] = i;
        }
        ++nFused;
    }
    (void)nT;
    return nFused;
}

// SearchBySim3: the agreement pass src/ORBmatcher.cc:1308-1323.  match12out[i1] = idx2 for the mutual matches.
int eaoo_sim3_agreement(int n1, const int* vnMatch1, int n2, const int* vnMatch2, int* match12out) {
    int nFound = 0;
    for (int i1 = 0; i1 < n1; ++i1) {
        match12out[i1] = -1;
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && idx2 < n2) {
            const int idx1 = vnMatch2[idx2];
            if (idx1 == i1) { match12out[i1] = idx2; ++nFound; }
        }
    }
    return nFound;
}

// MapPoint::ComputeDistinctiveDescriptors  src/MapPoint.cc:273-301 for one map point with n observation descriptors:
// returns the index of the descriptor with the least median distance to the others (first on ties), -1 when n == 0.
int eaoo_distinctive_descriptor(int n, const uint8_t* desc, int* medianOut) {
    if (n <= 0) return -1;
    std::vector<float> D((size_t)n * n);
    for (int i = 0; i < n; ++i) {
        D[(size_t)i * n + i] = 0;
        for (int j = i + 1; j < n; ++j) {
            const int d = descriptor_distance(desc + 32 * (size_t)i, desc + 32 * (size_t)j);
            D[(size_t)i * n + j] = d;
            D[(size_t)j * n + i] = d;
        }
    }
    int BestMedian = 0x7fffffff, BestIdx = 0;
    for (int i = 0; i < n; ++i) {
        std::vector<int> vDists(D.begin() + (size_t)i * n, D.begin() + (size_t)(i + 1) * n);
        std::sort(vDists.begin(), vDists.end());
        const int median = vDists[(size_t)(0.5 * (n - 1))];
        if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    if (medianOut) *medianOut = BestMedian;
    return BestIdx;
}

}  // extern "C"

// ---- Frame::ComputeStereoMatches  src/Frame.cc:841-1013 -------------------------------------------------------------------
// Left / right keypoints (level-0 coordinates, octave), descriptors, both image pyramids as bordered level buffers back
// to back (level l at off[l], (w[l]+38) x (h[l]+38) bytes, the level image at (19,19): src/ORBextractor.cc:1114-1116).
// Outputs mvuRight / mvDepth (-1 where unmatched).  sadOut (optional): the SAD of the accepted matches (-1 elsewhere), before
// the median filter.
extern "C" void eaoo_stereo_matches(int nL, const float* xL, const float* yL, const int* octL, const uint8_t* descL, int nR,
                                    const float* xR, const float* yR, const int* octR, const uint8_t* descR, int nLevels,
                                    const float* scale, const float* invScale, const uint8_t* pyrL, const uint8_t* pyrR,
                                    const int* off, const int* w, const int* h, float mb, float mbf, float* uRight,
                                    float* depth, int* sadOut) {
    (void)nLevels;
    for (int i = 0; i < nL; ++i) { uRight[i] = -1.0f; depth[i] = -1.0f; if (sadOut) sadOut[i] = -1; }
    const int nRows = h[0];
    std::vector<std::vector<int>> vRowIndices(nRows);
    for (int iR = 0; iR < nR; ++iR) {  // :855-866
        const float kpY = yR[iR];
        const float r = 2.0f * scale[octR[iR]];
        const int maxr = (int)ceilf(kpY + r);
        const int minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; ++yi)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);  // the reference indexes unchecked
    }
    const float minZ = mb;
    const float minD = -3;
    const float maxD = mbf / minZ;
    auto px = [&](const uint8_t* pyr, int l, int y, int x) -> float {  // mvImagePyramid[l].at<uchar>(y, x) as float
        return (float)pyr[off[l] + (size_t)(y + 19) * (w[l] + 38) + (x + 19)];
    };
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < nL; ++iL) {
        const int levelL = octL[iL];
        const float vL = yL[iL], uL = xL[iL];
        const std::vector<int>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD;
        const float maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        int bestIdxR = 0;
        for (int iR : vCandidates) {
            if (octR[iR] < levelL - 1 || octR[iR] > levelL + 1) continue;
            const float uR = xR[iR];
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(descL + 32 * (size_t)iL, descR + 32 * (size_t)iR);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < TH_HIGH) {  // subpixel match by correlation, :913-1000
            const float uR0 = xR[bestIdxR];
            const float scaleFactor = invScale[levelL];
            const float scaleduL = roundf(uL * scaleFactor);
            const float scaledvL = roundf(vL * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int wp = 5;
            float IL[11][11];
            const int y0 = (int)(scaledvL - wp), x0 = (int)(scaleduL - wp);
            for (int a = 0; a < 11; ++a)
                for (int b = 0; b < 11; ++b) IL[a][b] = px(pyrL, levelL, y0 + a, x0 + b);
            const float cL = IL[wp][wp];
            for (int a = 0; a < 11; ++a)
                for (int b = 0; b < 11; ++b) IL[a][b] = IL[a][b] - cL * 1.0f;
            int bestDistS = 0x7fffffff;
            int bestincR = 0;
            const int L = 5;
            float vDists[11];
            const float iniu = scaleduR0 + L - wp;
            const float endu = scaleduR0 + L + wp + 1;
            if (iniu < 0 || endu >= w[levelL]) continue;
            for (int incR = -L; incR <= +L; ++incR) {
                const int xr0 = (int)(scaleduR0 + incR - wp);
                const float cR = px(pyrR, levelL, y0 + wp, xr0 + wp);
                double acc = 0;  // cv::norm(IL, IR, NORM_L1): double accumulator over exact integers
                for (int a = 0; a < 11; ++a)
                    for (int b = 0; b < 11; ++b) acc += fabs((double)IL[a][b] - (double)(px(pyrR, levelL, y0 + a, xr0 + b) - cR * 1.0f));
                const float dist = (float)acc;
                if (dist < bestDistS) { bestDistS = dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1];
            const float dist2 = vDists[L + bestincR];
            const float dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scale[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= 0 && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
                if (sadOut) sadOut[iL] = bestDistS;
            }
        }
    }
    if (vDistIdx.empty()) return;  // the reference reads vDistIdx[0] of an empty vector here
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        uRight[vDistIdx[i].second] = -1;
        depth[vDistIdx[i].second] = -1;
    }
}

// ---- Frame::UndistortKeyPoints  src/Frame.cc:773-803 = cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) -----------
// OpenCV's cvUndistortPoints in double arithmetic: normalise with the camera matrix, five fixed-point iterations of the
// Brown model inverse (k1 k2 p1 p2 k3 [k4 k5 k6 s1 s2 s3 s4], missing coefficients are 0), re-project with P = mK, store as
// float.  guard = 1 adds OpenCV 4.x's "icdist < 0 -> keep the normalised point" exit (verified against cv2 4.13 in
// tests/test_oracle_primitives.py); OpenCV 3.3.1, the reference's pinned version, has no such exit (guard = 0).
// The k1 == 0 shortcut of the reference (:775-779) is the caller's.
extern "C" void eaoo_undistort_points(int n, const float* xs, const float* ys, float fx, float fy, float cx, float cy,
                                      const float* dist, int nDist, int guard, float* xo, float* yo) {
    double k[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nDist && i < 12; ++i) k[i] = dist[i];
    const double dfx = fx, dfy = fy, dcx = cx, dcy = cy;
    const double ifx = 1. / dfx, ify = 1. / dfy;
    for (int i = 0; i < n; ++i) {
        double x = xs[i], y = ys[i];
        const double u = x, v = y;
        x = (x - dcx) * ifx;
        y = (y - dcy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (guard && icdist < 0) { x = (u - dcx) * ifx; y = (v - dcy) * ify; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = dfx * x + 0.0 * y + dcx, yy = 0.0 * x + dfy * y + dcy, ww = 1. / (0.0 * x + 0.0 * y + 1.0);
        xo[i] = (float)(xx * ww);
        yo[i] = (float)(yy * ww);
    }
}
