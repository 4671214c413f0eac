/*
 * eaof_match.h — C ABI of the Hamming matcher in libeaof_orb.so (B200, sm_100a).
 *
 * Replaces the descriptor search loops of the reference's ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:37-102,
 * src/ORBmatcher.cc).  MapPoint* / KeyFrame* / Frame objects stay on the reference host path; what crosses this
 * boundary are plain arrays: 32-byte descriptors, keypoint angles/octaves/positions, validity flags in place of
 * "has a good map point", DBoW2::FeatureVector as CSR (sorted node ids, starts, feature indices) and the projected
 * pixel of each map point.  Results are bit-exact with the reference loops, including the TH_LOW/TH_HIGH gates,
 * the fp32 ratio test, the greedy "already matched" exclusion (which makes results depend on query order) and the
 * rotation-histogram pruning with each function's own histogram factor.
 *
 * Same conventions as eaof_orb.h: 0 / negative EAOF_ERR_*, eaof_last_error(), no CPU fallback.
 */
#ifndef EAOF_MATCH_H
#define EAOF_MATCH_H

#include <stddef.h>
#include <stdint.h>

#include "eaof_orb.h"

#ifdef __cplusplus
extern "C" {
#endif

#define EAOF_TH_HIGH 100      /* ORBmatcher::TH_HIGH      src/ORBmatcher.cc:37 */
#define EAOF_TH_LOW 50        /* ORBmatcher::TH_LOW       src/ORBmatcher.cc:38 */
#define EAOF_HISTO_LENGTH 30  /* ORBmatcher::HISTO_LENGTH src/ORBmatcher.cc:39 */

enum {
    EAOF_BOW_KF_FRAME = 0, /* SearchByBoW(KeyFrame*, Frame&, ...)    src/ORBmatcher.cc:159-288: best <= TH_LOW,
                              output per TARGET (Frame) feature = index of the matched query (KeyFrame) feature */
    EAOF_BOW_KF_KF = 1     /* SearchByBoW(KeyFrame*, KeyFrame*, ...) src/ORBmatcher.cc:522-655: best <  TH_LOW, targets
                              need valid_t, output per QUERY feature = index of the matched target feature */
};

typedef struct eaof_matcher eaof_matcher;

/* One handle = one stream + workspace for up to max_pairs pairs of up to max_features features each. */
int eaof_matcher_create(int device, int max_pairs, int max_features, eaof_matcher** out);
void eaof_matcher_destroy(eaof_matcher* m);
void* eaof_matcher_stream(eaof_matcher* m);
int eaof_matcher_sync(eaof_matcher* m);

/* static int ORBmatcher::DescriptorDistance(a, b)  src/ORBmatcher.cc:1649-1665, for n descriptor pairs (host
 * buffers; a and b hold n x 32 bytes). */
int eaof_hamming_distances(eaof_matcher* m, const uint8_t* a, const uint8_t* b, int n, int* dist_out);

/* SearchByBoW for one pair, HOST buffers.  Q = pKF/pKF1 (outer loop), T = F/pKF2 (inner loop).
 * valid_q / valid_t: 1 where the feature has a map point that is not bad (NULL = all valid; valid_t is only
 * consulted in EAOF_BOW_KF_KF mode, as in the reference).  Node lists: node ids ascending, start arrays have
 * n_nodes+1 entries, idx arrays list feature indices in FeatureVector order.  match_out / dist_out: n_t entries in
 * KF_FRAME mode, n_q entries in KF_KF mode (-1 = no match).  *n_matches = the function's return value. */
int eaof_match_bow(eaof_matcher* m, int mode, float nnratio, int check_orientation, int n_q, const uint8_t* desc_q,
                   const float* angle_q, const uint8_t* valid_q, int n_t, const uint8_t* desc_t, const float* angle_t,
                   const uint8_t* valid_t, int n_nodes_q, const int* node_id_q, const int* node_start_q,
                   const int* node_idx_q, int n_nodes_t, const int* node_id_t, const int* node_start_t,
                   const int* node_idx_t, int* match_out, int* dist_out, int* n_matches);

/* SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)  src/ORBmatcher.cc:657-823 (with
 * CheckDistEpipolarLine :140-157), one pair, HOST buffers.  free1/free2: 1 where the feature has NO map point
 * (only those are matched); stereo1/stereo2: mvuRight >= 0 (NULL = monocular).  Feature vectors as CSR like
 * eaof_match_bow.  F12: 3x3 row-major; (ex, ey): epipole of KF1's centre in KF2 (:664-670, computed by the caller);
 * scale_factors2 / level_sigma2_2: mvScaleFactors / mvLevelSigma2 of KF2.  match12 / dist12: n1 entries (index of
 * the KF2 feature, -1 = none) — vMatchedPairs is the list of (i, match12[i]) with match12[i] >= 0 in ascending i. */
int eaof_match_triangulation(eaof_matcher* m, int check_orientation, int only_stereo, int n1, const uint8_t* desc1,
                             const float* x1, const float* y1, const float* angle1, const uint8_t* free1,
                             const uint8_t* stereo1, int n2, const uint8_t* desc2, const float* x2, const float* y2,
                             const int* octave2, const float* angle2, const uint8_t* free2, const uint8_t* stereo2,
                             int n_nodes1, const int* node_id1, const int* node_start1, const int* node_idx1,
                             int n_nodes2, const int* node_id2, const int* node_start2, const int* node_idx2,
                             const float* F12, float ex, float ey, const float* scale_factors2,
                             const float* level_sigma2_2, int n_levels, int* match12, int* dist12, int* n_matches);

/* SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)  src/ORBmatcher.cc:1328-1472, one pair, HOST buffers.
 * Cur: undistorted keypoint positions/octaves/angles, descriptors, optional mvuRight (NULL = monocular) and
 * optional `taken` flags (Cur feature already holds a map point with observations).  The 64x48 grid of
 * Frame::AssignFeaturesToGrid (src/Frame.cc:599-614) is rebuilt on the device from (min_x, min_y, grid_inv_w,
 * grid_inv_h).  Last: per feature the projected pixel (u, v) and 1/z computed by the caller (src/ORBmatcher.cc:1364-1377
 * stays on the host), validity (map point present and not an outlier), octave, angle, the map point's descriptor
 * and obs = pMP->Observations()>0 (NULL = all).  search_mode 0: octave-1..octave+1, 1: bForward, 2: bBackward.
 * match_cur / dist_cur: n_cur entries (index of the Last feature, -1 = none).  dist_cur is -2 where a match was made and
 * then removed by the rotation check: the reference leaves NULL in mvpMapPoints there (:1463), not the previous pointer. */
int eaof_match_projection(eaof_matcher* m, int n_cur, const float* cur_x, const float* cur_y, const int* cur_octave,
                          const float* cur_angle, const uint8_t* cur_desc, const float* cur_uright,
                          const uint8_t* cur_taken, float min_x, float max_x, float min_y, float max_y,
                          float grid_inv_w, float grid_inv_h, int n_last, const uint8_t* last_valid,
                          const float* last_u, const float* last_v, const float* last_invz, const int* last_octave,
                          const float* last_angle, const uint8_t* last_desc, const uint8_t* last_obs,
                          const float* scale_factors, int n_levels, float th, float mbf, int search_mode,
                          int check_orientation, int* match_cur, int* dist_cur, int* n_matches);

/* Window search on the 64x48 feature grid — the common shape of the projection matchers other than (Cur,Last)
 * (SURVEY.md §8 a14), one call, HOST buffers.  Targets = the features of a Frame (undistorted position, octave, angle,
 * descriptor, optional mvuRight and `taken` flags); queries = map points already projected by the caller: pixel
 * (q_u, q_v), search radius, level window [q_min_level, q_max_level] as passed to Frame::GetFeaturesInArea
 * (src/Frame.cc:696-749), optional predicted right coordinate q_ur (stereo gate |q_ur - mvuRight| > radius), angle,
 * descriptor, q_obs = pMP->Observations()>0.  Queries are processed in index order with the reference's greedy
 * exclusion (a target that received a map point with observations is no longer a candidate).
 *   rule EAOF_WIN_BEST: best distance only, accept best <= th_accept.
 *        SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist)  src/ORBmatcher.cc:1474-1601:
 *        q_valid = map point present, not bad, not already found, distance inside the scale pyramid; radius =
 *        th*mvScaleFactors[level]; levels level-1..level+1; ttaken = mvpMapPoints[i] != NULL; th_accept = ORBdist;
 *        hist_mode 1; check_bounds 1.
 *   rule EAOF_WIN_RATIO_SAME_LEVEL: best and second best; rejected when both lie on the same level and
 *        best > nnratio*second.  SearchByProjection(Frame&, const vector<MapPoint*>&, th)  src/ORBmatcher.cc:45-129:
 *        q_valid = mbTrackInView && !isBad(); radius = RadiusByViewingCos(mTrackViewCos)*th*mvScaleFactors[level];
 *        levels level-1..level; q_ur = mTrackProjXR; th_accept = TH_HIGH; hist_mode 0; check_bounds 0.  *n_matches
 *        counts accepted queries like the reference's return value.
 * hist_mode: 0 none, 1 factor 1/HISTO_LENGTH, 2 factor HISTO_LENGTH/360.
 * match_t / dist_t: n_t entries (index of the query matched to the target, -1 = none; dist_t -2 = matched, then pruned
 * by the rotation check). */
enum { EAOF_WIN_BEST = 0, EAOF_WIN_RATIO_SAME_LEVEL = 1 };
int eaof_match_windows(eaof_matcher* m, int rule, int n_t, const float* t_x, const float* t_y, const int* t_octave,
                       const float* t_angle, const uint8_t* t_desc, const float* t_uright, const uint8_t* t_taken,
                       float min_x, float max_x, float min_y, float max_y, float grid_inv_w, float grid_inv_h, int n_q,
                       const uint8_t* q_valid, const float* q_u, const float* q_v, const float* q_radius,
                       const int* q_min_level, const int* q_max_level, const float* q_ur, const float* q_angle,
                       const uint8_t* q_desc, const uint8_t* q_obs, int th_accept, float nnratio, int hist_mode,
                       int check_bounds, int* match_t, int* dist_t, int* n_matches);

/* Independent window queries on the 64x48 grid of a KeyFrame (KeyFrame::GetFeaturesInArea, src/KeyFrame.cc:608-647) —
 * the search step of the map-side matchers whose loop never excludes a target because an earlier query took it, one
 * call, HOST buffers.  Per query (a map point projected by the caller): the best target by Hamming distance among
 * the window candidates on levels [q_min_level, q_max_level], first candidate in the reference's visiting order on
 * ties, accepted when best <= th_accept.  match_q / dist_q: n_q entries (index of the target, -1 = none); *n_matches
 * = number of accepted queries.  The caller applies the map mutations in query order.
 *   Fuse(KeyFrame*, const vector<MapPoint*>&, th)            src/ORBmatcher.cc:825-961: gate EAOF_GATE_FUSE_CHI2 (a
 *        candidate is dropped when its reprojection error e2*mvInvLevelSigma2[level] exceeds 7.8 with mvuRight >= 0 —
 *        e2 includes (q_ur - mvuRight)^2 — or 5.99 without, :901-925), levels level-1..level, th_accept = TH_LOW.
 *   Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)       :963-1100: no gate, th_accept = TH_LOW.
 *   SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) :1102-1326: once per direction, no gate, th_accept =
 *        TH_HIGH; the mutual-agreement pass (:1308-1323) is the caller's.
 * (SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) :290-403 does exclude matched targets: it is
 * eaof_match_windows with rule EAOF_WIN_BEST, t_taken = vpMatched[idx] != NULL, hist_mode 0, check_bounds 0.)
 * t_uright (mvuRight, NULL = all monocular) and inv_level_sigma2 / q_ur are only read by the chi-square gate. */
enum { EAOF_GATE_NONE = 0, EAOF_GATE_FUSE_CHI2 = 1 };
int eaof_match_windows_independent(eaof_matcher* m, int gate, int n_t, const float* t_x, const float* t_y,
                                   const int* t_octave, const uint8_t* t_desc, const float* t_uright, float min_x,
                                   float min_y, float grid_inv_w, float grid_inv_h, const float* inv_level_sigma2,
                                   int n_levels, int n_q, const uint8_t* q_valid, const float* q_u, const float* q_v,
                                   const float* q_radius, const int* q_min_level, const int* q_max_level,
                                   const float* q_ur, const uint8_t* q_desc, int th_accept, int* match_q, int* dist_q,
                                   int* n_matches);

/* Batched MapPoint::ComputeDistinctiveDescriptors  src/MapPoint.cc:242-307, HOST buffers: map point p owns the
 * observation descriptors rows [start[p], start[p+1]) of `desc`; for each point the N x N Hamming distances, the
 * median of every row (sorted row, element 0.5*(N-1) truncated, :288-291) and the FIRST row with the smallest median
 * (:293-297).  best_out[p] = index of that row relative to start[p] (-1 for a point without observations);
 * median_out (optional) = its median. */
int eaof_distinctive_descriptors(eaof_matcher* m, int n_points, const int* start, const uint8_t* desc, int* best_out,
                                 int* median_out);

/* SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  src/ORBmatcher.cc:405-520, HOST buffers.
 * F1: octave, angle, descriptor of every feature (only level-0 features are matched) and prev_matched (x,y pairs, in/out:
 * updated to the matched F2 position like vbPrevMatched).  F2: undistorted positions, octaves, angles, descriptors; its
 * 64x48 grid is rebuilt on the device.  A later F1 feature takes over an F2 feature when its distance is strictly
 * smaller (the earlier match is undone).  matches12: n1 entries. */
int eaof_match_initialization(eaof_matcher* m, float nnratio, int check_orientation, int n1, const int* octave1,
                              const float* angle1, const uint8_t* desc1, float* prev_matched, int n2, const float* x2,
                              const float* y2, const int* octave2, const float* angle2, const uint8_t* desc2, float min_x,
                              float max_x, float min_y, float max_y, float grid_inv_w, float grid_inv_h, int window_size,
                              int* matches12, int* n_matches);

/* ---- batched, device-resident forms used for sequences (BASELINE.json configs[1] and [4]) ------------------ */

/* Consecutive-frame SearchByProjection over the results an extractor handle holds on the device: pair p matches
 * Cur = frame cur_frame[p] against Last = frame last_frame[p]; the "projection" of Last keypoint i is its own
 * position shifted by (shift_x[p], shift_y[p]) (known inter-frame motion plays the role of Rcw*X+tcw), every Last
 * feature counts as a map point with observations, monocular, search_mode 0, orientation check on.
 * d_match / d_dist: n_pairs x cap ints on the device, d_nmatches: n_pairs ints.  Asynchronous on the matcher's
 * stream after waiting for the extractor's stream. */
int eaof_match_projection_batch_device(eaof_matcher* m, eaof_orb* ex, int n_pairs, const int* last_frame,
                                       const int* cur_frame, const float* shift_x, const float* shift_y, float th,
                                       int* d_match, int* d_dist, int* d_nmatches);

/* Brute-force SearchByBoW semantics (one vocabulary node holding every feature) over pairs of descriptor blocks
 * resident on the device: block f = d_desc + f*block_stride*32 with d_counts[f] features and angles
 * d_angle + f*block_stride.  Pair p: queries = block pair_q[p], targets = block pair_t[p].  Outputs as
 * eaof_match_bow, laid out [pair][block_stride].  pair_q/pair_t are host arrays.
 * The distances are computed on the tensor cores (tcgen05.mma kind::i8 on a +-1 expansion of blocks 0..max index, made on
 * the matcher's stream at every call since the caller may rewrite the array in between; 256 B of scratch per descriptor)
 * when the pairs outnumber the blocks they touch, by the XOR + POPC kernel otherwise or with EAOF_BOW_UMMA=0; both give the
 * same matches. */
int eaof_match_bruteforce_batch_device(eaof_matcher* m, int mode, float nnratio, int check_orientation, int n_pairs,
                                       const int* pair_q, const int* pair_t, const uint8_t* d_desc,
                                       const float* d_angle, const int* d_counts, int block_stride, int* d_match,
                                       int* d_dist, int* d_nmatches);

/* SearchByBoW over frames that never leave the device: descriptors and keypoint angles of the extractor handle's last
 * batch (frames 0..n_frames-1), FeatureVectors as eaof_voc_transform_orb_device (include/eaof_voc.h) left them
 * (d_n_fnodes[f], d_node_ids[f*cap + j], d_node_start[f*(cap+1) + j], d_feat_idx[f*cap + i], cap =
 * eaof_orb_max_keypoints).  Pair p: queries = frame pair_q[p] (the KeyFrame side), targets = frame pair_t[p]; every
 * feature counts as holding a good map point.  The merge-walk of the two FeatureVectors (src/ORBmatcher.cc:182-264) is
 * planned on the device.  Outputs as eaof_match_bow, laid out [pair][cap].  Asynchronous on the matcher's stream after
 * the extractor's; the caller orders it after the vocabulary's stream (eaof_voc_sync, or an event). */
int eaof_match_bow_orb_device(eaof_matcher* m, eaof_orb* ex, int n_frames, int mode, float nnratio, int check_orientation,
                              int n_pairs, const int* pair_q, const int* pair_t, const int* d_n_fnodes,
                              const uint32_t* d_node_ids, const int* d_node_start, const uint32_t* d_feat_idx, int* d_match,
                              int* d_dist, int* d_nmatches);

/* Micro-benchmark: POPC32 thread-instructions per second this device sustains with every SM busy (the issue-rate
 * roofline of the Hamming kernels); negative on error. */
double eaof_debug_popc_rate(int device);

/* Number of descriptor-pair distances the last batched call evaluated (for the matches/s metric). */
long long eaof_matcher_last_distance_count(const eaof_matcher* m);

#ifdef __cplusplus
}
#endif
#endif /* EAOF_MATCH_H */
