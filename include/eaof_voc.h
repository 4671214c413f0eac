/*
 * eaof_voc.h — C ABI of the bag-of-words conversion in libeaof_orb.so (B200, sm_100a).
 *
 * Replaces ORBVocabulary::transform(features, BowVector&, FeatureVector&, levelsup) of the DBoW2 the reference vendors
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1138-1205 with the per-feature tree descent :1230-1271 and
 * FORB::distance FORB.cpp:81-101; include/ORBVocabulary.h:32-33), as called by Frame::ComputeBoW (src/Frame.cc:764-771)
 * and KeyFrame::ComputeBoW (src/KeyFrame.cc:93-102) with levelsup = 4.  The vocabulary tree crosses the boundary as plain
 * arrays taken from the loaded TemplatedVocabulary (the drop-in subclass eao-fusion_b200/dropin/ORBVocabulary.h does
 * that); loading / saving / scoring stay on the reference host path.  Results are bit-exact: word ids, the double word
 * values (accumulated and normalised in the reference's order), feature-vector node ids and feature order.
 *
 * Same conventions as eaof_orb.h: 0 / negative EAOF_ERR_*, eaof_last_error(), no CPU fallback.
 */
#ifndef EAOF_VOC_H
#define EAOF_VOC_H

#include <stddef.h>
#include <stdint.h>

#include "eaof_orb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* DBoW2::WeightingType / ScoringType  Thirdparty/DBoW2/DBoW2/BowVector.h:29-49 */
enum { EAOF_VOC_TF_IDF = 0, EAOF_VOC_TF = 1, EAOF_VOC_IDF = 2, EAOF_VOC_BINARY = 3 };
enum {
    EAOF_VOC_L1_NORM = 0, EAOF_VOC_L2_NORM = 1, EAOF_VOC_CHI_SQUARE = 2, EAOF_VOC_KL = 3, EAOF_VOC_BHATTACHARYYA = 4,
    EAOF_VOC_DOT_PRODUCT = 5
};

typedef struct eaof_voc eaof_voc;

/* The tree of a loaded vocabulary (m_nodes, TemplatedVocabulary.h:435): node 0 is the root; the children of node i are
 * child_idx[child_start[i] .. child_start[i+1]) in the order of m_nodes[i].children (child_start has n_nodes+1 entries);
 * node_desc: n_nodes x 32 bytes (the root's row is ignored); weight / word_id: per node, read on leaves only.  L = m_L.
 * max_features = largest descriptor set a call will pass, max_sets = sets per call. */
int eaof_voc_create(int device, int L, int n_nodes, const int* child_start, const int* child_idx, const uint8_t* node_desc,
                    const double* weight, const int* word_id, int weighting, int scoring, int max_features, int max_sets,
                    eaof_voc** out);
void eaof_voc_destroy(eaof_voc* v);

/* transform() for n_sets descriptor sets (frames), HOST buffers: set s = rows [set_start[s], set_start[s+1]) of desc.
 * Outputs, all indexed with the set's own offset o = set_start[s]:
 *   BowVector      n_words[s] entries word_ids[o + j] (ascending) / word_vals[o + j];
 *   FeatureVector  n_fnodes[s] entries node_ids[o + j] (ascending) with features
 *                  feat_idx[o + node_start[o + s + j] .. o + node_start[o + s + j + 1]) — indices relative to the set, in
 *                  ascending order; node_start holds n_fnodes[s] + 1 entries per set (so the array needs total + n_sets).
 * Features whose word has weight <= 0 ("stopped") are left out of both, as in the reference. */
int eaof_voc_transform(eaof_voc* v, int n_sets, const int* set_start, const uint8_t* desc, int levelsup, int* n_words,
                       uint32_t* word_ids, double* word_vals, int* n_fnodes, uint32_t* node_ids, int* node_start,
                       uint32_t* feat_idx);

/* The same over the descriptors an extractor handle holds on the device after a batched call (frame f = one set), results
 * left in device memory with a per-frame stride of cap = eaof_orb_max_keypoints(): d_n_words[f], d_word_ids[f*cap + j],
 * d_word_vals[f*cap + j], d_n_fnodes[f], d_node_ids[f*cap + j], d_node_start[f*(cap+1) + j], d_feat_idx[f*cap + i].
 * Asynchronous on the vocabulary's stream after waiting for the extractor's; eaof_voc_sync() waits for it. */
int eaof_voc_transform_orb_device(eaof_voc* v, eaof_orb* ex, int n_frames, int levelsup, int* d_n_words,
                                  uint32_t* d_word_ids, double* d_word_vals, int* d_n_fnodes, uint32_t* d_node_ids,
                                  int* d_node_start, uint32_t* d_feat_idx);
int eaof_voc_sync(eaof_voc* v);
void* eaof_voc_stream(eaof_voc* v);

#ifdef __cplusplus
}
#endif
#endif /* EAOF_VOC_H */
