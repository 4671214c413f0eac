// ORACLE — test infrastructure only.  Nothing here is linked into, imported by or executed from the product path.
//
// Restatement, on plain arrays, of DBoW2's bag-of-words conversion as the reference vendors it
// (/root/reference/Thirdparty/DBoW2/DBoW2): TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)
// TemplatedVocabulary.h:1138-1205, the single-feature tree descent :1230-1271, FORB::distance FORB.cpp:81-101,
// BowVector::addWeight / addIfNotExist / normalize BowVector.cpp:33-91, FeatureVector::addFeature FeatureVector.cpp:33-47.
// The tree is given as arrays (node 0 = root; children of node i = childIdx[childStart[i] .. childStart[i+1]) in the
// order of m_nodes[i].children; 32-byte descriptors; double weights; wordId >= 0 on leaves).
//
// PARITY STATUS: pinned.  The vendored DBoW2 compiles here unmodified (oracle/voc_ref_harness.cc ->
// oracle/_ref/libvoc_ref.so); tests/test_oracle_voc_vs_ref.py loads synthetic vocabularies with the reference's own
// loadFromTextFile, dumps the tree it built, and requires identical word ids, bit-identical double word values and
// identical feature vectors from both.  The reference's ORBvoc file itself is not in the tree (parity on it: not run).
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

namespace {
// FORB::distance (FORB.cpp:81-101): the same SWAR popcount as ORBmatcher::DescriptorDistance, returned as double
int forb_distance(const uint8_t* a, const uint8_t* b) {
    const int32_t* pa = reinterpret_cast<const int32_t*>(a);
    const int32_t* pb = reinterpret_cast<const int32_t*>(b);
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}
enum { TF_IDF = 0, TF = 1, IDF = 2, BINARY = 3 };                                            // BowVector.h WeightingType
enum { L1_NORM = 0, L2_NORM = 1, CHI_SQUARE = 2, KL = 3, BHATTACHARYYA = 4, DOT_PRODUCT = 5 };  // ScoringType
}  // namespace

extern "C" {

// Single feature, TemplatedVocabulary.h:1230-1271.  Returns the leaf node; *nid = node on the path at level L-levelsup.
int eaoo_voc_descend(int L, const int* childStart, const int* childIdx, const uint8_t* nodeDesc, const uint8_t* feature,
                     int levelsup, int* nid) {
    const int nid_level = L - levelsup;
    if (nid_level <= 0 && nid) *nid = 0;  // root
    int final_id = 0, current_level = 0;
    do {
        ++current_level;
        const int b = childStart[final_id], e = childStart[final_id + 1];
        final_id = childIdx[b];
        double best_d = forb_distance(feature, nodeDesc + 32 * (size_t)final_id);
        for (int j = b + 1; j < e; ++j) {
            const int id = childIdx[j];
            const double d = forb_distance(feature, nodeDesc + 32 * (size_t)id);
            if (d < best_d) { best_d = d; final_id = id; }
        }
        if (nid && current_level == nid_level) *nid = final_id;
    } while (childStart[final_id + 1] > childStart[final_id]);  // !isLeaf()
    return final_id;
}

// transform(features, v, fv, levelsup), :1138-1205.  Outputs in std::map order: counts[0] words (wordIds, wordVals),
// counts[1] feature-vector nodes (nodeIds, nodeStart with counts[1]+1 entries, featIdx).
void eaoo_voc_transform(int L, int nNodes, const int* childStart, const int* childIdx, const uint8_t* nodeDesc,
                        const double* weight, const int* wordId, int weighting, int scoring, int nFeat,
                        const uint8_t* feat, int levelsup, int* counts, unsigned* wordIds, double* wordVals,
                        unsigned* nodeIds, int* nodeStart, unsigned* featIdx) {
    std::map<unsigned, double> v;
    std::map<unsigned, std::vector<unsigned>> fv;
    counts[0] = counts[1] = 0;
    nodeStart[0] = 0;
    if (nNodes <= 1) return;  // empty(): m_words.empty()
    const bool must = scoring != DOT_PRODUCT;                 // ScoringObject.h:74-89
    const bool l2 = scoring == L2_NORM;
    for (int i = 0; i < nFeat; ++i) {
        int nid = 0;
        const int leaf = eaoo_voc_descend(L, childStart, childIdx, nodeDesc, feat + 32 * (size_t)i, levelsup, &nid);
        const unsigned id = (unsigned)wordId[leaf];
        const double w = weight[leaf];
        if (w > 0) {  // not stopped
            if (weighting == TF || weighting == TF_IDF) {
                auto it = v.find(id);                          // addWeight
                if (it != v.end()) it->second += w; else v[id] = w;
            } else if (!v.count(id)) v[id] = w;                // addIfNotExist
            fv[(unsigned)nid].push_back((unsigned)i);
        }
    }
    if ((weighting == TF || weighting == TF_IDF) && !v.empty() && !must) {
        const double nd = v.size();
        for (auto& e : v) e.second /= nd;
    }
    if (must) {  // BowVector::normalize
        double norm = 0.0;
        if (!l2) for (auto& e : v) norm += fabs(e.second);
        else { for (auto& e : v) norm += e.second * e.second; norm = sqrt(norm); }
        if (norm > 0.0) for (auto& e : v) e.second /= norm;
    }
    int w = 0;
    for (auto& e : v) { wordIds[w] = e.first; wordVals[w] = e.second; ++w; }
    int a = 0, o = 0;
    for (auto& e : fv) {
        nodeIds[a] = e.first;
        nodeStart[a++] = o;
        for (unsigned f : e.second) featIdx[o++] = f;
    }
    nodeStart[a] = o;
    counts[0] = w;
    counts[1] = a;
}

}  // extern "C"
