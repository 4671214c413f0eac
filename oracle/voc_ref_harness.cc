// ORACLE — test infrastructure only (never linked into the product library).
//
// C harness around the UNMODIFIED vendored DBoW2 of the reference (/root/reference/Thirdparty/DBoW2: TemplatedVocabulary.h,
// FORB.cpp, BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp, DUtils/Random.cpp, DUtils/Timestamp.cpp), compiled in
// place by `make -C oracle vocref` against oracle/matchshim/mcv.h into oracle/_ref/libvoc_ref.so.  It pins the
// restatement of ORBVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)
// (TemplatedVocabulary.h:1138-1205, 1230-1271; called by Frame::ComputeBoW src/Frame.cc:764-771 and KeyFrame::ComputeBoW
// src/KeyFrame.cc:93-102 with levelsup = 4) to the reference's code as run here.  The reference's vocabulary file
// (ORBvoc) is not in the tree (.MISSING_LARGE_BLOBS); tests build synthetic trees in the same text format and load them
// with the reference's own loadFromTextFile.
//
// With -DEAOF_VOC_DROPIN the same harness instantiates the drop-in subclass (eao-fusion_b200/dropin/ORBVocabulary.h),
// whose transform runs on the GPU: tests/cpp/_build/libvoc_dropin.so.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "FORB.h"
#include "TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVocabulary;

#ifdef EAOF_VOC_DROPIN
#include "ORBVocabulary.h"  // eao-fusion_b200/dropin
typedef ORB_SLAM2::EaofORBVocabulary VocImpl;
#else
typedef RefVocabulary VocImpl;
#endif

namespace {
struct Peek : public VocImpl {  // read access to the protected tree
    using VocImpl::m_nodes;
    using VocImpl::m_words;
};
}  // namespace

extern "C" {

void* vref_load(const char* path) {
    VocImpl* v = new VocImpl();
    if (!v->loadFromTextFile(path) || v->empty()) { delete v; return nullptr; }
#ifdef EAOF_VOC_DROPIN
    v->Upload();
#endif
    return v;
}
void vref_free(void* h) { delete static_cast<VocImpl*>(h); }

// k, L, number of nodes (incl. root), number of words, scoring type, weighting type
void vref_info(void* h, int* out6) {
    VocImpl* v = static_cast<VocImpl*>(h);
    Peek* p = static_cast<Peek*>(v);
    out6[0] = v->getBranchingFactor(); out6[1] = v->getDepthLevels(); out6[2] = (int)p->m_nodes.size();
    out6[3] = (int)v->size(); out6[4] = (int)v->getScoringType(); out6[5] = (int)v->getWeightingType();
}

// The tree as the reference holds it after loading: per node parent id, 32 descriptor bytes, weight, word id (-1 for
// inner nodes), children in m_nodes[i].children order as a CSR (childStart has n+1 entries).
void vref_tree(void* h, int* parent, uint8_t* desc, double* weight, int* wordId, int* childStart, int* childIdx) {
    Peek* p = static_cast<Peek*>(static_cast<VocImpl*>(h));
    const size_t n = p->m_nodes.size();
    int o = 0;
    for (size_t i = 0; i < n; ++i) {
        const auto& nd = p->m_nodes[i];
        parent[i] = i ? (int)nd.parent : -1;
        if (!nd.descriptor.empty()) memcpy(desc + 32 * i, nd.descriptor.data, 32); else memset(desc + 32 * i, 0, 32);
        weight[i] = nd.weight;
        wordId[i] = nd.isLeaf() && i ? (int)nd.word_id : -1;
        childStart[i] = o;
        for (auto c : nd.children) childIdx[o++] = (int)c;
    }
    childStart[n] = o;
}

// transform(features, v, fv, levelsup).  Outputs: BowVector as (word id, value) in map order, FeatureVector as CSR in
// map order (node ids, starts with nFNodes+1 entries, feature indices).  counts[0] = number of words, counts[1] = nodes.
void vref_transform(void* h, int n, const uint8_t* desc, int levelsup, int* counts, unsigned* wordIds, double* wordVals,
                    unsigned* nodeIds, int* nodeStart, unsigned* featIdx) {
    VocImpl* v = static_cast<VocImpl*>(h);
    std::vector<cv::Mat> feats(n);
    for (int i = 0; i < n; ++i) {
        feats[i].create(1, 32, CV_8U);
        memcpy(feats[i].data, desc + 32 * (size_t)i, 32);
    }
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    static_cast<RefVocabulary*>(v)->transform(feats, bv, fv, levelsup);  // through the base class: virtual dispatch
    int w = 0;
    for (auto& e : bv) { wordIds[w] = e.first; wordVals[w] = e.second; ++w; }
    int a = 0, o = 0;
    for (auto& e : fv) {
        nodeIds[a] = e.first;
        nodeStart[a++] = o;
        for (auto f : e.second) featIdx[o++] = f;
    }
    nodeStart[a] = o;
    counts[0] = w;
    counts[1] = a;
}

// BowVector similarity of the vocabulary's scoring object (ORBVocabulary::score, used by KeyFrameDatabase), for two
// vectors given as (ids, values): lets the tests check that transform outputs are usable as the reference uses them.
double vref_score(void* h, int n1, const unsigned* id1, const double* v1, int n2, const unsigned* id2, const double* v2) {
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; ++i) a.insert(a.end(), DBoW2::BowVector::value_type(id1[i], v1[i]));
    for (int i = 0; i < n2; ++i) b.insert(b.end(), DBoW2::BowVector::value_type(id2[i], v2[i]));
    return static_cast<VocImpl*>(h)->score(a, b);
}

}  // extern "C"
