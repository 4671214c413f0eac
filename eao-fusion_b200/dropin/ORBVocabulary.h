/**
 * Drop-in replacement for EAO-Fusion's include/ORBVocabulary.h: ORB_SLAM2::ORBVocabulary keeps being a DBoW2
 * TemplatedVocabulary<FORB::TDescriptor, FORB> (loading, saving, scoring, the single-feature transforms are the vendored
 * DBoW2 code), but the virtual
 *     transform(const std::vector<TDescriptor>& features, BowVector& v, FeatureVector& fv, int levelsup)
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:145-146, :1138-1205) — what Frame::ComputeBoW / KeyFrame::ComputeBoW call
 * (src/Frame.cc:764-771, src/KeyFrame.cc:93-102) — runs on the GPU through libeaof_orb.so (include/eaof_voc.h).
 * No source change in System.cc / Frame.cc / KeyFrame.cc: `new ORBVocabulary()` + loadFromTextFile / loadFromBinaryFile
 * work as before; the tree is handed to the device on the first transform.  No CPU fallback: a library failure throws.
 */
#ifndef ORBVOCABULARY_H
#define ORBVOCABULARY_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "FORB.h"
#include "TemplatedVocabulary.h"
#include "eaof_voc.h"

namespace ORB_SLAM2
{

class EaofORBVocabulary : public DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB>
{
    typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Base;

public:
    EaofORBVocabulary(int k = 10, int L = 5, DBoW2::WeightingType weighting = DBoW2::TF_IDF,
                      DBoW2::ScoringType scoring = DBoW2::L1_NORM)
        : Base(k, L, weighting, scoring), mpVoc(NULL), mnCap(0) {}
    virtual ~EaofORBVocabulary() { Release(); }

    using Base::transform;  // the overloads that stay on the host

    virtual void transform(const std::vector<DBoW2::FORB::TDescriptor>& features, DBoW2::BowVector& v,
                           DBoW2::FeatureVector& fv, int levelsup) const
    {
        v.clear();
        fv.clear();
        if(this->empty())
            return;
        const int n = (int)features.size();
        std::vector<unsigned char> desc(32 * (size_t)(n ? n : 1));
        for(int i = 0; i < n; i++)
            memcpy(&desc[32 * (size_t)i], features[i].ptr<unsigned char>(), 32);
        const int setStart[2] = {0, n};
        int nWords = 0, nNodes = 0;
        std::vector<uint32_t> wordIds(n + 1), nodeIds(n + 1), featIdx(n + 1);
        std::vector<double> wordVals(n + 1);
        std::vector<int> nodeStart(n + 2);
        {
            // Tracking (Frame::ComputeBoW) and LocalMapping / LoopClosing (KeyFrame::ComputeBoW) share the one vocabulary: the
            // handle lookup — which may destroy and re-create the device copy for a larger frame — and the call that uses the
            // handle (one stream / staging area per vocabulary) sit under the same lock
            std::unique_lock<std::mutex> lock(mMutex);
            eaof_voc* voc = DeviceLocked(n);
            if(eaof_voc_transform(voc, 1, setStart, &desc[0], levelsup, &nWords, &wordIds[0], &wordVals[0], &nNodes,
                                  &nodeIds[0], &nodeStart[0], &featIdx[0]) != EAOF_OK)
                Throw("eaof_voc_transform");
        }
        for(int j = 0; j < nWords; j++)
            v.insert(v.end(), DBoW2::BowVector::value_type(wordIds[j], wordVals[j]));
        for(int j = 0; j < nNodes; j++)
        {
            DBoW2::FeatureVector::iterator it =
                fv.insert(fv.end(), DBoW2::FeatureVector::value_type(nodeIds[j], std::vector<unsigned int>()));
            it->second.assign(featIdx.begin() + nodeStart[j], featIdx.begin() + nodeStart[j + 1]);
        }
    }

    /// Hands the loaded tree to the device now (otherwise done by the first transform).
    void Upload(int maxFeatures = 4096) const
    {
        std::unique_lock<std::mutex> lock(mMutex);
        DeviceLocked(maxFeatures);
    }

private:
    static void Throw(const char* what)
    {
        std::string msg = std::string("ORBVocabulary(eaof): ") + what + ": " + eaof_last_error();
        fprintf(stderr, "%s\n", msg.c_str());
        throw std::runtime_error(msg);
    }

    void Release() const
    {
        if(mpVoc)
            eaof_voc_destroy(mpVoc);
        mpVoc = NULL;
    }

    // caller holds mMutex
    eaof_voc* DeviceLocked(int nFeatures) const
    {
        if(mpVoc && nFeatures <= mnCap)
            return mpVoc;
        const int kLibraryMax = 12000;  // eaof_voc_create's max_features limit
        if(nFeatures > kLibraryMax)
        {
            const std::string msg = "ORBVocabulary(eaof): a frame with " + std::to_string(nFeatures) +
                                    " features exceeds the 12000 a device transform takes (eaof_voc_create max_features)";
            fprintf(stderr, "%s\n", msg.c_str());
            throw std::runtime_error(msg);
        }
        Release();
        int cap = 4096;
        while(cap < nFeatures)
            cap *= 2;
        if(cap > kLibraryMax)
            cap = kLibraryMax;
        const size_t n = this->m_nodes.size();
        std::vector<int> childStart(n + 1), childIdx(n ? n - 1 : 0), wordId(n, -1);
        std::vector<unsigned char> desc(32 * n, 0);
        std::vector<double> weight(n, 0.0);
        size_t o = 0;
        for(size_t i = 0; i < n; i++)
        {
            const Node& nd = this->m_nodes[i];
            childStart[i] = (int)o;
            for(size_t c = 0; c < nd.children.size() && o < childIdx.size(); c++)
                childIdx[o++] = (int)nd.children[c];
            if(i && !nd.descriptor.empty())
                memcpy(&desc[32 * i], nd.descriptor.template ptr<unsigned char>(), 32);
            weight[i] = nd.weight;
            if(i && nd.isLeaf())
                wordId[i] = (int)nd.word_id;
        }
        childStart[n] = (int)o;
        const char* dev = getenv("EAOF_DEVICE");
        if(eaof_voc_create(dev && *dev ? atoi(dev) : 0, this->m_L, (int)n, &childStart[0], childIdx.empty() ? NULL : &childIdx[0],
                           &desc[0], &weight[0], &wordId[0], (int)this->m_weighting, (int)this->m_scoring, cap, 1,
                           &mpVoc) != EAOF_OK)
            Throw("eaof_voc_create");
        mnCap = cap;
        return mpVoc;
    }

    mutable std::mutex mMutex;
    mutable eaof_voc* mpVoc;
    mutable int mnCap;
};

typedef EaofORBVocabulary ORBVocabulary;

} //namespace ORB_SLAM

#endif // ORBVOCABULARY_H
