// Bag-of-words conversion of libeaof_orb.so (include/eaof_voc.h): ORBVocabulary::transform of the reference's vendored
// DBoW2 (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1138-1205, :1230-1271) as two kernels.
//
//   k_voc_descend   a group of 16 (k <= 16) or 32 lanes walks one descriptor down the tree: lane j takes child j of the
//                   current node (children are stored next to each other, so a level is one coalesced 32*k byte read),
//                   Hamming distance on POPC, group argmin by (distance, child order) = the reference's first-minimum
//                   rule.  Output per feature: word id, leaf (for its weight), node at level L - levelsup.
//   k_voc_assemble  one CTA per descriptor set builds both std::map-ordered outputs without sorting networks: the rank of
//                   a feature's key (id << 32 | feature index) among the set's keys is counted directly (n^2/1024
//                   shared-memory reads per thread, n ~ 1000), run heads give the distinct ids.  Word values are doubles
//                   accumulated and normalised in the reference's order (BowVector::addWeight in feature order,
//                   normalize in ascending word order), so one thread does the norm sum sequentially: bit-exact.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/eaof_voc.h"

extern "C" int eaof_internal_fail(int code, const char* msg);
extern "C" int eaof_internal_orb_view(eaof_orb* ex, const eaof_kp** kps, const uint8_t** desc, const int** counts, int* cap,
                                      int* w, int* h, const float** scale, int* nlevels, void** stream);
extern "C" int eaof_internal_orb_note_reader(eaof_orb* ex, void* readerStream);

namespace {

int vfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return eaof_internal_fail(code, buf);
}
#define VCK(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return vfail(EAOF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int ASM_THREADS = 1024;
constexpr uint32_t DROPPED = 0xffffffffu;

__device__ __forceinline__ int hamming256(const uint32_t* a, const uint4 b0, const uint4 b1) {
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) + __popc(a[4] ^ b1.x) +
           __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

// Tree in "position" order: position p holds node childIdx[p], so the children of any node are a contiguous run.
struct Tree {
    const uint4* desc;       // [nPos][2]
    const int2* kids;        // [nPos] children run (begin, end) of the node at this position; begin == end on a leaf
    const uint32_t* node;    // [nPos] original node id
    const uint32_t* word;    // [nPos] word id (leaves)
    const double* weight;    // [nPos]
    int rootBegin, rootEnd, L;
};

// base == nullptr: set s starts at row s*stride; count[s] features.
template <int G>
__global__ void __launch_bounds__(256) k_voc_descend(Tree T, const uint8_t* __restrict__ feats, const int* __restrict__ base,
                                                     const int* __restrict__ count, int stride, int levelsup,
                                                     uint32_t* __restrict__ fWord, uint32_t* __restrict__ fNode,
                                                     double* __restrict__ fWeight) {
    const int s = blockIdx.y;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / G, sub = threadIdx.x & (G - 1);
    if (i >= count[s]) return;  // whole groups leave together
    const unsigned grp = G == 32 ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));
    const size_t row = (size_t)(base ? base[s] : s * stride) + i;
    uint32_t q[8];
    {
        const uint4* p = reinterpret_cast<const uint4*>(feats + 32 * row);
        const uint4 a = p[0], b = p[1];
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    const int nidLevel = T.L - levelsup;
    uint32_t nid = 0;  // root when nid_level <= 0 (:1243)
    int b = T.rootBegin, e = T.rootEnd, level = 0, pos = -1;
    while (e > b) {
        ++level;
        unsigned best = 0xffffffffu;  // (distance << 8) | child ordinal: the first child with the smallest distance wins (:1256-1264)
        for (int j = sub; j < e - b; j += G) {
            const uint4 d0 = __ldg(T.desc + 2 * (size_t)(b + j)), d1 = __ldg(T.desc + 2 * (size_t)(b + j) + 1);
            const unsigned key = ((unsigned)hamming256(q, d0, d1) << 8) | (unsigned)j;
            best = key < best ? key : best;
        }
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            const unsigned other = __shfl_xor_sync(grp, best, o);
            best = other < best ? other : best;
        }
        pos = b + (int)(best & 0xffu);
        if (level == nidLevel) nid = T.node[pos];
        const int2 k = T.kids[pos];
        b = k.x; e = k.y;
    }
    if (sub == 0) {
        const size_t o = (size_t)s * stride + i;
        const double w = pos >= 0 ? T.weight[pos] : 0.0;
        fWeight[o] = w;
        fWord[o] = (pos >= 0 && w > 0) ? T.word[pos] : DROPPED;  // "stopped" words are dropped (:1168)
        fNode[o] = nid;
    }
}

// exclusive scan of one int per thread over the CTA (ASM_THREADS threads)
__device__ __forceinline__ int cta_excl_scan(int v, int* warpSums, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warpSums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int s = warpSums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        warpSums[lane] = s;
    }
    __syncthreads();
    const int r = (wid ? warpSums[wid - 1] : 0) + incl - v;
    *total = warpSums[31];
    __syncthreads();
    return r;
}

// Sorts the n keys of `keys` (DROPPED-high keys sink to the end) into `sorted` by rank counting; keys are distinct.
__device__ __forceinline__ void rank_sort(const unsigned long long* keys, unsigned long long* sorted, int n) {
    for (int i = threadIdx.x; i < n; i += ASM_THREADS) {
        const unsigned long long k = keys[i];
        int r = 0;
        for (int j = 0; j < n; ++j) r += keys[j] < k;
        sorted[r] = k;
    }
}

// ids[slot] / starts[slot] for every run of equal high words in sorted[0..m); returns the number of runs.
__device__ __forceinline__ int run_heads(const unsigned long long* sorted, int m, int per, int* warpSums, uint32_t* ids,
                                         int* starts) {
    const int r0 = threadIdx.x * per;
    int heads = 0;
    for (int r = r0; r < min(r0 + per, m); ++r) heads += r == 0 || (sorted[r] >> 32) != (sorted[r - 1] >> 32);
    int total;
    int slot = cta_excl_scan(heads, warpSums, &total);
    for (int r = r0; r < min(r0 + per, m); ++r)
        if (r == 0 || (sorted[r] >> 32) != (sorted[r - 1] >> 32)) {
            ids[slot] = (uint32_t)(sorted[r] >> 32);
            starts[slot] = r;
            ++slot;
        }
    return total;
}

__global__ void __launch_bounds__(ASM_THREADS) k_voc_assemble(const int* __restrict__ count, int stride, int weighting,
                                                              int scoring, const uint32_t* __restrict__ fWord,
                                                              const uint32_t* __restrict__ fNode,
                                                              const double* __restrict__ fWeight, int* __restrict__ runStart,
                                                              int* __restrict__ nWords, uint32_t* __restrict__ wordIds,
                                                              double* __restrict__ wordVals, int* __restrict__ nFNodes,
                                                              uint32_t* __restrict__ nodeIds, int* __restrict__ nodeStart,
                                                              uint32_t* __restrict__ featIdx) {
    extern __shared__ unsigned long long smemV[];
    __shared__ int warpSums[32];
    __shared__ int sKept;
    __shared__ double sNorm;
    const int s = blockIdx.x, n = count[s], tid = threadIdx.x;
    const size_t o = (size_t)s * stride;
    unsigned long long* keys = smemV;
    unsigned long long* sorted = smemV + stride;
    const int per = (n + ASM_THREADS - 1) / ASM_THREADS;
    int* rs = runStart + o;
    if (n == 0) {
        if (tid == 0) { nWords[s] = 0; nFNodes[s] = 0; nodeStart[(size_t)s * (stride + 1)] = 0; }
        return;
    }
    // ---- BowVector: words in ascending id, value = weight added once per feature in feature order
    if (tid == 0) sKept = 0;
    __syncthreads();
    int kept = 0;
    for (int i = tid; i < n; i += ASM_THREADS) {
        const uint32_t w = fWord[o + i];
        keys[i] = ((unsigned long long)w << 32) | (unsigned)i;
        kept += w != DROPPED;
    }
    atomicAdd(&sKept, kept);
    __syncthreads();
    const int m = sKept;
    rank_sort(keys, sorted, n);
    __syncthreads();
    const int nw = run_heads(sorted, m, per, warpSums, wordIds + o, rs);
    __syncthreads();
    double* vals = reinterpret_cast<double*>(keys);  // keys are no longer needed
    for (int j = tid; j < nw; j += ASM_THREADS) {
        const int b = rs[j], e = j + 1 < nw ? rs[j + 1] : m;
        const double w = fWeight[o + (uint32_t)(sorted[b] & 0xffffffffu)];  // every feature of the word carries the same weight
        double v = w;
        if (weighting == EAOF_VOC_TF || weighting == EAOF_VOC_TF_IDF)
            for (int c = b + 1; c < e; ++c) v += w;  // BowVector::addWeight, one rounding per feature
        vals[j] = v;
    }
    __syncthreads();
    const bool must = scoring != EAOF_VOC_DOT_PRODUCT;  // ScoringObject.h:74-89
    if (tid == 0) {
        double norm = 0.0;
        if (must) {  // BowVector::normalize, ascending word order
            if (scoring == EAOF_VOC_L2_NORM) { for (int j = 0; j < nw; ++j) norm += vals[j] * vals[j]; norm = sqrt(norm); }
            else for (int j = 0; j < nw; ++j) norm += fabs(vals[j]);
        } else if ((weighting == EAOF_VOC_TF || weighting == EAOF_VOC_TF_IDF) && nw > 0) {
            norm = (double)nw;  // :1176-1181
        }
        sNorm = norm;
        nWords[s] = nw;
    }
    __syncthreads();
    {
        const double norm = sNorm;
        for (int j = tid; j < nw; j += ASM_THREADS) wordVals[o + j] = norm > 0.0 ? vals[j] / norm : vals[j];
    }
    __syncthreads();
    // ---- FeatureVector: nodes in ascending id, features of a node in ascending index
    for (int i = tid; i < n; i += ASM_THREADS) {
        const bool drop = fWord[o + i] == DROPPED;
        keys[i] = ((unsigned long long)(drop ? DROPPED : fNode[o + i]) << 32) | (unsigned)i;
    }
    __syncthreads();
    rank_sort(keys, sorted, n);
    __syncthreads();
    int* ns = nodeStart + (size_t)s * (stride + 1);
    const int nn = run_heads(sorted, m, per, warpSums, nodeIds + o, ns);
    for (int r = tid; r < m; r += ASM_THREADS) featIdx[o + r] = (uint32_t)(sorted[r] & 0xffffffffu);
    if (tid == 0) { ns[nn] = m; nFNodes[s] = nn; }
}

}  // namespace

struct eaof_voc {
    int device = 0, L = 0, nPos = 0, maxFeat = 0, maxSets = 0, weighting = 0, scoring = 0, maxKids = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evDep = nullptr;
    // tree
    uint4* dDesc = nullptr; int2* dKids = nullptr; uint32_t* dNode = nullptr; uint32_t* dWord = nullptr; double* dWeight = nullptr;
    int rootBegin = 0, rootEnd = 0;
    // per-feature scratch [maxSets][maxFeat]
    uint32_t *fWord = nullptr, *fNode = nullptr;
    double* fWeight = nullptr;
    int* runStart = nullptr;
    // staging for the host-buffer call
    uint8_t* dFeat = nullptr;
    int *dBase = nullptr, *dCount = nullptr;
    int* oNW = nullptr; uint32_t* oWI = nullptr; double* oWV = nullptr; int* oNN = nullptr; uint32_t* oNI = nullptr;
    int* oNS = nullptr; uint32_t* oFI = nullptr;
    Tree tree() const { return Tree{dDesc, dKids, dNode, dWord, dWeight, rootBegin, rootEnd, L}; }
    size_t asm_smem(int stride) const { return 16 * (size_t)stride; }
};

namespace {
int launch(eaof_voc* v, const uint8_t* feats, const int* base, const int* count, int nSets, int stride, int levelsup, int* nW,
           uint32_t* wI, double* wV, int* nN, uint32_t* nI, int* nS, uint32_t* fI) {
    cudaStream_t s = v->stream;
    if (v->maxKids <= 16)
        k_voc_descend<16><<<dim3((stride * 16 + 255) / 256, nSets), 256, 0, s>>>(v->tree(), feats, base, count, stride, levelsup,
                                                                                v->fWord, v->fNode, v->fWeight);
    else
        k_voc_descend<32><<<dim3((stride * 32 + 255) / 256, nSets), 256, 0, s>>>(v->tree(), feats, base, count, stride, levelsup,
                                                                                v->fWord, v->fNode, v->fWeight);
    k_voc_assemble<<<nSets, ASM_THREADS, v->asm_smem(stride), s>>>(count, stride, v->weighting, v->scoring, v->fWord, v->fNode,
                                                                  v->fWeight, v->runStart, nW, wI, wV, nN, nI, nS, fI);
    VCK(cudaGetLastError());
    return EAOF_OK;
}
// workspace arrays start out zeroed: tails that no kernel writes are never stale memory when an output block is downloaded
template <typename T> cudaError_t dalloc(T** p, size_t n) {
    const size_t bytes = sizeof(T) * (n ? n : 1);
    cudaError_t e = cudaMalloc(p, bytes);
    return e != cudaSuccess ? e : cudaMemset(*p, 0, bytes);
}
}  // namespace

extern "C" {

int eaof_voc_create(int device, int L, int nNodes, const int* childStart, const int* childIdx, const uint8_t* nodeDesc,
                    const double* weight, const int* wordId, int weighting, int scoring, int maxFeat, int maxSets,
                    eaof_voc** out) {
    if (!out) return vfail(EAOF_ERR_ARG, "null argument");
    *out = nullptr;
    if (L < 1 || nNodes < 2 || !childStart || !childIdx || !nodeDesc || !weight || !wordId)
        return vfail(EAOF_ERR_ARG, "empty vocabulary or null array");
    if (weighting < EAOF_VOC_TF_IDF || weighting > EAOF_VOC_BINARY || scoring < EAOF_VOC_L1_NORM || scoring > EAOF_VOC_DOT_PRODUCT)
        return vfail(EAOF_ERR_ARG, "unknown weighting / scoring type");
    if (maxFeat < 1 || maxFeat > 12000 || maxSets < 1) return vfail(EAOF_ERR_ARG, "max_features must be in [1,12000], max_sets >= 1");
    const int nPos = childStart[nNodes];
    if (childStart[0] != 0 || nPos != nNodes - 1) return vfail(EAOF_ERR_ARG, "child lists must hold every node but the root exactly once");
    int maxKids = 0;
    std::vector<char> seen(nNodes, 0);
    for (int i = 0; i < nNodes; ++i) {
        if (childStart[i + 1] < childStart[i]) return vfail(EAOF_ERR_ARG, "child_start must be non-decreasing");
        maxKids = std::max(maxKids, childStart[i + 1] - childStart[i]);
    }
    for (int p = 0; p < nPos; ++p) {
        const int c = childIdx[p];
        if (c <= 0 || c >= nNodes || seen[c]) return vfail(EAOF_ERR_ARG, "child_idx[%d] = %d is not a valid, unique node id", p, c);
        seen[c] = 1;
    }
    if (maxKids < 1 || maxKids > 255) return vfail(EAOF_ERR_ARG, "branching factor %d not supported", maxKids);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return vfail(EAOF_ERR_CUDA, "no CUDA device: libeaof_orb has no CPU fallback");
    if (device < 0 || device >= ndev) return vfail(EAOF_ERR_ARG, "device out of range");
    VCK(cudaSetDevice(device));
    eaof_voc* v = new eaof_voc;
    v->device = device; v->L = L; v->nPos = nPos; v->maxFeat = maxFeat; v->maxSets = maxSets; v->weighting = weighting;
    v->scoring = scoring; v->maxKids = maxKids; v->rootBegin = childStart[0]; v->rootEnd = childStart[1];
    std::vector<uint8_t> hDesc(32 * (size_t)nPos);
    std::vector<int2> hKids(nPos);
    std::vector<uint32_t> hNode(nPos), hWord(nPos);
    std::vector<double> hW(nPos);
    for (int p = 0; p < nPos; ++p) {
        const int c = childIdx[p];
        memcpy(&hDesc[32 * (size_t)p], nodeDesc + 32 * (size_t)c, 32);
        hKids[p] = make_int2(childStart[c], childStart[c + 1]);
        hNode[p] = (uint32_t)c;
        hWord[p] = (uint32_t)wordId[c];
        hW[p] = weight[c];
    }
    const size_t PF = (size_t)maxSets * maxFeat;
    cudaError_t e = cudaSuccess;
#define A_(x) if (e == cudaSuccess) e = (x)
    A_(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
    A_(cudaEventCreateWithFlags(&v->evDep, cudaEventDisableTiming));
    A_(dalloc(&v->dDesc, 2 * (size_t)nPos)); A_(dalloc(&v->dKids, nPos)); A_(dalloc(&v->dNode, nPos)); A_(dalloc(&v->dWord, nPos));
    A_(dalloc(&v->dWeight, nPos));
    A_(dalloc(&v->fWord, PF)); A_(dalloc(&v->fNode, PF)); A_(dalloc(&v->fWeight, PF)); A_(dalloc(&v->runStart, PF));
    A_(dalloc(&v->dFeat, 32 * PF)); A_(dalloc(&v->dBase, maxSets)); A_(dalloc(&v->dCount, maxSets));
    A_(dalloc(&v->oNW, maxSets)); A_(dalloc(&v->oWI, PF)); A_(dalloc(&v->oWV, PF)); A_(dalloc(&v->oNN, maxSets));
    A_(dalloc(&v->oNI, PF)); A_(dalloc(&v->oNS, PF + maxSets)); A_(dalloc(&v->oFI, PF));
    A_(cudaMemcpy(v->dDesc, hDesc.data(), hDesc.size(), cudaMemcpyHostToDevice));
    A_(cudaMemcpy(v->dKids, hKids.data(), sizeof(int2) * nPos, cudaMemcpyHostToDevice));
    A_(cudaMemcpy(v->dNode, hNode.data(), sizeof(uint32_t) * nPos, cudaMemcpyHostToDevice));
    A_(cudaMemcpy(v->dWord, hWord.data(), sizeof(uint32_t) * nPos, cudaMemcpyHostToDevice));
    A_(cudaMemcpy(v->dWeight, hW.data(), sizeof(double) * nPos, cudaMemcpyHostToDevice));
    A_(cudaFuncSetAttribute(k_voc_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(16 * 12000)));
#undef A_
    if (e != cudaSuccess) {
        vfail(EAOF_ERR_CUDA, "vocabulary allocation failed: %s", cudaGetErrorString(e));
        eaof_voc_destroy(v);
        return EAOF_ERR_CUDA;
    }
    *out = v;
    return EAOF_OK;
}

void eaof_voc_destroy(eaof_voc* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->stream) cudaStreamSynchronize(v->stream);
    cudaFree(v->dDesc); cudaFree(v->dKids); cudaFree(v->dNode); cudaFree(v->dWord); cudaFree(v->dWeight);
    cudaFree(v->fWord); cudaFree(v->fNode); cudaFree(v->fWeight); cudaFree(v->runStart); cudaFree(v->dFeat);
    cudaFree(v->dBase); cudaFree(v->dCount); cudaFree(v->oNW); cudaFree(v->oWI); cudaFree(v->oWV); cudaFree(v->oNN);
    cudaFree(v->oNI); cudaFree(v->oNS); cudaFree(v->oFI);
    if (v->evDep) cudaEventDestroy(v->evDep);
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
}

void* eaof_voc_stream(eaof_voc* v) { return v ? (void*)v->stream : nullptr; }
int eaof_voc_sync(eaof_voc* v) {
    if (!v) return vfail(EAOF_ERR_ARG, "null handle");
    VCK(cudaStreamSynchronize(v->stream));
    return EAOF_OK;
}

int eaof_voc_transform(eaof_voc* v, int nSets, const int* setStart, const uint8_t* desc, int levelsup, int* nWords,
                       uint32_t* wordIds, double* wordVals, int* nFNodes, uint32_t* nodeIds, int* nodeStart, uint32_t* featIdx) {
    if (!v || nSets < 0 || !setStart || !nWords || !wordIds || !wordVals || !nFNodes || !nodeIds || !nodeStart || !featIdx)
        return vfail(EAOF_ERR_ARG, "bad argument");
    if (nSets == 0) return EAOF_OK;
    if (nSets > v->maxSets) return vfail(EAOF_ERR_ARG, "n_sets %d exceeds max_sets=%d", nSets, v->maxSets);
    std::vector<int> cnt(nSets);
    for (int s = 0; s < nSets; ++s) {
        cnt[s] = setStart[s + 1] - setStart[s];
        if (cnt[s] < 0 || cnt[s] > v->maxFeat) return vfail(EAOF_ERR_ARG, "set %d has %d features (max_features=%d)", s, cnt[s], v->maxFeat);
    }
    const int total = setStart[nSets] - setStart[0];
    if (total > 0 && !desc) return vfail(EAOF_ERR_ARG, "null descriptors");
    VCK(cudaSetDevice(v->device));
    cudaStream_t st = v->stream;
    std::vector<int> rel(nSets);
    for (int s = 0; s < nSets; ++s) rel[s] = setStart[s] - setStart[0];
    VCK(cudaMemcpyAsync(v->dBase, rel.data(), sizeof(int) * nSets, cudaMemcpyHostToDevice, st));
    VCK(cudaMemcpyAsync(v->dCount, cnt.data(), sizeof(int) * nSets, cudaMemcpyHostToDevice, st));
    if (total) VCK(cudaMemcpyAsync(v->dFeat, desc + 32 * (size_t)setStart[0], 32 * (size_t)total, cudaMemcpyHostToDevice, st));
    const int S = v->maxFeat;
    int rc = launch(v, v->dFeat, v->dBase, v->dCount, nSets, S, levelsup, v->oNW, v->oWI, v->oWV, v->oNN, v->oNI, v->oNS, v->oFI);
    if (rc) return rc;
    VCK(cudaMemcpyAsync(nWords, v->oNW, sizeof(int) * nSets, cudaMemcpyDeviceToHost, st));
    VCK(cudaMemcpyAsync(nFNodes, v->oNN, sizeof(int) * nSets, cudaMemcpyDeviceToHost, st));
    VCK(cudaStreamSynchronize(st));
    for (int s = 0; s < nSets; ++s) {
        const size_t o = (size_t)setStart[s];
        const size_t d = (size_t)s * S;
        if (nWords[s]) {
            VCK(cudaMemcpyAsync(wordIds + o, v->oWI + d, sizeof(uint32_t) * nWords[s], cudaMemcpyDeviceToHost, st));
            VCK(cudaMemcpyAsync(wordVals + o, v->oWV + d, sizeof(double) * nWords[s], cudaMemcpyDeviceToHost, st));
        }
        if (nFNodes[s]) VCK(cudaMemcpyAsync(nodeIds + o, v->oNI + d, sizeof(uint32_t) * nFNodes[s], cudaMemcpyDeviceToHost, st));
        VCK(cudaMemcpyAsync(nodeStart + o + s, v->oNS + (size_t)s * (S + 1), sizeof(int) * (nFNodes[s] + 1), cudaMemcpyDeviceToHost, st));
        if (cnt[s]) VCK(cudaMemcpyAsync(featIdx + o, v->oFI + d, sizeof(uint32_t) * cnt[s], cudaMemcpyDeviceToHost, st));
    }
    VCK(cudaStreamSynchronize(st));
    return EAOF_OK;
}

int eaof_voc_transform_orb_device(eaof_voc* v, eaof_orb* ex, int nFrames, int levelsup, int* dNW, uint32_t* dWI, double* dWV,
                                  int* dNN, uint32_t* dNI, int* dNS, uint32_t* dFI) {
    if (!v || !ex || nFrames < 1 || !dNW || !dWI || !dWV || !dNN || !dNI || !dNS || !dFI) return vfail(EAOF_ERR_ARG, "bad argument");
    const eaof_kp* kps; const uint8_t* desc; const int* counts; const float* scale; int cap, w, h, nlev; void* exStream;
    int rc = eaof_internal_orb_view(ex, &kps, &desc, &counts, &cap, &w, &h, &scale, &nlev, &exStream);
    if (rc) return rc;
    if (nFrames > v->maxSets || cap > v->maxFeat)
        return vfail(EAOF_ERR_ARG, "vocabulary handle sized for %d sets of %d features, asked for %d x %d", v->maxSets, v->maxFeat, nFrames, cap);
    VCK(cudaSetDevice(v->device));
    VCK(cudaEventRecord(v->evDep, (cudaStream_t)exStream));
    VCK(cudaStreamWaitEvent(v->stream, v->evDep, 0));
    rc = launch(v, desc, nullptr, counts, nFrames, cap, levelsup, dNW, dWI, dWV, dNN, dNI, dNS, dFI);
    if (rc) return rc;
    return eaof_internal_orb_note_reader(ex, v->stream);
}

}  // extern "C"
