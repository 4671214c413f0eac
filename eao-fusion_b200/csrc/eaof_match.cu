// Hamming matcher of libeaof_orb.so (include/eaof_match.h): the descriptor search loops of the reference's
// ORBmatcher (src/ORBmatcher.cc) as two-phase GPU algorithms.
//
//   phase 1 (dense, parallel over every query of every pair): all Hamming distances of a query against its
//            candidate set on the INT pipe (8 x XOR + POPC per pair, ORBmatcher::DescriptorDistance :1649-1665),
//            keeping only what the acceptance rule can ever look at: the candidates with dist < D ("near list", in
//            candidate order) for the ratio-test loops, the 4 best candidates for the best-only projection loop.
//   phase 2 (one warp per pair, queries in the reference's order): resolves the greedy "skip targets that are
//            already matched" exclusion (:209-210, :576, :1405-1407), applies TH_LOW/TH_HIGH and the fp32 ratio
//            test, fills the rotation histogram, runs ComputeThreeMaxima (:1603-1644) and prunes.
//
// Why the near list is exact: a match needs best <= TH and (float)best < ratio*(float)second.  With
// s_min = min{s : (float)TH < ratio*(float)s} every second-best >= s_min passes the ratio test for every admissible
// best, so only candidates with dist < D = max(s_min, TH+1) can influence the outcome; if a query has more such
// candidates than the list holds, phase 2 re-scans that query exactly.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/eaof_match.h"
#include "bow_umma.cuh"
#include <cuda.h>

extern "C" int eaof_internal_fail(int code, const char* msg);  // sets eaof_last_error (eaof_orb.cu)
// device-resident results of an extractor handle, and the hook that makes its next batch wait for a reader (eaof_orb.cu)
extern "C" int eaof_internal_orb_view(eaof_orb* ex, const eaof_kp** kps, const uint8_t** desc, const int** counts, int* cap,
                                      int* w, int* h, const float** scale, int* nlevels, void** stream);
extern "C" int eaof_internal_orb_note_reader(eaof_orb* ex, void* readerStream);

namespace {

int mfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return eaof_internal_fail(code, buf);
}
#define MCK(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return mfail(EAOF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int GRID_COLS = 64, GRID_ROWS = 48, GRID_CELLS = GRID_COLS * GRID_ROWS;  // include/Frame.h:89-90
constexpr int NEAR_K = 7;                                                          // near-list entries per query
constexpr int TOP_K = 6;                                                           // projection: best candidates kept

__device__ __forceinline__ int hamming256(const uint32_t* a, const uint32_t* b) {
    int d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d += __popc(a[i] ^ b[i]);
    return d;
}

// The same distance with 5 POPC instead of 8: POPC issues at a quarter of the ALU rate (it is the pipe the brute-force
// kernels saturate, profiles/r01_bowdense_full_summary.txt), so three carry-save adders (sum = a^b^c, carry = maj(a,b,c),
// two LOP3 each) fold seven of the eight XOR words into one weight-1 word and three weight-2 words first.
__device__ __forceinline__ int hamming256_csa(const uint32_t* a, const uint32_t* b) {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a[i] ^ b[i];
    const uint32_t s1 = x[0] ^ x[1] ^ x[2], c1 = (x[0] & x[1]) | (x[2] & (x[0] ^ x[1]));
    const uint32_t s2 = x[3] ^ x[4] ^ x[5], c2 = (x[3] & x[4]) | (x[5] & (x[3] ^ x[4]));
    const uint32_t s3 = s1 ^ s2 ^ x[6], c3 = (s1 & s2) | (x[6] & (s1 ^ s2));
    return __popc(s3) + __popc(x[7]) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

__global__ void k_hamming_pairs(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* pa = reinterpret_cast<const uint4*>(a + 32 * (size_t)i);
    const uint4* pb = reinterpret_cast<const uint4*>(b + 32 * (size_t)i);
    const uint4 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1];
    out[i] = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
             __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---------------------------------------------------------------------------------------------------------
// SearchByBoW
struct BowSeg { int qOff, qCnt, tOff, tCnt; };  // one vocabulary node present in both feature vectors

struct BowArgs {
    int spec;  // k_bow_resolve: 32-queries-at-a-time resolution (needs stride * 4 bytes of shared memory for the proposal owners)
    // per pair p: descriptors of the query/target blocks, optional CSR index lists (NULL = identity)
    const uint8_t* desc;      // base
    const float* angle;       // base, same indexing as desc rows
    const int* counts;        // per block (brute force) or NULL
    const int* pairQ;         // block index per pair (device) or NULL (=> host API: block 0 = Q, block 1 = T)
    const int* pairT;
    int blockStride;          // rows per block
    const uint8_t* validQ;    // [pair][stride] or NULL
    const uint8_t* validT;
    const int* idxQ;          // [pair][stride] feature index lists (NULL = identity)
    const int* idxT;
    const BowSeg* segs;       // segments of all pairs (NULL = one segment covering everything)
    const int* segStart;      // [pair+1] (NULL with segs)
    // device-built plans (eaof_match_bow_orb_device): segments laid out [pair][stride] with segCount[pair] entries, and
    // node-sorted feature lists per block (FeatureVector of every frame) from which idxQ / idxT derive through pairQ / pairT
    const int* segCount;
    const uint32_t* featIdx;
    int featStride;
    int nQhost, nThost;       // used when counts == NULL
    int stride;               // per-pair stride of the workspace arrays (= max_features)
    int mode;
    int thEff;                // accept best <= thEff
    int D;                    // near-list threshold
    float ratio;
    int checkOri;
};

#define BOW_QT 128  // queries per CTA
#define BOW_TT 64   // targets staged per step

// phase 1: blockIdx.y = pair, blockIdx.x = tile of BOW_QT list positions inside segment `seg` (brute force: seg 0)
__global__ void __launch_bounds__(BOW_QT) k_bow_dense(BowArgs A, const int2* __restrict__ tiles, uint32_t* __restrict__ nearBuf) {
    __shared__ __align__(16) uint32_t sT[BOW_TT][8];
    __shared__ int sIdx[BOW_TT];
    int pair, qBeg, qEnd, tBeg, tEnd;
    if (tiles) {  // host API: explicit (segment, first list position) tiles of pair 0
        const int2 t = tiles[blockIdx.x];
        pair = 0;
        const BowSeg s = A.segs[t.x];
        qBeg = t.y;
        qEnd = min(s.qOff + s.qCnt, t.y + BOW_QT);
        tBeg = s.tOff;
        tEnd = s.tOff + s.tCnt;
    } else {
        pair = blockIdx.y;
        const int nq = A.counts[A.pairQ[pair]], nt = A.counts[A.pairT[pair]];
        qBeg = blockIdx.x * BOW_QT;
        if (qBeg >= nq) return;
        qEnd = min(nq, qBeg + BOW_QT);
        tBeg = 0;
        tEnd = nt;
    }
    const int bq = A.pairQ ? A.pairQ[pair] : 0, bt = A.pairT ? A.pairT[pair] : 1;
    const uint8_t* dQ = A.desc + (size_t)bq * A.blockStride * 32;
    const uint8_t* dT = A.desc + (size_t)bt * A.blockStride * 32;
    const int* idxQ = A.idxQ ? A.idxQ + (size_t)pair * A.stride : nullptr;
    const int* idxT = A.idxT ? A.idxT + (size_t)pair * A.stride : nullptr;

    const int qpos = qBeg + threadIdx.x;
    const bool active = qpos < qEnd;
    const int q = active ? (idxQ ? idxQ[qpos] : qpos) : 0;
    uint32_t qd[8];
    {
        const uint4* p = reinterpret_cast<const uint4*>(dQ + 32 * (size_t)q);
        const uint4 a = p[0], b = p[1];
        qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
    }
    uint32_t near[NEAR_K];
#pragma unroll
    for (int i = 0; i < NEAR_K; ++i) near[i] = 0xffffffffu;
    int cnt = 0;
    for (int t0 = tBeg; t0 < tEnd; t0 += BOW_TT) {
        const int tt = min(BOW_TT, tEnd - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tt * 2; i += BOW_QT) {
            const int j = i >> 1, h = i & 1;
            const int t = idxT ? idxT[t0 + j] : t0 + j;
            reinterpret_cast<uint4*>(&sT[j][0])[h] = reinterpret_cast<const uint4*>(dT + 32 * (size_t)t)[h];
            if (h == 0) sIdx[j] = t;
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int j = 0; j < tt; ++j) {
                const int d = hamming256_csa(qd, sT[j]);
                if (d < A.D) {
                    if (cnt < NEAR_K) {
                        const uint32_t e = (uint32_t)sIdx[j] | ((uint32_t)d << 16);
#pragma unroll
                        for (int i = 0; i < NEAR_K; ++i) if (i == cnt) near[i] = e;
                    }
                    ++cnt;
                }
            }
        }
    }
    if (active) {
        uint32_t* o = nearBuf + ((size_t)pair * A.stride + q) * 8;
        reinterpret_cast<uint4*>(o)[0] = make_uint4(near[0], near[1], near[2], near[3]);
        reinterpret_cast<uint4*>(o)[1] = make_uint4(near[4], near[5], near[6], (uint32_t)cnt);
    }
}

// Device-built plan of a SearchByBoW pair (src/ORBmatcher.cc:182-264: the merge-walk of the two FeatureVectors): warp =
// pair; lanes take the query frame's nodes 32 at a time, look each one up in the target frame's ascending node list and
// the hits are appended in node order.  FeatureVectors in the layout eaof_voc_transform_orb_device leaves.
__global__ void __launch_bounds__(32) k_bow_plan(BowArgs A, const int* __restrict__ nFNodes, const uint32_t* __restrict__ nodeIds,
                                                 const int* __restrict__ nodeStart, BowSeg* __restrict__ segs,
                                                 int* __restrict__ segCount) {
    const int pair = blockIdx.x, lane = threadIdx.x;
    const int bq = A.pairQ[pair], bt = A.pairT[pair];
    const int nNQ = nFNodes[bq], nNT = nFNodes[bt];
    const uint32_t* idQ = nodeIds + (size_t)bq * A.featStride;
    const uint32_t* idT = nodeIds + (size_t)bt * A.featStride;
    const int* stQ = nodeStart + (size_t)bq * (A.featStride + 1);
    const int* stT = nodeStart + (size_t)bt * (A.featStride + 1);
    BowSeg* out = segs + (size_t)pair * A.stride;
    const unsigned below = (1u << lane) - 1;
    int n = 0;
    for (int a0 = 0; a0 < nNQ; a0 += 32) {
        const int a = a0 + lane;
        BowSeg s{0, 0, 0, 0};
        bool hit = false;
        if (a < nNQ) {
            const uint32_t id = idQ[a];
            int lo = 0, hi = nNT;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (idT[mid] < id) lo = mid + 1; else hi = mid; }
            if (lo < nNT && idT[lo] == id) {
                s = BowSeg{stQ[a], stQ[a + 1] - stQ[a], stT[lo], stT[lo + 1] - stT[lo]};
                hit = s.qCnt > 0 && s.tCnt > 0;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) out[n + __popc(m & below)] = s;
        n += __popc(m);
    }
    if (lane == 0) segCount[pair] = n;
}

// phase 1 for planned pairs: thread = one position of the query frame's node-sorted feature list; the thread finds the
// segment its position falls in and scans that node's target features straight from global memory (a vocabulary node
// holds about ten features of a frame: nothing to stage).  Same near-list output as k_bow_dense.
__global__ void __launch_bounds__(BOW_QT) k_bow_dense_nodes(BowArgs A, uint32_t* __restrict__ nearBuf) {
    const int pair = blockIdx.y, p = blockIdx.x * BOW_QT + threadIdx.x;
    const int nSeg = A.segCount[pair];
    const BowSeg* segs = A.segs + (size_t)pair * A.stride;
    int lo = 0, hi = nSeg;  // last segment with qOff <= p
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (segs[mid].qOff <= p) lo = mid + 1; else hi = mid; }
    if (lo == 0) return;
    const BowSeg s = segs[lo - 1];
    if (p >= s.qOff + s.qCnt) return;
    const int bq = A.pairQ[pair], bt = A.pairT[pair];
    const uint8_t* dQ = A.desc + (size_t)bq * A.blockStride * 32;
    const uint8_t* dT = A.desc + (size_t)bt * A.blockStride * 32;
    const int* idxQ = reinterpret_cast<const int*>(A.featIdx + (size_t)bq * A.featStride);
    const int* idxT = reinterpret_cast<const int*>(A.featIdx + (size_t)bt * A.featStride);
    const int q = idxQ[p];
    uint32_t qd[8];
    {
        const uint4* pq = reinterpret_cast<const uint4*>(dQ + 32 * (size_t)q);
        const uint4 a = pq[0], b = pq[1];
        qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
    }
    uint32_t near[NEAR_K];
#pragma unroll
    for (int i = 0; i < NEAR_K; ++i) near[i] = 0xffffffffu;
    int cnt = 0;
    for (int tp = s.tOff; tp < s.tOff + s.tCnt; ++tp) {
        const int t = idxT[tp];
        uint32_t td[8];
        const uint4* pt = reinterpret_cast<const uint4*>(dT + 32 * (size_t)t);
        const uint4 a = __ldg(pt), b = __ldg(pt + 1);
        td[0] = a.x; td[1] = a.y; td[2] = a.z; td[3] = a.w; td[4] = b.x; td[5] = b.y; td[6] = b.z; td[7] = b.w;
        const int d = hamming256_csa(qd, td);
        if (d < A.D) {
            if (cnt < NEAR_K) {
                const uint32_t e = (uint32_t)t | ((uint32_t)d << 16);
#pragma unroll
                for (int i = 0; i < NEAR_K; ++i) if (i == cnt) near[i] = e;
            }
            ++cnt;
        }
    }
    uint32_t* o = nearBuf + ((size_t)pair * A.stride + q) * 8;
    reinterpret_cast<uint4*>(o)[0] = make_uint4(near[0], near[1], near[2], near[3]);
    reinterpret_cast<uint4*>(o)[1] = make_uint4(near[4], near[5], near[6], (uint32_t)cnt);
}

__global__ void __launch_bounds__(256) k_kp_angles(const eaof_kp* __restrict__ kps, int n, float* __restrict__ angle) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) angle[i] = kps[i].angle;
}

// ComputeThreeMaxima, src/ORBmatcher.cc:1603-1644
__device__ void three_maxima(const int* hist, int& i1, int& i2, int& i3) {
    int max1 = 0, max2 = 0, max3 = 0;
    i1 = i2 = i3 = -1;
    for (int i = 0; i < EAOF_HISTO_LENGTH; ++i) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
        else if (s > max3) { max3 = s; i3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3 = -1; }
}

__device__ __forceinline__ int rot_bin(float a1, float a2, float factor) {
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == EAOF_HISTO_LENGTH) bin = 0;
    return bin;
}

// phase 2: one warp per pair
__global__ void __launch_bounds__(32) k_bow_resolve(BowArgs A, const uint32_t* __restrict__ nearBuf,
                                                    uint32_t* __restrict__ accBuf, int* __restrict__ matchOut,
                                                    int* __restrict__ distOut, int* __restrict__ nMatches) {
    extern __shared__ uint32_t smem[];  // matched-target bitmap [ (stride+31)/32 ]
    __shared__ int hist[EAOF_HISTO_LENGTH];
    const int pair = blockIdx.x, lane = threadIdx.x;
    const int bq = A.pairQ ? A.pairQ[pair] : 0, bt = A.pairT ? A.pairT[pair] : 1;
    const int nq = A.counts ? A.counts[bq] : A.nQhost, nt = A.counts ? A.counts[bt] : A.nThost;
    const uint8_t* dQ = A.desc + (size_t)bq * A.blockStride * 32;
    const uint8_t* dT = A.desc + (size_t)bt * A.blockStride * 32;
    const float* angQ = A.angle + (size_t)bq * A.blockStride;
    const float* angT = A.angle + (size_t)bt * A.blockStride;
    const int* idxQ = A.featIdx ? reinterpret_cast<const int*>(A.featIdx + (size_t)bq * A.featStride)
                                : (A.idxQ ? A.idxQ + (size_t)pair * A.stride : nullptr);
    const int* idxT = A.featIdx ? reinterpret_cast<const int*>(A.featIdx + (size_t)bt * A.featStride)
                                : (A.idxT ? A.idxT + (size_t)pair * A.stride : nullptr);
    const uint8_t* validQ = A.validQ ? A.validQ + (size_t)pair * A.stride : nullptr;
    const uint8_t* validT = (A.validT && A.mode == EAOF_BOW_KF_KF) ? A.validT + (size_t)pair * A.stride : nullptr;
    const int nOut = A.mode == EAOF_BOW_KF_FRAME ? nt : nq;
    int* mOut = matchOut + (size_t)pair * A.stride;
    int* dOut = distOut ? distOut + (size_t)pair * A.stride : nullptr;
    uint32_t* acc = accBuf + (size_t)pair * A.stride;
    const uint32_t* nearP = nearBuf + (size_t)pair * A.stride * 8;

    const int words = (A.stride + 31) >> 5;
    uint32_t* owner = smem + words;  // [stride] lowest lane of the current batch that proposes this target (0xffffffff: none)
    for (int i = lane; i < words; i += 32) smem[i] = 0;
    if (A.spec)
        for (int i = lane; i < A.stride; i += 32) owner[i] = 0xffffffffu;
    for (int i = lane; i < EAOF_HISTO_LENGTH; i += 32) hist[i] = 0;
    for (int i = lane; i < nOut; i += 32) { mOut[i] = -1; if (dOut) dOut[i] = -1; }
    __syncwarp();
    if (validT)  // targets without a good map point are never candidates (:573-580)
        for (int i = lane; i < nt; i += 32) if (!validT[i]) atomicOr(&smem[i >> 5], 1u << (i & 31));
    __syncwarp();

    const float factor = 1.0f / EAOF_HISTO_LENGTH;  // :172, :541
    int nAcc = 0;
    const int nSeg = A.segCount ? A.segCount[pair] : (A.segs ? (A.segStart[pair + 1] - A.segStart[pair]) : 1);
    const size_t segBase = A.segCount ? (size_t)pair * A.stride : (A.segs ? (size_t)A.segStart[pair] : 0);
    for (int si = 0; si < nSeg; ++si) {
        BowSeg s;
        if (A.segs) s = A.segs[segBase + si];
        else s = BowSeg{0, nq, 0, nt};
        for (int q0 = s.qOff; q0 < s.qOff + s.qCnt; q0 += 32) {
            const int myPos = q0 + lane;
            const bool have = myPos < s.qOff + s.qCnt;
            const int myQ = have ? (idxQ ? idxQ[myPos] : myPos) : 0;
            uint4 w0 = make_uint4(0, 0, 0, 0), w1 = make_uint4(0, 0, 0, 0);
            bool ok = have && (!validQ || validQ[myQ]);
            if (ok) {
                const uint4* p = reinterpret_cast<const uint4*>(nearP + (size_t)myQ * 8);
                w0 = p[0];
                w1 = p[1];
            }
            const int myCnt = ok ? (int)w1.w : 0;
            const unsigned todo = __ballot_sync(0xffffffffu, myCnt > 0);
            if (todo == 0) continue;
            if (A.spec && !__any_sync(0xffffffffu, myCnt > NEAR_K)) {
                // 32 queries at once, exact: every lane resolves its own query against the matched-target bitmap plus the
                // targets proposed by EARLIER lanes of this batch (owner[t] < lane), proposals are republished and the
                // evaluation repeated until nothing changes.  Lane 0 is never blocked by a proposal, lane l only by lanes
                // < l: by induction the fixed point is the result of walking the 32 queries in order, reached after as many
                // rounds as the longest chain of stolen targets (one or two in practice).  Batches with a query that needs
                // the exact re-scan (> NEAR_K near candidates) take the sequential loop below.
                const uint32_t e[NEAR_K] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z};
                int accT = -1, accD = 0;
                for (;;) {
                    int best1 = 256, best2 = 256, bestIdx = -1;
#pragma unroll
                    for (int k = 0; k < NEAR_K; ++k) {
                        if (k < myCnt) {
                            const int t = e[k] & 0xffff, d = (int)(e[k] >> 16);
                            if (!((smem[t >> 5] >> (t & 31)) & 1u) && owner[t] >= (uint32_t)lane) {
                                if (d < best1) { best2 = best1; best1 = d; bestIdx = t; }
                                else if (d < best2) best2 = d;
                            }
                        }
                    }
                    const int newT = (bestIdx >= 0 && best1 <= A.thEff && (float)best1 < __fmul_rn(A.ratio, (float)best2)) ? bestIdx : -1;
                    const bool changed = newT != accT;
                    if (!__any_sync(0xffffffffu, changed)) break;
                    __syncwarp();
                    if (accT >= 0) owner[accT] = 0xffffffffu;  // retract (a proposal of the same target by another lane is re-published below)
                    __syncwarp();
                    accT = newT;
                    accD = best1;
                    if (accT >= 0) atomicMin(&owner[accT], (uint32_t)lane);
                    __syncwarp();
                }
                // commit in query order
                __syncwarp();  // every lane has finished reading the bitmap and the owners
                const unsigned am = __ballot_sync(0xffffffffu, accT >= 0);
                if (accT >= 0) {
                    const int outIdx = A.mode == EAOF_BOW_KF_FRAME ? accT : myQ;
                    atomicOr(&smem[accT >> 5], 1u << (accT & 31));
                    owner[accT] = 0xffffffffu;
                    mOut[outIdx] = A.mode == EAOF_BOW_KF_FRAME ? myQ : accT;
                    if (dOut) dOut[outIdx] = accD;
                    acc[nAcc + __popc(am & ((1u << lane) - 1u))] = (uint32_t)myQ | ((uint32_t)accT << 16);
                }
                nAcc += __popc(am);
                __syncwarp();
                continue;
            }
            unsigned rem = todo;
            while (rem) {  // queries with at least one near candidate, in list order
                const int j = __ffs(rem) - 1;
                rem &= rem - 1;
                const int q = __shfl_sync(0xffffffffu, myQ, j);
                const int cnt = __shfl_sync(0xffffffffu, myCnt, j);
                int best1 = 256, best2 = 256, bestIdx = -1;
                if (cnt <= NEAR_K) {
                    const uint32_t e[NEAR_K] = {__shfl_sync(0xffffffffu, w0.x, j), __shfl_sync(0xffffffffu, w0.y, j),
                                                __shfl_sync(0xffffffffu, w0.z, j), __shfl_sync(0xffffffffu, w0.w, j),
                                                __shfl_sync(0xffffffffu, w1.x, j), __shfl_sync(0xffffffffu, w1.y, j),
                                                __shfl_sync(0xffffffffu, w1.z, j)};
#pragma unroll
                    for (int k = 0; k < NEAR_K; ++k) {
                        if (k < cnt) {
                            const int t = e[k] & 0xffff, d = (int)(e[k] >> 16);
                            if (!((smem[t >> 5] >> (t & 31)) & 1u)) {
                                if (d < best1) { best2 = best1; best1 = d; bestIdx = t; }
                                else if (d < best2) best2 = d;
                            }
                        }
                    }
                    // an unseen second-best is >= D, which passes the ratio test for every best <= thEff
                } else {
                    // exact re-scan of this query over the whole node (rare: > NEAR_K near candidates)
                    uint32_t qd[8];
                    {
                        const uint4* p = reinterpret_cast<const uint4*>(dQ + 32 * (size_t)q);
                        const uint4 a = p[0], b = p[1];
                        qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
                    }
                    int a1 = 256, a2 = 256, aPos = 0x7fffffff, aIdx = -1;  // this lane's two smallest, first position of the min
                    for (int tp = s.tOff + lane; tp < s.tOff + s.tCnt; tp += 32) {
                        const int t = idxT ? idxT[tp] : tp;
                        if ((smem[t >> 5] >> (t & 31)) & 1u) continue;
                        uint32_t td[8];
                        const uint4* p = reinterpret_cast<const uint4*>(dT + 32 * (size_t)t);
                        const uint4 a = p[0], b = p[1];
                        td[0] = a.x; td[1] = a.y; td[2] = a.z; td[3] = a.w; td[4] = b.x; td[5] = b.y; td[6] = b.z; td[7] = b.w;
                        const int d = hamming256(qd, td);
                        if (d < a1) { a2 = a1; a1 = d; aPos = tp; aIdx = t; }
                        else if (d < a2) a2 = d;
                    }
                    // warp merge: global min by (dist, position), second = 2nd order statistic of the multiset
                    int g1 = a1, gPos = aPos, gIdx = aIdx, gLane = lane;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const int o1 = __shfl_xor_sync(0xffffffffu, g1, o), oP = __shfl_xor_sync(0xffffffffu, gPos, o);
                        const int oI = __shfl_xor_sync(0xffffffffu, gIdx, o), oL = __shfl_xor_sync(0xffffffffu, gLane, o);
                        if (o1 < g1 || (o1 == g1 && oP < gPos)) { g1 = o1; gPos = oP; gIdx = oI; gLane = oL; }
                    }
                    int sec = (lane == gLane) ? a2 : a1;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sec = min(sec, __shfl_xor_sync(0xffffffffu, sec, o));
                    best1 = g1; bestIdx = gIdx; best2 = sec;
                }
                if (bestIdx >= 0 && best1 <= A.thEff && (float)best1 < __fmul_rn(A.ratio, (float)best2)) {
                    if (lane == 0) {
                        const int outIdx = A.mode == EAOF_BOW_KF_FRAME ? bestIdx : q;
                        smem[bestIdx >> 5] |= 1u << (bestIdx & 31);
                        mOut[outIdx] = A.mode == EAOF_BOW_KF_FRAME ? q : bestIdx;
                        if (dOut) dOut[outIdx] = best1;
                        acc[nAcc] = (uint32_t)q | ((uint32_t)bestIdx << 16);
                    }
                    ++nAcc;
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();
    // rotation histogram over the accepted matches (bins do not influence acceptance, so they are computed here, in
    // parallel, instead of inside the sequential loop), ComputeThreeMaxima, pruning (:266-285)
    int removed = 0;
    if (A.checkOri) {
        for (int k = lane; k < nAcc; k += 32) {
            const uint32_t a = acc[k];
            atomicAdd(&hist[rot_bin(angQ[a & 0xffff], angT[a >> 16], factor)], 1);
        }
        __syncwarp();
        int i1, i2, i3;
        three_maxima(hist, i1, i2, i3);
        for (int k = lane; k < nAcc; k += 32) {
            const uint32_t a = acc[k];
            const int q = a & 0xffff, t = a >> 16;
            const int bin = rot_bin(angQ[q], angT[t], factor);
            if (bin != i1 && bin != i2 && bin != i3) {
                const int outIdx = A.mode == EAOF_BOW_KF_FRAME ? t : q;
                mOut[outIdx] = -1;
                if (dOut) dOut[outIdx] = -1;
                ++removed;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    }
    if (lane == 0) nMatches[pair] = nAcc - removed;
}

// ---------------------------------------------------------------------------------------------------------
// SearchForTriangulation  src/ORBmatcher.cc:657-823 (+ CheckDistEpipolarLine :140-157).  There is no greedy exclusion
// in this loop (vbMatched2 is never set inside it), so every query is independent: one thread per KF1 feature of a
// shared vocabulary node scans the node's KF2 features staged in shared memory.  bestDist starts at TH_LOW and a
// candidate replaces the best when dist <= bestDist, i.e. on equal distance the LATER candidate wins.
struct TriArgs {
    const uint8_t* desc;  // block 0 = KF1 rows, block 1 = KF2 rows (blockStride rows apart)
    int blockStride;
    const float* x1; const float* y1; const float* angle1; const uint8_t* free1; const uint8_t* stereo1;
    const float* x2; const float* y2; const int* oct2; const float* angle2; const uint8_t* free2; const uint8_t* stereo2;
    const int* idx1; const int* idx2;
    const BowSeg* segs;
    float F[9];
    float ex, ey;
    float scale2[EAOF_MAX_LEVELS], sigma2[EAOF_MAX_LEVELS];
    int onlyStereo, checkOri, n1;
};

struct TriCand { float x, y; int idx; int flags; };  // flags: bit0 stereo, bits 8.. octave

__global__ void __launch_bounds__(BOW_QT) k_tri_dense(TriArgs A, const int2* __restrict__ tiles, int* __restrict__ match12,
                                                      int* __restrict__ dist12) {
    __shared__ __align__(16) uint32_t sT[BOW_TT][8];
    __shared__ TriCand sC[BOW_TT];
    const int2 tile = tiles[blockIdx.x];
    const BowSeg s = A.segs[tile.x];
    const int qpos = tile.y + threadIdx.x;
    const int qEnd = min(s.qOff + s.qCnt, tile.y + BOW_QT);
    const int q = qpos < qEnd ? A.idx1[qpos] : -1;
    const uint8_t* d1 = A.desc;
    const uint8_t* d2 = A.desc + (size_t)A.blockStride * 32;
    bool active = q >= 0 && A.free1[q];                       // :697-701 a feature that already has a map point is skipped
    const bool bStereo1 = active && A.stereo1 && A.stereo1[q];
    if (A.onlyStereo && !bStereo1) active = false;            // :705-707
    uint32_t qd[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float la = 0.f, lb = 0.f, lc = 0.f, den = 0.f;
    if (active) {
        const uint4* p = reinterpret_cast<const uint4*>(d1 + 32 * (size_t)q);
        const uint4 a = p[0], b = p[1];
        qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
        const float x = A.x1[q], y = A.y1[q];                 // epipolar line l = x1' F12, :143-145
        la = __fadd_rn(__fadd_rn(__fmul_rn(x, A.F[0]), __fmul_rn(y, A.F[3])), A.F[6]);
        lb = __fadd_rn(__fadd_rn(__fmul_rn(x, A.F[1]), __fmul_rn(y, A.F[4])), A.F[7]);
        lc = __fadd_rn(__fadd_rn(__fmul_rn(x, A.F[2]), __fmul_rn(y, A.F[5])), A.F[8]);
        den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
    }
    int bestDist = EAOF_TH_LOW, bestIdx = -1;
    for (int t0 = s.tOff; t0 < s.tOff + s.tCnt; t0 += BOW_TT) {
        const int tt = min(BOW_TT, s.tOff + s.tCnt - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tt * 2; i += BOW_QT) {
            const int j = i >> 1, h = i & 1;
            const int t = A.idx2[t0 + j];
            reinterpret_cast<uint4*>(&sT[j][0])[h] = reinterpret_cast<const uint4*>(d2 + 32 * (size_t)t)[h];
            if (h == 0) {
                const bool st = A.stereo2 && A.stereo2[t];
                // a KF2 feature that has a map point is never a candidate (:725); idx -1 marks it
                const bool usable = A.free2[t] && !(A.onlyStereo && !st);
                sC[j] = TriCand{A.x2[t], A.y2[t], usable ? t : -1, (st ? 1 : 0) | (A.oct2[t] << 8)};
            }
        }
        __syncthreads();
        if (active) {
            for (int j = 0; j < tt; ++j) {
                const TriCand c = sC[j];
                if (c.idx < 0) continue;
                const int d = hamming256_csa(qd, sT[j]);
                if (d > EAOF_TH_LOW || d > bestDist) continue;   // :738
                const int oct = c.flags >> 8;
                if (!bStereo1 && !(c.flags & 1)) {               // :743-749 too close to the epipole
                    const float dx = __fsub_rn(A.ex, c.x), dy = __fsub_rn(A.ey, c.y);
                    if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, A.scale2[oct])) continue;
                }
                const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, c.x), __fmul_rn(lb, c.y)), lc);   // :147-156
                if (den == 0.f) continue;
                const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                if ((double)dsqr < __dmul_rn(3.84, (double)A.sigma2[oct])) { bestIdx = c.idx; bestDist = d; }
            }
        }
    }
    if (q >= 0 && bestIdx >= 0) {  // every feature sits in at most one node, so q is written by one thread only
        match12[q] = bestIdx;
        dist12[q] = bestDist;
    }
}

// rotation histogram (factor 1/HISTO_LENGTH, :684), ComputeThreeMaxima, pruning and the match count (:786-808)
__global__ void __launch_bounds__(256) k_tri_finish(TriArgs A, int* __restrict__ match12, int* __restrict__ dist12,
                                                    int* __restrict__ nMatches) {
    __shared__ int hist[EAOF_HISTO_LENGTH];
    __shared__ int total, keep[3];
    const float factor = 1.0f / EAOF_HISTO_LENGTH;
    if (threadIdx.x < EAOF_HISTO_LENGTH) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < A.n1; i += blockDim.x) {
        const int t = match12[i];
        if (t < 0) continue;
        ++mine;
        if (A.checkOri) atomicAdd(&hist[rot_bin(A.angle1[i], A.angle2[t], factor)], 1);
    }
    atomicAdd(&total, mine);
    __syncthreads();
    if (A.checkOri) {
        if (threadIdx.x == 0) three_maxima(hist, keep[0], keep[1], keep[2]);
        __syncthreads();
        int removed = 0;
        for (int i = threadIdx.x; i < A.n1; i += blockDim.x) {
            const int t = match12[i];
            if (t < 0) continue;
            const int bin = rot_bin(A.angle1[i], A.angle2[t], factor);
            if (bin != keep[0] && bin != keep[1] && bin != keep[2]) { match12[i] = -1; dist12[i] = -1; ++removed; }
        }
        atomicSub(&total, removed);
        __syncthreads();
    }
    if (threadIdx.x == 0) *nMatches = total;
}

// ---------------------------------------------------------------------------------------------------------
// SearchByProjection(Cur, Last)
__device__ __forceinline__ void level_window(int searchMode, int oct, int& minLevel, int& maxLevel);
struct ProjArgs {
    // Cur side, [pair][stride]
    const float* cx; const float* cy; const int* coct; const float* cangle; const float* curight; const uint8_t* ctaken;
    const int* nC;           // [pair]
    // Last side, [pair][stride]
    const float* lu; const float* lv; const float* linvz; const int* loct; const float* langle;
    const uint8_t* lvalid; const uint8_t* lobs;
    const int* nL;           // [pair]
    const uint8_t* desc;     // descriptor base; rows of pair p: Cur at cRow[p], Last at lRow[p]
    const int* cRow; const int* lRow;
    int stride;
    float minX, maxX, minY, maxY, invW, invH;
    float scale[EAOF_MAX_LEVELS];
    float th, mbf;
    int searchMode, checkOri;
    // generalised window queries (eaof_match_windows): per-query radius / level window / predicted right coordinate
    // replace th*scale[octave], the searchMode window and u - mbf/z; NULL = SearchByProjection(Cur,Last) behaviour
    const float* qRadius; const int* qMinL; const int* qMaxL; const float* qUr;
    int thAccept;      // accept best <= thAccept (TH_HIGH, or ORBdist of SearchByProjection(Cur,KF))
    int cut;           // phase 1 keeps candidates with dist <= cut
    int histMode;      // 0: no rotation check, 1: factor 1/HISTO_LENGTH, 2: factor HISTO_LENGTH/360
    int checkBounds;   // skip queries projected outside [minX,maxX]x[minY,maxY]
    float ratio;       // rule EAOF_WIN_RATIO_SAME_LEVEL
    float invSigma2[EAOF_MAX_LEVELS];  // gate EAOF_GATE_FUSE_CHI2: mvInvLevelSigma2 of the target keyframe
};

// window of query i of pair `pair`; false when the query is skipped before the search
__device__ __forceinline__ bool query_window(const ProjArgs& A, size_t po, int i, float& u, float& v, float& r, int& minLevel,
                                             int& maxLevel, float& ur) {
    const float invz = A.linvz ? A.linvz[po + i] : 1.f;
    u = A.lu[po + i];
    v = A.lv[po + i];
    if (A.lvalid && !A.lvalid[po + i]) return false;
    if (invz < 0) return false;
    if (A.checkBounds && ((u < A.minX || u > A.maxX) || (v < A.minY || v > A.maxY))) return false;
    if (A.qRadius) {
        r = A.qRadius[po + i];
        minLevel = A.qMinL[po + i];
        maxLevel = A.qMaxL[po + i];
    } else {
        const int oct = A.loct[po + i];
        r = __fmul_rn(A.th, A.scale[oct]);
        level_window(A.searchMode, oct, minLevel, maxLevel);
    }
    ur = A.qUr ? A.qUr[po + i] : __fsub_rn(u, __fmul_rn(A.mbf, invz));
    return true;
}

// Frame::AssignFeaturesToGrid / PosInGrid (src/Frame.cc:599-614,751-761): one CTA per pair builds the CSR grid of
// the Cur frame; cell lists are sorted ascending so they equal the reference's push_back order.
__global__ void __launch_bounds__(256) k_build_grid(ProjArgs A, int* __restrict__ cellStart, int* __restrict__ cellIdx,
                                                    float4* __restrict__ cellPack) {
    __shared__ int cnt[GRID_CELLS];
    __shared__ int warpTot[8];
    const int pair = blockIdx.x, tid = threadIdx.x;
    const int n = A.nC[pair];
    const float* x = A.cx + (size_t)pair * A.stride;
    const float* y = A.cy + (size_t)pair * A.stride;
    int* cs = cellStart + (size_t)pair * (GRID_CELLS + 1);
    int* ci = cellIdx + (size_t)pair * A.stride;
    for (int i = tid; i < GRID_CELLS; i += 256) cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(x[i], A.minX), A.invW));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(y[i], A.minY), A.invH));
        if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) atomicAdd(&cnt[px * GRID_ROWS + py], 1);
    }
    __syncthreads();
    // exclusive scan of 3072 counts: 12 per thread
    int local[12], sum = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) { local[k] = cnt[tid * 12 + k]; sum += local[k]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
    if ((tid & 31) == 31) warpTot[tid >> 5] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < (tid >> 5); ++w) base += warpTot[w];
    int run = base + incl - sum;
#pragma unroll
    for (int k = 0; k < 12; ++k) { cs[tid * 12 + k] = run; cnt[tid * 12 + k] = run; run += local[k]; }
    if (tid == 255) cs[GRID_CELLS] = run;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(x[i], A.minX), A.invW));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(y[i], A.minY), A.invH));
        if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) ci[atomicAdd(&cnt[px * GRID_ROWS + py], 1)] = i;
    }
    __syncthreads();
    for (int c = tid; c < GRID_CELLS; c += 256) {  // ascending feature index inside each cell
        const int b = cs[c], e = cnt[c];
        for (int i = b + 1; i < e; ++i) {
            const int v = ci[i];
            int j = i - 1;
            while (j >= b && ci[j] > v) { ci[j + 1] = ci[j]; --j; }
            ci[j + 1] = v;
        }
    }
    __syncthreads();
    // the same lists with the fields phase 1 filters on next to each other: (x, y, octave, feature index) per entry, so
    // that a query streams over one contiguous run per grid column instead of chasing cellIdx -> x/y/octave
    float4* pk = cellPack + (size_t)pair * A.stride;
    const int* oct = A.coct + (size_t)pair * A.stride;
    const int total = cs[GRID_CELLS];
    for (int j = tid; j < total; j += 256) {
        const int k = ci[j];
        pk[j] = make_float4(x[k], y[k], __int_as_float(oct[k]), __int_as_float(k));
    }
}

// Walks Frame::GetFeaturesInArea (src/Frame.cc:696-749) for one query and calls f(k) for every candidate, in the
// reference's order (ix, iy, insertion order).  Returns false if the query is skipped before the search.
template <typename F>
__device__ __forceinline__ void for_each_candidate(const ProjArgs& A, const int* cs, const int* ci, const float* cx,
                                                    const float* cy, const int* coct, float u, float v, float r,
                                                    int minLevel, int maxLevel, F f) {
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, A.minX), r), A.invW)));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, A.minX), r), A.invW)));
    if (nMaxCellX < 0) return;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, A.minY), r), A.invH)));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, A.minY), r), A.invH)));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
        for (int iy = nMinCellY; iy <= nMaxCellY; ++iy) {
            const int c = ix * GRID_ROWS + iy;
            for (int j = cs[c]; j < cs[c + 1]; ++j) {
                const int k = ci[j];
                if (bCheckLevels) {
                    if (coct[k] < minLevel) continue;
                    if (maxLevel >= 0 && coct[k] > maxLevel) continue;
                }
                const float dx = __fsub_rn(cx[k], u), dy = __fsub_rn(cy[k], v);
                if (fabsf(dx) < r && fabsf(dy) < r) f(k);
            }
        }
}

__device__ __forceinline__ void level_window(int searchMode, int oct, int& minLevel, int& maxLevel) {
    if (searchMode == 1) { minLevel = oct; maxLevel = -1; }        // bForward  :1387
    else if (searchMode == 2) { minLevel = 0; maxLevel = oct; }    // bBackward :1389
    else { minLevel = oct - 1; maxLevel = oct + 1; }               // :1391
}

// phase 1: PROJ_LANES lanes per query.  The candidate sequence of Frame::GetFeaturesInArea (cells in (ix, iy) order,
// features of a cell in insertion order) is dealt round-robin to the lanes by its running ordinal, so that the scattered
// position / descriptor loads of different candidates are in flight together; every lane keeps its TOP_K best
// (distance, then ordinal), and the lanes' sorted lists are merged by PROJ_LANES-way selection.  Output per query: the
// TOP_K best candidates by (dist asc, arrival asc) + the number of candidates with dist <= cut.
// GATE selects the per-candidate test between the window test and the distance: 0 = the stereo gate of the Frame
// searches (src/ORBmatcher.cc:1409-1415, :93-98), 1 = the reprojection chi-square of Fuse(KeyFrame*, vpMapPoints, th)
// (:901-929).
#define PROJ_LANES 4
template <int GATE>
__global__ void __launch_bounds__(128) k_proj_dense(ProjArgs A, const int* __restrict__ cellStart,
                                                    const float4* __restrict__ cellPack, uint32_t* __restrict__ topBuf) {
    const int pair = blockIdx.y;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / PROJ_LANES, sub = threadIdx.x & (PROJ_LANES - 1);
    const unsigned grp = 0xfu << (threadIdx.x & 28);  // the lanes of this query inside the warp
    if (i >= A.nL[pair]) return;                      // whole groups leave together
    const size_t po = (size_t)pair * A.stride;
    int count = 0;
    uint32_t top[TOP_K], ordv[TOP_K];
#pragma unroll
    for (int k = 0; k < TOP_K; ++k) { top[k] = 0xffffffffu; ordv[k] = 0xffffffffu; }
    float u, v, r, ur;
    int minLevel, maxLevel;
    const bool ok = query_window(A, po, i, u, v, r, minLevel, maxLevel, ur);
    if (ok) {
        uint32_t qd[8];
        {
            const uint4* p = reinterpret_cast<const uint4*>(A.desc + 32 * ((size_t)A.lRow[pair] + i));
            const uint4 a = p[0], b = p[1];
            qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
        }
        const uint8_t* cd = A.desc + 32 * (size_t)A.cRow[pair];
        const float* cur = A.curight ? A.curight + po : nullptr;
        const int* cs = cellStart + (size_t)pair * (GRID_CELLS + 1);
        const float4* pk = cellPack + po;
        // Frame::GetFeaturesInArea, src/Frame.cc:696-749.  Cells (ix, iy0..iy1) are adjacent in the CSR order, so one grid
        // column is one contiguous run of entries, in the reference's visiting order.
        const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, A.minX), r), A.invW)));
        const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, A.minX), r), A.invW)));
        const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, A.minY), r), A.invH)));
        const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, A.minY), r), A.invH)));
        const bool any = !(nMinCellX >= GRID_COLS || nMaxCellX < 0 || nMinCellY >= GRID_ROWS || nMaxCellY < 0);
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        int ord = 0;
        if (any)
            for (int ix = nMinCellX; ix <= nMaxCellX; ++ix) {
                {
                    const int b = cs[ix * GRID_ROWS + nMinCellY], e = cs[ix * GRID_ROWS + nMaxCellY + 1];
                    for (int j = b + ((sub - ord) & (PROJ_LANES - 1)); j < e; j += PROJ_LANES) {
                        const float4 ent = __ldg(pk + j);
                        const int k = __float_as_int(ent.w);
                        if (bCheckLevels) {
                            const int o = __float_as_int(ent.z);
                            if (o < minLevel) continue;
                            if (maxLevel >= 0 && o > maxLevel) continue;
                        }
                        const float dx = __fsub_rn(ent.x, u), dy = __fsub_rn(ent.y, v);
                        if (!(fabsf(dx) < r && fabsf(dy) < r)) continue;
                        if (GATE == 0) {
                            if (cur && cur[k] > 0 && fabsf(__fsub_rn(ur, cur[k])) > r) continue;  // :1409-1415
                        } else {
                            const float kr = cur ? cur[k] : -1.f, inv = A.invSigma2[__float_as_int(ent.z)];
                            const float ex = __fsub_rn(u, ent.x), ey = __fsub_rn(v, ent.y);
                            float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                            if (kr >= 0) {  // :901-914
                                const float er = __fsub_rn(ur, kr);
                                e2 = __fadd_rn(e2, __fmul_rn(er, er));
                                if ((double)__fmul_rn(e2, inv) > 7.8) continue;
                            } else if ((double)__fmul_rn(e2, inv) > 5.99) continue;  // :915-925
                        }
                        uint32_t td[8];
                        const uint4* p = reinterpret_cast<const uint4*>(cd + 32 * (size_t)k);
                        const uint4 a = p[0], b2 = p[1];
                        td[0] = a.x; td[1] = a.y; td[2] = a.z; td[3] = a.w; td[4] = b2.x; td[5] = b2.y; td[6] = b2.z; td[7] = b2.w;
                        const uint32_t d = (uint32_t)hamming256(qd, td);
                        if (d > (uint32_t)A.cut) continue;  // can never influence the outcome: not kept, counted or re-scanned
                        uint32_t e2 = (d << 16) | (uint32_t)k, o2 = (uint32_t)(ord + (j - b));
                        // sorted insertion by distance; this lane's candidates arrive in ordinal order, ties keep the earlier
                        bool shifting = false;
#pragma unroll
                        for (int s2 = 0; s2 < TOP_K; ++s2) {
                            if (shifting || (e2 >> 16) < (top[s2] >> 16)) {
                                const uint32_t t = top[s2], to = ordv[s2];
                                top[s2] = e2; ordv[s2] = o2; e2 = t; o2 = to; shifting = true;
                            }
                        }
                        ++count;
                    }
                    ord += e - b;
                }
            }
    }
    // merge: TOP_K rounds, each takes the smallest (dist, ordinal) head among the lanes of the group
    uint32_t out[TOP_K];
    int head = 0;
#pragma unroll
    for (int round = 0; round < TOP_K; ++round) {
        uint32_t hv = 0xffffffffu, ho = 0xffffffffu;
#pragma unroll
        for (int s2 = 0; s2 < TOP_K; ++s2)
            if (s2 == head) { hv = top[s2]; ho = ordv[s2]; }
        unsigned long long key = hv == 0xffffffffu ? ~0ull : (((unsigned long long)(hv >> 16)) << 40) | ((unsigned long long)ho << 8) | sub;
        unsigned long long best = key;
#pragma unroll
        for (int o = 1; o < PROJ_LANES; o <<= 1) {
            const unsigned long long other = __shfl_xor_sync(grp, best, o);
            best = other < best ? other : best;
        }
        const int winner = (int)(best & 0xff);
        const uint32_t wv = __shfl_sync(grp, hv, (threadIdx.x & 28) | (winner & (PROJ_LANES - 1)));
        out[round] = best == ~0ull ? 0xffffffffu : wv;
        if (best != ~0ull && winner == sub) ++head;
    }
#pragma unroll
    for (int o = 1; o < PROJ_LANES; o <<= 1) count += __shfl_xor_sync(grp, count, o);
    if (sub == 0) {
        uint32_t* o = topBuf + (po + i) * 8;
        reinterpret_cast<uint4*>(o)[0] = make_uint4(out[0], out[1], out[2], out[3]);
        reinterpret_cast<uint4*>(o)[1] = make_uint4(out[4], out[5], 0, ok ? (uint32_t)count : 0xffffffffu);
    }
}

// Independent window queries (Fuse x2, SearchBySim3: no "already matched" exclusion inside the loop, so the result of a
// query is the head of its phase-1 list): thread = query.
__global__ void __launch_bounds__(128) k_win_pick(int nQ, int thAccept, const uint32_t* __restrict__ topBuf,
                                                  int* __restrict__ matchQ, int* __restrict__ distQ, int* __restrict__ nMatches) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nQ) return;
    const uint32_t head = topBuf[(size_t)i * 8], cnt = topBuf[(size_t)i * 8 + 7];
    const bool hit = cnt != 0xffffffffu && head != 0xffffffffu && (int)(head >> 16) <= thAccept;
    matchQ[i] = hit ? (int)(head & 0xffffu) : -1;
    distQ[i] = hit ? (int)(head >> 16) : -1;
    const unsigned b = __ballot_sync(__activemask(), hit);
    if (hit && (threadIdx.x & 31) == __ffs(b) - 1) atomicAdd(nMatches, __popc(b));
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-307), batched: CTA = map point, warp = one row of its
// N x N distance matrix at a time.  The row median (element int(0.5*(N-1)) of the sorted row) is read off a 257-bin
// histogram of the row's distances instead of sorting; rows compete by (median, row index), so the first row with the
// smallest median wins as in :293-297.
__global__ void __launch_bounds__(128) k_distinctive(const int* __restrict__ start, const uint8_t* __restrict__ desc,
                                                     int* __restrict__ bestOut, int* __restrict__ medOut) {
    __shared__ int hist[4][288];
    __shared__ unsigned long long warpBest[4];
    const int p = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = start[p], n = start[p + 1] - s;
    if (n <= 0) {
        if (threadIdx.x == 0) { bestOut[p] = -1; medOut[p] = -1; }
        return;
    }
    const int k = (n - 1) >> 1;
    const uint8_t* D = desc + 32 * (size_t)s;
    unsigned long long best = ~0ull;
    for (int row = w; row < n; row += 4) {
#pragma unroll
        for (int b = 0; b < 9; ++b) hist[w][lane * 9 + b] = 0;
        __syncwarp();
        uint32_t qd[8];
        {
            const uint4* q = reinterpret_cast<const uint4*>(D + 32 * (size_t)row);
            const uint4 a = q[0], b = q[1];
            qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
        }
        for (int j = lane; j < n; j += 32) {
            uint32_t td[8];
            const uint4* t = reinterpret_cast<const uint4*>(D + 32 * (size_t)j);
            const uint4 a = t[0], b = t[1];
            td[0] = a.x; td[1] = a.y; td[2] = a.z; td[3] = a.w; td[4] = b.x; td[5] = b.y; td[6] = b.z; td[7] = b.w;
            atomicAdd(&hist[w][hamming256_csa(qd, td)], 1);
        }
        __syncwarp();
        int bins[9], sum = 0;
#pragma unroll
        for (int b = 0; b < 9; ++b) { bins[b] = hist[w][lane * 9 + b]; sum += bins[b]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int owner = __ffs(__ballot_sync(0xffffffffu, incl > k)) - 1;  // some lane always qualifies: total = n > k
        int med = 0;
        if (lane == owner) {
            int c = incl - sum;
            bool found = false;
#pragma unroll
            for (int b = 0; b < 9; ++b) {
                c += bins[b];
                if (!found && c > k) { med = lane * 9 + b; found = true; }
            }
        }
        med = __shfl_sync(0xffffffffu, med, owner);
        const unsigned long long key = ((unsigned long long)med << 32) | (unsigned)row;
        best = key < best ? key : best;
        __syncwarp();
    }
    if (lane == 0) warpBest[w] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long b = warpBest[0];
        for (int i = 1; i < 4; ++i) b = warpBest[i] < b ? warpBest[i] : b;
        bestOut[p] = (int)(b & 0xffffffffu);
        medOut[p] = (int)(b >> 32);
    }
}

// phase 2: one warp per pair, Last features in index order
__global__ void __launch_bounds__(32) k_proj_resolve(ProjArgs A, const int* __restrict__ cellStart,
                                                     const float4* __restrict__ cellPack, const uint32_t* __restrict__ topBuf,
                                                     uint32_t* __restrict__ accBuf, int* __restrict__ matchOut,
                                                     int* __restrict__ distOut, int* __restrict__ nMatches) {
    extern __shared__ uint32_t smem[];  // taken bitmap
    __shared__ int hist[EAOF_HISTO_LENGTH];
    const int pair = blockIdx.x, lane = threadIdx.x;
    const size_t po = (size_t)pair * A.stride;
    const int nC = A.nC[pair], nL = A.nL[pair];
    int* mOut = matchOut + po;
    int* dOut = distOut ? distOut + po : nullptr;
    uint32_t* acc = accBuf + po;
    const int words = (A.stride + 31) >> 5;
    for (int i = lane; i < words; i += 32) smem[i] = 0;
    for (int i = lane; i < EAOF_HISTO_LENGTH; i += 32) hist[i] = 0;
    for (int i = lane; i < nC; i += 32) { mOut[i] = -1; if (dOut) dOut[i] = -1; }
    __syncwarp();
    if (A.ctaken)
        for (int i = lane; i < nC; i += 32) if (A.ctaken[po + i]) atomicOr(&smem[i >> 5], 1u << (i & 31));
    __syncwarp();
    const float factor = A.histMode == 2 ? EAOF_HISTO_LENGTH / 360.0f : 1.0f / EAOF_HISTO_LENGTH;  // :1337 vs :1486
    // The reference walks the Last features in index order and each accepted match may make its Cur feature
    // unavailable to every later query (:1405-1407).  32 queries are resolved at a time: every lane proposes the
    // first candidate of its list that is still free, all lanes below the first lane whose proposal collides with a
    // lower lane's exclusive claim commit (their choice cannot depend on anything still undecided), and the rest
    // propose again against the updated bitmap.  The bitmap only ever gains bits, so a candidate seen taken stays
    // ruled out; a fixed point of this is exactly the sequential result.
    const unsigned below = (1u << lane) - 1;
    int nAcc = 0;
    // the lists of the next 32 queries are fetched while the current 32 are being resolved
    uint4 nw0 = make_uint4(~0u, ~0u, ~0u, ~0u), nw1 = make_uint4(~0u, ~0u, 0u, 0u);
    int nObs = 1;
    auto fetch = [&](int q) {
        nw0 = make_uint4(~0u, ~0u, ~0u, ~0u); nw1 = make_uint4(~0u, ~0u, 0u, 0u); nObs = 1;
        if (q < nL) {
            const uint4* p = reinterpret_cast<const uint4*>(topBuf + (po + q) * 8);
            nw0 = p[0]; nw1 = p[1];
            if (A.lobs) nObs = A.lobs[po + q] != 0;
        }
    };
    fetch(lane);
    for (int i0 = 0; i0 < nL; i0 += 32) {
        const int mine = i0 + lane;
        const uint32_t e[TOP_K] = {nw0.x, nw0.y, nw0.z, nw0.w, nw1.x, nw1.y};
        const int myCnt = mine < nL ? (int)nw1.w : 0, myObs = nObs;
        fetch(mine + 32);
        bool pending = myCnt > 0;
        int next = 0;  // entries before `next` are known to be taken
        while (__ballot_sync(0xffffffffu, pending)) {
            int prop = -1, propDist = 256;
            bool needScan = false;
            if (pending) {
#pragma unroll
                for (int k = 0; k < TOP_K; ++k) {
                    if (prop < 0 && k >= next && k < myCnt) {
                        const int t = e[k] & 0xffff;
                        if (!((smem[t >> 5] >> (t & 31)) & 1u)) { prop = t; propDist = (int)(e[k] >> 16); }
                        next = k + (prop < 0);
                    }
                }
                if (prop < 0) {
                    if (myCnt > TOP_K) needScan = true;  // every kept candidate is taken: exact re-scan when it is my turn
                    else pending = false;                // no candidate left: this query stays unmatched
                }
            }
            const bool proposing = pending && prop >= 0;
            const unsigned same = __match_any_sync(0xffffffffu, proposing ? prop : -1 - lane);
            const unsigned excl = __ballot_sync(0xffffffffu, proposing && myObs);
            const bool conflict = (proposing && (same & below & excl)) || (pending && needScan);
            const unsigned cm = __ballot_sync(0xffffffffu, conflict);
            const int c = cm ? __ffs(cm) - 1 : 32;
            const bool commit = proposing && lane < c;
            const unsigned cmask = __ballot_sync(0xffffffffu, commit);
            if (commit) {
                // several non-exclusive claims of one target: the reference lets the last query overwrite the earlier ones
                const unsigned grp = same & cmask;
                if ((grp >> lane) <= 1u) {
                    mOut[prop] = mine;
                    if (dOut) dOut[prop] = propDist;
                }
                if (myObs) atomicOr(&smem[prop >> 5], 1u << (prop & 31));
                acc[nAcc + __popc(cmask & below)] = (uint32_t)mine | ((uint32_t)prop << 16);
                pending = false;
            }
            nAcc += __popc(cmask);
            __syncwarp();
            const unsigned scanMask = __ballot_sync(0xffffffffu, pending && needScan);
            if (c < 32 && ((scanMask >> c) & 1u)) {
                // lane c's query is now the lowest undecided one and the bitmap is final for it: exact re-scan of its window,
                // the whole warp sharing the candidate runs; the winner is the smallest (distance, arrival ordinal)
                const int i = i0 + c;
                float u, v, r, ur;
                int minLevel, maxLevel;
                query_window(A, po, i, u, v, r, minLevel, maxLevel, ur);
                uint32_t qd[8];
                {
                    const uint4* p = reinterpret_cast<const uint4*>(A.desc + 32 * ((size_t)A.lRow[pair] + i));
                    const uint4 a = p[0], b = p[1];
                    qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
                }
                const uint8_t* cd = A.desc + 32 * (size_t)A.cRow[pair];
                const float* cur = A.curight ? A.curight + po : nullptr;
                const int* cs = cellStart + (size_t)pair * (GRID_CELLS + 1);
                const float4* pk = cellPack + po;
                const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, A.minX), r), A.invW)));
                const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, A.minX), r), A.invW)));
                const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, A.minY), r), A.invH)));
                const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, A.minY), r), A.invH)));
                const bool any = !(nMinCellX >= GRID_COLS || nMaxCellX < 0 || nMinCellY >= GRID_ROWS || nMaxCellY < 0);
                const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
                int ordBase = 0;
                unsigned long long bestKey = ~0ull;  // dist << 40 | ordinal << 16 | candidate index
                // The window's grid columns are contiguous runs of the packed cell list.  Lane x fetches the run of column
                // nMinCellX + x, a warp scan turns the run lengths into arrival ordinals, and the candidates of ALL columns are
                // then dealt to the lanes in one flat loop: the chain entry -> (right coordinate, descriptor) of dependent L2
                // loads is walked once or twice per re-scan instead of once or twice per column.
                const int nColsW = any ? nMaxCellX - nMinCellX + 1 : 0;
                for (int cx0 = 0; cx0 < nColsW; cx0 += 32) {
                    int rb = 0, rn = 0;
                    if (cx0 + lane < nColsW) {
                        const int ix = nMinCellX + cx0 + lane;
                        rb = cs[ix * GRID_ROWS + nMinCellY];
                        rn = cs[ix * GRID_ROWS + nMaxCellY + 1] - rb;
                    }
                    int incl = rn;
#pragma unroll
                    for (int d2 = 1; d2 < 32; d2 <<= 1) {
                        const int v2 = __shfl_up_sync(0xffffffffu, incl, d2);
                        if (lane >= d2) incl += v2;
                    }
                    const int totalC = __shfl_sync(0xffffffffu, incl, 31);
                    const int nHere = min(32, nColsW - cx0);
                    for (int t0 = 0; t0 < totalC; t0 += 32) {
                        const int t = t0 + lane;
                        // column of flat position t = number of columns whose inclusive count is <= t (every lane takes part
                        // in the shuffles, lanes past the end with a clamped column)
                        int col = 0;
                        for (int x = 0; x < nHere; ++x) col += (__shfl_sync(0xffffffffu, incl, x) <= t) ? 1 : 0;
                        col = min(col, nHere - 1);
                        const int cb = __shfl_sync(0xffffffffu, rb, col), ce = __shfl_sync(0xffffffffu, incl, col);
                        const int cn = __shfl_sync(0xffffffffu, rn, col);
                        if (t < totalC) {
                            const int jj = cb + (t - (ce - cn));
                            const float4 ent = __ldg(pk + jj);
                            const int k = __float_as_int(ent.w);
                            bool ok = true;
                            if (bCheckLevels) {
                                const int o = __float_as_int(ent.z);
                                if (o < minLevel) ok = false;
                                if (maxLevel >= 0 && o > maxLevel) ok = false;
                            }
                            const float dx = __fsub_rn(ent.x, u), dy = __fsub_rn(ent.y, v);
                            if (!(fabsf(dx) < r && fabsf(dy) < r)) ok = false;
                            if (ok && ((smem[k >> 5] >> (k & 31)) & 1u)) ok = false;
                            if (ok && cur && cur[k] > 0 && fabsf(__fsub_rn(ur, cur[k])) > r) ok = false;
                            if (ok) {
                                uint32_t td[8];
                                const uint4* pp = reinterpret_cast<const uint4*>(cd + 32 * (size_t)k);
                                const uint4 a2 = pp[0], b2 = pp[1];
                                td[0] = a2.x; td[1] = a2.y; td[2] = a2.z; td[3] = a2.w; td[4] = b2.x; td[5] = b2.y; td[6] = b2.z; td[7] = b2.w;
                                const unsigned long long key = ((unsigned long long)hamming256(qd, td) << 40) |
                                                               ((unsigned long long)(ordBase + t) << 16) | (unsigned long long)k;
                                bestKey = key < bestKey ? key : bestKey;
                            }
                        }
                    }
                    ordBase += totalC;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, bestKey, o);
                    bestKey = other < bestKey ? other : bestKey;
                }
                int bestDist = bestKey == ~0ull ? 256 : (int)(bestKey >> 40), bestIdx = bestKey == ~0ull ? -1 : (int)(bestKey & 0xffff);
                if (bestDist > A.thAccept) bestIdx = -1;  // :1428 / :1558
                if (lane == c) {
                    if (bestIdx >= 0) {
                        mOut[bestIdx] = i;
                        if (dOut) dOut[bestIdx] = bestDist;
                        if (myObs) smem[bestIdx >> 5] |= 1u << (bestIdx & 31);
                        acc[nAcc] = (uint32_t)i | ((uint32_t)bestIdx << 16);
                    }
                    pending = false;
                }
                nAcc += bestIdx >= 0 ? 1 : 0;
                __syncwarp();
            }
        }
    }
    __syncwarp();
    // rotation histogram, ComputeThreeMaxima and pruning (:1448-1469), bins computed in parallel after the loop
    int removed = 0;
    if (A.checkOri && A.histMode) {
        for (int k = lane; k < nAcc; k += 32) {
            const uint32_t a = acc[k];
            atomicAdd(&hist[rot_bin(A.langle[po + (a & 0xffff)], A.cangle[po + (a >> 16)], factor)], 1);
        }
        __syncwarp();
        int i1, i2, i3;
        three_maxima(hist, i1, i2, i3);
        __syncwarp();
        // rotHist[bin] lists Cur indices; an index overwritten by a later query (only possible when the earlier map
        // point had no observations) can sit in two bins and is cleared if either is pruned, as in the reference
        for (int k = lane; k < nAcc; k += 32) {
            const uint32_t a = acc[k];
            const int bin = rot_bin(A.langle[po + (a & 0xffff)], A.cangle[po + (a >> 16)], factor);
            if (bin != i1 && bin != i2 && bin != i3) {
                mOut[a >> 16] = -1;
                if (dOut) dOut[a >> 16] = -2;  // matched, then pruned: the reference leaves NULL here, not the old pointer
                ++removed;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    }
    if (lane == 0) nMatches[pair] = nAcc - removed;
}

// phase 2 of the ratio rule (SearchByProjection(Frame&, vector<MapPoint*>&, th), src/ORBmatcher.cc:45-129): one warp,
// queries strictly in order.  Phase 1 left, per query, the candidates with dist <= cut sorted by (dist, arrival) —
// exactly the order in which the reference's best / second-best bookkeeping ranks them — so best and second are the
// first two entries whose target is still free.  A second that phase 1 did not keep is > cut and passes the ratio test
// for every acceptable best.  A match is rejected only when best and second lie on the same pyramid level and
// best > ratio*second (:118-121).
__global__ void __launch_bounds__(32) k_win_resolve_ratio(ProjArgs A, const int* __restrict__ cellStart,
                                                          const int* __restrict__ cellIdx, const uint32_t* __restrict__ topBuf,
                                                          int* __restrict__ matchOut, int* __restrict__ distOut,
                                                          int* __restrict__ nMatches) {
    extern __shared__ uint32_t smem[];  // taken bitmap
    const int pair = blockIdx.x, lane = threadIdx.x;
    const size_t po = (size_t)pair * A.stride;
    const int nC = A.nC[pair], nL = A.nL[pair];
    int* mOut = matchOut + po;
    int* dOut = distOut ? distOut + po : nullptr;
    const int* coct = A.coct + po;
    const int words = (A.stride + 31) >> 5;
    for (int i = lane; i < words; i += 32) smem[i] = 0;
    for (int i = lane; i < nC; i += 32) { mOut[i] = -1; if (dOut) dOut[i] = -1; }
    __syncwarp();
    if (A.ctaken)
        for (int i = lane; i < nC; i += 32) if (A.ctaken[po + i]) atomicOr(&smem[i >> 5], 1u << (i & 31));
    __syncwarp();
    int nAcc = 0;
    for (int i0 = 0; i0 < nL; i0 += 32) {
        const int mine = i0 + lane;
        uint4 w0 = make_uint4(~0u, ~0u, ~0u, ~0u), w1 = make_uint4(~0u, ~0u, 0u, 0u);
        if (mine < nL) {
            const uint4* p = reinterpret_cast<const uint4*>(topBuf + (po + mine) * 8);
            w0 = p[0];
            w1 = p[1];
        }
        const int myCnt = mine < nL ? (int)w1.w : 0;
        unsigned rem = __ballot_sync(0xffffffffu, myCnt > 0);
        while (rem) {
            const int j = __ffs(rem) - 1;
            rem &= rem - 1;
            const int q = i0 + j;
            const int cnt = __shfl_sync(0xffffffffu, myCnt, j);
            const uint32_t e[TOP_K] = {__shfl_sync(0xffffffffu, w0.x, j), __shfl_sync(0xffffffffu, w0.y, j),
                                       __shfl_sync(0xffffffffu, w0.z, j), __shfl_sync(0xffffffffu, w0.w, j),
                                       __shfl_sync(0xffffffffu, w1.x, j), __shfl_sync(0xffffffffu, w1.y, j)};
            int best = 256, best2 = 256, bestIdx = -1, lvl = -1, lvl2 = -1, found = 0;
#pragma unroll
            for (int k = 0; k < TOP_K; ++k) {
                if (k < cnt && found < 2) {
                    const int t = e[k] & 0xffff, d = (int)(e[k] >> 16);
                    if (!((smem[t >> 5] >> (t & 31)) & 1u)) {
                        if (found == 0) { best = d; bestIdx = t; lvl = coct[t]; }
                        else { best2 = d; lvl2 = coct[t]; }
                        ++found;
                    }
                }
            }
            if (cnt > TOP_K && found < 2) {
                // more near candidates than the list holds and too many of the kept ones are taken: exact sequential
                // re-scan of this query (rare), every lane redundantly so the result is warp-uniform
                float u, v, r, ur;
                int minLevel, maxLevel;
                query_window(A, po, q, u, v, r, minLevel, maxLevel, ur);
                uint32_t qd[8];
                const uint4* p = reinterpret_cast<const uint4*>(A.desc + 32 * ((size_t)A.lRow[pair] + q));
                const uint4 a = p[0], b = p[1];
                qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
                const uint8_t* cd = A.desc + 32 * (size_t)A.cRow[pair];
                const float* cur = A.curight ? A.curight + po : nullptr;
                best = 256; best2 = 256; bestIdx = -1; lvl = -1; lvl2 = -1;
                for_each_candidate(A, cellStart + (size_t)pair * (GRID_CELLS + 1), cellIdx + po, A.cx + po, A.cy + po, coct,
                                   u, v, r, minLevel, maxLevel, [&](int k) {
                    if ((smem[k >> 5] >> (k & 31)) & 1u) return;
                    if (cur && cur[k] > 0 && fabsf(__fsub_rn(ur, cur[k])) > r) return;
                    uint32_t td[8];
                    const uint4* pp = reinterpret_cast<const uint4*>(cd + 32 * (size_t)k);
                    const uint4 a2 = pp[0], b2 = pp[1];
                    td[0] = a2.x; td[1] = a2.y; td[2] = a2.z; td[3] = a2.w; td[4] = b2.x; td[5] = b2.y; td[6] = b2.z; td[7] = b2.w;
                    const int d = hamming256(qd, td);
                    if (d < best) { best2 = best; best = d; lvl2 = lvl; lvl = coct[k]; bestIdx = k; }
                    else if (d < best2) { lvl2 = coct[k]; best2 = d; }
                });
            }
            if (bestIdx >= 0 && best <= A.thAccept) {
                if (!(lvl == lvl2 && (float)best > __fmul_rn(A.ratio, (float)best2))) {
                    const int obs = A.lobs ? (A.lobs[po + q] != 0) : 1;
                    if (lane == 0) {
                        mOut[bestIdx] = q;
                        if (dOut) dOut[bestIdx] = best;
                        if (obs) smem[bestIdx >> 5] |= 1u << (bestIdx & 31);
                    }
                    ++nAcc;
                    __syncwarp();
                }
            }
        }
    }
    // the reference counts accepted queries (nmatches++ per assignment, :123-124), also when a later query overwrites
    if (lane == 0) nMatches[pair] = nAcc;
}

// phase 2 of SearchForInitialization (src/ORBmatcher.cc:405-520): one warp, F1 features in index order.  A target
// that is already matched stays a candidate for a later query whose distance is strictly smaller than the stored one
// (:443-444), in which case the earlier match is undone (:465-469).  Phase 1 left the candidates with dist <= cut
// sorted by (dist, arrival); best / second best are the first two that pass the matched-distance filter.
__global__ void __launch_bounds__(32) k_init_resolve(ProjArgs A, const int* __restrict__ cellStart,
                                                     const int* __restrict__ cellIdx, const uint32_t* __restrict__ topBuf,
                                                     int* __restrict__ matched21, int* __restrict__ matchedDist,
                                                     int* __restrict__ binOf, int* __restrict__ match12,
                                                     int* __restrict__ nMatches) {
    __shared__ int hist[EAOF_HISTO_LENGTH];
    const int lane = threadIdx.x;
    const int nC = A.nC[0], nL = A.nL[0];
    const size_t po = 0;
    for (int i = lane; i < EAOF_HISTO_LENGTH; i += 32) hist[i] = 0;
    for (int i = lane; i < nC; i += 32) { matched21[i] = -1; matchedDist[i] = 0x7fffffff; }
    for (int i = lane; i < nL; i += 32) { match12[i] = -1; binOf[i] = -1; }
    __syncwarp();
    const float factor = 1.0f / EAOF_HISTO_LENGTH;  // :413
    int nm = 0;
    for (int i0 = 0; i0 < nL; i0 += 32) {
        const int mine = i0 + lane;
        uint4 w0 = make_uint4(~0u, ~0u, ~0u, ~0u), w1 = make_uint4(~0u, ~0u, 0u, 0u);
        if (mine < nL) {
            const uint4* p = reinterpret_cast<const uint4*>(topBuf + (po + mine) * 8);
            w0 = p[0];
            w1 = p[1];
        }
        const int myCnt = mine < nL ? (int)w1.w : 0;
        unsigned rem = __ballot_sync(0xffffffffu, myCnt > 0);
        while (rem) {
            const int j = __ffs(rem) - 1;
            rem &= rem - 1;
            const int q = i0 + j;
            const int cnt = __shfl_sync(0xffffffffu, myCnt, j);
            const uint32_t e[TOP_K] = {__shfl_sync(0xffffffffu, w0.x, j), __shfl_sync(0xffffffffu, w0.y, j),
                                       __shfl_sync(0xffffffffu, w0.z, j), __shfl_sync(0xffffffffu, w0.w, j),
                                       __shfl_sync(0xffffffffu, w1.x, j), __shfl_sync(0xffffffffu, w1.y, j)};
            int best = 0x7fffffff, best2 = 0x7fffffff, bestIdx = -1, found = 0;
#pragma unroll
            for (int k = 0; k < TOP_K; ++k) {
                if (k < cnt && found < 2) {
                    const int t = e[k] & 0xffff, d = (int)(e[k] >> 16);
                    if (!(matchedDist[t] <= d)) {
                        if (found == 0) { best = d; bestIdx = t; } else best2 = d;
                        ++found;
                    }
                }
            }
            if (cnt > TOP_K && found < 2) {  // exact re-scan (rare)
                float u, v, r, ur;
                int minLevel, maxLevel;
                query_window(A, po, q, u, v, r, minLevel, maxLevel, ur);
                uint32_t qd[8];
                const uint4* p = reinterpret_cast<const uint4*>(A.desc + 32 * ((size_t)A.lRow[0] + q));
                const uint4 a = p[0], b = p[1];
                qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
                const uint8_t* cd = A.desc + 32 * (size_t)A.cRow[0];
                best = 0x7fffffff; best2 = 0x7fffffff; bestIdx = -1;
                for_each_candidate(A, cellStart, cellIdx, A.cx, A.cy, A.coct, u, v, r, minLevel, maxLevel, [&](int k) {
                    uint32_t td[8];
                    const uint4* pp = reinterpret_cast<const uint4*>(cd + 32 * (size_t)k);
                    const uint4 a2 = pp[0], b2 = pp[1];
                    td[0] = a2.x; td[1] = a2.y; td[2] = a2.z; td[3] = a2.w; td[4] = b2.x; td[5] = b2.y; td[6] = b2.z; td[7] = b2.w;
                    const int d = hamming256(qd, td);
                    if (matchedDist[k] <= d) return;
                    if (d < best) { best2 = best; best = d; bestIdx = k; }
                    else if (d < best2) best2 = d;
                });
            }
            // an unseen second best is > cut, which passes the ratio test for every best <= TH_LOW
            const float second = best2 == 0x7fffffff ? 2147483648.f : (float)best2;
            if (bestIdx >= 0 && best <= EAOF_TH_LOW && (float)best < __fmul_rn(second, A.ratio)) {
                const int prev = matched21[bestIdx];
                const int bin = A.checkOri ? rot_bin(A.langle[q], A.cangle[bestIdx], factor) : -1;
                __syncwarp();
                if (lane == 0) {
                    if (prev >= 0) match12[prev] = -1;
                    match12[q] = bestIdx;
                    matched21[bestIdx] = q;
                    matchedDist[bestIdx] = best;
                    if (bin >= 0) { binOf[q] = bin; ++hist[bin]; }
                }
                nm += prev >= 0 ? 0 : 1;
                __syncwarp();
            }
        }
    }
    __syncwarp();
    if (A.checkOri) {
        int i1, i2, i3;
        three_maxima(hist, i1, i2, i3);
        int removed = 0;
        for (int i = lane; i < nL; i += 32) {
            const int bin = binOf[i];
            if (bin >= 0 && bin != i1 && bin != i2 && bin != i3 && match12[i] >= 0) { match12[i] = -1; ++removed; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nm -= removed;
    }
    if (lane == 0) *nMatches = nm;
}

// Queries of the consecutive-frame path: Last keypoints of an extractor batch shifted by the known motion.
__global__ void k_proj_prepare(const eaof_kp* __restrict__ kps, const int* __restrict__ counts, int cap,
                               const int* __restrict__ lastFrame, const int* __restrict__ curFrame,
                               const float* __restrict__ shiftX, const float* __restrict__ shiftY, int stride,
                               float* cx, float* cy, int* coct, float* cangle, float* lu, float* lv, int* loct,
                               float* langle, int* nC, int* nL, int* cRow, int* lRow) {
    const int pair = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const int fl = lastFrame[pair], fc = curFrame[pair];
    const size_t po = (size_t)pair * stride;
    if (i == 0) { nC[pair] = counts[fc]; nL[pair] = counts[fl]; cRow[pair] = fc * cap; lRow[pair] = fl * cap; }
    if (i < counts[fc]) {
        const eaof_kp k = kps[(size_t)fc * cap + i];
        cx[po + i] = k.x; cy[po + i] = k.y; coct[po + i] = k.octave; cangle[po + i] = k.angle;
    }
    if (i < counts[fl]) {
        const eaof_kp k = kps[(size_t)fl * cap + i];
        lu[po + i] = __fadd_rn(k.x, shiftX[pair]); lv[po + i] = __fadd_rn(k.y, shiftY[pair]);
        loct[po + i] = k.octave; langle[po + i] = k.angle;
    }
}


// POPC issue-rate probe (the roofline of the Hamming kernels): 8 independent XOR+POPC+ADD chains per
// thread, enough warps to saturate every scheduler.
__global__ void __launch_bounds__(256) k_popc_probe(uint32_t seed, int iters, uint32_t* out) {
    uint32_t a[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + 1u) + 0x9e3779b9u * (i + 1u); acc[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += __popc(a[i] ^ acc[i]);
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= acc[i];
    if (r == 0x12345678u) out[0] = r;  // never true in practice; keeps the chains alive
}
}  // namespace

struct eaof_matcher {
    int device = 0, maxPairs = 0, maxFeat = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evDep = nullptr;
    uint32_t *nearBuf = nullptr, *accBuf = nullptr;
    // tcgen05 path of the brute-force matcher (bow_umma.cuh): +-1-expanded copy of a descriptor array and its tensor maps
    int8_t* umma8 = nullptr;
    size_t ummaCap = 0;                 // descriptors the buffer holds
    const uint8_t* ummaSrc = nullptr;   // array the expansion was made from (valid between prepare and release)
    int ummaBlocks = 0, ummaStride = 0;
    eaof_umma::TMap ummaMapA{}, ummaMapB{};
    void* encodeTiled = nullptr;
    int sms = 0;
    bool ummaOn = true;                 // EAOF_BOW_UMMA=0: POPC kernel (k_bow_dense)
    int *cellStart = nullptr, *cellIdx = nullptr;
    float4* cellPack = nullptr;  // grid lists as (x, y, octave, index) entries
    // SoA staging, [maxPairs][maxFeat]
    float *cx = nullptr, *cy = nullptr, *cangle = nullptr, *curight = nullptr, *lu = nullptr, *lv = nullptr,
          *linvz = nullptr, *langle = nullptr;
    int *coct = nullptr, *loct = nullptr, *nC = nullptr, *nL = nullptr, *cRow = nullptr, *lRow = nullptr;
    uint8_t *ctaken = nullptr, *lvalid = nullptr, *lobs = nullptr;
    float* qRadius = nullptr;   // eaof_match_windows: per-query radius and upper level
    int* qMaxL = nullptr;
    int* initBin = nullptr;     // eaof_match_initialization: rotation bin per accepted F1 feature
    // single-pair host API staging
    uint8_t* desc2 = nullptr;   // 2 blocks of maxFeat descriptors
    float* angle2 = nullptr;    // 2 blocks of maxFeat angles
    uint8_t *validQ = nullptr, *validT = nullptr;
    int *idxQ = nullptr, *idxT = nullptr;
    BowSeg* segs = nullptr;
    int* segStart = nullptr;
    int2* tiles = nullptr;
    int *pairIdx = nullptr;     // [4][maxPairs] device copies of pair arrays / shifts
    float* pairShift = nullptr; // [2][maxPairs]
    int *outMatch = nullptr, *outDist = nullptr, *outN = nullptr;
    // eaof_match_bow_orb_device: per-pair plans and the extractor's keypoint angles as a plain array (lazily allocated)
    BowSeg* segsBatch = nullptr;  // [maxPairs][maxFeat]
    int* segCountBatch = nullptr; // [maxPairs]
    float* orbAngle = nullptr;    // [(maxPairs+1) * maxFeat]
    long long lastDistances = 0;
    // Small per-call host arrays (pair lists, shifts) go through a pinned staging buffer: cudaMemcpyAsync from pageable
    // memory would synchronise the stream first and stall the caller behind the extraction the stream waits for.
    // Unchanged arrays (a sequence matched batch after batch) are not uploaded again.
    // single-pair host API: every input array of a call is gathered in ONE pinned block and uploaded with ONE copy (a
    // cudaMemcpyAsync per array out of pageable memory costs more than the kernels), outputs come back the same way
    uint8_t *upH = nullptr, *upD = nullptr;
    size_t upCap = 0, upOff = 0;
    int *outArenaD = nullptr, *outArenaH = nullptr;  // [4 + 2*maxFeat]: n | match | dist
    template <typename T> T* stage(const T* src, size_t n) {
        const size_t bytes = sizeof(T) * n;
        T* dst = reinterpret_cast<T*>(upD + upOff);
        memcpy(upH + upOff, src, bytes);
        upOff += (bytes + 15) & ~(size_t)15;
        return dst;
    }
    cudaError_t flush_stage() {
        const cudaError_t e = cudaMemcpyAsync(upD, upH, upOff, cudaMemcpyHostToDevice, stream);
        upOff = 0;
        return e;
    }
    int* hStage = nullptr;            // [4][maxPairs] words
    cudaEvent_t evStage = nullptr;    // last upload out of hStage has completed
    std::vector<int> lastUp[4];
    int upload_words(int slot, const void* src, int n, void* dst) {
        const int* w = static_cast<const int*>(src);
        if ((int)lastUp[slot].size() == n && memcmp(lastUp[slot].data(), w, sizeof(int) * (size_t)n) == 0) return 0;
        if (cudaEventSynchronize(evStage) != cudaSuccess) return -1;
        memcpy(hStage + (size_t)slot * maxPairs, w, sizeof(int) * (size_t)n);
        if (cudaMemcpyAsync(dst, hStage + (size_t)slot * maxPairs, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, stream) != cudaSuccess) return -1;
        if (cudaEventRecord(evStage, stream) != cudaSuccess) return -1;
        lastUp[slot].assign(w, w + n);
        return 0;
    }
};

namespace {
// workspace arrays start out zeroed: tails that no kernel writes are never stale memory when an output block is downloaded
template <typename T> cudaError_t dalloc(T** p, size_t n) {
    const size_t bytes = sizeof(T) * (n ? n : 1);
    cudaError_t e = cudaMalloc(p, bytes);
    return e != cudaSuccess ? e : cudaMemset(*p, 0, bytes);
}

int near_threshold(int thEff, float ratio) {
    int s = 0;
    while (s <= 256 && !((float)thEff < ratio * (float)s)) ++s;  // s_min
    int D = s > thEff + 1 ? s : thEff + 1;
    return D > 257 ? 257 : D;
}
}  // namespace

extern "C" {

int eaof_matcher_create(int device, int maxPairs, int maxFeat, eaof_matcher** out) {
    if (!out || maxPairs < 1 || maxFeat < 1 || maxFeat > 65535) return mfail(EAOF_ERR_ARG, "bad matcher size (max_features must be in [1,65535])");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return mfail(EAOF_ERR_CUDA, "no CUDA device: libeaof_orb has no CPU fallback");
    if (device < 0 || device >= ndev) return mfail(EAOF_ERR_ARG, "device out of range");
    MCK(cudaSetDevice(device));
    eaof_matcher* m = new eaof_matcher;
    m->device = device; m->maxPairs = maxPairs; m->maxFeat = maxFeat;
    const size_t PF = (size_t)maxPairs * maxFeat;
    cudaError_t e = cudaSuccess;
#define A_(x) if (e == cudaSuccess) e = (x)
    A_(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            m->encodeTiled = fn;
        else
            cudaGetLastError();
        cudaDeviceGetAttribute(&m->sms, cudaDevAttrMultiProcessorCount, device);
        if (const char* e = getenv("EAOF_BOW_UMMA")) if (*e) m->ummaOn = atoi(e) != 0;
    }
    A_(cudaEventCreateWithFlags(&m->evDep, cudaEventDisableTiming));
    A_(cudaEventCreateWithFlags(&m->evStage, cudaEventDisableTiming));
    A_(cudaMallocHost(&m->hStage, sizeof(int) * 4 * (size_t)maxPairs));
    A_(dalloc(&m->nearBuf, PF * 8)); A_(dalloc(&m->accBuf, PF));
    A_(dalloc(&m->cellStart, (size_t)maxPairs * (GRID_CELLS + 1))); A_(dalloc(&m->cellIdx, PF)); A_(dalloc(&m->cellPack, PF));
    A_(dalloc(&m->cx, PF)); A_(dalloc(&m->cy, PF)); A_(dalloc(&m->cangle, PF)); A_(dalloc(&m->curight, PF));
    A_(dalloc(&m->lu, PF)); A_(dalloc(&m->lv, PF)); A_(dalloc(&m->linvz, PF)); A_(dalloc(&m->langle, PF));
    A_(dalloc(&m->coct, PF)); A_(dalloc(&m->loct, PF));
    A_(dalloc(&m->nC, maxPairs)); A_(dalloc(&m->nL, maxPairs)); A_(dalloc(&m->cRow, maxPairs)); A_(dalloc(&m->lRow, maxPairs));
    A_(dalloc(&m->ctaken, PF)); A_(dalloc(&m->lvalid, PF)); A_(dalloc(&m->lobs, PF));
    A_(dalloc(&m->qRadius, (size_t)maxFeat)); A_(dalloc(&m->qMaxL, (size_t)maxFeat)); A_(dalloc(&m->initBin, (size_t)maxFeat));
    A_(dalloc(&m->desc2, (size_t)2 * maxFeat * 32)); A_(dalloc(&m->angle2, (size_t)2 * maxFeat));
    A_(dalloc(&m->validQ, maxFeat)); A_(dalloc(&m->validT, maxFeat)); A_(dalloc(&m->idxQ, maxFeat)); A_(dalloc(&m->idxT, maxFeat));
    A_(dalloc(&m->segs, maxFeat)); A_(dalloc(&m->segStart, 2)); A_(dalloc(&m->tiles, (size_t)2 * maxFeat));
    A_(dalloc(&m->pairIdx, (size_t)4 * maxPairs)); A_(dalloc(&m->pairShift, (size_t)2 * maxPairs));
    A_(dalloc(&m->outMatch, PF)); A_(dalloc(&m->outDist, PF)); A_(dalloc(&m->outN, maxPairs));
    m->upCap = (size_t)maxFeat * 160 + 4096;  // 2 x 32-byte descriptors + up to 20 4-byte arrays per feature, 16-byte aligned each
    A_(cudaMallocHost(&m->upH, m->upCap)); A_(dalloc(&m->upD, m->upCap));
    A_(cudaMallocHost(&m->outArenaH, sizeof(int) * (4 + 2 * (size_t)maxFeat))); A_(dalloc(&m->outArenaD, 4 + 2 * (size_t)maxFeat));
#undef A_
    if (e != cudaSuccess) {
        mfail(EAOF_ERR_CUDA, "matcher allocation failed: %s", cudaGetErrorString(e));
        eaof_matcher_destroy(m);
        return EAOF_ERR_CUDA;
    }
    *out = m;
    return EAOF_OK;
}

void eaof_matcher_destroy(eaof_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    void* ptrs[] = {m->nearBuf, m->accBuf, m->cellStart, m->cellIdx, m->cx, m->cy, m->cangle, m->curight, m->lu, m->lv,
                    m->linvz, m->langle, m->coct, m->loct, m->nC, m->nL, m->cRow, m->lRow, m->ctaken, m->lvalid, m->lobs,
                    m->desc2, m->angle2, m->validQ, m->validT, m->idxQ, m->idxT, m->segs, m->segStart, m->tiles,
                    m->pairIdx, m->pairShift, m->outMatch, m->outDist, m->outN, m->qRadius, m->qMaxL, m->initBin, m->cellPack,
                    m->segsBatch, m->segCountBatch, m->orbAngle};
    for (void* p : ptrs) cudaFree(p);
    cudaFree(m->umma8);
    if (m->evDep) cudaEventDestroy(m->evDep);
    if (m->evStage) cudaEventDestroy(m->evStage);
    cudaFreeHost(m->hStage);
    cudaFreeHost(m->upH); cudaFree(m->upD); cudaFreeHost(m->outArenaH); cudaFree(m->outArenaD);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

void* eaof_matcher_stream(eaof_matcher* m) { return m ? (void*)m->stream : nullptr; }
int eaof_matcher_sync(eaof_matcher* m) {
    if (!m) return mfail(EAOF_ERR_ARG, "null matcher");
    MCK(cudaStreamSynchronize(m->stream));
    return EAOF_OK;
}
long long eaof_matcher_last_distance_count(const eaof_matcher* m) { return m ? m->lastDistances : 0; }

int eaof_hamming_distances(eaof_matcher* m, const uint8_t* a, const uint8_t* b, int n, int* out) {
    if (!m || !a || !b || !out || n < 0) return mfail(EAOF_ERR_ARG, "bad argument");
    if (n == 0) return EAOF_OK;
    MCK(cudaSetDevice(m->device));
    uint8_t *da = nullptr, *db = nullptr;
    int* dout = nullptr;
    MCK(cudaMalloc(&da, 32 * (size_t)n)); MCK(cudaMalloc(&db, 32 * (size_t)n)); MCK(cudaMalloc(&dout, sizeof(int) * (size_t)n));
    MCK(cudaMemcpyAsync(da, a, 32 * (size_t)n, cudaMemcpyHostToDevice, m->stream));
    MCK(cudaMemcpyAsync(db, b, 32 * (size_t)n, cudaMemcpyHostToDevice, m->stream));
    k_hamming_pairs<<<(n + 255) / 256, 256, 0, m->stream>>>(da, db, n, dout);
    MCK(cudaMemcpyAsync(out, dout, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    m->lastDistances = n;
    return EAOF_OK;
}

int eaof_match_bow(eaof_matcher* m, int mode, float ratio, int checkOri, int nQ, const uint8_t* descQ, const float* angleQ,
                   const uint8_t* validQ, int nT, const uint8_t* descT, const float* angleT, const uint8_t* validT,
                   int nNodesQ, const int* nodeIdQ, const int* nodeStartQ, const int* nodeIdxQ, int nNodesT,
                   const int* nodeIdT, const int* nodeStartT, const int* nodeIdxT, int* matchOut, int* distOut,
                   int* nMatches) {
    if (!m || !matchOut || !nMatches || nQ < 0 || nT < 0) return mfail(EAOF_ERR_ARG, "bad argument");
    if (mode != EAOF_BOW_KF_FRAME && mode != EAOF_BOW_KF_KF) return mfail(EAOF_ERR_ARG, "unknown mode");
    if (nQ > m->maxFeat || nT > m->maxFeat) return mfail(EAOF_ERR_ARG, "feature count exceeds max_features=%d", m->maxFeat);
    if ((nQ && (!descQ || !angleQ)) || (nT && (!descT || !angleT))) return mfail(EAOF_ERR_ARG, "null descriptor/angle array");
    const int nOut = mode == EAOF_BOW_KF_FRAME ? nT : nQ;
    *nMatches = 0;
    for (int i = 0; i < nOut; ++i) { matchOut[i] = -1; if (distOut) distOut[i] = -1; }
    if (nQ == 0 || nT == 0) return EAOF_OK;
    // merge-walk of the two feature vectors (:182-264): segments of nodes present on both sides, ascending node id
    std::vector<BowSeg> segs;
    std::vector<int2> tiles;
    long long dists = 0;
    {
        int a = 0, b = 0;
        while (a < nNodesQ && b < nNodesT) {
            if (nodeIdQ[a] == nodeIdT[b]) {
                BowSeg s{nodeStartQ[a], nodeStartQ[a + 1] - nodeStartQ[a], nodeStartT[b], nodeStartT[b + 1] - nodeStartT[b]};
                if (s.qCnt > 0 && s.tCnt > 0) {
                    for (int q0 = s.qOff; q0 < s.qOff + s.qCnt; q0 += BOW_QT) tiles.push_back(make_int2((int)segs.size(), q0));
                    segs.push_back(s);
                    dists += (long long)s.qCnt * s.tCnt;
                }
                ++a; ++b;
            } else if (nodeIdQ[a] < nodeIdT[b]) ++a;
            else ++b;
        }
    }
    if (segs.empty()) return EAOF_OK;
    const int nIdxQ = nodeStartQ[nNodesQ], nIdxT = nodeStartT[nNodesT];
    if (nIdxQ > m->maxFeat || nIdxT > m->maxFeat || (int)segs.size() > m->maxFeat || (int)tiles.size() > 2 * m->maxFeat)
        return mfail(EAOF_ERR_ARG, "feature vector larger than max_features");
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    MCK(cudaMemcpyAsync(m->desc2, descQ, 32 * (size_t)nQ, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(m->desc2 + 32 * (size_t)m->maxFeat, descT, 32 * (size_t)nT, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(m->angle2, angleQ, sizeof(float) * nQ, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(m->angle2 + m->maxFeat, angleT, sizeof(float) * nT, cudaMemcpyHostToDevice, s));
    if (validQ) MCK(cudaMemcpyAsync(m->validQ, validQ, nQ, cudaMemcpyHostToDevice, s));
    if (validT) MCK(cudaMemcpyAsync(m->validT, validT, nT, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(m->idxQ, nodeIdxQ, sizeof(int) * nIdxQ, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(m->idxT, nodeIdxT, sizeof(int) * nIdxT, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(m->segs, segs.data(), sizeof(BowSeg) * segs.size(), cudaMemcpyHostToDevice, s));
    const int segStart[2] = {0, (int)segs.size()};
    MCK(cudaMemcpyAsync(m->segStart, segStart, sizeof segStart, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(m->tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, s));
    BowArgs A{};
    A.desc = m->desc2; A.angle = m->angle2; A.counts = nullptr; A.pairQ = nullptr; A.pairT = nullptr;
    A.blockStride = m->maxFeat; A.validQ = validQ ? m->validQ : nullptr; A.validT = validT ? m->validT : nullptr;
    A.idxQ = m->idxQ; A.idxT = m->idxT; A.segs = m->segs; A.segStart = m->segStart; A.nQhost = nQ; A.nThost = nT;
    A.stride = m->maxFeat; A.mode = mode; A.thEff = mode == EAOF_BOW_KF_FRAME ? EAOF_TH_LOW : EAOF_TH_LOW - 1;
    A.D = near_threshold(A.thEff, ratio); A.ratio = ratio; A.checkOri = checkOri;
    MCK(cudaMemsetAsync(m->nearBuf, 0, sizeof(uint32_t) * 8 * (size_t)m->maxFeat, s));
    k_bow_dense<<<dim3((unsigned)tiles.size(), 1), BOW_QT, 0, s>>>(A, m->tiles, m->nearBuf);
    A.spec = m->maxFeat <= 11000;
    const size_t bm = sizeof(uint32_t) * ((m->maxFeat + 31) / 32 + (A.spec ? (size_t)m->maxFeat : 0));  // matched bitmap + proposal owners
    k_bow_resolve<<<1, 32, bm, s>>>(A, m->nearBuf, m->accBuf, m->outMatch, m->outDist, m->outN);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(matchOut, m->outMatch, sizeof(int) * nOut, cudaMemcpyDeviceToHost, s));
    if (distOut) MCK(cudaMemcpyAsync(distOut, m->outDist, sizeof(int) * nOut, cudaMemcpyDeviceToHost, s));
    MCK(cudaMemcpyAsync(nMatches, m->outN, sizeof(int), cudaMemcpyDeviceToHost, s));
    MCK(cudaStreamSynchronize(s));
    m->lastDistances = dists;
    return EAOF_OK;
}

int eaof_match_triangulation(eaof_matcher* m, int checkOri, int onlyStereo, int n1, const uint8_t* desc1, const float* x1,
                             const float* y1, const float* angle1, const uint8_t* free1, const uint8_t* stereo1, int n2,
                             const uint8_t* desc2, const float* x2, const float* y2, const int* oct2, const float* angle2,
                             const uint8_t* free2, const uint8_t* stereo2, int nNodes1, const int* nodeId1,
                             const int* nodeStart1, const int* nodeIdx1, int nNodes2, const int* nodeId2,
                             const int* nodeStart2, const int* nodeIdx2, const float* F12, float ex, float ey,
                             const float* scaleFactors2, const float* levelSigma2, int nLevels, int* match12, int* dist12,
                             int* nMatches) {
    if (!m || !match12 || !nMatches || n1 < 0 || n2 < 0) return mfail(EAOF_ERR_ARG, "bad argument");
    if (n1 > m->maxFeat || n2 > m->maxFeat) return mfail(EAOF_ERR_ARG, "feature count exceeds max_features=%d", m->maxFeat);
    if (nLevels < 1 || nLevels > EAOF_MAX_LEVELS || !scaleFactors2 || !levelSigma2 || !F12) return mfail(EAOF_ERR_ARG, "bad scale table / F12");
    *nMatches = 0;
    for (int i = 0; i < n1; ++i) { match12[i] = -1; if (dist12) dist12[i] = -1; }
    if (n1 == 0 || n2 == 0) return EAOF_OK;
    if (!desc1 || !x1 || !y1 || !angle1 || !free1 || !desc2 || !x2 || !y2 || !oct2 || !angle2 || !free2) return mfail(EAOF_ERR_ARG, "null array");
    for (int i = 0; i < n2; ++i) if (oct2[i] < 0 || oct2[i] >= nLevels) return mfail(EAOF_ERR_ARG, "octave2[%d] out of range", i);
    std::vector<BowSeg> segs;
    std::vector<int2> tiles;
    long long dists = 0;
    {
        int a = 0, b = 0;  // merge-walk of the two feature vectors, :690-783
        while (a < nNodes1 && b < nNodes2) {
            if (nodeId1[a] == nodeId2[b]) {
                BowSeg s{nodeStart1[a], nodeStart1[a + 1] - nodeStart1[a], nodeStart2[b], nodeStart2[b + 1] - nodeStart2[b]};
                if (s.qCnt > 0 && s.tCnt > 0) {
                    for (int q0 = s.qOff; q0 < s.qOff + s.qCnt; q0 += BOW_QT) tiles.push_back(make_int2((int)segs.size(), q0));
                    segs.push_back(s);
                    dists += (long long)s.qCnt * s.tCnt;
                }
                ++a; ++b;
            } else if (nodeId1[a] < nodeId2[b]) ++a;
            else ++b;
        }
    }
    if (segs.empty()) return EAOF_OK;
    const int nIdx1 = nodeStart1[nNodes1], nIdx2 = nodeStart2[nNodes2];
    if (nIdx1 > m->maxFeat || nIdx2 > m->maxFeat || (int)segs.size() > m->maxFeat || (int)tiles.size() > 2 * m->maxFeat)
        return mfail(EAOF_ERR_ARG, "feature vector larger than max_features");
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
#define UP(dst, src, n, T) MCK(cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)(n), cudaMemcpyHostToDevice, s))
    UP(m->desc2, desc1, 32 * (size_t)n1, uint8_t);
    UP(m->desc2 + 32 * (size_t)m->maxFeat, desc2, 32 * (size_t)n2, uint8_t);
    UP(m->cx, x1, n1, float); UP(m->cy, y1, n1, float); UP(m->cangle, angle1, n1, float); UP(m->validQ, free1, n1, uint8_t);
    if (stereo1) UP(m->ctaken, stereo1, n1, uint8_t);
    UP(m->lu, x2, n2, float); UP(m->lv, y2, n2, float); UP(m->loct, oct2, n2, int); UP(m->langle, angle2, n2, float);
    UP(m->validT, free2, n2, uint8_t);
    if (stereo2) UP(m->lvalid, stereo2, n2, uint8_t);
    UP(m->idxQ, nodeIdx1, nIdx1, int); UP(m->idxT, nodeIdx2, nIdx2, int);
    UP(m->segs, segs.data(), segs.size(), BowSeg); UP(m->tiles, tiles.data(), tiles.size(), int2);
#undef UP
    TriArgs A{};
    A.desc = m->desc2; A.blockStride = m->maxFeat;
    A.x1 = m->cx; A.y1 = m->cy; A.angle1 = m->cangle; A.free1 = m->validQ; A.stereo1 = stereo1 ? m->ctaken : nullptr;
    A.x2 = m->lu; A.y2 = m->lv; A.oct2 = m->loct; A.angle2 = m->langle; A.free2 = m->validT; A.stereo2 = stereo2 ? m->lvalid : nullptr;
    A.idx1 = m->idxQ; A.idx2 = m->idxT; A.segs = m->segs;
    for (int i = 0; i < 9; ++i) A.F[i] = F12[i];
    A.ex = ex; A.ey = ey;
    for (int i = 0; i < nLevels; ++i) { A.scale2[i] = scaleFactors2[i]; A.sigma2[i] = levelSigma2[i]; }
    A.onlyStereo = onlyStereo; A.checkOri = checkOri; A.n1 = n1;
    MCK(cudaMemsetAsync(m->outMatch, 0xff, sizeof(int) * (size_t)n1, s));
    MCK(cudaMemsetAsync(m->outDist, 0xff, sizeof(int) * (size_t)n1, s));
    k_tri_dense<<<(unsigned)tiles.size(), BOW_QT, 0, s>>>(A, m->tiles, m->outMatch, m->outDist);
    k_tri_finish<<<1, 256, 0, s>>>(A, m->outMatch, m->outDist, m->outN);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(match12, m->outMatch, sizeof(int) * n1, cudaMemcpyDeviceToHost, s));
    if (dist12) MCK(cudaMemcpyAsync(dist12, m->outDist, sizeof(int) * n1, cudaMemcpyDeviceToHost, s));
    MCK(cudaMemcpyAsync(nMatches, m->outN, sizeof(int), cudaMemcpyDeviceToHost, s));
    MCK(cudaStreamSynchronize(s));
    m->lastDistances = dists;
    return EAOF_OK;
}

// Measured POPC32 thread-instructions per second on `device` (all SMs busy), or a negative value on error.
double eaof_debug_popc_rate(int device) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, iters = 4096;
    k_popc_probe<<<blocks, 256>>>(1u, 64, d);
    double best = -1.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_popc_probe<<<blocks, 256>>>(2u + rep, iters, d);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)blocks * 256 * 8.0 * iters / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}

// ---- tcgen05 path of phase 1 (bow_umma.cuh) ----------------------------------------------------------------------------
// prepare: +-1 expansion of blocks [0, nBlocks) of a descriptor array on the matcher's stream + tensor maps over it; valid
// until release (the array may be rewritten by the caller between calls, so nothing is cached across API calls).
// Returns false when the path is unavailable (no cuTensorMapEncodeTiled, switched off, allocation failure): callers fall
// back to k_bow_dense.
static bool umma_prepare(eaof_matcher* m, const uint8_t* dDesc, int nBlocks, int blockStride) {
    m->ummaSrc = nullptr;
    if (!m->ummaOn || !m->encodeTiled || nBlocks < 1 || m->sms < 1) return false;
    const size_t nDesc = (size_t)nBlocks * blockStride;
    if (nDesc >= ((size_t)1 << 31)) return false;
    if (nDesc > m->ummaCap) {
        cudaStreamSynchronize(m->stream);
        cudaFree(m->umma8);
        m->umma8 = nullptr; m->ummaCap = 0;
        if (cudaMalloc(&m->umma8, nDesc * 256) != cudaSuccess) { cudaGetLastError(); return false; }
        m->ummaCap = nDesc;
    }
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    for (int which = 0; which < 2; ++which) {
        const cuuint64_t dims[2] = {256, (cuuint64_t)nDesc};
        const cuuint64_t strides[1] = {256};
        const cuuint32_t box[2] = {128, (cuuint32_t)(which ? eaof_umma::kTileT : eaof_umma::kTileQ)}, es[2] = {1, 1};
        if (((EncodeTiled)m->encodeTiled)(reinterpret_cast<CUtensorMap*>(which ? &m->ummaMapB : &m->ummaMapA), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                                          m->umma8, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    static bool attrSet[64] = {};
    if (m->device < 64 && !attrSet[m->device]) {
        if (cudaFuncSetAttribute(eaof_umma::k_bow_dense_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, eaof_umma::kSmemBytes) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attrSet[m->device] = true;
    }
    eaof_umma::k_expand_pm1<<<(unsigned)((nDesc * 16 + 255) / 256), 256, 0, m->stream>>>(dDesc, m->umma8, nDesc);
    if (cudaGetLastError() != cudaSuccess) return false;
    m->ummaSrc = dDesc; m->ummaBlocks = nBlocks; m->ummaStride = blockStride;
    return true;
}
static void umma_release(eaof_matcher* m) { m->ummaSrc = nullptr; }
// phase 1 of a brute-force chunk: tensor cores when the expansion of this array is in place, POPC kernel otherwise
static void bruteforce_dense(eaof_matcher* m, const BowArgs& A, int nPairs, const uint8_t* dDesc, int blockStride) {
    cudaStream_t s = m->stream;
    if (m->ummaSrc == dDesc && m->ummaStride == blockStride) {
        eaof_umma::Args U{A.pairQ, A.pairT, A.counts, blockStride, nPairs, (blockStride + eaof_umma::kTileQ - 1) / eaof_umma::kTileQ, A.D};
        eaof_umma::k_bow_dense_umma<<<m->sms, eaof_umma::kThreads, eaof_umma::kSmemBytes, s>>>(m->ummaMapA, m->ummaMapB, U, m->nearBuf);
    } else {
        k_bow_dense<<<dim3((blockStride + BOW_QT - 1) / BOW_QT, nPairs), BOW_QT, 0, s>>>(A, nullptr, m->nearBuf);
    }
}
extern "C" int eaof_internal_bruteforce_prepare(eaof_matcher* m, const uint8_t* dDesc, int nBlocks, int blockStride) {
    if (!m || !dDesc) return mfail(EAOF_ERR_ARG, "null argument");
    MCK(cudaSetDevice(m->device));
    umma_prepare(m, dDesc, nBlocks, blockStride);  // false = the POPC kernel serves the chunks
    return EAOF_OK;
}
extern "C" int eaof_internal_bruteforce_release(eaof_matcher* m) {
    if (!m) return mfail(EAOF_ERR_ARG, "null argument");
    umma_release(m);
    return EAOF_OK;
}

int eaof_match_bruteforce_batch_device(eaof_matcher* m, int mode, float ratio, int checkOri, int nPairs, const int* pairQ,
                                       const int* pairT, const uint8_t* dDesc, const float* dAngle, const int* dCounts,
                                       int blockStride, int* dMatch, int* dDist, int* dN) {
    if (!m || !pairQ || !pairT || !dDesc || !dAngle || !dCounts || !dMatch || !dN) return mfail(EAOF_ERR_ARG, "null argument");
    if (nPairs < 1 || nPairs > m->maxPairs) return mfail(EAOF_ERR_ARG, "n_pairs %d outside [1,%d]", nPairs, m->maxPairs);
    if (blockStride < 1 || blockStride > m->maxFeat) return mfail(EAOF_ERR_ARG, "block_stride exceeds max_features");
    if (mode != EAOF_BOW_KF_FRAME && mode != EAOF_BOW_KF_KF) return mfail(EAOF_ERR_ARG, "unknown mode");
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    if (m->upload_words(0, pairQ, nPairs, m->pairIdx) || m->upload_words(1, pairT, nPairs, m->pairIdx + m->maxPairs))
        return mfail(EAOF_ERR_CUDA, "pair list upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    BowArgs A{};
    A.desc = dDesc; A.angle = dAngle; A.counts = dCounts; A.pairQ = m->pairIdx; A.pairT = m->pairIdx + m->maxPairs;
    A.blockStride = blockStride; A.stride = blockStride; A.mode = mode;
    A.thEff = mode == EAOF_BOW_KF_FRAME ? EAOF_TH_LOW : EAOF_TH_LOW - 1;
    A.D = near_threshold(A.thEff, ratio); A.ratio = ratio; A.checkOri = checkOri;
    {
        int maxBlock = 0;
        for (int p = 0; p < nPairs; ++p) {
            if (pairQ[p] < 0 || pairT[p] < 0) return mfail(EAOF_ERR_ARG, "negative block index in pair %d", p);
            maxBlock = std::max(maxBlock, std::max(pairQ[p], pairT[p]));
        }
        // worth the expansion (256 B written per descriptor) when the pairs outnumber the blocks they touch
        if (nPairs >= 4 && nPairs * 2 >= maxBlock + 1) umma_prepare(m, dDesc, maxBlock + 1, blockStride);
    }
    bruteforce_dense(m, A, nPairs, dDesc, blockStride);
    umma_release(m);
    A.spec = blockStride <= 11000;
    const size_t bm = sizeof(uint32_t) * ((blockStride + 31) / 32 + (A.spec ? (size_t)blockStride : 0));
    k_bow_resolve<<<nPairs, 32, bm, s>>>(A, m->nearBuf, m->accBuf, dMatch, dDist, dN);
    MCK(cudaGetLastError());
    m->lastDistances = -1;  // counts live on the device; the caller knows nQ*nT per pair
    return EAOF_OK;
}

// internal (eaof_sweep.cu): the same two kernels over a chunk of a pair list that already lives on the device
int eaof_internal_bruteforce_pairs_device(eaof_matcher* m, int mode, float ratio, int checkOri, int nPairs, const int* dPairQ,
                                          const int* dPairT, const uint8_t* dDesc, const float* dAngle, const int* dCounts,
                                          int blockStride, int* dMatch, int* dDist, int* dN) {
    if (!m || !dPairQ || !dPairT || !dDesc || !dAngle || !dCounts || !dMatch || !dN) return mfail(EAOF_ERR_ARG, "null argument");
    if (nPairs < 1 || nPairs > m->maxPairs) return mfail(EAOF_ERR_ARG, "n_pairs %d outside [1,%d]", nPairs, m->maxPairs);
    if (blockStride < 1 || blockStride > m->maxFeat) return mfail(EAOF_ERR_ARG, "block_stride exceeds max_features");
    if (mode != EAOF_BOW_KF_FRAME && mode != EAOF_BOW_KF_KF) return mfail(EAOF_ERR_ARG, "unknown mode");
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    BowArgs A{};
    A.desc = dDesc; A.angle = dAngle; A.counts = dCounts; A.pairQ = dPairQ; A.pairT = dPairT;
    A.blockStride = blockStride; A.stride = blockStride; A.mode = mode;
    A.thEff = mode == EAOF_BOW_KF_FRAME ? EAOF_TH_LOW : EAOF_TH_LOW - 1;
    A.D = near_threshold(A.thEff, ratio); A.ratio = ratio; A.checkOri = checkOri;
    bruteforce_dense(m, A, nPairs, dDesc, blockStride);  // tensor cores when eaof_internal_bruteforce_prepare expanded this array
    A.spec = blockStride <= 11000;
    const size_t bm = sizeof(uint32_t) * ((blockStride + 31) / 32 + (A.spec ? (size_t)blockStride : 0));
    k_bow_resolve<<<nPairs, 32, bm, s>>>(A, m->nearBuf, m->accBuf, dMatch, dDist, dN);
    MCK(cudaGetLastError());
    m->lastDistances = -1;
    return EAOF_OK;
}
int eaof_internal_matcher_limits(const eaof_matcher* m, int* maxPairs, int* maxFeat, int* device) {
    if (!m) return mfail(EAOF_ERR_ARG, "null matcher handle");
    *maxPairs = m->maxPairs; *maxFeat = m->maxFeat; *device = m->device;
    return EAOF_OK;
}

int eaof_match_bow_orb_device(eaof_matcher* m, eaof_orb* ex, int nFrames, int mode, float ratio, int checkOri, int nPairs,
                              const int* pairQ, const int* pairT, const int* dNFNodes, const uint32_t* dNodeIds,
                              const int* dNodeStart, const uint32_t* dFeatIdx, int* dMatch, int* dDist, int* dN) {
    if (!m || !ex || !pairQ || !pairT || !dNFNodes || !dNodeIds || !dNodeStart || !dFeatIdx || !dMatch || !dN)
        return mfail(EAOF_ERR_ARG, "null argument");
    if (nPairs < 1 || nPairs > m->maxPairs) return mfail(EAOF_ERR_ARG, "n_pairs %d outside [1,%d]", nPairs, m->maxPairs);
    if (mode != EAOF_BOW_KF_FRAME && mode != EAOF_BOW_KF_KF) return mfail(EAOF_ERR_ARG, "unknown mode");
    const eaof_kp* kps; const uint8_t* desc; const int* counts; const float* scale;
    int cap, W, H, nlevels; void* exStream;
    int rc = eaof_internal_orb_view(ex, &kps, &desc, &counts, &cap, &W, &H, &scale, &nlevels, &exStream);
    if (rc) return rc;
    if (cap > m->maxFeat) return mfail(EAOF_ERR_ARG, "extractor keypoint capacity %d exceeds matcher max_features %d", cap, m->maxFeat);
    if (nFrames < 1 || (size_t)nFrames * cap > (size_t)(m->maxPairs + 1) * m->maxFeat)
        return mfail(EAOF_ERR_ARG, "n_frames %d does not fit a matcher of %d pairs", nFrames, m->maxPairs);
    for (int p = 0; p < nPairs; ++p)
        if (pairQ[p] < 0 || pairQ[p] >= nFrames || pairT[p] < 0 || pairT[p] >= nFrames) return mfail(EAOF_ERR_ARG, "pair %d names a frame outside [0,%d)", p, nFrames);
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    if (!m->segsBatch) {
        MCK(cudaMalloc(&m->segsBatch, sizeof(BowSeg) * (size_t)m->maxPairs * m->maxFeat));
        MCK(cudaMalloc(&m->segCountBatch, sizeof(int) * (size_t)m->maxPairs));
        MCK(cudaMalloc(&m->orbAngle, sizeof(float) * (size_t)(m->maxPairs + 1) * m->maxFeat));
    }
    if (m->upload_words(0, pairQ, nPairs, m->pairIdx) || m->upload_words(1, pairT, nPairs, m->pairIdx + m->maxPairs))
        return mfail(EAOF_ERR_CUDA, "pair list upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    MCK(cudaEventRecord(m->evDep, (cudaStream_t)exStream));
    MCK(cudaStreamWaitEvent(s, m->evDep, 0));
    k_kp_angles<<<(nFrames * cap + 255) / 256, 256, 0, s>>>(kps, nFrames * cap, m->orbAngle);
    BowArgs A{};
    A.desc = desc; A.angle = m->orbAngle; A.counts = counts; A.pairQ = m->pairIdx; A.pairT = m->pairIdx + m->maxPairs;
    A.blockStride = cap; A.stride = cap; A.mode = mode; A.segs = m->segsBatch; A.segCount = m->segCountBatch;
    A.featIdx = dFeatIdx; A.featStride = cap;
    A.thEff = mode == EAOF_BOW_KF_FRAME ? EAOF_TH_LOW : EAOF_TH_LOW - 1;
    A.D = near_threshold(A.thEff, ratio); A.ratio = ratio; A.checkOri = checkOri;
    k_bow_plan<<<nPairs, 32, 0, s>>>(A, dNFNodes, dNodeIds, dNodeStart, m->segsBatch, m->segCountBatch);
    k_bow_dense_nodes<<<dim3((cap + BOW_QT - 1) / BOW_QT, nPairs), BOW_QT, 0, s>>>(A, m->nearBuf);
    A.spec = cap <= 11000;
    const size_t bm = sizeof(uint32_t) * ((cap + 31) / 32 + (A.spec ? (size_t)cap : 0));
    k_bow_resolve<<<nPairs, 32, bm, s>>>(A, m->nearBuf, m->accBuf, dMatch, dDist, dN);
    MCK(cudaGetLastError());
    m->lastDistances = -1;
    return eaof_internal_orb_note_reader(ex, (void*)s);
}

static int run_projection(eaof_matcher* m, ProjArgs& A, int nPairs, int maxL, int* dMatch, int* dDist, int* dN) {
    cudaStream_t s = m->stream;
    k_build_grid<<<nPairs, 256, 0, s>>>(A, m->cellStart, m->cellIdx, m->cellPack);
    k_proj_dense<0><<<dim3((maxL * PROJ_LANES + 127) / 128, nPairs), 128, 0, s>>>(A, m->cellStart, m->cellPack, m->nearBuf);
    const size_t bm = sizeof(uint32_t) * ((A.stride + 31) / 32);
    k_proj_resolve<<<nPairs, 32, bm, s>>>(A, m->cellStart, m->cellPack, m->nearBuf, m->accBuf, dMatch, dDist, dN);
    MCK(cudaGetLastError());
    return EAOF_OK;
}

int eaof_match_projection(eaof_matcher* m, int nC, const float* cx, const float* cy, const int* coct, const float* cangle,
                          const uint8_t* cdesc, const float* curight, const uint8_t* ctaken, float minX, float maxX,
                          float minY, float maxY, float invW, float invH, int nL, const uint8_t* lvalid, const float* lu,
                          const float* lv, const float* linvz, const int* loct, const float* langle, const uint8_t* ldesc,
                          const uint8_t* lobs, const float* scaleFactors, int nLevels, float th, float mbf, int searchMode,
                          int checkOri, int* matchCur, int* distCur, int* nMatches) {
    if (!m || !matchCur || !nMatches || nC < 0 || nL < 0) return mfail(EAOF_ERR_ARG, "bad argument");
    if (nC > m->maxFeat || nL > m->maxFeat) return mfail(EAOF_ERR_ARG, "feature count exceeds max_features=%d", m->maxFeat);
    if (nLevels < 1 || nLevels > EAOF_MAX_LEVELS || !scaleFactors) return mfail(EAOF_ERR_ARG, "bad scale table");
    *nMatches = 0;
    for (int i = 0; i < nC; ++i) { matchCur[i] = -1; if (distCur) distCur[i] = -1; }
    if (nC == 0 || nL == 0) return EAOF_OK;
    if (!cx || !cy || !coct || !cangle || !cdesc || !lu || !lv || !loct || !langle || !ldesc) return mfail(EAOF_ERR_ARG, "null array");
    for (int i = 0; i < nL; ++i) if (loct[i] < 0 || loct[i] >= nLevels) return mfail(EAOF_ERR_ARG, "last_octave[%d] out of range", i);
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    // one pinned block, one upload
    m->upOff = 0;
    ProjArgs A{};
    A.cx = m->stage(cx, nC); A.cy = m->stage(cy, nC); A.coct = m->stage(coct, nC); A.cangle = m->stage(cangle, nC);
    A.curight = curight ? m->stage(curight, nC) : nullptr;
    A.ctaken = ctaken ? m->stage(ctaken, nC) : nullptr;
    A.lu = m->stage(lu, nL); A.lv = m->stage(lv, nL); A.loct = m->stage(loct, nL); A.langle = m->stage(langle, nL);
    A.linvz = linvz ? m->stage(linvz, nL) : nullptr;
    A.lvalid = lvalid ? m->stage(lvalid, nL) : nullptr;
    A.lobs = lobs ? m->stage(lobs, nL) : nullptr;
    A.desc = m->stage(cdesc, 32 * (size_t)nC);          // Cur rows, then Last rows right behind them (32*nC is a multiple of 16)
    m->stage(ldesc, 32 * (size_t)nL);
    const int hdr[4] = {nC, nL, 0, nC};                 // counts, first descriptor row of Cur / of Last
    const int* dHdr = m->stage(hdr, 4);
    A.nC = dHdr; A.nL = dHdr + 1; A.cRow = dHdr + 2; A.lRow = dHdr + 3;
    MCK(m->flush_stage());
    A.stride = m->maxFeat;
    A.minX = minX; A.maxX = maxX; A.minY = minY; A.maxY = maxY; A.invW = invW; A.invH = invH;
    for (int i = 0; i < nLevels; ++i) A.scale[i] = scaleFactors[i];
    A.th = th; A.mbf = mbf; A.searchMode = searchMode; A.checkOri = checkOri;
    A.thAccept = EAOF_TH_HIGH; A.cut = EAOF_TH_HIGH; A.histMode = 2; A.checkBounds = 1;
    int* dN = m->outArenaD;
    int* dM = m->outArenaD + 4;
    int* dD = dM + nC;
    int rc = run_projection(m, A, 1, nL, dM, dD, dN);
    if (rc) return rc;
    MCK(cudaMemcpyAsync(m->outArenaH, m->outArenaD, sizeof(int) * (4 + 2 * (size_t)nC), cudaMemcpyDeviceToHost, s));
    MCK(cudaStreamSynchronize(s));
    *nMatches = m->outArenaH[0];
    memcpy(matchCur, m->outArenaH + 4, sizeof(int) * nC);
    if (distCur) memcpy(distCur, m->outArenaH + 4 + nC, sizeof(int) * nC);
    return EAOF_OK;
}

int eaof_match_windows(eaof_matcher* m, int rule, int nT, const float* tx, const float* ty, const int* toct,
                       const float* tangle, const uint8_t* tdesc, const float* turight, const uint8_t* ttaken, float minX,
                       float maxX, float minY, float maxY, float invW, float invH, int nQ, const uint8_t* qValid,
                       const float* qU, const float* qV, const float* qRadius, const int* qMinLevel, const int* qMaxLevel,
                       const float* qUr, const float* qAngle, const uint8_t* qDesc, const uint8_t* qObs, int thAccept,
                       float nnratio, int histMode, int checkBounds, int* matchT, int* distT, int* nMatches) {
    if (!m || !matchT || !nMatches || nT < 0 || nQ < 0) return mfail(EAOF_ERR_ARG, "bad argument");
    if (rule != EAOF_WIN_BEST && rule != EAOF_WIN_RATIO_SAME_LEVEL) return mfail(EAOF_ERR_ARG, "unknown rule");
    if (nT > m->maxFeat || nQ > m->maxFeat) return mfail(EAOF_ERR_ARG, "feature count exceeds max_features=%d", m->maxFeat);
    if (histMode < 0 || histMode > 2 || thAccept < 0 || thAccept > 256) return mfail(EAOF_ERR_ARG, "bad hist_mode / th_accept");
    *nMatches = 0;
    for (int i = 0; i < nT; ++i) { matchT[i] = -1; if (distT) distT[i] = -1; }
    if (nT == 0 || nQ == 0) return EAOF_OK;
    if (!tx || !ty || !toct || !tdesc || !qU || !qV || !qRadius || !qMinLevel || !qMaxLevel || !qDesc) return mfail(EAOF_ERR_ARG, "null array");
    if (histMode && (!tangle || !qAngle)) return mfail(EAOF_ERR_ARG, "rotation check needs angles");
    if (turight && !qUr) return mfail(EAOF_ERR_ARG, "t_uright given without q_ur");
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    // one pinned block, one upload (see eaof_match_projection)
    m->upOff = 0;
    ProjArgs A{};
    A.cx = m->stage(tx, nT); A.cy = m->stage(ty, nT); A.coct = m->stage(toct, nT);
    A.cangle = tangle ? m->stage(tangle, nT) : m->cangle;
    A.curight = turight ? m->stage(turight, nT) : nullptr;
    A.ctaken = ttaken ? m->stage(ttaken, nT) : nullptr;
    A.lu = m->stage(qU, nQ); A.lv = m->stage(qV, nQ); A.qRadius = m->stage(qRadius, nQ);
    A.loct = m->stage(qMinLevel, nQ); A.qMinL = A.loct; A.qMaxL = m->stage(qMaxLevel, nQ);
    A.qUr = qUr ? m->stage(qUr, nQ) : A.lu;  // without a stereo prediction qUr is never read
    A.langle = qAngle ? m->stage(qAngle, nQ) : m->langle;
    A.lvalid = qValid ? m->stage(qValid, nQ) : nullptr;
    A.lobs = qObs ? m->stage(qObs, nQ) : nullptr;
    A.linvz = nullptr;
    A.desc = m->stage(tdesc, 32 * (size_t)nT);
    m->stage(qDesc, 32 * (size_t)nQ);
    const int hdr[4] = {nT, nQ, 0, nT};
    const int* dHdr = m->stage(hdr, 4);
    A.nC = dHdr; A.nL = dHdr + 1; A.cRow = dHdr + 2; A.lRow = dHdr + 3;
    MCK(m->flush_stage());
    A.stride = m->maxFeat;
    A.minX = minX; A.maxX = maxX; A.minY = minY; A.maxY = maxY; A.invW = invW; A.invH = invH;
    A.checkOri = histMode != 0; A.histMode = histMode; A.checkBounds = checkBounds;
    A.thAccept = thAccept; A.ratio = nnratio;
    if (rule == EAOF_WIN_BEST) {
        A.cut = thAccept;
    } else {
        int sMin = 0;  // smallest second-best that passes "best <= ratio*second" for every best <= thAccept
        while (sMin <= 256 && (float)thAccept > nnratio * (float)sMin) ++sMin;
        A.cut = std::min(256, std::max(sMin, thAccept + 1)) - 1;
        if (A.cut < thAccept) A.cut = thAccept;
    }
    if (!turight) A.curight = nullptr;
    k_build_grid<<<1, 256, 0, s>>>(A, m->cellStart, m->cellIdx, m->cellPack);
    k_proj_dense<0><<<dim3((nQ * PROJ_LANES + 127) / 128, 1), 128, 0, s>>>(A, m->cellStart, m->cellPack, m->nearBuf);
    const size_t bm = sizeof(uint32_t) * ((A.stride + 31) / 32);
    int* dN = m->outArenaD;
    int* dM = m->outArenaD + 4;
    int* dD = dM + nT;
    if (rule == EAOF_WIN_BEST)
        k_proj_resolve<<<1, 32, bm, s>>>(A, m->cellStart, m->cellPack, m->nearBuf, m->accBuf, dM, dD, dN);
    else
        k_win_resolve_ratio<<<1, 32, bm, s>>>(A, m->cellStart, m->cellIdx, m->nearBuf, dM, dD, dN);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(m->outArenaH, m->outArenaD, sizeof(int) * (4 + 2 * (size_t)nT), cudaMemcpyDeviceToHost, s));
    MCK(cudaStreamSynchronize(s));
    *nMatches = m->outArenaH[0];
    memcpy(matchT, m->outArenaH + 4, sizeof(int) * nT);
    if (distT) memcpy(distT, m->outArenaH + 4 + nT, sizeof(int) * nT);
    return EAOF_OK;
}

int eaof_match_windows_independent(eaof_matcher* m, int gate, int nT, const float* tx, const float* ty, const int* toct,
                                   const uint8_t* tdesc, const float* turight, float minX, float minY, float invW,
                                   float invH, const float* invLevelSigma2, int nLevels, int nQ, const uint8_t* qValid,
                                   const float* qU, const float* qV, const float* qRadius, const int* qMinLevel,
                                   const int* qMaxLevel, const float* qUr, const uint8_t* qDesc, int thAccept,
                                   int* matchQ, int* distQ, int* nMatches) {
    if (!m || !matchQ || !nMatches || nT < 0 || nQ < 0) return mfail(EAOF_ERR_ARG, "bad argument");
    if (gate != EAOF_GATE_NONE && gate != EAOF_GATE_FUSE_CHI2) return mfail(EAOF_ERR_ARG, "unknown gate");
    if (nT > m->maxFeat || nQ > m->maxFeat) return mfail(EAOF_ERR_ARG, "feature count exceeds max_features=%d", m->maxFeat);
    if (thAccept < 0 || thAccept > 256) return mfail(EAOF_ERR_ARG, "bad th_accept");
    *nMatches = 0;
    for (int i = 0; i < nQ; ++i) { matchQ[i] = -1; if (distQ) distQ[i] = -1; }
    if (nT == 0 || nQ == 0) return EAOF_OK;
    if (!tx || !ty || !toct || !tdesc || !qU || !qV || !qRadius || !qMinLevel || !qMaxLevel || !qDesc) return mfail(EAOF_ERR_ARG, "null array");
    if (gate == EAOF_GATE_FUSE_CHI2) {
        if (!invLevelSigma2 || !qUr || nLevels < 1 || nLevels > EAOF_MAX_LEVELS) return mfail(EAOF_ERR_ARG, "chi-square gate needs inv_level_sigma2 and q_ur");
        for (int i = 0; i < nT; ++i) if (toct[i] < 0 || toct[i] >= nLevels) return mfail(EAOF_ERR_ARG, "t_octave[%d] out of range", i);
    }
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
#define UP(dst, src, n, T) MCK(cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)(n), cudaMemcpyHostToDevice, s))
    UP(m->cx, tx, nT, float); UP(m->cy, ty, nT, float); UP(m->coct, toct, nT, int);
    const bool useRight = gate == EAOF_GATE_FUSE_CHI2 && turight;
    if (useRight) UP(m->curight, turight, nT, float);
    UP(m->lu, qU, nQ, float); UP(m->lv, qV, nQ, float); UP(m->qRadius, qRadius, nQ, float);
    UP(m->loct, qMinLevel, nQ, int); UP(m->qMaxL, qMaxLevel, nQ, int);
    if (qUr) UP(m->linvz, qUr, nQ, float);
    if (qValid) UP(m->lvalid, qValid, nQ, uint8_t);
    UP(m->desc2, tdesc, 32 * (size_t)nT, uint8_t);
    UP(m->desc2 + 32 * (size_t)m->maxFeat, qDesc, 32 * (size_t)nQ, uint8_t);
    const int hdr[4] = {nT, nQ, 0, m->maxFeat};
    UP(m->nC, &hdr[0], 1, int); UP(m->nL, &hdr[1], 1, int); UP(m->cRow, &hdr[2], 1, int); UP(m->lRow, &hdr[3], 1, int);
#undef UP
    ProjArgs A{};
    A.cx = m->cx; A.cy = m->cy; A.coct = m->coct; A.curight = useRight ? m->curight : nullptr; A.nC = m->nC;
    A.lu = m->lu; A.lv = m->lv; A.loct = m->loct; A.lvalid = qValid ? m->lvalid : nullptr; A.nL = m->nL;
    A.desc = m->desc2; A.cRow = m->cRow; A.lRow = m->lRow; A.stride = m->maxFeat;
    A.minX = minX; A.minY = minY; A.invW = invW; A.invH = invH;
    A.qRadius = m->qRadius; A.qMinL = m->loct; A.qMaxL = m->qMaxL; A.qUr = qUr ? m->linvz : m->lu;
    A.thAccept = thAccept; A.cut = thAccept;
    if (gate == EAOF_GATE_FUSE_CHI2) for (int i = 0; i < nLevels; ++i) A.invSigma2[i] = invLevelSigma2[i];
    MCK(cudaMemsetAsync(m->outN, 0, sizeof(int), s));
    k_build_grid<<<1, 256, 0, s>>>(A, m->cellStart, m->cellIdx, m->cellPack);
    if (gate == EAOF_GATE_FUSE_CHI2)
        k_proj_dense<1><<<dim3((nQ * PROJ_LANES + 127) / 128, 1), 128, 0, s>>>(A, m->cellStart, m->cellPack, m->nearBuf);
    else
        k_proj_dense<0><<<dim3((nQ * PROJ_LANES + 127) / 128, 1), 128, 0, s>>>(A, m->cellStart, m->cellPack, m->nearBuf);
    k_win_pick<<<(nQ + 127) / 128, 128, 0, s>>>(nQ, thAccept, m->nearBuf, m->outMatch, m->outDist, m->outN);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(matchQ, m->outMatch, sizeof(int) * nQ, cudaMemcpyDeviceToHost, s));
    if (distQ) MCK(cudaMemcpyAsync(distQ, m->outDist, sizeof(int) * nQ, cudaMemcpyDeviceToHost, s));
    MCK(cudaMemcpyAsync(nMatches, m->outN, sizeof(int), cudaMemcpyDeviceToHost, s));
    MCK(cudaStreamSynchronize(s));
    return EAOF_OK;
}

int eaof_distinctive_descriptors(eaof_matcher* m, int nPoints, const int* start, const uint8_t* desc, int* bestOut,
                                 int* medianOut) {
    if (!m || nPoints < 0 || !bestOut) return mfail(EAOF_ERR_ARG, "bad argument");
    if (nPoints == 0) return EAOF_OK;
    if (!start || !desc) return mfail(EAOF_ERR_ARG, "null array");
    const int rowCap = 2 * m->maxFeat, ptCap = m->maxFeat - 1;  // staging: desc2 holds 2*maxFeat rows, idxQ maxFeat starts
    for (int p = 0; p < nPoints; ++p) {
        if (start[p + 1] < start[p]) return mfail(EAOF_ERR_ARG, "start[] must be non-decreasing");
        if (start[p + 1] - start[p] > rowCap) return mfail(EAOF_ERR_ARG, "map point %d has more than %d observations", p, rowCap);
    }
    if (ptCap < 1) return mfail(EAOF_ERR_ARG, "matcher too small");
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    std::vector<int> rel;
    for (int p0 = 0; p0 < nPoints;) {
        int p1 = p0;  // largest chunk of points whose rows fit the staging buffers
        while (p1 < nPoints && p1 - p0 < ptCap && start[p1 + 1] - start[p0] <= rowCap) ++p1;
        const int np = p1 - p0, rows = start[p1] - start[p0];
        rel.resize(np + 1);
        for (int i = 0; i <= np; ++i) rel[i] = start[p0 + i] - start[p0];
        MCK(cudaMemcpyAsync(m->idxQ, rel.data(), sizeof(int) * (np + 1), cudaMemcpyHostToDevice, s));
        if (rows) MCK(cudaMemcpyAsync(m->desc2, desc + 32 * (size_t)start[p0], 32 * (size_t)rows, cudaMemcpyHostToDevice, s));
        k_distinctive<<<np, 128, 0, s>>>(m->idxQ, m->desc2, m->outMatch, m->outDist);
        MCK(cudaGetLastError());
        MCK(cudaMemcpyAsync(bestOut + p0, m->outMatch, sizeof(int) * np, cudaMemcpyDeviceToHost, s));
        if (medianOut) MCK(cudaMemcpyAsync(medianOut + p0, m->outDist, sizeof(int) * np, cudaMemcpyDeviceToHost, s));
        MCK(cudaStreamSynchronize(s));  // rel / staging are reused by the next chunk
        p0 = p1;
    }
    return EAOF_OK;
}

int eaof_match_initialization(eaof_matcher* m, float nnratio, int checkOri, int n1, const int* oct1, const float* angle1,
                              const uint8_t* desc1, float* prevMatched, int n2, const float* x2, const float* y2,
                              const int* oct2, const float* angle2, const uint8_t* desc2, float minX, float maxX, float minY,
                              float maxY, float invW, float invH, int windowSize, int* matches12, int* nMatches) {
    if (!m || !matches12 || !nMatches || n1 < 0 || n2 < 0) return mfail(EAOF_ERR_ARG, "bad argument");
    if (n1 > m->maxFeat || n2 > m->maxFeat) return mfail(EAOF_ERR_ARG, "feature count exceeds max_features=%d", m->maxFeat);
    *nMatches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return EAOF_OK;
    if (!oct1 || !angle1 || !desc1 || !prevMatched || !x2 || !y2 || !oct2 || !angle2 || !desc2) return mfail(EAOF_ERR_ARG, "null array");
    (void)maxX; (void)maxY;
    std::vector<float> qu(n1), qv(n1), qr(n1, (float)windowSize);
    std::vector<int> ql(n1);
    std::vector<uint8_t> qval(n1);
    for (int i = 0; i < n1; ++i) {
        qu[i] = prevMatched[2 * i]; qv[i] = prevMatched[2 * i + 1];
        ql[i] = oct1[i];                 // GetFeaturesInArea(x, y, windowSize, level1, level1), :425
        qval[i] = oct1[i] > 0 ? 0 : 1;   // only level-0 features of F1 are matched (:421-423)
    }
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
#define UP(dst, src, n, T) MCK(cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)(n), cudaMemcpyHostToDevice, s))
    UP(m->cx, x2, n2, float); UP(m->cy, y2, n2, float); UP(m->coct, oct2, n2, int); UP(m->cangle, angle2, n2, float);
    UP(m->lu, qu.data(), n1, float); UP(m->lv, qv.data(), n1, float); UP(m->qRadius, qr.data(), n1, float);
    UP(m->loct, ql.data(), n1, int); UP(m->qMaxL, ql.data(), n1, int); UP(m->langle, angle1, n1, float);
    UP(m->lvalid, qval.data(), n1, uint8_t);
    UP(m->desc2, desc2, 32 * (size_t)n2, uint8_t);
    UP(m->desc2 + 32 * (size_t)m->maxFeat, desc1, 32 * (size_t)n1, uint8_t);
    const int hdr[4] = {n2, n1, 0, m->maxFeat};
    UP(m->nC, &hdr[0], 1, int); UP(m->nL, &hdr[1], 1, int); UP(m->cRow, &hdr[2], 1, int); UP(m->lRow, &hdr[3], 1, int);
#undef UP
    MCK(cudaStreamSynchronize(s));  // the staging vectors above live on this stack frame
    ProjArgs A{};
    A.cx = m->cx; A.cy = m->cy; A.coct = m->coct; A.cangle = m->cangle; A.nC = m->nC; A.lu = m->lu; A.lv = m->lv;
    A.loct = m->loct; A.langle = m->langle; A.lvalid = m->lvalid; A.nL = m->nL; A.desc = m->desc2; A.cRow = m->cRow;
    A.lRow = m->lRow; A.stride = m->maxFeat; A.minX = minX; A.maxX = maxX; A.minY = minY; A.maxY = maxY; A.invW = invW;
    A.invH = invH; A.qRadius = m->qRadius; A.qMinL = m->loct; A.qMaxL = m->qMaxL; A.qUr = m->lu; A.checkOri = checkOri;
    A.histMode = 1; A.checkBounds = 0; A.thAccept = EAOF_TH_LOW; A.ratio = nnratio;
    A.cut = near_threshold(EAOF_TH_LOW, nnratio) - 1;
    if (A.cut > 256) A.cut = 256;
    k_build_grid<<<1, 256, 0, s>>>(A, m->cellStart, m->cellIdx, m->cellPack);
    k_proj_dense<0><<<dim3((n1 * PROJ_LANES + 127) / 128, 1), 128, 0, s>>>(A, m->cellStart, m->cellPack, m->nearBuf);
    k_init_resolve<<<1, 32, 0, s>>>(A, m->cellStart, m->cellIdx, m->nearBuf, m->idxQ, m->idxT, m->initBin, m->outMatch, m->outN);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(matches12, m->outMatch, sizeof(int) * n1, cudaMemcpyDeviceToHost, s));
    MCK(cudaMemcpyAsync(nMatches, m->outN, sizeof(int), cudaMemcpyDeviceToHost, s));
    MCK(cudaStreamSynchronize(s));
    for (int i = 0; i < n1; ++i)  // :515-517 update prev matched
        if (matches12[i] >= 0) { prevMatched[2 * i] = x2[matches12[i]]; prevMatched[2 * i + 1] = y2[matches12[i]]; }
    return EAOF_OK;
}

// accessors implemented in eaof_orb.cu
int eaof_internal_orb_view(eaof_orb* ex, const eaof_kp** kps, const uint8_t** desc, const int** counts, int* cap, int* w,
                           int* h, const float** scale, int* nlevels, void** stream);
int eaof_internal_orb_note_reader(eaof_orb* ex, void* readerStream);

int eaof_match_projection_batch_device(eaof_matcher* m, eaof_orb* ex, int nPairs, const int* lastFrame, const int* curFrame,
                                       const float* shiftX, const float* shiftY, float th, int* dMatch, int* dDist, int* dN) {
    if (!m || !ex || !lastFrame || !curFrame || !shiftX || !shiftY || !dMatch || !dN) return mfail(EAOF_ERR_ARG, "null argument");
    if (nPairs < 1 || nPairs > m->maxPairs) return mfail(EAOF_ERR_ARG, "n_pairs %d outside [1,%d]", nPairs, m->maxPairs);
    const eaof_kp* kps; const uint8_t* desc; const int* counts; const float* scale;
    int cap, W, H, nlevels; void* exStream;
    int rc = eaof_internal_orb_view(ex, &kps, &desc, &counts, &cap, &W, &H, &scale, &nlevels, &exStream);
    if (rc) return rc;
    if (cap > m->maxFeat) return mfail(EAOF_ERR_ARG, "extractor keypoint capacity %d exceeds matcher max_features %d", cap, m->maxFeat);
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    int* dLast = m->pairIdx; int* dCur = m->pairIdx + m->maxPairs;
    float* dSx = m->pairShift; float* dSy = m->pairShift + m->maxPairs;
    if (m->upload_words(0, lastFrame, nPairs, dLast) || m->upload_words(1, curFrame, nPairs, dCur) ||
        m->upload_words(2, shiftX, nPairs, dSx) || m->upload_words(3, shiftY, nPairs, dSy))
        return mfail(EAOF_ERR_CUDA, "pair list upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    // order after the extraction that produced the keypoints
    MCK(cudaEventRecord(m->evDep, (cudaStream_t)exStream));
    MCK(cudaStreamWaitEvent(s, m->evDep, 0));
    const int stride = cap;
    k_proj_prepare<<<dim3((cap + 127) / 128, nPairs), 128, 0, s>>>(kps, counts, cap, dLast, dCur, dSx, dSy, stride, m->cx, m->cy,
                                                                  m->coct, m->cangle, m->lu, m->lv, m->loct, m->langle, m->nC,
                                                                  m->nL, m->cRow, m->lRow);
    ProjArgs A{};
    A.cx = m->cx; A.cy = m->cy; A.coct = m->coct; A.cangle = m->cangle; A.nC = m->nC; A.lu = m->lu; A.lv = m->lv;
    A.loct = m->loct; A.langle = m->langle; A.nL = m->nL; A.desc = desc; A.cRow = m->cRow; A.lRow = m->lRow; A.stride = stride;
    // Frame image bounds without distortion (src/Frame.cc:1040-1046): mnMinX=0, mnMaxX=cols, grid cell inverse sizes :213-214
    A.minX = 0.f; A.maxX = (float)W; A.minY = 0.f; A.maxY = (float)H;
    A.invW = (float)GRID_COLS / (A.maxX - A.minX); A.invH = (float)GRID_ROWS / (A.maxY - A.minY);
    for (int i = 0; i < nlevels; ++i) A.scale[i] = scale[i];
    A.th = th; A.mbf = 0.f; A.searchMode = 0; A.checkOri = 1;
    A.thAccept = EAOF_TH_HIGH; A.cut = EAOF_TH_HIGH; A.histMode = 2; A.checkBounds = 1;
    rc = run_projection(m, A, nPairs, cap, dMatch, dDist, dN);
    if (rc) return rc;
    return eaof_internal_orb_note_reader(ex, (void*)s);  // the extractor must not overwrite what these kernels still read
}

}  // extern "C"
