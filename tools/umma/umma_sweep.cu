// Stand-alone check + timing of eao-fusion_b200/csrc/bow_umma.cuh (tcgen05 int8 Hamming, phase 1 of the brute-force matcher)
// against a CPU popcount reference:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../eao-fusion_b200/csrc -o umma_sweep umma_sweep.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bow_umma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
using namespace eaof_umma;

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int stride = 2000, nBlocks = 24, D = 56;
    const int nPairsTime = argc > 1 ? atoi(argv[1]) : 1024;
    std::vector<int> counts(nBlocks, 2000);
    counts[3] = 1999; counts[5] = 130; counts[7] = 1; counts[9] = 0; counts[11] = 257; counts[13] = 128;
    std::vector<uint8_t> desc((size_t)nBlocks * stride * 32);
    srand(11);
    for (auto& b : desc) b = (uint8_t)(rand() & 0xff);
    // plant near-duplicates: descriptor i of block b+1 = descriptor (i*7 % n) of block b with a few flipped bits
    for (int b = 0; b + 1 < nBlocks; ++b)
        for (int i = 0; i < counts[b + 1]; ++i) {
            if (counts[b] == 0 || (i % 3) == 2) continue;
            const int j = (i * 7) % counts[b];
            uint8_t* dst = &desc[((size_t)(b + 1) * stride + i) * 32];
            memcpy(dst, &desc[((size_t)b * stride + j) * 32], 32);
            const int flips = rand() % 70;
            for (int f = 0; f < flips; ++f) { const int bit = rand() & 255; dst[bit >> 3] ^= (uint8_t)(1u << (bit & 7)); }
        }
    std::vector<int> pq, pt;
    for (int b = 0; b + 1 < nBlocks; ++b) { pq.push_back(b + 1); pt.push_back(b); }
    for (int b = 0; b + 2 < nBlocks; b += 3) { pq.push_back(b); pt.push_back(b + 2); }
    const int nPairs = (int)pq.size();

    uint8_t* dDesc; int8_t* dExp; int *dCounts, *dPq, *dPt; uint32_t* dNear;
    const size_t nDesc = (size_t)nBlocks * stride;
    CK(cudaMalloc(&dDesc, desc.size())); CK(cudaMalloc(&dExp, nDesc * 256)); CK(cudaMalloc(&dCounts, sizeof(int) * nBlocks));
    const int maxPairs = std::max(nPairs, nPairsTime);
    CK(cudaMalloc(&dPq, sizeof(int) * maxPairs)); CK(cudaMalloc(&dPt, sizeof(int) * maxPairs));
    CK(cudaMalloc(&dNear, (size_t)maxPairs * stride * 32));
    CK(cudaMemcpy(dDesc, desc.data(), desc.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dCounts, counts.data(), sizeof(int) * nBlocks, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dPq, pq.data(), sizeof(int) * nPairs, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dPt, pt.data(), sizeof(int) * nPairs, cudaMemcpyHostToDevice));
    CK(cudaMemset(dNear, 0xee, (size_t)maxPairs * stride * 32));

    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    TMap mA{}, mB{};
    auto enc = [&](TMap& m, int boxRows) {
        const cuuint64_t dims[2] = {256, (cuuint64_t)nDesc};
        const cuuint64_t strides[1] = {256};
        const cuuint32_t box[2] = {128, (cuuint32_t)boxRows}, es[2] = {1, 1};
        CUresult r = ((EncodeTiled)fn)((CUtensorMap*)&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dExp, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    };
    enc(mA, kTileQ);
    enc(mB, kTileT);
    k_expand_pm1<<<(unsigned)((nDesc * 16 + 255) / 256), 256>>>(dDesc, dExp, nDesc);
    CK(cudaGetLastError());
    CK(cudaFuncSetAttribute(k_bow_dense_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    Args U{dPq, dPt, dCounts, stride, nPairs, (stride + kTileQ - 1) / kTileQ, D};
    k_bow_dense_umma<<<sms, kThreads, kSmemBytes>>>(mA, mB, U, dNear);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> near((size_t)nPairs * stride * 8);
    CK(cudaMemcpy(near.data(), dNear, near.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0, cand = 0;
    for (int p = 0; p < nPairs; ++p) {
        const int bq = pq[p], bt = pt[p];
        for (int q = 0; q < counts[bq]; ++q) {
            uint32_t ref[8];
            for (int i = 0; i < 7; ++i) ref[i] = 0xffffffffu;
            int cnt = 0;
            const uint64_t* a = reinterpret_cast<const uint64_t*>(&desc[((size_t)bq * stride + q) * 32]);
            for (int t = 0; t < counts[bt]; ++t) {
                const uint64_t* b = reinterpret_cast<const uint64_t*>(&desc[((size_t)bt * stride + t) * 32]);
                const int d = __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) + __builtin_popcountll(a[2] ^ b[2]) +
                              __builtin_popcountll(a[3] ^ b[3]);
                if (d < D) { if (cnt < 7) ref[cnt] = (uint32_t)t | ((uint32_t)d << 16); ++cnt; }
            }
            ref[7] = (uint32_t)cnt;
            cand += cnt;
            const uint32_t* g = &near[((size_t)p * stride + q) * 8];
            if (memcmp(g, ref, 32) != 0) {
                if (bad < 5) printf("pair %d (q block %d, t block %d) query %d: got cnt %u first %08x, want cnt %u first %08x\n", p, bq, bt, q, g[7], g[0], ref[7], ref[0]);
                ++bad;
            }
        }
    }
    printf("UMMA_SWEEP check %s: %ld mismatching queries, %ld candidates below D over %d pairs\n", bad ? "FAIL" : "OK", bad, cand, nPairs);
    if (bad) return 2;
    // timing: nPairsTime pairs of full blocks
    std::vector<int> tq(nPairsTime), tt(nPairsTime);
    const int full[] = {0, 1, 2, 4, 6, 8, 10, 12, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23};
    for (int i = 0; i < nPairsTime; ++i) { tq[i] = full[i % 18]; tt[i] = full[(i * 5 + 1) % 18]; }
    CK(cudaMemcpy(dPq, tq.data(), sizeof(int) * nPairsTime, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dPt, tt.data(), sizeof(int) * nPairsTime, cudaMemcpyHostToDevice));
    U.nPairs = nPairsTime;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_bow_dense_umma<<<sms, kThreads, kSmemBytes>>>(mA, mB, U, dNear);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 3; ++r) k_bow_dense_umma<<<sms, kThreads, kSmemBytes>>>(mA, mB, U, dNear);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= 3;
    printf("UMMA_SWEEP timing: %d pairs of 2000 x 2000 in %.3f ms = %.3f T distances/s, %.0f pairs/s\n", nPairsTime, ms,
           (double)nPairsTime * 4e6 / (ms * 1e-3) / 1e12, nPairsTime / (ms * 1e-3));
    return 0;
}
