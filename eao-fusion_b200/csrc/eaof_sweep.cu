// libeaof_orb.so — the exchange step of the cross-frame Hamming sweep (include/eaof_sweep.h): ncclAllGather of the
// descriptor / angle / count blocks every rank extracted, then the brute-force SearchByBoW kernels of eaof_match.cu over
// this rank's share of the pair list.  The reference has no counterpart (single process); the semantics
// per pair are src/ORBmatcher.cc:159-288 with one vocabulary node holding every feature.
//
// NCCL is bound with dlopen at first use so that the library loads (and every other entry point works) where NCCL is
// absent; a process that already loaded a libnccl.so.2 (PyTorch ships one) gets that same copy.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; the symbols are resolved by dlsym

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/eaof_sweep.h"

extern "C" int eaof_internal_fail(int code, const char* msg);
extern "C" int eaof_internal_bruteforce_pairs_device(eaof_matcher* m, int mode, float ratio, int checkOri, int nPairs,
                                                     const int* dPairQ, const int* dPairT, const uint8_t* dDesc,
                                                     const float* dAngle, const int* dCounts, int blockStride, int* dMatch,
                                                     int* dDist, int* dN);
extern "C" int eaof_internal_matcher_limits(const eaof_matcher* m, int* maxPairs, int* maxFeat, int* device);
extern "C" int eaof_internal_bruteforce_prepare(eaof_matcher* m, const uint8_t* dDesc, int nBlocks, int blockStride);
extern "C" int eaof_internal_bruteforce_release(eaof_matcher* m);

namespace {

int sfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return eaof_internal_fail(code, buf);
}
#define SCK(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess) return sfail(EAOF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetLastError) GetLastError = nullptr;
};
NcclApi g_nccl;
std::mutex g_ncclMu;

int bind_nccl() {
    std::lock_guard<std::mutex> lk(g_ncclMu);
    if (g_nccl.lib) return EAOF_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return sfail(EAOF_ERR_NCCL, "libnccl.so.2 not found (%s): the multi-GPU sweep needs NCCL, there is no other transport", dlerror());
    NcclApi a;
    a.lib = h;
#define BIND(name)                                                                           \
    a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, "nccl" #name));                     \
    if (!a.name) return sfail(EAOF_ERR_NCCL, "libnccl has no symbol nccl" #name)
    BIND(GetVersion); BIND(GetUniqueId); BIND(CommInitRank); BIND(CommDestroy); BIND(AllGather); BIND(GroupStart);
    BIND(GroupEnd); BIND(GetErrorString); BIND(GetLastError);
#undef BIND
    g_nccl = a;
    return EAOF_OK;
}

int nfail(ncclResult_t r, ncclComm_t comm, const char* what) {
    const char* last = comm && g_nccl.GetLastError ? g_nccl.GetLastError(comm) : "";
    return sfail(EAOF_ERR_NCCL, "%s failed: %s%s%s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?",
                 last && *last ? " — " : "", last ? last : "");
}
#define NCK(call, comm)                                       \
    do {                                                      \
        ncclResult_t r_ = (call);                             \
        if (r_ != ncclSuccess) return nfail(r_, comm, #call); \
    } while (0)

}  // namespace

struct eaof_sweep {
    int rank = 0, world = 1, device = 0;
    ncclComm_t comm = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    long long lastBytes = 0;
    int *dPairs = nullptr, *hPairs = nullptr;  // [2][pairCap] pair list of the last eaof_sweep_match
    int pairCap = 0;
    cudaEvent_t evPairs = nullptr;             // the upload out of hPairs has completed
};

extern "C" {

int eaof_sweep_unique_id(uint8_t id[EAOF_SWEEP_ID_BYTES]) {
    if (!id) return sfail(EAOF_ERR_ARG, "null argument");
    int rc = bind_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == EAOF_SWEEP_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NCK(g_nccl.GetUniqueId(&u), nullptr);
    memcpy(id, &u, sizeof u);
    return EAOF_OK;
}

int eaof_sweep_nccl_version(int* version) {
    if (!version) return sfail(EAOF_ERR_ARG, "null argument");
    int rc = bind_nccl();
    if (rc) return rc;
    NCK(g_nccl.GetVersion(version), nullptr);
    return EAOF_OK;
}

int eaof_sweep_create(const uint8_t id[EAOF_SWEEP_ID_BYTES], int rank, int world, int device, eaof_sweep** out) {
    if (!out) return sfail(EAOF_ERR_ARG, "null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return sfail(EAOF_ERR_ARG, "rank %d outside a world of %d", rank, world);
    if (world > 1 && !id) return sfail(EAOF_ERR_ARG, "a world of %d ranks needs the unique id of eaof_sweep_unique_id", world);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return sfail(EAOF_ERR_CUDA, "no CUDA device: libeaof_orb has no CPU fallback");
    if (device < 0 || device >= ndev) return sfail(EAOF_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    SCK(cudaSetDevice(device));
    eaof_sweep* s = new eaof_sweep;
    s->rank = rank; s->world = world; s->device = device;
    if (world > 1) {
        int rc = bind_nccl();
        if (rc) { delete s; return rc; }
        ncclUniqueId u;
        memcpy(&u, id, sizeof u);
        ncclResult_t r = g_nccl.CommInitRank(&s->comm, world, u, rank);
        if (r != ncclSuccess) { delete s; return nfail(r, nullptr, "ncclCommInitRank"); }
    }
    if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->evPairs, cudaEventDisableTiming) != cudaSuccess) {
        eaof_sweep_destroy(s);
        return sfail(EAOF_ERR_CUDA, "cudaEventCreate failed");
    }
    *out = s;
    return EAOF_OK;
}

void eaof_sweep_destroy(eaof_sweep* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->evPairs) cudaEventDestroy(s->evPairs);
    cudaFree(s->dPairs);
    cudaFreeHost(s->hPairs);
    delete s;
}

int eaof_sweep_rank(const eaof_sweep* s) { return s ? s->rank : sfail(EAOF_ERR_ARG, "null handle"); }
int eaof_sweep_world(const eaof_sweep* s) { return s ? s->world : sfail(EAOF_ERR_ARG, "null handle"); }

int eaof_sweep_allgather_blocks(eaof_sweep* s, int blocksPerRank, int blockStride, const uint8_t* dDescLocal,
                                const float* dAngleLocal, const int* dCountLocal, uint8_t* dDescAll, float* dAngleAll,
                                int* dCountAll, void* stream) {
    if (!s || !dDescLocal || !dAngleLocal || !dCountLocal || !dDescAll || !dAngleAll || !dCountAll)
        return sfail(EAOF_ERR_ARG, "null argument");
    if (blocksPerRank < 1 || blockStride < 1) return sfail(EAOF_ERR_ARG, "blocks_per_rank and block_stride must be positive");
    SCK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nDesc = (size_t)blocksPerRank * blockStride * 32, nAng = (size_t)blocksPerRank * blockStride,
                 nCnt = (size_t)blocksPerRank;
    SCK(cudaEventRecord(s->ev0, st));
    if (s->world == 1) {
        if (dDescAll != dDescLocal) SCK(cudaMemcpyAsync(dDescAll, dDescLocal, nDesc, cudaMemcpyDeviceToDevice, st));
        if (dAngleAll != dAngleLocal) SCK(cudaMemcpyAsync(dAngleAll, dAngleLocal, nAng * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (dCountAll != dCountLocal) SCK(cudaMemcpyAsync(dCountAll, dCountLocal, nCnt * sizeof(int), cudaMemcpyDeviceToDevice, st));
        s->lastBytes = 0;
    } else {
        NCK(g_nccl.GroupStart(), s->comm);
        NCK(g_nccl.AllGather(dDescLocal, dDescAll, nDesc, ncclUint8, s->comm, st), s->comm);
        NCK(g_nccl.AllGather(dAngleLocal, dAngleAll, nAng, ncclFloat32, s->comm, st), s->comm);
        NCK(g_nccl.AllGather(dCountLocal, dCountAll, nCnt, ncclInt32, s->comm, st), s->comm);
        NCK(g_nccl.GroupEnd(), s->comm);
        s->lastBytes = (long long)(s->world - 1) * (long long)(nDesc + nAng * sizeof(float) + nCnt * sizeof(int));
    }
    SCK(cudaEventRecord(s->ev1, st));
    s->timed = true;
    return EAOF_OK;
}

int eaof_sweep_match(eaof_sweep* s, eaof_matcher* m, int mode, float nnratio, int checkOri, int blocksPerRank,
                     int blockStride, const uint8_t* dDescLocal, const float* dAngleLocal, const int* dCountLocal,
                     uint8_t* dDescAll, float* dAngleAll, int* dCountAll, int nPairs, const int* pairQ, const int* pairT,
                     int* dMatch, int* dDist, int* dN) {
    if (!s || !m || !dMatch || !dDist || !dN) return sfail(EAOF_ERR_ARG, "null argument");
    if (nPairs < 0 || (nPairs > 0 && (!pairQ || !pairT))) return sfail(EAOF_ERR_ARG, "bad pair list");
    int maxPairs = 0, maxFeat = 0, mdev = 0;
    int rc = eaof_internal_matcher_limits(m, &maxPairs, &maxFeat, &mdev);
    if (rc) return rc;
    if (mdev != s->device) return sfail(EAOF_ERR_ARG, "matcher lives on device %d, the sweep handle on %d", mdev, s->device);
    const int nBlocks = s->world * blocksPerRank;
    cudaStream_t st = (cudaStream_t)eaof_matcher_stream(m);
    // every rank takes part in the collective even when its share of the pair list is empty — or wrong: the pair list is
    // this rank's own data and is checked after the exchange, so that one rank's bad list cannot leave the others waiting
    rc = eaof_sweep_allgather_blocks(s, blocksPerRank, blockStride, dDescLocal, dAngleLocal, dCountLocal, dDescAll, dAngleAll,
                                     dCountAll, st);
    if (rc || nPairs == 0) return rc;
    for (int p = 0; p < nPairs; ++p)
        if (pairQ[p] < 0 || pairQ[p] >= nBlocks || pairT[p] < 0 || pairT[p] >= nBlocks)
            return sfail(EAOF_ERR_ARG, "pair %d names a block outside [0,%d)", p, nBlocks);
    if (nPairs > s->pairCap) {
        SCK(cudaStreamSynchronize(st));
        cudaFree(s->dPairs); cudaFreeHost(s->hPairs);
        s->dPairs = nullptr; s->hPairs = nullptr; s->pairCap = 0;
        SCK(cudaMalloc(&s->dPairs, sizeof(int) * 2 * (size_t)nPairs));
        SCK(cudaMallocHost(&s->hPairs, sizeof(int) * 2 * (size_t)nPairs));
        s->pairCap = nPairs;
    } else {
        SCK(cudaEventSynchronize(s->evPairs));  // the previous call's upload has left the staging buffer
    }
    memcpy(s->hPairs, pairQ, sizeof(int) * (size_t)nPairs);
    memcpy(s->hPairs + s->pairCap, pairT, sizeof(int) * (size_t)nPairs);
    SCK(cudaMemcpyAsync(s->dPairs, s->hPairs, sizeof(int) * (size_t)nPairs, cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(s->dPairs + s->pairCap, s->hPairs + s->pairCap, sizeof(int) * (size_t)nPairs, cudaMemcpyHostToDevice, st));
    SCK(cudaEventRecord(s->evPairs, st));
    // the gathered descriptors are expanded once for the tensor-core distance kernel, every chunk of pairs reads the expansion
    rc = eaof_internal_bruteforce_prepare(m, dDescAll, nBlocks, blockStride);
    if (rc) return rc;
    for (int p0 = 0; p0 < nPairs; p0 += maxPairs) {
        const int n = nPairs - p0 < maxPairs ? nPairs - p0 : maxPairs;
        rc = eaof_internal_bruteforce_pairs_device(m, mode, nnratio, checkOri, n, s->dPairs + p0, s->dPairs + s->pairCap + p0,
                                                   dDescAll, dAngleAll, dCountAll, blockStride,
                                                   dMatch + (size_t)p0 * blockStride, dDist + (size_t)p0 * blockStride, dN + p0);
        if (rc) break;
    }
    eaof_internal_bruteforce_release(m);
    return rc;
}

int eaof_sweep_last_allgather(eaof_sweep* s, float* ms, long long* bytes) {
    if (!s) return sfail(EAOF_ERR_ARG, "null handle");
    if (!s->timed) return sfail(EAOF_ERR_ARG, "no all-gather has run on this handle");
    if (ms) SCK(cudaEventElapsedTime(ms, s->ev0, s->ev1));
    if (bytes) *bytes = s->lastBytes;
    return EAOF_OK;
}

}  // extern "C"
