// Probe: one tcgen05.mma kind::i8 tile (M=128 queries, N=256 targets, K=256) on +-1 int8 operands staged by TMA (SWIZZLE_128B),
// accumulators in TMEM, read back with tcgen05.ld.  Hamming(a, b) = (256 - dot(a', b')) / 2 for a' = 2a - 1.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && timeout 60 ./umma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
// K-major operand, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);   // start address
    d |= (uint64_t)1 << 16;                       // leading byte offset (ignored for swizzled K-major), encoded >> 4
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                       // layout type: SWIZZLE_128B
    return d;
}

struct alignas(64) TMap { unsigned long long v[16]; };

__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ TMap mapA, const __grid_constant__ TMap mapB, int* __restrict__ C) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ uint32_t tmemBase;
    uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    uint8_t* sA = base;                 // 2 k-blocks x [128 rows x 128 B] = 32 KB
    uint8_t* sB = base + 2 * 16384;     // 2 k-blocks x [256 rows x 128 B] = 64 KB
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t barLoad = smem_u32(&bars[0]), barMma = smem_u32(&bars[1]);
    if (threadIdx.x == 0) {
        mbar_init(barLoad, 1);
        mbar_init(barMma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmemBase;
    if (threadIdx.x == 0) {
        mbar_expect_tx(barLoad, 2 * 16384 + 2 * 32768);
        for (int kb = 0; kb < 2; ++kb) {
            tma_load_2d(smem_u32(sA + kb * 16384), &mapA, kb * 128, 0, barLoad);
            tma_load_2d(smem_u32(sB + kb * 32768), &mapB, kb * 128, 0, barLoad);
        }
        mbar_wait(barLoad, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        // instruction descriptor: D = S32, A = B = signed 8-bit, both K-major, N = 256, M = 128
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        for (int kb = 0; kb < 2; ++kb)
            for (int k = 0; k < 4; ++k) {  // UMMA_K = 32 bytes inside the 128-byte swizzled row
                const uint64_t da = umma_desc_sw128(smem_u32(sA + kb * 16384) + k * 32);
                const uint64_t db = umma_desc_sw128(smem_u32(sB + kb * 32768) + k * 32);
                const uint32_t acc = (kb | k) ? 1u : 0u;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barMma) : "memory");
    }
    __syncwarp();
    mbar_wait(barMma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    // warp w reads TMEM lanes [32 w, 32 w + 32): thread = one row of D, 32 columns per load
    for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
            "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = 32 * warp + lane;
        for (int j = 0; j < 32; ++j) C[row * 256 + c0 + j] = (int)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int M = 128, N = 256, K = 256;
    std::vector<int8_t> A(M * K), B(N * K);
    srand(7);
    for (auto& x : A) x = (rand() & 1) ? 1 : -1;
    for (auto& x : B) x = (rand() & 1) ? 1 : -1;
    int8_t *dA, *dB;
    int* dC;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dC, sizeof(int) * M * N));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xff, sizeof(int) * M * N));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    TMap mA{}, mB{};
    auto enc = [&](TMap& m, void* ptr, int rows, int boxRows) {
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)K};
        const cuuint32_t box[2] = {128, (cuuint32_t)boxRows}, es[2] = {1, 1};
        CUresult r = ((EncodeTiled)fn)((CUtensorMap*)&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    };
    enc(mA, dA, M, 128);
    enc(mB, dB, N, 256);
    const size_t smem = 96 * 1024 + 1024;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<<<1, 128, smem>>>(mA, mB, dC);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<int> C(M * N);
    CK(cudaMemcpy(C.data(), dC, sizeof(int) * M * N, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            int s = 0;
            for (int k = 0; k < K; ++k) s += (int)A[m * K + k] * (int)B[n * K + k];
            if (s != C[m * N + n]) { if (bad < 5) printf("mismatch at (%d,%d): got %d want %d\n", m, n, C[m * N + n], s); ++bad; }
        }
    printf("UMMA_PROBE %s: %ld of %d mismatches\n", bad ? "FAIL" : "OK", bad, M * N);
    return bad ? 2 : 0;
}
