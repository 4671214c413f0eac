// Brute-force Hamming phase 1 on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::i8 on +-1-expanded descriptors.
//
// k_bow_dense computes 256-bit Hamming distances with XOR + carry-save adders + 5 POPC per pair of descriptors and is bound
// by the POPC / ALU pipes (0.75 T distances/s per B200).  With every bit b stored as the int8 value 2b - 1, the dot product
// of two expanded descriptors is 256 - 2 * Hamming: a 128 x 256 x 256 int8 GEMM tile yields 32,768 distances in eight
// tcgen05.mma instructions (ptxas expands the 1-bit mma.sync that PTX still offers into IMMA + unpacking on sm_100a:
// profiles/r01_b1_mma_sass_histogram.txt, so int8 it is).  Operands are 8x larger than the packed bits; they are expanded once
// per descriptor array (k_expand_pm1) and staged by TMA with the 128-byte swizzle the UMMA shared-memory descriptors name.
//
// Persistent CTA, six warps, one role each:
//   warp 0      TMA producer: the query tile of a work item (128 rows x 256 B, two 128-byte k-blocks) once, then the target
//               tiles (256 rows x 256 B) through a two-stage ring
//   warp 1      MMA issuer (one elected lane): 8 x tcgen05.mma.cta_group::1.kind::i8 (M 128, N 256, K 32) per target tile
//               into one of two 256-column TMEM accumulators; tcgen05.commit releases the smem stage and publishes the tile
//   warps 2..5  epilogue: warp w owns TMEM lanes [32 (w % 4), +32) = 32 query rows, thread = one query; tcgen05.ld 32 columns
//               at a time, dot > 256 - 2 D  <=>  distance < D, candidates appended in target order exactly like
//               k_bow_dense (first NEAR_K kept, all counted); the near list of a query lives in registers across the target
//               tiles of its work item and is written once
// A work item = (pair, tile of 128 queries).  Everything phase 2 (k_bow_resolve) reads has the same layout as before.
#pragma once
#include <stdint.h>

namespace eaof_umma {

struct alignas(64) TMap { unsigned long long v[16]; };

struct Args {
    const int* pairQ;      // [nPairs] block of the query frame
    const int* pairT;      // [nPairs] block of the target frame
    const int* counts;     // descriptors per block
    int blockStride;       // rows per block in the expanded array (= per-pair stride of nearBuf)
    int nPairs;
    int qTiles;            // ceil(blockStride / 128)
    int D;                 // near-list threshold: keep distance < D
};

constexpr int kNearK = 7;
constexpr int kTileQ = 128, kTileT = 256;
constexpr int kABytes = 2 * kTileQ * 128;   // two k-blocks of [128 rows x 128 B]
constexpr int kBBytes = 2 * kTileT * 128;   // two k-blocks of [256 rows x 128 B]
constexpr int kSmemBytes = kABytes + 2 * kBBytes + 1024;
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const void* tmap, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tmap), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
// K-major operand, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smemAddr) {
    return (uint64_t)((smemAddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmemD, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmemD), "l"(da),
                 "l"(db), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// max of 32 signed values in 16 three-input DPX instructions
__device__ __forceinline__ int max32(const uint32_t (&v)[32]) {
    int m = __vimax3_s32((int)v[0], (int)v[1], (int)v[2]);
#pragma unroll
    for (int j = 3; j + 1 < 32; j += 2) m = __vimax3_s32(m, (int)v[j], (int)v[j + 1]);
    return max(m, (int)v[31]);
}

// bit b of a 256-bit descriptor -> int8 2b - 1; thread = 16 bits -> 16 bytes
__global__ void __launch_bounds__(256) k_expand_pm1(const uint8_t* __restrict__ desc, int8_t* __restrict__ out, size_t nDesc) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;  // 16 chunks per descriptor
    if (i >= nDesc * 16) return;
    const unsigned bits = reinterpret_cast<const uint16_t*>(desc)[i];
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t x = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) x |= (((bits >> (4 * k + j)) & 1u) ? 0x01u : 0xffu) << (8 * j);
        w[k] = x;
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(kThreads, 1) k_bow_dense_umma(const __grid_constant__ TMap mapA, const __grid_constant__ TMap mapB,
                                                                const Args U, uint32_t* __restrict__ nearBuf) {
    extern __shared__ __align__(1024) uint8_t ummaSmem[];
    __shared__ __align__(8) unsigned long long bars[10];
    __shared__ uint32_t tmemBase;
    uint8_t* base = ummaSmem + ((1024u - (s32(ummaSmem) & 1023u)) & 1023u);
    const uint32_t sA = s32(base), sB0 = sA + kABytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // barriers: 0 fullA, 1 emptyA, 2-3 fullB[stage], 4-5 emptyB[stage], 6-7 tmemFull[acc], 8-9 tmemEmpty[acc]
    const uint32_t bar0 = s32(&bars[0]);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) bar_init(BAR(i), 1);
        bar_init(BAR(8), 4);
        bar_init(BAR(9), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmemBase)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmemBase;
    const int nItems = U.nPairs * U.qTiles;
    uint32_t aIter = 0, bIter = 0;  // per-role running counters: every role walks the same sequence of items and tiles

    if (warp == 0) {
        // ---------------- TMA producer
        if (lane == 0)
            for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
                const int pair = item / U.qTiles, qt = item - pair * U.qTiles;
                const int bq = U.pairQ[pair], bt = U.pairT[pair];
                const int nq = U.counts[bq], nt = U.counts[bt];
                if (qt * kTileQ >= nq || nt <= 0) continue;
                bar_wait(BAR(1), (aIter & 1u) ^ 1u);
                bar_expect_tx(BAR(0), kABytes);
                tma_2d(sA, &mapA, 0, bq * U.blockStride + qt * kTileQ, BAR(0));
                tma_2d(sA + kTileQ * 128, &mapA, 128, bq * U.blockStride + qt * kTileQ, BAR(0));
                ++aIter;
                for (int t0 = 0; t0 < nt; t0 += kTileT, ++bIter) {
                    const uint32_t st = bIter & 1u;
                    bar_wait(BAR(4 + st), ((bIter >> 1) & 1u) ^ 1u);
                    bar_expect_tx(BAR(2 + st), kBBytes);
                    tma_2d(sB0 + st * kBBytes, &mapB, 0, bt * U.blockStride + t0, BAR(2 + st));
                    tma_2d(sB0 + st * kBBytes + kTileT * 128, &mapB, 128, bt * U.blockStride + t0, BAR(2 + st));
                }
            }
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            // instruction descriptor: D = S32, A = B = signed 8-bit, both K-major, N = 256, M = 128 (cute::UMMA::InstrDescriptor)
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileT >> 3) << 17) | ((uint32_t)(kTileQ >> 4) << 24);
            for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
                const int pair = item / U.qTiles, qt = item - pair * U.qTiles;
                const int nq = U.counts[U.pairQ[pair]], nt = U.counts[U.pairT[pair]];
                if (qt * kTileQ >= nq || nt <= 0) continue;
                bar_wait(BAR(0), aIter & 1u);
                ++aIter;
                for (int t0 = 0; t0 < nt; t0 += kTileT, ++bIter) {
                    const uint32_t st = bIter & 1u, ph = (bIter >> 1) & 1u;
                    bar_wait(BAR(2 + st), ph);         // target tile landed
                    bar_wait(BAR(8 + st), ph ^ 1u);    // accumulator st drained by the epilogue
                    asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_i8(tmem + st * kTileT, desc_sw128(sA + kb * kTileQ * 128 + k * 32),
                                    desc_sw128(sB0 + st * kBBytes + kb * kTileT * 128 + k * 32), idesc, (kb | k) ? 1u : 0u);
                    umma_commit(BAR(4 + st));  // smem stage free once these MMAs have read it
                    umma_commit(BAR(6 + st));  // accumulator complete
                }
                umma_commit(BAR(1));           // query tile free
            }
        }
    } else {
        // ---------------- epilogue: TMEM lanes [32 (warp % 4), +32), thread = one query
        const int quarter = warp & 3;
        const int thr = 256 - 2 * U.D;  // dot > thr  <=>  distance < D
        for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
            const int pair = item / U.qTiles, qt = item - pair * U.qTiles;
            const int nq = U.counts[U.pairQ[pair]], nt = U.counts[U.pairT[pair]];
            if (qt * kTileQ >= nq) continue;
            const int q = qt * kTileQ + 32 * quarter + lane;
            if (nt <= 0) {  // no targets: empty lists (the other roles skip this item)
                if (q < nq) {
                    uint32_t* o = nearBuf + ((size_t)pair * U.blockStride + q) * 8;
                    reinterpret_cast<uint4*>(o)[0] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                    reinterpret_cast<uint4*>(o)[1] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u);
                }
                continue;
            }
            uint32_t near[kNearK];
#pragma unroll
            for (int i = 0; i < kNearK; ++i) near[i] = 0xffffffffu;
            int cnt = 0;
            for (int t0 = 0; t0 < nt; t0 += kTileT, ++bIter) {
                const uint32_t st = bIter & 1u, ph = (bIter >> 1) & 1u;
                bar_wait(BAR(6 + st), ph);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int cols = min(kTileT, nt - t0);
                // Candidates are rare (a query meets a handful of targets below D in a whole frame), and a branch per
                // element costs its resolve latency 256 times per tile with only one epilogue warp per scheduler to hide
                // it: 64 columns are loaded at a time, each half reduced to its maximum, and only a half whose maximum
                // exceeds the threshold is scanned element by element.
                auto scan = [&](const uint32_t (&v)[32], const int c0) {
                    unsigned m = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) m |= ((int)v[j] > thr ? 1u : 0u) << j;
                    if (cols - c0 < 32) m &= (1u << (cols - c0)) - 1u;
                    while (m) {  // ascending column = target order
                        const int b = __ffs((int)m) - 1;
                        m &= m - 1;
                        int dot = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j == b) dot = (int)v[j];
                        const uint32_t e = (uint32_t)(t0 + c0 + b) | ((uint32_t)((256 - dot) >> 1) << 16);
#pragma unroll
                        for (int i = 0; i < kNearK; ++i)
                            if (i == cnt) near[i] = e;
                        ++cnt;
                    }
                };
                const uint32_t tbase = tmem + ((uint32_t)(32 * quarter) << 16) + st * kTileT;
                // Reading the accumulator is the floor of this kernel (TMEM reads run at 64 B per clock and SM: 2048 clocks for
                // the 128 KB of a tile, twice the time of its eight MMAs), so the loads are kept back to back: the next 64
                // columns are requested before the current 64 are examined (tcgen05.wait::ld waits for everything outstanding,
                // hence wait -> request next -> examine).  Columns beyond `cols` hold other rows' products: bounded in scan().
                // (tcgen05.ld ... .pack::16b, two columns per register, was measured slower: 1.54 ms against 0.82 per 1024 pairs.)
                uint32_t a0[32], b0[32], a1[32], b1[32];
                tmem_ld32(tbase, a0);
                tmem_ld32(tbase + 32, b0);
#pragma unroll
                for (int step = 0; step < kTileT / 64; ++step) {
                    const int c0 = 64 * step;
                    if (c0 >= cols) break;
                    tmem_ld_wait();
                    const bool more = c0 + 64 < cols;
                    if (step & 1) {
                        if (more) { tmem_ld32(tbase + c0 + 64, a0); tmem_ld32(tbase + c0 + 96, b0); }
                        if (max32(a1) > thr) scan(a1, c0);
                        if (c0 + 32 < cols && max32(b1) > thr) scan(b1, c0 + 32);
                    } else {
                        if (more) { tmem_ld32(tbase + c0 + 64, a1); tmem_ld32(tbase + c0 + 96, b1); }
                        if (max32(a0) > thr) scan(a0, c0);
                        if (c0 + 32 < cols && max32(b0) > thr) scan(b0, c0 + 32);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) bar_arrive(BAR(8 + st));
            }
            if (q < nq) {
                uint32_t* o = nearBuf + ((size_t)pair * U.blockStride + q) * 8;
                reinterpret_cast<uint4*>(o)[0] = make_uint4(near[0], near[1], near[2], near[3]);
                reinterpret_cast<uint4*>(o)[1] = make_uint4(near[4], near[5], near[6], (uint32_t)cnt);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace eaof_umma
