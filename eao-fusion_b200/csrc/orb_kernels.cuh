// Hand-written sm_100a kernels of the ORB extraction path (one batch of frames per launch).
//
//   k_level0        ComputePyramid level 0: copyMakeBorder(REFLECT_101)            src/ORBextractor.cc:1125-1129
//   k_resize        ComputePyramid level l>0: resize(INTER_LINEAR)+copyMakeBorder  src/ORBextractor.cc:1118-1124
//   k_fast          per-cell FAST-9/16 + NMS + iniThFAST/minThFAST retry           src/ORBextractor.cc:789-829
//   k_octree        DistributeOctTree (+DivideNode), one CTA per (frame, level)    src/ORBextractor.cc:481-763
//   k_blur          GaussianBlur 7x7 sigma 2, integer arithmetic                   src/ORBextractor.cc:1085-1086
//   k_angle_desc    IC_Angle + computeOrbDescriptor + keypoint record              src/ORBextractor.cc:77-147,837-847,1095-1101
//
// Exactness rules: all pixel work is integer; the three float computations
// (fastAtan2, angle*pi/180 -> sincosf, pattern rotation) use explicit round-to-nearest intrinsics so nvcc can
// never contract a multiply-add, and sincosf restates glibc's algorithm because the reference calls
// std::cos(float)/std::sin(float) (= cosf/sinf), see DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <stdint.h>

#include "geom.h"
#include "orb_pattern.h"

namespace eaof {

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// ---- TMA / mbarrier primitives (sm_100a PTX) shared by k_pyramid_fused and the k_fast_tma variants
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int x, int y, int z, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

// Bulk asynchronous copies (the 1-D form of TMA: no tensor map, 16-byte aligned addresses and sizes).
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_commit_and_drain() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may be released once it has been read
}

// One CUtensorMap (128 bytes, 64-byte aligned) per pyramid level, passed as a __grid_constant__ kernel parameter (the
// canonical way: a descriptor that merely sits in global memory would need a fence.proxy.tensormap acquire first).
struct alignas(64) FastTmaMaps {
    unsigned long long m[EAOF_MAX_LEVELS][16];
};


// ------------------------------------------------------------------------------------------------
// Pyramid.  A level row is `pitch` bytes; inner pixel x sits at byte EAOF_INNER_X0 + x, so the bordered
// span is bytes [13, w+51).  
// Each thread produces one aligned 4-byte word of LEVEL0_ROWS consecutive bordered rows: the column part of the address
// arithmetic (reflected byte offsets, alignment test) is done once per thread.
#define LEVEL0_ROWS 8
__global__ void __launch_bounds__(256) k_level0(const uint8_t* __restrict__ in, size_t framePitch, size_t stride,
                                                uint8_t* __restrict__ pyr, const __grid_constant__ Geom g) {
    const LevelGeom& L = g.L[0];
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = (blockIdx.y * blockDim.y + threadIdx.y) * LEVEL0_ROWS;
    const int f = blockIdx.z;
    const int c0 = 12 + 4 * wi;
    if (r0 >= L.rows || c0 >= L.w + 52) return;
    const int bx = c0 - EAOF_INNER_X0;
    const uint8_t* base = in + (size_t)f * framePitch;
    const bool fast = bx >= 0 && bx + 3 < L.w && ((reinterpret_cast<uintptr_t>(base + bx) | stride) & 3) == 0;
    int xs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = reflect101(bx + j, L.w);
    uint8_t* dst = pyr + (size_t)f * g.pyrFrameBytes + L.off + c0;
#pragma unroll
    for (int k = 0; k < LEVEL0_ROWS; ++k) {
        const int by = r0 + k;
        if (by >= L.rows) break;
        const uint8_t* src = base + (size_t)reflect101(by - EAOF_EDGE, L.h) * stride;
        uint32_t v;
        if (fast) {
            v = __ldg(reinterpret_cast<const uint32_t*>(src + bx));
        } else {
            v = (uint32_t)__ldg(src + xs[0]) | ((uint32_t)__ldg(src + xs[1]) << 8) | ((uint32_t)__ldg(src + xs[2]) << 16) |
                ((uint32_t)__ldg(src + xs[3]) << 24);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)by * L.pitch) = v;
    }
}

// k_level0_bulk: the same copyMakeBorder as a pair of bulk asynchronous copies.  Level 0 is a copy of the input with a
// reflected border; a thread per 4 bytes spends ~100 issue slots on addresses for one load and one store and waits on the
// load (72 % of k_level0's stall samples).  Here a CTA takes `rowsPerCta` inner rows: one cp.async.bulk per row brings the
// pixels into shared memory at the byte offset they have inside a padded level row (all rows complete on one mbarrier),
// the threads write the 19 + 19 reflected border pixels of each row next to them, and one cp.async.bulk per row writes
// the whole bordered span back; rows that the top / bottom border mirrors are stored a second time from the same bytes.
// Needs 16-byte aligned input rows (base, stride, frame pitch) and width % 16 == 0, h >= 20: k_level0 otherwise.
__global__ void __launch_bounds__(128) k_level0_bulk(const uint8_t* __restrict__ in, size_t framePitch, size_t stride,
                                                     uint8_t* __restrict__ pyr, const __grid_constant__ Geom g, int rowsPerCta) {
    extern __shared__ __align__(128) uint8_t l0Smem[];
    __shared__ __align__(8) unsigned long long l0Bar;
    const LevelGeom& L = g.L[0];
    const int w = L.w, h = L.h;
    const int span = (EAOF_INNER_X0 + w + EAOF_EDGE + 15) & ~15;  // bytes [0, span) of a padded row hold the bordered pixels
    const int y0 = blockIdx.x * rowsPerCta, nr = min(rowsPerCta, h - y0);
    const int f = blockIdx.y;
    const uint32_t bar = smem_u32(&l0Bar), base = smem_u32(l0Smem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)(nr * w));
    __syncthreads();
    if (threadIdx.x < nr)
        bulk_load(base + threadIdx.x * span + EAOF_INNER_X0, in + (size_t)f * framePitch + (size_t)(y0 + threadIdx.x) * stride, (uint32_t)w, bar);
    for (int r = threadIdx.x + 128; r < nr; r += 128)
        bulk_load(base + r * span + EAOF_INNER_X0, in + (size_t)f * framePitch + (size_t)(y0 + r) * stride, (uint32_t)w, bar);
    mbar_wait(bar, 0);
    // everything outside the inner pixels: reflected border where the bordered image has pixels, zero elsewhere
    const int rightN = span - (EAOF_INNER_X0 + w), side = EAOF_INNER_X0 + rightN;
    for (int i = threadIdx.x; i < nr * side; i += blockDim.x) {
        const int r = i / side, k = i - r * side;
        uint8_t* row = l0Smem + r * span;
        if (k < EAOF_INNER_X0) {  // byte k: pixel x = k - 32 < 0 -> pixel -x
            const int x = EAOF_INNER_X0 - k;
            row[k] = x <= EAOF_EDGE ? row[EAOF_INNER_X0 + x] : (uint8_t)0;
        } else {                  // byte 32 + w + d: pixel w + d -> pixel w - 2 - d
            const int d = k - EAOF_INNER_X0;
            row[EAOF_INNER_X0 + w + d] = d < EAOF_EDGE ? row[EAOF_INNER_X0 + w - 2 - d] : (uint8_t)0;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes above -> async-proxy reads below
    __syncthreads();
    uint8_t* dst = pyr + (size_t)f * g.pyrFrameBytes + L.off;
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const int y = y0 + r;
        const uint32_t src = base + r * span;
        bulk_store(dst + (size_t)(y + EAOF_EDGE) * L.pitch, src, (uint32_t)span);
        if (y >= 1 && y <= EAOF_EDGE) bulk_store(dst + (size_t)(EAOF_EDGE - y) * L.pitch, src, (uint32_t)span);
        if (y >= h - 1 - EAOF_EDGE && y <= h - 2) bulk_store(dst + (size_t)(EAOF_EDGE + 2 * (h - 1) - y) * L.pitch, src, (uint32_t)span);
        bulk_store_commit_and_drain();
    }
}

// Colour ingest: cv::cvtColor(RGB/BGR/RGBA/BGRA -> GRAY) on 8U (src/Tracking.cc:324-337, OpenCV RGB2Gray<uchar>) fused
// with the level-0 copyMakeBorder.  gray = (c0*k0 + c1*k1 + c2*k2 + 2^(shift-1)) >> shift with the coefficients in
// channel order (the caller swaps them for RGB vs BGR); alpha is ignored.  Same work decomposition as k_level0.
template <int CH>
__global__ void __launch_bounds__(256) k_level0_color(const uint8_t* __restrict__ in, size_t framePitch, size_t stride,
                                                      uint8_t* __restrict__ pyr, const __grid_constant__ Geom g, int k0,
                                                      int k1, int k2, int shift) {
    const LevelGeom& L = g.L[0];
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int by = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    const int c0 = 12 + 4 * wi;
    if (by >= L.rows || c0 >= L.w + 52) return;
    const int y = reflect101(by - EAOF_EDGE, L.h);
    const uint8_t* src = in + (size_t)f * framePitch + (size_t)y * stride;
    const int bx = c0 - EAOF_INNER_X0;
    const int rnd = 1 << (shift - 1);
    uint32_t v = 0;
    if (bx >= 0 && bx + 3 < L.w && ((reinterpret_cast<uintptr_t>(src + (size_t)bx * CH) & (CH == 4 ? 15 : 3)) == 0)) {
        uint32_t w[CH];  // 4 pixels = CH aligned words
        if (CH == 4) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + (size_t)bx * 4));
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[CH - 1] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) w[j] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)bx * 3) + j);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int b = j * CH + k;
                c[k] = (w[b >> 2] >> (8 * (b & 3))) & 0xff;
            }
            v |= (uint32_t)((c[0] * k0 + c[1] * k1 + c[2] * k2 + rnd) >> shift) << (8 * j);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint8_t* p = src + (size_t)reflect101(bx + j, L.w) * CH;
            v |= (uint32_t)(((int)__ldg(p) * k0 + (int)__ldg(p + 1) * k1 + (int)__ldg(p + 2) * k2 + rnd) >> shift) << (8 * j);
        }
    }
    *reinterpret_cast<uint32_t*>(pyr + (size_t)f * g.pyrFrameBytes + L.off + (size_t)by * L.pitch + c0) = v;
}

// Frame::ComputeStereoFromRGBD (src/Frame.cc:1016-1037) over the keypoints a handle holds on the device: depth at the
// (truncated) raw keypoint position, mvDepth = d and mvuRight = x_undistorted - mbf/d where d > 0, -1 elsewhere.
// U16: raw sensor depth converted like Tracking's imDepth.convertTo(CV_32F, mDepthMapFactor) (src/Tracking.cc:340-341).
template <bool U16>
__global__ void __launch_bounds__(256) k_stereo_from_rgbd(const eaof_kp* __restrict__ kps, const int* __restrict__ counts,
                                                          int cap, const void* __restrict__ depth, size_t strideBytes,
                                                          size_t framePitchBytes, float depthScale, int w, int h,
                                                          const float* __restrict__ xUn, float mbf,
                                                          float* __restrict__ uRight, float* __restrict__ depthOut) {
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[f]) return;
    const size_t o = (size_t)f * cap + i;
    const eaof_kp kp = kps[o];
    const int u = (int)kp.x, v = (int)kp.y;  // cv::Mat::at<float>(float, float): implicit conversion, truncation
    float d = -1.f;
    if (u >= 0 && u < w && v >= 0 && v < h) {
        const uint8_t* row = static_cast<const uint8_t*>(depth) + (size_t)f * framePitchBytes + (size_t)v * strideBytes;
        d = U16 ? __fmul_rn((float)reinterpret_cast<const uint16_t*>(row)[u], depthScale) : reinterpret_cast<const float*>(row)[u];
    }
    const bool ok = d > 0;
    depthOut[o] = ok ? d : -1.f;
    uRight[o] = ok ? __fsub_rn(xUn ? xUn[o] : kp.x, __fdiv_rn(mbf, d)) : -1.f;
}

// Frame::ComputeStereoMatches (src/Frame.cc:841-1013) over the results two extractor handles (left / right camera) hold on
// the device.  k_stereo_match: warp = one left keypoint.  (1) the lanes sweep the right keypoints: row band of the right
// keypoint contains the left row (:855-866, :885), octave within +-1, u inside [uL - maxD, uL + 3], Hamming distance;
// warp argmin by (distance, right index) = the reference's first-minimum over ascending iR, accepted below TH_HIGH.
// (2) sub-pixel refinement on the level images: the 121 pixels of the 11x11 patch are dealt to the lanes, the 11 SADs of
// (patch - its centre) against the right patches at shifts -5..5 are integer sums reduced across the warp; lane 0 picks
// the first smallest, fits the parabola (fp32, same operation order) and applies the disparity gates.  sad = -1 where no
// match survives.  k_stereo_filter: CTA = frame, median of the SADs by rank counting, matches at or above
// 1.5f*1.4f*median are removed (:1002-1012).
struct StereoArgs {
    const eaof_kp* kpL; const uint8_t* descL; const int* cntL; const uint8_t* pyrL;
    const eaof_kp* kpR; const uint8_t* descR; const int* cntR; const uint8_t* pyrR;
    int capL, capR;
    float invScale[EAOF_MAX_LEVELS];
    float mb, mbf;
};

__global__ void __launch_bounds__(512) k_stereo_match(StereoArgs A, float* __restrict__ uRight, float* __restrict__ depth,
                                                      int* __restrict__ sad, const __grid_constant__ Geom g) {
    // the right frame's keypoints, staged once per CTA in the form the sweep tests: row band (minr | maxr << 16), u, octave
    extern __shared__ __align__(16) unsigned char stereoSmem[];
    const int nR = A.cntR[blockIdx.y];
    int* sBand = reinterpret_cast<int*>(stereoSmem);
    float* sU = reinterpret_cast<float*>(sBand + A.capR);
    unsigned char* sOct = reinterpret_cast<unsigned char*>(sU + A.capR);
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    for (int iR = threadIdx.x; iR < nR; iR += blockDim.x) {
        const eaof_kp kR = A.kpR[(size_t)f * A.capR + iR];
        const float r = __fmul_rn(2.0f, g.L[kR.octave].scale);
        const int maxr = (int)ceilf(__fadd_rn(kR.y, r)), minr = (int)floorf(__fsub_rn(kR.y, r));
        sBand[iR] = (max(minr, 0) & 0xffff) | (max(maxr, -1) << 16);  // rows are >= 0; a band ending above row 0 holds none
        sU[iR] = kR.x;
        sOct[iR] = (unsigned char)kR.octave;
    }
    __syncthreads();
    const int iL = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (iL >= A.cntL[f]) return;  // no CTA-wide barrier below
    const size_t oL = (size_t)f * A.capL + iL;
    const eaof_kp kL = A.kpL[oL];
    const int levelL = kL.octave;
    const float uL = kL.x, vL = kL.y;
    const float maxD = __fdiv_rn(A.mbf, A.mb);
    const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, -3.f);
    const int row = (int)vL;
    uint32_t q[8];
    {
        const uint4* p = reinterpret_cast<const uint4*>(A.descL + 32 * oL);
        const uint4 a = p[0], b = p[1];
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    unsigned best = 0xffffffffu;  // (distance << 16) | iR
    if (!(maxU < 0))
        for (int iR = lane; iR < nR; iR += 32) {
            const int band = sBand[iR];
            if (row < (band & 0xffff) || row > (band >> 16)) continue;
            const int oR = sOct[iR];
            if (oR < levelL - 1 || oR > levelL + 1) continue;
            const float uR = sU[iR];
            if (!(uR >= minU && uR <= maxU)) continue;
            const uint4* p = reinterpret_cast<const uint4*>(A.descR + 32 * ((size_t)f * A.capR + iR));
            const uint4 a = p[0], b = p[1];
            const int d = __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
                          __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
            const unsigned key = ((unsigned)d << 16) | (unsigned)iR;
            best = key < best ? key : best;
        }
#pragma unroll
    for (int o = 16; o; o >>= 1) { const unsigned t = __shfl_xor_sync(0xffffffffu, best, o); best = t < best ? t : best; }
    float outU = -1.f, outD = -1.f;
    int outS = -1;
    if (best != 0xffffffffu && (int)(best >> 16) < 100) {  // bestDist < ORBmatcher::TH_HIGH, :913
        const LevelGeom& L = g.L[levelL];
        const float uR0 = A.kpR[(size_t)f * A.capR + (best & 0xffffu)].x;
        const float sc = A.invScale[levelL];
        const float scaleduL = roundf(__fmul_rn(uL, sc)), scaledvL = roundf(__fmul_rn(vL, sc)), scaleduR0 = roundf(__fmul_rn(uR0, sc));
        const float iniu = scaleduR0, endu = __fadd_rn(scaleduR0, 11.f);  // scaleduR0+L-w, scaleduR0+L+w+1 with L = w = 5
        if (!(iniu < 0 || endu >= (float)L.w)) {
            const int y0 = (int)(scaledvL - 5.f), x0 = (int)(scaleduL - 5.f), xr = (int)(scaleduR0 - 5.f);
            const uint8_t* bL = A.pyrL + (size_t)f * g.pyrFrameBytes + L.off + (size_t)(y0 + EAOF_EDGE) * L.pitch + EAOF_INNER_X0;
            const uint8_t* bR = A.pyrR + (size_t)f * g.pyrFrameBytes + L.off + (size_t)(y0 + EAOF_EDGE) * L.pitch + EAOF_INNER_X0;
            const int cL = bL[5 * L.pitch + x0 + 5];
            int acc[11], cR[11];
#pragma unroll
            for (int k = 0; k < 11; ++k) { acc[k] = 0; cR[k] = bR[5 * L.pitch + xr + k]; }  // centre of the patch at shift k-5
            for (int p = lane; p < 121; p += 32) {
                const int a = p / 11, b = p - a * 11;
                const int il = (int)bL[a * L.pitch + x0 + b] - cL;
                const uint8_t* rr = bR + a * L.pitch + xr + b - 5;
#pragma unroll
                for (int k = 0; k < 11; ++k) acc[k] += abs(il - ((int)rr[k] - cR[k]));
            }
#pragma unroll
            for (int k = 0; k < 11; ++k)
#pragma unroll
                for (int o = 16; o; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            if (lane == 0) {
                int bestS = 0x7fffffff, bestInc = 0;
#pragma unroll
                for (int k = 0; k < 11; ++k)
                    if ((float)acc[k] < (float)bestS) { bestS = acc[k]; bestInc = k - 5; }
                if (bestInc != -5 && bestInc != 5) {
                    float d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
                    for (int k = 1; k < 10; ++k)
                        if (k - 5 == bestInc) { d1 = (float)acc[k - 1]; d2 = (float)acc[k]; d3 = (float)acc[k + 1]; }
                    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3),
                                                   __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                    if (!(deltaR < -1 || deltaR > 1)) {
                        float bestuR = __fmul_rn(L.scale, __fadd_rn(__fadd_rn(scaleduR0, (float)bestInc), deltaR));
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= 0 && disparity < maxD) {
                            if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                            outD = __fdiv_rn(A.mbf, disparity);
                            outU = bestuR;
                            outS = bestS;
                        }
                    }
                }
            }
        }
    }
    if (lane == 0) { uRight[oL] = outU; depth[oL] = outD; sad[oL] = outS; }
}

__global__ void __launch_bounds__(1024) k_stereo_filter(const int* __restrict__ cntL, int cap, float* __restrict__ uRight,
                                                        float* __restrict__ depth, const int* __restrict__ sad) {
    extern __shared__ int sS[];
    __shared__ int sMedian, sCount;
    const int f = blockIdx.x, n = cntL[f], tid = threadIdx.x;
    const size_t o = (size_t)f * cap;
    if (tid == 0) { sMedian = -1; sCount = 0; }
    __syncthreads();
    int cnt = 0;
    for (int i = tid; i < n; i += 1024) { const int v = sad[o + i]; sS[i] = v; cnt += v >= 0; }
    atomicAdd(&sCount, cnt);
    __syncthreads();
    const int m = sCount;
    if (m == 0) return;
    const int target = m / 2;  // vDistIdx[vDistIdx.size()/2] of the pairs sorted by (SAD, left index)
    for (int i = tid; i < n; i += 1024) {
        const int v = sS[i];
        if (v < 0) continue;
        int r = 0;
        for (int j = 0; j < n; ++j) { const int u = sS[j]; r += u >= 0 && (u < v || (u == v && j < i)); }
        if (r == target) sMedian = v;
    }
    __syncthreads();
    const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)sMedian);
    for (int i = tid; i < n; i += 1024) {
        const int v = sS[i];
        if (v >= 0 && !((float)v < thDist)) { uRight[o + i] = -1.f; depth[o + i] = -1.f; }
    }
}

// Frame::UndistortKeyPoints (src/Frame.cc:773-803) = cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) over
// the keypoints a handle holds on the device: OpenCV's cvUndistortPoints in double arithmetic (the library is built with
// --fmad=false, double division is IEEE): normalise, five fixed-point iterations of the inverse Brown model, re-project
// with P = mK, round to float.  thread = keypoint.
struct UndistortArgs {
    double fx, fy, cx, cy;
    double k[12];
    int guard;  // 1: OpenCV 4.x "icdist < 0" exit
};

__global__ void __launch_bounds__(256) k_undistort(const eaof_kp* __restrict__ kps, const int* __restrict__ counts, int cap,
                                                   UndistortArgs U, float* __restrict__ xo, float* __restrict__ yo) {
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[f]) return;
    const size_t o = (size_t)f * cap + i;
    const double u = kps[o].x, v = kps[o].y;
    const double ifx = 1. / U.fx, ify = 1. / U.fy;
    double x = (u - U.cx) * ifx, y = (v - U.cy) * ify;
    const double x0 = x, y0 = y;
    const double* k = U.k;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (U.guard && icdist < 0) { x = (u - U.cx) * ifx; y = (v - U.cy) * ify; break; }
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = U.fx * x + 0.0 * y + U.cx, yy = 0.0 * x + U.fy * y + U.cy, ww = 1. / (0.0 * x + 0.0 * y + 1.0);
    xo[o] = (float)(xx * ww);
    yo[o] = (float)(yy * ww);
}

__global__ void __launch_bounds__(256) k_copy_xy(const eaof_kp* __restrict__ kps, const int* __restrict__ counts, int cap,
                                                 float* __restrict__ xo, float* __restrict__ yo) {
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[f]) return;
    const size_t o = (size_t)f * cap + i;
    xo[o] = kps[o].x;
    yo[o] = kps[o].y;
}

// cv::resize 8UC1 INTER_LINEAR with 11-bit fixed-point coefficients; the border pixel at
// bordered position (bx,by) equals the resized pixel at the reflected inner position, so resize and
// copyMakeBorder are one pass.  tabs: per destination column [sx, a0|a1<<16], per row [sy, b0|b1<<16].
__global__ void __launch_bounds__(256) k_resize_generic(uint8_t* __restrict__ pyr, const int* __restrict__ tabs,
                                                        const __grid_constant__ Geom g, int l) {
    const LevelGeom& D = g.L[l];
    const LevelGeom& S = g.L[l - 1];
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int by = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    const int c0 = 12 + 4 * wi;
    if (by >= D.rows || c0 >= D.w + 52) return;
    uint8_t* frame = pyr + (size_t)f * g.pyrFrameBytes;
    const int y = reflect101(by - EAOF_EDGE, D.h);
    const int sy = tabs[D.yTab + 2 * y];
    const int bb = tabs[D.yTab + 2 * y + 1];
    const int b0 = (short)(bb & 0xffff), b1 = bb >> 16;
    const int sy0 = min(max(sy, 0), S.h - 1), sy1 = min(max(sy + 1, 0), S.h - 1);
    const uint8_t* sIn = frame + S.off + (size_t)EAOF_EDGE * S.pitch + EAOF_INNER_X0;
    const uint8_t* r0 = sIn + (size_t)sy0 * S.pitch;
    const uint8_t* r1 = sIn + (size_t)sy1 * S.pitch;
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = reflect101(c0 - EAOF_INNER_X0 + j, D.w);
        const int sx = tabs[D.xTab + 2 * x];
        const int aa = tabs[D.xTab + 2 * x + 1];
        const int a0 = (short)(aa & 0xffff), a1 = aa >> 16;
        const int sx1 = min(sx + 1, S.w - 1);
        const int h0 = r0[sx] * a0 + r0[sx1] * a1;
        const int h1 = r1[sx] * a0 + r1[sx1] * a1;
        const int d = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v |= (uint32_t)(d & 0xff) << (8 * j);
    }
    *reinterpret_cast<uint32_t*>(frame + D.off + (size_t)by * D.pitch + c0) = v;
}

// Same arithmetic, organised for throughput (levels with at least 40 rows): one thread owns 4 bordered columns —
// their source word offsets, byte selectors and packed (a0, a1) coefficients are per-thread constants — and walks
// down RSZ_ROWS inner rows (32, 16 or 8: the host picks the largest that still gives the launch enough threads).  A horizontal pass of one source row is 4 x (2 word loads + PRMT + IDP.2A); the lower
// source row of one output row is usually the upper row of the next, so it is kept.  Rows that the REFLECT_101
// border mirrors (inner rows 1..19 and h-20..h-2) are stored twice.
#ifndef RSZ_THREADS
#define RSZ_THREADS 128
#endif
__device__ __forceinline__ unsigned mad_hi_u32(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// The 4 destination columns of a thread read source columns sx_j, sx_j + 1 that lie within 8 consecutive bytes (any scale
// factor up to 2; eaof_orb_create checks the tables and sends larger steps to k_resize_generic): three aligned words from
// ONE row pointer (immediate offsets 0 / 4 / 8), two funnel shifts that bring the window's first byte to bit 0, and one
// PRMT per column with a per-thread constant selector — instead of two loads from a 64-bit address of its own per column.
template <int RSZ_ROWS>
__global__ void __launch_bounds__(RSZ_THREADS) k_resize(uint8_t* __restrict__ pyr, const int* __restrict__ tabs,
                                                        const __grid_constant__ Geom g, int l) {
    const LevelGeom& D = g.L[l];
    const LevelGeom& S = g.L[l - 1];
    const int nCW = (D.w + 43) >> 2;
    const int t = blockIdx.x * RSZ_THREADS + threadIdx.x;
    const int rc = t / nCW, cwd = t - rc * nCW;
    const int y0 = rc * RSZ_ROWS;
    if (y0 >= D.h) return;
    const int f = blockIdx.y;
    uint8_t* frame = pyr + (size_t)f * g.pyrFrameBytes;
    const int c0 = 12 + 4 * cwd;
    int sx[4];
    unsigned sel[4], coef[4];
    int sxMin = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = reflect101(c0 - EAOF_INNER_X0 + j, D.w);
        const int2 e = __ldg(reinterpret_cast<const int2*>(tabs + D.xTab) + x);
        sx[j] = e.x;
        coef[j] = (unsigned)e.y;
        sxMin = min(sxMin, e.x);
    }
    const int wofs = sxMin >> 2;
    const unsigned sh = 8u * (unsigned)(sxMin & 3);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned dlt = (unsigned)(sx[j] - sxMin);  // <= 6: bytes dlt, dlt + 1 of the shifted window; a1 == 0 whenever
        sel[j] = dlt | ((dlt + 1u) << 4);                // sx + 1 would leave the image
    }
    const uint32_t* sIn = reinterpret_cast<const uint32_t*>(frame + S.off + (size_t)EAOF_EDGE * S.pitch + EAOF_INNER_X0) + wofs;
    const int sPitchW = S.pitch >> 2;
    uint8_t* dOut = frame + D.off + c0;
    int cached = -1;
    unsigned hB[4] = {0, 0, 0, 0};
    auto hrow = [&](int sy, unsigned (&h)[4]) {
        const uint32_t* r = sIn + (ptrdiff_t)sy * sPitchW;
        const unsigned w0 = __ldg(r), w1 = __ldg(r + 1), w2 = __ldg(r + 2);
        const unsigned X = __funnelshift_r(w0, w1, sh), Y = __funnelshift_r(w1, w2, sh);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __dp2a_lo(coef[j], __byte_perm(X, Y, sel[j]), 0u) >> 4;  // every use is (sum >> 4), A.2
    };
    const int yEnd = min(y0 + RSZ_ROWS, D.h);
    const int2* yt = reinterpret_cast<const int2*>(tabs + D.yTab);
    for (int y = y0; y < yEnd; ++y) {
        const int2 e = __ldg(yt + y);
        const int sy = e.x, bb = e.y;
        // (b * (S >> 4)) >> 16 == high word of (b << 16) * (S >> 4): one IMAD.HI (FMA pipe) per product, the second one
        // adding the first and the rounding constant; b in [0, 2048]
        const unsigned b0s = (unsigned)bb << 16, b1s = (unsigned)bb & 0xffff0000u;
        const int sy0 = min(max(sy, 0), S.h - 1), sy1 = min(max(sy + 1, 0), S.h - 1);
        unsigned hA[4];
        if (sy0 == cached) {
#pragma unroll
            for (int j = 0; j < 4; ++j) hA[j] = hB[j];
        } else {
            hrow(sy0, hA);
        }
        if (sy1 != sy0) hrow(sy1, hB);
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) hB[j] = hA[j];
        }
        cached = sy1;
        unsigned d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = mad_hi_u32(b1s, hB[j], mad_hi_u32(b0s, hA[j], 2u)) >> 2;  // <= 255
        const uint32_t v = ((d[3] * 256u + d[2]) * 256u + d[1]) * 256u + d[0];
        *reinterpret_cast<uint32_t*>(dOut + (size_t)(y + EAOF_EDGE) * D.pitch) = v;
        if (y >= 1 && y <= EAOF_EDGE) *reinterpret_cast<uint32_t*>(dOut + (size_t)(EAOF_EDGE - y) * D.pitch) = v;
        if (y >= D.h - 1 - EAOF_EDGE && y <= D.h - 2)
            *reinterpret_cast<uint32_t*>(dOut + (size_t)(EAOF_EDGE + 2 * (D.h - 1) - y) * D.pitch) = v;
    }
}

// k_resize_bulk: k_resize with the source rows of a CTA staged in shared memory by bulk asynchronous copies.
// k_resize is bound by the latency of its global loads (60 % of its stall samples wait on them; a register prefetch queue
// cost more than it hid).  Here a CTA owns RSZ_ROWS destination rows of one frame over the whole width: warp 0 requests the
// source rows those need ([sy(y0), sy(yEnd-1) + 1], about 1.2 RSZ_ROWS + 1 rows of S.w + 12 bytes) with one cp.async.bulk per
// row, all completing on one mbarrier; the threads meanwhile fetch their per-column constants, then run k_resize's row loop
// with three LDS per source row instead of three LDG (32-bit addresses, ~30 cycles instead of ~600).  Destination rows go
// straight to global memory as before (one aligned word per thread and row, coalesced), mirrored rows twice.
template <int RSZ_ROWS>
__global__ void __launch_bounds__(256) k_resize_bulk(uint8_t* __restrict__ pyr, const int* __restrict__ tabs,
                                                     const __grid_constant__ Geom g, int l) {
    extern __shared__ __align__(128) uint8_t rszSmem[];
    __shared__ __align__(8) unsigned long long rszBar;
    __shared__ int4 rszRows[16];  // per destination row of this CTA: source rows sy0, sy1 (clamped) and the coefficients b0, b1
    const LevelGeom& D = g.L[l];
    const LevelGeom& S = g.L[l - 1];
    const int y0 = blockIdx.x * RSZ_ROWS, yEnd = min(y0 + RSZ_ROWS, D.h);
    const int f = blockIdx.y;
    uint8_t* frame = pyr + (size_t)f * g.pyrFrameBytes;
    const int2* yt = reinterpret_cast<const int2*>(tabs + D.yTab);
    const int syA = min(max(__ldg(yt + y0).x, 0), S.h - 1), syB = min(max(__ldg(yt + yEnd - 1).x + 1, 0), S.h - 1);
    const int nS = syB - syA + 1;
    const int srcBytes = (S.w + 12 + 15) & ~15;
    const uint32_t bar = smem_u32(&rszBar), base = smem_u32(rszSmem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ((int)threadIdx.x < yEnd - y0) {  // the row constants are the same for every thread of the CTA: fetched and clamped once
        const int2 e = __ldg(yt + y0 + threadIdx.x);
        rszRows[threadIdx.x] = make_int4(min(max(e.x, 0), S.h - 1), min(max(e.x + 1, 0), S.h - 1), e.y & 0xffff, (int)((unsigned)e.y >> 16));
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)(nS * srcBytes));
        __syncwarp();
        const uint8_t* src = frame + S.off + (size_t)(EAOF_EDGE + syA) * S.pitch + EAOF_INNER_X0;
        for (int r = threadIdx.x; r < nS; r += 32) bulk_load(base + r * srcBytes, src + (size_t)r * S.pitch, (uint32_t)srcBytes, bar);
    }
    const int nCW = (D.w + 43) >> 2;
    const int srcWords = srcBytes >> 2;
    const uint32_t* sW = reinterpret_cast<const uint32_t*>(rszSmem);
    bool waited = false;
    for (int cwd = threadIdx.x; cwd < nCW; cwd += blockDim.x) {
        const int c0 = 12 + 4 * cwd;
        int sx[4];
        unsigned sel[4], coef[4];
        int sxMin = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = reflect101(c0 - EAOF_INNER_X0 + j, D.w);
            const int2 e = __ldg(reinterpret_cast<const int2*>(tabs + D.xTab) + x);
            sx[j] = e.x;
            coef[j] = (unsigned)e.y;
            sxMin = min(sxMin, e.x);
        }
        const unsigned sh = 8u * (unsigned)(sxMin & 3);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned dlt = (unsigned)(sx[j] - sxMin);
            sel[j] = dlt | ((dlt + 1u) << 4);
        }
        const int wofs = sxMin >> 2;
        uint8_t* dOut = frame + D.off + c0;
        if (!waited) {
            mbar_wait(bar, 0);
            waited = true;
        }
        int cached = -1;
        unsigned hB[4] = {0, 0, 0, 0};
        auto hrow = [&](int sy, unsigned (&h)[4]) {
            const uint32_t* r = sW + (sy - syA) * srcWords + wofs;
            const unsigned w0 = r[0], w1 = r[1], w2 = r[2];
            const unsigned X = __funnelshift_r(w0, w1, sh), Y = __funnelshift_r(w1, w2, sh);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __dp2a_lo(coef[j], __byte_perm(X, Y, sel[j]), 0u) >> 4;  // every use is (sum >> 4), A.2
        };
        // one destination row: horizontal pass of the source rows not yet held, vertical pass, packed word
        auto row = [&](int y) -> uint32_t {
            const int4 e = rszRows[y - y0];
            // b in [0, 2048], h < 2^15: the products fit 32 bits, (b * h) >> 16 is a plain shift (folded into the adds)
            const unsigned b0 = (unsigned)e.z, b1 = (unsigned)e.w;
            const int sy0 = e.x, sy1 = e.y;
            unsigned hA[4];
            if (sy0 == cached) {
#pragma unroll
                for (int j = 0; j < 4; ++j) hA[j] = hB[j];
            } else {
                hrow(sy0, hA);
            }
            if (sy1 != sy0) hrow(sy1, hB);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) hB[j] = hA[j];
            }
            cached = sy1;
            unsigned d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = (((b0 * hA[j]) >> 16) + ((b1 * hB[j]) >> 16) + 2u) >> 2;  // <= 255
            return ((d[3] * 256u + d[2]) * 256u + d[1]) * 256u + d[0];
        };
        uint8_t* o = dOut + (size_t)(y0 + EAOF_EDGE) * D.pitch;
        if (y0 > EAOF_EDGE && yEnd - 1 < D.h - 1 - EAOF_EDGE) {  // no row of this CTA is mirrored into the border (CTA-uniform)
            for (int y = y0; y < yEnd; ++y, o += D.pitch) *reinterpret_cast<uint32_t*>(o) = row(y);
        } else {
            for (int y = y0; y < yEnd; ++y, o += D.pitch) {
                const uint32_t v = row(y);
                *reinterpret_cast<uint32_t*>(o) = v;
                if (y >= 1 && y <= EAOF_EDGE) *reinterpret_cast<uint32_t*>(dOut + (size_t)(EAOF_EDGE - y) * D.pitch) = v;
                if (y >= D.h - 1 - EAOF_EDGE && y <= D.h - 2)
                    *reinterpret_cast<uint32_t*>(dOut + (size_t)(EAOF_EDGE + 2 * (D.h - 1) - y) * D.pitch) = v;
            }
        }
    }
}

// ---- k_pyramid_fused: ComputePyramid in ONE launch ---------------------------------------------------------------
// src/ORBextractor.cc:1107-1132 is a dependent chain: level l is cv::resize of level l-1.  Instead of one launch per level
// (each a pass over global memory, each a launch on the critical path of a single frame), a CTA owns a tile of the LAST
// level and carries the part of the image that tile descends from down through all levels in shared memory:
//   * the tile plan is built on the host from the reference's own resize tables: own ranges [a, b) per level nest exactly
//     (a_{l-1} = sx_l(a_l)), so every pixel of every level is written by exactly one CTA; what a CTA needs beyond its own
//     range is a halo on the right / bottom only (1 px at the last level, growing by the scale factor per level: 17 px at
//     level 0 of an 8-level 1.2 pyramid), recomputed, never exchanged;
//   * the level-0 region arrives with one cp.async.bulk.tensor box of the input frames' {width, height, frames} tensor
//     (box start rounded down to a 16-byte column, TMA's rule) completing on an mbarrier; inputs TMA cannot take
//     (row stride or base not multiples of 16) are fetched with plain loads into the same layout;
//   * per level: thread = (destination column, chunk of rows); the horizontal pass of the two source rows lives in
//     registers (the lower row of one output row is usually the upper row of the next), vertical pass and rounding exactly
//     as k_resize (11-bit coefficients, (b*(h>>4))>>16 per term, +2 >> 2);
//   * every level is written to the bordered pyramid from shared memory: own pixels, and for tiles on the image edge their
//     REFLECT_101 mirror images (the source of a border pixel within 19 px of the edge always lies in the edge tile).
// Tiles are stored with their x origin rounded down to a multiple of 4, which makes a shared-memory word and the global
// word it goes to share their alignment (global inner pixel x sits at byte 32 + x).
struct FusedAxis { short a, b, lo, n; };  // per (level, tile index): own [a, b), needed [lo, lo + n)
struct FusedArgs {
    const FusedAxis* ax;   // [nlevels][nTx]
    const FusedAxis* ay;   // [nlevels][nTy]
    int nTx, nTy;
    int pitchT[EAOF_MAX_LEVELS];  // shared-memory row pitch of a tile of level l (bytes, multiple of 16)
    int bufBytes[2];       // even levels / odd levels
    int boxW0, boxH0;      // TMA box of level 0 (= pitchT[0] x rows)
    int useTma;
    const uint8_t* in;     // input frames (device)
    size_t stride, framePitch;
    int f0;                // first frame of this launch inside the handle's pyramid buffer
};

#define EAOF_FUSED_MAX_ROWS 160
__global__ void __launch_bounds__(512) k_pyramid_fused(const __grid_constant__ FastTmaMaps inMap, const FusedArgs A,
                                                       uint8_t* __restrict__ pyr, const int* __restrict__ tabs,
                                                       const __grid_constant__ Geom g) {
    extern __shared__ __align__(128) uint8_t fusedSmem[];
    __shared__ __align__(8) unsigned long long fusedBar;
    __shared__ int4 fusedRows[EAOF_FUSED_MAX_ROWS];  // per destination row of the level in flight: (source row 0, source row 1, b0|b1<<16)
    const int tid = threadIdx.x, nThreads = blockDim.x, lane = tid & 31, warp = tid >> 5, nWarps = nThreads >> 5;
    const int ty = blockIdx.x / A.nTx, tx = blockIdx.x - ty * A.nTx;
    const int f = blockIdx.y;
    uint8_t* const buf0 = fusedSmem + ((128u - (smem_u32(fusedSmem) & 127u)) & 127u);  // even levels; odd levels behind it
    auto buf = [&](int parity) { return buf0 + (parity ? A.bufBytes[0] : 0); };
    uint8_t* frame = pyr + (size_t)(A.f0 + f) * g.pyrFrameBytes;

    // writes level l from its shared-memory tile: own pixels + mirrored border.  Own ranges start on multiples of 4 in x
    // (host plan), so all words but the ones touching the image border are plain word copies (pass 1, no divergence);
    // the border words of edge tiles go byte by byte through the reflection (pass 2, lanes = those words only).
    auto write_level = [&](int l, const uint8_t* tile, int pitch, int ox, int oy) {
        const LevelGeom& L = g.L[l];
        const FusedAxis X = A.ax[l * A.nTx + tx], Y = A.ay[l * A.nTy + ty];
        const int bx0 = X.a == 0 ? -EAOF_EDGE : X.a, bx1 = X.b == L.w ? L.w + EAOF_EDGE : X.b;   // bordered own span, inner coordinates
        const int by0 = Y.a == 0 ? -EAOF_EDGE : Y.a, by1 = Y.b == L.h ? L.h + EAOF_EDGE : Y.b;
        const int nRows = by1 - by0;
        uint8_t* out = frame + L.off + EAOF_INNER_X0;                   // inner pixel (x, bordered row r) at out[r*pitch + x]
        auto refl = [](int p, int len) { return p < 0 ? -p : p >= len ? 2 * (len - 1) - p : p; };  // one fold: levels are >= 40 px
        // pass 1: whole words inside both the own range and the image
        const int fx0 = (max(bx0, 0) + 3) & ~3, fx1 = min(bx1, L.w) & ~3;
        const int nFast = max(fx1 - fx0, 0) >> 2;
        for (int r = warp; r < nRows; r += nWarps) {
            const int by = by0 + r;
            const uint32_t* srow = reinterpret_cast<const uint32_t*>(tile + (refl(by, L.h) - oy) * pitch + (fx0 - ox));
            uint32_t* drow = reinterpret_cast<uint32_t*>(out + (size_t)(by + EAOF_EDGE) * L.pitch + fx0);
            for (int wi = lane; wi < nFast; wi += 32) drow[wi] = srow[wi];
        }
        // pass 2: the bytes left and right of the whole words (image border columns, and a ragged last word)
        const int nLeft = fx0 - bx0, nRight = bx1 - max(fx1, fx0);
        const int nSlow = nLeft + nRight;
        if (nSlow > 0) {
            for (int i = tid; i < nSlow * nRows; i += nThreads) {
                const int r = i / nSlow, k = i - r * nSlow;
                const int xx = k < nLeft ? bx0 + k : max(fx1, fx0) + (k - nLeft);
                const int by = by0 + r;
                out[(size_t)(by + EAOF_EDGE) * L.pitch + xx] = tile[(refl(by, L.h) - oy) * pitch + refl(xx, L.w) - ox];
            }
        }
    };

    // ---- level 0: the needed region of the input frame
    const FusedAxis X0 = A.ax[tx], Y0 = A.ay[ty];
    int ox = X0.lo & ~15, oy = Y0.lo;    // tile origin (x rounded down: TMA's 16-byte rule, and word alignment)
    {
        uint8_t* t0 = buf(0);
        const int pitch0 = A.pitchT[0];
        if (A.useTma) {
            const uint32_t bar = smem_u32(&fusedBar);
            if (tid == 0) {
                mbar_init(bar, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, (uint32_t)(A.boxW0 * A.boxH0));
                tma_load_3d(smem_u32(t0), &inMap.m[0][0], ox, oy, f, bar);
            }
            __syncthreads();
            mbar_wait(bar, 0);
        } else {
            const uint8_t* src = A.in + (size_t)f * A.framePitch;
            const int w = g.L[0].w, h = g.L[0].h;
            const int nx = min(X0.lo + X0.n, w) - ox, ny = min(Y0.lo + Y0.n, h) - oy;
            for (int r = warp; r < ny; r += nWarps)
                for (int cx = lane; cx < nx; cx += 32) t0[r * pitch0 + cx] = __ldg(src + (size_t)(oy + r) * A.stride + ox + cx);
            __syncthreads();
        }
        write_level(0, t0, pitch0, ox, oy);
    }

    // ---- levels 1 .. n-1
    for (int l = 1; l < g.nlevels; ++l) {
        const LevelGeom& D = g.L[l];
        const LevelGeom& S = g.L[l - 1];
        const FusedAxis X = A.ax[l * A.nTx + tx], Y = A.ay[l * A.nTy + ty];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(buf((l - 1) & 1));
        uint8_t* dst = buf(l & 1);
        const int pitchSW = A.pitchT[l - 1] >> 2, pitchD = A.pitchT[l];
        const int oxD = X.lo & ~3, oyD = Y.lo;
        const int nRows = Y.n;
        // rows of this tile: word offsets of the two source rows inside the source tile and the packed (b0, b1)
        for (int r = tid; r < nRows; r += nThreads) {
            const int y = Y.lo + r;
            const int sy = __ldg(tabs + D.yTab + 2 * y);
            const int sy0 = min(max(sy, 0), S.h - 1), sy1 = min(max(sy + 1, 0), S.h - 1);
            fusedRows[r] = make_int4((sy0 - oy) * pitchSW, (sy1 - oy) * pitchSW, __ldg(tabs + D.yTab + 2 * y + 1), 0);
        }
        __syncthreads();
        // work item = (word of 4 destination columns, chunk of rows); per-thread constants: source word offsets, byte
        // selectors and packed (a0, a1) of its 4 columns — the arithmetic of k_resize on shared memory
        const int nWordsD = (X.lo + X.n - oxD + 3) >> 2;
        int chunks = nThreads / nWordsD;
        if (chunks < 1) chunks = 1;
        if (chunks > nRows) chunks = nRows;
        const int rowsPer = (nRows + chunks - 1) / chunks;
        for (int task = tid; task < nWordsD * chunks; task += nThreads) {
            const int ch = task / nWordsD, wd = task - ch * nWordsD;
            int wofs[4];
            unsigned sel[4], coef[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = min(max(oxD + 4 * wd + j, (int)X.lo), X.lo + X.n - 1);  // columns outside the needed range repeat its edge
                const int sx = __ldg(tabs + D.xTab + 2 * x) - ox;
                coef[j] = (unsigned)__ldg(tabs + D.xTab + 2 * x + 1);
                wofs[j] = sx >> 2;
                sel[j] = (unsigned)(sx & 3) | ((unsigned)((sx & 3) + 1) << 4);  // bytes sx, sx+1 of the word pair; a1 == 0
            }                                                                    // whenever sx+1 would leave the image
            auto hrow = [&](int rowW, unsigned (&h)[4]) {
                const uint32_t* r = src + rowW;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    h[j] = __dp2a_lo(coef[j], __byte_perm(r[wofs[j]], r[wofs[j] + 1], sel[j]), 0u) >> 4;  // every use is (sum >> 4), A.2
            };
            const int r0 = ch * rowsPer, r1 = min(r0 + rowsPer, nRows);
            int cached = -1;
            unsigned hB[4] = {0, 0, 0, 0};
            uint32_t* dcol = reinterpret_cast<uint32_t*>(dst) + wd;
            const int pitchDW = pitchD >> 2;
            for (int r = r0; r < r1; ++r) {
                const int4 rw = fusedRows[r];
                const unsigned b0s = (unsigned)rw.z << 16, b1s = (unsigned)rw.z & 0xffff0000u;
                unsigned hA[4];
                if (rw.x == cached) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) hA[j] = hB[j];
                } else {
                    hrow(rw.x, hA);
                }
                if (rw.y != rw.x) hrow(rw.y, hB);
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) hB[j] = hA[j];
                }
                cached = rw.y;
                unsigned d[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) d[j] = mad_hi_u32(b1s, hB[j], mad_hi_u32(b0s, hA[j], 2u)) >> 2;  // <= 255
                dcol[r * pitchDW] = ((d[3] * 256u + d[2]) * 256u + d[1]) * 256u + d[0];
            }
        }
        __syncthreads();
        write_level(l, dst, pitchD, oxD, oyD);
        ox = oxD;
        oy = oyD;
        // the tile just written becomes the source of the next level; its buffer is not touched again until the level
        // after that is computed, by which time every thread has passed the barriers above once more
    }
}

// ------------------------------------------------------------------------------------------------
// FAST-9/16.  Ring offsets k=0..15: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)
// (-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3).
#define FAST_TP 72  // smem tile pitch (bytes)
#define RING_OFF(k, P)                                                                                               \
    ((k) == 0 ? 3 * (P) : (k) == 1 ? 3 * (P) + 1 : (k) == 2 ? 2 * (P) + 2 : (k) == 3 ? (P) + 3 : (k) == 4 ? 3           \
     : (k) == 5 ? -(P) + 3 : (k) == 6 ? -2 * (P) + 2 : (k) == 7 ? -3 * (P) + 1 : (k) == 8 ? -3 * (P)                    \
     : (k) == 9 ? -3 * (P)-1 : (k) == 10 ? -2 * (P)-2 : (k) == 11 ? -(P)-3 : (k) == 12 ? -3                             \
     : (k) == 13 ? (P)-3 : (k) == 14 ? 2 * (P)-2 : 3 * (P)-1)

// max over the 16 contiguous 9-arcs of min(ring - v) (brighter) and of min(v - ring) (darker).
// corner at threshold t  <=>  result > t ;  OpenCV's cornerScore<16> == result - 1 for a corner.
// Both polarities are evaluated at once on packed s16x2 lanes (low half d = ring - v, high half -d) with the
// DPX min/max instructions; sliding minima: m2 covers 2 ring pixels, m4 covers 4, m9 = min3(m4[k], m4[k+4], w[k+8]).
// NOTE: a plain-int formulation of the same min/max tree (scalar min()/max() on int arrays) is MISCOMPILED by
// nvcc 12.9 for sm_100a (wrong VIMNMX3 fusion; 97% of results wrong on a B200, identical source is right on the
// host) — scratch repro kept as tests/cuda/arc_best_miscompile.cu; tests/test_gpu_extract.py::test_stages pins this.
__device__ __forceinline__ int arc_best(const uint8_t* p) {
    const int v = p[0];
    unsigned w[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int d = (int)p[RING_OFF(k, FAST_TP)] - v;
        w[k] = ((unsigned)d & 0xffffu) | ((unsigned)(-d) << 16);
    }
    unsigned m2[16], m4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m2[k] = __vmins2(w[k], w[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m4[k] = __vmins2(m2[k], m2[(k + 2) & 15]);
    unsigned best = 0x80008000u;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        best = __vmaxs2(best, __vimin3_s16x2(m4[k], m4[(k + 4) & 15], w[(k + 8) & 15]));
    const int lo = (short)(best & 0xffff), hi = (short)(best >> 16);
    return lo > hi ? lo : hi;
}

// ---- packed 16x2 helpers: one 32-bit register carries the same quantity for two pixels, and VIMNMX(3).U16x2
// works on both halves in one issue slot.  Differences ring - centre are kept biased by +256 (range 1..511) so that
// they are formed by a plain 32-bit add (no borrow can cross the lanes) and compare as unsigned.
#define FAST_BIAS2 0x01000100u

// Arc score of two pixels at once.  d[k] = 256 + ring_k - centre on u16x2 lanes.  Returns 256 + max(bright, dark)
// where bright = max over the 16 arcs of min(ring - centre over 9 contiguous) and dark = max over arcs of
// min(centre - ring) = -(min over arcs of max(ring - centre)).  Sliding 9-window = min3 of three 3-windows.
__device__ __forceinline__ unsigned arc_best2(const unsigned (&d)[16]) {
    unsigned lo3[16], hi3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        lo3[k] = __vimin3_u16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        hi3[k] = __vimax3_u16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    }
    unsigned best = 0u, worst = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned a = __vimin3_u16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
        const unsigned b = __vimin3_u16x2(lo3[k + 1], lo3[(k + 4) & 15], lo3[(k + 7) & 15]);
        best = __vimax3_u16x2(best, a, b);
        const unsigned c = __vimax3_u16x2(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
        const unsigned e = __vimax3_u16x2(hi3[k + 1], hi3[(k + 4) & 15], hi3[(k + 7) & 15]);
        worst = __vimin3_u16x2(worst, c, e);
    }
    return __vmaxu2(best, 2u * FAST_BIAS2 - worst);  // worst <= 511 per lane: no borrow
}

// One WARP per cell (no CTA-wide barrier anywhere: the warps of a CTA only share the launch).  Candidates are
// appended to the (frame, level) list in arbitrary order; DistributeOctTree only needs their cell-raster rank for
// tie-breaking, which k_octree recomputes from the coordinates.
// cand word: x | y<<12 | score<<24 (detection-window coordinates, src/ORBextractor.cc:822-824).
// Shared memory per warp (sized for the largest cell of the handle, Geom::fast*): the cell tile, the score map in
// the same layout, and the list of surviving pixels.
// One warp per CTA: cells differ widely in cost (flat cells stop after the pre-test, busy ones score hundreds of pixels,
// empty ones run twice), and the block scheduler balances single warps far better than groups of four that retire
// together — 0.680 -> 0.602 ms per 250 frames (4 warps x 8 CTAs and 2 x 16 measured the same as each other).
#ifndef FAST_WARPS
#define FAST_WARPS 1
#endif
#ifndef FAST_MINB
#define FAST_MINB 24
#endif
#define FAST_CLST 128
// Phases (A), (B), (C) of one cell whose tile already sits in shared memory: pixel (row r, byte column c) at byte
// (r*PW + PAD)*4 + c of `tile`, cell pixel x at byte column mis + x; Bm (zeroed by the caller) has the same layout.
template <int PAD>
__device__ __forceinline__ void fast_cell(uint32_t* tile, uint32_t* Bm, uint16_t* clst, uint16_t* lst, const int lstCap,
                                          const CellDesc c, const int f, const int mis, const int PW, const int lane,
                                          uint32_t* __restrict__ cand, uint32_t* __restrict__ candCount, const Geom& g) {
    // NMS survivors overwrite the tile: the first one is written only when the attempt that produced it is the last
    // one, and phase (C) reads nothing but the score map.
    uint32_t* outl = tile;
    const LevelGeom& L = g.L[c.level];
    const int cw = c.cw, ch = c.ch;
    const int ih = ch - 6;
    // words holding at least one pixel of the cell's inner area x in [3, cw-3)
    const int cLo = mis + 3, cHi = mis + cw - 4;  // first / last valid tile byte column
    const int wLo = cLo >> 2, nW = (cHi >> 2) - wLo + 1;
    const int nTasks = ih * nW;
    const unsigned rcpW = 65536u / (unsigned)nW + 1u;
    const unsigned below = (1u << lane) - 1;
    
    // The cell is first searched at iniThFAST; only a cell with no keypoint after NMS is searched again at
    // minThFAST (:808-816).  Each attempt: (A) a necessary condition on the 4 compass ring pixels (every 9-arc holds
    // one of the pixels {0, 8} and one of {4, 12}) at that threshold, one lane per aligned 4-pixel word on byte lanes,
    // the surviving pixels compacted with warp ballots; (B) exact arc score of the survivors, any two of them per lane
    // on u16x2; (C) 8-neighbour NMS restricted to the cell.  Arc scores are threshold independent, so what the first attempt
    // wrote into the score map stays valid for the second.
    int no = 0;
    for (int attempt = 0; attempt < 2 && no == 0; ++attempt) {
        const int th = attempt == 0 ? g.iniTh : g.minTh;
            const bool thHigh = th >= 128;
        const unsigned thK = (unsigned)(127 - (th & 127)) * 0x01010101u;
        // (A) and (B) alternate in rounds: (A) appends the surviving PIXELS (row << 8 | tile byte column) until the list
        // could overflow, (B) drains it two pixels per lane.  One round for all but noise-like cells.
        int ncorn = 0;  // corners found by (B); the first FAST_CLST of them are listed for (C)
        const uint8_t* tb = reinterpret_cast<const uint8_t*>(tile);
        uint8_t* mb = reinterpret_cast<uint8_t*>(Bm);
        const int pitchB = 4 * PW;
        for (int i0 = 0; i0 < nTasks;) {
            // ---- (A)
            int nl = 0;
            for (; i0 < nTasks && nl + 128 <= lstCap; i0 += 32) {
                const int i = i0 + lane;
                unsigned m = 0;
                int y = 0, w = 0;
                if (i < nTasks) {
                    const int r = (int)(((unsigned)i * rcpW) >> 16);
                    w = wLo + (i - r * nW);
                    y = r + 3;
                    const uint32_t* t = tile + y * PW + PAD + w;
                    const unsigned W0 = t[0], Wm = t[-1], Wp = t[1], Wu = t[-3 * PW], Wd = t[3 * PW];
                    const unsigned V4 = __byte_perm(W0, Wp, 0x6543), V12 = __byte_perm(Wm, W0, 0x4321);
                    // Every 9-arc of the ring holds one of the pixels {0, 8} and one of {4, 12}, so a corner at threshold
                    // th has |ring - centre| > th on one pixel of each pair.  Four pixels at once: VABSDIFF4 against the
                    // words 3 rows below / above and 3 columns right / left, "byte > th" as a carry into bit 7 of each
                    // byte (low 7 bits + 127 - th', combined with the byte's own top bit: OR for th < 128, AND above).
                    const unsigned a0 = __vabsdiffu4(W0, Wd), a8 = __vabsdiffu4(W0, Wu);
                    const unsigned a4 = __vabsdiffu4(W0, V4), a12 = __vabsdiffu4(W0, V12);
                    const unsigned t0 = (a0 & 0x7f7f7f7fu) + thK, t8 = (a8 & 0x7f7f7f7fu) + thK;
                    const unsigned t4 = (a4 & 0x7f7f7f7fu) + thK, t12 = (a12 & 0x7f7f7f7fu) + thK;
                    m = thHigh ? ((t0 & a0) | (t8 & a8)) & ((t4 & a4) | (t12 & a12))
                               : ((t0 | a0) | (t8 | a8)) & ((t4 | a4) | (t12 | a12));
                    m &= 0x80808080u;  // bit 7 of byte j: pixel j of the word survives
                }
                const unsigned e = (unsigned)((y << 8) | (w << 2));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool pj = (m >> (8 * j + 7)) & 1u;
                    const unsigned mj = __ballot_sync(0xffffffffu, pj);
                    if (pj) lst[nl + __popc(mj & below)] = (uint16_t)(e | (unsigned)j);
                    nl += __popc(mj);
                }
            }
            if (nl == 0) continue;
            __syncwarp();
            // ---- (B): lane takes two list entries; their ring pixels are read as bytes (no lane extraction) and packed
            // into u16x2 for the arc score
            for (int p0 = 0; p0 < nl; p0 += 64) {
                const int iA = p0 + 2 * lane;
                const bool actA = iA < nl, actB = iA + 1 < nl;
                const int eA = lst[actA ? iA : 0], eB = lst[actB ? iA + 1 : (actA ? iA : 0)];
                const int oA = ((eA >> 8) * PW + PAD) * 4 + (eA & 255), oB = ((eB >> 8) * PW + PAD) * 4 + (eB & 255);
                const uint8_t* pA = tb + oA;
                const uint8_t* pB = tb + oB;
                const unsigned nc = FAST_BIAS2 - ((unsigned)pA[0] | ((unsigned)pB[0] << 16));
                unsigned d[16];
#define RING2(k, off) d[k] = ((unsigned)pA[off] | ((unsigned)pB[off] << 16)) + nc;
                RING2(0, 3 * pitchB) RING2(1, 3 * pitchB + 1) RING2(2, 2 * pitchB + 2) RING2(3, pitchB + 3)
                RING2(4, 3) RING2(5, -pitchB + 3) RING2(6, -2 * pitchB + 2) RING2(7, -3 * pitchB + 1)
                RING2(8, -3 * pitchB) RING2(9, -3 * pitchB - 1) RING2(10, -2 * pitchB - 2) RING2(11, -pitchB - 3)
                RING2(12, -3) RING2(13, pitchB - 3) RING2(14, 2 * pitchB - 2) RING2(15, 3 * pitchB - 1)
#undef RING2
                const unsigned b2 = arc_best2(d);
                // a pixel that is not a corner at this threshold, or lies outside the inner area of the cell (cv::FAST
                // never scores those), keeps 0: the NMS only ever asks "corner ? score : 0"
                const int cA = eA & 255, cB = eB & 255;
                const int bLo = (int)(b2 & 0xffffu) - 256, bHi = (int)(b2 >> 16) - 256;
                const bool k0 = actA && bLo > th && cA >= cLo && cA <= cHi;
                const bool k2 = actB && bHi > th && cB >= cLo && cB <= cHi;
                if (k0) mb[oA] = (uint8_t)bLo;
                if (k2) mb[oB] = (uint8_t)bHi;
                const unsigned m0 = __ballot_sync(0xffffffffu, k0), m2 = __ballot_sync(0xffffffffu, k2);
                const int p0c = ncorn + __popc(m0 & below), p2c = ncorn + __popc(m0) + __popc(m2 & below);
                if (k0 && p0c < FAST_CLST) clst[p0c] = (uint16_t)eA;
                if (k2 && p2c < FAST_CLST) clst[p2c] = (uint16_t)eB;
                ncorn += __popc(m0) + __popc(m2);
            }
            __syncwarp();
        }
        if (ncorn == 0) continue;
        // ---- (C)
        if (ncorn <= FAST_CLST) {
            for (int i0 = 0; i0 < ncorn; i0 += 32) {
                const bool act = i0 + lane < ncorn;
                const int e = clst[act ? i0 + lane : 0], y = e >> 8, col = e & 255;
                const uint8_t* q = reinterpret_cast<const uint8_t*>(Bm + y * PW + PAD) + col;
                const int s = (int)q[0] - 1;
                int nbMax = 0;  // stored scores are either 0 or > th
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (dx == 0 && dy == 0) continue;
                        nbMax = max(nbMax, (int)q[dy * pitchB + dx]);
                    }
                const bool keep = act && s > (nbMax > 0 ? nbMax - 1 : 0);  // s > (neighbour corner ? its score : 0), all 8
                const unsigned mk = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int x = col - mis;  // cell coordinates
                    const int wx = x + c.iniX - EAOF_MIN_BORDER, wy = y + c.iniY - EAOF_MIN_BORDER;
                    outl[no + __popc(mk & below)] = (uint32_t)wx | ((uint32_t)wy << 12) | ((uint32_t)s << 24);
                }
                no += __popc(mk);
            }
        } else  // more corners than the list holds (noise-like cells): scan the score map of the inner area
        for (int i0 = 0; i0 < nTasks; i0 += 32) {
            const int i = i0 + lane;
            const int r = (int)(((unsigned)min(i, nTasks - 1) * rcpW) >> 16);
            const int w = wLo + (min(i, nTasks - 1) - r * nW), y = r + 3;
            const uint8_t* q0 = reinterpret_cast<const uint8_t*>(Bm + y * PW + PAD + w);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint8_t* q = q0 + j;
                const int s = (i < nTasks ? (int)q[0] : 0) - 1;
                bool keep = false;
                if (s >= 0) {
                    int nbMax = 0;  // stored scores are either 0 or > th
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            if (dx == 0 && dy == 0) continue;
                            nbMax = max(nbMax, (int)q[dy * pitchB + dx]);
                        }
                    keep = s > (nbMax > 0 ? nbMax - 1 : 0);  // s > (neighbour is a corner ? its score : 0) for all 8
                }
                const unsigned mk = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int x = 4 * w + j - mis;  // cell coordinates
                    const int wx = x + c.iniX - EAOF_MIN_BORDER, wy = y + c.iniY - EAOF_MIN_BORDER;
                    outl[no + __popc(mk & below)] = (uint32_t)wx | ((uint32_t)wy << 12) | ((uint32_t)s << 24);
                }
                no += __popc(mk);
            }
        }
        __syncwarp();
    }
    if (no == 0) return;
    uint32_t gBase = 0;
    if (lane == 0) gBase = atomicAdd(&candCount[f * g.nlevels + c.level], (uint32_t)no);
    gBase = __shfl_sync(0xffffffffu, gBase, 0);
    uint32_t* dst = cand + (size_t)f * g.candPerFrame + L.candOff + gBase;
    for (int i = lane; i < no; i += 32) dst[i] = outl[i];
}

// Arc score of two pixels from RAW ring values: r[k] = ring_k of pixel A | ring_k of pixel B << 16, cc = the two centres
// packed the same way.  max over arcs of min(ring) - centre (brighter) and centre - min over arcs of max(ring) (darker)
// are translation invariant, so the centre is subtracted once at the end instead of from each of the 16 ring values.
// Returns 256 + max(bright, dark) per lane like arc_best2.
__device__ __forceinline__ unsigned arc_best2_raw(const unsigned (&r)[16], const unsigned cc) {
    unsigned lo3[16], hi3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        lo3[k] = __vimin3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
        hi3[k] = __vimax3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    }
    unsigned best = 0u, worst = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned a = __vimin3_u16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
        const unsigned b = __vimin3_u16x2(lo3[k + 1], lo3[(k + 4) & 15], lo3[(k + 7) & 15]);
        best = __vimax3_u16x2(best, a, b);
        const unsigned c = __vimax3_u16x2(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
        const unsigned e = __vimax3_u16x2(hi3[k + 1], hi3[(k + 4) & 15], hi3[(k + 7) & 15]);
        worst = __vimin3_u16x2(worst, c, e);
    }
    // every lane value is in 0..255: 256 + best - c and 256 + c - worst stay in 1..511, no borrow or carry crosses the lanes
    return __vmaxu2(best + (FAST_BIAS2 - cc), (FAST_BIAS2 + cc) - worst);
}

// ---- fast_cell_rows: the same search with LANE = ROW in the pre-test -------------------------------------------------
// Phase (A) of fast_cell spends more issue slots on bookkeeping than on the test: a task index -> (row, word) division,
// five shared-memory loads per word, and four ballots + prefix + store per word to compact the survivors.  Here a lane
// walks one row of the cell from left to right instead: the word to the right becomes the centre word of the next step
// (three loads per word), and the survivors of the row are collected as a bit mask in two registers — word k of the row,
// pixel j -> bit 8j + (k & 7) of register k >> 3 — with one shift + OR per word and NO warp vote.  The odd tile pitch
// makes "same column, 32 consecutive rows" conflict-free.  One warp scan over the per-row counts then gives every lane its
// place in the survivor list, and a lane emits its own row's pixels (one FLO + a few ALU per survivor).  Columns outside
// the inner area of the cell are masked before emission, so phase (B) no longer scores pixels cv::FAST never looks at.
// The tile pitch PW (words) is a template parameter: every ring / neighbour offset is then an immediate of the load
// instead of a register (phase (B) alone formed 14 such addresses per pixel pair).
// The code is kept small on purpose — one instance of every phase serves both attempts and all row rounds: with one-warp
// CTAs at unrelated program counters an unrolled 6000-instruction version of this function spent a third of its stall
// samples waiting for instruction fetch.
// Requires nW <= 16 words per row and thresholds < 128: eaof_orb_create picks k_fast_generic otherwise.
#define FAST_CLST2 128  // corner list entries (u16: row << 8 | tile byte column)
template <int PW>
__device__ __forceinline__ void fast_cell_rows(uint32_t* tile, uint32_t* Bm, uint16_t* clst, uint16_t* lst, const int lstCap,
                                               const CellDesc c, const int f, const int mis, const int lane,
                                               uint32_t* __restrict__ cand, uint32_t* __restrict__ candCount, const Geom& g) {
    constexpr int PAD = 1;
    constexpr int pitchB = 4 * PW;
    uint32_t* outl = tile;
    const LevelGeom& L = g.L[c.level];
    const int cw = c.cw, ch = c.ch;
    const int ih = ch - 6;
    const int cLo = mis + 3, cHi = mis + cw - 4;  // first / last valid tile byte column
    const int wLo = cLo >> 2, nW = (cHi >> 2) - wLo + 1;
    const unsigned below = (1u << lane) - 1;
    const uint8_t* tb = reinterpret_cast<const uint8_t*>(tile);
    uint8_t* mb = reinterpret_cast<uint8_t*>(Bm);
    // valid-column masks in the row-mask layout (warp-uniform)
    unsigned vm0, vm1;
    {
        const int n0 = min(nW, 8), n1 = nW - 8;
        vm0 = 0x01010101u * ((1u << n0) - 1u);
        vm1 = n1 > 0 ? 0x01010101u * ((1u << n1) - 1u) : 0u;
        vm0 &= ~(((1u << (8 * (cLo & 3))) - 1u) & 0x01010101u);  // first word: pixels j < cLo & 3 lie left of the inner area
        const int kl = (nW - 1) & 7, jl = cHi & 3;               // last word: pixels j > cHi & 3 lie right of it
        const unsigned clr = jl < 3 ? (0x01010101u << kl) & ~((2u << (8 * jl + kl)) - 1u) : 0u;
        if (nW > 8) vm1 &= ~clr; else vm0 &= ~clr;
    }
    const int nGroups = (nW + 3) >> 2;

    // The cell is first searched at iniThFAST; only a cell with no keypoint after NMS is searched again at minThFAST
    // (:808-816).  Arc scores are threshold independent, so what the first attempt wrote into the score map stays valid.
    int no = 0;
#pragma unroll 1
    for (int attempt = 0; attempt < 2 && no == 0; ++attempt) {
        const int th = attempt == 0 ? g.iniTh : g.minTh;
        const unsigned thK = (unsigned)(127 - th) * 0x01010101u;
        // ---- (A) + emission + (B), 32 rows at a time
        int ncorn = 0;  // corners found by (B); the first FAST_CLST2 of them are listed for (C)
#pragma unroll 1
        for (int r0 = 0; r0 < ih; r0 += 32) {
            const int r = r0 + lane;
            unsigned acc0 = 0u, acc1 = 0u;
            if (r < ih) {
                // words beyond nW (the unrolled groups of four) read the rest of the tile row or the row below: masked by vm
                const uint32_t* t = tile + (r + 3) * PW + PAD + wLo;
                unsigned Wm = t[-1], W0 = t[0];
                unsigned Um = t[-1 - 2 * PW], U0 = t[-2 * PW], Dm = t[-1 + 2 * PW], D0 = t[2 * PW];  // rows y -+ 2 (second attempt)
#pragma unroll 1
                for (int gi = 0; gi < nGroups; ++gi, t += 4) {
                    unsigned x = 0u;
                    unsigned Wc[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        Wc[i] = W0;
                        const unsigned Wp = t[i + 1], Wu = t[i - 3 * PW], Wd = t[i + 3 * PW];
                        const unsigned V4 = __byte_perm(W0, Wp, 0x6543), V12 = __byte_perm(Wm, W0, 0x4321);
                        // Every 9-arc of the ring holds one of the pixels {0, 8} and one of {4, 12}: a corner at threshold th
                        // has |ring - centre| > th on one pixel of each pair.  "byte > th" = carry into bit 7 of
                        // (low 7 bits + 127 - th), or the byte's own top bit.
                        const unsigned a0 = __vabsdiffu4(W0, Wd), a8 = __vabsdiffu4(W0, Wu);
                        const unsigned a4 = __vabsdiffu4(W0, V4), a12 = __vabsdiffu4(W0, V12);
                        const unsigned t0 = (a0 & 0x7f7f7f7fu) + thK, t8 = (a8 & 0x7f7f7f7fu) + thK;
                        const unsigned t4 = (a4 & 0x7f7f7f7fu) + thK, t12 = (a12 & 0x7f7f7f7fu) + thK;
                        const unsigned m = ((t0 | a0) | (t8 | a8)) & ((t4 | a4) | (t12 | a12)) & 0x80808080u;
                        x |= m >> (3 - i);  // word 4 gi + i -> bit 4 + i of every byte
                        Wm = W0;
                        W0 = Wp;
                    }
                    if (attempt) {
                        // The minThFAST attempt lets a third of all pixels through the compass test, and each of them costs a
                        // full arc score.  The same argument holds for the diagonal ring pixels: every 9-arc holds one of
                        // {2, 10} and one of {6, 14}.  max(a, b) > th is tested on a | b >= max(a, b): conservative for any th
                        // (exact when th + 1 is a power of two, e.g. the default 7), and a pre-test only has to be necessary.
                        unsigned y = 0u;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const unsigned Up = t[i + 1 - 2 * PW], Dp = t[i + 1 + 2 * PW];
                            const unsigned a2 = __vabsdiffu4(Wc[i], __byte_perm(D0, Dp, 0x5432));
                            const unsigned a14 = __vabsdiffu4(Wc[i], __byte_perm(Dm, D0, 0x5432));
                            const unsigned a6 = __vabsdiffu4(Wc[i], __byte_perm(U0, Up, 0x5432));
                            const unsigned a10 = __vabsdiffu4(Wc[i], __byte_perm(Um, U0, 0x5432));
                            const unsigned o1 = a2 | a10, o2 = a6 | a14;
                            const unsigned u1 = (o1 & 0x7f7f7f7fu) + thK, u2 = (o2 & 0x7f7f7f7fu) + thK;
                            const unsigned m = (u1 | o1) & (u2 | o2) & 0x80808080u;
                            y |= m >> (3 - i);
                            Um = U0; U0 = Up;
                            Dm = D0; D0 = Dp;
                        }
                        x &= y;
                    }
                    x = (gi & 1) ? x : x >> 4;
                    if (gi < 2) acc0 |= x; else acc1 |= x;
                }
                acc0 &= vm0;
                acc1 &= vm1;
            }
            const int cnt = __popc(acc0) + __popc(acc1);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31), lowHalf = __shfl_sync(0xffffffffu, incl, 15);
            if (total == 0) continue;
            // The survivor list holds 16 rows' worth of pixels (it is the largest per-warp array and decides how many warps an
            // SM keeps resident): a round whose 32 rows have more survivors than that — noise-like cells — is listed and
            // scored in two halves.
            const int nsub = total > lstCap ? 2 : 1;
#pragma unroll 1
            for (int sub = 0; sub < nsub; ++sub) {
                const bool mineNow = nsub == 1 || (lane >> 4) == sub;
                const int nl = nsub == 1 ? total : (sub ? total - lowHalf : lowHalf);
                if (mineNow) {
                    uint16_t* p = lst + (incl - cnt) - (sub ? lowHalf : 0);
                    const unsigned e0 = ((unsigned)(r + 3) << 8) | ((unsigned)wLo << 2);
#pragma unroll 1
                    for (int q = 0; q < 2; ++q) {
                        unsigned m = q ? acc1 : acc0;
                        const unsigned eq = e0 + 32u * q;
                        while (m) {  // two survivors per trip
                            unsigned b, b2;
                            asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(m));
                            m ^= 1u << b;
                            p[0] = (uint16_t)(eq + ((b & 7u) << 2) + (b >> 3));
                            const bool more = m != 0u;
                            asm("bfind.u32 %0, %1;" : "=r"(b2) : "r"(m));
                            if (more) {
                                m ^= 1u << b2;
                                p[1] = (uint16_t)(eq + ((b2 & 7u) << 2) + (b2 >> 3));
                            }
                            p += more ? 2 : 1;
                        }
                    }
                }
                __syncwarp();
                // ---- (B): two list entries per lane on u16x2, ring pixels read as bytes
                #pragma unroll 1
                for (int p0 = 0; p0 < nl; p0 += 64) {
                    const int iA = p0 + 2 * lane;
                    const bool actA = iA < nl, actB = iA + 1 < nl;
                    const int eA = lst[actA ? iA : 0], eB = lst[actB ? iA + 1 : (actA ? iA : 0)];
                    const int oA = ((eA >> 8) * PW + PAD) * 4 + (eA & 255), oB = ((eB >> 8) * PW + PAD) * 4 + (eB & 255);
                    const uint8_t* pA = tb + oA;
                    const uint8_t* pB = tb + oB;
                    const unsigned cc = (unsigned)pA[0] | ((unsigned)pB[0] << 16);
                    unsigned d[16];
        #define RING2(k, off) d[k] = (unsigned)pA[off] | ((unsigned)pB[off] << 16);
                    RING2(0, 3 * pitchB) RING2(1, 3 * pitchB + 1) RING2(2, 2 * pitchB + 2) RING2(3, pitchB + 3)
                    RING2(4, 3) RING2(5, -pitchB + 3) RING2(6, -2 * pitchB + 2) RING2(7, -3 * pitchB + 1)
                    RING2(8, -3 * pitchB) RING2(9, -3 * pitchB - 1) RING2(10, -2 * pitchB - 2) RING2(11, -pitchB - 3)
                    RING2(12, -3) RING2(13, pitchB - 3) RING2(14, 2 * pitchB - 2) RING2(15, 3 * pitchB - 1)
        #undef RING2
                    const unsigned b2 = arc_best2_raw(d, cc);
                    const int bLo = (int)(b2 & 0xffffu) - 256, bHi = (int)(b2 >> 16) - 256;
                    const bool k0 = actA && bLo > th, k2 = actB && bHi > th;
                    if (k0) mb[oA] = (uint8_t)bLo;
                    if (k2) mb[oB] = (uint8_t)bHi;
                    const unsigned m0 = __ballot_sync(0xffffffffu, k0), m2 = __ballot_sync(0xffffffffu, k2);
                    const int p0c = ncorn + __popc(m0 & below), p2c = ncorn + __popc(m0) + __popc(m2 & below);
                    if (k0 && p0c < FAST_CLST2) clst[p0c] = (uint16_t)eA;
                    if (k2 && p2c < FAST_CLST2) clst[p2c] = (uint16_t)eB;
                    ncorn += __popc(m0) + __popc(m2);
                }
                __syncwarp();  // the list is rewritten by the next sub-pass / round
            }
        }
        __syncwarp();  // the list is rewritten by the next attempt, the score map read by (C)
        if (ncorn == 0) continue;
        // ---- (C)
        if (ncorn <= FAST_CLST2) {
#pragma unroll 1
            for (int i0 = 0; i0 < ncorn; i0 += 32) {
                const bool act = i0 + lane < ncorn;
                const int e = clst[act ? i0 + lane : 0], y = e >> 8, col = e & 255;
                const uint8_t* q = mb + (y * PW + PAD) * 4 + col;
                const int s = (int)q[0] - 1;
                int nbMax = 0;  // stored scores are either 0 or > th
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (dx == 0 && dy == 0) continue;
                        nbMax = max(nbMax, (int)q[dy * pitchB + dx]);
                    }
                const bool keep = act && s > (nbMax > 0 ? nbMax - 1 : 0);  // s > (neighbour corner ? its score : 0), all 8
                const unsigned mk = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int x = col - mis;  // cell coordinates
                    const int wx = x + c.iniX - EAOF_MIN_BORDER, wy = y + c.iniY - EAOF_MIN_BORDER;
                    outl[no + __popc(mk & below)] = (uint32_t)wx | ((uint32_t)wy << 12) | ((uint32_t)s << 24);
                }
                no += __popc(mk);
            }
        } else {  // more corners than the list holds (noise-like cells): scan the score map of the inner area
            const int nTasks = ih * nW * 4;
            const unsigned rcpW = (1u << 20) / (unsigned)(4 * nW) + 1u;  // exact for i < 4096, divisor <= 64
#pragma unroll 1
            for (int i0 = 0; i0 < nTasks; i0 += 32) {
                const int i = min(i0 + lane, nTasks - 1);
                const int r = (int)(((unsigned)i * rcpW) >> 20);
                const int bc = 4 * wLo + (i - r * 4 * nW), y = r + 3;
                const uint8_t* q = mb + (y * PW + PAD) * 4 + bc;
                const int s = (i0 + lane < nTasks ? (int)q[0] : 0) - 1;
                bool keep = false;
                if (s >= 0) {
                    int nbMax = 0;
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            if (dx == 0 && dy == 0) continue;
                            nbMax = max(nbMax, (int)q[dy * pitchB + dx]);
                        }
                    keep = s > (nbMax > 0 ? nbMax - 1 : 0);
                }
                const unsigned mk = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int wx = bc - mis + c.iniX - EAOF_MIN_BORDER, wy = y + c.iniY - EAOF_MIN_BORDER;
                    outl[no + __popc(mk & below)] = (uint32_t)wx | ((uint32_t)wy << 12) | ((uint32_t)s << 24);
                }
                no += __popc(mk);
            }
        }
        __syncwarp();
    }
    if (no == 0) return;
    uint32_t gBase = 0;
    if (lane == 0) gBase = atomicAdd(&candCount[f * g.nlevels + c.level], (uint32_t)no);
    gBase = __shfl_sync(0xffffffffu, gBase, 0);
    uint32_t* dst = cand + (size_t)f * g.candPerFrame + L.candOff + gBase;
    for (int i = lane; i < no; i += 32) dst[i] = outl[i];
}

// Tile of one cell -> shared memory (pitch PW words, one pad word on the left), score map cleared (one warp).
__device__ __forceinline__ void fast_stage_tile(const uint8_t* __restrict__ pyr, uint32_t* tile, uint32_t* Bm, const CellDesc c,
                                                const int f, const int lane, const int PW, const Geom& g) {
    const LevelGeom& L = g.L[c.level];
    const int ch = c.ch;
    const int col0 = EAOF_INNER_X0 + c.iniX;
    const int mis = col0 & 3;  // cell pixel x sits at tile byte column mis + x
    const int nwords = (mis + c.cw + 3) >> 2;
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(
        pyr + (size_t)f * g.pyrFrameBytes + L.off + (size_t)(EAOF_EDGE + c.iniY) * L.pitch + (col0 - mis));
    const int pitchW = L.pitch >> 2;
    if (2 * nwords <= 32) {  // two rows per pass, four passes per loop trip (their loads in flight together)
        const int half = lane >= 16, wl = lane & 15;
        if (wl < nwords) {
            const uint32_t* s = src32 + (size_t)half * pitchW + wl;
            uint32_t* d = tile + half * PW + 1 + wl;
            const size_t step = (size_t)2 * pitchW;
            int r = half;
            for (; r + 6 < ch; r += 8, s += 4 * step, d += 8 * PW) {
                const uint32_t v0 = __ldg(s), v1 = __ldg(s + step), v2 = __ldg(s + 2 * step), v3 = __ldg(s + 3 * step);
                d[0] = v0;
                d[2 * PW] = v1;
                d[4 * PW] = v2;
                d[6 * PW] = v3;
            }
            for (; r < ch; r += 2, s += step, d += 2 * PW) d[0] = __ldg(s);
        }
    } else if (lane < nwords) {
        for (int r = 0; r < ch; ++r) tile[r * PW + 1 + lane] = __ldg(src32 + r * pitchW + lane);
    }
    for (int i = lane; i < g.fastMapWords / 4; i += 32) reinterpret_cast<uint4*>(Bm)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
}

template <int PW>
__global__ void __launch_bounds__(FAST_WARPS * 32, FAST_MINB) k_fast(const uint8_t* __restrict__ pyr,
                                                                     const CellDesc* __restrict__ cells,
                                                                     uint32_t* __restrict__ cand,
                                                                     uint32_t* __restrict__ candCount,
                                                                     const __grid_constant__ Geom g, const int cellBegin,
                                                                     const int cellEnd) {
    extern __shared__ __align__(16) uint32_t fastSmem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cell = cellBegin + blockIdx.x * FAST_WARPS + warp;  // a launch covers cells [cellBegin, cellEnd): all, or one level
    if (cell >= cellEnd) return;
    const int mapWords = g.fastMapWords;         // multiple of 4
    uint32_t* tile = fastSmem + (size_t)warp * g.fastWarpWords;  // pixel (row r, tile byte column c) at byte (r*PW + 1)*4 + c
    uint32_t* Bm = tile + mapWords;              // arc score of corners (0 elsewhere), same layout
    uint16_t* clst = reinterpret_cast<uint16_t*>(Bm + mapWords);  // corners found by (B): row << 8 | tile byte column
    uint16_t* lst = clst + FAST_CLST2;                            // surviving pixels, same encoding
    const CellDesc c = cells[cell];
    const int f = blockIdx.y;
    if (c.cw <= 6 || c.ch <= 6) return;
    fast_stage_tile(pyr, tile, Bm, c, f, lane, PW, g);
    const int lstCap = 2 * (g.fastWarpWords - 2 * mapWords) - FAST_CLST2;  // entries the survivor list holds (>= 16 rows of pixels)
    fast_cell_rows<PW>(tile, Bm, clst, lst, lstCap, c, f, (EAOF_INNER_X0 + c.iniX) & 3, lane, cand, candCount, g);
}

// The task-per-word search (fast_cell) for handles whose geometry or thresholds fast_cell_rows does not cover: cells wider
// than 16 inner words, minThFAST >= iniThFAST, iniThFAST >= 128.  Same shared-memory carve-up as k_fast.
__global__ void __launch_bounds__(FAST_WARPS * 32, FAST_MINB) k_fast_generic(const uint8_t* __restrict__ pyr,
                                                                             const CellDesc* __restrict__ cells,
                                                                             uint32_t* __restrict__ cand,
                                                                             uint32_t* __restrict__ candCount,
                                                                             const __grid_constant__ Geom g, const int cellBegin,
                                                                             const int cellEnd) {
    extern __shared__ __align__(16) uint32_t fastSmem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cell = cellBegin + blockIdx.x * FAST_WARPS + warp;
    if (cell >= cellEnd) return;
    const int mapWords = g.fastMapWords;
    uint32_t* tile = fastSmem + (size_t)warp * g.fastWarpWords;
    uint32_t* Bm = tile + mapWords;
    uint16_t* clst = reinterpret_cast<uint16_t*>(Bm + mapWords);
    uint16_t* lst = clst + FAST_CLST;
    const CellDesc c = cells[cell];
    const int f = blockIdx.y;
    if (c.cw <= 6 || c.ch <= 6) return;
    fast_stage_tile(pyr, tile, Bm, c, f, lane, g.fastPW, g);
    const int lstCap = 2 * (g.fastWarpWords - 2 * mapWords) - FAST_CLST;  // entries the survivor list holds
    fast_cell<1>(tile, Bm, clst, lst, lstCap, c, f, (EAOF_INNER_X0 + c.iniX) & 3, g.fastPW, lane, cand, candCount, g);
}

// ---- k_fast_tma: the same cell search with the tile staged by TMA --------------------------------------------------
// Persistent warps (no CTA-wide barrier): every warp pulls (frame, cell) work items off a global counter and runs a
// four-stage software pipeline over them — item j is claimed (atomicAdd) in iteration j-3, its CellDesc is fetched in
// j-2, its tile is requested in j-1 with ONE cp.async.bulk.tensor (a 3-D box {FAST box width, box height, 1} of the
// level's u8 tensor {pitch, rows, frames}, landing densely in one of the warp's two tile buffers and completing on that
// buffer's mbarrier), and it is searched in iteration j.  That replaces ~19 x (LDG + STS + address) instructions per lane
// of k_fast, and the load latency of the next cell hides under the search of the current one.
// TMA rule met the hard way (tools/tma_probe2.cu): the byte offset of the box's innermost start coordinate must be a
// multiple of 16 — any other x raises "illegal instruction" (asynchronously, and not on every request).  The box
// therefore starts at the cell's column rounded down to 16 and is 15 bytes wider than the widest cell; the cell's pixel x
// sits at tile byte column mis + x with mis = column & 15, which fast_cell handles like k_fast's 4-byte alignment.
// Shared memory per warp: 2 tile buffers + score map (box geometry), corner list, survivor list, 2 mbarriers.
#ifndef FASTT_WARPS
#define FASTT_WARPS 4
#endif
#ifndef FASTT_MINB
#define FASTT_MINB 6
#endif

struct FastTmaArgs {
    int boxW, boxH;      // box = tile geometry: boxW bytes per row (multiple of 16), boxH rows
    int tileBytes;       // boxW*boxH rounded up to 128
    int warpBytes;       // shared memory per warp (multiple of 128)
    int lstCap;          // survivor-list entries
    int f0;              // first frame of this launch inside the handle's pyramid buffer (chunked batches)
    int nItems;          // frames of this launch * cells per frame
    int* dbg;            // EAOF_TMA_DEBUG builds: host-mapped record per warp of the last request
};

__global__ void __launch_bounds__(FASTT_WARPS * 32, FASTT_MINB) k_fast_tma(const __grid_constant__ FastTmaMaps maps, const FastTmaArgs A,
                                                                          const CellDesc* __restrict__ cells,
                                                                          uint32_t* __restrict__ cand,
                                                                          uint32_t* __restrict__ candCount,
                                                                          unsigned int* __restrict__ workCounter,
                                                                          const __grid_constant__ Geom g) {
    extern __shared__ __align__(128) uint8_t fastTmaSmem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the dynamic shared memory window is 16-byte aligned by contract; TMA destinations need 128
    uint8_t* base = fastTmaSmem + ((128u - (smem_u32(fastTmaSmem) & 127u)) & 127u) + (size_t)warp * A.warpBytes;
    uint32_t* Bm = reinterpret_cast<uint32_t*>(base + 2 * A.tileBytes);
    uint16_t* clst = reinterpret_cast<uint16_t*>(base + 3 * A.tileBytes);
    uint16_t* lst = clst + FAST_CLST;
    const uint32_t bar0 = smem_u32(base + 3 * A.tileBytes + 2 * FAST_CLST + 2 * A.lstCap);  // 8-byte aligned: see host sizing
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int PW = A.boxW >> 2;
    const int mapWords4 = A.tileBytes >> 4;
    const int nItems = A.nItems, cpf = g.cellsPerFrame;

    auto claim = [&]() -> int {
        unsigned v = 0;
        if (lane == 0) v = atomicAdd(workCounter, 1u);
        return (int)__shfl_sync(0xffffffffu, v, 0);
    };
    auto fetch = [&](int item, int& f, CellDesc& c) {
        f = item / cpf;
        c = cells[item - f * cpf];  // one 16-byte load
    };
    auto request = [&](const CellDesc& c, int f, int buf) {
        if (lane == 0) {
            // this buffer was read (and its head overwritten with NMS survivors) through the generic proxy two iterations
            // ago: order those accesses before the async-proxy write of the new box
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t bar = bar0 + 8 * buf;
#ifdef EAOF_TMA_DEBUG
            if (A.dbg) {
                int* r = A.dbg + 8 * (blockIdx.x * FASTT_WARPS + warp);
                r[0] = c.level; r[1] = (EAOF_INNER_X0 + c.iniX) & ~15; r[2] = EAOF_EDGE + c.iniY; r[3] = A.f0 + f;
                r[4] = (int)smem_u32(base + buf * A.tileBytes); r[5] = (int)bar; r[6] = buf; r[7] += 1;
                __threadfence_system();
            }
#endif
            mbar_expect_tx(bar, (uint32_t)(A.boxW * A.boxH));
            tma_load_3d(smem_u32(base + buf * A.tileBytes), &maps.m[c.level][0], (EAOF_INNER_X0 + c.iniX) & ~15, EAOF_EDGE + c.iniY,
                        A.f0 + f, bar);
        }
    };

    // pipeline registers: item claimed (i3), item with descriptor (i2: c2, f2), item with tile in flight (i1: c1, f1)
    int i1 = claim(), i2 = claim(), i3 = claim();
    CellDesc c1{}, c2{};
    int f1 = 0, f2 = 0;
    if (i1 < nItems) { fetch(i1, f1, c1); request(c1, f1, 0); }
    if (i2 < nItems) fetch(i2, f2, c2);
    int buf = 0;
    uint32_t phases = 0;  // bit b = parity the next wait on buffer b's mbarrier expects
    while (i1 < nItems) {
        const CellDesc c = c1;
        const int f = f1;
        // advance the pipeline before searching: tile of the next item, descriptor of the one after, claim of the third
        i1 = i2; c1 = c2; f1 = f2;
        if (i1 < nItems) request(c1, f1, buf ^ 1);
        i2 = i3;
        if (i2 < nItems) fetch(i2, f2, c2);
        i3 = claim();
        // score map of this cell
        for (int i = lane; i < mapWords4; i += 32) reinterpret_cast<uint4*>(Bm)[i] = make_uint4(0, 0, 0, 0);
        mbar_wait(bar0 + 8 * buf, (phases >> buf) & 1u);
        phases ^= 1u << buf;
        __syncwarp();
        if (c.cw > 6 && c.ch > 6)
            fast_cell<0>(reinterpret_cast<uint32_t*>(base + buf * A.tileBytes), Bm, clst, lst, A.lstCap, c, f,
                         (EAOF_INNER_X0 + c.iniX) & 15, PW, lane, cand, candCount, g);
        __syncwarp();
        buf ^= 1;
    }
}

// k_fast_tma1: the smallest TMA form — one warp and one cell per CTA like k_fast, the tile requested with one
// cp.async.bulk.tensor (no double buffering: the other resident warps cover the latency, the score map is cleared while the
// box is in flight).  Same occupancy as k_fast (32 one-warp CTAs per SM), minus its tile-load instructions.
__global__ void __launch_bounds__(32, 32) k_fast_tma1(const __grid_constant__ FastTmaMaps maps, const FastTmaArgs A,
                                                      const CellDesc* __restrict__ cells, uint32_t* __restrict__ cand,
                                                      uint32_t* __restrict__ candCount, const __grid_constant__ Geom g) {
    extern __shared__ __align__(128) uint8_t fastTma1Smem[];
    const int lane = threadIdx.x;
    uint8_t* base = fastTma1Smem + ((128u - (smem_u32(fastTma1Smem) & 127u)) & 127u);
    uint32_t* Bm = reinterpret_cast<uint32_t*>(base + A.tileBytes);
    uint16_t* clst = reinterpret_cast<uint16_t*>(base + 2 * A.tileBytes);
    uint16_t* lst = clst + FAST_CLST;
    const uint32_t bar = smem_u32(base + 2 * A.tileBytes + 2 * FAST_CLST + 2 * A.lstCap);
    const CellDesc c = cells[blockIdx.x];
    const int f = blockIdx.y;
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, (uint32_t)(A.boxW * A.boxH));
        tma_load_3d(smem_u32(base), &maps.m[c.level][0], (EAOF_INNER_X0 + c.iniX) & ~15, EAOF_EDGE + c.iniY, A.f0 + f, bar);
    }
    for (int i = lane; i < (A.tileBytes >> 4); i += 32) reinterpret_cast<uint4*>(Bm)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    mbar_wait(bar, 0);
    fast_cell<0>(reinterpret_cast<uint32_t*>(base), Bm, clst, lst, A.lstCap, c, f, (EAOF_INNER_X0 + c.iniX) & 15, A.boxW >> 2, lane, cand,
                 candCount, g);
}

// ------------------------------------------------------------------------------------------------
// DistributeOctTree.  Parallel formulation (validated on the CPU by oracle/orb_oracle.cc against the
// reference's std::list code): keys never move, each key carries the list position of its node, a pass
// (a) counts the four quadrant populations of every node that may be split, (b) decides which nodes are split
// and where their children land, (c) relabels the keys.  push_front puts children in front of the list in
// reverse creation order, so "later created" == "smaller list position", which is also the canonical
// tie-break for the reference's (size, pointer) sort.
// 256 threads per (frame, level): the passes are barrier-bound, and twice as many resident CTAs hide more of it than
// 512-thread CTAs do (measured 0.130 -> 0.102 ms per 250 frames; splitting small and large levels into two launches of
// different widths was slower than one launch)
// The CTA width is a template parameter chosen per launch: with few (frame, level) CTAs the launch lasts as long as the CTA of
// the largest level, which then wants 1024 threads; with thousands of CTAs 256-thread ones keep more of them resident.
// Keys and labels are copied to shared memory when the level's candidates fit the launch's budget (keyCap): every pass
// reads all keys twice, and from global memory each of those sweeps is a chain of L2 round trips (at 1920x1080 / 4000
// features the level-0 CTA spent 0.43 ms that way).

struct OctShared {
    int n;        // list size
    int cPrev;    // children created by the previous pass (= candidates live in [0, cPrev))
    int m;        // nodes in processing order
    int nsplit;   // how many of them are split
    int created;  // children created by this pass
    int flag;
    int warp[32];
};

template <int OCT_THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* warpSums) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warpSums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int s = lane < (OCT_THREADS / 32) ? warpSums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        warpSums[lane] = s;
    }
    __syncthreads();
    const int base = wid > 0 ? warpSums[wid - 1] : 0;
    *total = warpSums[OCT_THREADS / 32 - 1];
    const int r = base + incl - v;
    __syncthreads();
    return r;
}

__device__ __forceinline__ int ceil_half(int a) { return (a + 1) >> 1; }  // ceil((float)a/2) for a >= 0

template <int OCT_THREADS>
__global__ void __launch_bounds__(OCT_THREADS, OCT_THREADS == 256 ? 6 : OCT_THREADS == 512 ? 2 : 1) k_octree(const uint32_t* __restrict__ cand,
                                                         const uint32_t* __restrict__ candCount,
                                                         uint16_t* __restrict__ label, uint32_t* __restrict__ slotXY,
                                                         uint8_t* __restrict__ slotScore, int* __restrict__ lvlCount,
                                                         const __grid_constant__ Geom g, const int keyCap, const int levelBegin) {
    extern __shared__ __align__(16) uint8_t smemRaw[];
    __shared__ OctShared S;
    const int l = levelBegin + blockIdx.x, f = blockIdx.y;  // a launch covers levels [levelBegin, levelBegin + gridDim.x)
    const LevelGeom& L = g.L[l];
    const int tid = threadIdx.x, lane = tid & 31;
    const int cap = g.maxNodeCap;
    const int N = L.quota;

    // shared arrays
    short4* bndA = reinterpret_cast<short4*>(smemRaw);                    // x0,y0,x1,y1
    short4* bndB = bndA + cap;
    uint32_t* cntA = reinterpret_cast<uint32_t*>(bndB + cap);
    uint32_t* cntB = cntA + cap;
    uint32_t* cc = cntB + cap;                                            // [cap*4] quadrant counts / final best (u64 x cap*2)
    uint16_t* childPos = reinterpret_cast<uint16_t*>(cc + 4 * (size_t)cap);  // [cap*4]
    uint16_t* keepPos = childPos + 4 * (size_t)cap;                       // [cap]
    uint16_t* firstChild = keepPos + cap;                                 // [cap]
    uint16_t* proc = firstChild + cap;                                    // [cap] processing order -> list position
    uint16_t* tmpIdx = proc + cap;                                        // [cap]
    uint8_t* wanted = reinterpret_cast<uint8_t*>(tmpIdx + cap);           // [cap]
    uint8_t* split = wanted + cap;                                        // [cap]
    uint8_t* ne = split + cap;                                            // [cap] non-empty children

    uint32_t* tmpCnt = reinterpret_cast<uint32_t*>(childPos);            // sorted pass only: key counts in tmpIdx order
    uint32_t* rnk = reinterpret_cast<uint32_t*>(keepPos);                 // sorted pass only: rank accumulators (keepPos + firstChild)

    const int nk = (int)candCount[f * g.nlevels + l];
    const uint32_t* gkeys = cand + (size_t)f * g.candPerFrame + L.candOff;
    int* outCount = lvlCount + f * g.nlevels + l;

    if (nk == 0 || L.nIni < 1) {
        if (tid == 0) *outCount = 0;
        return;
    }
    // keys / labels: shared memory behind the node arrays when they fit, else the global candidate / label arrays
    uint32_t* sKeys = reinterpret_cast<uint32_t*>(smemRaw + ((((size_t)cap * 59 + 64) + 15) & ~(size_t)15));
    const bool inS = nk <= keyCap;
    const uint32_t* keys = inS ? sKeys : gkeys;
    uint16_t* lab = inS ? reinterpret_cast<uint16_t*>(sKeys + keyCap) : label + (size_t)f * g.candPerFrame + L.candOff;
    if (inS)
        for (int k = tid; k < nk; k += OCT_THREADS) sKeys[k] = gkeys[k];

    // ---- roots (:543-585)
    for (int i = tid; i < cap; i += OCT_THREADS) cc[i] = 0;
    __syncthreads();
    for (int k = tid; k < nk; k += OCT_THREADS) {
        const int x = keys[k] & 0xfff;
        const int r = (int)__fdiv_rn((float)x, L.hX);
        atomicAdd(&cc[r], 1u);
    }
    __syncthreads();
    {
        // nIni is tiny (1..4 for any sane aspect ratio): thread 0 compacts the non-empty roots
        if (tid == 0) {
            int n = 0;
            for (int i = 0; i < L.nIni; ++i) {
                if (cc[i] > 0) {
                    bndA[n] = make_short4((short)(int)__fmul_rn(L.hX, (float)i), 0,
                                          (short)(int)__fmul_rn(L.hX, (float)(i + 1)), (short)L.winH);
                    cntA[n] = cc[i];
                    tmpIdx[i] = (uint16_t)n;
                    ++n;
                }
            }
            S.n = n;
            S.cPrev = 0;
        }
    }
    __syncthreads();
    for (int k = tid; k < nk; k += OCT_THREADS) {
        const int x = keys[k] & 0xfff;
        lab[k] = tmpIdx[(int)__fdiv_rn((float)x, L.hX)];
    }
    __syncthreads();

    short4* bnd = bndA;
    short4* bnd2 = bndB;
    uint32_t* cnt = cntA;
    uint32_t* cnt2 = cntB;

    // One pass.  sorted == false: sweep, every node with more than one key is split (:594-665).
    // sorted == true: candidates are the >1-key children of the previous pass, split in (size desc,
    // position asc) order until the list reaches N nodes (:676-737).
    auto run_pass = [&](bool sorted) {
        const int n = S.n;
        const int cPrev = S.cPrev;
        // 1. processing order
        int mTotal = 0;
        for (int i0 = 0; i0 < n; i0 += OCT_THREADS) {
            const int i = i0 + tid;
            int w = 0;
            if (i < n) {
                w = (cnt[i] > 1 && (!sorted || i < cPrev)) ? 1 : 0;
                wanted[i] = (uint8_t)w;
                split[i] = 0;
                cc[4 * i] = cc[4 * i + 1] = cc[4 * i + 2] = cc[4 * i + 3] = 0;
            }
            int tot;
            const int e = block_excl_scan<OCT_THREADS>(w, &tot, S.warp);
            if (w) {
                (sorted ? tmpIdx : proc)[mTotal + e] = (uint16_t)i;
                if (sorted) {
                    tmpCnt[mTotal + e] = cnt[i];
                    rnk[mTotal + e] = 0u;
                }
            }
            mTotal += tot;
        }
        __syncthreads();
        const int m = mTotal;
        if (sorted) {
            // rank by (cnt desc, position asc); ranks are a permutation.  tmpIdx is ascending in position, so "position of b <
            // position of a" is b < a.  The m x m comparisons are dealt to ALL threads: element a x a chunk of b's per thread
            // (m is a few dozen to a few hundred: one thread per element left most of the CTA idle).
            const int nCh = m < OCT_THREADS ? OCT_THREADS / m : 1, C = (m + nCh - 1) / nCh;
            for (int t = tid; t < m * nCh; t += OCT_THREADS) {
                const int ch = t / m, a = t - ch * m;
                const int b0 = ch * C, b1 = min(m, b0 + C);
                const uint32_t ca = tmpCnt[a];
                unsigned r = 0;
                for (int b = b0; b < b1; ++b) {
                    const uint32_t cb = tmpCnt[b];
                    r += (cb > ca) || (cb == ca && b < a);
                }
                if (nCh > 1) atomicAdd(&rnk[a], r); else rnk[a] = r;
            }
            __syncthreads();
            for (int a = tid; a < m; a += OCT_THREADS) proc[rnk[a]] = tmpIdx[a];
            __syncthreads();
        }
        // 2. quadrant populations
        for (int k0 = 0; k0 < nk; k0 += OCT_THREADS) {
            const int k = k0 + tid;
            int slot = -1;
            if (k < nk) {
                const int p = lab[k];
                if (wanted[p]) {
                    const uint32_t kw = keys[k];
                    const int x = kw & 0xfff, y = (kw >> 12) & 0xfff;
                    const short4 b = bnd[p];
                    const int midX = b.x + ceil_half(b.z - b.x), midY = b.y + ceil_half(b.w - b.y);
                    slot = 4 * p + (x < midX ? 0 : 1) + (y < midY ? 0 : 2);
                }
            }
            // warp-aggregated shared atomics while there are few nodes (early passes put thousands of keys on four
            // counters); with many nodes the lanes of a warp rarely meet and match_any costs more than it saves
            if (n <= 16) {
                const unsigned act = __ballot_sync(0xffffffffu, slot >= 0);
                if (slot >= 0) {
                    const unsigned peers = __match_any_sync(act, slot);
                    if ((int)(__ffs(peers) - 1) == lane) atomicAdd(&cc[slot], (uint32_t)__popc(peers));
                }
            } else if (slot >= 0) {
                atomicAdd(&cc[slot], 1u);
            }
        }
        __syncthreads();
        // 3. children per node, in processing order; stop rule for the sorted pass
        if (tid == 0) { S.nsplit = m; }
        __syncthreads();
        if (sorted) {
            int carry = 0;
            for (int r0 = 0; r0 < m; r0 += OCT_THREADS) {
                const int r = r0 + tid;
                int v = 0;
                if (r < m) {
                    const int p = proc[r];
                    v = (cc[4 * p] > 0) + (cc[4 * p + 1] > 0) + (cc[4 * p + 2] > 0) + (cc[4 * p + 3] > 0) - 1;
                }
                int tot;
                const int e = block_excl_scan<OCT_THREADS>(v, &tot, S.warp);
                if (r < m) {
                    const int incl = carry + e + v;       // list growth after splitting r
                    const int before = carry + e;
                    if (n + incl >= N && n + before < N) atomicMin(&S.nsplit, r + 1);
                }
                carry += tot;
            }
            __syncthreads();
        }
        const int nsplit = S.nsplit;
        int created = 0;
        for (int r0 = 0; r0 < nsplit; r0 += OCT_THREADS) {
            const int r = r0 + tid;
            int v = 0, p = 0;
            if (r < nsplit) {
                p = proc[r];
                v = (cc[4 * p] > 0) + (cc[4 * p + 1] > 0) + (cc[4 * p + 2] > 0) + (cc[4 * p + 3] > 0);
            }
            int tot;
            const int e = block_excl_scan<OCT_THREADS>(v, &tot, S.warp);
            if (r < nsplit) {
                split[p] = 1;
                ne[p] = (uint8_t)v;
                firstChild[p] = (uint16_t)(created + e);
            }
            created += tot;
        }
        __syncthreads();
        // 4. surviving nodes keep their relative order behind the new children
        int kept = 0;
        for (int i0 = 0; i0 < n; i0 += OCT_THREADS) {
            const int i = i0 + tid;
            const int v = (i < n && !split[i]) ? 1 : 0;
            int tot;
            const int e = block_excl_scan<OCT_THREADS>(v, &tot, S.warp);
            if (v) keepPos[i] = (uint16_t)(created + kept + e);
            kept += tot;
        }
        __syncthreads();
        // 5. new list
        for (int i = tid; i < n; i += OCT_THREADS) {
            const short4 b = bnd[i];
            if (!split[i]) {
                bnd2[keepPos[i]] = b;
                cnt2[keepPos[i]] = cnt[i];
            } else {
                const int mx = b.x + ceil_half(b.z - b.x), my = b.y + ceil_half(b.w - b.y);
                int c = firstChild[i];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t kc = cc[4 * i + q];
                    if (kc == 0) continue;
                    const int np = created - 1 - c;
                    bnd2[np] = make_short4((short)((q & 1) ? mx : b.x), (short)((q & 2) ? my : b.y),
                                           (short)((q & 1) ? b.z : mx), (short)((q & 2) ? b.w : my));
                    cnt2[np] = kc;
                    childPos[4 * i + q] = (uint16_t)np;
                    ++c;
                }
            }
        }
        __syncthreads();
        // 6. relabel
        for (int k = tid; k < nk; k += OCT_THREADS) {
            const int p = lab[k];
            if (!split[p]) {
                lab[k] = keepPos[p];
            } else {
                const uint32_t kw = keys[k];
                const int x = kw & 0xfff, y = (kw >> 12) & 0xfff;
                const short4 b = bnd[p];
                const int midX = b.x + ceil_half(b.z - b.x), midY = b.y + ceil_half(b.w - b.y);
                lab[k] = childPos[4 * p + (x < midX ? 0 : 1) + (y < midY ? 0 : 2)];
            }
        }
        __syncthreads();
        if (tid == 0) {
            S.n = created + kept;
            S.cPrev = created;
        }
        short4* tb = bnd; bnd = bnd2; bnd2 = tb;
        uint32_t* tc = cnt; cnt = cnt2; cnt2 = tc;
        __syncthreads();
    };

    auto count_expandable = [&]() -> int {  // children of the last pass with more than one key
        const int cPrev = S.cPrev;
        int v = 0;
        for (int i = tid; i < cPrev; i += OCT_THREADS) v += cnt[i] > 1;
        int tot;
        block_excl_scan<OCT_THREADS>(v, &tot, S.warp);
        return tot;
    };

    bool finish = false;
    while (!finish) {
        const int prevSize = S.n;
        __syncthreads();
        run_pass(false);
        const int size = S.n;
        if (size >= N || size == prevSize) {
            finish = true;
        } else {
            const int nToExpand = count_expandable();
            if (size + 3 * nToExpand > N) {
                while (!finish) {
                    const int prev2 = S.n;
                    __syncthreads();
                    if (count_expandable() > 0) run_pass(true);
                    if (S.n >= N || S.n == prev2) finish = true;
                }
            }
        }
    }
    __syncthreads();

    // ---- best response per node, first in cell-raster order wins (:741-760)
    const int n = S.n;
    unsigned long long* best = reinterpret_cast<unsigned long long*>(cc);
    for (int i = tid; i < n; i += OCT_THREADS) best[i] = 0ull;
    __syncthreads();
    for (int k = tid; k < nk; k += OCT_THREADS) {
        const uint32_t kw = keys[k];
        const int x = kw & 0xfff, y = (kw >> 12) & 0xfff;
        const uint32_t sc = kw >> 24;
        const int cj = (x - 3) / L.wCell, ci = (y - 3) / L.hCell;
        const uint32_t ord = ((uint32_t)(ci * L.nCols + cj) << 12) | ((uint32_t)(y - 3 - ci * L.hCell) << 6) |
                             (uint32_t)(x - 3 - cj * L.wCell);
        atomicMax(&best[lab[k]], ((unsigned long long)(sc + 1) << 32) | (0xffffffffu - ord));
    }
    __syncthreads();
    uint32_t* oXY = slotXY + (size_t)f * g.slotsPerFrame + L.slotOff;
    uint8_t* oSc = slotScore + (size_t)f * g.slotsPerFrame + L.slotOff;
    for (int i = tid; i < n; i += OCT_THREADS) {
        const unsigned long long b = best[i];
        const uint32_t ord = 0xffffffffu - (uint32_t)(b & 0xffffffffu);
        const int cell = ord >> 12, ci = cell / L.nCols, cj = cell - ci * L.nCols;
        const int y = ci * L.hCell + 3 + ((ord >> 6) & 63), x = cj * L.wCell + 3 + (ord & 63);
        oXY[i] = (uint32_t)x | ((uint32_t)y << 16);
        oSc[i] = (uint8_t)((b >> 32) - 1);
    }
    if (tid == 0) *outCount = n;
}

// ------------------------------------------------------------------------------------------------
// Gaussian blur 7x7, sigma 2 on the inner level (borders come from the REFLECT_101 pyramid border, which is
// what cv::GaussianBlur(BORDER_REFLECT_101) on the cloned ROI sees).  Integer arithmetic, SURVEY.md A.6.
// One thread per (column word, chunk of BLUR_ROWS rows): it walks down its 4 columns with the last seven
// horizontally filtered rows in registers.  The horizontal pass runs on packed 16x2 lanes (the 7-tap sum of
// bytes times 8-bit taps is <= 255*257 = 65535, so two pixels share a register and a plain IMAD never carries from
// one lane into the other); the vertical pass needs 25 bits and runs per pixel.
#ifndef BLUR_ROWS
#define BLUR_ROWS 16
#endif
#ifndef BLUR_THREADS
#define BLUR_THREADS 128
#endif
#ifndef BLUR_AHEAD
#define BLUR_AHEAD 3
#endif

__global__ void __launch_bounds__(BLUR_THREADS) k_blur(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur,
                                                       const __grid_constant__ Geom g) {
    const int t = blockIdx.x * BLUR_THREADS + threadIdx.x;
    if (t >= g.blurTasksPerFrame) return;
    const int f = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && t >= g.L[l + 1].blurTaskOff) ++l;
    const LevelGeom& L = g.L[l];
    const int nCW = (L.w + 3) >> 2;
    const int tt = t - L.blurTaskOff;
    const int rc = tt / nCW, cwd = tt - rc * nCW;
    const int x = 4 * cwd, y0 = rc * BLUR_ROWS;
    const bool cv4 = g.blurMode == 1;
    const unsigned k0 = 18, k1 = 34, k2 = cv4 ? 48 : 49, k3 = cv4 ? 56 : 55;
    // taps as byte vectors for the integer dot-product instructions
    const unsigned TL = k0 | (k1 << 8) | (k2 << 16) | (k3 << 24);  // pixels x-3 .. x
    const unsigned TR = k2 | (k1 << 8) | (k0 << 16);                // pixels x+1 .. x+3
    const unsigned TV = k2 | (k1 << 8);                              // rows y+1, y+2 (TL's halves serve rows y-3..y)
    const int simdW = g.blurMode == 2 ? (L.w & ~3) : 0;
    const bool halfEven = x < simdW;  // simdW is a multiple of 4, so the whole word rounds the same way
    const size_t lvl = (size_t)f * g.pyrFrameBytes + L.off;
    const uint8_t* in = pyr + lvl + EAOF_INNER_X0 + x;         // word-aligned
    uint8_t* out = blur + lvl + (size_t)EAOF_EDGE * L.pitch + EAOF_INNER_X0 + x;

    // Horizontal pass: 7 taps of one pixel = two DP4A over byte groups picked by PRMT (sum <= 255*257 fits 16 bits).
    // Vertical pass: the horizontally filtered values of vertically adjacent rows are kept as 16x2 pairs
    // pr[r] = (h[r-1], h[r]); an output row is three DP2A over the pairs created 5, 3 and 1 rows ago plus k0 times the
    // newest value (25 bits).
    unsigned prA[4] = {0, 0, 0, 0}, prB[4] = {0, 0, 0, 0}, prC[4] = {0, 0, 0, 0}, prD[4] = {0, 0, 0, 0}, prE[4] = {0, 0, 0, 0};
    unsigned hPrev[4] = {0, 0, 0, 0};
    // The stage waits on its loads (ncu: long-scoreboard stalls dominate, issue slots half used), so the three words of
    // a row are fetched BLUR_AHEAD rows before they are filtered.
    unsigned qm[BLUR_AHEAD], q0[BLUR_AHEAD], qp[BLUR_AHEAD];
#pragma unroll
    for (int r = 0; r < BLUR_AHEAD; ++r) {
        const int by = min(y0 + r - 3 + EAOF_EDGE, L.rows - 1);
        const uint32_t* p = reinterpret_cast<const uint32_t*>(in + (size_t)by * L.pitch);
        qm[r] = __ldg(p - 1); q0[r] = __ldg(p); qp[r] = __ldg(p + 1);
    }
#pragma unroll
    for (int r = 0; r < BLUR_ROWS + 6; ++r) {
        const unsigned Wm = qm[r % BLUR_AHEAD], W0 = q0[r % BLUR_AHEAD], Wp = qp[r % BLUR_AHEAD];
        if (r + BLUR_AHEAD < BLUR_ROWS + 6) {
            const int by = min(y0 + r + BLUR_AHEAD - 3 + EAOF_EDGE, L.rows - 1);
            const uint32_t* p = reinterpret_cast<const uint32_t*>(in + (size_t)by * L.pitch);
            qm[r % BLUR_AHEAD] = __ldg(p - 1); q0[r % BLUR_AHEAD] = __ldg(p); qp[r % BLUR_AHEAD] = __ldg(p + 1);
        }
        unsigned h[4];
        h[0] = __dp4a(__byte_perm(Wm, W0, 0x4321), TL, __dp4a(__byte_perm(W0, Wp, 0x4321), TR, 0u));
        h[1] = __dp4a(__byte_perm(Wm, W0, 0x5432), TL, __dp4a(__byte_perm(W0, Wp, 0x5432), TR, 0u));
        h[2] = __dp4a(__byte_perm(Wm, W0, 0x6543), TL, __dp4a(__byte_perm(W0, Wp, 0x6543), TR, 0u));
        h[3] = __dp4a(W0, TL, __dp4a(Wp, TR, 0u));
        // pairs created at rows r-5 (prA) .. r-1 (prE); shift the history and append (h[r-1], h[r])
        unsigned a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned prNew = hPrev[j] | (h[j] << 16);
            // window rows r-6..r: (r-6, r-5) = prA, (r-4, r-3) = prC, (r-2, r-1) = prE, single r
            a[j] = __dp2a_lo(prA[j], TL, __dp2a_hi(prC[j], TL, __dp2a_lo(prE[j], TV, k0 * h[j])));
            prA[j] = prB[j]; prB[j] = prC[j]; prC[j] = prD[j]; prD[j] = prE[j]; prE[j] = prNew;
            hPrev[j] = h[j];
        }
        if (r >= 6) {
            const int y = y0 + r - 6;
            if (y < L.h) {
                uint32_t v;
                if (halfEven) {
                    v = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int o = (int)(a[j] >> 16);
                        const int rem = (int)(a[j] & 0xffff);
                        o += (rem > 32768) || (rem == 32768 && (o & 1));
                        v |= (uint32_t)min(o, 255) << (8 * j);
                    }
                } else {
                    // (a + 2^15) >> 16 saturated to 255 = byte 2 of min(a + 2^15, 0xffffff)
                    const unsigned b0 = min(a[0] + 32768u, 0xffffffu), b1 = min(a[1] + 32768u, 0xffffffu);
                    const unsigned b2 = min(a[2] + 32768u, 0xffffffu), b3 = min(a[3] + 32768u, 0xffffffu);
                    v = __byte_perm(__byte_perm(b0, b1, 0x0062), __byte_perm(b2, b3, 0x0062), 0x5410);
                }
                // the padded row always has room for a full word (pitch >= w + 64)
                *reinterpret_cast<uint32_t*>(out + (size_t)y * L.pitch) = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fastAtan2 (degrees), SURVEY.md A.5 — every operation rounded to fp32 separately.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float k180pi = (float)(180 / 3.1415926535897932384626433832795);
    const float p1 = __fmul_rn(0.9997878412794807f, k180pi), p3 = __fmul_rn(-0.3258083974640975f, k180pi),
                p5 = __fmul_rn(0.1555786518463281f, k180pi), p7 = __fmul_rn(-0.04432655554792128f, k180pi);
    const float eps = 2.2204460492503131e-16f;  // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// glibc 2.39 sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h) for 0 <= y < 120, in double arithmetic.
__device__ __forceinline__ float sincos_poly(double x, double x2, bool neg, int n) {
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = __dmul_rn(x, x2);
        const double s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
        const double x7 = __dmul_rn(x3, x2);
        const double s = __dadd_rn(x, __dmul_rn(x3, S1));
        return __double2float_rn(__dadd_rn(s, __dmul_rn(x7, s1)));
    }
    const double sg = neg ? -1.0 : 1.0;
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __dadd_rn(sg * C3, __dmul_rn(x2, sg * C4));
    const double c1 = __dadd_rn(sg * C0, __dmul_rn(x2, sg * C1));
    const double x6 = __dmul_rn(x4, x2);
    const double c = __dadd_rn(c1, __dmul_rn(x4, sg * C2));
    return __double2float_rn(__dadd_rn(c, __dmul_rn(x6, c2)));
}
__device__ __forceinline__ void glibc_sincosf(float y, float* sp, float* cp) {
    double x = (double)y;
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ff;
    if (top < ((0x3f490fdbu >> 20) & 0x7ff)) {  // |y| < pi/4 (abstop12 compare)
        const double x2 = __dmul_rn(x, x);
        if (top < ((0x39800000u >> 20) & 0x7ff)) {  // |y| < 2^-12
            *sp = y;
            *cp = 1.0f;
            return;
        }
        *sp = sincos_poly(x, x2, false, 0);
        *cp = sincos_poly(x, x2, false, 1);
        return;
    }
    const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __dsub_rn(x, __dmul_rn((double)n, 0x1.921FB54442D18p0));
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const bool neg = (n & 2) != 0;
    const double xs = __dmul_rn(x, s), x2 = __dmul_rn(x, x);
    *sp = sincos_poly(xs, x2, neg, n);
    *cp = sincos_poly(xs, x2, neg, n ^ 1);
}

__global__ void k_debug_sincosf(uint32_t firstBits, uint32_t stride, uint32_t n, float* s, float* c) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    glibc_sincosf(__uint_as_float(firstBits + i * stride), &s[i], &c[i]);
}

#define EAOF_HALF_PATCH 15                  // HALF_PATCH_SIZE, src/ORBextractor.cc:73
#define EAOF_ANGLE_TASKS (31 * 9)           // (row, aligned word) pairs covering the 31x31 patch at any alignment
#define EAOF_ANGLE_TASKS_PAD 288
__device__ __align__(16) signed char d_pattern[EAOF_ORB_PATTERN_INTS];

// One warp per keypoint slot.  IC_Angle on the unblurred bordered level, descriptor on the blurred level.
__global__ void __launch_bounds__(256) k_angle_desc(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                                                    const uint32_t* __restrict__ slotXY,
                                                    const uint8_t* __restrict__ slotScore,
                                                    const int* __restrict__ lvlCount, const uint2* __restrict__ angleTab,
                                                    void* __restrict__ kpsOut,
                                                    uint8_t* __restrict__ descOut, int* __restrict__ kpCount,
                                                    int kpCap, const __grid_constant__ Geom g) {
    const int f = blockIdx.y;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= g.slotsPerFrame) return;
    int l = 0;
    while (l + 1 < g.nlevels && warp >= g.L[l + 1].slotOff) ++l;
    const LevelGeom& L = g.L[l];
    const int p = warp - L.slotOff;
    const int* lc = lvlCount + f * g.nlevels;
    if (warp == 0 && lane == 0) {
        int tot = 0;
        for (int i = 0; i < g.nlevels; ++i) tot += lc[i];
        kpCount[f] = tot;
    }
    if (p >= lc[l]) return;
    int outIdx = p;
    for (int i = 0; i < l; ++i) outIdx += lc[i];

    const uint32_t xy = slotXY[(size_t)f * g.slotsPerFrame + warp];
    const int X = (int)(xy & 0xffff) + EAOF_MIN_BORDER, Y = (int)(xy >> 16) + EAOF_MIN_BORDER;  // :841-842
    const size_t lvlBase = (size_t)f * g.pyrFrameBytes + L.off + (size_t)EAOF_EDGE * L.pitch + EAOF_INNER_X0;

    // IC_Angle: m10 = sum u*I, m01 = sum v*I over the radius-15 disc (749 pixels).  The 31 rows are read as 9 aligned
    // words each; for every (row, word) the table holds the four u weights and the four v weights as signed bytes
    // (0 outside the disc), one table per alignment of the patch, so a task is one pixel load and two DP4As.
    int m10 = 0, m01 = 0;
    {
        const int cx = EAOF_INNER_X0 + X - EAOF_HALF_PATCH;  // byte column of u = -15 inside the padded row
        const int a = cx & 3;
        const uint2* wt = angleTab + a * EAOF_ANGLE_TASKS_PAD;
        const uint8_t* base = pyr + (size_t)f * g.pyrFrameBytes + L.off + (size_t)(EAOF_EDGE + Y) * L.pitch + (cx - a);
#pragma unroll
        for (int k = 0; k < (EAOF_ANGLE_TASKS + 31) / 32; ++k) {
            const int i = lane + 32 * k;
            if (i < EAOF_ANGLE_TASKS) {
                const int r = (i * 57) >> 9;  // i / 9 for i < 288
                const int j = i - 9 * r;
                const unsigned pix = __ldg(reinterpret_cast<const uint32_t*>(base + (r - EAOF_HALF_PATCH) * L.pitch) + j);
                const uint2 w = __ldg(wt + i);
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m10) : "r"(pix), "r"(w.x));
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m01) : "r"(pix), "r"(w.y));
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // computeOrbDescriptor: lane i produces byte i (pairs 8i..8i+7)
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float a, b;
    glibc_sincosf(__fmul_rn(angle, factorPI), &b, &a);
    const uint8_t* bc = blur + lvlBase + (size_t)Y * L.pitch + X;
    const int4 pa = reinterpret_cast<const int4*>(d_pattern)[2 * lane];
    const int4 pb = reinterpret_cast<const int4*>(d_pattern)[2 * lane + 1];
    const int words[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x0 = (float)(signed char)(words[k] & 0xff), y0 = (float)(signed char)((words[k] >> 8) & 0xff);
        const float x1 = (float)(signed char)((words[k] >> 16) & 0xff), y1 = (float)(signed char)(words[k] >> 24);
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = bc[r0 * L.pitch + c0], t1 = bc[r1 * L.pitch + c1];
        val |= (t0 < t1) << k;
    }
    if (outIdx < kpCap) {
        descOut[((size_t)f * kpCap + outIdx) * 32 + lane] = (uint8_t)val;
        if (lane == 0) {
            float* o = reinterpret_cast<float*>(kpsOut) + ((size_t)f * kpCap + outIdx) * 6;
            float px = (float)X, py = (float)Y;
            if (l != 0) { px = __fmul_rn(px, L.scale); py = __fmul_rn(py, L.scale); }
            o[0] = px;
            o[1] = py;
            o[2] = L.kpSize;
            o[3] = angle;
            o[4] = (float)slotScore[(size_t)f * g.slotsPerFrame + warp];
            reinterpret_cast<int*>(o)[5] = l;
        }
    }
}

// ---- k_angle_desc_tma: the same stage with both patches of a keypoint staged by TMA ---------------------------------
// k_angle_desc reads the 749-pixel disc (unblurred level) and the 512 rotated rBRIEF taps (blurred level) straight from
// global memory: 1-byte gathers that touch ~20 different 128-byte lines per warp request and keep the L1 tag stage at 85 %.
// Here lane 0 of the warp that owns a keypoint requests two boxes — {48 x 31} of the unblurred level around the disc and
// {64 x 39} of the blurred level around the tap area (x origins rounded down to 16 bytes, TMA's rule; the taps reach at most
// 18.4 px from the centre) — with two cp.async.bulk.tensor on the warp's own mbarrier, and every later access is a
// shared-memory load: aligned words for the moments, bytes for the taps.  4 KB of shared memory per warp.
#define DESC_BOXA_W 48
#define DESC_BOXA_H 31
#define DESC_BOXB_W 64
#define DESC_BOXB_H 39
#define DESC_BOXA_BYTES 1536  // 48 * 31 = 1488, padded to a multiple of 128
#define DESC_WARP_BYTES 4096  // + 64 * 39 = 2496
#define DESC_TMA_WARPS 8
__global__ void __launch_bounds__(DESC_TMA_WARPS * 32) k_angle_desc_tma(const __grid_constant__ FastTmaMaps mapsPyr,
                                                                       const __grid_constant__ FastTmaMaps mapsBlur, const int f0,
                                                                       const uint32_t* __restrict__ slotXY,
                                                                       const uint8_t* __restrict__ slotScore,
                                                                       const int* __restrict__ lvlCount,
                                                                       const uint2* __restrict__ angleTab,
                                                                       const uint8_t* __restrict__ slotLevel, void* __restrict__ kpsOut,
                                                                       uint8_t* __restrict__ descOut, int* __restrict__ kpCount,
                                                                       int kpCap, const __grid_constant__ Geom g) {
    extern __shared__ __align__(128) uint8_t adSmem[];
    const int f = blockIdx.y;
    const int wIn = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = blockIdx.x * DESC_TMA_WARPS + wIn;
    if (warp >= g.slotsPerFrame) return;
    const int l = slotLevel[warp];
    const LevelGeom& L = g.L[l];
    const int p = warp - L.slotOff;
    // keypoints of the levels below this one (output position) and of the frame: one load per lane + a warp scan
    const int* lc = lvlCount + f * g.nlevels;
    const int mine = lane < g.nlevels ? lc[lane] : 0;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < EAOF_MAX_LEVELS; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const int outIdx = p + __shfl_sync(0xffffffffu, incl - mine, l);
    if (warp == 0 && lane == g.nlevels - 1) kpCount[f] = incl;
    if (p >= __shfl_sync(0xffffffffu, mine, l)) return;
    uint8_t* smem = adSmem + ((128u - (smem_u32(adSmem) & 127u)) & 127u);
    uint8_t* boxA = smem + wIn * DESC_WARP_BYTES;
    uint8_t* boxB = boxA + DESC_BOXA_BYTES;
    const uint32_t bar = smem_u32(smem + DESC_TMA_WARPS * DESC_WARP_BYTES + 8 * wIn);

    const uint32_t xy = slotXY[(size_t)f * g.slotsPerFrame + warp];
    const int X = (int)(xy & 0xffff) + EAOF_MIN_BORDER, Y = (int)(xy >> 16) + EAOF_MIN_BORDER;  // :841-842
    const int cxA = EAOF_INNER_X0 + X - EAOF_HALF_PATCH;  // byte column of u = -15 inside the padded row
    const int cxB = EAOF_INNER_X0 + X - EAOF_EDGE;        // byte column of the leftmost tap column
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(bar, (uint32_t)(DESC_BOXA_W * DESC_BOXA_H + DESC_BOXB_W * DESC_BOXB_H));
        tma_load_3d(smem_u32(boxA), &mapsPyr.m[l][0], cxA & ~15, EAOF_EDGE + Y - EAOF_HALF_PATCH, f0 + f, bar);
        tma_load_3d(smem_u32(boxB), &mapsBlur.m[l][0], cxB & ~15, Y, f0 + f, bar);
    }
    __syncwarp();
    const int a = cxA & 3;
    const uint2* wt = angleTab + a * EAOF_ANGLE_TASKS_PAD;
    uint2 wgt[(EAOF_ANGLE_TASKS + 31) / 32];
#pragma unroll
    for (int k = 0; k < (EAOF_ANGLE_TASKS + 31) / 32; ++k) wgt[k] = __ldg(wt + min(lane + 32 * k, EAOF_ANGLE_TASKS_PAD - 1));
    const int4 pa = reinterpret_cast<const int4*>(d_pattern)[2 * lane];
    const int4 pb = reinterpret_cast<const int4*>(d_pattern)[2 * lane + 1];
    mbar_wait(bar, 0);

    int m10 = 0, m01 = 0;
    {
        const uint32_t* wA = reinterpret_cast<const uint32_t*>(boxA) + (((cxA & 15) - a) >> 2);
#pragma unroll
        for (int k = 0; k < (EAOF_ANGLE_TASKS + 31) / 32; ++k) {
            const int i = lane + 32 * k;
            if (i < EAOF_ANGLE_TASKS) {
                const int r = (i * 57) >> 9;  // i / 9 for i < 288
                const int j = i - 9 * r;
                const unsigned pix = wA[r * (DESC_BOXA_W / 4) + j];
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m10) : "r"(pix), "r"(wgt[k].x));
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m01) : "r"(pix), "r"(wgt[k].y));
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float ca, sb;
    glibc_sincosf(__fmul_rn(angle, factorPI), &sb, &ca);
    // Tap coordinates: round-to-nearest-even of an fp32 value below 2^22 = low bits of (value + 1.5 * 2^23): one FADD on the
    // FMA pipe instead of an F2I on the conversion pipe (70 % busy once the gathers were cheap).  The 0x4B400000 biases of row
    // and column are folded into the base address (shared-memory addresses are 32-bit: the sum wraps).
    const float magic = 12582912.f;
    const uint32_t bcAddr = smem_u32(boxB) + EAOF_EDGE * DESC_BOXB_W + (cxB & 15) + EAOF_EDGE - 0x4B400000u * (DESC_BOXB_W + 1u);
    const int words[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x0 = (float)(signed char)(words[k] & 0xff), y0 = (float)(signed char)((words[k] >> 8) & 0xff);
        const float x1 = (float)(signed char)((words[k] >> 16) & 0xff), y1 = (float)(signed char)(words[k] >> 24);
        const uint32_t r0 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(x0, sb), __fmul_rn(y0, ca)), magic));
        const uint32_t c0 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sb)), magic));
        const uint32_t r1 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(x1, sb), __fmul_rn(y1, ca)), magic));
        const uint32_t c1 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sb)), magic));
        uint32_t t0, t1;
        asm("ld.shared.u8 %0, [%1];" : "=r"(t0) : "r"(bcAddr + r0 * DESC_BOXB_W + c0));
        asm("ld.shared.u8 %0, [%1];" : "=r"(t1) : "r"(bcAddr + r1 * DESC_BOXB_W + c1));
        val |= (t0 < t1) << k;
    }
    if (outIdx < kpCap) {
        descOut[((size_t)f * kpCap + outIdx) * 32 + lane] = (uint8_t)val;
        if (lane == 0) {
            float* o = reinterpret_cast<float*>(kpsOut) + ((size_t)f * kpCap + outIdx) * 6;
            float px = (float)X, py = (float)Y;
            if (l != 0) { px = __fmul_rn(px, L.scale); py = __fmul_rn(py, L.scale); }
            o[0] = px;
            o[1] = py;
            o[2] = L.kpSize;
            o[3] = angle;
            o[4] = (float)slotScore[(size_t)f * g.slotsPerFrame + warp];
            reinterpret_cast<int*>(o)[5] = l;
        }
    }
}

}  // namespace eaof
