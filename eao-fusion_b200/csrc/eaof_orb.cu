// libeaof_orb.so — host side of the extractor C ABI (include/eaof_orb.h).  Builds the geometry the reference
// derives per frame (src/ORBextractor.cc:410-470 constructor tables, :1107-1118 level sizes, :773-806 cell grid),
// owns the device workspace and issues the kernel sequence of orb_kernels.cuh on the handle's stream.
// There is deliberately no CPU implementation here: every failure to reach the GPU is an error.
#include <cuda.h>  // CUtensorMap and its enums; cuTensorMapEncodeTiled is fetched through the runtime (no -lcuda)
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only; ranges are no-ops unless a tool (ncu --nvtx, nsys) injects itself

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/eaof_orb.h"
#include "orb_kernels.cuh"

namespace {

thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(EAOF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline int cv_round_f(float v) { return (int)lrintf(v); }  // cvRound: round-half-even
inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

}  // namespace

#ifdef EAOF_TMA_DEBUG
static int* g_tmaDbgHost = nullptr;
extern "C" int* eaof_debug_tma_records() { return g_tmaDbgHost; }
#endif
// Which FAST kernel a handle uses unless $EAOF_FAST_TMA says otherwise (0 LDG-staged, 1 persistent TMA, 2 one-shot TMA);
// the measurements behind the default are in DESIGN.md §4.
#ifndef EAOF_FUSED_MAX_CTAS
#define EAOF_FUSED_MAX_CTAS 296  // two tiles per SM
#endif
#ifndef EAOF_FAST_TMA_DEFAULT
#define EAOF_FAST_TMA_DEFAULT 0
#endif
struct eaof_orb {
    eaof_orb_params p{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // side stream: the blur runs beside FAST + quadtree (both only need the pyramid)
    cudaEvent_t evPyr = nullptr, evBlur = nullptr;
    // host-buffer pipeline: upload, kernels and download of consecutive chunks of a batch overlap on three streams
    cudaStream_t streamIn = nullptr, streamOut = nullptr;
    static const int kMaxChunks = 64;
    cudaEvent_t evIn[kMaxChunks] = {}, evDone[kMaxChunks] = {};
    cudaEvent_t evOutIdle = nullptr;
    int chunkFrames = 0;
    // a matcher that reads dKps/dDesc/dKpCount in place on its own stream leaves this event behind; the next batch's
    // k_angle_desc (the only writer of those buffers) waits for it
    cudaEvent_t evReader = nullptr;
    bool readerPending = false;
    // the same for a consumer of the pyramid block (eaof_stereo_matches reads both cameras' level images): the next
    // batch's level-0 pass waits for it
    cudaEvent_t evPyrReader = nullptr;
    bool pyrReaderPending = false;
    int* dSad = nullptr;  // eaof_stereo_matches scratch, [max_batch][kpCap], lazily allocated
    // batch issued by eaof_orb_extract_batch_async and not yet collected
    int pendN = 0, pendCap = 0;
    eaof_kp* pendKps = nullptr;
    uint8_t* pendDesc = nullptr;
    bool pendDirect = false;
    bool pendGraph = false;
    // single-frame latency path: the kernel sequence + the three result downloads of one frame captured once as a CUDA
    // graph (one cudaGraphLaunch per frame instead of ~15 launches / copies / event calls)
    cudaGraph_t graph1 = nullptr;
    cudaGraphExec_t graphExec1 = nullptr;
    bool graphTried = false;
    bool capturing = false;
    Geom g{};
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> quota;
    int kpCap = 0;
    // device
    uint8_t* dIn = nullptr;      // staging for host-input calls: max_batch frames
    uint8_t* dAux = nullptr;     // lazily allocated staging for host colour frames / host depth maps
    size_t auxBytes = 0;
    float *dURight = nullptr, *dDepthKp = nullptr;  // ComputeStereoFromRGBD outputs, [max_batch][kpCap], lazily allocated
    int colorCh = 0, colorK[3] = {0, 0, 0}, colorShift = 0;  // set by the colour entry points around run_batch
    uint8_t* dPyr = nullptr;     // max_batch pyramid blocks
    uint8_t* dBlur = nullptr;    // same layout, blurred inner levels
    int* dTabs = nullptr;        // resize coefficient tables
    uint2* dAngleTab = nullptr;  // IC_Angle weights: [alignment 0..3][row*9 + word] = (u weights, v weights) as s8x4
    CellDesc* dCells = nullptr;
    uint32_t* dCand = nullptr;
    uint16_t* dLabel = nullptr;
    uint32_t* dCandCount = nullptr;  // [batch][nlevels]
    uint32_t* dSlotXY = nullptr;
    uint8_t* dSlotScore = nullptr;
    int* dLvlCount = nullptr;
    eaof_kp* dKps = nullptr;
    uint8_t* dDesc = nullptr;
    int* dKpCount = nullptr;
    // pinned host staging
    uint8_t* hIn = nullptr;
    eaof_kp* hKps = nullptr;
    uint8_t* hDesc = nullptr;
    int* hKpCount = nullptr;
    size_t octSmem = 0;
    int octKeyCap[3] = {0, 0, 0};  // keys k_octree<256/512/1024> holds in shared memory
    int octForce = -1;
    // k_fast_tma: one tensor map per pyramid level over the handle's whole pyramid buffer, the work counter, launch shape
    eaof::FastTmaMaps fastMaps{};
    unsigned int* dFastCtr = nullptr;   // [kMaxChunks]
    eaof::FastTmaArgs fastT{};
    // k_pyramid_fused: tile plan (built at create), input tensor map (re-encoded when the input pointer / pitches change)
    bool fused = false;
    eaof::FusedArgs fusedA{};
    eaof::FusedAxis* dFusedAxes = nullptr;
    size_t fusedSmem = 0;
    eaof::FastTmaMaps fusedInMap{};
    const uint8_t* fusedMapPtr = nullptr;
    size_t fusedMapStride = 0, fusedMapPitch = 0;
    int fusedMapN = 0;
    void* encodeTiled = nullptr;  // cuTensorMapEncodeTiled
    int fastTma = 0;  // 0: k_fast (LDG-staged tile), 1: k_fast_tma (persistent, double-buffered TMA), 2: k_fast_tma1
    void (*fastKernel)(const uint8_t*, const CellDesc*, uint32_t*, uint32_t*, const Geom, int, int) = nullptr;  // k_fast<PW> of this geometry
    // single-frame latency path: per-level DAG inside the captured graph (level l -> FAST of level l -> quadtree of level l on
    // a stream of its own), EAOF_LAT_DAG=0 keeps the stage-by-stage order
    bool latDag = true;
    cudaStream_t dagStream[EAOF_MAX_LEVELS] = {};
    cudaEvent_t evLvl[EAOF_MAX_LEVELS] = {}, evOct[EAOF_MAX_LEVELS] = {};
    bool rszWindow[EAOF_MAX_LEVELS] = {};  // level l: k_resize's 8-byte source window covers every group of 4 columns
    int rszBulkRows[EAOF_MAX_LEVELS] = {};      // k_resize_bulk: destination rows per CTA (0: level not eligible)
    size_t rszBulkSmem[EAOF_MAX_LEVELS] = {};
    bool bulkPyr = true;                        // EAOF_PYR_BULK=0: per-thread loads (k_level0 / k_resize) for A/B runs
    eaof::FastTmaMaps descMapsPyr{}, descMapsBlur{};  // k_angle_desc_tma: patch boxes of the unblurred / blurred levels
    double latT[8] = {};  // single-frame path: host timestamps (s) at enter / upload queued / graph launched / wait entered / synced / copied out
    int fusedMode = 1;  // EAOF_PYR_FUSED: 0 never, 1 for small batches, 2 always (read at create)
    bool descTma = false;
    uint8_t* dSlotLevel = nullptr;
    bool fastGeneric = false;  // k_fast_generic instead of k_fast (geometry / thresholds outside what fast_cell_rows covers)
    int fastTmaGrid = 0;
    size_t fastTmaSmem = 0;
    bool profiling = false;
    cudaEvent_t ev[7] = {};
    float stageMs[6] = {};
    int lastLaunches = 0;
    int lastFrames = 0;
};

namespace {

// ORBextractor::ORBextractor, src/ORBextractor.cc:410-446 (scale tables + per-level quotas)
void build_tables(eaof_orb* c) {
    const int n = c->p.nlevels;
    const double scaleFactor = c->p.scale_factor;  // the member is a double initialised from the float argument
    c->scale.assign(n, 1.f);
    c->sigma2.assign(n, 1.f);
    for (int i = 1; i < n; ++i) {
        c->scale[i] = (float)(c->scale[i - 1] * scaleFactor);
        c->sigma2[i] = c->scale[i] * c->scale[i];
    }
    c->invScale.resize(n);
    c->invSigma2.resize(n);
    for (int i = 0; i < n; ++i) {
        c->invScale[i] = 1.0f / c->scale[i];
        c->invSigma2[i] = 1.0f / c->sigma2[i];
    }
    c->quota.assign(n, 0);
    const float factor = (float)(1.0f / scaleFactor);
    float nDesired = c->p.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n));
    int sum = 0;
    for (int l = 0; l < n - 1; ++l) {
        c->quota[l] = cv_round_f(nDesired);
        sum += c->quota[l];
        nDesired *= factor;
    }
    c->quota[n - 1] = c->p.nfeatures - sum > 0 ? c->p.nfeatures - sum : 0;
}

int build_geometry(eaof_orb* c, std::vector<int>& tabs, std::vector<CellDesc>& cells) {
    Geom& g = c->g;
    memset(&g, 0, sizeof g);
    g.nlevels = c->p.nlevels;
    g.W = c->p.width;
    g.H = c->p.height;
    g.iniTh = c->p.ini_th_fast;
    g.minTh = c->p.min_th_fast;
    g.blurMode = c->p.blur_mode;
    uint64_t off = 0;
    uint32_t candOff = 0;
    int slotOff = 0;
    int fastRows = 7, fastWords = 2, fastList = 160, fastOut = 1, fastInnerWords = 1;
    for (int l = 0; l < g.nlevels; ++l) {
        LevelGeom& L = g.L[l];
        L.w = cv_round_f((float)g.W * c->invScale[l]);  // :1112
        L.h = cv_round_f((float)g.H * c->invScale[l]);
        if (L.w < 1 || L.h < 1) return fail(EAOF_ERR_UNSUPPORTED, "pyramid level %d is empty (%dx%d)", l, L.w, L.h);
        L.pitch = (L.w + 64 + 63) & ~63;
        L.rows = L.h + 2 * EAOF_EDGE;
        L.off = (uint32_t)off;
        off += ((uint64_t)L.pitch * L.rows + 255) & ~(uint64_t)255;
        // cell grid, :773-787
        const int maxBX = L.w - EAOF_EDGE + 3, maxBY = L.h - EAOF_EDGE + 3;
        L.winW = maxBX - EAOF_MIN_BORDER;
        L.winH = maxBY - EAOF_MIN_BORDER;
        const float width = (float)L.winW, height = (float)L.winH;
        L.nCols = (int)(width / 30.f);
        L.nRows = (int)(height / 30.f);
        if (L.nCols < 1 || L.nRows < 1) {
            L.nCols = L.nRows = 0;  // the reference's cell loops do not execute
            L.wCell = L.hCell = 1;
        } else {
            L.wCell = (int)ceilf(width / L.nCols);
            L.hCell = (int)ceilf(height / L.nRows);
        }
        // root nodes, :543-545.  winH == 0 or nIni < 1 is undefined behaviour in the reference.
        if (L.winH == 0) return fail(EAOF_ERR_UNSUPPORTED, "level %d: detection window height 0 (reference divides by zero)", l);
        L.nIni = (int)roundf((float)L.winW / (float)L.winH);
        if (L.nIni < 1) {
            if (L.nCols > 0) return fail(EAOF_ERR_UNSUPPORTED, "level %d: aspect ratio gives 0 root nodes (reference indexes an empty vector)", l);
            L.nIni = 0;
        }
        L.hX = L.nIni > 0 ? (float)L.winW / L.nIni : 1.f;
        L.quota = c->quota[l];
        L.nodeCap = L.quota + 3 > 4 * L.nIni ? L.quota + 3 : 4 * L.nIni;
        if (L.nodeCap < 4) L.nodeCap = 4;
        L.scale = c->scale[l];
        L.kpSize = (float)(int)(31 * c->scale[l]);  // :835
        L.slotOff = slotOff;
        slotOff += L.nodeCap;
        if (L.nodeCap > g.maxNodeCap) g.maxNodeCap = L.nodeCap;
        // cells, :789-806 (same skips and clipping)
        L.cellOff = (int)cells.size();
        uint32_t cap = 0;
        for (int i = 0; i < L.nRows; ++i) {
            const int iniY = EAOF_MIN_BORDER + i * L.hCell;
            int maxY = iniY + L.hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = maxBY;
            for (int j = 0; j < L.nCols; ++j) {
                const int iniX = EAOF_MIN_BORDER + j * L.wCell;
                int maxX = iniX + L.wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = maxBX;
                const int cw = maxX - iniX, ch = maxY - iniY;
                if (cw < 7 || ch < 7) continue;  // cv::FAST finds nothing in fewer than 7 rows/cols
                if (cw > 66 || ch > 66) return fail(EAOF_ERR_UNSUPPORTED, "cell larger than 66 px");
                cells.push_back(CellDesc{(short)l, (short)iniX, (short)iniY, (short)cw, (short)ch, 0});
                {
                    // k_fast shared-memory extents: tile rows/words, pair-list entries, NMS survivors
                    const int mis = (EAOF_INNER_X0 + iniX) & 3;
                    const int nwords = (mis + cw + 3) >> 2;
                    const int nW = ((mis + cw - 4) >> 2) - ((mis + 3) >> 2) + 1;
                    fastRows = std::max(fastRows, ch);
                    fastWords = std::max(fastWords, nwords);
                    fastInnerWords = std::max(fastInnerWords, nW);
#ifdef FAST_LIST_CAP
                    fastList = FAST_LIST_CAP;
#else
                    fastList = std::max(fastList, std::max(64 * nW, 160));  // 16 rows of pixels: fast_cell_rows lists 32 rows at once, or two halves
#endif
                    fastOut = std::max(fastOut, ((cw - 6 + 1) / 2) * ((ch - 6 + 1) / 2));
                }
                cap += (uint32_t)(((cw - 6 + 1) / 2) * ((ch - 6 + 1) / 2));  // NMS keeps no two 8-adjacent pixels
            }
        }
        L.candOff = candOff;
        L.candCap = (cap + 63) & ~63u;
        candOff += L.candCap;
        // resize tables (A.2): per destination column [sx, a0|a1<<16], per row [sy, b0|b1<<16]
        L.xTab = (int)tabs.size();
        if (l > 0) {
            const LevelGeom& S = g.L[l - 1];
            const double sx_ = 1.0 / ((double)L.w / S.w), sy_ = 1.0 / ((double)L.h / S.h);
            for (int dx = 0; dx < L.w; ++dx) {
                float fx = (float)((dx + 0.5) * sx_ - 0.5);
                int sx = cv_floor_f(fx);
                fx -= sx;
                if (sx < 0) { fx = 0; sx = 0; }
                if (sx >= S.w - 1) { fx = 0; sx = S.w - 1; }
                const int a0 = sat_short(cv_round_f((1.f - fx) * 2048.f)), a1 = sat_short(cv_round_f(fx * 2048.f));
                tabs.push_back(sx);
                tabs.push_back((a0 & 0xffff) | (a1 << 16));
            }
            L.yTab = (int)tabs.size();
            for (int dy = 0; dy < L.h; ++dy) {
                float fy = (float)((dy + 0.5) * sy_ - 0.5);
                const int sy = cv_floor_f(fy);
                fy -= sy;
                const int b0 = sat_short(cv_round_f((1.f - fy) * 2048.f)), b1 = sat_short(cv_round_f(fy * 2048.f));
                tabs.push_back(sy);
                tabs.push_back((b0 & 0xffff) | (b1 << 16));
            }
        } else {
            L.yTab = L.xTab;
        }
        // k_resize reads the source columns of 4 destination columns from an 8-byte window
        c->rszWindow[l] = true;
        if (l > 0) {
            auto refl = [&](int p) { const int n = L.w; if (n == 1) return 0; while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p; return p; };
            for (int c0 = 12; c0 < L.w + 52 && c->rszWindow[l]; c0 += 4) {
                int lo = 1 << 30, hi = -1;
                for (int j = 0; j < 4; ++j) {
                    const int sx = tabs[L.xTab + 2 * refl(c0 - EAOF_INNER_X0 + j)];
                    lo = std::min(lo, sx);
                    hi = std::max(hi, sx);
                }
                if (hi + 1 - lo > 7) c->rszWindow[l] = false;
            }
        }
        // k_resize_bulk: shared memory for the source rows of a chunk of 16 (or 8) destination rows
        c->rszBulkRows[l] = 0;
        if (l > 0) {
            const LevelGeom& S = g.L[l - 1];
            const int srcBytes = (S.w + 12 + 15) & ~15;
            for (int rows = 16; rows >= 8 && !c->rszBulkRows[l]; rows >>= 1) {
                int nsMax = 0;
                for (int y0 = 0; y0 < L.h; y0 += rows) {
                    const int y1 = std::min(y0 + rows, L.h) - 1;
                    const int a = std::min(std::max(tabs[L.yTab + 2 * y0], 0), S.h - 1), b = std::min(std::max(tabs[L.yTab + 2 * y1] + 1, 0), S.h - 1);
                    nsMax = std::max(nsMax, b - a + 1);
                }
                if ((size_t)nsMax * srcBytes <= 40 * 1024) {
                    c->rszBulkRows[l] = rows;
                    c->rszBulkSmem[l] = (size_t)nsMax * srcBytes;
                }
            }
        }
        L.blurTaskOff = g.blurTasksPerFrame;
        g.blurTasksPerFrame += ((L.w + 3) / 4) * ((L.h + BLUR_ROWS - 1) / BLUR_ROWS);
        if (L.winW + 3 > 4095 || L.winH + 3 > 4095) return fail(EAOF_ERR_UNSUPPORTED, "frames larger than 4096 px are not supported");
    }
    g.fastPW = std::max((fastWords + 2) | 1, 11);  // one pad word on the left, one spare on the right (phase A reads word w+1), odd pitch; k_fast<PW> is instantiated for 11..21
    g.fastMapWords = std::max(fastRows * g.fastPW + 2, fastOut);
    g.fastMapWords = (g.fastMapWords + 3) & ~3;
    g.fastWarpWords = (2 * g.fastMapWords + FAST_CLST2 / 2 + (fastList + 1) / 2 + 3) & ~3;
    // fast_cell_rows (k_fast) covers rows of <= 16 inner words and minThFAST < iniThFAST < 128; anything else runs k_fast_generic
    c->fastGeneric = fastInnerWords > 16 || g.fastPW > 21 || g.iniTh >= 128 || g.minTh >= 128 || g.iniTh < 0 || g.minTh < 0;
    if (const char* e = getenv("EAOF_FAST_GENERIC")) if (*e) c->fastGeneric = atoi(e) != 0;
    if (const char* e = getenv("EAOF_PYR_BULK")) if (*e) c->bulkPyr = atoi(e) != 0;
    g.pyrFrameBytes = off + 4096;  // slack: tile loaders may read a few bytes past the last row
    g.candPerFrame = candOff;
    g.slotsPerFrame = slotOff;
    g.cellsPerFrame = (int)cells.size();
    if (g.maxNodeCap > 60000) return fail(EAOF_ERR_UNSUPPORTED, "nfeatures too large for 16-bit node labels");
    return EAOF_OK;
}

// Tile plan of k_pyramid_fused along one axis: own ranges [a, b) per (level, tile) that nest through the resize tables
// (a_{l-1} = source index of a_l), and the needed ranges [lo, lo+n) = own range + what the deeper levels read.
bool fused_axis(const Geom& g, bool xAxis, const std::vector<int>& tabs, int& nT, std::vector<eaof::FusedAxis>& out) {
    const int nl = g.nlevels;
    auto size = [&](int l) { return xAxis ? g.L[l].w : g.L[l].h; };
    auto src = [&](int l, int d) {  // source index in level l-1 of destination index d of level l, clamped like the kernels do
        const int v = tabs[(xAxis ? g.L[l].xTab : g.L[l].yTab) + 2 * d];
        return std::min(std::max(v, 0), size(l - 1) - 1);
    };
    const int last = nl - 1;
    const int nLast = size(last);
    static const int tilePx = getenv("EAOF_FUSED_TILE") ? std::max(20, atoi(getenv("EAOF_FUSED_TILE"))) : 22;  // last-level tile edge: 22 px = 48 CTAs per 640x480 frame (41 -> 35 us for one frame against 32 px)
    nT = std::max(1, std::min((nLast + tilePx / 2) / tilePx, nLast / 20));
    std::vector<std::vector<int>> a(nl, std::vector<int>(nT + 1));
    for (int t = 0; t <= nT; ++t) {
        a[last][t] = (int)((long long)t * nLast / nT);
        if (xAxis && t > 0 && t < nT) a[last][t] &= ~3;
    }
    for (int l = last; l >= 1; --l)
        for (int t = 0; t <= nT; ++t)  // x boundaries on multiples of 4: whole words change hands between tiles
            a[l - 1][t] = t == 0 ? 0 : t == nT ? size(l - 1) : (xAxis ? src(l, a[l][t]) & ~3 : src(l, a[l][t]));
    out.assign((size_t)nl * nT, eaof::FusedAxis{0, 0, 0, 0});
    for (int t = 0; t < nT; ++t) {
        int lo = a[last][t], hi = a[last][t + 1] - 1;
        for (int l = last; l >= 0; --l) {
            if (a[l][t + 1] - a[l][t] < 1) return false;
            // the mirror sources of the 19-px border must lie inside the edge tiles' own ranges
            if (nT > 1 && (t == 0 || t == nT - 1) && a[l][t + 1] - a[l][t] < EAOF_EDGE + 1) return false;
            if (hi - lo + 1 > 30000) return false;
            out[(size_t)l * nT + t] = eaof::FusedAxis{(short)a[l][t], (short)a[l][t + 1], (short)lo, (short)(hi - lo + 1)};
            if (l >= 1) {
                const int lo2 = std::min(a[l - 1][t], src(l, lo));
                const int hi2 = std::max(a[l - 1][t + 1] - 1, std::min(src(l, hi) + 1, size(l - 1) - 1));
                lo = lo2;
                hi = hi2;
            }
        }
    }
    return true;
}

int build_fused_plan(eaof_orb* c, const std::vector<int>& tabs, std::vector<eaof::FusedAxis>& axes) {
    const Geom& g = c->g;
    c->fused = false;
    const char* e = getenv("EAOF_PYR_FUSED");
    c->fusedMode = e && *e ? atoi(e) : 1;
    if (c->fusedMode == 0) return EAOF_OK;
    if (g.nlevels < 2) return EAOF_OK;
    for (int l = 0; l < g.nlevels; ++l)
        if (g.L[l].w < 2 * EAOF_EDGE + 2 || g.L[l].h < 2 * EAOF_EDGE + 2) return EAOF_OK;  // borders fold more than once: per-level kernels
    std::vector<eaof::FusedAxis> ax, ay;
    int nTx = 0, nTy = 0;
    if (!fused_axis(g, true, tabs, nTx, ax) || !fused_axis(g, false, tabs, nTy, ay)) return EAOF_OK;
    eaof::FusedArgs& A = c->fusedA;
    A = eaof::FusedArgs{};
    A.nTx = nTx; A.nTy = nTy;
    size_t bytes[2] = {0, 0};
    for (int l = 0; l < g.nlevels; ++l) {
        int wMax = 0, hMax = 0;
        for (int t = 0; t < nTx; ++t) {
            const eaof::FusedAxis& X = ax[(size_t)l * nTx + t];
            const int ox = l == 0 ? (X.lo & ~15) : (X.lo & ~3);
            wMax = std::max(wMax, X.lo + X.n - ox);
        }
        for (int t = 0; t < nTy; ++t) hMax = std::max(hMax, (int)ay[(size_t)l * nTy + t].n);
        A.pitchT[l] = (wMax + 8 + 15) & ~15;  // + 8: the horizontal pass reads the word after the one holding column sx (zero coefficient at the image edge)
        if (l == 0) { A.boxW0 = A.pitchT[0]; A.boxH0 = hMax; }
        if (hMax > EAOF_FUSED_MAX_ROWS) return EAOF_OK;
        bytes[l & 1] = std::max(bytes[l & 1], (size_t)A.pitchT[l] * (hMax + 1));
    }
    if (A.boxW0 > 256 || A.boxH0 > 256) return EAOF_OK;
    A.bufBytes[0] = (int)((bytes[0] + 127) & ~(size_t)127);
    A.bufBytes[1] = (int)((bytes[1] + 127) & ~(size_t)127);
    c->fusedSmem = (size_t)A.bufBytes[0] + A.bufBytes[1] + 128;
    if (c->fusedSmem > 110 * 1024) return EAOF_OK;
    axes = ax;
    axes.insert(axes.end(), ay.begin(), ay.end());
    c->fused = true;
    return EAOF_OK;
}

// Issues the kernel sequence for frames [f0, f0+n) of the workspace; dImgs points at frame f0's image.
// Kernels of different handles that run at the same time slow each other down on B200 (measured: two handles
// alternating full batches reach 115 k frames/s against 132 k for one, tools/lanes_experiment.py), while uploads and
// downloads overlap with kernels for free.  The host-buffer pipeline therefore passes a per-device token from batch to
// batch: a batch's kernels start only after the kernels of the batch issued before it, whichever handle issued it.
std::mutex g_tokMu;
cudaEvent_t g_tok[64] = {};
bool g_tokSet[64] = {};

int run_batch(eaof_orb* c, const uint8_t* dImgs, int n, size_t stride, size_t framePitch, int f0 = 0, int chunkIdx = 0) {
    const Geom& g = c->g;
    cudaStream_t s = c->stream;
    // every per-frame array is frame-major, so a chunk is addressed by offsetting the base pointers
    uint8_t* const dPyr = c->dPyr + (size_t)f0 * g.pyrFrameBytes;
    uint8_t* const dPyrBase = c->dPyr;  // k_pyramid_fused offsets by f0 itself
    uint8_t* const dBlur = c->dBlur + (size_t)f0 * g.pyrFrameBytes;
    uint32_t* const dCand = c->dCand + (size_t)f0 * g.candPerFrame;
    uint16_t* const dLabel = c->dLabel + (size_t)f0 * g.candPerFrame;
    uint32_t* const dCandCount = c->dCandCount + (size_t)f0 * g.nlevels;
    uint32_t* const dSlotXY = c->dSlotXY + (size_t)f0 * g.slotsPerFrame;
    uint8_t* const dSlotScore = c->dSlotScore + (size_t)f0 * g.slotsPerFrame;
    int* const dLvlCount = c->dLvlCount + (size_t)f0 * g.nlevels;
    eaof_kp* const dKps = c->dKps + (size_t)f0 * c->kpCap;
    uint8_t* const dDesc = c->dDesc + (size_t)f0 * c->kpCap * 32;
    int* const dKpCount = c->dKpCount + f0;
    int launches = 0;
    const bool prof = c->profiling;
    if (prof) CK(cudaEventRecord(c->ev[0], s));
    nvtxRangePushA("eaof:pyramid");
    CK(cudaMemsetAsync(dCandCount, 0, sizeof(uint32_t) * (size_t)n * g.nlevels, s));
    if (c->pyrReaderPending && !c->capturing) {
        CK(cudaStreamWaitEvent(s, c->evPyrReader, 0));
        c->pyrReaderPending = false;
    }
    // One launch for the whole pyramid pays off while the frames of a call cannot fill the GPU level by level (the chain of
    // per-level launches is then pure latency); big batches keep the per-level kernels, which do less redundant work
    // (measured crossover: DESIGN.md §4).  EAOF_PYR_FUSED=2 forces the fused kernel for every batch size.
    // latency path (graph capture of one frame): per-level DAG, which needs the per-level pyramid kernels
    const bool dagNow = c->capturing && n == 1 && c->latDag && c->fastTma == 0 && !prof && g.cellsPerFrame > 0;
    const bool fusedNow = !dagNow && c->fused && c->colorCh == 0 && (c->fusedMode == 2 || n * c->fusedA.nTx * c->fusedA.nTy <= EAOF_FUSED_MAX_CTAS);
    if (fusedNow) {
        eaof::FusedArgs A = c->fusedA;
        A.in = dImgs; A.stride = stride; A.framePitch = framePitch; A.f0 = f0;
        // TMA takes the level-0 boxes when the frames are 16-byte aligned in every dimension; the map is re-encoded only
        // when the input changes (the host-buffer paths always read the handle's own staging buffer)
        A.useTma = c->encodeTiled && (reinterpret_cast<uintptr_t>(dImgs) & 15) == 0 && (stride & 15) == 0 && (framePitch & 15) == 0;
        if (A.useTma && (c->fusedMapPtr != dImgs || c->fusedMapStride != stride || c->fusedMapPitch != framePitch || c->fusedMapN < n)) {
            typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            CUtensorMap m;
            const cuuint64_t dims[3] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)n};
            const cuuint64_t strides[2] = {(cuuint64_t)stride, (cuuint64_t)framePitch};
            const cuuint32_t box[3] = {(cuuint32_t)A.boxW0, (cuuint32_t)A.boxH0, 1}, es[3] = {1, 1, 1};
            const CUresult r = ((EncodeTiled)c->encodeTiled)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(dImgs), dims, strides,
                                                             box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) {
                memcpy(&c->fusedInMap.m[0][0], &m, sizeof m);
                c->fusedMapPtr = dImgs; c->fusedMapStride = stride; c->fusedMapPitch = framePitch; c->fusedMapN = n;
            } else {
                A.useTma = 0;
            }
        }
        // few frames: wide CTAs shorten the per-level critical path of a tile; many frames: narrow ones fill the SMs better
        static const int forced = getenv("EAOF_FUSED_THREADS") ? atoi(getenv("EAOF_FUSED_THREADS")) : 0;
        const int threads = forced ? forced : (n * A.nTx * A.nTy < 1024 ? 512 : 256);
        eaof::k_pyramid_fused<<<dim3(A.nTx * A.nTy, n), threads, c->fusedSmem, s>>>(c->fusedInMap, A, dPyrBase, c->dTabs, g);
        ++launches;
    } else {
        const LevelGeom& L = g.L[0];
        dim3 b(64, 4), gr(((L.w + 43) / 4 + 63) / 64, (L.rows + 3) / 4, n);
        const dim3 gr8(gr.x, (L.rows + 4 * LEVEL0_ROWS - 1) / (4 * LEVEL0_ROWS), n);  // k_level0: LEVEL0_ROWS rows per thread
        if (c->colorCh == 3)
            eaof::k_level0_color<3><<<gr, b, 0, s>>>(dImgs, framePitch, stride, dPyr, g, c->colorK[0], c->colorK[1], c->colorK[2], c->colorShift);
        else if (c->colorCh == 4)
            eaof::k_level0_color<4><<<gr, b, 0, s>>>(dImgs, framePitch, stride, dPyr, g, c->colorK[0], c->colorK[1], c->colorK[2], c->colorShift);
        else if (c->bulkPyr && L.w % 16 == 0 && L.w >= 20 && L.h >= 20 && ((reinterpret_cast<uintptr_t>(dImgs) | stride | framePitch) & 15) == 0) {
            // copyMakeBorder as bulk asynchronous copies through shared memory
            const int span = (EAOF_INNER_X0 + L.w + EAOF_EDGE + 15) & ~15;
            int rows = span * 16 <= 40 * 1024 ? 16 : span * 8 <= 40 * 1024 ? 8 : 4;
            while (rows > 4 && (long long)n * ((L.h + rows - 1) / rows) < 296) rows >>= 1;
            eaof::k_level0_bulk<<<dim3((L.h + rows - 1) / rows, n), 128, (size_t)rows * span, s>>>(dImgs, framePitch, stride, dPyr, g, rows);
        } else
            eaof::k_level0<<<gr8, b, 0, s>>>(dImgs, framePitch, stride, dPyr, g);
        ++launches;
    }
    auto launch_resize = [&](int l, cudaStream_t st) {
        const LevelGeom& L = g.L[l];
        if (L.h >= 40 && c->rszWindow[l] && c->bulkPyr && c->rszBulkRows[l]) {
            // source rows staged in shared memory by bulk asynchronous copies, CTA = chunk of rows over the whole width
            const int nCW = (L.w + 43) / 4;
            int rows = c->rszBulkRows[l];  // few frames: shorter chunks, so that the level still fills the SMs
            while (rows > 4 && (long long)n * ((L.h + rows - 1) / rows) < 296) rows >>= 1;
            const int threads = std::min(256, (nCW + 31) & ~31);
            const dim3 gr((L.h + rows - 1) / rows, n);
            if (rows == 16) eaof::k_resize_bulk<16><<<gr, threads, c->rszBulkSmem[l], st>>>(dPyr, c->dTabs, g, l);
            else if (rows == 8) eaof::k_resize_bulk<8><<<gr, threads, c->rszBulkSmem[l], st>>>(dPyr, c->dTabs, g, l);
            else eaof::k_resize_bulk<4><<<gr, threads, c->rszBulkSmem[l], st>>>(dPyr, c->dTabs, g, l);
        } else if (L.h >= 40 && c->rszWindow[l]) {
            // rows per thread: long walks reuse source rows, but small levels / small batches need the threads
#ifndef RSZ_WANT
#define RSZ_WANT 600000
#endif
            const long long want = RSZ_WANT;
            const int nCW = (L.w + 43) / 4;
            int rows = 32;
            while (rows > 8 && (long long)nCW * ((L.h + rows - 1) / rows) * n < want) rows >>= 1;
            const int tasks = nCW * ((L.h + rows - 1) / rows);
            const dim3 gr((tasks + RSZ_THREADS - 1) / RSZ_THREADS, n);
            if (rows == 32) eaof::k_resize<32><<<gr, RSZ_THREADS, 0, st>>>(dPyr, c->dTabs, g, l);
            else if (rows == 16) eaof::k_resize<16><<<gr, RSZ_THREADS, 0, st>>>(dPyr, c->dTabs, g, l);
            else eaof::k_resize<8><<<gr, RSZ_THREADS, 0, st>>>(dPyr, c->dTabs, g, l);
        } else {  // tiny levels: border rows may fold more than once
            dim3 b(64, 4), gr(((L.w + 43) / 4 + 63) / 64, (L.rows + 3) / 4, n);
            eaof::k_resize_generic<<<gr, b, 0, st>>>(dPyr, c->dTabs, g, l);
        }
        ++launches;
        };
    auto launch_fast = [&](int cellBegin, int cellEnd, cudaStream_t st) {
        const size_t smem = (size_t)FAST_WARPS * g.fastWarpWords * 4;
        const dim3 gr((cellEnd - cellBegin + FAST_WARPS - 1) / FAST_WARPS, n);
        if (c->fastGeneric) eaof::k_fast_generic<<<gr, FAST_WARPS * 32, smem, st>>>(dPyr, c->dCells, dCand, dCandCount, g, cellBegin, cellEnd);
        else c->fastKernel<<<gr, FAST_WARPS * 32, smem, st>>>(dPyr, c->dCells, dCand, dCandCount, g, cellBegin, cellEnd);
    };
    auto launch_octree = [&](int levelBegin, int nLevels, cudaStream_t st) {
        // CTA width by how many (frame, level) CTAs there are; keys + labels in shared memory up to the width's budget
        const int ctas = n * g.nlevels;
        const int v = c->octForce >= 0 ? c->octForce : ctas <= 296 ? 2 : ctas <= 592 ? 1 : 0;
        const int keyCap = c->octKeyCap[v];
        const size_t smem = c->octSmem + (size_t)keyCap * 6;
        const dim3 gr(nLevels, n);
        if (v == 2) eaof::k_octree<1024><<<gr, 1024, smem, st>>>(dCand, dCandCount, dLabel, dSlotXY, dSlotScore, dLvlCount, g, keyCap, levelBegin);
        else if (v == 1) eaof::k_octree<512><<<gr, 512, smem, st>>>(dCand, dCandCount, dLabel, dSlotXY, dSlotScore, dLvlCount, g, keyCap, levelBegin);
        else eaof::k_octree<256><<<gr, 256, smem, st>>>(dCand, dCandCount, dLabel, dSlotXY, dSlotScore, dLvlCount, g, keyCap, levelBegin);
    };
    if (dagNow) CK(cudaEventRecord(c->evLvl[0], s));
    for (int l = 1; l < g.nlevels && !fusedNow; ++l) {
        launch_resize(l, s);
        if (dagNow) CK(cudaEventRecord(c->evLvl[l], s));
    }
    nvtxRangePop();
    if (prof) CK(cudaEventRecord(c->ev[1], s));
    nvtxRangePushA("eaof:fast+blur+octree");
    // FAST is bound by the integer ALU pipe, the blur by the FMA pipe, the quadtree by latency: outside profiling mode
    // (which serialises the stages to time them) the blur runs on the side stream beside FAST + quadtree.
    static const bool serialBlur = getenv("EAOF_SERIAL_BLUR") != nullptr;  // experiment knob
    const bool side = !prof && !serialBlur;
    cudaStream_t sb = side ? c->stream2 : s;
    if (side) {
        CK(cudaEventRecord(c->evPyr, s));
        CK(cudaStreamWaitEvent(sb, c->evPyr, 0));
        eaof::k_blur<<<dim3((g.blurTasksPerFrame + BLUR_THREADS - 1) / BLUR_THREADS, n), BLUR_THREADS, 0, sb>>>(dPyr, dBlur, g);
        ++launches;
        CK(cudaEventRecord(c->evBlur, sb));
    }
    if (g.cellsPerFrame > 0 && c->fastTma == 2) {
        // one cell per one-warp CTA, tile by TMA, no persistent scheduling
        eaof::FastTmaArgs A = c->fastT;
        A.f0 = f0;
        const size_t smem = (size_t)2 * A.tileBytes + 2 * FAST_CLST + 2 * A.lstCap + 16 + 128;
        eaof::k_fast_tma1<<<dim3(g.cellsPerFrame, n), 32, smem, s>>>(c->fastMaps, A, c->dCells, dCand, dCandCount, g);
        ++launches;
    } else if (g.cellsPerFrame > 0 && c->fastTma) {
        // persistent warps pulling (frame, cell) items off a counter; the tile of every item arrives by TMA
        unsigned int* ctr = c->dFastCtr + (chunkIdx % eaof_orb::kMaxChunks);
        CK(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), s));
        eaof::FastTmaArgs A = c->fastT;
        A.f0 = f0;
        A.nItems = n * g.cellsPerFrame;
        const int grid = std::min(c->fastTmaGrid, (A.nItems + FASTT_WARPS - 1) / FASTT_WARPS);
        eaof::k_fast_tma<<<grid, FASTT_WARPS * 32, c->fastTmaSmem, s>>>(c->fastMaps, A, c->dCells, dCand, dCandCount, ctr, g);
        ++launches;
    } else if (g.cellsPerFrame > 0 && !dagNow) {
        launch_fast(0, g.cellsPerFrame, s);
        ++launches;
    }
    if (prof) CK(cudaEventRecord(c->ev[2], s));
    if (!dagNow) {
        launch_octree(0, g.nlevels, s);
        ++launches;
    } else {
        // per-level branches: FAST + quadtree of level l start as soon as level l exists (events recorded by the pyramid loop)
        for (int l = 0; l < g.nlevels; ++l) {
            cudaStream_t st = c->dagStream[l];
            CK(cudaStreamWaitEvent(st, c->evLvl[l], 0));
            const int cb = g.L[l].cellOff, ce = l + 1 < g.nlevels ? g.L[l + 1].cellOff : g.cellsPerFrame;
            if (ce > cb) { launch_fast(cb, ce, st); ++launches; }
            launch_octree(l, 1, st);
            ++launches;
            CK(cudaEventRecord(c->evOct[l], st));
        }
        for (int l = 0; l < g.nlevels; ++l) CK(cudaStreamWaitEvent(s, c->evOct[l], 0));
    }
    if (prof) CK(cudaEventRecord(c->ev[3], s));
    if (!side) {
        eaof::k_blur<<<dim3((g.blurTasksPerFrame + BLUR_THREADS - 1) / BLUR_THREADS, n), BLUR_THREADS, 0, s>>>(dPyr, dBlur, g);
        ++launches;
    } else {
        CK(cudaStreamWaitEvent(s, c->evBlur, 0));
    }
    nvtxRangePop();
    if (prof) CK(cudaEventRecord(c->ev[4], s));
    nvtxRangePushA("eaof:angle+descriptor");
    if (c->readerPending && !c->capturing) {
        CK(cudaStreamWaitEvent(s, c->evReader, 0));
        c->readerPending = false;
    }
    {
#ifndef DESC_WARPS
#define DESC_WARPS 8
#endif
        const int warpsPerBlock = DESC_WARPS;
        dim3 gr((g.slotsPerFrame + warpsPerBlock - 1) / warpsPerBlock, n);
        if (c->descTma) {
            const dim3 grT((g.slotsPerFrame + DESC_TMA_WARPS - 1) / DESC_TMA_WARPS, n);
            eaof::k_angle_desc_tma<<<grT, DESC_TMA_WARPS * 32, DESC_TMA_WARPS * (DESC_WARP_BYTES + 8) + 128, s>>>(
                c->descMapsPyr, c->descMapsBlur, f0, dSlotXY, dSlotScore, dLvlCount, c->dAngleTab, c->dSlotLevel, dKps, dDesc, dKpCount, c->kpCap, g);
        } else
        eaof::k_angle_desc<<<gr, warpsPerBlock * 32, 0, s>>>(dPyr, dBlur, dSlotXY, dSlotScore, dLvlCount,
                                                             c->dAngleTab, dKps, dDesc, dKpCount, c->kpCap, g);
        ++launches;
    }
    nvtxRangePop();
    if (prof) CK(cudaEventRecord(c->ev[5], s));
    CK(cudaGetLastError());
    c->lastLaunches = (f0 == 0 ? 0 : c->lastLaunches) + launches;
    c->lastFrames = f0 + n;
    return EAOF_OK;
}

int check_shape(const eaof_orb* c, int width, int height, int n) {
    if (!c) return fail(EAOF_ERR_ARG, "null handle");
    if (width != c->p.width || height != c->p.height)
        return fail(EAOF_ERR_ARG, "frame is %dx%d but the handle was created for %dx%d", width, height, c->p.width, c->p.height);
    if (n < 1 || n > c->p.max_batch) return fail(EAOF_ERR_ARG, "n_frames %d outside [1, max_batch=%d]", n, c->p.max_batch);
    return EAOF_OK;
}

}  // namespace

struct eaof_ring {
    int slots = 0, width = 0, height = 0, channels = 0;
    size_t stride = 0, slotBytes = 0;
    uint8_t* base = nullptr;  // cudaHostAlloc: slots x slotBytes
    std::vector<double> stamps;
    std::atomic<uint64_t> head{0}, tail{0};
};

extern "C" {

const char* eaof_last_error(void) { return g_err.c_str(); }
int eaof_abi_version(void) { return EAOF_ABI_VERSION; }

int eaof_orb_create(const eaof_orb_params* params, int device, eaof_orb** out) {
    if (!params || !out) return fail(EAOF_ERR_ARG, "null argument");
    *out = nullptr;
    const eaof_orb_params& p = *params;
    if (p.nlevels < 1 || p.nlevels > EAOF_MAX_LEVELS) return fail(EAOF_ERR_ARG, "nlevels must be in [1,%d]", EAOF_MAX_LEVELS);
    if (p.nfeatures < 1 || !(p.scale_factor > 1.0f)) return fail(EAOF_ERR_ARG, "nfeatures >= 1 and scaleFactor > 1 required");
    if (p.ini_th_fast < 0 || p.min_th_fast < 0 || p.ini_th_fast > 254 || p.min_th_fast > 254)
        return fail(EAOF_ERR_ARG, "FAST thresholds must be in [0,254]");
    if (p.blur_mode < 0 || p.blur_mode > 2) return fail(EAOF_ERR_ARG, "unknown blur_mode");
    if (p.width < 1 || p.height < 1 || p.max_batch < 1) return fail(EAOF_ERR_ARG, "width, height, max_batch must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(EAOF_ERR_CUDA, "no CUDA device: libeaof_orb has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(EAOF_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));

    eaof_orb* c = new eaof_orb;
    c->p = p;
    c->device = device;
    build_tables(c);
    std::vector<int> tabs;
    std::vector<CellDesc> cells;
    int rc = build_geometry(c, tabs, cells);
    if (rc != EAOF_OK) { delete c; return rc; }
    const Geom& g = c->g;
    c->kpCap = g.slotsPerFrame;
    const size_t B = (size_t)p.max_batch;
#define CKD(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            fail(EAOF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                        \
            eaof_orb_destroy(c);                                                                        \
            return EAOF_ERR_CUDA;                                                                       \
        }                                                                                               \
    } while (0)
    CKD(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CKD(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CKD(cudaEventCreateWithFlags(&c->evPyr, cudaEventDisableTiming));
    CKD(cudaEventCreateWithFlags(&c->evBlur, cudaEventDisableTiming));
    if (const char* e = getenv("EAOF_LAT_DAG")) if (*e) c->latDag = atoi(e) != 0;
    if (c->latDag)
        for (int l = 0; l < p.nlevels && l < EAOF_MAX_LEVELS; ++l) {
            CKD(cudaStreamCreateWithFlags(&c->dagStream[l], cudaStreamNonBlocking));
            CKD(cudaEventCreateWithFlags(&c->evLvl[l], cudaEventDisableTiming));
            CKD(cudaEventCreateWithFlags(&c->evOct[l], cudaEventDisableTiming));
        }
    CKD(cudaStreamCreateWithFlags(&c->streamIn, cudaStreamNonBlocking));
    CKD(cudaStreamCreateWithFlags(&c->streamOut, cudaStreamNonBlocking));
    CKD(cudaEventCreateWithFlags(&c->evOutIdle, cudaEventDisableTiming));
    CKD(cudaEventCreateWithFlags(&c->evReader, cudaEventDisableTiming));
    CKD(cudaEventCreateWithFlags(&c->evPyrReader, cudaEventDisableTiming));
    for (int i = 0; i < eaof_orb::kMaxChunks; ++i) {
        CKD(cudaEventCreateWithFlags(&c->evIn[i], cudaEventDisableTiming));
        CKD(cudaEventCreateWithFlags(&c->evDone[i], cudaEventDisableTiming));
    }
    {
        // chunk size of the host-buffer pipeline: large enough to fill the GPU (the quadtree launches nlevels CTAs per
        // frame, ~3 resident per SM), small enough that the first upload and last download are a small share
        const char* e = getenv("EAOF_CHUNK");
        int cf = e && *e ? atoi(e) : 0;
        if (cf < 1) cf = std::max(8, 444 / std::max(1, p.nlevels));
        cf = std::max(cf, (p.max_batch + eaof_orb::kMaxChunks - 1) / eaof_orb::kMaxChunks);
        c->chunkFrames = std::min(cf, p.max_batch);
    }
    CKD(cudaMalloc(&c->dIn, B * (size_t)p.width * p.height));
    CKD(cudaMalloc(&c->dPyr, B * g.pyrFrameBytes));
    CKD(cudaMalloc(&c->dBlur, B * g.pyrFrameBytes));
    CKD(cudaMalloc(&c->dTabs, sizeof(int) * (tabs.size() + 4)));
    CKD(cudaMalloc(&c->dCells, sizeof(CellDesc) * (cells.size() + 1)));
    {
        // umax as the constructor derives it for HALF_PATCH_SIZE = 15 (src/ORBextractor.cc:454-469; the oracle
        // checks these 16 values against the reference object), then the per-alignment DP4A weights
        static const int umax[EAOF_HALF_PATCH + 1] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        std::vector<uint2> at(4 * EAOF_ANGLE_TASKS_PAD, make_uint2(0, 0));
        for (int a = 0; a < 4; ++a)
            for (int r = 0; r < 31; ++r)
                for (int j = 0; j < 9; ++j) {
                    const int v = r - EAOF_HALF_PATCH;
                    uint32_t uw = 0, vw = 0;
                    for (int b = 0; b < 4; ++b) {
                        const int u = 4 * j - a - EAOF_HALF_PATCH + b;
                        if (abs(u) <= EAOF_HALF_PATCH && abs(u) <= umax[abs(v)]) {
                            uw |= (uint32_t)(uint8_t)(int8_t)u << (8 * b);
                            vw |= (uint32_t)(uint8_t)(int8_t)v << (8 * b);
                        }
                    }
                    at[(size_t)a * EAOF_ANGLE_TASKS_PAD + r * 9 + j] = make_uint2(uw, vw);
                }
        CKD(cudaMalloc(&c->dAngleTab, sizeof(uint2) * at.size()));
        CKD(cudaMemcpy(c->dAngleTab, at.data(), sizeof(uint2) * at.size(), cudaMemcpyHostToDevice));
    }
    CKD(cudaMalloc(&c->dCand, sizeof(uint32_t) * B * (size_t)(g.candPerFrame + 64)));
    CKD(cudaMalloc(&c->dLabel, sizeof(uint16_t) * B * (size_t)(g.candPerFrame + 64)));
    CKD(cudaMalloc(&c->dCandCount, sizeof(uint32_t) * B * g.nlevels));
    CKD(cudaMalloc(&c->dSlotXY, sizeof(uint32_t) * B * g.slotsPerFrame));
    CKD(cudaMalloc(&c->dSlotScore, B * g.slotsPerFrame));
    CKD(cudaMalloc(&c->dLvlCount, sizeof(int) * B * g.nlevels));
    CKD(cudaMalloc(&c->dKps, sizeof(eaof_kp) * B * c->kpCap));
    CKD(cudaMalloc(&c->dDesc, 32 * B * c->kpCap));
    CKD(cudaMalloc(&c->dKpCount, sizeof(int) * B));
    CKD(cudaMemset(c->dPyr, 0, B * g.pyrFrameBytes));
    CKD(cudaMemset(c->dBlur, 0, B * g.pyrFrameBytes));
    CKD(cudaMemset(c->dKpCount, 0, sizeof(int) * B));
    // outputs are downloaded with all their cap slots: what the kernels do not fill stays zero instead of whatever the
    // allocation held before (and compute-sanitizer initcheck stays quiet about the copies)
    CKD(cudaMemset(c->dKps, 0, sizeof(eaof_kp) * B * c->kpCap));
    CKD(cudaMemset(c->dDesc, 0, 32 * B * c->kpCap));
    CKD(cudaMemset(c->dCand, 0, sizeof(uint32_t) * B * (size_t)(g.candPerFrame + 64)));
    CKD(cudaMemset(c->dLabel, 0, sizeof(uint16_t) * B * (size_t)(g.candPerFrame + 64)));
    CKD(cudaMemset(c->dCandCount, 0, sizeof(uint32_t) * B * g.nlevels));
    CKD(cudaMemset(c->dSlotXY, 0, sizeof(uint32_t) * B * g.slotsPerFrame));
    CKD(cudaMemset(c->dSlotScore, 0, B * g.slotsPerFrame));
    CKD(cudaMemset(c->dLvlCount, 0, sizeof(int) * B * g.nlevels));
    CKD(cudaMemcpy(c->dTabs, tabs.data(), sizeof(int) * tabs.size(), cudaMemcpyHostToDevice));
    if (!cells.empty()) CKD(cudaMemcpy(c->dCells, cells.data(), sizeof(CellDesc) * cells.size(), cudaMemcpyHostToDevice));
    CKD(cudaMemcpyToSymbol(eaof::d_pattern, kOrbPattern31, EAOF_ORB_PATTERN_INTS));
    CKD(cudaMallocHost(&c->hIn, B * (size_t)p.width * p.height));
    CKD(cudaMallocHost(&c->hKps, sizeof(eaof_kp) * B * c->kpCap));
    CKD(cudaMallocHost(&c->hDesc, 32 * B * c->kpCap));
    CKD(cudaMallocHost(&c->hKpCount, sizeof(int) * B));
    for (auto& e : c->ev) CKD(cudaEventCreate(&e));
    c->octSmem = (size_t)g.maxNodeCap * 59 + 64;
    c->octSmem = (c->octSmem + 15) & ~(size_t)15;
    {
        // shared-memory budget per CTA width (256 / 512 / 1024 threads): ~7 / 3 / 2 CTAs per SM
        const size_t budget[3] = {30 * 1024, 64 * 1024, 110 * 1024};
        for (int v = 0; v < 3; ++v) c->octKeyCap[v] = budget[v] > c->octSmem ? (int)((budget[v] - c->octSmem) / 6) & ~7 : 0;
        const char* e = getenv("EAOF_OCT_WIDTH");  // 0 / 1 / 2: force a CTA width (A/B runs)
        c->octForce = e && *e ? atoi(e) : -1;
        if (const char* k = getenv("EAOF_OCT_KEYS_GLOBAL")) if (*k == '1') c->octKeyCap[0] = c->octKeyCap[1] = c->octKeyCap[2] = 0;
    }
    if (c->octSmem > 200 * 1024) {
        fail(EAOF_ERR_UNSUPPORTED, "nfeatures too large: quadtree needs %zu B of shared memory", c->octSmem);
        eaof_orb_destroy(c);
        return EAOF_ERR_UNSUPPORTED;
    }
    // k_fast: a few KB of shared memory per warp — ask for the largest carve-out so that shared memory does not cap the
    // resident warps below what the register file allows
    switch (g.fastPW) {  // tile pitch as a template parameter (ring offsets become immediates)
        case 11: c->fastKernel = eaof::k_fast<11>; break;
        case 13: c->fastKernel = eaof::k_fast<13>; break;
        case 15: c->fastKernel = eaof::k_fast<15>; break;
        case 17: c->fastKernel = eaof::k_fast<17>; break;
        case 19: c->fastKernel = eaof::k_fast<19>; break;
        case 21: c->fastKernel = eaof::k_fast<21>; break;
        default: c->fastKernel = eaof::k_fast_generic; c->fastGeneric = true; break;
    }
    cudaFuncSetAttribute(c->fastKernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(eaof::k_fast_generic, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if ((size_t)FAST_WARPS * g.fastWarpWords * 4 > 48 * 1024) {
        static std::mutex muF;
        std::lock_guard<std::mutex> lk(muF);
        CKD(cudaFuncSetAttribute(c->fastKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CKD(cudaFuncSetAttribute(eaof::k_fast_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    {
        // cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            c->encodeTiled = fn;
        else
            cudaGetLastError();
        // k_pyramid_fused: tile plan + shared-memory size
        std::vector<eaof::FusedAxis> axes;
        build_fused_plan(c, tabs, axes);
        if (c->fused) {
            CKD(cudaMalloc(&c->dFusedAxes, sizeof(eaof::FusedAxis) * axes.size()));
            CKD(cudaMemcpy(c->dFusedAxes, axes.data(), sizeof(eaof::FusedAxis) * axes.size(), cudaMemcpyHostToDevice));
            c->fusedA.ax = c->dFusedAxes;
            c->fusedA.ay = c->dFusedAxes + (size_t)g.nlevels * c->fusedA.nTx;
            static std::mutex muP;
            std::lock_guard<std::mutex> lk(muP);
            CKD(cudaFuncSetAttribute(eaof::k_pyramid_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        }
    }
    {
        // k_angle_desc_tma: per level one map over the unblurred pyramid (box 48 x 31: the IC_Angle disc) and one over the
        // blurred pyramid (box 64 x 39: the rotated rBRIEF taps).  EAOF_DESC_TMA=0 keeps the gather kernel (A/B runs).
        const char* e = getenv("EAOF_DESC_TMA");
        const int want = e && *e ? atoi(e) : 1;
        c->descTma = false;
        if (want && c->encodeTiled) {
            typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            bool ok = true;
            for (int which = 0; which < 2 && ok; ++which) {
                std::vector<CUtensorMap> maps(EAOF_MAX_LEVELS);
                memset(maps.data(), 0, sizeof(CUtensorMap) * maps.size());
                for (int l = 0; l < g.nlevels && ok; ++l) {
                    const LevelGeom& L = g.L[l];
                    const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.rows, (cuuint64_t)B};
                    const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.pyrFrameBytes};
                    const cuuint32_t box[3] = {(cuuint32_t)(which ? DESC_BOXB_W : DESC_BOXA_W), (cuuint32_t)(which ? DESC_BOXB_H : DESC_BOXA_H), 1};
                    const cuuint32_t estr[3] = {1, 1, 1};
                    ok = ((EncodeTiled)c->encodeTiled)(&maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (which ? c->dBlur : c->dPyr) + L.off, dims, strides,
                                                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                }
                memcpy(which ? &c->descMapsBlur : &c->descMapsPyr, maps.data(), sizeof(eaof::FastTmaMaps));
            }
            if (ok) {  // slot -> level table of this geometry
                std::vector<uint8_t> sl((size_t)g.slotsPerFrame);
                for (int l = 0; l < g.nlevels; ++l)
                    for (int i = 0; i < g.L[l].nodeCap; ++i) sl[(size_t)g.L[l].slotOff + i] = (uint8_t)l;
                CKD(cudaMalloc(&c->dSlotLevel, sl.size() + 16));
                CKD(cudaMemcpy(c->dSlotLevel, sl.data(), sl.size(), cudaMemcpyHostToDevice));
            }
            c->descTma = ok;
        }
    }
    {
        // k_fast_tma: tensor maps (CU_TENSOR_MAP_DATA_TYPE_UINT8, rank 3: byte column, row, frame) and shared-memory shape.
        // EAOF_FAST_TMA=0 keeps the LDG-staged k_fast (A/B measurements).
        const char* e = getenv("EAOF_FAST_TMA");
        const int want = e && *e ? atoi(e) : EAOF_FAST_TMA_DEFAULT;
        int maxCw = 16, maxCh = 8;
        for (const CellDesc& cd : cells) { maxCw = std::max<int>(maxCw, cd.cw); maxCh = std::max<int>(maxCh, cd.ch); }
        eaof::FastTmaArgs& T = c->fastT;
        T.boxW = (maxCw + 15 + 15) & ~15;  // the box starts on a 16-byte column (TMA rule), up to 15 bytes left of the cell
        T.boxH = (maxCh + 7) & ~7;
        T.tileBytes = (T.boxW * T.boxH + 127) & ~127;
        T.lstCap = 248;  // phase (A) refills the list in rounds of <= 128 entries
        T.warpBytes = (3 * T.tileBytes + 2 * FAST_CLST + 2 * T.lstCap + 16 + 127) & ~127;
        c->fastTmaSmem = (size_t)FASTT_WARPS * T.warpBytes + 128;
        if (want > 0 && !cells.empty() && T.boxW <= 256 && T.boxH <= 256 && (c->fastTmaSmem + 1024) * FASTT_MINB <= 227 * 1024) {
            typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            void* fn = c->encodeTiled;
            if (!fn) {
                fail(EAOF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
                eaof_orb_destroy(c);
                return EAOF_ERR_CUDA;
            }
            std::vector<CUtensorMap> maps(EAOF_MAX_LEVELS);
            memset(maps.data(), 0, sizeof(CUtensorMap) * maps.size());
            for (int l = 0; l < g.nlevels; ++l) {
                const LevelGeom& L = g.L[l];
                const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.rows, (cuuint64_t)B};
                const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.pyrFrameBytes};
                const cuuint32_t box[3] = {(cuuint32_t)T.boxW, (cuuint32_t)T.boxH, 1};
                const cuuint32_t estr[3] = {1, 1, 1};
                const CUresult r = ((EncodeTiled)fn)(&maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->dPyr + L.off, dims, strides, box, estr,
                                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (getenv("EAOF_TMA_DEBUG"))
                    fprintf(stderr, "tmap level %d: base %p dims %llu x %llu x %llu strides %llu %llu box %u x %u -> %d\n", l,
                            (void*)(c->dPyr + L.off), (unsigned long long)dims[0], (unsigned long long)dims[1],
                            (unsigned long long)dims[2], (unsigned long long)strides[0], (unsigned long long)strides[1], box[0], box[1], (int)r);
                if (r != CUDA_SUCCESS) {
                    fail(EAOF_ERR_CUDA, "cuTensorMapEncodeTiled failed for level %d (CUresult %d)", l, (int)r);
                    eaof_orb_destroy(c);
                    return EAOF_ERR_CUDA;
                }
            }
            static_assert(sizeof(CUtensorMap) == 128 && sizeof(eaof::FastTmaMaps) == 128 * EAOF_MAX_LEVELS, "tensor map size");
            memcpy(&c->fastMaps, maps.data(), sizeof(c->fastMaps));
            CKD(cudaMalloc(&c->dFastCtr, sizeof(unsigned int) * eaof_orb::kMaxChunks));
            int sms = 0;
            CKD(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
            c->fastTmaGrid = sms * FASTT_MINB;
            static std::mutex muT;
            std::lock_guard<std::mutex> lk(muT);
            CKD(cudaFuncSetAttribute(eaof::k_fast_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            CKD(cudaFuncSetAttribute(eaof::k_fast_tma, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            CKD(cudaFuncSetAttribute(eaof::k_fast_tma1, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            c->fastTma = want;
#ifdef EAOF_TMA_DEBUG
            {
                int* h = nullptr;
                CKD(cudaHostAlloc(&h, sizeof(int) * 8 * 8192, cudaHostAllocMapped));
                memset(h, 0, sizeof(int) * 8 * 8192);
                int* d = nullptr;
                CKD(cudaHostGetDevicePointer(&d, h, 0));
                T.dbg = d;
                g_tmaDbgHost = h;
            }
#endif
        }
    }
    {
        // the attribute is per function, not per handle: only ever raise it (handles with different nfeatures coexist)
        static std::mutex mu;
        static size_t maxSet[64] = {};
        std::lock_guard<std::mutex> lk(mu);
        size_t need = c->octSmem;
        for (int v = 0; v < 3; ++v) need = std::max(need, c->octSmem + (size_t)c->octKeyCap[v] * 6);
        if (device < 64 && need > maxSet[device]) {
            CKD(cudaFuncSetAttribute(eaof::k_octree<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            CKD(cudaFuncSetAttribute(eaof::k_octree<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            CKD(cudaFuncSetAttribute(eaof::k_octree<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            maxSet[device] = need;
        }
    }
#undef CKD
    *out = c;
    return EAOF_OK;
}

void eaof_orb_destroy(eaof_orb* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->dAux); cudaFree(c->dURight); cudaFree(c->dDepthKp);
    cudaFree(c->dIn); cudaFree(c->dPyr); cudaFree(c->dBlur); cudaFree(c->dTabs); cudaFree(c->dAngleTab); cudaFree(c->dSlotLevel); cudaFree(c->dCells);
    cudaFree(c->dCand); cudaFree(c->dLabel); cudaFree(c->dCandCount); cudaFree(c->dSlotXY); cudaFree(c->dSlotScore);
    cudaFree(c->dLvlCount); cudaFree(c->dKps); cudaFree(c->dDesc); cudaFree(c->dKpCount);
    cudaFreeHost(c->hIn); cudaFreeHost(c->hKps); cudaFreeHost(c->hDesc); cudaFreeHost(c->hKpCount);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->evPyr) cudaEventDestroy(c->evPyr);
    if (c->evBlur) cudaEventDestroy(c->evBlur);
    for (int l = 0; l < EAOF_MAX_LEVELS; ++l) {
        if (c->dagStream[l]) cudaStreamDestroy(c->dagStream[l]);
        if (c->evLvl[l]) cudaEventDestroy(c->evLvl[l]);
        if (c->evOct[l]) cudaEventDestroy(c->evOct[l]);
    }
    if (c->stream2) cudaStreamDestroy(c->stream2);
    for (int i = 0; i < eaof_orb::kMaxChunks; ++i) {
        if (c->evIn[i]) cudaEventDestroy(c->evIn[i]);
        if (c->evDone[i]) cudaEventDestroy(c->evDone[i]);
    }
    if (c->evOutIdle) cudaEventDestroy(c->evOutIdle);
    if (c->evReader) cudaEventDestroy(c->evReader);
    if (c->evPyrReader) cudaEventDestroy(c->evPyrReader);
    cudaFree(c->dSad);
    cudaFree(c->dFastCtr);
    cudaFree(c->dFusedAxes);
    if (c->graphExec1) cudaGraphExecDestroy(c->graphExec1);
    if (c->graph1) cudaGraphDestroy(c->graph1);
    if (c->streamIn) { cudaStreamSynchronize(c->streamIn); cudaStreamDestroy(c->streamIn); }
    if (c->streamOut) { cudaStreamSynchronize(c->streamOut); cudaStreamDestroy(c->streamOut); }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int eaof_orb_max_keypoints(const eaof_orb* c) { return c ? c->kpCap : fail(EAOF_ERR_ARG, "null handle"); }

int eaof_orb_scale_tables(const eaof_orb* c, float* scale, float* inv, float* s2, float* is2, int* fpl) {
    if (!c) return fail(EAOF_ERR_ARG, "null handle");
    for (int i = 0; i < c->p.nlevels; ++i) {
        if (scale) scale[i] = c->scale[i];
        if (inv) inv[i] = c->invScale[i];
        if (s2) s2[i] = c->sigma2[i];
        if (is2) is2[i] = c->invSigma2[i];
        if (fpl) fpl[i] = c->quota[i];
    }
    return EAOF_OK;
}

int eaof_orb_level_size(const eaof_orb* c, int level, int* w, int* h) {
    if (!c || level < 0 || level >= c->p.nlevels) return fail(EAOF_ERR_ARG, "bad level");
    if (w) *w = c->g.L[level].w;
    if (h) *h = c->g.L[level].h;
    return EAOF_OK;
}

int eaof_orb_extract_batch_device(eaof_orb* c, const uint8_t* dImgs, int n, int width, int height, size_t stride,
                                  size_t framePitch) {
    int rc = check_shape(c, width, height, n);
    if (rc) return rc;
    if (!dImgs || stride < (size_t)width) return fail(EAOF_ERR_ARG, "bad image pointer/stride");
    CK(cudaSetDevice(c->device));
    return run_batch(c, dImgs, n, stride, framePitch);
}

namespace {
int set_color(eaof_orb* c, int color, int grayMode) {
    if (color < EAOF_COLOR_BGR || color > EAOF_COLOR_RGBA) return fail(EAOF_ERR_ARG, "unknown colour layout");
    if (grayMode != EAOF_GRAY_CV331 && grayMode != EAOF_GRAY_CV4) return fail(EAOF_ERR_ARG, "unknown gray_mode");
    // OpenCV RGB2Gray<uchar>: B2Y, G2Y, R2Y at yuv_shift 14 (3.3.1) / BY15, GY15, RY15 at 15 bits (4.x)
    const int kB = grayMode == EAOF_GRAY_CV331 ? 1868 : 3735, kG = grayMode == EAOF_GRAY_CV331 ? 9617 : 19235,
              kR = grayMode == EAOF_GRAY_CV331 ? 4899 : 9798;
    const bool rgb = color == EAOF_COLOR_RGB || color == EAOF_COLOR_RGBA;
    c->colorCh = (color == EAOF_COLOR_BGRA || color == EAOF_COLOR_RGBA) ? 4 : 3;
    c->colorK[0] = rgb ? kR : kB; c->colorK[1] = kG; c->colorK[2] = rgb ? kB : kR;
    c->colorShift = grayMode == EAOF_GRAY_CV331 ? 14 : 15;
    return EAOF_OK;
}
int need_aux(eaof_orb* c, size_t bytes) {
    if (c->auxBytes >= bytes) return EAOF_OK;
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(c->dAux);
    c->dAux = nullptr; c->auxBytes = 0;
    CK(cudaMalloc(&c->dAux, bytes));
    c->auxBytes = bytes;
    return EAOF_OK;
}
}  // namespace

int eaof_orb_extract_batch_device_color(eaof_orb* c, const uint8_t* dImgs, int n, int width, int height, size_t stride,
                                        size_t framePitch, int color, int grayMode) {
    int rc = check_shape(c, width, height, n);
    if (rc) return rc;
    if ((rc = set_color(c, color, grayMode))) return rc;
    if (!dImgs || stride < (size_t)width * c->colorCh) { c->colorCh = 0; return fail(EAOF_ERR_ARG, "bad image pointer/stride"); }
    CK(cudaSetDevice(c->device));
    rc = run_batch(c, dImgs, n, stride, framePitch);
    c->colorCh = 0;
    return rc;
}

int eaof_orb_extract_batch_color(eaof_orb* c, const uint8_t* imgs, int n, int width, int height, size_t stride,
                                 size_t framePitch, int color, int grayMode, eaof_kp* kps, uint8_t* desc, int cap, int* nOut) {
    int rc = check_shape(c, width, height, n);
    if (rc) return rc;
    if (!imgs || !nOut) return fail(EAOF_ERR_ARG, "bad argument");
    const int ch = (color == EAOF_COLOR_BGRA || color == EAOF_COLOR_RGBA) ? 4 : 3;
    if (stride < (size_t)width * ch) return fail(EAOF_ERR_ARG, "bad stride");
    CK(cudaSetDevice(c->device));
    const size_t rowBytes = (size_t)width * ch, tight = rowBytes * height;
    if ((rc = need_aux(c, tight * n))) return rc;
    for (int f = 0; f < n; ++f)
        CK(cudaMemcpy2DAsync(c->dAux + (size_t)f * tight, rowBytes, imgs + (size_t)f * framePitch, stride, rowBytes, height,
                             cudaMemcpyHostToDevice, c->stream));
    if ((rc = eaof_orb_extract_batch_device_color(c, c->dAux, n, width, height, rowBytes, tight, color, grayMode))) return rc;
    return eaof_orb_fetch_results(c, n, kps, desc, cap, nOut);
}

int eaof_orb_stereo_from_rgbd_device(eaof_orb* c, int n, const void* dDepth, int depthType, float depthScale,
                                     size_t strideBytes, size_t framePitchBytes, const float* dXUn, float mbf,
                                     float* dURight, float* dDepthOut) {
    if (!c || !dDepth || !dURight || !dDepthOut || n < 1 || n > c->p.max_batch) return fail(EAOF_ERR_ARG, "bad argument");
    if (depthType != EAOF_DEPTH_F32 && depthType != EAOF_DEPTH_U16) return fail(EAOF_ERR_ARG, "unknown depth_type");
    CK(cudaSetDevice(c->device));
    const dim3 gr((c->kpCap + 255) / 256, n);
    if (depthType == EAOF_DEPTH_U16)
        eaof::k_stereo_from_rgbd<true><<<gr, 256, 0, c->stream>>>(c->dKps, c->dKpCount, c->kpCap, dDepth, strideBytes, framePitchBytes,
                                                                  depthScale, c->p.width, c->p.height, dXUn, mbf, dURight, dDepthOut);
    else
        eaof::k_stereo_from_rgbd<false><<<gr, 256, 0, c->stream>>>(c->dKps, c->dKpCount, c->kpCap, dDepth, strideBytes, framePitchBytes,
                                                                   depthScale, c->p.width, c->p.height, dXUn, mbf, dURight, dDepthOut);
    CK(cudaGetLastError());
    return EAOF_OK;
}

int eaof_orb_stereo_from_rgbd(eaof_orb* c, int n, const void* depth, int depthType, float depthScale, size_t strideBytes,
                              size_t framePitchBytes, float mbf, float* uRight, float* depthOut, int cap) {
    if (!c || !depth || !uRight || !depthOut || n < 1 || n > c->p.max_batch) return fail(EAOF_ERR_ARG, "bad argument");
    if (depthType != EAOF_DEPTH_F32 && depthType != EAOF_DEPTH_U16) return fail(EAOF_ERR_ARG, "unknown depth_type");
    const size_t px = depthType == EAOF_DEPTH_U16 ? 2 : 4, rowBytes = px * c->p.width, tight = rowBytes * c->p.height;
    if (strideBytes < rowBytes) return fail(EAOF_ERR_ARG, "bad stride");
    CK(cudaSetDevice(c->device));
    int rc = need_aux(c, tight * n);
    if (rc) return rc;
    const size_t outN = (size_t)c->p.max_batch * c->kpCap;
    if (!c->dURight) { CK(cudaMalloc(&c->dURight, sizeof(float) * outN)); CK(cudaMalloc(&c->dDepthKp, sizeof(float) * outN)); }
    for (int f = 0; f < n; ++f)
        CK(cudaMemcpy2DAsync(c->dAux + (size_t)f * tight, rowBytes, static_cast<const uint8_t*>(depth) + (size_t)f * framePitchBytes,
                             strideBytes, rowBytes, c->p.height, cudaMemcpyHostToDevice, c->stream));
    if ((rc = eaof_orb_stereo_from_rgbd_device(c, n, c->dAux, depthType, depthScale, rowBytes, tight, nullptr, mbf, c->dURight, c->dDepthKp)))
        return rc;
    std::vector<float> hu((size_t)n * c->kpCap), hd((size_t)n * c->kpCap);
    std::vector<int> cnt(n);
    CK(cudaMemcpyAsync(hu.data(), c->dURight, sizeof(float) * hu.size(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(hd.data(), c->dDepthKp, sizeof(float) * hd.size(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(cnt.data(), c->dKpCount, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int f = 0; f < n; ++f) {
        if (cnt[f] > cap) return fail(EAOF_ERR_ARG, "frame %d has %d keypoints but cap is %d", f, cnt[f], cap);
        memcpy(uRight + (size_t)f * cap, hu.data() + (size_t)f * c->kpCap, sizeof(float) * cnt[f]);
        memcpy(depthOut + (size_t)f * cap, hd.data() + (size_t)f * c->kpCap, sizeof(float) * cnt[f]);
    }
    return EAOF_OK;
}

int eaof_stereo_matches_device(eaof_orb* l, eaof_orb* r, int n, float mb, float mbf, float* dURight, float* dDepth) {
    if (!l || !r || !dURight || !dDepth || n < 1) return fail(EAOF_ERR_ARG, "bad argument");
    if (l == r) return fail(EAOF_ERR_ARG, "left and right must be two extractor handles");
    if (l->device != r->device) return fail(EAOF_ERR_ARG, "both handles must live on one device");
    if (l->p.width != r->p.width || l->p.height != r->p.height || l->p.nlevels != r->p.nlevels ||
        l->p.scale_factor != r->p.scale_factor)
        return fail(EAOF_ERR_ARG, "left and right handles must share frame size and pyramid parameters");
    if (n > l->lastFrames || n > r->lastFrames) return fail(EAOF_ERR_ARG, "n_frames exceeds the last batch of a handle");
    if (!(mb > 0)) return fail(EAOF_ERR_ARG, "mb must be positive");
    CK(cudaSetDevice(l->device));
    if (!l->dSad) CK(cudaMalloc(&l->dSad, sizeof(int) * (size_t)l->p.max_batch * l->kpCap));
    cudaStream_t s = l->stream;
    CK(cudaEventRecord(r->evPyr, r->stream));  // the right camera's batch is complete
    CK(cudaStreamWaitEvent(s, r->evPyr, 0));
    eaof::StereoArgs A{};
    A.kpL = l->dKps; A.descL = l->dDesc; A.cntL = l->dKpCount; A.pyrL = l->dPyr;
    A.kpR = r->dKps; A.descR = r->dDesc; A.cntR = r->dKpCount; A.pyrR = r->dPyr;
    A.capL = l->kpCap; A.capR = r->kpCap;
    for (int i = 0; i < l->p.nlevels; ++i) A.invScale[i] = l->invScale[i];
    A.mb = mb; A.mbf = mbf;
    const size_t smemMatch = (size_t)r->kpCap * 9 + 16;
    if (smemMatch > 48 * 1024) CK(cudaFuncSetAttribute(eaof::k_stereo_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemMatch));
    eaof::k_stereo_match<<<dim3((l->kpCap + 15) / 16, n), 512, smemMatch, s>>>(A, dURight, dDepth, l->dSad, l->g);
    eaof::k_stereo_filter<<<n, 1024, sizeof(int) * (size_t)l->kpCap, s>>>(l->dKpCount, l->kpCap, dURight, dDepth, l->dSad);
    CK(cudaGetLastError());
    // the right handle's next batch must not overwrite what these kernels read
    CK(cudaEventRecord(r->evReader, s));
    r->readerPending = true;
    CK(cudaEventRecord(r->evPyrReader, s));
    r->pyrReaderPending = true;
    return EAOF_OK;
}

int eaof_stereo_matches(eaof_orb* l, eaof_orb* r, int n, float mb, float mbf, float* uRight, float* depthOut, int cap) {
    if (!l || !r || !uRight || !depthOut || n < 1 || n > l->p.max_batch) return fail(EAOF_ERR_ARG, "bad argument");
    CK(cudaSetDevice(l->device));
    const size_t outN = (size_t)l->p.max_batch * l->kpCap;
    if (!l->dURight) { CK(cudaMalloc(&l->dURight, sizeof(float) * outN)); CK(cudaMalloc(&l->dDepthKp, sizeof(float) * outN)); }
    int rc = eaof_stereo_matches_device(l, r, n, mb, mbf, l->dURight, l->dDepthKp);
    if (rc) return rc;
    std::vector<float> hu((size_t)n * l->kpCap), hd((size_t)n * l->kpCap);
    std::vector<int> cnt(n);
    CK(cudaMemcpyAsync(hu.data(), l->dURight, sizeof(float) * hu.size(), cudaMemcpyDeviceToHost, l->stream));
    CK(cudaMemcpyAsync(hd.data(), l->dDepthKp, sizeof(float) * hd.size(), cudaMemcpyDeviceToHost, l->stream));
    CK(cudaMemcpyAsync(cnt.data(), l->dKpCount, sizeof(int) * n, cudaMemcpyDeviceToHost, l->stream));
    CK(cudaStreamSynchronize(l->stream));
    for (int f = 0; f < n; ++f) {
        if (cnt[f] > cap) return fail(EAOF_ERR_ARG, "frame %d has %d keypoints but cap is %d", f, cnt[f], cap);
        memcpy(uRight + (size_t)f * cap, hu.data() + (size_t)f * l->kpCap, sizeof(float) * cnt[f]);
        memcpy(depthOut + (size_t)f * cap, hd.data() + (size_t)f * l->kpCap, sizeof(float) * cnt[f]);
    }
    return EAOF_OK;
}

int eaof_orb_undistort_keypoints_device(eaof_orb* c, int n, float fx, float fy, float cx, float cy, const float* dist,
                                        int nDist, int mode, float* dXUn, float* dYUn) {
    if (!c || !dXUn || !dYUn || n < 1 || n > c->p.max_batch) return fail(EAOF_ERR_ARG, "bad argument");
    if (nDist < 0 || nDist > 12 || (nDist && !dist)) return fail(EAOF_ERR_ARG, "0..12 distortion coefficients expected");
    if (mode != EAOF_UNDISTORT_CV331 && mode != EAOF_UNDISTORT_CV4) return fail(EAOF_ERR_ARG, "unknown undistort mode");
    CK(cudaSetDevice(c->device));
    const dim3 gr((c->kpCap + 255) / 256, n);
    if (nDist == 0 || dist[0] == 0.0f) {  // mvKeysUn = mvKeys, src/Frame.cc:775-779
        eaof::k_copy_xy<<<gr, 256, 0, c->stream>>>(c->dKps, c->dKpCount, c->kpCap, dXUn, dYUn);
    } else {
        eaof::UndistortArgs U{};
        U.fx = fx; U.fy = fy; U.cx = cx; U.cy = cy;
        for (int i = 0; i < nDist; ++i) U.k[i] = dist[i];
        U.guard = mode == EAOF_UNDISTORT_CV4;
        eaof::k_undistort<<<gr, 256, 0, c->stream>>>(c->dKps, c->dKpCount, c->kpCap, U, dXUn, dYUn);
    }
    CK(cudaGetLastError());
    return EAOF_OK;
}

int eaof_orb_undistort_keypoints(eaof_orb* c, int n, float fx, float fy, float cx, float cy, const float* dist, int nDist,
                                 int mode, float* xUn, float* yUn, int cap) {
    if (!c || !xUn || !yUn || n < 1 || n > c->p.max_batch) return fail(EAOF_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    const size_t outN = (size_t)c->p.max_batch * c->kpCap;
    if (!c->dURight) { CK(cudaMalloc(&c->dURight, sizeof(float) * outN)); CK(cudaMalloc(&c->dDepthKp, sizeof(float) * outN)); }
    int rc = eaof_orb_undistort_keypoints_device(c, n, fx, fy, cx, cy, dist, nDist, mode, c->dURight, c->dDepthKp);
    if (rc) return rc;
    std::vector<float> hx((size_t)n * c->kpCap), hy((size_t)n * c->kpCap);
    std::vector<int> cnt(n);
    CK(cudaMemcpyAsync(hx.data(), c->dURight, sizeof(float) * hx.size(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(hy.data(), c->dDepthKp, sizeof(float) * hy.size(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(cnt.data(), c->dKpCount, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int f = 0; f < n; ++f) {
        if (cnt[f] > cap) return fail(EAOF_ERR_ARG, "frame %d has %d keypoints but cap is %d", f, cnt[f], cap);
        memcpy(xUn + (size_t)f * cap, hx.data() + (size_t)f * c->kpCap, sizeof(float) * cnt[f]);
        memcpy(yUn + (size_t)f * cap, hy.data() + (size_t)f * c->kpCap, sizeof(float) * cnt[f]);
    }
    return EAOF_OK;
}

int eaof_orb_sync(eaof_orb* c) {
    if (!c) return fail(EAOF_ERR_ARG, "null handle");
    CK(cudaStreamSynchronize(c->stream));
    if (c->profiling && c->lastFrames > 0) {
        float tot = 0;
        for (int i = 0; i < 5; ++i) {
            CK(cudaEventElapsedTime(&c->stageMs[i], c->ev[i], c->ev[i + 1]));
            tot += c->stageMs[i];
        }
        c->stageMs[5] = tot;
    }
    return EAOF_OK;
}

int eaof_orb_device_results(eaof_orb* c, const eaof_kp** k, const uint8_t** d, const int** n, int* cap) {
    if (!c) return fail(EAOF_ERR_ARG, "null handle");
    if (k) *k = c->dKps;
    if (d) *d = c->dDesc;
    if (n) *n = c->dKpCount;
    if (cap) *cap = c->kpCap;
    return EAOF_OK;
}

int eaof_orb_fetch_results(eaof_orb* c, int n, eaof_kp* kps, uint8_t* desc, int cap, int* nOut) {
    if (!c || !nOut || n < 1 || n > c->p.max_batch) return fail(EAOF_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    CK(cudaMemcpyAsync(c->hKpCount, c->dKpCount, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    if (kps) CK(cudaMemcpyAsync(c->hKps, c->dKps, sizeof(eaof_kp) * (size_t)n * c->kpCap, cudaMemcpyDeviceToHost, s));
    if (desc) CK(cudaMemcpyAsync(c->hDesc, c->dDesc, 32 * (size_t)n * c->kpCap, cudaMemcpyDeviceToHost, s));
    int rc = eaof_orb_sync(c);
    if (rc) return rc;
    for (int f = 0; f < n; ++f) {
        const int k = c->hKpCount[f];
        nOut[f] = k;
        if (k > cap && (kps || desc)) return fail(EAOF_ERR_ARG, "frame %d has %d keypoints but cap is %d", f, k, cap);
        if (kps) memcpy(kps + (size_t)f * cap, c->hKps + (size_t)f * c->kpCap, sizeof(eaof_kp) * k);
        if (desc) memcpy(desc + (size_t)f * cap * 32, c->hDesc + (size_t)f * c->kpCap * 32, 32 * (size_t)k);
    }
    return EAOF_OK;
}

static inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int eaof_orb_extract_batch_async(eaof_orb* c, const uint8_t* imgs, int n, int width, int height, size_t stride,
                                 size_t framePitch, eaof_kp* kps, uint8_t* desc, int cap) {
    int rc = check_shape(c, width, height, n);
    if (rc) return rc;
    if (!imgs || stride < (size_t)width) return fail(EAOF_ERR_ARG, "bad argument");
    if (c->pendN) return fail(EAOF_ERR_ARG, "a batch is already in flight on this handle: call eaof_orb_extract_batch_wait first");
    CK(cudaSetDevice(c->device));
    const size_t frameBytes = (size_t)width * height;
    const bool packed = stride == (size_t)width && framePitch == frameBytes;
    // When the caller's output layout is the device layout (cap == eaof_orb_max_keypoints) results are downloaded
    // straight into the caller's buffers; otherwise through the pinned staging buffers and compacted on the host.
    const bool direct = cap == c->kpCap;
    eaof_kp* hK = kps ? (direct ? kps : c->hKps) : nullptr;
    uint8_t* hD = desc ? (direct ? desc : c->hDesc) : nullptr;
    static const bool noGraph = getenv("EAOF_NO_GRAPH") != nullptr;  // experiment knob
    if (n == 1 && !c->profiling && !noGraph && !(c->graphTried && !c->graphExec1)) {
        // ---- latency path: upload on the compute stream, then the captured graph (kernels + downloads into the pinned staging)
        cudaStream_t s = c->stream;
        c->latT[0] = now_s();
        if (c->pyrReaderPending) { CK(cudaStreamWaitEvent(s, c->evPyrReader, 0)); c->pyrReaderPending = false; }
        if (c->readerPending) { CK(cudaStreamWaitEvent(s, c->evReader, 0)); c->readerPending = false; }
        if (packed)
            CK(cudaMemcpyAsync(c->dIn, imgs, frameBytes, cudaMemcpyHostToDevice, s));
        else
            CK(cudaMemcpy2DAsync(c->dIn, (size_t)width, imgs, stride, (size_t)width, (size_t)height, cudaMemcpyHostToDevice, s));
        c->latT[1] = now_s();
        if (!c->graphExec1) {
            c->graphTried = true;
            c->capturing = true;
            cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
            int rcb = EAOF_OK;
            if (e == cudaSuccess) {
                rcb = run_batch(c, c->dIn, 1, (size_t)width, frameBytes, 0, 0);
                if (rcb == EAOF_OK) {
                    cudaMemcpyAsync(c->hKpCount, c->dKpCount, sizeof(int), cudaMemcpyDeviceToHost, s);
                    cudaMemcpyAsync(c->hKps, c->dKps, sizeof(eaof_kp) * (size_t)c->kpCap, cudaMemcpyDeviceToHost, s);
                    cudaMemcpyAsync(c->hDesc, c->dDesc, 32 * (size_t)c->kpCap, cudaMemcpyDeviceToHost, s);
                }
                e = cudaStreamEndCapture(s, &c->graph1);
                if (e == cudaSuccess && rcb == EAOF_OK) e = cudaGraphInstantiate(&c->graphExec1, c->graph1, 0);
            }
            c->capturing = false;
            if (e != cudaSuccess || rcb != EAOF_OK || !c->graphExec1) {
                // not capturable here: this handle stays on the stream path below (graphTried keeps it there)
                cudaGetLastError();
                if (c->graph1) { cudaGraphDestroy(c->graph1); c->graph1 = nullptr; }
                c->graphExec1 = nullptr;
            }
        }
        if (c->graphExec1) {
            CK(cudaGraphLaunch(c->graphExec1, s));
            c->latT[2] = now_s();
            c->lastFrames = 1;
            c->pendN = 1; c->pendCap = cap; c->pendKps = kps; c->pendDesc = desc; c->pendDirect = false; c->pendGraph = true;
            return EAOF_OK;
        }
    }
    const bool pipelined = !c->profiling && n > c->chunkFrames;
    const int chunk = pipelined ? c->chunkFrames : n;
    // the download stream may still be busy with the previous call's copies out of the same device buffers
    CK(cudaEventRecord(c->evOutIdle, c->streamOut));
    CK(cudaStreamWaitEvent(c->stream, c->evOutIdle, 0));
    const int dev = c->device < 64 ? c->device : 63;
    std::lock_guard<std::mutex> tokLock(g_tokMu);
    if (g_tokSet[dev]) CK(cudaStreamWaitEvent(c->stream, g_tok[dev], 0));
    int ci = 0;
    for (int f0 = 0; f0 < n; f0 += chunk, ++ci) {
        const int m = std::min(chunk, n - f0);
        // upload (H2D straight from the caller's buffer: truly asynchronous only when it is pinned)
        uint8_t* dst = c->dIn + (size_t)f0 * frameBytes;
        if (packed) {
            CK(cudaMemcpyAsync(dst, imgs + (size_t)f0 * framePitch, (size_t)m * frameBytes, cudaMemcpyHostToDevice, c->streamIn));
        } else {
            for (int f = 0; f < m; ++f)
                CK(cudaMemcpy2DAsync(dst + (size_t)f * frameBytes, (size_t)width, imgs + (size_t)(f0 + f) * framePitch, stride,
                                     (size_t)width, (size_t)height, cudaMemcpyHostToDevice, c->streamIn));
        }
        CK(cudaEventRecord(c->evIn[ci], c->streamIn));
        CK(cudaStreamWaitEvent(c->stream, c->evIn[ci], 0));
        rc = run_batch(c, dst, m, (size_t)width, frameBytes, f0, ci);
        if (rc) return rc;
        CK(cudaEventRecord(c->evDone[ci], c->stream));
        // download
        CK(cudaStreamWaitEvent(c->streamOut, c->evDone[ci], 0));
        CK(cudaMemcpyAsync(c->hKpCount + f0, c->dKpCount + f0, sizeof(int) * m, cudaMemcpyDeviceToHost, c->streamOut));
        if (hK) CK(cudaMemcpyAsync(hK + (size_t)f0 * c->kpCap, c->dKps + (size_t)f0 * c->kpCap, sizeof(eaof_kp) * (size_t)m * c->kpCap,
                                   cudaMemcpyDeviceToHost, c->streamOut));
        if (hD) CK(cudaMemcpyAsync(hD + (size_t)f0 * c->kpCap * 32, c->dDesc + (size_t)f0 * c->kpCap * 32, 32 * (size_t)m * c->kpCap,
                                   cudaMemcpyDeviceToHost, c->streamOut));
    }
    if (!g_tok[dev]) CK(cudaEventCreateWithFlags(&g_tok[dev], cudaEventDisableTiming));
    CK(cudaEventRecord(g_tok[dev], c->stream));
    g_tokSet[dev] = true;
    c->pendN = n; c->pendCap = cap; c->pendKps = kps; c->pendDesc = desc; c->pendDirect = direct; c->pendGraph = false;
    return EAOF_OK;
}

int eaof_orb_extract_batch_wait(eaof_orb* c, int* nOut) {
    if (!c || !nOut) return fail(EAOF_ERR_ARG, "null argument");
    if (!c->pendN) return fail(EAOF_ERR_ARG, "no batch in flight on this handle");
    const int n = c->pendN, cap = c->pendCap;
    eaof_kp* kps = c->pendKps;
    uint8_t* desc = c->pendDesc;
    const bool direct = c->pendDirect;
    c->pendN = 0;
    c->latT[3] = now_s();
    CK(cudaSetDevice(c->device));
    if (!c->pendGraph) CK(cudaStreamSynchronize(c->streamOut));  // the graph downloads on the compute stream itself
    int rc = eaof_orb_sync(c);
    if (rc) return rc;
    c->latT[4] = now_s();
    for (int f = 0; f < n; ++f) {
        const int k = c->hKpCount[f];
        nOut[f] = k;
        if (k > cap && (kps || desc)) return fail(EAOF_ERR_ARG, "frame %d has %d keypoints but cap is %d", f, k, cap);
        if (!direct) {
            if (kps) memcpy(kps + (size_t)f * cap, c->hKps + (size_t)f * c->kpCap, sizeof(eaof_kp) * k);
            if (desc) memcpy(desc + (size_t)f * cap * 32, c->hDesc + (size_t)f * c->kpCap * 32, 32 * (size_t)k);
        }
    }
    c->latT[5] = now_s();
    return EAOF_OK;
}

int eaof_debug_latency_trace(eaof_orb* c, double* out6) {
    if (!c || !out6) return fail(EAOF_ERR_ARG, "null argument");
    for (int i = 0; i < 6; ++i) out6[i] = c->latT[i];
    return EAOF_OK;
}

int eaof_orb_extract_batch(eaof_orb* c, const uint8_t* imgs, int n, int width, int height, size_t stride,
                           size_t framePitch, eaof_kp* kps, uint8_t* desc, int cap, int* nOut) {
    if (!nOut) return fail(EAOF_ERR_ARG, "bad argument");
    int rc = eaof_orb_extract_batch_async(c, imgs, n, width, height, stride, framePitch, kps, desc, cap);
    if (rc) return rc;
    return eaof_orb_extract_batch_wait(c, nOut);
}

int eaof_orb_set_pipeline_chunk(eaof_orb* c, int frames) {
    if (!c || frames < 0) return fail(EAOF_ERR_ARG, "bad argument");
    if (frames == 0) frames = std::max(8, 444 / std::max(1, c->p.nlevels));
    frames = std::max(frames, (c->p.max_batch + eaof_orb::kMaxChunks - 1) / eaof_orb::kMaxChunks);
    c->chunkFrames = std::min(frames, c->p.max_batch);
    return EAOF_OK;
}

int eaof_orb_extract(eaof_orb* c, const uint8_t* img, int width, int height, size_t stride, eaof_kp* kps, uint8_t* desc,
                     int cap, int* nOut) {
    if (!img || width <= 0 || height <= 0) return fail(EAOF_ERR_EMPTY, "empty image");
    return eaof_orb_extract_batch(c, img, 1, width, height, stride, stride * (size_t)height, kps, desc, cap, nOut);
}

int eaof_orb_pyramid_level(eaof_orb* c, int frame, int level, int withBorder, uint8_t* dst, size_t dstStride) {
    if (!c || !dst || level < 0 || level >= c->p.nlevels || frame < 0 || frame >= c->p.max_batch)
        return fail(EAOF_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    const LevelGeom& L = c->g.L[level];
    const uint8_t* base = c->dPyr + (size_t)frame * c->g.pyrFrameBytes + L.off;
    CK(cudaStreamSynchronize(c->stream));
    if (withBorder)
        CK(cudaMemcpy2D(dst, dstStride, base + EAOF_INNER_X0 - EAOF_EDGE, L.pitch, L.w + 2 * EAOF_EDGE, L.rows, cudaMemcpyDeviceToHost));
    else
        CK(cudaMemcpy2D(dst, dstStride, base + (size_t)EAOF_EDGE * L.pitch + EAOF_INNER_X0, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return EAOF_OK;
}

int eaof_orb_debug_blurred_level(eaof_orb* c, int frame, int level, uint8_t* dst, size_t dstStride) {
    if (!c || !dst || level < 0 || level >= c->p.nlevels || frame < 0 || frame >= c->p.max_batch)
        return fail(EAOF_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    const LevelGeom& L = c->g.L[level];
    const uint8_t* base = c->dBlur + (size_t)frame * c->g.pyrFrameBytes + L.off;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy2D(dst, dstStride, base + (size_t)EAOF_EDGE * L.pitch + EAOF_INNER_X0, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return EAOF_OK;
}

int eaof_orb_debug_candidates(eaof_orb* c, int frame, int level, int* xys, int cap, int* nOut) {
    if (!c || !nOut || level < 0 || level >= c->p.nlevels || frame < 0 || frame >= c->p.max_batch)
        return fail(EAOF_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    uint32_t n = 0;
    CK(cudaMemcpy(&n, c->dCandCount + (size_t)frame * c->g.nlevels + level, sizeof n, cudaMemcpyDeviceToHost));
    *nOut = (int)n;
    const int m = (int)n < cap ? (int)n : cap;
    if (m > 0 && xys) {
        std::vector<uint32_t> tmp(m);
        CK(cudaMemcpy(tmp.data(), c->dCand + (size_t)frame * c->g.candPerFrame + c->g.L[level].candOff, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
        for (int i = 0; i < m; ++i) {
            xys[3 * i] = tmp[i] & 0xfff;
            xys[3 * i + 1] = (tmp[i] >> 12) & 0xfff;
            xys[3 * i + 2] = tmp[i] >> 24;
        }
    }
    return EAOF_OK;
}

// Device sinf/cosf restatement vs this host's libm over every `stride`-th float in [0, hi]; returns mismatches.
long eaof_debug_sincosf_mismatches(float hi, uint32_t stride) {
    uint32_t ul;
    memcpy(&ul, &hi, 4);
    if (stride == 0) stride = 1;
    const uint32_t chunk = 1u << 24;
    float *ds = nullptr, *dc = nullptr;
    if (cudaMalloc(&ds, sizeof(float) * chunk) != cudaSuccess || cudaMalloc(&dc, sizeof(float) * chunk) != cudaSuccess) {
        fail(EAOF_ERR_CUDA, "cudaMalloc failed");
        return -1;
    }
    std::vector<float> hs(chunk), hc(chunk);
    long bad = 0;
    const uint64_t total = (uint64_t)ul / stride + 1;
    for (uint64_t base = 0; base < total; base += chunk) {
        const uint32_t n = (uint32_t)((total - base) < chunk ? (total - base) : chunk);
        eaof::k_debug_sincosf<<<(n + 255) / 256, 256>>>((uint32_t)(base * stride), stride, n, ds, dc);
        if (cudaMemcpy(hs.data(), ds, sizeof(float) * n, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(hc.data(), dc, sizeof(float) * n, cudaMemcpyDeviceToHost) != cudaSuccess) {
            fail(EAOF_ERR_CUDA, "sincosf sweep failed: %s", cudaGetErrorString(cudaGetLastError()));
            bad = -1;
            break;
        }
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t u = (uint32_t)(base * stride) + i * stride;
            float f;
            memcpy(&f, &u, 4);
            const float rs = sinf(f), rc = cosf(f);
            bad += memcmp(&rs, &hs[i], 4) != 0;
            bad += memcmp(&rc, &hc[i], 4) != 0;
        }
    }
    cudaFree(ds);
    cudaFree(dc);
    return bad;
}

int eaof_orb_set_profiling(eaof_orb* c, int on) {
    if (!c) return fail(EAOF_ERR_ARG, "null handle");
    c->profiling = on != 0;
    return EAOF_OK;
}
int eaof_orb_stage_times(eaof_orb* c, float* ms6) {
    if (!c || !ms6) return fail(EAOF_ERR_ARG, "null argument");
    for (int i = 0; i < 6; ++i) ms6[i] = c->stageMs[i];
    return EAOF_OK;
}
// ---- pinned frame ring (include/eaof_orb.h; ros_test/src/message_flow.cc:250-254 -> src/Tracking.cc:324-337) ----------
// head counts the frames the producer committed, tail the frames the consumer released; slot of frame i = i % slots.  The
// producer only writes head (and the slot it acquired), the consumer only writes tail: two atomics, no lock.
int eaof_ring_create(int slots, int width, int height, int channels, eaof_ring** out) {
    if (!out) return fail(EAOF_ERR_ARG, "null argument");
    *out = nullptr;
    if (slots < 1 || slots > 65536 || width <= 0 || height <= 0 || (channels != 1 && channels != 3 && channels != 4))
        return fail(EAOF_ERR_ARG, "bad ring shape: %d slots of %dx%dx%d", slots, width, height, channels);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(EAOF_ERR_CUDA, "no CUDA device: libeaof_orb has no CPU fallback");
    }
    eaof_ring* r = new eaof_ring();
    r->slots = slots; r->width = width; r->height = height; r->channels = channels;
    r->stride = (size_t)width * channels;
    r->slotBytes = (r->stride * height + 4095) & ~(size_t)4095;
    cudaError_t e = cudaHostAlloc((void**)&r->base, r->slotBytes * slots, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete r;
        return fail(EAOF_ERR_CUDA, "cudaHostAlloc of %zu ring bytes: %s", (size_t)slots * width * height * channels, cudaGetErrorString(e));
    }
    r->stamps.assign(slots, 0.0);
    *out = r;
    return EAOF_OK;
}

void eaof_ring_destroy(eaof_ring* r) {
    if (!r) return;
    if (r->base) cudaFreeHost(r->base);
    delete r;
}

int eaof_ring_acquire(eaof_ring* r, uint8_t** slot, size_t* stride) {
    if (!r || !slot) return fail(EAOF_ERR_ARG, "null argument");
    const uint64_t h = r->head.load(std::memory_order_relaxed), t = r->tail.load(std::memory_order_acquire);
    if (h - t >= (uint64_t)r->slots) return fail(EAOF_ERR_BUSY, "frame ring full: %d frames pending", r->slots);
    *slot = r->base + (size_t)(h % r->slots) * r->slotBytes;
    if (stride) *stride = r->stride;
    return EAOF_OK;
}

int eaof_ring_commit(eaof_ring* r, double timestamp) {
    if (!r) return fail(EAOF_ERR_ARG, "null argument");
    const uint64_t h = r->head.load(std::memory_order_relaxed);
    if (h - r->tail.load(std::memory_order_acquire) >= (uint64_t)r->slots) return fail(EAOF_ERR_BUSY, "commit without a free slot");
    r->stamps[h % r->slots] = timestamp;
    r->head.store(h + 1, std::memory_order_release);  // publishes the pixels and the time stamp
    return EAOF_OK;
}

int eaof_ring_pending(const eaof_ring* r) {
    if (!r) return fail(EAOF_ERR_ARG, "null argument");
    return (int)(r->head.load(std::memory_order_acquire) - r->tail.load(std::memory_order_relaxed));
}

int eaof_ring_peek(const eaof_ring* r, int k, const uint8_t** slot, double* timestamp) {
    if (!r) return fail(EAOF_ERR_ARG, "null argument");
    const uint64_t h = r->head.load(std::memory_order_acquire), t = r->tail.load(std::memory_order_relaxed);
    if (k < 0 || (uint64_t)k >= h - t) return fail(EAOF_ERR_ARG, "frame %d of %d pending", k, (int)(h - t));
    const size_t i = (size_t)((t + k) % r->slots);
    if (slot) *slot = r->base + i * r->slotBytes;
    if (timestamp) *timestamp = r->stamps[i];
    return EAOF_OK;
}

int eaof_ring_release(eaof_ring* r, int n) {
    if (!r) return fail(EAOF_ERR_ARG, "null argument");
    const uint64_t h = r->head.load(std::memory_order_acquire), t = r->tail.load(std::memory_order_relaxed);
    if (n < 0 || (uint64_t)n > h - t) return fail(EAOF_ERR_ARG, "release of %d frames, %d pending", n, (int)(h - t));
    r->tail.store(t + n, std::memory_order_release);
    return EAOF_OK;
}

int eaof_orb_extract_ring(eaof_orb* c, eaof_ring* r, int maxFrames, int color, int grayMode, eaof_kp* kps, uint8_t* desc,
                          int cap, int* nOut, double* timestamps, int* nFrames) {
    if (!c || !r || !nOut || !nFrames) return fail(EAOF_ERR_ARG, "null argument");
    *nFrames = 0;
    if (r->width != c->p.width || r->height != c->p.height)
        return fail(EAOF_ERR_ARG, "ring of %dx%d frames, extractor built for %dx%d", r->width, r->height, c->p.width, c->p.height);
    const uint64_t h = r->head.load(std::memory_order_acquire), t = r->tail.load(std::memory_order_relaxed);
    const size_t first = (size_t)(t % r->slots);
    const int n = (int)std::min<uint64_t>({h - t, (uint64_t)std::max(maxFrames, 0), (uint64_t)c->p.max_batch, (uint64_t)(r->slots - first)});
    if (n == 0) return EAOF_OK;
    if (r->channels != 1 && (r->channels == 4) != (color == EAOF_COLOR_BGRA || color == EAOF_COLOR_RGBA))
        return fail(EAOF_ERR_ARG, "colour order %d does not fit a ring of %d-channel frames", color, r->channels);
    const uint8_t* imgs = r->base + first * r->slotBytes;
    const int rc = r->channels == 1
        ? eaof_orb_extract_batch(c, imgs, n, r->width, r->height, r->stride, r->slotBytes, kps, desc, cap, nOut)
        : eaof_orb_extract_batch_color(c, imgs, n, r->width, r->height, r->stride, r->slotBytes, color, grayMode, kps, desc, cap, nOut);
    if (rc) return rc;
    if (timestamps)
        for (int i = 0; i < n; ++i) timestamps[i] = r->stamps[first + i];
    *nFrames = n;
    return EAOF_OK;
}

// ---- internal hooks for eaof_match.cu (not part of the public ABI)
int eaof_internal_fail(int code, const char* msg) { return fail(code, "%s", msg); }
int eaof_internal_orb_view(eaof_orb* c, const eaof_kp** kps, const uint8_t** desc, const int** counts, int* cap, int* w,
                           int* h, const float** scale, int* nlevels, void** stream) {
    if (!c) return fail(EAOF_ERR_ARG, "null extractor handle");
    *kps = c->dKps; *desc = c->dDesc; *counts = c->dKpCount; *cap = c->kpCap; *w = c->p.width; *h = c->p.height;
    *scale = c->scale.data(); *nlevels = c->p.nlevels; *stream = (void*)c->stream;
    return EAOF_OK;
}

// called by a consumer right after it enqueued its last read of the result buffers on `readerStream`
int eaof_internal_orb_note_reader(eaof_orb* c, void* readerStream) {
    if (!c) return fail(EAOF_ERR_ARG, "null extractor handle");
    CK(cudaEventRecord(c->evReader, (cudaStream_t)readerStream));
    c->readerPending = true;
    return EAOF_OK;
}

int eaof_orb_last_launch_count(const eaof_orb* c) { return c ? c->lastLaunches : 0; }
void* eaof_orb_stream(eaof_orb* c) { return c ? (void*)c->stream : nullptr; }

}  // extern "C"
