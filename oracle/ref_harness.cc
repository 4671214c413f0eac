// ORACLE — test infrastructure only (never linked into the product library).
//
// C harness around the UNMODIFIED reference translation unit
// /root/reference/src/ORBextractor.cc, which oracle/Makefile compiles in place
// against oracle/cvshim (no reference source is copied into this repo).  Built
// into oracle/_ref/liborb_ref.so and used by tests/ as the end-to-end oracle and
// by bench.py as the "reference" CPU baseline.
//
// Canonical tie-break: DistributeOctTree sorts
// pair<int, ExtractorNode*> (src/ORBextractor.cc:591,681-685), so equal-size nodes
// are ordered by heap address.  With `canonical` on, every allocation made while
// the reference runs comes from a monotonic never-reuse arena, so address order ==
// creation order, which is the rule the CUDA path and oracle/orb_oracle.cc implement.
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <thread>
#include <vector>

#include "ORBextractor.h"  // the reference's own header, from /root/reference/include

// ---------------------------------------------------------------- arena operator new
namespace {
struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0, peak = 0;
    bool active = false;
    void ensure() {
        if (base) return;
        cap = (size_t)1 << 34;  // 16 GiB of address space, touched lazily
        void* p = mmap(nullptr, cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { fprintf(stderr, "orb_ref: arena mmap failed\n"); abort(); }
        base = (char*)p;
    }
    void reset() {
        if (off > peak) peak = off;
        // give pages back now and then so long runs do not pin memory
        if (off > ((size_t)256 << 20)) madvise(base, off, MADV_DONTNEED);
        off = 0;
    }
    bool owns(const void* p) const { return base && (const char*)p >= base && (const char*)p < base + cap; }
    void* alloc(size_t n) {
        const size_t a = (off + 15) & ~(size_t)15;
        if (a + n > cap) { fprintf(stderr, "orb_ref: arena exhausted\n"); abort(); }
        off = a + n;
        return base + a;
    }
};
thread_local Arena g_arena;
std::atomic<long> g_arena_allocs{0};
}  // namespace

void* operator new(size_t n) {
    if (g_arena.active) { g_arena_allocs.fetch_add(1, std::memory_order_relaxed); return g_arena.alloc(n ? n : 1); }
    void* p = malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { if (p && !g_arena.owns(p)) free(p); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

// ---------------------------------------------------------------- shim globals
namespace cv {
static int g_blur_mode = cvprim::BLUR_CV331;
int eaof_shim_blur_mode() { return g_blur_mode; }
void eaof_shim_set_blur_mode(int m) { g_blur_mode = m; }
}  // namespace cv

namespace {
struct Probe : public ORB_SLAM2::ORBextractor {  // exposes protected tables read-only
    using ORB_SLAM2::ORBextractor::ORBextractor;
    const std::vector<int>& quotas() const { return mnFeaturesPerLevel; }
    const std::vector<int>& umax_table() const { return umax; }
    const std::vector<cv::Point>& pattern_table() const { return pattern; }
};
struct Handle {
    Probe* ex;
    int nlevels;
    bool canonical = true;
    std::vector<std::vector<uint8_t>> pyr;  // bordered level copies of the last frame
    std::vector<int> lw, lh;
};
}  // namespace

extern "C" {

struct orbref_kp { float x, y, size, angle, response; int octave; };

void orbref_set_blur_mode(int m) { cv::eaof_shim_set_blur_mode(m); }
long orbref_arena_allocs() { return g_arena_allocs.load(); }

void* orbref_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    Handle* h = new Handle;
    h->ex = new Probe(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
    h->nlevels = nlevels;
    h->pyr.resize(nlevels);
    h->lw.assign(nlevels, 0);
    h->lh.assign(nlevels, 0);
    return h;
}
void orbref_destroy(void* hp) {
    Handle* h = (Handle*)hp;
    if (!h) return;
    delete h->ex;
    delete h;
}
void orbref_set_canonical(void* hp, int on) { ((Handle*)hp)->canonical = on != 0; }

void orbref_tables(void* hp, float* sf, float* isf, float* s2, float* is2, int* quotas, int* umax16) {
    Handle* h = (Handle*)hp;
    std::vector<float> a = h->ex->GetScaleFactors(), b = h->ex->GetInverseScaleFactors(),
                       c = h->ex->GetScaleSigmaSquares(), d = h->ex->GetInverseScaleSigmaSquares();
    for (int i = 0; i < h->nlevels; ++i) {
        if (sf) sf[i] = a[i];
        if (isf) isf[i] = b[i];
        if (s2) s2[i] = c[i];
        if (is2) is2[i] = d[i];
        if (quotas) quotas[i] = h->ex->quotas()[i];
    }
    if (umax16) for (int i = 0; i < 16; ++i) umax16[i] = h->ex->umax_table()[i];
}
void orbref_pattern(void* hp, int* xy1024) {
    Handle* h = (Handle*)hp;
    for (int i = 0; i < 512; ++i) { xy1024[2 * i] = h->ex->pattern_table()[i].x; xy1024[2 * i + 1] = h->ex->pattern_table()[i].y; }
}

// Runs ORBextractor::operator() (src/ORBextractor.cc:1043).  Returns the keypoint count (may exceed cap;
// only min(n,cap) entries are written), or -1 for an empty image (the reference returns silently).
int orbref_extract(void* hp, const uint8_t* img, int w, int hgt, size_t stride, orbref_kp* kps, uint8_t* desc,
                   int cap, int keep_pyramid) {
    Handle* h = (Handle*)hp;
    if (!img || w <= 0 || hgt <= 0) return -1;
    Arena& A = g_arena;
    if (h->canonical) { A.ensure(); A.reset(); A.active = true; }
    int n = 0;
    {
        std::vector<cv::KeyPoint> k;
        cv::Mat d;
        cv::Mat image(hgt, w, CV_8UC1, (void*)img, stride);
        (*h->ex)(image, cv::Mat(), k, d);
        A.active = false;
        n = (int)k.size();
        const int m = n < cap ? n : cap;
        for (int i = 0; i < m; ++i) {
            if (kps) kps[i] = {k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave};
            if (desc) memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
        }
        if (keep_pyramid) {
            for (int l = 0; l < h->nlevels; ++l) {
                const cv::Mat& m0 = h->ex->mvImagePyramid[l];
                h->lw[l] = m0.cols;
                h->lh[l] = m0.rows;
                const int bw = m0.cols + 38, bh = m0.rows + 38;
                h->pyr[l].resize((size_t)bw * bh);
                const uint8_t* base = m0.data - 19 * (size_t)m0.step - 19;
                for (int y = 0; y < bh; ++y) memcpy(&h->pyr[l][(size_t)y * bw], base + (size_t)y * m0.step, bw);
            }
        }
        // drop every arena-backed object the extractor still references before the arena is recycled
        h->ex->mvImagePyramid.assign(h->nlevels, cv::Mat());
    }
    if (h->canonical) A.reset();
    return n;
}

int orbref_level_size(void* hp, int level, int* w, int* hgt) {
    Handle* h = (Handle*)hp;
    if (level < 0 || level >= h->nlevels) return -1;
    *w = h->lw[level];
    *hgt = h->lh[level];
    return 0;
}
// with_border: copies the (w+38)x(h+38) buffer, else the inner w x h ROI
int orbref_copy_level(void* hp, int level, uint8_t* dst, size_t dstride, int with_border) {
    Handle* h = (Handle*)hp;
    if (level < 0 || level >= h->nlevels || h->pyr[level].empty()) return -1;
    const int bw = h->lw[level] + 38;
    if (with_border) {
        for (int y = 0; y < h->lh[level] + 38; ++y) memcpy(dst + (size_t)y * dstride, &h->pyr[level][(size_t)y * bw], bw);
    } else {
        for (int y = 0; y < h->lh[level]; ++y)
            memcpy(dst + (size_t)y * dstride, &h->pyr[level][(size_t)(y + 19) * bw + 19], h->lw[level]);
    }
    return 0;
}

// CPU baseline: extracts frames[0..n) (contiguous w*h each), frame i on thread i % n_threads, one extractor
// instance per thread (the reference runs one instance on the Tracking thread, src/Frame.cc:616-622).
// Returns wall seconds for `repeat` passes; total_kp receives the keypoint count of one pass.
double orbref_bench(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* frames,
                    int n, int w, int hgt, int n_threads, int canonical, int repeat, long* total_kp) {
    if (n_threads < 1) n_threads = 1;
    std::vector<long> counts(n_threads, 0);
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&, t]() {
            void* hp = orbref_create(nfeatures, scaleFactor, nlevels, iniTh, minTh);
            orbref_set_canonical(hp, canonical);
            long c = 0;
            for (int r = 0; r < repeat; ++r)
                for (int i = t; i < n; i += n_threads) {
                    int k = orbref_extract(hp, frames + (size_t)i * w * hgt, w, hgt, (size_t)w, nullptr, nullptr, 0, 0);
                    if (r == 0) c += k;
                }
            counts[t] = c;
            orbref_destroy(hp);
        });
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    long tot = 0;
    for (long c : counts) tot += c;
    if (total_kp) *total_kp = tot;
    return std::chrono::duration<double>(t1 - t0).count();
}

// Frame-parallel extraction that keeps the results (parity checks over whole sequences, and the single-thread latency
// distribution of the reference: per_frame_secs[i] = wall time of frame i's operator()).  Frame i runs on thread
// i % n_threads; kps / desc are laid out [n][cap]; counts[i] = keypoints of frame i (may exceed cap: then truncated).
double orbref_extract_many(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* frames,
                           int n, int w, int hgt, int n_threads, int canonical, orbref_kp* kps, uint8_t* desc, int cap,
                           int* counts, double* per_frame_secs) {
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&, t]() {
            void* hp = orbref_create(nfeatures, scaleFactor, nlevels, iniTh, minTh);
            orbref_set_canonical(hp, canonical);
            for (int i = t; i < n; i += n_threads) {
                auto a = std::chrono::steady_clock::now();
                const int k = orbref_extract(hp, frames + (size_t)i * w * hgt, w, hgt, (size_t)w,
                                             kps ? kps + (size_t)i * cap : nullptr, desc ? desc + (size_t)i * cap * 32 : nullptr,
                                             cap, 0);
                auto b = std::chrono::steady_clock::now();
                if (counts) counts[i] = k;
                if (per_frame_secs) per_frame_secs[i] = std::chrono::duration<double>(b - a).count();
            }
            orbref_destroy(hp);
        });
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
