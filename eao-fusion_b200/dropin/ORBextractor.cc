/**
 * Drop-in ORB_SLAM2::ORBextractor: host-side glue of ORBextractor::operator() (reference src/ORBextractor.cc:1043-1105)
 * over the C ABI of libeaof_orb.so.  Everything numerical happens in the CUDA kernels; this file only converts
 * between cv:: types and the plain buffers of include/eaof_orb.h and reproduces the reference's edge behaviour
 * (empty image -> silent return, non-8UC1 -> assert, zero keypoints -> descriptors.release()).
 */
#include "ORBextractor.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "eaof_orb.h"

namespace ORB_SLAM2
{

static const int EDGE_THRESHOLD = 19;  // src/ORBextractor.cc:74

static void Throw(const char* what)
{
    std::string msg = std::string("ORBextractor(eaof): ") + what + ": " + eaof_last_error();
    fprintf(stderr, "%s\n", msg.c_str());
    throw std::runtime_error(msg);  // no CPU fallback, by contract
}

static int EnvInt(const char* name, int dflt)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// ---- which GaussianBlur does the OpenCV this file is compiled against compute? ---------------------------------------
// cv::GaussianBlur(7x7, sigma 2) on 8U is integer arithmetic whose taps and rounding changed between OpenCV versions and
// builds; the library implements the three known forms (EAOF_BLUR_*).  Instead of trusting
// a default, the constructor blurs a fixed probe image with the maintainer's own cv::GaussianBlur, evaluates the three
// integer formulas on the host and selects the one that reproduces it — or throws when none does (an IPP-dispatched
// build, for instance): descriptors would silently differ from the CPU build otherwise.  $EAOF_BLUR_MODE overrides.
namespace
{

int Reflect101(int p, int n)
{
    if(n == 1)
        return 0;
    while(p < 0 || p >= n)
        p = p < 0 ? -p : 2*(n - 1) - p;
    return p;
}

// the integer 7x7 formula in blur mode `mode` at one pixel; `tie` reports an exact .5 before rounding
int BlurFormulaAt(const std::vector<unsigned char>& im, int w, int h, int x, int y, int mode, bool* tie)
{
    static const int t331[7] = {18, 34, 49, 55, 49, 34, 18}, t4[7] = {18, 34, 48, 56, 48, 34, 18};
    const int* k = mode == EAOF_BLUR_CV4 ? t4 : t331;
    int acc = 0;
    for(int j = 0; j < 7; ++j)
    {
        const unsigned char* row = &im[(size_t)Reflect101(y + j - 3, h)*w];
        int hsum = 0;
        for(int i = 0; i < 7; ++i)
            hsum += k[i]*row[Reflect101(x + i - 3, w)];
        acc += k[j]*hsum;
    }
    const int rem = acc & 0xffff;
    if(tie)
        *tie = rem == 32768;
    int q;
    if(mode == EAOF_BLUR_CV331_SSE2 && x < (w & ~3))
    {
        q = acc >> 16;
        q += (rem > 32768) || (rem == 32768 && (q & 1));  // cvtps2dq: round half to even
    }
    else
        q = (acc + 32768) >> 16;
    return q > 255 ? 255 : q;
}

// 64x64 probe: LCG noise, a saturated block, and 7x7 patches searched so that the 3.3.1 taps land exactly on .5 with an
// even integer part — the only place where the two 3.3.1 roundings differ
void MakeBlurProbe(std::vector<unsigned char>& im, int w, int h)
{
    im.assign((size_t)w*h, 0);
    unsigned s = 12345u;
    for(size_t i = 0; i < im.size(); ++i)
    {
        s = s*1664525u + 1013904223u;
        im[i] = (unsigned char)(s >> 24);
    }
    for(int y = 40; y < 52; ++y)
        for(int x = 4; x < 16; ++x)
            im[(size_t)y*w + x] = 255;
    int placed = 0;
    for(int attempt = 0; attempt < 4000000 && placed < 6; ++attempt)
    {
        const int cx = 8 + 9*(placed % 5), cy = 8 + 9*(placed / 5);  // patch centres, 9 px apart: the 7x7 supports are disjoint
        unsigned char patch[49];
        for(int i = 0; i < 49; ++i)
        {
            s = s*1664525u + 1013904223u;
            patch[i] = (unsigned char)(s >> 24);
        }
        for(int j = 0; j < 7; ++j)
            for(int i = 0; i < 7; ++i)
                im[(size_t)(cy + j - 3)*w + cx + i - 3] = patch[7*j + i];
        bool tie = false;
        const int q = BlurFormulaAt(im, w, h, cx, cy, EAOF_BLUR_CV331, &tie);
        if(tie && ((q - 1) & 1) == 0)  // half-up gave q: the integer part q-1 is even, so half-to-even gives q-1
            ++placed;
    }
}

int ProbeBlurMode()
{
    const int w = 64, h = 64;
    std::vector<unsigned char> probe;
    MakeBlurProbe(probe, w, h);
    cv::Mat src(h, w, CV_8UC1), dst;
    for(int y = 0; y < h; ++y)
        memcpy(src.ptr(y), &probe[(size_t)y*w], w);
    cv::GaussianBlur(src, dst, cv::Size(7, 7), 2, 2, cv::BORDER_REFLECT_101);
    const int order[3] = {EAOF_BLUR_CV331, EAOF_BLUR_CV331_SSE2, EAOF_BLUR_CV4};
    for(int m = 0; m < 3; ++m)
    {
        bool same = true;
        for(int y = 0; y < h && same; ++y)
            for(int x = 0; x < w; ++x)
                if(dst.ptr(y)[x] != BlurFormulaAt(probe, w, h, x, y, order[m], NULL))
                {
                    same = false;
                    break;
                }
        if(same)
            return order[m];
    }
    return -1;
}

}  // namespace

#if defined(CV_VERSION_MAJOR) || defined(CV_MAJOR_VERSION)
// The colour entry points of the library (eaof_orb_extract_batch_color, include/eaof_orb.h) take the integer formula of
// cv::cvtColor(BGR2GRAY) explicitly as well; this probe tells a caller which one its OpenCV computes (-1: neither).
// Compiled only against a real OpenCV (the test shim of this repository has no 3-channel Mat).
int ORBextractor::ProbeGrayMode()
{
    cv::Mat bgr(16, 256, CV_8UC3), gray;
    unsigned s = 777u;
    for(int y = 0; y < bgr.rows; ++y)
        for(int x = 0; x < bgr.cols*3; ++x)
        {
            s = s*1664525u + 1013904223u;
            bgr.ptr(y)[x] = (unsigned char)(s >> 24);
        }
    cv::cvtColor(bgr, gray, cv::COLOR_BGR2GRAY);
    bool ok14 = true, ok15 = true;
    for(int y = 0; y < bgr.rows; ++y)
        for(int x = 0; x < bgr.cols; ++x)
        {
            const unsigned char* p = bgr.ptr(y) + 3*x;
            ok14 = ok14 && gray.ptr(y)[x] == ((p[0]*1868 + p[1]*9617 + p[2]*4899 + (1 << 13)) >> 14);
            ok15 = ok15 && gray.ptr(y)[x] == ((p[0]*3735 + p[1]*19235 + p[2]*9798 + (1 << 14)) >> 15);
        }
    return ok14 ? EAOF_GRAY_CV331 : ok15 ? EAOF_GRAY_CV4 : -1;
}
#else
int ORBextractor::ProbeGrayMode() { return -1; }
#endif

ORBextractor::ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
{
    mSetup.features = nfeatures;
    mSetup.factor = scaleFactor;  // a double member initialised from the float argument, like the reference's
    mSetup.levels = nlevels;
    mSetup.fastHigh = iniThFAST;
    mSetup.fastLow = minThFAST;
    mGpu.ctx = NULL;
    mGpu.width = mGpu.height = 0;
    mGpu.device = EnvInt("EAOF_DEVICE", 0);
    mGpu.blurMode = EnvInt("EAOF_BLUR_MODE", -1);
    if(mGpu.blurMode < 0)
    {
        mGpu.blurMode = ProbeBlurMode();
        if(mGpu.blurMode < 0)
        {
            const std::string msg = "ORBextractor(eaof): cv::GaussianBlur(7x7, sigma 2) of this OpenCV build matches none of the "
                                    "integer formulas the library implements (EAOF_BLUR_CV331 / _CV331_SSE2 / _CV4); set "
                                    "EAOF_BLUR_MODE to accept a documented deviation";
            fprintf(stderr, "%s\n", msg.c_str());
            throw std::runtime_error(msg);
        }
    }
    mGpu.downloadPyramid = EnvInt("EAOF_PYRAMID", 1) != 0;

    // The getters must answer before the first frame, so the scale tables are restated here with the arithmetic of
    // src/ORBextractor.cc:415-433 (float table entry times the double factor, narrowed); the library computes the same
    // tables for its kernels and Prepare() cross-checks the two.
    Tables& T = mTables;
    T.scale.assign(nlevels, 1.0f);
    T.sigma2.assign(nlevels, 1.0f);
    for(int l = 1; l < nlevels; ++l)
    {
        T.scale[l] = (float)(T.scale[l-1]*mSetup.factor);
        T.sigma2[l] = T.scale[l]*T.scale[l];
    }
    T.invScale.resize(nlevels);
    T.invSigma2.resize(nlevels);
    for(int l = 0; l < nlevels; ++l)
    {
        T.invScale[l] = 1.0f/T.scale[l];
        T.invSigma2[l] = 1.0f/T.sigma2[l];
    }
    T.quota.assign(nlevels, 0);
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor()
{
    if(mGpu.ctx)
        eaof_orb_destroy(mGpu.ctx);
}

void ORBextractor::SetDevice(int d) { mGpu.device = d; }
void ORBextractor::SetBlurMode(int m) { mGpu.blurMode = m; }
void ORBextractor::SetPyramidDownload(bool on) { mGpu.downloadPyramid = on; }

// (Re)creates the library handle for this frame size and checks its tables against the host restatement.
void ORBextractor::Prepare(int width, int height)
{
    Gpu& G = mGpu;
    if(G.ctx && width == G.width && height == G.height)
        return;
    if(G.ctx)
    {
        eaof_orb_destroy(G.ctx);
        G.ctx = NULL;
    }
    eaof_orb_params p;
    p.nfeatures = mSetup.features;
    p.scale_factor = (float)mSetup.factor;
    p.nlevels = mSetup.levels;
    p.ini_th_fast = mSetup.fastHigh;
    p.min_th_fast = mSetup.fastLow;
    p.blur_mode = G.blurMode;
    p.width = width;
    p.height = height;
    p.max_batch = 1;
    if(eaof_orb_create(&p, G.device, &G.ctx) != EAOF_OK)
        Throw("eaof_orb_create");
    G.width = width;
    G.height = height;

    const int n = mSetup.levels;
    const size_t bytes = sizeof(float)*n;
    std::vector<float> sf(n), isf(n), s2(n), is2(n);
    if(eaof_orb_scale_tables(G.ctx, &sf[0], &isf[0], &s2[0], &is2[0], &mTables.quota[0]) != EAOF_OK)
        Throw("eaof_orb_scale_tables");
    if(memcmp(&sf[0], &mTables.scale[0], bytes) || memcmp(&isf[0], &mTables.invScale[0], bytes) ||
       memcmp(&s2[0], &mTables.sigma2[0], bytes) || memcmp(&is2[0], &mTables.invSigma2[0], bytes))
        Throw("scale tables of the library differ from the host restatement");

    const size_t cap = (size_t)eaof_orb_max_keypoints(G.ctx);
    G.kpStage.resize(sizeof(eaof_kp)*cap);
    G.descStage.resize(32*cap);
}

void ORBextractor::operator()(cv::InputArray imageArray, cv::InputArray, std::vector<cv::KeyPoint>& keypoints,
                              cv::OutputArray descriptorArray)
{
    if(imageArray.empty())
        return;  // the reference returns without touching its outputs (src/ORBextractor.cc:1046-1047)
    cv::Mat image = imageArray.getMat();
    assert(image.type() == CV_8UC1);

    Prepare(image.cols, image.rows);
    Gpu& G = mGpu;
    const int cap = eaof_orb_max_keypoints(G.ctx);
    eaof_kp* kps = reinterpret_cast<eaof_kp*>(&G.kpStage[0]);
    int n = 0;
    if(eaof_orb_extract(G.ctx, image.data, image.cols, image.rows, (size_t)image.step, kps, &G.descStage[0], cap, &n) != EAOF_OK)
        Throw("eaof_orb_extract");

    if(G.downloadPyramid)  // bordered host copies with the level as an ROI view, the layout ComputePyramid leaves
        for(int level = 0; level < mSetup.levels; ++level)
        {
            int w = 0, h = 0;
            eaof_orb_level_size(G.ctx, level, &w, &h);
            cv::Mat bordered(h + 2*EDGE_THRESHOLD, w + 2*EDGE_THRESHOLD, CV_8UC1);
            if(eaof_orb_pyramid_level(G.ctx, 0, level, 1, bordered.data, (size_t)bordered.step) != EAOF_OK)
                Throw("eaof_orb_pyramid_level");
            mvImagePyramid[level] = bordered(cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, w, h));
        }

    keypoints.clear();
    if(n == 0)
    {
        descriptorArray.release();
        return;
    }
    descriptorArray.create(n, 32, CV_8U);
    cv::Mat descriptors = descriptorArray.getMat();
    keypoints.resize(n);
    for(int i = 0; i < n; ++i)
    {
        const eaof_kp& k = kps[i];
        cv::KeyPoint& o = keypoints[i];
        o.pt.x = k.x;
        o.pt.y = k.y;
        o.size = k.size;
        o.angle = k.angle;
        o.response = k.response;
        o.octave = k.octave;
        o.class_id = -1;
    }
    if(descriptors.isContinuous())
        memcpy(descriptors.data, &G.descStage[0], 32*(size_t)n);
    else
        for(int i = 0; i < n; ++i)
            memcpy(descriptors.ptr(i), &G.descStage[32*(size_t)i], 32);
}

}  // namespace ORB_SLAM2
