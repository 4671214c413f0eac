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
    mGpu.blurMode = EnvInt("EAOF_BLUR_MODE", EAOF_BLUR_CV331);
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
    keypoints.reserve(n);
    for(int i = 0; i < n; ++i)
    {
        const eaof_kp& k = kps[i];
        keypoints.push_back(cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, -1));
        memcpy(descriptors.ptr(i), &G.descStage[32*(size_t)i], 32);
    }
}

}  // namespace ORB_SLAM2
