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

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels,
         int _iniThFAST, int _minThFAST):
    nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels),
    iniThFAST(_iniThFAST), minThFAST(_minThFAST),
    mpCtx(NULL), mnCtxWidth(0), mnCtxHeight(0),
    mnDevice(EnvInt("EAOF_DEVICE", 0)), mnBlurMode(EnvInt("EAOF_BLUR_MODE", EAOF_BLUR_CV331)),
    mbDownloadPyramid(EnvInt("EAOF_PYRAMID", 1) != 0)
{
    // The getters must answer before the first frame (Frame's constructor reads them right after ExtractORB, but
    // Tracking may query earlier), so the scale tables are restated on the host exactly as
    // src/ORBextractor.cc:415-446 computes them; the library computes the same tables for the kernels and
    // EnsureWorkspace cross-checks the two.
    mvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvScaleFactor[0]=1.0f;
    mvLevelSigma2[0]=1.0f;
    for(int i=1; i<nlevels; i++)
    {
        mvScaleFactor[i]=(float)(mvScaleFactor[i-1]*scaleFactor);
        mvLevelSigma2[i]=mvScaleFactor[i]*mvScaleFactor[i];
    }
    mvInvScaleFactor.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    for(int i=0; i<nlevels; i++)
    {
        mvInvScaleFactor[i]=1.0f/mvScaleFactor[i];
        mvInvLevelSigma2[i]=1.0f/mvLevelSigma2[i];
    }
    mvImagePyramid.resize(nlevels);
    mnFeaturesPerLevel.assign(nlevels, 0);
}

ORBextractor::~ORBextractor()
{
    if(mpCtx)
        eaof_orb_destroy(mpCtx);
}

void ORBextractor::SetDevice(int d) { mnDevice = d; }
void ORBextractor::SetBlurMode(int m) { mnBlurMode = m; }
void ORBextractor::SetPyramidDownload(bool on) { mbDownloadPyramid = on; }

void ORBextractor::EnsureWorkspace(int width, int height)
{
    if(mpCtx && width==mnCtxWidth && height==mnCtxHeight)
        return;
    if(mpCtx)
    {
        eaof_orb_destroy(mpCtx);
        mpCtx = NULL;
    }
    eaof_orb_params p;
    p.nfeatures = nfeatures;
    p.scale_factor = (float)scaleFactor;
    p.nlevels = nlevels;
    p.ini_th_fast = iniThFAST;
    p.min_th_fast = minThFAST;
    p.blur_mode = mnBlurMode;
    p.width = width;
    p.height = height;
    p.max_batch = 1;
    if(eaof_orb_create(&p, mnDevice, &mpCtx)!=EAOF_OK)
        Throw("eaof_orb_create");
    mnCtxWidth = width;
    mnCtxHeight = height;

    std::vector<float> sf(nlevels), isf(nlevels), s2(nlevels), is2(nlevels);
    if(eaof_orb_scale_tables(mpCtx, &sf[0], &isf[0], &s2[0], &is2[0], &mnFeaturesPerLevel[0])!=EAOF_OK)
        Throw("eaof_orb_scale_tables");
    if(memcmp(&sf[0], &mvScaleFactor[0], sizeof(float)*nlevels) || memcmp(&isf[0], &mvInvScaleFactor[0], sizeof(float)*nlevels) ||
       memcmp(&s2[0], &mvLevelSigma2[0], sizeof(float)*nlevels) || memcmp(&is2[0], &mvInvLevelSigma2[0], sizeof(float)*nlevels))
        Throw("scale tables of the library differ from the host restatement");

    const int cap = eaof_orb_max_keypoints(mpCtx);
    mvKpStage.resize(sizeof(eaof_kp)*(size_t)cap);
    mvDescStage.resize(32*(size_t)cap);
}

void ORBextractor::operator()( cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                      cv::OutputArray _descriptors)
{
    (void)_mask;
    if(_image.empty())
        return;

    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1 );

    EnsureWorkspace(image.cols, image.rows);

    const int cap = eaof_orb_max_keypoints(mpCtx);
    eaof_kp* kps = reinterpret_cast<eaof_kp*>(&mvKpStage[0]);
    int n = 0;
    if(eaof_orb_extract(mpCtx, image.data, image.cols, image.rows, (size_t)image.step, kps, &mvDescStage[0], cap, &n)!=EAOF_OK)
        Throw("eaof_orb_extract");

    // mvImagePyramid: bordered host copies with the level as an ROI view, as ComputePyramid builds them
    if(mbDownloadPyramid)
    {
        for(int level=0; level<nlevels; ++level)
        {
            int w=0, h=0;
            eaof_orb_level_size(mpCtx, level, &w, &h);
            cv::Mat temp(h+EDGE_THRESHOLD*2, w+EDGE_THRESHOLD*2, CV_8UC1);
            if(eaof_orb_pyramid_level(mpCtx, 0, level, 1, temp.data, (size_t)temp.step)!=EAOF_OK)
                Throw("eaof_orb_pyramid_level");
            mvImagePyramid[level] = temp(cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, w, h));
        }
    }

    _keypoints.clear();
    if(n==0)
    {
        _descriptors.release();
        return;
    }
    _descriptors.create(n, 32, CV_8U);
    cv::Mat descriptors = _descriptors.getMat();
    _keypoints.reserve(n);
    for(int i=0; i<n; ++i)
    {
        const eaof_kp& k = kps[i];
        _keypoints.push_back(cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, -1));
        memcpy(descriptors.ptr(i), &mvDescStage[32*(size_t)i], 32);
    }
}

} //namespace ORB_SLAM
