/**
 * Drop-in ORB_SLAM2::ORBextractor for EAO-Fusion, backed by libeaof_orb.so (B200, sm_100a).
 *
 * Replaces include/ORBextractor.h + src/ORBextractor.cc of the reference: the public surface below is the one
 * Frame::ExtractORB (src/Frame.cc:616-622), the Frame constructors (scale-table getters) and
 * Frame::ComputeStereoMatches (mvImagePyramid, src/Frame.cc:846,936-953) use, so those files re-link unchanged.
 * All pixel and descriptor work runs on the GPU through the C ABI of include/eaof_orb.h; there is no CPU path —
 * when the library cannot reach a CUDA device the constructor of the first frame's workspace throws.
 *
 * What differs from the reference header, deliberately:
 *   - class ExtractorNode and the protected Compute* / DistributeOctTree members do not exist (they are the
 *     implementation being replaced; nothing outside ORBextractor.cc names them);
 *   - three optional knobs that default to the reference behaviour: SetDevice, SetBlurMode, SetPyramidDownload.
 */
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>
#include <opencv/cv.h>

struct eaof_orb;  // include/eaof_orb.h

namespace ORB_SLAM2
{

class ORBextractor
{
public:

    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels,
                 int iniThFAST, int minThFAST);

    ~ORBextractor();

    // Compute the ORB features and descriptors on an image (8-bit, single channel).
    // Mask is ignored, as in the reference (include/ORBextractor.h:58).
    void operator()( cv::InputArray image, cv::InputArray mask,
      std::vector<cv::KeyPoint>& keypoints,
      cv::OutputArray descriptors);

    int inline GetLevels(){
        return nlevels;}

    float inline GetScaleFactor(){
        return scaleFactor;}

    std::vector<float> inline GetScaleFactors(){
        return mvScaleFactor;
    }

    std::vector<float> inline GetInverseScaleFactors(){
        return mvInvScaleFactor;
    }

    std::vector<float> inline GetScaleSigmaSquares(){
        return mvLevelSigma2;
    }

    std::vector<float> inline GetInverseScaleSigmaSquares(){
        return mvInvLevelSigma2;
    }

    // Level l is a (cols x rows) view at (19,19) inside a host copy of the (cols+38)x(rows+38) bordered buffer,
    // exactly the layout ComputePyramid leaves behind (src/ORBextractor.cc:1114-1116).
    std::vector<cv::Mat> mvImagePyramid;

    // ---- knobs that do not exist in the reference (defaults reproduce it)
    void SetDevice(int cudaDevice);          // before the first frame; default 0 or $EAOF_DEVICE
    void SetBlurMode(int eaofBlurMode);      // EAOF_BLUR_* of eaof_orb.h; default OpenCV 3.3.1 taps or $EAOF_BLUR_MODE
    void SetPyramidDownload(bool on);        // default on; off skips the device->host copy of mvImagePyramid
                                             // (only Frame::ComputeStereoMatches reads it) or $EAOF_PYRAMID=0
    // The library handle holding this extractor's last frame on the device (NULL before the first frame): what the
    // device-side Frame helpers take (eaof_orb_stereo_from_rgbd, eaof_stereo_matches, eaof_voc_transform_orb_device).
    eaof_orb* Handle() const { return mpCtx; }

protected:

    void EnsureWorkspace(int width, int height);

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;

    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    eaof_orb* mpCtx;
    int mnCtxWidth, mnCtxHeight;
    int mnDevice, mnBlurMode;
    bool mbDownloadPyramid;
    std::vector<unsigned char> mvKpStage;   // eaof_kp[cap]
    std::vector<unsigned char> mvDescStage; // cap x 32

private:
    ORBextractor(const ORBextractor&);
    ORBextractor& operator=(const ORBextractor&);
};

} //namespace ORB_SLAM

#endif
