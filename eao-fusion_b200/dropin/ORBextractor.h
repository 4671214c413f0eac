/**
 * Drop-in ORB_SLAM2::ORBextractor for EAO-Fusion, backed by libeaof_orb.so (B200, sm_100a).
 *
 * Replaces include/ORBextractor.h + src/ORBextractor.cc of the reference.  The public members are exactly the ones the
 * rest of the reference names — Frame::ExtractORB (src/Frame.cc:616-622), the Frame constructors' table getters
 * (:104-110, :184-190) and Frame::ComputeStereoMatches reading mvImagePyramid (:846, :936-953) — with the same
 * signatures, so those translation units re-link unchanged.  All pixel and descriptor work runs on the GPU through the
 * C ABI of include/eaof_orb.h; there is no CPU path: when the library cannot reach a CUDA device, the first frame throws.
 *
 * Not here, deliberately: class ExtractorNode and the Compute* / DistributeOctTree members of the reference header (they
 * are the implementation being replaced; nothing outside ORBextractor.cc names them).  Added: knobs that default to
 * the reference behaviour (SetDevice, SetBlurMode, SetPyramidDownload, Handle).  The blur arithmetic is selected by probing
 * the cv::GaussianBlur of the OpenCV this file is compiled against (see ORBextractor.cc).
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
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();

    /// ORB keypoints + descriptors of an 8-bit single-channel image; the mask is ignored, as in the reference.
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                    cv::OutputArray descriptors);

    // pyramid tables, returned by value like the reference's inline getters
    int GetLevels() { return mSetup.levels; }
    float GetScaleFactor() { return mSetup.factor; }
    std::vector<float> GetScaleFactors() { return mTables.scale; }
    std::vector<float> GetInverseScaleFactors() { return mTables.invScale; }
    std::vector<float> GetScaleSigmaSquares() { return mTables.sigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mTables.invSigma2; }

    /// Level l = a (cols x rows) view at (19,19) of a host copy of the bordered (cols+38) x (rows+38) buffer.
    std::vector<cv::Mat> mvImagePyramid;

    // ---- not in the reference; the defaults reproduce it
    void SetDevice(int cudaDevice);      ///< before the first frame; default 0 or $EAOF_DEVICE
    void SetBlurMode(int eaofBlurMode);  ///< EAOF_BLUR_*; default: probed from this build's cv::GaussianBlur, or $EAOF_BLUR_MODE
    int BlurMode() const { return mGpu.blurMode; }
    /// EAOF_GRAY_* that reproduces this build's cv::cvtColor(BGR2GRAY) (-1: neither), for the colour entry points
    static int ProbeGrayMode();
    void SetPyramidDownload(bool on);    ///< off: mvImagePyramid is not copied back (only stereo reads it); $EAOF_PYRAMID=0
    /// Library handle that holds the last frame on the device (NULL before the first frame): what the device-side
    /// Frame helpers take (eaof_orb_stereo_from_rgbd, eaof_stereo_matches, eaof_voc_transform_orb_device, ...).
    eaof_orb* Handle() const { return mGpu.ctx; }

private:
    ORBextractor(const ORBextractor&);
    ORBextractor& operator=(const ORBextractor&);
    void Prepare(int width, int height);

    struct Setup { int features; double factor; int levels, fastHigh, fastLow; } mSetup;
    struct Tables { std::vector<float> scale, invScale, sigma2, invSigma2; std::vector<int> quota; } mTables;
    struct Gpu
    {
        eaof_orb* ctx;
        int width, height, device, blurMode;
        bool downloadPyramid;
        std::vector<unsigned char> kpStage, descStage;  // eaof_kp[cap], cap x 32 bytes
    } mGpu;
};

}  // namespace ORB_SLAM2

#endif  // ORBEXTRACTOR_H
