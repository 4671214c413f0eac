/*
 * eaof_orb.h — C ABI of libeaof_orb.so, the B200 (sm_100a) ORB front-end.
 *
 * The reference (ChenJiahao031008/EAO-Fusion) has no plugin/FFI layer: the boundary of its ORB hot path is the
 * C++ class interface in include/ORBextractor.h:45-111 and include/ORBmatcher.h:37-102.  Every entry point below
 * names the reference interface it replaces; eao-fusion_b200/dropin/ re-creates those two classes on top of this
 * ABI so that Frame.cc / Tracking.cc re-link without source changes.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative EAOF_ERR_* code,
 * with a human-readable reason available from eaof_last_error() (thread-local).  There is NO CPU fallback: without
 * a CUDA device eaof_orb_create fails with EAOF_ERR_CUDA.  One handle = one CUDA stream + one device workspace;
 * calls on one handle must be serialised by the caller, different handles are independent (the reference runs two
 * extractor instances concurrently for stereo, src/Frame.cc:113-116).
 */
#ifndef EAOF_ORB_H
#define EAOF_ORB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EAOF_ABI_VERSION 1
#define EAOF_MAX_LEVELS 16

enum {
    EAOF_OK = 0,
    EAOF_ERR_ARG = -1,         /* null pointer, bad size, capacity too small */
    EAOF_ERR_CUDA = -2,        /* CUDA runtime/driver failure (no device, launch error, out of memory) */
    EAOF_ERR_UNSUPPORTED = -3, /* shape the reference itself has undefined behaviour for */
    EAOF_ERR_EMPTY = -4,       /* empty image: the reference returns silently, src/ORBextractor.cc:1046-1047 */
    EAOF_ERR_NCCL = -5,
    EAOF_ERR_BUSY = -6         /* frame ring full (the consumer is behind): drop the frame, like a ROS queue_size does */
};

/* Gaussian-blur arithmetic: OpenCV-version-dependent, so explicit. */
enum {
    EAOF_BLUR_CV331 = 0,      /* OpenCV 3.3.1 8U path: taps 18,34,49,55,49,34,18 (sum 257), (S+2^15)>>16 [default] */
    EAOF_BLUR_CV4 = 1,        /* OpenCV >=3.4.1/4.x: taps 18,34,48,56,48,34,18 (sum 256), (S+2^15)>>16 */
    EAOF_BLUR_CV331_SSE2 = 2  /* 3.3.1 taps, columns x < 4*floor(w/4) rounded half-to-even (SSE2 float column pass) */
};

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * include/ORBextractor.h:51-52, src/ORBextractor.cc:410-470 — plus workspace sizing. */
typedef struct eaof_orb_params {
    int nfeatures;
    float scale_factor;
    int nlevels;
    int ini_th_fast;
    int min_th_fast;
    int blur_mode;  /* EAOF_BLUR_* */
    int width;      /* frame size the workspace is built for (all frames of a handle share it) */
    int height;
    int max_batch;  /* frames per batched call (>= 1) */
} eaof_orb_params;

/* One output keypoint: the fields of cv::KeyPoint the reference fills (src/ORBextractor.cc:837-847,1095-1101);
 * class_id is always -1 there and is not transported. */
typedef struct eaof_kp {
    float x, y;     /* pt, already multiplied by mvScaleFactor[octave] */
    float size;     /* (int)(31 * mvScaleFactor[octave]) */
    float angle;    /* IC_Angle, degrees in [0,360) */
    float response; /* FAST score */
    int octave;
} eaof_kp;

typedef struct eaof_orb eaof_orb;

const char* eaof_last_error(void);
int eaof_abi_version(void);

int eaof_orb_create(const eaof_orb_params* params, int device, eaof_orb** out);
void eaof_orb_destroy(eaof_orb* ctx);

/* Upper bound on keypoints per frame (sum over levels of max(quota+3, 4*nIni)); size kps/desc with it. */
int eaof_orb_max_keypoints(const eaof_orb* ctx);

/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:66-80) and mnFeaturesPerLevel; arrays of nlevels entries, any may be NULL. */
int eaof_orb_scale_tables(const eaof_orb* ctx, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                          int* features_per_level);
/* Inner size of pyramid level `level` (src/ORBextractor.cc:1112). */
int eaof_orb_level_size(const eaof_orb* ctx, int level, int* width, int* height);

/* ORBextractor::operator()(image, mask, keypoints, descriptors)  src/ORBextractor.cc:1043-1105, one frame,
 * HOST buffers, synchronous.  img: 8-bit single channel, `stride` bytes per row.  kps/desc hold `cap` entries
 * (desc: cap x 32 bytes, row i belongs to kps[i]).  *n_out receives the count. */
int eaof_orb_extract(eaof_orb* ctx, const uint8_t* img, int width, int height, size_t stride, eaof_kp* kps,
                     uint8_t* desc, int cap, int* n_out);

/* Batched form for frame sequences: n_frames <= max_batch images of width x height, image f at
 * imgs + f*frame_pitch.  Outputs: frame f's keypoints at kps + f*cap, descriptors at desc + f*cap*32,
 * count in n_out[f].  HOST buffers (pinned memory recommended), synchronous. */
int eaof_orb_extract_batch(eaof_orb* ctx, const uint8_t* imgs, int n_frames, int width, int height, size_t stride,
                           size_t frame_pitch, eaof_kp* kps, uint8_t* desc, int cap, int* n_out);

/* The same call in two halves, so that a caller can keep several handles busy (upload of one batch under the kernels of
 * another): _async enqueues upload, kernels and download and returns; _wait blocks until the outputs passed to _async
 * are complete and delivers the per-frame counts.  One batch in flight per handle.  Buffers should be pinned. */
int eaof_orb_extract_batch_async(eaof_orb* ctx, const uint8_t* imgs, int n_frames, int width, int height, size_t stride,
                                 size_t frame_pitch, eaof_kp* kps, uint8_t* desc, int cap);
int eaof_orb_extract_batch_wait(eaof_orb* ctx, int* n_out);
/* Frames per chunk of the upload / kernels / download pipeline inside the host-buffer calls (0 = default heuristic;
 * >= n_frames disables chunking).  Also settable at create time through $EAOF_CHUNK. */
int eaof_orb_set_pipeline_chunk(eaof_orb* ctx, int frames);

/* Same, with the frames already resident in device memory (d_imgs, tightly packed rows of `stride` bytes); results
 * stay on the device.  Asynchronous on the handle's stream; use eaof_orb_sync or the accessors below. */
int eaof_orb_extract_batch_device(eaof_orb* ctx, const uint8_t* d_imgs, int n_frames, int width, int height,
                                  size_t stride, size_t frame_pitch);
int eaof_orb_sync(eaof_orb* ctx);

/* Colour ingest: Tracking::GrabImageRGBD / GrabImageMonocular convert the camera image with
 * cv::cvtColor(CV_RGB2GRAY | CV_BGR2GRAY | CV_RGBA2GRAY | CV_BGRA2GRAY) before the Frame constructor calls the extractor
 * (src/Tracking.cc:324-337).  These entry points take the interleaved 8-bit colour frames and do that conversion on the
 * device inside the level-0 pass; everything after it is the gray path.  The gray image (mImGray) is pyramid level 0,
 * eaof_orb_pyramid_level(ctx, frame, 0, 0, ...).  `stride` / `frame_pitch` in bytes.  gray_mode: the integer formula of
 * OpenCV's RGB2Gray<uchar>, which changed between versions like the blur taps did. */
enum { EAOF_COLOR_BGR = 0, EAOF_COLOR_RGB = 1, EAOF_COLOR_BGRA = 2, EAOF_COLOR_RGBA = 3 };
enum {
    EAOF_GRAY_CV331 = 0, /* OpenCV 3.3.1: (B*1868 + G*9617 + R*4899 + 2^13) >> 14 [the reference's pinned version] */
    EAOF_GRAY_CV4 = 1    /* OpenCV 4.x:   (B*3735 + G*19235 + R*9798 + 2^14) >> 15 (verified against cv2 4.13) */
};
int eaof_orb_extract_batch_device_color(eaof_orb* ctx, const uint8_t* d_imgs, int n_frames, int width, int height,
                                        size_t stride, size_t frame_pitch, int color, int gray_mode);
int eaof_orb_extract_batch_color(eaof_orb* ctx, const uint8_t* imgs, int n_frames, int width, int height, size_t stride,
                                 size_t frame_pitch, int color, int gray_mode, eaof_kp* kps, uint8_t* desc, int cap,
                                 int* n_out);

/* Pinned frame ring (the ingest side).  The reference's camera callback copies every ROS image message
 * into a fresh pageable cv::Mat (cv_bridge::toCvCopy, ros_test/src/message_flow.cc:250-254) that Tracking::GrabImageRGBD
 * converts and hands to the Frame constructor (src/Tracking.cc:324-337, src/Frame.cc:193).  With a ring the callback writes
 * the frame into a slot of page-locked memory instead, so the extractor's upload is an asynchronous DMA straight from the
 * slot (a pageable image costs 19-30 us of driver staging per 640x480 frame) and a tracker that has fallen
 * behind takes every pending frame in ONE batched call.  Single producer (acquire -> fill -> commit), single consumer
 * (pending -> peek / extract_ring -> release); the two sides may be different threads, no lock is taken.
 * channels = 1 (gray), 3 or 4 (interleaved colour, converted on the device by the colour path above).
 * Rows are tightly packed (stride = width * channels), slots are 4096-byte aligned. */
typedef struct eaof_ring eaof_ring;
int eaof_ring_create(int slots, int width, int height, int channels, eaof_ring** out);
void eaof_ring_destroy(eaof_ring* ring);
/* producer: the next free slot (EAOF_ERR_BUSY when all slots hold unreleased frames), then publish it with its time stamp */
int eaof_ring_acquire(eaof_ring* ring, uint8_t** slot, size_t* stride);
int eaof_ring_commit(eaof_ring* ring, double timestamp);
/* consumer: number of committed, unreleased frames; the k-th oldest of them; give the n oldest back to the producer */
int eaof_ring_pending(const eaof_ring* ring);
int eaof_ring_peek(const eaof_ring* ring, int k, const uint8_t** slot, double* timestamp);
int eaof_ring_release(eaof_ring* ring, int n);
/* Extract the oldest pending frames: up to max_frames of them (and at most max_batch, and only up to the end of the slot
 * array: a run that wraps around is taken by the next call).  *n_frames receives how many were taken (0: nothing pending);
 * outputs as in eaof_orb_extract_batch; color / gray_mode are used by colour rings only.  The frames stay in the ring
 * until eaof_ring_release (Tracking keeps mImGray for the viewer and the semantic thread). */
int eaof_orb_extract_ring(eaof_orb* ctx, eaof_ring* ring, int max_frames, int color, int gray_mode, eaof_kp* kps,
                          uint8_t* desc, int cap, int* n_out, double* timestamps, int* n_frames);

/* Frame::UndistortKeyPoints()  src/Frame.cc:773-803 for the keypoints of the last batch: mvKeysUn
 * positions = cv::undistortPoints(mvKeys, mK, mDistCoef, cv::Mat(), mK) with mK = (fx, fy, cx, cy) and `dist` = k1 k2 p1 p2
 * [k3 k4 k5 k6 s1 s2 s3 s4] (n_dist of them), computed in double like OpenCV; with no coefficients or k1 == 0 the
 * positions are copied, which is the reference's shortcut (:775-779).  mode: EAOF_UNDISTORT_CV331 = OpenCV 3.3.1 (five
 * iterations, no exit), EAOF_UNDISTORT_CV4 = OpenCV 4.x (adds the "icdist < 0" exit; checked against cv2 4.13) — they
 * agree wherever the distortion polynomial stays positive, i.e. for every real calibration.
 * _device: outputs [n_frames][eaof_orb_max_keypoints()] floats in device memory (what eaof_orb_stereo_from_rgbd_device
 * takes as d_x_undistorted), asynchronous; host form: [n_frames][cap]. */
enum { EAOF_UNDISTORT_CV331 = 0, EAOF_UNDISTORT_CV4 = 1 };
int eaof_orb_undistort_keypoints_device(eaof_orb* ctx, int n_frames, float fx, float fy, float cx, float cy,
                                        const float* dist, int n_dist, int mode, float* d_x_un, float* d_y_un);
int eaof_orb_undistort_keypoints(eaof_orb* ctx, int n_frames, float fx, float fy, float cx, float cy, const float* dist,
                                 int n_dist, int mode, float* x_un, float* y_un, int cap);

/* Frame::ComputeStereoFromRGBD(imDepth)  src/Frame.cc:1016-1037 for the keypoints of the last batch:
 * per keypoint the depth d at its raw position (coordinates truncated like cv::Mat::at<float>(float, float)); where
 * d > 0: mvDepth = d, mvuRight = x_undistorted - mbf/d; -1 elsewhere.  depth_type EAOF_DEPTH_F32: the CV_32F map the
 * reference holds; EAOF_DEPTH_U16: the raw 16-bit sensor map, converted as Tracking does with
 * imDepth.convertTo(CV_32F, mDepthMapFactor) (src/Tracking.cc:340-341): d = (float)raw * depth_scale.
 * _device: depth map and outputs in device memory ([n_frames][eaof_orb_max_keypoints()] floats), asynchronous on the
 * handle's stream; d_x_undistorted = mvKeysUn x per keypoint in the same layout, NULL when the camera has no distortion
 * (mvKeysUn = mvKeys, src/Frame.cc:775-779).  Host form: depth maps and outputs ([n_frames][cap]) in host memory. */
enum { EAOF_DEPTH_F32 = 0, EAOF_DEPTH_U16 = 1 };
int eaof_orb_stereo_from_rgbd_device(eaof_orb* ctx, int n_frames, const void* d_depth, int depth_type, float depth_scale,
                                     size_t stride_bytes, size_t frame_pitch_bytes, const float* d_x_undistorted,
                                     float mbf, float* d_uright, float* d_depth_out);
int eaof_orb_stereo_from_rgbd(eaof_orb* ctx, int n_frames, const void* depth, int depth_type, float depth_scale,
                              size_t stride_bytes, size_t frame_pitch_bytes, float mbf, float* uright, float* depth_out,
                              int cap);

/* Frame::ComputeStereoMatches()  src/Frame.cc:841-1013 for frames 0..n_frames-1 of the last batches of
 * two handles, `left` = mpORBextractorLeft and `right` = mpORBextractorRight (same frame size and pyramid parameters,
 * same device): per left keypoint the best right keypoint by Hamming distance in its row band (+-2*scale rows, octave
 * within +-1, u in [uL - mbf/mb, uL + 3], distance < TH_HIGH), refined on the two image pyramids — the only readers of
 * mvImagePyramid — by the 11x11 SAD over shifts -5..5 and a parabola fit; then the 1.5*1.4*median SAD filter.
 * mvuRight / mvDepth per left keypoint, -1 where unmatched, laid out like eaof_orb_stereo_from_rgbd's outputs.
 * _device: outputs in device memory ([n_frames][eaof_orb_max_keypoints(left)]), asynchronous on left's stream (which waits
 * for right's batch; right's next batch waits for this call).  No pyramid ever travels to the host. */
int eaof_stereo_matches_device(eaof_orb* left, eaof_orb* right, int n_frames, float mb, float mbf, float* d_uright,
                               float* d_depth);
int eaof_stereo_matches(eaof_orb* left, eaof_orb* right, int n_frames, float mb, float mbf, float* uright, float* depth_out,
                        int cap);

/* Device-resident results of the last batch: kps[f*cap_out + i], desc[(f*cap_out + i)*32], counts[f].
 * cap_out == eaof_orb_max_keypoints(). */
int eaof_orb_device_results(eaof_orb* ctx, const eaof_kp** d_kps, const uint8_t** d_desc, const int** d_counts,
                            int* cap_out);
/* Copies the last batch's results to host buffers laid out as in eaof_orb_extract_batch. */
int eaof_orb_fetch_results(eaof_orb* ctx, int n_frames, eaof_kp* kps, uint8_t* desc, int cap, int* n_out);

/* mvImagePyramid[level] of frame `frame` of the last call (include/ORBextractor.h:85): copies the inner
 * width x height image (with_border = 0) or the (w+38)x(h+38) bordered buffer to host memory. */
int eaof_orb_pyramid_level(eaof_orb* ctx, int frame, int level, int with_border, uint8_t* dst, size_t dst_stride);

/* ---- stage dumps for parity tests (debug; synchronous) ------------------------------------------------- */
/* Blurred inner level (input of computeOrbDescriptor, src/ORBextractor.cc:1085-1086). */
int eaof_orb_debug_blurred_level(eaof_orb* ctx, int frame, int level, uint8_t* dst, size_t dst_stride);
/* FAST candidates of (frame, level) before DistributeOctTree, unordered: triples (x, y, score) in detection-window
 * coordinates (src/ORBextractor.cc:822-824).  Returns the count through *n_out (may exceed cap). */
int eaof_orb_debug_candidates(eaof_orb* ctx, int frame, int level, int* xys, int cap, int* n_out);
/* Runs the device restatement of glibc sinf/cosf (used by the descriptor rotation) over every `stride`-th float in
 * [0, hi] and compares with this host's libm; returns the number of mismatching results (0 expected), -1 on error. */
long eaof_debug_sincosf_mismatches(float hi, uint32_t stride);
/* Host timestamps (seconds, steady clock) of the last single-frame call on this handle: entered / upload queued / graph
 * launched / wait entered / stream synchronised / results copied out.  Measurement aid for the latency path. */
int eaof_debug_latency_trace(eaof_orb* h, double* out6);
/* Per-stage GPU time of the last batched call in milliseconds (CUDA events on the handle's stream):
 * [0] pyramid, [1] FAST, [2] octree, [3] blur, [4] angle+descriptor, [5] total.  Requires
 * eaof_orb_set_profiling(ctx, 1) before the call. */
int eaof_orb_set_profiling(eaof_orb* ctx, int on);
int eaof_orb_stage_times(eaof_orb* ctx, float* ms6);
/* Number of kernel launches the last batched call issued. */
int eaof_orb_last_launch_count(const eaof_orb* ctx);
/* The handle's cudaStream_t, so callers can record their own CUDA events around batched calls. */
void* eaof_orb_stream(eaof_orb* ctx);

#ifdef __cplusplus
}
#endif
#endif /* EAOF_ORB_H */
