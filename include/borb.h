/* borb.h — C ABI of the B200-native ORB front-end (libborb.so).
 *
 * Drop-in boundary for the ONE hot path of raulmur/ORB_SLAM2:
 *   ORBextractor::operator()              include/ORBextractor.h:59-61, src/ORBextractor.cc:1043
 *   Frame::ComputeStereoMatches           include/Frame.h:89,            src/Frame.cc:466
 *   ORBmatcher::SearchByProjection (F,MPs) include/ORBmatcher.h:46,       src/ORBmatcher.cc:45
 *   ORBmatcher::SearchByBoW               include/ORBmatcher.h:61-62,    src/ORBmatcher.cc:159,522
 *   ORBmatcher::SearchForTriangulation    include/ORBmatcher.h:69-70,    src/ORBmatcher.cc:657
 *   ORBmatcher::DescriptorDistance        include/ORBmatcher.h:44,       src/ORBmatcher.cc:1647
 *   TemplatedVocabulary::transform (feeder) Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127
 * The reference has no FFI layer; these entry points are what the C++ adapters in
 * include/borb_adapters.hpp (same class signatures as the reference) forward to.
 *
 * Conventions: plain pointers and sizes only; every function returns a borb_status; no exception
 * or process exit crosses the boundary; no CPU fallback exists — without a CUDA device every
 * compute entry point returns BORB_ERR_NO_DEVICE / BORB_ERR_CUDA.  A handle owns one CUDA stream
 * and its scratch; calls on DIFFERENT handles are thread-safe and run concurrently, calls on the
 * same handle must be serialised by the caller (the reference creates one extractor per camera and
 * one matcher per call site/thread, src/Tracking.cc:119-125).
 */
#ifndef BORB_H
#define BORB_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define BORB_API __attribute__((visibility("default")))
#else
#define BORB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BORB_VERSION 1
#define BORB_MAX_LEVELS 16
#define BORB_MAX_DIM 4095 /* image width/height limit (candidates pack x,y in 12 bits) */

typedef enum borb_status {
    BORB_OK = 0,
    BORB_ERR_INVALID_ARG = 1,
    BORB_ERR_NO_DEVICE = 2,   /* no CUDA device / driver: there is no CPU path */
    BORB_ERR_CUDA = 3,        /* a CUDA call failed; see borb_last_error() */
    BORB_ERR_UNSUPPORTED = 4, /* shape / quota outside the supported envelope */
    BORB_ERR_CAPACITY = 5,    /* caller buffer too small; *n_out still holds the required count */
    BORB_ERR_STATE = 6        /* call order violated (e.g. stereo match before extract) */
} borb_status;

/* Layout-identical to cv::KeyPoint (28 bytes) so adapters can memcpy (src/ORBextractor.cc:1103). */
typedef struct borb_keypoint {
    float x, y;      /* pt, level-0 pixel units (level px * mvScaleFactor[octave], :1095-1101) */
    float size;      /* PATCH_SIZE * mvScaleFactor[octave], int-truncated (:837,:846) */
    float angle;     /* degrees [0,360), IC_Angle (:77-104) */
    float response;  /* FAST score */
    int32_t octave;  /* pyramid level */
    int32_t class_id;/* -1 */
} borb_keypoint;

/* ORBextractor ctor arguments (include/ORBextractor.h:53-54; YAML keys ORBextractor.*). */
typedef struct borb_extractor_cfg {
    int32_t n_features;   /* ORBextractor.nFeatures  */
    float scale_factor;   /* ORBextractor.scaleFactor */
    int32_t n_levels;     /* ORBextractor.nLevels (<= BORB_MAX_LEVELS) */
    int32_t ini_th_fast;  /* ORBextractor.iniThFAST */
    int32_t min_th_fast;  /* ORBextractor.minThFAST */
} borb_extractor_cfg;

typedef struct borb_extractor borb_extractor;

BORB_API const char* borb_last_error(void);     /* thread-local description of the last failure */
BORB_API const char* borb_status_str(borb_status s);
BORB_API int borb_version(void);
BORB_API borb_status borb_device_count(int* n);

/* Pinned host memory for asynchronous / overlapped transfers. */
BORB_API borb_status borb_host_alloc(void** p, size_t bytes);
BORB_API borb_status borb_host_free(void* p);

/* ---------------------------------------------------------------- extractor ------------------ */
/* Replaces ORBextractor::ORBextractor (src/ORBextractor.cc:410-470). */
BORB_API borb_status borb_extractor_create(const borb_extractor_cfg* cfg, int device, borb_extractor** out);
BORB_API borb_status borb_extractor_destroy(borb_extractor* e);
/* Getters of include/ORBextractor.h:63-83; arrays hold n_levels entries. */
BORB_API borb_status borb_extractor_tables(const borb_extractor* e, float* scale, float* inv_scale, float* sigma2,
                                  float* inv_sigma2, int32_t* features_per_level);
/* Upper bound of keypoints one image can return for this cfg and image size (quota may be
 * exceeded by <=3 per level and is never trimmed, src/ORBextractor.cc:730). */
BORB_API borb_status borb_extractor_capacity(const borb_extractor* e, int width, int height, int* cap);
/* Pre-allocates device scratch for batches of up to max_images images of width x height. */
BORB_API borb_status borb_extractor_reserve(borb_extractor* e, int width, int height, int max_images);

/* ORBextractor::operator() (src/ORBextractor.cc:1043) for ONE host image.  gray: 8-bit, `stride`
 * bytes per row.  kps/desc: caller buffers of `cap` entries (desc: cap x 32 bytes, row-major like the
 * reference's N x 32 CV_8U Mat).  Empty image => BORB_OK with *n_out = 0 (:1046-1047). */
BORB_API borb_status borb_extract(borb_extractor* e, const uint8_t* gray, int width, int height, int stride,
                         borb_keypoint* kps, uint8_t* desc, int cap, int* n_out);

/* The same for n images of identical size in one launch sequence (the throughput path: independent
 * frames / camera streams batched per launch).  gray[i] are host pointers; image i writes
 * kps + i*cap, desc + i*cap*32 and n_out[i]. */
BORB_API borb_status borb_extract_batch(borb_extractor* e, const uint8_t* const* gray, int n_images, int width, int height,
                               int stride, borb_keypoint* kps, uint8_t* desc, int cap, int* n_out);

/* Asynchronous halves of the batch call: `_enqueue` returns once the work is queued on the handle's
 * stream (host buffers must stay valid and should be pinned); `borb_sync` waits for it.  Results of
 * the last batch stay resident in HBM until the next enqueue on this handle. */
BORB_API borb_status borb_extract_batch_enqueue(borb_extractor* e, const uint8_t* const* gray, int n_images, int width,
                                       int height, int stride, borb_keypoint* kps, uint8_t* desc, int cap, int* n_out);
BORB_API borb_status borb_sync(borb_extractor* e);

/* Device-resident input variant: d_gray points to n_images images already in HBM on the handle's
 * device (image i at d_gray + i*image_stride, rows `pitch` bytes apart).  No host copies unless the
 * output pointers are non-NULL.  Used for the HBM-resident throughput measurement. */
BORB_API borb_status borb_extract_batch_device(borb_extractor* e, const uint8_t* d_gray, int n_images, int width, int height,
                                      size_t pitch, size_t image_stride, borb_keypoint* kps, uint8_t* desc, int cap,
                                      int* n_out);

/* mvImagePyramid[level] of image `image` of the last batch (include/ORBextractor.h:85), copied to
 * a caller buffer of at least h*w bytes (tight rows).  Pass dst=NULL to query w/h only. */
BORB_API borb_status borb_extractor_pyramid(borb_extractor* e, int image, int level, uint8_t* dst, int* w, int* h);

/* ---------------------------------------------------------------- stereo --------------------- */
/* Frame::ComputeStereoMatches (src/Frame.cc:466-640) for the images of the LAST batch of `e`:
 * pair p uses image left_idx[p] as left and right_idx[p] as right (NULL index arrays mean
 * left=2p, right=2p+1).  bf = Camera.bf, b = mb = bf/fx (src/Frame.cc:114; the reference reads mb
 * before initialising it, :496 — the intended value is used here).  Outputs per pair p at
 * u_right + p*cap and depth + p*cap, entries [0, n_left): -1.0f means "no match" (:468-469). */
BORB_API borb_status borb_stereo_match(borb_extractor* e, int n_pairs, const int* left_idx, const int* right_idx, float bf,
                              float b, float* u_right, float* depth, int cap);

/* Same, when the left and right images were extracted by two different handles on the same device
 * (the reference's mpORBextractorLeft / mpORBextractorRight, src/Frame.cc:78-81): image 0 of each. */
BORB_API borb_status borb_stereo_match2(borb_extractor* left, borb_extractor* right, float bf, float b, float* u_right,
                               float* depth, int cap);

/* Frame::Frame stereo constructor hot path in one call (src/Frame.cc:61-117): extract L+R for
 * n_pairs frames and associate them; one H2D (images) and one D2H (results) per call.
 * Any output pointer may be NULL to skip its copy. */
BORB_API borb_status borb_stereo_frames(borb_extractor* e, const uint8_t* const* left, const uint8_t* const* right, int n_pairs,
                               int width, int height, int stride, float bf, float b, borb_keypoint* kps_left,
                               uint8_t* desc_left, int* n_left, borb_keypoint* kps_right, uint8_t* desc_right,
                               int* n_right, float* u_right, float* depth, int cap);
BORB_API borb_status borb_stereo_frames_enqueue(borb_extractor* e, const uint8_t* const* left, const uint8_t* const* right,
                                       int n_pairs, int width, int height, int stride, float bf, float b,
                                       borb_keypoint* kps_left, uint8_t* desc_left, int* n_left,
                                       borb_keypoint* kps_right, uint8_t* desc_right, int* n_right, float* u_right,
                                       float* depth, int cap);
/* HBM-resident variants (inputs as in borb_extract_batch_device: image 2p = left, 2p+1 = right). */
BORB_API borb_status borb_stereo_frames_device_enqueue(borb_extractor* e, const uint8_t* d_gray, int n_pairs, int width,
                                              int height, size_t pitch, size_t image_stride, float bf, float b,
                                              int* n_left, int* n_right, float* u_right, float* depth, int cap);
BORB_API borb_status borb_stereo_frames_device(borb_extractor* e, const uint8_t* d_gray, int n_pairs, int width, int height,
                                      size_t pitch, size_t image_stride, float bf, float b, int* n_left, int* n_right,
                                      float* u_right, float* depth, int cap);

/* ---------------------------------------------------------------- introspection -------------- */
/* Per-stage intermediates of the last batch, for parity tests (tests/ compare each stage with the
 * oracle).  xys: (x, y, score) int32 triples in level pixel coordinates. */
BORB_API borb_status borb_debug_candidates(borb_extractor* e, int image, int level, int32_t* xys, int cap, int* n_out);
BORB_API borb_status borb_debug_selected(borb_extractor* e, int image, int level, int32_t* xys, int cap, int* n_out);
BORB_API borb_status borb_debug_blurred(borb_extractor* e, int image, int level, uint8_t* dst, int* w, int* h);
/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
BORB_API borb_status borb_launch_count(const borb_extractor* e, uint64_t* n);
/* Device time (ms, CUDA events on the handle's stream) of each stage of the last batch:
 * [0] upload, [1] pyramid, [2] FAST/NMS, [3] quadtree, [4] blur, [5] orient+rBRIEF, [6] stereo, [7] download. */
BORB_API borb_status borb_stage_times(borb_extractor* e, float* ms8);
BORB_API borb_status borb_set_timing(borb_extractor* e, int enable);   /* also resets the accumulators */
/* Per-stage device time summed over every step completed (synced) since borb_set_timing(e,1); steps may
 * be queued back to back without host syncs (a ring of CUDA events on the handle's stream). */
BORB_API borb_status borb_stage_times_total(borb_extractor* e, double* ms8, uint64_t* steps);
/* The handle's cudaStream_t, so callers can bracket work with their own CUDA events. */
BORB_API borb_status borb_extractor_stream(borb_extractor* e, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* BORB_H */
