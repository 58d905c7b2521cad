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
 *   Frame::AssignFeaturesToGrid / GetFeaturesInArea  src/Frame.cc:230-245, 327-380
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

#define BORB_VERSION 2
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

/* Pixel format of the HOST images given to the extract / stereo entry points of this handle (default 1 = CV_8UC1).
 * With 3 or 4 channels the conversion Tracking::GrabImageStereo/RGBD/Monocular performs before calling the extractor —
 * cv::cvtColor(RGB2GRAY | BGR2GRAY | RGBA2GRAY | BGRA2GRAY), src/Tracking.cc:172-197,211-223,243-255 — is fused into the
 * upload: only the raw camera frame crosses PCIe.  rgb_order = Tracking::mbRGB (1: R first, 0: B first).
 * `stride` arguments are then bytes per row of the interleaved image. */
BORB_API borb_status borb_extractor_set_input_format(borb_extractor* e, int channels, int rgb_order);

/* Stereo rectification fused into the upload: cv::remap(imLeft, imLeftRect, M1l, M2l, cv::INTER_LINEAR) /
 * (imRight, ..., M1r, M2r, ...) of Examples/Stereo/stereo_euroc.cc:136-137 with the CV_32FC1 maps the caller built once
 * with cv::initUndistortRectifyMap (:96-98).  which = 0: mono / left, 1: right (install 0 first; in the stereo entry
 * points even images use set 0 and odd images set 1).  map_x / map_y: dst_h x dst_w floats; NULL removes the set(s).
 * Afterwards the extract / stereo entry points take RAW src_w x src_h CV_8UC1 frames and work on dst_w x dst_h. */
BORB_API borb_status borb_extractor_set_rectify_maps(borb_extractor* e, int which, const float* map_x, const float* map_y, int src_w,
                                                     int src_h, int dst_w, int dst_h);

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

/* ---------------------------------------------------------------- matchers ------------------- */
/* ORB_SLAM2::ORBmatcher is a stateless stack object created at every call site on three threads
 * (include/ORBmatcher.h:37-102; src/Tracking.cc:599,764,869,1184,1357,1396, src/LocalMapping.cc:215,483,
 * src/LoopClosing.cc:239,589).  A borb_matcher handle owns a CUDA stream and scratch: keep one per thread.
 * The pointer graphs the reference walks (Frame, KeyFrame, MapPoint) are snapshotted by the adapter into the
 * plain views below on the calling thread; results come back as indices. */
typedef struct borb_matcher borb_matcher;
/* Size envelope of every matcher / database call (the reference has no limits; these return BORB_ERR_INVALID_ARG instead of
 * failing inside a launch): at most BORB_MATCH_MAX_FEATURES features per frame or keyframe, and at most that many VALID query
 * points per call (MapPoints of SearchByProjection, last-frame features, world points).  The feature grid sort, the claim
 * bitsets and the resolve kernel's per-query lists live in shared memory, which is what bounds them.  borb_search_local_points
 * compacts the valid points before the check, so the limit applies to the points that reach Frame::isInFrustum, not to the
 * length of Tracking::mvpLocalMapPoints (a KITTI-scale local map of 10^4+ points with a few hundred candidates is fine). */
#define BORB_MATCH_MAX_FEATURES 8192
BORB_API borb_status borb_matcher_create(int device, borb_matcher** out);
BORB_API borb_status borb_matcher_destroy(borb_matcher* m);

/* Frame snapshot for SearchByProjection (include/Frame.h): undistorted keypoints, descriptors, stereo
 * coordinate, image bounds of the 64x48 feature grid (mnMinX.. / src/Frame.cc:97-102), scale factors. */
typedef struct borb_frame borb_frame;   /* device-resident Frame, see borb_frame_create below */
typedef struct borb_frame_view {
    int32_t n;                     /* N */
    const borb_keypoint* keys_un;  /* mvKeysUn */
    const uint8_t* desc;           /* mDescriptors, N x 32 */
    const float* u_right;          /* mvuRight (NULL: monocular, no stereo check) */
    const uint8_t* occupied;       /* 1 if mvpMapPoints[i] && Observations()>0 (src/ORBmatcher.cc:87-89); NULL: none */
    float min_x, min_y, max_x, max_y;
    int32_t n_levels;
    const float* scale_factors;    /* mvScaleFactors */
    const borb_frame* resident;    /* NULL, or a device-resident copy of this frame: then n, keys_un, desc, u_right, the bounds and
                                      scale_factors above are ignored (taken from the resident frame, whose feature grid is already
                                      built) and only `occupied` is read from the host */
} borb_frame_view;

/* Device-resident Frame: uploads the view once (keypoints, descriptors, mvuRight, scale factors) and builds the 64x48 feature
 * grid of Frame::AssignFeaturesToGrid (src/Frame.cc:230-245) ONCE, so that the matcher calls of one Track() — SearchLocalPoints
 * (src/Tracking.cc:1148-1194), SearchByProjection(CurrentFrame, LastFrame) (:867-898), relocalisation — stop re-uploading
 * ~100 KB and re-sorting the grid per call: put the handle into borb_frame_view::resident.  The frame may be used by any
 * matcher on the same device (creation records an event the users wait on).  Destroyed frames are recycled. */
BORB_API borb_status borb_frame_create(borb_matcher* m, const borb_frame_view* view, borb_frame** out);
BORB_API borb_status borb_frame_destroy(borb_frame* f);

/* Camera.* of the settings file as Tracking builds mK / mDistCoef / mbf (src/Tracking.cc:54-90). */
typedef struct borb_camera {
    float fx, fy, cx, cy;
    float k1, k2, p1, p2, k3;      /* k1 == 0 => mvKeysUn = mvKeys (src/Frame.cc:406-410) */
    float bf;                      /* Camera.bf */
} borb_camera;

/* The tail of the Frame constructors (src/Frame.cc:61-117 stereo, :119-178 RGB-D, :180-233 monocular) ON THE DEVICE for images
 * of the LAST batch of extractor `e` — keypoints and descriptors go from the extractor's workspace into resident frames without
 * crossing PCIe:  Frame::UndistortKeyPoints (:404-434, cv::undistortPoints restated: double arithmetic, 5 iterations),
 * Frame::ComputeStereoFromRGBD (:643-664) and Frame::AssignFeaturesToGrid (:230-245).
 *   images[i], n_keys[i]   image index inside the batch and its keypoint count (n_out of the extract call)
 *   mode 0 monocular; 1 stereo: mvuRight / mvDepth are the association borb_stereo_frames computed for pair images[i]/2
 *        (images[i] must be the LEFT image, an even index); 2 RGB-D: depth[i] = host depth map of that frame, registered to
 *        the image (same width x height), depth_type 0 = CV_32F metres, 1 = CV_16U raw with depth_factor = mDepthMapFactor
 *        (+4: depth[i] are DEVICE pointers to tightly packed maps, no copy)
 *        (the convertTo of Tracking::GrabImageRGBD, src/Tracking.cc:227-228, is fused into the lookup)
 *   keys_un / u_right / depth_out   optional host copies (mvKeysUn, mvuRight, mvDepth), n_frames x cap entries
 *   bounds4   mnMinX, mnMinY, mnMaxX, mnMaxY (Frame::ComputeImageBounds, :436-464)
 *   frames    n_frames resident frames; use them through borb_frame_view::resident, release with borb_frame_destroy. */
BORB_API borb_status borb_frames_from_extractor(borb_matcher* m, borb_extractor* e, const int32_t* images, int n_frames,
                                                const int32_t* n_keys, const borb_camera* cam, int mode, const void* const* depth,
                                                int depth_type, float depth_factor, int depth_stride_bytes, borb_keypoint* keys_un,
                                                float* u_right, float* depth_out, int cap, float* bounds4, borb_frame** frames);
BORB_API borb_status borb_frame_info(const borb_frame* f, int32_t* n, int32_t* n_levels, int32_t* has_u_right);

/* Local map points that passed Frame::isInFrustum (src/Frame.cc:269-325), in vpMapPoints order. */
typedef struct borb_mappoint_view {
    int32_t n;
    const float* proj_x;           /* mTrackProjX */
    const float* proj_y;           /* mTrackProjY */
    const float* proj_xr;          /* mTrackProjXR */
    const int32_t* level;          /* mnTrackScaleLevel */
    const float* view_cos;         /* mTrackViewCos */
    const uint8_t* desc;           /* GetDescriptor(), n x 32 */
    const uint8_t* valid;          /* mbTrackInView && !isBad() (NULL: all valid) */
    const uint8_t* has_obs;        /* Observations()>0 (NULL: all) */
} borb_mappoint_view;

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) — src/ORBmatcher.cc:45-129.
 * match_feat[i] = index of the frame feature that received map point i (F.mvpMapPoints[idx]=pMP), or -1. */
BORB_API borb_status borb_search_by_projection(borb_matcher* m, const borb_frame_view* frame, const borb_mappoint_view* mps,
                                               float th, float nnratio, int32_t* match_feat, int32_t* n_matches);
/* The same search for n_jobs INDEPENDENT (frame, MapPoint list) pairs — e.g. the current frames of n camera streams — in one
 * launch pair and one synchronisation: frames[j] (must be device-resident: borb_frame_view::resident) is searched with points[j],
 * results go to match_feat[j][0 .. points[j].n) and n_matches[j], each identical to what the single call returns.  A single call
 * is ~6 us of kernels behind ~30 us of launch and synchronisation latency; the batch amortises the latter over the streams, the
 * way borb_extract_batch does for the images (no reference counterpart: the reference tracks one camera on one thread). */
BORB_API borb_status borb_search_by_projection_batch(borb_matcher* m, const borb_frame_view* frames, const borb_mappoint_view* points,
                                                     int n_jobs, float th, float nnratio, int32_t* const* match_feat, int32_t* n_matches);

/* LastFrame snapshot for the motion-model search: per last-frame feature i the keypoint (octave, angle of mvKeysUn), the
 * world position and representative descriptor of its MapPoint, valid[i] = mvpMapPoints[i] && !mvbOutlier[i],
 * has_obs[i] = mvpMapPoints[i]->Observations()>0. */
typedef struct borb_lastframe_view {
    int32_t n;
    const borb_keypoint* keys_un;  /* LastFrame.mvKeysUn (octave == mvKeys[i].octave) */
    const float* world_pos;        /* n x 3, pMP->GetWorldPos() */
    const uint8_t* desc;           /* n x 32, pMP->GetDescriptor() */
    const uint8_t* valid;          /* NULL: all valid */
    const uint8_t* has_obs;        /* NULL: all */
} borb_lastframe_view;

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) — src/ORBmatcher.cc:1328-1470
 * (Tracking::TrackWithMotionModel, src/Tracking.cc:885,891).  Tcw = CurrentFrame.mTcw rows 0..2 (3x4 row-major);
 * forward/backward = the bForward/bBackward flags of :1348-1349 (a few float ops on the two poses, computed by the caller).
 * state_cur[i2] (cur->n entries): >=0 = index of the LastFrame feature whose MapPoint now sits in
 * CurrentFrame.mvpMapPoints[i2]; -1 = untouched; -2 = set to NULL by the rotation-consistency cull (:1456-1466). */
BORB_API borb_status borb_search_by_projection_last(borb_matcher* m, const borb_frame_view* cur, const borb_lastframe_view* last,
                                                    const float* Tcw, float fx, float fy, float cx, float cy, float bf, float th,
                                                    int forward, int backward, int check_orientation, int32_t* state_cur,
                                                    int32_t* n_matches);

/* A list of MapPoints with their world-frame data, for the two pose-projection searches below. */
typedef struct borb_worldpoints_view {
    int32_t n;
    const float* world_pos;        /* n x 3, pMP->GetWorldPos() */
    const uint8_t* desc;           /* n x 32, pMP->GetDescriptor() */
    const float* max_distance;     /* n, MapPoint::mfMaxDistance (GetMaxDistanceInvariance() = 1.2f * this, src/MapPoint.cc:379-383) */
    const float* min_distance;     /* n, MapPoint::mfMinDistance (GetMinDistanceInvariance() = 0.8f * this, :373-377) */
    const float* normal;           /* n x 3, pMP->GetNormal(); only read by borb_search_by_projection_sim3 */
    const float* angle;            /* n, pKF->mvKeysUn[i].angle of the keyframe feature that observes the point; only read
                                      by borb_search_by_projection_kf with check_orientation */
    const uint8_t* valid;          /* NULL: all valid */
} borb_worldpoints_view;

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
 * — src/ORBmatcher.cc:1472-1599 (Tracking::Relocalization, src/Tracking.cc:1396,1410).
 * pts[i] = pKF->GetMapPointMatches()[i] with valid[i] = pMP && !pMP->isBad() && !sAlreadyFound.count(pMP);
 * cur->occupied[i2] = CurrentFrame.mvpMapPoints[i2] != NULL.  Tcw = CurrentFrame.mTcw rows 0..2, Ow = -Rcw.t()*tcw
 * (:1476-1478, the caller's cv::Mat lines); log_scale_factor = CurrentFrame.mfLogScaleFactor (PredictScale,
 * src/MapPoint.cc:402-417).  state_cur as in borb_search_by_projection_last (index into pts, -1, -2). */
BORB_API borb_status borb_search_by_projection_kf(borb_matcher* m, const borb_frame_view* cur, const borb_worldpoints_view* pts,
                                                  const float* Tcw, const float* Ow, float fx, float fy, float cx, float cy,
                                                  float log_scale_factor, float th, int orb_dist, int check_orientation,
                                                  int32_t* state_cur, int32_t* n_matches);

/* ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th)
 * — src/ORBmatcher.cc:290-403 (LoopClosing::ComputeSim3, src/LoopClosing.cc:391).
 * kf: the keyframe's features; kf->occupied[idx] = vpMatched[idx] != NULL on entry.  Tcw = [Rcw | tcw] after the
 * Sim3 scale has been divided out and Ow = -Rcw.t()*tcw (:298-303, the caller's cv::Mat lines).
 * pts->valid[i] = !vpPoints[i]->isBad() && !spAlreadyFound.count(vpPoints[i]).
 * state_kf[idx] = index into pts of the point now in vpMatched[idx], -1 = vpMatched[idx] untouched. */
BORB_API borb_status borb_search_by_projection_sim3(borb_matcher* m, const borb_frame_view* kf, const borb_worldpoints_view* pts,
                                                    const float* Tcw, const float* Ow, float fx, float fy, float cx, float cy,
                                                    float log_scale_factor, int th, int32_t* state_kf, int32_t* n_matches);

/* The search part of ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th)
 * — src/ORBmatcher.cc:825-970 (LocalMapping::SearchInNeighbors, src/LocalMapping.cc:483-511) — with scw_variant = 0, and of
 * ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, vpPoints, th, vpReplacePoint) — :972-1100 (LoopClosing::SearchAndFuse,
 * src/LoopClosing.cc:589) — with scw_variant = 1 (Tcw = [Rcw|tcw] with the scale divided out, :983-987).
 * best_idx[i] = feature of pKF selected for point i (bestDist <= TH_LOW), else -1; n_found = how many.  The MapPoint
 * bookkeeping that follows in the reference (:947-966 / :1077-1090: Replace, AddObservation, AddMapPoint, vpReplacePoint)
 * stays with the caller, applied in order; it feeds back into the search only through isBad()/IsInKeyFrame() of a
 * point listed twice, which the caller re-tests when it applies best_idx.
 * kf->u_right = pKF->mvuRight (scw_variant 0; NULL = monocular), inv_level_sigma2 = pKF->mvInvLevelSigma2 (scw_variant 0),
 * pts->valid[i] = pMP && !isBad() && !IsInKeyFrame(pKF) (resp. !spAlreadyFound.count(pMP)). kf->occupied is ignored. */
BORB_API borb_status borb_fuse(borb_matcher* m, const borb_frame_view* kf, const float* inv_level_sigma2,
                               const borb_worldpoints_view* pts, const float* Tcw, const float* Ow, float fx, float fy, float cx,
                               float cy, float bf, float log_scale_factor, float th, int scw_variant, int32_t* best_idx,
                               int32_t* n_found);

/* ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) — src/ORBmatcher.cc:1102-1326
 * (LoopClosing::ComputeSim3, src/LoopClosing.cc:375-378).  pts1 / pts2 = GetMapPointMatches() of the two keyframes
 * (one slot per feature: pts->n == kf->n) with valid[i] = pMP && !vbAlreadyMatched[i] && !isBad() (:1130-1141,:1151-1155).
 * T1w / T2w = the keyframe poses (3x4); S12 = [s12*R12 | t12], S21 = [(1/s12)*R12^T | -sR21*t12] as the reference's
 * cv::Mat lines produce them (:1119-1122).  (fx,fy,cx,cy) = pKF1's intrinsics, used for both directions (:1105-1108).
 * match12[i1] = index in KF2 whose MapPoint goes to vpMatches12[i1], or -1; n_found = return value. */
BORB_API borb_status borb_search_by_sim3(borb_matcher* m, const borb_frame_view* kf1, const borb_frame_view* kf2,
                                         const borb_worldpoints_view* pts1, const borb_worldpoints_view* pts2, const float* T1w,
                                         const float* T2w, const float* S12, const float* S21, float fx, float fy, float cx, float cy,
                                         float log_scale_factor1, float log_scale_factor2, float th, int32_t* match12,
                                         int32_t* n_found);

/* Tracking::SearchLocalPoints (src/Tracking.cc:1148-1194) in one call: Frame::isInFrustum (src/Frame.cc:269-325) for every
 * candidate MapPoint, then ORBmatcher::SearchByProjection(F, vpMapPoints, th) (src/ORBmatcher.cc:45-129) on those in view —
 * the projections never leave the device.  pts->valid[i] = the point reaches isInFrustum (:1171-1175: not already matched
 * in this frame, not bad); has_obs[i] = Observations()>0 (NULL: all).  Tcw = [mRcw | mtcw], Ow = mOw, mbf, log_scale_factor =
 * Frame members; viewing_cos_limit = 0.5 at the call site.  in_view[i] = mbTrackInView (the caller runs IncreaseVisible on
 * it); proj_x/proj_y/proj_xr/level/view_cos (each may be NULL) = mTrackProjX/Y/XR, mnTrackScaleLevel, mTrackViewCos
 * (0 where not in view); match_feat / n_matches as in borb_search_by_projection. */
BORB_API borb_status borb_search_local_points(borb_matcher* m, const borb_frame_view* frame, const borb_worldpoints_view* pts,
                                              const uint8_t* has_obs, const float* Tcw, const float* Ow, float fx, float fy, float cx,
                                              float cy, float mbf, float viewing_cos_limit, float log_scale_factor, float th,
                                              float nnratio, uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr,
                                              int32_t* level, float* view_cos, int32_t* match_feat, int32_t* n_matches);

/* ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12,
 * int windowSize) — src/ORBmatcher.cc:405-520 (Tracking::MonocularInitialization, src/Tracking.cc:599).
 * f1/f2: keys_un, desc (and f2's grid bounds) are read; prev_matched = vbPrevMatched as f1->n x 2 floats, updated in
 * place (:513-517); matches12[i1] = index in F2 or -1; n_matches = return value. */
BORB_API borb_status borb_search_for_initialization(borb_matcher* m, const borb_frame_view* f1, const borb_frame_view* f2,
                                                    float* prev_matched, int window_size, float nnratio, int check_orientation,
                                                    int32_t* matches12, int32_t* n_matches);

/* MapPoint::ComputeDistinctiveDescriptors — src/MapPoint.cc:242-307, for n_points MapPoints in one launch.
 * desc: the observing keyframes' descriptors (pKF->mDescriptors.row(idx) of every non-bad observation, in
 * mObservations order), MapPoint p owning rows offsets[p] .. offsets[p+1]-1.  best_idx[p] = row (relative to
 * offsets[p]) of the descriptor with the least median distance to the others, -1 if the point has none. */
BORB_API borb_status borb_distinctive_descriptors(borb_matcher* m, const uint8_t* desc, const int32_t* offsets, int n_points,
                                                  int32_t* best_idx);

/* DBoW2::FeatureVector (ordered map NodeId -> feature indices) as CSR; node_id ascending. */
typedef struct borb_featvec_view {
    int32_t n_nodes;
    const uint32_t* node_id;
    const int32_t* start;          /* n_nodes + 1 */
    const uint32_t* feat_idx;
} borb_featvec_view;

/* KeyFrame (or Frame) snapshot for the BoW-guided searches. */
typedef struct borb_keyframe_view {
    int32_t n;
    const borb_keypoint* keys_un;  /* mvKeysUn (angle, octave, pt) */
    const uint8_t* desc;           /* mDescriptors */
    const uint8_t* has_mp;         /* per feature: MapPoint present && !isBad() (NULL: none) */
    const float* u_right;          /* mvuRight (NULL: all -1) */
    borb_featvec_view fv;          /* mFeatVec */
    int32_t n_levels;
    const float* scale_factors;    /* mvScaleFactors */
    const float* level_sigma2;     /* mvLevelSigma2 */
} borb_keyframe_view;

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) — src/ORBmatcher.cc:159-288, for n_kf keyframes
 * against one frame in one launch (relocalisation / loop candidates).  match[k*frame->n + j] = index of the
 * feature of keyframe k whose MapPoint frame feature j received, or -1; n_matches[k] = return value. */
BORB_API borb_status borb_search_by_bow(borb_matcher* m, const borb_keyframe_view* kfs, int n_kf, const borb_keyframe_view* frame,
                                        float nnratio, int check_orientation, int32_t* match, int32_t* n_matches);
/* ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) — src/ORBmatcher.cc:522-655.
 * match12[i] = index in kf2 of the MapPoint matched to feature i of kf1, or -1. */
BORB_API borb_status borb_search_by_bow_kf(borb_matcher* m, const borb_keyframe_view* kf1, const borb_keyframe_view* kf2,
                                           float nnratio, int check_orientation, int32_t* match12, int32_t* n_matches);
/* ORBmatcher::SearchForTriangulation — src/ORBmatcher.cc:657-823.  F12 row-major 3x3; (ex,ey) = projection of
 * kf1's camera centre into kf2 (:663-670, computed by the caller).  pairs: 2*cap ints (idx1, idx2), ascending idx1. */
BORB_API borb_status borb_search_for_triangulation(borb_matcher* m, const borb_keyframe_view* kf1, const borb_keyframe_view* kf2,
                                                   const float* F12, float ex, float ey, int only_stereo, int check_orientation,
                                                   int32_t* pairs, int cap, int32_t* n_pairs);

/* ---- device-resident keyframe database -------------------------------------------------------------------------
 * KeyFrameDatabase (include/KeyFrameDatabase.h, src/KeyFrameDatabase.cc) re-designed for the GPU: instead of an inverted
 * file walked per query word, every keyframe's BowVector (and the keyframe-side inputs of SearchByBoW: keypoints,
 * descriptors, FeatureVector, MapPoint mask) stays in HBM, and one launch computes for ALL keyframes what the walk and
 * the scoring loop produce: the number of words shared with the query (mnRelocWords / mnLoopWords, :91-108,:211-224) and
 * DBoW2's L1 score (:127,:240; Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-71), bit-identical (terms added in word
 * order).  The covisibility accumulation that follows (:134-190,:251-307) walks the keyframe graph and stays with the
 * caller; first_word gives it the reference's list order (keyframes are met in order of their first shared word,
 * then of insertion into that word's list). */
typedef struct borb_kfdb borb_kfdb;
BORB_API borb_status borb_kfdb_create(int device, borb_kfdb** out);
BORB_API borb_status borb_kfdb_destroy(borb_kfdb* db);
BORB_API borb_status borb_kfdb_clear(borb_kfdb* db);                                   /* KeyFrameDatabase::clear :68-73 */
/* KeyFrameDatabase::add (:41-47).  bow_word ascending (std::map order), bow_value = BowVector weights (double).
 * *slot_out identifies the keyframe from now on (slots are never reused). */
BORB_API borb_status borb_kfdb_add(borb_kfdb* db, const borb_keyframe_view* kf, const uint32_t* bow_word, const double* bow_value,
                                   int n_bow, int32_t* slot_out);
BORB_API borb_status borb_kfdb_erase(borb_kfdb* db, int32_t slot);                     /* KeyFrameDatabase::erase :49-66 */
BORB_API borb_status borb_kfdb_set_has_mp(borb_kfdb* db, int32_t slot, const uint8_t* has_mp);   /* MapPoints culled / added since add() */
BORB_API borb_status borb_kfdb_size(const borb_kfdb* db, int32_t* n_slots, uint64_t* device_bytes);
/* One query BowVector against every keyframe.  Outputs have one entry per slot (erased slots: 0 common words):
 * common_words[s], score[s] = (float)L1 score, first_word[s] = smallest shared word id (0xFFFFFFFF if none). */
BORB_API borb_status borb_kfdb_query(borb_matcher* m, borb_kfdb* db, const uint32_t* bow_word, const double* bow_value, int n_bow,
                                     int32_t* common_words, float* score, uint32_t* first_word, int cap, int32_t* n_slots);
/* borb_search_by_bow with the keyframes taken from the database (only the frame is uploaded). */
BORB_API borb_status borb_search_by_bow_db(borb_matcher* m, borb_kfdb* db, const int32_t* slots, int n_kf,
                                           const borb_keyframe_view* frame, float nnratio, int check_orientation, int32_t* match,
                                           int32_t* n_matches);

/* The same search returning COMPACT results — the form to use against many keyframes (BASELINE configs[4]: one frame
 * against a 2000-keyframe database; the dense matrix above would be 9.7 MB of D2H per query).  slots == NULL searches
 * slots 0..n_kf-1 (n_kf must equal the slot count; erased slots give 0 matches).  n_matches[k] = return value of
 * SearchByBoW for keyframe k.  If pairs != NULL: keyframe k's matches are pairs[pair_offset[k] .. + n_matches[k]), each
 * (frame feature index) | (keyframe feature index) << 16, in FeatureVector order of the frame (node id, then feature index);
 * the blocks of different keyframes are packed in no particular order.  *n_pairs_total = sum of n_matches; more than pairs_cap
 * => BORB_ERR_CAPACITY (counts and offsets are still valid). */
BORB_API borb_status borb_search_by_bow_db_pairs(borb_matcher* m, borb_kfdb* db, const int32_t* slots, int n_kf,
                                                 const borb_keyframe_view* frame, float nnratio, int check_orientation,
                                                 int32_t* n_matches, int32_t* pair_offset, uint32_t* pairs, int pairs_cap,
                                                 int32_t* n_pairs_total);

/* ---------------------------------------------------------------- vocabulary (BoW feeder) ---- */
/* ORBVocabulary = DBoW2::TemplatedVocabulary<FORB> (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h).  The tree lives
 * in HBM as one packed blob (so it can be broadcast over NCCL once and shared by every stream of a GPU). */
typedef struct borb_voc borb_voc;
/* Nodes in id order, node 0 = root; children keep the order in which they appear (loadFromTextFile :1378-1420). */
BORB_API borb_status borb_voc_create(const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight,
                                     int n_nodes, int k, int L, int device, borb_voc** out);
/* "k L scoring weighting" + one "parent isLeaf d0..d31 weight" line per node (ORBvoc.txt, :1338-1424). */
BORB_API borb_status borb_voc_load_text(const char* path, int device, borb_voc** out);
BORB_API borb_status borb_voc_destroy(borb_voc* v);
/* Device address and size of the packed blob (root rank: source of the NCCL broadcast). */
BORB_API borb_status borb_voc_blob(const borb_voc* v, void** d_blob, size_t* bytes);
/* Adopt a packed blob that already sits in this device's memory (receiver side of the broadcast). */
BORB_API borb_status borb_voc_from_blob(void* d_blob, size_t bytes, int device, borb_voc** out);
/* NCCL without torch, for a C++ Tracking host: the vocabulary is parsed by ONE rank (src/System.cc:65 loads the 145 MB
 * text on every process) and broadcast over NVLink into every GPU's HBM.  libnccl.so.2 is resolved with dlopen at the first call
 * (BORB_ERR_UNSUPPORTED if absent); libborb.so itself does not link NCCL.
 *   borb_nccl_unique_id: ncclGetUniqueId (128 bytes) on one rank, handed to the others by the host's own means;
 *   borb_nccl_comm_create / _destroy: ncclCommInitRank / ncclCommDestroy — or pass an ncclComm_t the host already owns;
 *   borb_voc_broadcast: two ncclBroadcast calls (size, blob).  On `root` pass the loaded vocabulary and get it back in *out;
 *   elsewhere pass NULL and receive a vocabulary backed by the received blob.  nccl_comm = the rank's ncclComm_t. */
BORB_API borb_status borb_nccl_unique_id(uint8_t* id128);
BORB_API borb_status borb_nccl_comm_create(const uint8_t* id128, int world_size, int rank, int device, void** nccl_comm);
BORB_API borb_status borb_nccl_comm_destroy(void* nccl_comm);
BORB_API borb_status borb_voc_broadcast(borb_voc* root_voc, void* nccl_comm, int root, int rank, int device, borb_voc** out);

/* TemplatedVocabulary::transform(feature, id, weight, nid, levelsup) for n descriptors (:1218-1259): word id,
 * word weight and the node id at level L-levelsup per feature.  Frame::ComputeBoW (src/Frame.cc:395-402) builds
 * mBowVec / mFeatVec from these on the host side of the adapter. */
BORB_API borb_status borb_bow_transform(borb_voc* v, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight,
                                        int32_t* node);

/* Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:395-402, src/KeyFrame.cc:59-68): mBowVec and mFeatVec of n descriptors in
 * one call — tree descent on the GPU, the ordered-map bookkeeping of TemplatedVocabulary::transform (:1150-1194) on the host.
 * bow_word / bow_value: BowVector in word order (capacity n), *n_bow entries; fv_node / fv_start / fv_idx: FeatureVector as CSR
 * (capacities n, n + 1, n), *n_nodes nodes — the layout borb_featvec_view and borb_kfdb_add take. */
BORB_API borb_status borb_compute_bow(borb_voc* v, const uint8_t* desc, int n, int levelsup, uint32_t* bow_word, double* bow_value,
                                      int32_t* n_bow, uint32_t* fv_node, int32_t* fv_start, uint32_t* fv_idx, int32_t* n_nodes);

/* ---------------------------------------------------------------- introspection -------------- */
/* Device time (CUDA events on the matcher's stream) of the kernels of the last borb_search_by_bow_db* call on this handle. */
BORB_API borb_status borb_matcher_set_timing(borb_matcher* m, int enable);
BORB_API borb_status borb_matcher_last_kernel_ms(borb_matcher* m, float* ms);
BORB_API borb_status borb_matcher_launch_count(const borb_matcher* m, uint64_t* n);
/* Per-stage intermediates of the last batch, for parity tests (tests/ compare each stage with the
 * oracle).  xys: (x, y, score) int32 triples in level pixel coordinates. */
BORB_API borb_status borb_debug_candidates(borb_extractor* e, int image, int level, int32_t* xys, int cap, int* n_out);
BORB_API borb_status borb_debug_selected(borb_extractor* e, int image, int level, int32_t* xys, int cap, int* n_out);
BORB_API borb_status borb_debug_blurred(borb_extractor* e, int image, int level, uint8_t* dst, int* w, int* h);
/* Ablation of fast_kernel for the speed-of-light table in profiles/ (0 = full kernel, the only mode that produces
 * keypoints; 1 = TMA tile load only, 2 = + packed reject pass, 3 = + exact scores without NMS / emit). */
BORB_API borb_status borb_debug_set_fast_mode(borb_extractor* e, int mode);
/* Distance arithmetic of the database SearchByBoW kernel: 2 (default) = three 3:2 compressors + 5 POPC per 256-bit distance,
 * 1 = full carry-save adder tree + 4 POPC, 0 = 8 POPC.  Same results; kept switchable for the measurement in profiles/. */
BORB_API borb_status borb_debug_set_bow_csa(int mode);
/* Work-item size of the same kernel: keyframes per item = target / (bucket width)^2, clamped to [1, 32] (default 2560); a negative
 * target selects the static item-to-warp schedule instead of the atomic work counter. */
BORB_API borb_status borb_debug_set_bow_item_target(int target);
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
