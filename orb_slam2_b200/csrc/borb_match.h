// Internal declarations for the matcher / vocabulary part of libborb.  Not part of the C ABI.
#pragma once
#include "borb_internal.h"

// Device-resident Frame (include/borb.h: borb_frame_create / borb_frame_from_extractor): what the windowed searches read of a
// Frame — mvKeysUn, mDescriptors, mvuRight, mvScaleFactors and the 64x48 feature grid (Frame::AssignFeaturesToGrid) — kept in
// HBM across the matcher calls of one Track().
struct borb_frame {
    int device = 0;
    int n = 0, n_levels = 0;
    float min_x = 0, min_y = 0, max_x = 0, max_y = 0;
    uint8_t* block = nullptr;       // one allocation: keys | desc | u_right | depth | scale factors | cell_start | cell_idx
    size_t block_bytes = 0;
    int cap = 0;                    // features the block can hold
    borb_keypoint* keys = nullptr;
    uint8_t* desc = nullptr;
    float* u_right = nullptr;       // null: monocular
    float* depth = nullptr;
    float* ur_store = nullptr;      // storage behind u_right / depth (always allocated)
    float* depth_store = nullptr;
    float* sf = nullptr;
    int* cell_start = nullptr;
    int* cell_idx = nullptr;
    cudaEvent_t ready = nullptr;    // recorded after the last kernel that writes the block
};

extern "C" void borb_voc_adopt_ownership(borb_voc* v);     // internal (borb_nccl.cu): the vocabulary now owns its blob

namespace borb {

constexpr int GRID_COLS = 64, GRID_ROWS = 48;     // FRAME_GRID_COLS / FRAME_GRID_ROWS (include/Frame.h:37-38)
constexpr int GRID_CELLS = GRID_COLS * GRID_ROWS;
constexpr int CAND_UNSORTED = 0x40000000;           // flag in cand_cnt: list longer than the sort capacity, kept in position order
constexpr int CAND_COUNT_MASK = 0x3FFFFFFF;
constexpr int MATCH_MAX_FEATURES = 8192;          // per frame / keyframe (grid sort and claim bitsets live in smem)
static_assert(MATCH_MAX_FEATURES == BORB_MATCH_MAX_FEATURES, "include/borb.h documents this limit");

struct ProjArgs {                 // device pointers
    int n;                        // frame features
    const borb_keypoint* keys;
    const uint8_t* desc;
    const float* u_right;         // may be null
    const uint8_t* occupied;      // may be null
    float minX, minY, invW, invH;
    const float* scale_factors;
    const int* cell_start;
    const int* cell_idx;
    int n_mp;
    const float *proj_x, *proj_y, *proj_xr, *view_cos;
    const int32_t* level;
    const uint8_t* mp_desc;
    const uint8_t* mp_valid;      // may be null
    const uint8_t* mp_has_obs;    // may be null
    float th, nnratio;
    uint32_t* cand;               // n_mp x n
    int* cand_cnt;                // n_mp
    // mode 1 (SearchByProjection(CurrentFrame, LastFrame)): the window is precomputed per query by project_points_kernel
    const float* q_radius;        // null in mode 0
    const int32_t* q_minl;
    const int32_t* q_maxl;
    int mode;                     // 0: local map points (ratio test), 1: last frame (best only, rotation histogram)
    int check_ori;
    const float* q_angle;         // mode 1: mvKeysUn[i].angle of the query
    int th_dist;                  // mode 1: accept bestDist <= th_dist (TH_HIGH / ORBdist / TH_LOW)
    int chi2;                     // 1: Fuse(pKF, vpMapPoints, th) reprojection gates per candidate (:907-931)
    const float* inv_sigma2;      // chi2: mvInvLevelSigma2
    uint8_t* q_valid_out;         // mode 1: validity written by project_points_kernel (aliases mp_valid)
    // batched launches (borb_search_by_projection_batch): where this job's results go
    int32_t* out_match;           // n_mp entries, then the match count
};

struct LastArgs {                 // inputs of project_points_kernel
    int variant;                  // 0: (CurrentFrame, LastFrame) :1328   1: (CurrentFrame, KeyFrame) :1472   2: (KeyFrame, Scw) :290
                                  // 3: Frame::isInFrustum (src/Frame.cc:269-325), feeding SearchByProjection(F, vpMapPoints)
    float view_cos_limit;         // variant 3
    int32_t* level_out;           // variant 3: mnTrackScaleLevel
    float* viewcos_out;           // variant 3: mTrackViewCos
    int n_last;
    const borb_keypoint* last_keys;   // variant 0 only (octave, angle of the last-frame feature)
    const float* q_angle_in;      // variant 1: angle of the observing keyframe feature (may be null)
    const float* max_distance;    // variants 1, 2
    const float* min_distance;
    const float* normal;          // variant 2, n x 3
    float Ow[3];
    float log_scale;
    int n_levels;
    // variant 2 options (Fuse x2, SearchBySim3 directions share its code path)
    int invz_double;              // invz = 1.0/z evaluated in double (:1014, :1164) instead of 1/z in float (:333, :861)
    int use_normal;               // viewing-angle gate PO.dot(Pn) < 0.5*dist
    int chain;                    // second transform T2 applied to the camera-frame point; distance = |point in camera 2| (:1157-1177)
    float T2[12];
    const float* world_pos;       // n_last x 3
    const uint8_t* valid_in;      // may be null
    float T[12];                  // Tcw rows 0..2
    float fx, fy, cx, cy, bf, th;
    float minX, minY, maxX, maxY;
    const float* scale_factors;
    int forward, backward;
    float *proj_x, *proj_y, *proj_xr, *radius, *angle;
    int32_t *minl, *maxl;
    uint8_t* valid_out;
};

struct KfDev {                    // device-side borb_keyframe_view
    int n, nn;
    const borb_keypoint* keys;
    const uint8_t* desc;
    const uint8_t* has_mp;        // may be null
    const float* u_right;         // may be null
    const uint32_t* node;
    const int32_t* start;
    const uint32_t* idx;
    const float* scale_factors;
    const float* level_sigma2;
};

struct BowDev {                   // one keyframe's BowVector in the device-resident database (null / 0 when erased)
    const uint32_t* word;         // ascending
    const double* value;
    int n;
};

struct KfStream {                  // one keyframe of the device-resident database, features permuted into FeatureVector order
    const uint32_t* node;         // nn node ids, ascending
    const int32_t* start;         // nn + 1 row offsets
    const uint2* meta;            // m rows: x = feature index (mFeatVec order: node, then feature index) | good-MapPoint flag << 16
                                  //             | index of the row's node in node[] << 17 (nn <= MATCH_MAX_FEATURES < 2^15),
                                  //         y = bits of mvKeysUn[feature].angle
    const uint8_t* desc;          // m x 32, 16-byte aligned
    int nn, m, n, pad;
    const void* pad2[2];
};

struct BowDbArgs {                // bowdb_match_kernel (k_bowdb.cu)
    const KfStream* table;        // one entry per database slot (nn = 0: erased)
    const int32_t* slots;         // n_kf slots to search, or null = slots 0..n_kf-1
    int n_kf, n_items;            // keyframes to search; work items (frame node x keyframe range, the work list inside frame_block)
    const uint8_t* frame_block;   // packed query frame (FrameBlockHdr + sections), 16-byte aligned, frame_bytes % 16 == 0
    int frame_bytes, frame_in_smem;
    int static_sched;             // 1: item i -> warp i mod (warps), 0: atomic work counter
    float nnratio;
    int check_ori;
    uint32_t* table_out;          // n_kf x m_frame, preset to 0xFFFFFFFF
    int* hist_out;                // n_kf x 32 rotation-histogram counters, preset to 0
    int* work_counter;            // preset to 0
};

struct BowDbFinal {               // bowdb_finalize_kernel
    const uint32_t* table_out;
    const int* hist;              // n_kf x 32 (BowDbArgs::hist_out)
    int n_kf, mf, check_ori;
    const uint16_t* forig;        // frame feature index per frame position
    int32_t* n_matches;           // n_kf
    int32_t* pair_off;            // n_kf (may be null)
    uint32_t* pairs;              // frame feature | keyframe feature << 16 (may be null)
    int pairs_cap;
    int* cursor;                  // preset to 0
    int32_t* dense;               // n_kf x dense_stride preset to -1 (may be null): match[k][frame feature] = keyframe feature
    int dense_stride;
};

struct FrameJob {                 // one frame of borb_frames_from_extractor (k_frame.cu)
    const borb_keypoint* src_keys;   // the extractor's mvKeys of that image
    const uint8_t* src_desc;
    const float* src_ur;             // stereo: the extractor's mvuRight / mvDepth of the pair
    const float* src_depth;
    const void* depth_img;           // RGB-D: depth map in HBM (w x h, tight rows)
    borb_keypoint* keys;             // destination borb_frame fields
    uint8_t* desc;
    float* u_right;
    float* depth;
    int* cell_start;
    int* cell_idx;
    int n;
    float min_x, min_y, inv_w, inv_h;
};

struct TriArgs { float F[9]; float ex, ey; int only_stereo, check_ori; };

struct VocDev {                   // views into the packed blob
    int n_nodes, k, L;
    const uint8_t* desc;          // n_nodes x 32
    const double* weight;
    const int32_t* word_id;       // -1 for inner nodes
    const int32_t* child_start;   // n_nodes + 1
    const int32_t* child_ids;     // n_nodes - 1
};

int launch_grid_sort(const borb_keypoint* keys, int n, float minX, float minY, float invW, float invH, int* cell_start, int* cell_idx,
                     cudaStream_t s);
void launch_candidates(const ProjArgs& A, cudaStream_t s);
int launch_projection_batch(const ProjArgs* d_jobs, int n_jobs, int max_n, int max_n_mp, cudaStream_t s);
void launch_resolve(const ProjArgs& A, bool last, int32_t* out, int32_t* ev_idx, uint8_t* ev_bin, int* n_matches, cudaStream_t s);
int launch_projection(const ProjArgs& A, int32_t* match_feat, int* n_matches, cudaStream_t s);
int launch_projection_last(const LastArgs& L, const ProjArgs& A, int32_t* state_cur, int32_t* hist_idx, uint8_t* hist_bin, int* n_matches,
                           cudaStream_t s);
int launch_initialization(const ProjArgs& A, const borb_keypoint* keys1, int n1, int32_t* match12, int32_t* ev_idx, uint8_t* ev_bin,
                          float* prev, int* n_matches, cudaStream_t s);
int launch_kfdb_score(const BowDev* table, int n_slots, const uint32_t* qword, const double* qvalue, int nq, int32_t* common, float* score,
                      uint32_t* first_word, cudaStream_t s);
int launch_distinctive(const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best_idx, cudaStream_t s);
int launch_frustum_projection(const LastArgs& L, const ProjArgs& A, int32_t* match_feat, int* n_matches, cudaStream_t s);
int launch_projection_argmin(const LastArgs& L, const ProjArgs& A, int32_t* best_idx, int* n_found, cudaStream_t s);
int launch_sim3_agree(const int32_t* match1, const int32_t* match2, int n1, int n2, int32_t* match12, int* n_found, cudaStream_t s);
int launch_bow_match(const KfDev* qs, const KfDev* ts, int n_pairs, int mode, float nnratio, int check_ori, int32_t* match,
                     int out_stride, uint8_t* bins, int32_t* n_matches, int max_t, cudaStream_t s);
void host_image_bounds(int w, int h, const borb_camera& c, float* b4);
int launch_frame_build(const FrameJob* d_jobs, int n_jobs, int max_n, const borb_camera& cam, int mode, int depth_type, float depth_factor, int w,
                       int h, int out_cap, borb_keypoint* keys_out, float* ur_out, float* depth_out, cudaStream_t s);
int launch_bowdb(const BowDbArgs& A, const BowDbFinal& F, int csa, int n_sm, cudaStream_t s);
bool bowdb_frame_fits_smem(int frame_bytes);
int launch_triangulation(const KfDev& q, const KfDev& t, const TriArgs& T, int32_t* vmatch, uint8_t* bins, int32_t* pairs, int cap,
                         int32_t* n_pairs, cudaStream_t s);
int launch_bow_transform(const VocDev& V, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight, int32_t* node,
                         cudaStream_t s);

}  // namespace borb
