// TEST INFRASTRUCTURE ONLY (oracle/).  C entry points around the reference's src/ORBmatcher.cc, which oracle/Makefile (target
// `matchref`) compiles VERBATIM where it lies under /root/reference against oracle/matchshim/ (nothing is copied).  Each
// function builds the Frame / KeyFrame / MapPoint objects of the shim from flat arrays (the same arrays the restatements in
// orb_port_match.cpp and the CUDA library take), calls the reference method, and flattens what it wrote.  Used by
// tests/test_oracle_match_ref.py to pin the restatements — and through them the CUDA kernels — to the reference source.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "ORBmatcher.h"
#include "orb_port.h"
#ifdef BORB_ADAPTER_NO_EXTRACTOR          // the build of the PRODUCT's adapters (oracle/Makefile: libadaptmatch.so), not of the reference
#include <cstdlib>
#include "borb_matcher_adapters.hpp"
#endif

using namespace ORB_SLAM2;

namespace {

std::vector<cv::KeyPoint> make_keys(const orbport_kp* k, int n) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbport_kp), "keypoint layout");
    std::vector<cv::KeyPoint> v(n);
    if (n) std::memcpy(v.data(), k, (size_t)n * sizeof(orbport_kp));
    return v;
}
cv::Mat make_desc(const uint8_t* d, int n) {
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n) std::memcpy(m.data, d, (size_t)n * 32);
    return m;
}
cv::Mat vec3(const float* p) {
    cv::Mat m(3, 1, CV_32F);
    for (int i = 0; i < 3; i++) m.at<float>(i) = p[i];
    return m;
}
cv::Mat mat3(const float* p, int stride) {
    cv::Mat m(3, 3, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) m.at<float>(r, c) = p[r * stride + c];
    return m;
}
cv::Mat pose44(const float* T12) {                    // 3x4 row-major -> 4x4
    cv::Mat m(4, 4, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) m.at<float>(r, c) = T12[4 * r + c];
    m.at<float>(3, 3) = 1.f;
    return m;
}
void fill_featvec(DBoW2::FeatureVector& fv, int nn, const uint32_t* node, const int32_t* start, const uint32_t* idx) {
    for (int a = 0; a < nn; a++) {
        std::vector<unsigned int>& v = fv[node[a]];
        for (int e = start[a]; e < start[a + 1]; e++) v.push_back(idx[e]);
    }
}

struct FrameArgs {
    const orbport_kp* keys; const uint8_t* desc; const float* u_right; int n;
    float minX, minY, maxX, maxY; const float* sf; int n_levels; float log_scale;
};
void build_frame(Frame& F, const FrameArgs& a) {
    F.N = a.n;
    F.mvKeysUn = make_keys(a.keys, a.n);
    F.mvKeys = F.mvKeysUn;
    F.mDescriptors = make_desc(a.desc, a.n);
    F.mvuRight.assign(a.n, -1.f);
    if (a.u_right) F.mvuRight.assign(a.u_right, a.u_right + a.n);
    if (a.sf) F.mvScaleFactors.assign(a.sf, a.sf + a.n_levels);
    F.mnScaleLevels = a.n_levels; F.mfLogScaleFactor = a.log_scale;
    F.mnMinX = a.minX; F.mnMinY = a.minY; F.mnMaxX = a.maxX; F.mnMaxY = a.maxY;
    F.mvpMapPoints.assign(a.n, nullptr);
    F.mvbOutlier.assign(a.n, false);
    F.grid.build(F.mvKeysUn, a.minX, a.minY, a.maxX, a.maxY);
#ifdef BORB_ADAPTER_NO_EXTRACTOR
    // tests/test_gpu_adapters.py runs every fixture a second time with the current frame device-resident (borb_frame)
    if (std::getenv("BORB_ADAPT_RESIDENT") && a.n > 0 && a.sf) borb::adapt::make_resident(F);
    else borb::adapt::unbind_resident(&F);
#endif
}
void build_keyframe(KeyFrame& K, const FrameArgs& a) {
    K.N = a.n;
    K.mvKeysUn = make_keys(a.keys, a.n);
    K.mDescriptors = make_desc(a.desc, a.n);
    K.mvuRight.assign(a.n, -1.f);
    if (a.u_right) K.mvuRight.assign(a.u_right, a.u_right + a.n);
    if (a.sf) K.mvScaleFactors.assign(a.sf, a.sf + a.n_levels);
    K.mnScaleLevels = a.n_levels; K.mfLogScaleFactor = a.log_scale;
    K.mnMinX = a.minX; K.mnMinY = a.minY; K.mnMaxX = a.maxX; K.mnMaxY = a.maxY;
    K.mvpMapPoints.assign(a.n, nullptr);
    K.grid.build(K.mvKeysUn, a.minX, a.minY, a.maxX, a.maxY);
}
// one MapPoint per query with world-frame data
void build_points(std::vector<MapPoint>& pool, int n, const float* wp, const uint8_t* desc, const float* maxd, const float* mind,
                  const float* normal) {
    pool.resize(n);
    for (int i = 0; i < n; i++) {
        if (wp) pool[i].mWorldPos = vec3(wp + 3 * (size_t)i);
        if (normal) pool[i].mNormalVector = vec3(normal + 3 * (size_t)i);
        pool[i].mDescriptor = cv::Mat(1, 32, CV_8U);
        std::memcpy(pool[i].mDescriptor.data, desc + (size_t)i * 32, 32);
        if (maxd) pool[i].mfMaxDistance = maxd[i];
        if (mind) pool[i].mfMinDistance = mind[i];
    }
}
int index_of(const std::vector<MapPoint>& pool, const MapPoint* p) {
    if (!p || pool.empty() || p < &pool[0] || p > &pool[pool.size() - 1]) return -1;
    return (int)(p - &pool[0]);
}

}  // namespace

extern "C" {

int matchref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    cv::Mat ma(1, 32, CV_8U), mb(1, 32, CV_8U);
    std::memcpy(ma.data, a, 32); std::memcpy(mb.data, b, 32);
    return ORBmatcher::DescriptorDistance(ma, mb);
}

// ---- helpers replaying the few cv::Mat lines that precede the loops (so that tests can hand the SAME numbers to both sides)
// ORBmatcher.cc:298-303 / :985-990: decomposition of Scw (3x4 row-major in) -> [Rcw|tcw] and Ow
void matchref_decompose_scw(const float* Scw12, float* T12, float* Ow3) {
    cv::Mat Scw = pose44(Scw12);
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T12[4 * r + c] = Rcw.at<float>(r, c); T12[4 * r + 3] = tcw.at<float>(r); Ow3[r] = Ow.at<float>(r); }
}
// ORBmatcher.cc:1119-1122: sR12, sR21, t21 -> S12 = [sR12|t12], S21 = [sR21|t21]
void matchref_sim3_mats(float s12, const float* R12_9, const float* t12_3, float* S12, float* S21) {
    cv::Mat R12 = mat3(R12_9, 3), t12 = vec3(t12_3);
    cv::Mat sR12 = s12 * R12;
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) { S12[4 * r + c] = sR12.at<float>(r, c); S21[4 * r + c] = sR21.at<float>(r, c); }
        S12[4 * r + 3] = t12.at<float>(r); S21[4 * r + 3] = t21.at<float>(r);
    }
}
// ORBmatcher.cc:1476-1478 (and :1338-1341): Ow = -Rcw.t()*tcw of a 3x4 pose
void matchref_camera_center(const float* T12, float* Ow3) {
    cv::Mat T = pose44(T12);
    const cv::Mat Rcw = T.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = T.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;
    for (int r = 0; r < 3; r++) Ow3[r] = Ow.at<float>(r);
}
// ORBmatcher.cc:663-670: epipole of camera 1 in image 2
void matchref_epipole(const float* Ow1, const float* T2w12, float fx2, float fy2, float cx2, float cy2, float* ex, float* ey) {
    cv::Mat Cw = vec3(Ow1);
    cv::Mat T = pose44(T2w12);
    cv::Mat R2w = T.rowRange(0, 3).colRange(0, 3).clone();
    cv::Mat t2w = T.rowRange(0, 3).col(3).clone();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    *ex = fx2 * C2.at<float>(0) * invz + cx2;
    *ey = fy2 * C2.at<float>(1) * invz + cy2;
}
// ORBmatcher.cc:1338-1349: bForward / bBackward
void matchref_forward_backward(const float* TcwCur12, const float* TcwLast12, float mb, int bMono, int* fwd, int* bwd) {
    cv::Mat Tc = pose44(TcwCur12), Tl = pose44(TcwLast12);
    const cv::Mat Rcw = Tc.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = Tc.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat Rlw = Tl.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tlw = Tl.rowRange(0, 3).col(3);
    const cv::Mat tlc = Rlw * twc + tlw;
    *fwd = tlc.at<float>(2) > mb && !bMono;
    *bwd = -tlc.at<float>(2) > mb && !bMono;
}

// ---- SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) — ORBmatcher.cc:45-129
// owner[idx] = index of the map point sitting in F.mvpMapPoints[idx] afterwards, -3 = the MapPoint that was there before, -1 = none
int matchref_search_by_projection(const orbport_kp* keys_un, const uint8_t* desc, const float* u_right, const uint8_t* occupied, int N,
                                  float minX, float minY, float maxX, float maxY, const float* scale_factors, int n_levels, int n_mp,
                                  const float* proj_x, const float* proj_y, const float* proj_xr, const int32_t* level,
                                  const float* view_cos, const uint8_t* mp_desc, const uint8_t* mp_valid, const uint8_t* mp_has_obs,
                                  float th, float nnratio, int32_t* owner) {
    Frame F;
    build_frame(F, FrameArgs{keys_un, desc, u_right, N, minX, minY, maxX, maxY, scale_factors, n_levels, 1.f});
    MapPoint prior; prior.nObs = 1;
    for (int i = 0; i < N; i++) if (occupied && occupied[i]) F.mvpMapPoints[i] = &prior;
    std::vector<MapPoint> pool;
    build_points(pool, n_mp, nullptr, mp_desc, nullptr, nullptr, nullptr);
    std::vector<MapPoint*> vp(n_mp);
    for (int i = 0; i < n_mp; i++) {
        MapPoint& p = pool[i];
        p.mbTrackInView = !mp_valid || mp_valid[i];
        p.mTrackProjX = proj_x[i]; p.mTrackProjY = proj_y[i]; p.mTrackProjXR = proj_xr[i];
        p.mnTrackScaleLevel = level[i]; p.mTrackViewCos = view_cos[i];
        p.nObs = (!mp_has_obs || mp_has_obs[i]) ? 1 : 0;
        vp[i] = &p;
    }
    ORBmatcher m(nnratio, true);
    const int n = m.SearchByProjection(F, vp, th);
    for (int i = 0; i < N; i++) owner[i] = F.mvpMapPoints[i] == &prior ? -3 : index_of(pool, F.mvpMapPoints[i]);
    return n;
}

// ---- SearchByProjection(CurrentFrame, LastFrame, th, bMono) — :1328-1470.  Same owner[] convention (index into LastFrame).
int matchref_search_by_projection_last(const orbport_kp* cur_keys_un, const uint8_t* cur_desc, const float* cur_u_right,
                                       const uint8_t* cur_occupied, int n_cur, float minX, float minY, float maxX, float maxY,
                                       const float* scale_factors, int n_levels, const orbport_kp* last_keys, const float* world_pos,
                                       const uint8_t* last_desc, const uint8_t* valid, const uint8_t* has_obs, int n_last,
                                       const float* TcwCur, const float* TcwLast, float fx, float fy, float cx, float cy, float bf,
                                       float mb, float th, int bMono, int check_ori, int32_t* owner) {
    Frame Cur, Last;
    build_frame(Cur, FrameArgs{cur_keys_un, cur_desc, cur_u_right, n_cur, minX, minY, maxX, maxY, scale_factors, n_levels, 1.f});
    Cur.mTcw = pose44(TcwCur); Cur.fx = fx; Cur.fy = fy; Cur.cx = cx; Cur.cy = cy; Cur.mbf = bf; Cur.mb = mb;
    MapPoint prior; prior.nObs = 1;
    for (int i = 0; i < n_cur; i++) if (cur_occupied && cur_occupied[i]) Cur.mvpMapPoints[i] = &prior;
    build_frame(Last, FrameArgs{last_keys, last_desc, nullptr, n_last, minX, minY, maxX, maxY, scale_factors, n_levels, 1.f});
    Last.mTcw = pose44(TcwLast);
    std::vector<MapPoint> pool;
    build_points(pool, n_last, world_pos, last_desc, nullptr, nullptr, nullptr);
    for (int i = 0; i < n_last; i++) {
        pool[i].nObs = (!has_obs || has_obs[i]) ? 1 : 0;
        if (!valid || valid[i]) Last.mvpMapPoints[i] = &pool[i];
    }
    ORBmatcher m(0.9f, check_ori != 0);
    const int n = m.SearchByProjection(Cur, Last, th, bMono != 0);
    for (int i = 0; i < n_cur; i++) owner[i] = Cur.mvpMapPoints[i] == &prior ? -3 : index_of(pool, Cur.mvpMapPoints[i]);
    return n;
}

// ---- SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) — :1472-1599
int matchref_search_by_projection_kf(const orbport_kp* cur_keys_un, const uint8_t* cur_desc, const uint8_t* cur_occupied, int n_cur,
                                     float minX, float minY, float maxX, float maxY, const float* scale_factors, int n_levels,
                                     float log_scale_factor, const float* kf_angle, const float* world_pos, const uint8_t* mp_desc,
                                     const float* max_distance, const float* min_distance, const uint8_t* valid, int n_q,
                                     const float* Tcw, float fx, float fy, float cx, float cy, float th, int ORBdist, int check_ori,
                                     int32_t* owner) {
    Frame Cur;
    build_frame(Cur, FrameArgs{cur_keys_un, cur_desc, nullptr, n_cur, minX, minY, maxX, maxY, scale_factors, n_levels, log_scale_factor});
    Cur.mTcw = pose44(Tcw); Cur.fx = fx; Cur.fy = fy; Cur.cx = cx; Cur.cy = cy;
    MapPoint prior;
    for (int i = 0; i < n_cur; i++) if (cur_occupied && cur_occupied[i]) Cur.mvpMapPoints[i] = &prior;
    KeyFrame K;
    std::vector<orbport_kp> kk(n_q > 0 ? n_q : 1);
    std::memset(kk.data(), 0, kk.size() * sizeof(orbport_kp));
    for (int i = 0; i < n_q; i++) kk[i].angle = kf_angle ? kf_angle[i] : 0.f;
    std::vector<uint8_t> zero((size_t)(n_q > 0 ? n_q : 1) * 32, 0);
    build_keyframe(K, FrameArgs{kk.data(), zero.data(), nullptr, n_q, minX, minY, maxX, maxY, scale_factors, n_levels, log_scale_factor});
    std::vector<MapPoint> pool;
    build_points(pool, n_q, world_pos, mp_desc, max_distance, min_distance, nullptr);
    for (int i = 0; i < n_q; i++) if (!valid || valid[i]) K.mvpMapPoints[i] = &pool[i];
    ORBmatcher m(0.9f, check_ori != 0);
    const int n = m.SearchByProjection(Cur, &K, std::set<MapPoint*>(), th, ORBdist);
    for (int i = 0; i < n_cur; i++) owner[i] = Cur.mvpMapPoints[i] == &prior ? -3 : index_of(pool, Cur.mvpMapPoints[i]);
    return n;
}

// ---- SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) — :290-403 (takes Scw itself; see matchref_decompose_scw)
int matchref_search_by_projection_sim3(const orbport_kp* kf_keys_un, const uint8_t* kf_desc, const uint8_t* kf_matched, int n_kf,
                                       float minX, float minY, float maxX, float maxY, const float* scale_factors, int n_levels,
                                       float log_scale_factor, const float* world_pos, const uint8_t* mp_desc,
                                       const float* max_distance, const float* min_distance, const float* normal,
                                       const uint8_t* valid, int n_q, const float* Scw12, float fx, float fy, float cx, float cy,
                                       int th, int32_t* owner) {
    KeyFrame K;
    build_keyframe(K, FrameArgs{kf_keys_un, kf_desc, nullptr, n_kf, minX, minY, maxX, maxY, scale_factors, n_levels, log_scale_factor});
    K.fx = fx; K.fy = fy; K.cx = cx; K.cy = cy;
    MapPoint prior;
    std::vector<MapPoint*> vpMatched(n_kf, nullptr);
    for (int i = 0; i < n_kf; i++) if (kf_matched && kf_matched[i]) vpMatched[i] = &prior;
    std::vector<MapPoint> pool;
    build_points(pool, n_q, world_pos, mp_desc, max_distance, min_distance, normal);
    std::vector<MapPoint*> vp(n_q);
    for (int i = 0; i < n_q; i++) { pool[i].mbBad = valid && !valid[i]; vp[i] = &pool[i]; }
    ORBmatcher m(0.75f, true);
    const int n = m.SearchByProjection(&K, pose44(Scw12), vp, vpMatched, th);
    for (int i = 0; i < n_kf; i++) owner[i] = vpMatched[i] == &prior ? -3 : index_of(pool, vpMatched[i]);
    return n;
}

// ---- SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) — :159-288
int matchref_search_by_bow_kf_f(const orbport_kp* kf_keys, const uint8_t* kf_desc, const uint8_t* kf_has_mp, int kf_n, int kf_nn,
                                const uint32_t* kf_node, const int32_t* kf_start, const uint32_t* kf_idx, const orbport_kp* f_keys,
                                const uint8_t* f_desc, int f_n, int f_nn, const uint32_t* f_node, const int32_t* f_start,
                                const uint32_t* f_idx, float nnratio, int check_ori, int32_t* match_f) {
    KeyFrame K; Frame F;
    build_keyframe(K, FrameArgs{kf_keys, kf_desc, nullptr, kf_n, 0, 0, 1, 1, nullptr, 0, 1.f});
    build_frame(F, FrameArgs{f_keys, f_desc, nullptr, f_n, 0, 0, 1, 1, nullptr, 0, 1.f});
    fill_featvec(K.mFeatVec, kf_nn, kf_node, kf_start, kf_idx);
    fill_featvec(F.mFeatVec, f_nn, f_node, f_start, f_idx);
    std::vector<MapPoint> pool(kf_n);
    for (int i = 0; i < kf_n; i++) if (kf_has_mp && kf_has_mp[i]) K.mvpMapPoints[i] = &pool[i];
    std::vector<MapPoint*> out;
    ORBmatcher m(nnratio, check_ori != 0);
    const int n = m.SearchByBoW(&K, F, out);
    for (int j = 0; j < f_n; j++) match_f[j] = index_of(pool, out[j]);
    return n;
}

// ---- SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) — :522-655
int matchref_search_by_bow_kf_kf(const orbport_kp* k1, const uint8_t* d1, const uint8_t* has_mp1, int n1, int nn1, const uint32_t* node1,
                                 const int32_t* start1, const uint32_t* idx1, const orbport_kp* k2, const uint8_t* d2,
                                 const uint8_t* has_mp2, int n2, int nn2, const uint32_t* node2, const int32_t* start2,
                                 const uint32_t* idx2, float nnratio, int check_ori, int32_t* match12) {
    KeyFrame K1, K2;
    build_keyframe(K1, FrameArgs{k1, d1, nullptr, n1, 0, 0, 1, 1, nullptr, 0, 1.f});
    build_keyframe(K2, FrameArgs{k2, d2, nullptr, n2, 0, 0, 1, 1, nullptr, 0, 1.f});
    fill_featvec(K1.mFeatVec, nn1, node1, start1, idx1);
    fill_featvec(K2.mFeatVec, nn2, node2, start2, idx2);
    std::vector<MapPoint> p1(n1), p2(n2);
    for (int i = 0; i < n1; i++) if (has_mp1 && has_mp1[i]) K1.mvpMapPoints[i] = &p1[i];
    for (int i = 0; i < n2; i++) if (has_mp2 && has_mp2[i]) K2.mvpMapPoints[i] = &p2[i];
    std::vector<MapPoint*> out;
    ORBmatcher m(nnratio, check_ori != 0);
    const int n = m.SearchByBoW(&K1, &K2, out);
    for (int i = 0; i < n1; i++) match12[i] = index_of(p2, out[i]);
    return n;
}

// ---- SearchForTriangulation — :657-823 (computes the epipole itself from Ow1 and KF2's pose; see matchref_epipole)
int matchref_search_for_triangulation(const orbport_kp* k1, const uint8_t* d1, const uint8_t* has_mp1, const float* ur1, int n1, int nn1,
                                      const uint32_t* node1, const int32_t* start1, const uint32_t* idx1, const orbport_kp* k2,
                                      const uint8_t* d2, const uint8_t* has_mp2, const float* ur2, int n2, int nn2,
                                      const uint32_t* node2, const int32_t* start2, const uint32_t* idx2, const float* F12,
                                      const float* Ow1, const float* T2w, float fx2, float fy2, float cx2, float cy2, const float* sf2,
                                      const float* sigma2_2, int n_levels, int only_stereo, int check_ori, int32_t* pairs) {
    KeyFrame K1, K2;
    build_keyframe(K1, FrameArgs{k1, d1, ur1, n1, 0, 0, 1, 1, nullptr, 0, 1.f});
    build_keyframe(K2, FrameArgs{k2, d2, ur2, n2, 0, 0, 1, 1, sf2, n_levels, 1.f});
    K2.mvLevelSigma2.assign(sigma2_2, sigma2_2 + n_levels);
    fill_featvec(K1.mFeatVec, nn1, node1, start1, idx1);
    fill_featvec(K2.mFeatVec, nn2, node2, start2, idx2);
    K1.Ow = vec3(Ow1);
    cv::Mat T = pose44(T2w);
    K2.Rcw = T.rowRange(0, 3).colRange(0, 3).clone(); K2.tcw = T.rowRange(0, 3).col(3).clone();
    K2.fx = fx2; K2.fy = fy2; K2.cx = cx2; K2.cy = cy2;
    std::vector<MapPoint> p1(n1), p2(n2);
    for (int i = 0; i < n1; i++) if (has_mp1 && has_mp1[i]) K1.mvpMapPoints[i] = &p1[i];
    for (int i = 0; i < n2; i++) if (has_mp2 && has_mp2[i]) K2.mvpMapPoints[i] = &p2[i];
    std::vector<pair<size_t, size_t> > v;
    ORBmatcher m(0.6f, check_ori != 0);
    const int n = m.SearchForTriangulation(&K1, &K2, mat3(F12, 3), v, only_stereo != 0);
    for (size_t i = 0; i < v.size(); i++) { pairs[2 * i] = (int32_t)v[i].first; pairs[2 * i + 1] = (int32_t)v[i].second; }
    return n;
}

// ---- SearchForInitialization — :405-520
int matchref_search_for_initialization(const orbport_kp* k1, const uint8_t* d1, int n1, const orbport_kp* k2, const uint8_t* d2, int n2,
                                       float minX, float minY, float maxX, float maxY, float* prev_matched, int windowSize,
                                       float nnratio, int check_ori, int32_t* vnMatches12) {
    Frame F1, F2;
    build_frame(F1, FrameArgs{k1, d1, nullptr, n1, minX, minY, maxX, maxY, nullptr, 0, 1.f});
    build_frame(F2, FrameArgs{k2, d2, nullptr, n2, minX, minY, maxX, maxY, nullptr, 0, 1.f});
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher m(nnratio, check_ori != 0);
    const int n = m.SearchForInitialization(F1, F2, prev, m12, windowSize);
    for (int i = 0; i < n1; i++) { vnMatches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
    return n;
}

// ---- SearchBySim3 — :1102-1326 (takes s12, R12, t12 themselves; see matchref_sim3_mats)
int matchref_search_by_sim3(const orbport_kp* k1, const uint8_t* d1, int n1, const float* bounds1, const float* sf1, float log_scale1,
                            const orbport_kp* k2, const uint8_t* d2, int n2, const float* bounds2, const float* sf2, float log_scale2,
                            int n_levels, const float* wp1, const uint8_t* md1, const float* max1, const float* min1,
                            const uint8_t* valid1, const float* wp2, const uint8_t* md2, const float* max2, const float* min2,
                            const uint8_t* valid2, const float* T1w, const float* T2w, float s12, const float* R12, const float* t12,
                            float fx, float fy, float cx, float cy, float th, int32_t* match12) {
    KeyFrame K1, K2;
    build_keyframe(K1, FrameArgs{k1, d1, nullptr, n1, bounds1[0], bounds1[1], bounds1[2], bounds1[3], sf1, n_levels, log_scale1});
    build_keyframe(K2, FrameArgs{k2, d2, nullptr, n2, bounds2[0], bounds2[1], bounds2[2], bounds2[3], sf2, n_levels, log_scale2});
    K1.fx = fx; K1.fy = fy; K1.cx = cx; K1.cy = cy;
    cv::Mat Ta = pose44(T1w), Tb = pose44(T2w);
    K1.Rcw = Ta.rowRange(0, 3).colRange(0, 3).clone(); K1.tcw = Ta.rowRange(0, 3).col(3).clone();
    K2.Rcw = Tb.rowRange(0, 3).colRange(0, 3).clone(); K2.tcw = Tb.rowRange(0, 3).col(3).clone();
    std::vector<MapPoint> p1, p2;
    build_points(p1, n1, wp1, md1, max1, min1, nullptr);
    build_points(p2, n2, wp2, md2, max2, min2, nullptr);
    for (int i = 0; i < n1; i++) if (!valid1 || valid1[i]) K1.mvpMapPoints[i] = &p1[i];
    for (int i = 0; i < n2; i++) if (!valid2 || valid2[i]) K2.mvpMapPoints[i] = &p2[i];
    std::vector<MapPoint*> vpMatches12(n1, nullptr);
    ORBmatcher m(0.75f, true);
    const int n = m.SearchBySim3(&K1, &K2, vpMatches12, s12, mat3(R12, 3), vec3(t12), th);
    for (int i = 0; i < n1; i++) match12[i] = index_of(p2, vpMatches12[i]);
    return n;
}

// ---- Fuse(pKF, vpMapPoints, th) :825-970 (scw_variant 0; Tcw = pose, Ow = camera centre) and
//      Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) :972-1100 (scw_variant 1; Tcw = Scw itself)
int matchref_fuse(const orbport_kp* kf_keys_un, const uint8_t* kf_desc, const float* kf_u_right, const float* inv_level_sigma2, int n_kf,
                  float minX, float minY, float maxX, float maxY, const float* scale_factors, int n_levels, float log_scale_factor,
                  const float* world_pos, const uint8_t* mp_desc, const float* max_distance, const float* min_distance,
                  const float* normal, const uint8_t* valid, int n_q, const float* Tcw, const float* Ow, float fx, float fy, float cx,
                  float cy, float bf, float th, int scw_variant, int32_t* best_idx) {
    KeyFrame K;
    build_keyframe(K, FrameArgs{kf_keys_un, kf_desc, kf_u_right, n_kf, minX, minY, maxX, maxY, scale_factors, n_levels, log_scale_factor});
    K.fx = fx; K.fy = fy; K.cx = cx; K.cy = cy; K.mbf = bf; K.fuseMode = true;
    if (inv_level_sigma2) K.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + n_levels);
    std::vector<MapPoint> pool;
    build_points(pool, n_q, world_pos, mp_desc, max_distance, min_distance, normal);
    std::vector<MapPoint*> vp(n_q);
    for (int i = 0; i < n_q; i++) { pool[i].mbBad = valid && !valid[i]; vp[i] = &pool[i]; }
    ORBmatcher m(0.6f, true);
    int n;
    if (scw_variant) {
        std::vector<MapPoint*> vpReplace(n_q, nullptr);
        n = m.Fuse(&K, pose44(Tcw), vp, th, vpReplace);
    } else {
        cv::Mat T = pose44(Tcw);
        K.Rcw = T.rowRange(0, 3).colRange(0, 3).clone(); K.tcw = T.rowRange(0, 3).col(3).clone(); K.Ow = vec3(Ow);
        n = m.Fuse(&K, vp, th);
    }
    for (int i = 0; i < n_q; i++) best_idx[i] = pool[i].fusedIdx;
    return n;
}

// ---- bench.py's CPU arm for BASELINE configs[4]: the reference keeps its KeyFrames in memory, so the stand-ins are built
// once (outside the timed region) and the timed sweep is nothing but ORBmatcher::SearchByBoW(pKF, F, out) per keyframe.
struct KfHandle { KeyFrame K; std::vector<MapPoint> pool; };
struct FrHandle { Frame F; };
void* matchref_kf_create(const orbport_kp* keys, const uint8_t* desc, const uint8_t* has_mp, int n, int nn, const uint32_t* node,
                         const int32_t* start, const uint32_t* idx) {
    KfHandle* h = new KfHandle();
    build_keyframe(h->K, FrameArgs{keys, desc, nullptr, n, 0, 0, 1, 1, nullptr, 0, 1.f});
    fill_featvec(h->K.mFeatVec, nn, node, start, idx);
    h->pool.resize(n);
    for (int i = 0; i < n; i++) if (has_mp && has_mp[i]) h->K.mvpMapPoints[i] = &h->pool[i];
    return h;
}
void matchref_kf_destroy(void* h) { delete (KfHandle*)h; }
void* matchref_frame_create(const orbport_kp* keys, const uint8_t* desc, int n, int nn, const uint32_t* node, const int32_t* start,
                            const uint32_t* idx) {
    FrHandle* h = new FrHandle();
    build_frame(h->F, FrameArgs{keys, desc, nullptr, n, 0, 0, 1, 1, nullptr, 0, 1.f});
    fill_featvec(h->F.mFeatVec, nn, node, start, idx);
    return h;
}
void matchref_frame_destroy(void* h) { delete (FrHandle*)h; }
// match_f: n_kf x f_n (may be null): index of the keyframe feature matched to frame feature j, or -1
int matchref_search_by_bow_sweep(void* const* kfs, int n_kf, void* frame, float nnratio, int check_ori, int32_t* nmatches, int32_t* match_f) {
    FrHandle* f = (FrHandle*)frame;
    ORBmatcher m(nnratio, check_ori != 0);
    int total = 0;
    for (int k = 0; k < n_kf; k++) {
        KfHandle* h = (KfHandle*)kfs[k];
        std::vector<MapPoint*> out;
        nmatches[k] = m.SearchByBoW(&h->K, f->F, out);
        total += nmatches[k];
        if (match_f) for (int j = 0; j < f->F.N; j++) match_f[(size_t)k * f->F.N + j] = index_of(h->pool, out[j]);
    }
    return total;
}

}  // extern "C"
