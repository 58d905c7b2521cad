// borb_matcher_adapters.hpp — header-only adapters for every ORBmatcher method (reference include/ORBmatcher.h:44-83) and for
// Frame::ComputeBoW: they snapshot the reference's pointer graph (Frame / KeyFrame / MapPoint, whose getters take mutexes —
// src/MapPoint.cc:309-313,373-383) into plain arrays on the calling thread, call the C ABI of borb.h, and write the results back the
// way the reference's own loops do.  Templated on the reference's types so that the header needs nothing but their public
// members; integration/ORBmatcher_borb.cc instantiates them with ORB_SLAM2::Frame / KeyFrame / MapPoint and is the drop-in
// replacement of src/ORBmatcher.cc.  The few cv::Mat lines of pose algebra the reference performs before its loops (camera centre,
// Sim3 decomposition, epipole — a handful of 3x3 products) stay in that file and arrive here as plain float arrays.
//
// Include AFTER <opencv2/core/core.hpp> and borb_adapters.hpp.  tests/test_gpu_adapters.py EXECUTES every adapter on the GPU: the
// oracle's plain-data Frame / KeyFrame / MapPoint stand-ins (oracle/matchshim) are run through integration/ORBmatcher_borb.cc by
// the same C wrappers that drive the verbatim src/ORBmatcher.cc, and the results must equal the reference's.
//
// MapPoint needs two one-line getters next to src/MapPoint.cc:373-383 — float GetMaxDistance() / GetMinDistance() returning
// mfMaxDistance / mfMinDistance (the library applies the 1.2f / 0.8f of the *Invariance getters and evaluates PredictScale itself).
#pragma once
#include <cstdint>
#include <cstring>
#include <mutex>
#include <set>
#include <unordered_map>
#include <utility>
#include <vector>

#include "borb_adapters.hpp"

namespace borb {
namespace adapt {

// One matcher handle (CUDA stream + device scratch) per calling thread: ORBmatcher objects are created on the stack at
// every call site of Tracking / LocalMapping / LoopClosing, i.e. on three different threads.
inline borb_matcher* thread_matcher(int device = 0) {
    static thread_local borb_matcher* m = nullptr;
    if (!m) check(borb_matcher_create(device, &m), "borb_matcher_create");
    return m;
}

// Device-resident copies of Frames (borb_frame: keypoints, descriptors, mvuRight and the 64x48 grid stay in HBM between the
// matcher calls of one Tracking::Track()).  The code that builds a Frame registers the handle it got from
// borb_frames_from_extractor / borb_frame_create under the Frame's address; every adapter below then searches the resident
// copy (borb_frame_view::resident) and only the query side crosses PCIe.  Frame's copy constructor / destructor call
// bind_resident(this, handle_of(other)) / unbind_resident(this) — or, without touching Frame, Tracking does it where it
// assigns mCurrentFrame / mLastFrame.  A binding does not own the handle unless `owned` is set.
struct ResidentRegistry {
    std::mutex mu;
    std::unordered_map<const void*, std::pair<borb_frame*, bool>> map;
};
inline ResidentRegistry& resident_registry() { static ResidentRegistry r; return r; }
inline void unbind_resident(const void* frame) {
    ResidentRegistry& r = resident_registry();
    std::lock_guard<std::mutex> lk(r.mu);
    auto it = r.map.find(frame);
    if (it == r.map.end()) return;
    if (it->second.second) borb_frame_destroy(it->second.first);
    r.map.erase(it);
}
inline void bind_resident(const void* frame, borb_frame* h, bool owned = false) {
    unbind_resident(frame);
    ResidentRegistry& r = resident_registry();
    std::lock_guard<std::mutex> lk(r.mu);
    r.map[frame] = std::make_pair(h, owned);
}
inline const borb_frame* resident_of(const void* frame) {
    ResidentRegistry& r = resident_registry();
    std::lock_guard<std::mutex> lk(r.mu);
    auto it = r.map.find(frame);
    return it == r.map.end() ? nullptr : it->second.first;
}

template <class FrameT>
inline borb_frame_view frame_view(const FrameT& F, const uint8_t* occupied) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(borb_keypoint), "cv::KeyPoint must be the 28-byte POD layout");
    borb_frame_view v;
    v.n = F.N;
    v.keys_un = reinterpret_cast<const borb_keypoint*>(F.mvKeysUn.data());
    v.desc = F.mDescriptors.data;                                  // N x 32, continuous (ORBextractor output)
    v.u_right = F.mvuRight.empty() ? nullptr : F.mvuRight.data();
    v.occupied = occupied;
    v.min_x = F.mnMinX; v.min_y = F.mnMinY; v.max_x = F.mnMaxX; v.max_y = F.mnMaxY;   // static members of Frame in the reference
    v.n_levels = (int32_t)F.mvScaleFactors.size();
    v.scale_factors = F.mvScaleFactors.data();
    v.resident = resident_of(&F);
    return v;
}

// Uploads a Frame's own host members once (for Frames that were not built by borb_frames_from_extractor) and binds the copy.
template <class FrameT>
inline borb_frame* make_resident(const FrameT& F, int device = 0) {
    unbind_resident(&F);
    const borb_frame_view v = frame_view(F, nullptr);
    borb_frame* h = nullptr;
    check(borb_frame_create(thread_matcher(device), &v, &h), "borb_frame_create");
    bind_resident(&F, h, true);
    return h;
}

// DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>) -> CSR, in map order
template <class FeatVec>
struct FlatFeatVec {
    std::vector<uint32_t> node, idx;
    std::vector<int32_t> start;
    explicit FlatFeatVec(const FeatVec& fv) {
        for (const auto& kv : fv) {
            node.push_back((uint32_t)kv.first);
            start.push_back((int32_t)idx.size());
            idx.insert(idx.end(), kv.second.begin(), kv.second.end());
        }
        start.push_back((int32_t)idx.size());
    }
    borb_featvec_view view() const { return borb_featvec_view{(int32_t)node.size(), node.data(), start.data(), idx.data()}; }
};

// A list of MapPoints with world-frame data (GetWorldPos / GetDescriptor / distances / normal), snapshotted once.
template <class MapPointT>
struct PointSnapshot {
    std::vector<float> wp, nrm, maxd, mind, angle;
    std::vector<uint8_t> desc, valid;
    int n = 0;
    explicit PointSnapshot(int n_) : wp((size_t)3 * n_), nrm((size_t)3 * n_), maxd(n_), mind(n_), angle(n_), desc((size_t)32 * n_), valid(n_), n(n_) {}
    void take(int i, MapPointT* p, bool want_normal) {
        const cv::Mat P = p->GetWorldPos();
        for (int k = 0; k < 3; k++) wp[(size_t)3 * i + k] = P.template at<float>(k);
        if (want_normal) { const cv::Mat N = p->GetNormal(); for (int k = 0; k < 3; k++) nrm[(size_t)3 * i + k] = N.template at<float>(k); }
        maxd[i] = p->GetMaxDistance(); mind[i] = p->GetMinDistance();
        const cv::Mat d = p->GetDescriptor();
        std::memcpy(&desc[(size_t)32 * i], d.data, 32);
    }
    borb_worldpoints_view view() const {
        borb_worldpoints_view v = {};
        v.n = n; v.world_pos = wp.data(); v.desc = desc.data(); v.max_distance = maxd.data(); v.min_distance = mind.data();
        v.normal = nrm.data(); v.angle = angle.data(); v.valid = valid.data();
        return v;
    }
};

template <class KeyFrameT>
inline borb_frame_view keyframe_as_frame_view(KeyFrameT* pKF, const uint8_t* occupied) {
    borb_frame_view v;
    v.n = pKF->N;
    v.keys_un = reinterpret_cast<const borb_keypoint*>(pKF->mvKeysUn.data());
    v.desc = pKF->mDescriptors.data;
    v.u_right = pKF->mvuRight.empty() ? nullptr : pKF->mvuRight.data();
    v.occupied = occupied;
    v.min_x = pKF->mnMinX; v.min_y = pKF->mnMinY; v.max_x = pKF->mnMaxX; v.max_y = pKF->mnMaxY;
    v.n_levels = (int32_t)pKF->mvScaleFactors.size();
    v.scale_factors = pKF->mvScaleFactors.data();
    v.resident = nullptr;
    return v;
}

template <class KeyFrameT>
inline borb_keyframe_view keyframe_view(KeyFrameT* pKF, const uint8_t* has_mp, const borb_featvec_view& fv) {
    borb_keyframe_view v = {};
    v.n = pKF->N;
    v.keys_un = reinterpret_cast<const borb_keypoint*>(pKF->mvKeysUn.data());
    v.desc = pKF->mDescriptors.data;
    v.has_mp = has_mp;
    v.u_right = pKF->mvuRight.empty() ? nullptr : pKF->mvuRight.data();
    v.fv = fv;
    v.n_levels = (int32_t)pKF->mvScaleFactors.size();
    v.scale_factors = pKF->mvScaleFactors.data();
    v.level_sigma2 = pKF->mvLevelSigma2.empty() ? nullptr : pKF->mvLevelSigma2.data();
    return v;
}

// ---- ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) — src/ORBmatcher.cc:45-129
template <class FrameT, class MapPointT>
int SearchByProjection(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, float th, float nnratio) {
    const int n = (int)vpMapPoints.size();
    std::vector<float> px(n), py(n), pxr(n), vc(n);
    std::vector<int32_t> lvl(n);
    std::vector<uint8_t> valid(n), obs(n), desc((size_t)32 * n), occ(F.N);
    for (int i = 0; i < n; i++) {
        MapPointT* p = vpMapPoints[i];
        valid[i] = p->mbTrackInView && !p->isBad();                 // :54-58
        px[i] = p->mTrackProjX; py[i] = p->mTrackProjY; pxr[i] = p->mTrackProjXR;
        lvl[i] = p->mnTrackScaleLevel; vc[i] = p->mTrackViewCos;
        obs[i] = p->Observations() > 0;
        const cv::Mat d = p->GetDescriptor();
        std::memcpy(&desc[(size_t)32 * i], d.data, 32);
    }
    for (int i = 0; i < F.N; i++) occ[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;   // :87-89
    const borb_frame_view fv = frame_view(F, occ.data());
    const borb_mappoint_view mv = {n, px.data(), py.data(), pxr.data(), lvl.data(), vc.data(), desc.data(), valid.data(), obs.data()};
    std::vector<int32_t> match(n > 0 ? n : 1);
    int32_t nmatches = 0;
    check(borb_search_by_projection(thread_matcher(), &fv, &mv, th, nnratio, match.data(), &nmatches), "borb_search_by_projection");
    for (int i = 0; i < n; i++)
        if (match[i] >= 0) F.mvpMapPoints[match[i]] = vpMapPoints[i];   // :123, same order => same overwrites
    return nmatches;
}

// ---- ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
// — src/ORBmatcher.cc:1328-1470.  bForward / bBackward (:1338-1349) are computed by the caller (integration/ORBmatcher_borb.cc).
template <class FrameT>
int SearchByProjectionLast(FrameT& CurrentFrame, const FrameT& LastFrame, float th, bool bForward, bool bBackward, bool checkOrientation) {
    const int nl = LastFrame.N;
    std::vector<float> wp((size_t)3 * nl);
    std::vector<uint8_t> desc((size_t)32 * nl), valid(nl), obs(nl), occ(CurrentFrame.N);
    for (int i = 0; i < nl; i++) {
        auto* pMP = LastFrame.mvpMapPoints[i];
        valid[i] = pMP && !LastFrame.mvbOutlier[i];                 // :1353-1357
        if (!valid[i]) continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        for (int k = 0; k < 3; k++) wp[(size_t)3 * i + k] = x3Dw.template at<float>(k);
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&desc[(size_t)32 * i], d.data, 32);
        obs[i] = pMP->Observations() > 0;
    }
    for (int i = 0; i < CurrentFrame.N; i++)
        occ[i] = CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0;        // :1401-1403
    float Tcw[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) Tcw[4 * r + c] = CurrentFrame.mTcw.template at<float>(r, c);
    const borb_frame_view cur = frame_view(CurrentFrame, occ.data());
    const borb_lastframe_view last = {nl, reinterpret_cast<const borb_keypoint*>(LastFrame.mvKeysUn.data()), wp.data(), desc.data(),
                                      valid.data(), obs.data()};
    std::vector<int32_t> state(CurrentFrame.N > 0 ? CurrentFrame.N : 1);
    int32_t nmatches = 0;
    check(borb_search_by_projection_last(thread_matcher(), &cur, &last, Tcw, CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx,
                                         CurrentFrame.cy, CurrentFrame.mbf, th, bForward, bBackward, checkOrientation, state.data(),
                                         &nmatches), "borb_search_by_projection_last");
    for (int i2 = 0; i2 < CurrentFrame.N; i2++) {
        if (state[i2] >= 0) CurrentFrame.mvpMapPoints[i2] = LastFrame.mvpMapPoints[state[i2]];             // :1428
        else if (state[i2] == -2) CurrentFrame.mvpMapPoints[i2] = nullptr;                                  // rotation cull :1456-1466
    }
    return nmatches;
}

// ---- ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
// — src/ORBmatcher.cc:1472-1599 (Tracking::Relocalization).  Tcw = CurrentFrame.mTcw rows 0..2, Ow = -Rcw.t()*tcw (:1476-1478).
template <class FrameT, class KeyFrameT, class MapPointT>
int SearchByProjectionKF(FrameT& CurrentFrame, KeyFrameT* pKF, const std::set<MapPointT*>& sAlreadyFound, float th, int ORBdist,
                         const float* Tcw, const float* Ow, bool checkOrientation) {
    const std::vector<MapPointT*> vpMPs = pKF->GetMapPointMatches();                                       // :1487
    const int n = (int)vpMPs.size();
    PointSnapshot<MapPointT> S(n);
    for (int i = 0; i < n; i++) {
        MapPointT* p = vpMPs[i];
        S.valid[i] = p && !p->isBad() && !sAlreadyFound.count(p);                                          // :1493-1495
        if (!S.valid[i]) continue;
        S.take(i, p, false);
        S.angle[i] = pKF->mvKeysUn[i].angle;                                                               // :1555
    }
    std::vector<uint8_t> occ(CurrentFrame.N);
    for (int i = 0; i < CurrentFrame.N; i++) occ[i] = CurrentFrame.mvpMapPoints[i] != nullptr;            // :1539-1540
    const borb_frame_view cur = frame_view(CurrentFrame, occ.data());
    const borb_worldpoints_view pts = S.view();
    std::vector<int32_t> state(CurrentFrame.N > 0 ? CurrentFrame.N : 1);
    int32_t nmatches = 0;
    check(borb_search_by_projection_kf(thread_matcher(), &cur, &pts, Tcw, Ow, CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx, CurrentFrame.cy,
                                       CurrentFrame.mfLogScaleFactor, th, ORBdist, checkOrientation, state.data(), &nmatches),
          "borb_search_by_projection_kf");
    for (int i2 = 0; i2 < CurrentFrame.N; i2++) {
        if (state[i2] >= 0) CurrentFrame.mvpMapPoints[i2] = vpMPs[state[i2]];                              // :1552
        else if (state[i2] == -2) CurrentFrame.mvpMapPoints[i2] = nullptr;                                  // :1586-1595
    }
    return nmatches;
}

// ---- ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th)
// — src/ORBmatcher.cc:290-403 (LoopClosing::ComputeSim3).  Tcw = [Rcw | tcw] with the Sim3 scale divided out, Ow = -Rcw.t()*tcw (:298-303).
template <class KeyFrameT, class MapPointT>
int SearchByProjectionSim3(KeyFrameT* pKF, const std::vector<MapPointT*>& vpPoints, std::vector<MapPointT*>& vpMatched, int th,
                           const float* Tcw, const float* Ow) {
    std::set<MapPointT*> spAlreadyFound(vpMatched.begin(), vpMatched.end());                               // :306-307
    spAlreadyFound.erase(static_cast<MapPointT*>(nullptr));
    const int n = (int)vpPoints.size();
    PointSnapshot<MapPointT> S(n);
    for (int i = 0; i < n; i++) {
        MapPointT* p = vpPoints[i];
        S.valid[i] = !p->isBad() && !spAlreadyFound.count(p);                                              // :316-318
        if (S.valid[i]) S.take(i, p, true);
    }
    std::vector<uint8_t> occ(pKF->N);
    for (int i = 0; i < pKF->N; i++) occ[i] = vpMatched[i] != nullptr;                                     // :374-375
    const borb_frame_view kf = keyframe_as_frame_view(pKF, occ.data());
    const borb_worldpoints_view pts = S.view();
    std::vector<int32_t> state(pKF->N > 0 ? pKF->N : 1);
    int32_t nmatches = 0;
    check(borb_search_by_projection_sim3(thread_matcher(), &kf, &pts, Tcw, Ow, pKF->fx, pKF->fy, pKF->cx, pKF->cy, pKF->mfLogScaleFactor, th,
                                         state.data(), &nmatches), "borb_search_by_projection_sim3");
    for (int idx = 0; idx < pKF->N; idx++)
        if (state[idx] >= 0) vpMatched[idx] = vpPoints[state[idx]];                                        // :396
    return nmatches;
}

// ---- ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches) — src/ORBmatcher.cc:159-288
template <class KeyFrameT, class FrameT, class MapPointT>
int SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches, float nnratio, bool checkOrientation) {
    const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches();                              // :161
    vpMapPointMatches = std::vector<MapPointT*>(F.N, static_cast<MapPointT*>(nullptr));                  // :163
    const int nk = (int)vpMapPointsKF.size();
    std::vector<uint8_t> has_mp(nk);
    for (int i = 0; i < nk; i++) has_mp[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();              // :196-202
    const FlatFeatVec<decltype(pKF->mFeatVec)> fk(pKF->mFeatVec);
    const FlatFeatVec<decltype(F.mFeatVec)> ff(F.mFeatVec);
    borb_keyframe_view kv = {};
    kv.n = nk; kv.keys_un = reinterpret_cast<const borb_keypoint*>(pKF->mvKeysUn.data()); kv.desc = pKF->mDescriptors.data;
    kv.has_mp = has_mp.data(); kv.fv = fk.view();
    borb_keyframe_view fv = {};
    fv.n = F.N; fv.keys_un = reinterpret_cast<const borb_keypoint*>(F.mvKeysUn.data()); fv.desc = F.mDescriptors.data; fv.fv = ff.view();
    std::vector<int32_t> match(F.N > 0 ? F.N : 1);
    int32_t nmatches = 0;
    check(borb_search_by_bow(thread_matcher(), &kv, 1, &fv, nnratio, checkOrientation, match.data(), &nmatches), "borb_search_by_bow");
    for (int j = 0; j < F.N; j++)
        if (match[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[match[j]];                               // :232
    return nmatches;
}

// ---- ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint *> &vpMatches12) — src/ORBmatcher.cc:522-655
template <class KeyFrameT, class MapPointT>
int SearchByBoWKF(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, float nnratio, bool checkOrientation) {
    const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();   // :524-532
    const int n1 = (int)vpMapPoints1.size(), n2 = (int)vpMapPoints2.size();
    vpMatches12 = std::vector<MapPointT*>(n1, static_cast<MapPointT*>(nullptr));                          // :534
    std::vector<uint8_t> hm1(n1), hm2(n2);
    for (int i = 0; i < n1; i++) hm1[i] = vpMapPoints1[i] && !vpMapPoints1[i]->isBad();                   // :561-565
    for (int i = 0; i < n2; i++) hm2[i] = vpMapPoints2[i] && !vpMapPoints2[i]->isBad();                   // :577-583
    const FlatFeatVec<decltype(pKF1->mFeatVec)> f1(pKF1->mFeatVec), f2(pKF2->mFeatVec);
    const borb_keyframe_view k1 = keyframe_view(pKF1, hm1.data(), f1.view()), k2 = keyframe_view(pKF2, hm2.data(), f2.view());
    std::vector<int32_t> match(n1 > 0 ? n1 : 1);
    int32_t nmatches = 0;
    check(borb_search_by_bow_kf(thread_matcher(), &k1, &k2, nnratio, checkOrientation, match.data(), &nmatches), "borb_search_by_bow_kf");
    for (int i = 0; i < n1; i++)
        if (match[i] >= 0) vpMatches12[i] = vpMapPoints2[match[i]];                                       // :602
    return nmatches;
}

// ---- ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12, int windowSize)
// — src/ORBmatcher.cc:405-520
template <class FrameT>
int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize,
                            float nnratio, bool checkOrientation) {
    const int n1 = (int)F1.mvKeysUn.size();
    vnMatches12 = std::vector<int>(n1, -1);                                                                // :408
    std::vector<float> prev((size_t)2 * (n1 > 0 ? n1 : 1));
    for (int i = 0; i < n1; i++) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    const borb_frame_view v1 = frame_view(F1, nullptr), v2 = frame_view(F2, nullptr);
    std::vector<int32_t> m12(n1 > 0 ? n1 : 1);
    int32_t nmatches = 0;
    check(borb_search_for_initialization(thread_matcher(), &v1, &v2, prev.data(), windowSize, nnratio, checkOrientation, m12.data(), &nmatches),
          "borb_search_for_initialization");
    for (int i = 0; i < n1; i++) {
        vnMatches12[i] = m12[i];
        vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1];                           // :513-517 (updated in place)
    }
    return nmatches;
}

// ---- ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) — src/ORBmatcher.cc:657-823.
// F12 row-major 3x3; (ex, ey) = projection of pKF1's camera centre into pKF2 (:663-670, computed by the caller).
template <class KeyFrameT>
int SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, const float* F12, float ex, float ey,
                           std::vector<std::pair<size_t, size_t> >& vMatchedPairs, bool bOnlyStereo, bool checkOrientation) {
    const int n1 = pKF1->N, n2 = pKF2->N;
    std::vector<uint8_t> hm1(n1), hm2(n2);
    for (int i = 0; i < n1; i++) hm1[i] = pKF1->GetMapPoint(i) != nullptr;                                // :697-703: "if(pMP1) continue"
    for (int i = 0; i < n2; i++) hm2[i] = pKF2->GetMapPoint(i) != nullptr;                                // :722-727
    const FlatFeatVec<decltype(pKF1->mFeatVec)> f1(pKF1->mFeatVec), f2(pKF2->mFeatVec);
    const borb_keyframe_view k1 = keyframe_view(pKF1, hm1.data(), f1.view()), k2 = keyframe_view(pKF2, hm2.data(), f2.view());
    std::vector<int32_t> pairs((size_t)2 * (n1 > 0 ? n1 : 1));
    int32_t np = 0;
    check(borb_search_for_triangulation(thread_matcher(), &k1, &k2, F12, ex, ey, bOnlyStereo, checkOrientation, pairs.data(), n1 > 0 ? n1 : 1, &np),
          "borb_search_for_triangulation");
    vMatchedPairs.clear();
    vMatchedPairs.reserve(np);
    for (int i = 0; i < np; i++) vMatchedPairs.push_back(std::make_pair((size_t)pairs[2 * i], (size_t)pairs[2 * i + 1]));   // :812-820
    return np;
}

// ---- ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) — src/ORBmatcher.cc:1102-1326.  T1w / T2w = the keyframe
// poses (3x4), S12 = [s12*R12 | t12], S21 = [(1/s12)*R12^T | -sR21*t12] as the reference's cv::Mat lines produce them (:1119-1122).
template <class KeyFrameT, class MapPointT>
int SearchBySim3(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const float* T1w, const float* T2w, const float* S12,
                 const float* S21, float th) {
    const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);                          // :1130-1143
    for (int i = 0; i < N1; i++) {
        MapPointT* pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    PointSnapshot<MapPointT> S1(N1), S2(N2);
    for (int i = 0; i < N1; i++) {
        MapPointT* p = vpMapPoints1[i];
        S1.valid[i] = p && !vbAlreadyMatched1[i] && !p->isBad();                                          // :1151-1155
        if (S1.valid[i]) S1.take(i, p, false);
    }
    for (int i = 0; i < N2; i++) {
        MapPointT* p = vpMapPoints2[i];
        S2.valid[i] = p && !vbAlreadyMatched2[i] && !p->isBad();                                          // :1229-1233
        if (S2.valid[i]) S2.take(i, p, false);
    }
    const borb_frame_view k1 = keyframe_as_frame_view(pKF1, nullptr), k2 = keyframe_as_frame_view(pKF2, nullptr);
    const borb_worldpoints_view p1 = S1.view(), p2 = S2.view();
    std::vector<int32_t> match12(N1 > 0 ? N1 : 1);
    int32_t nFound = 0;
    check(borb_search_by_sim3(thread_matcher(), &k1, &k2, &p1, &p2, T1w, T2w, S12, S21, pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy,
                              pKF1->mfLogScaleFactor, pKF2->mfLogScaleFactor, th, match12.data(), &nFound), "borb_search_by_sim3");
    for (int i1 = 0; i1 < N1; i1++)
        if (match12[i1] >= 0) vpMatches12[i1] = vpMapPoints2[match12[i1]];                                // :1316
    return nFound;
}

// ---- ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th) — src/ORBmatcher.cc:825-970.
// The search runs on the GPU for all points at once; the MapPoint bookkeeping (:947-966) is applied here in the reference's order,
// re-testing isBad() / IsInKeyFrame() at the top of every iteration exactly where the reference tests them (:842-849), because an
// earlier iteration's Replace / AddObservation can change them for a point that is listed twice.
template <class KeyFrameT, class MapPointT>
int Fuse(KeyFrameT* pKF, const std::vector<MapPointT*>& vpMapPoints, float th, const float* Tcw, const float* Ow) {
    const int n = (int)vpMapPoints.size();
    PointSnapshot<MapPointT> S(n);
    for (int i = 0; i < n; i++) {
        MapPointT* p = vpMapPoints[i];
        S.valid[i] = p && !p->isBad() && !p->IsInKeyFrame(pKF);                                           // :842-849
        if (S.valid[i]) S.take(i, p, true);
    }
    const borb_frame_view kf = keyframe_as_frame_view(pKF, nullptr);
    const borb_worldpoints_view pts = S.view();
    std::vector<int32_t> best(n > 0 ? n : 1);
    int32_t nFound = 0;
    check(borb_fuse(thread_matcher(), &kf, pKF->mvInvLevelSigma2.data(), &pts, Tcw, Ow, pKF->fx, pKF->fy, pKF->cx, pKF->cy, pKF->mbf,
                    pKF->mfLogScaleFactor, th, 0, best.data(), &nFound), "borb_fuse");
    int nFused = 0;
    for (int i = 0; i < n; i++) {
        MapPointT* pMP = vpMapPoints[i];
        if (!pMP || best[i] < 0) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        MapPointT* pMPinKF = pKF->GetMapPoint(best[i]);                                                    // :947-966
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, best[i]);
            pKF->AddMapPoint(pMP, best[i]);
        }
        nFused++;
    }
    return nFused;
}

// ---- ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, float th, vector<MapPoint *> &vpReplacePoint)
// — src/ORBmatcher.cc:972-1100 (LoopClosing::SearchAndFuse).  Tcw = [Rcw | tcw] with the scale divided out, Ow = -Rcw.t()*tcw (:983-990).
template <class KeyFrameT, class MapPointT>
int FuseSim3(KeyFrameT* pKF, const std::vector<MapPointT*>& vpPoints, float th, std::vector<MapPointT*>& vpReplacePoint, const float* Tcw,
             const float* Ow) {
    const std::set<MapPointT*> spAlreadyFound = pKF->GetMapPoints();                                       // :993
    const int n = (int)vpPoints.size();
    PointSnapshot<MapPointT> S(n);
    for (int i = 0; i < n; i++) {
        MapPointT* p = vpPoints[i];
        S.valid[i] = !p->isBad() && !spAlreadyFound.count(p);                                             // :1005-1007
        if (S.valid[i]) S.take(i, p, true);
    }
    const borb_frame_view kf = keyframe_as_frame_view(pKF, nullptr);
    const borb_worldpoints_view pts = S.view();
    std::vector<int32_t> best(n > 0 ? n : 1);
    int32_t nFound = 0;
    check(borb_fuse(thread_matcher(), &kf, nullptr, &pts, Tcw, Ow, pKF->fx, pKF->fy, pKF->cx, pKF->cy, pKF->mbf, pKF->mfLogScaleFactor, th, 1,
                    best.data(), &nFound), "borb_fuse");
    int nFused = 0;
    for (int i = 0; i < n; i++) {
        MapPointT* pMP = vpPoints[i];
        if (best[i] < 0 || pMP->isBad()) continue;
        MapPointT* pMPinKF = pKF->GetMapPoint(best[i]);                                                    // :1077-1090
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, best[i]);
            pKF->AddMapPoint(pMP, best[i]);
        }
        nFused++;
    }
    return nFused;
}

// ---- Tracking::SearchLocalPoints — src/Tracking.cc:1148-1194: Frame::isInFrustum for every local MapPoint + SearchByProjection,
// one call.  `alreadyMatched(pMP)` = the first loop's bookkeeping (:1151-1168: pMP->mnLastFrameSeen == mCurrentFrame.mnId).
template <class FrameT, class MapPointT, class Pred>
int SearchLocalPoints(FrameT& F, const std::vector<MapPointT*>& vpLocalMapPoints, float th, float nnratio, Pred alreadyMatched) {
    const int n = (int)vpLocalMapPoints.size();
    PointSnapshot<MapPointT> S(n);
    std::vector<uint8_t> obs(n), occ(F.N), in_view(n > 0 ? n : 1);
    for (int i = 0; i < n; i++) {
        MapPointT* p = vpLocalMapPoints[i];
        S.valid[i] = !alreadyMatched(p) && !p->isBad();             // :1171-1175
        if (!S.valid[i]) continue;
        S.take(i, p, true);
        obs[i] = p->Observations() > 0;
    }
    for (int i = 0; i < F.N; i++) occ[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;
    float Tcw[12], Ow[3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) Tcw[4 * r + c] = F.mTcw.template at<float>(r, c);
    const cv::Mat O = F.GetCameraCenter();
    for (int k = 0; k < 3; k++) Ow[k] = O.template at<float>(k);
    const borb_frame_view fv = frame_view(F, occ.data());
    const borb_worldpoints_view pv = S.view();
    std::vector<int32_t> match(n > 0 ? n : 1);
    int32_t nmatches = 0;
    check(borb_search_local_points(thread_matcher(), &fv, &pv, obs.data(), Tcw, Ow, F.fx, F.fy, F.cx, F.cy, F.mbf, 0.5f, F.mfLogScaleFactor,
                                   th, nnratio, in_view.data(), nullptr, nullptr, nullptr, nullptr, nullptr, match.data(), &nmatches),
          "borb_search_local_points");
    for (int i = 0; i < n; i++) {
        if (in_view[i]) vpLocalMapPoints[i]->IncreaseVisible();     // :1177
        if (match[i] >= 0) F.mvpMapPoints[match[i]] = vpLocalMapPoints[i];
    }
    return nmatches;
}

// ---- Frame::ComputeBoW / KeyFrame::ComputeBoW — src/Frame.cc:395-402, src/KeyFrame.cc:59-68: mBowVec / mFeatVec from mDescriptors.
// BowVecT = DBoW2::BowVector (std::map<WordId, WordValue>), FeatVecT = DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>).
template <class BowVecT, class FeatVecT>
void ComputeBoW(borb_voc* voc, const cv::Mat& mDescriptors, BowVecT& mBowVec, FeatVecT& mFeatVec, int levelsup = 4) {
    if (!mBowVec.empty()) return;                                   // :397
    const int n = mDescriptors.rows;
    std::vector<uint32_t> bw(n > 0 ? n : 1), fn(n > 0 ? n : 1), fi(n > 0 ? n : 1);
    std::vector<double> bv(n > 0 ? n : 1);
    std::vector<int32_t> fs(n + 1);
    int32_t nb = 0, nn = 0;
    check(borb_compute_bow(voc, mDescriptors.data, n, levelsup, bw.data(), bv.data(), &nb, fn.data(), fs.data(), fi.data(), &nn), "borb_compute_bow");
    for (int k = 0; k < nb; k++) mBowVec.insert(mBowVec.end(), std::make_pair(bw[k], bv[k]));
    for (int a = 0; a < nn; a++) {
        auto it = mFeatVec.insert(mFeatVec.end(), std::make_pair(fn[a], typename FeatVecT::mapped_type()));
        it->second.assign(fi.begin() + fs[a], fi.begin() + fs[a + 1]);
    }
}

}  // namespace adapt
}  // namespace borb
