// borb_matcher_adapters.hpp — header-only adapters for the per-frame ORBmatcher paths: they snapshot the reference's
// pointer graph (Frame / KeyFrame / MapPoint, whose getters take mutexes — src/MapPoint.cc:309-313,373-383) into plain
// arrays on the calling thread, call the C ABI of borb.h, and write the results back the way the reference's own loops do.
// Templated on the reference's types so that the header needs nothing but their public members; in the reference tree
// instantiate with ORB_SLAM2::Frame / KeyFrame / MapPoint, e.g. the whole body of src/ORBmatcher.cc:45-129 becomes
//
//     int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th)
//     { return borb::adapt::SearchByProjection(F, vpMapPoints, th, mfNNratio); }
//
// Include AFTER <opencv2/core/core.hpp> and borb_adapters.hpp.  tests/test_adapters_compile.py instantiates every template
// below with mock types that carry the same member names and links the result against libborb.so.
// Not covered here (same pattern): SearchForTriangulation, SearchBySim3, Fuse, the Sim3 / relocalisation
// projection overloads — their cv::Mat pose algebra (a few 3x3 products) stays in ORBmatcher.cc and is passed in.
#pragma once
#include <cstdint>
#include <cstring>
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

template <class FrameT>
inline borb_frame_view frame_view(const FrameT& F, const uint8_t* occupied) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(borb_keypoint), "cv::KeyPoint must be the 28-byte POD layout");
    borb_frame_view v;
    v.n = F.N;
    v.keys_un = reinterpret_cast<const borb_keypoint*>(F.mvKeysUn.data());
    v.desc = F.mDescriptors.data;                                  // N x 32, continuous (ORBextractor output)
    v.u_right = F.mvuRight.empty() ? nullptr : F.mvuRight.data();
    v.occupied = occupied;
    v.min_x = FrameT::mnMinX; v.min_y = FrameT::mnMinY; v.max_x = FrameT::mnMaxX; v.max_y = FrameT::mnMaxY;
    v.n_levels = (int32_t)F.mvScaleFactors.size();
    v.scale_factors = F.mvScaleFactors.data();
    return v;
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
// — src/ORBmatcher.cc:1328-1470.  The eight cv::Mat lines that derive bForward / bBackward (:1338-1349) stay in ORBmatcher.cc.
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
        for (int k = 0; k < 3; k++) wp[(size_t)3 * i + k] = x3Dw.template at<float>(k, 0);
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

// ---- Tracking::SearchLocalPoints — src/Tracking.cc:1148-1194: Frame::isInFrustum for every local MapPoint + SearchByProjection,
// one call.  `alreadyMatched(pMP)` = the first loop's bookkeeping (:1151-1168: pMP->mnLastFrameSeen == mCurrentFrame.mnId).
template <class FrameT, class MapPointT, class Pred>
int SearchLocalPoints(FrameT& F, const std::vector<MapPointT*>& vpLocalMapPoints, float th, float nnratio, Pred alreadyMatched) {
    const int n = (int)vpLocalMapPoints.size();
    std::vector<float> wp((size_t)3 * n), nrm((size_t)3 * n), maxd(n), mind(n);
    std::vector<uint8_t> desc((size_t)32 * n), valid(n), obs(n), occ(F.N), in_view(n > 0 ? n : 1);
    for (int i = 0; i < n; i++) {
        MapPointT* p = vpLocalMapPoints[i];
        valid[i] = !alreadyMatched(p) && !p->isBad();               // :1171-1175
        if (!valid[i]) continue;
        const cv::Mat P = p->GetWorldPos(), Pn = p->GetNormal();
        for (int k = 0; k < 3; k++) { wp[(size_t)3 * i + k] = P.template at<float>(k, 0); nrm[(size_t)3 * i + k] = Pn.template at<float>(k, 0); }
        maxd[i] = p->GetMaxDistance();                              // mfMaxDistance / mfMinDistance themselves (the library applies the
        mind[i] = p->GetMinDistance();                              // 1.2f / 0.8f of the *Invariance getters): two one-line getters a
                                                                    // maintainer adds next to src/MapPoint.cc:373-383
        obs[i] = p->Observations() > 0;
        const cv::Mat d = p->GetDescriptor();
        std::memcpy(&desc[(size_t)32 * i], d.data, 32);
    }
    for (int i = 0; i < F.N; i++) occ[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;
    float Tcw[12], Ow[3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) Tcw[4 * r + c] = F.mTcw.template at<float>(r, c);
    const cv::Mat O = F.GetCameraCenter();
    for (int k = 0; k < 3; k++) Ow[k] = O.template at<float>(k, 0);
    const borb_frame_view fv = frame_view(F, occ.data());
    borb_worldpoints_view pv = {};
    pv.n = n; pv.world_pos = wp.data(); pv.desc = desc.data(); pv.max_distance = maxd.data(); pv.min_distance = mind.data();
    pv.normal = nrm.data(); pv.valid = valid.data();
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

}  // namespace adapt
}  // namespace borb
