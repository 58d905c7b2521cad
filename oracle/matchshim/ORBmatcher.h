// TEST INFRASTRUCTURE ONLY (oracle/).  Stand-in for the reference's include/ORBmatcher.h + Frame.h + KeyFrame.h + MapPoint.h so
// that /root/reference/src/ORBmatcher.cc compiles VERBATIM where it lies (oracle/Makefile target `matchref`): the class
// declaration repeats the reference's method signatures (include/ORBmatcher.h:37-102 — an interface, it has to match), the
// three data classes carry exactly the members ORBmatcher.cc touches, as plain data filled by oracle/matchref_wrap.cpp.
// What is NOT the reference's code here (and therefore not "verbatim-pinned"): GetFeaturesInArea / the 64x48 grid
// (src/Frame.cc:230-245,327-392, src/KeyFrame.cc:569-608), MapPoint::PredictScale (src/MapPoint.cc:385-417) and the
// *DistanceInvariance getters (:373-383) — restated below in a few lines each — and the cv::Mat arithmetic of the cv shim.
#pragma once
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <map>
#include <set>
#include <vector>

#include <opencv2/core/core.hpp>

#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"     // the reference's own header (pure STL), found through -I/root/reference

namespace ORB_SLAM2 {
using std::pair;
using std::vector;

class KeyFrame;
class Frame;

constexpr int FRAME_GRID_ROWS = 48, FRAME_GRID_COLS = 64;     // include/Frame.h:37-38

// 64x48 bucket grid over [minX,maxX) x [minY,maxY) — AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea
struct FeatureGrid {
    float minX = 0, minY = 0, maxX = 1, maxY = 1, invW = 1, invH = 1;
    std::vector<size_t> cell[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    void build(const std::vector<cv::KeyPoint>& keys, float x0, float y0, float x1, float y1) {
        minX = x0; minY = y0; maxX = x1; maxY = y1;
        invW = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(maxX - minX);
        invH = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(maxY - minY);
        for (auto& col : cell) for (auto& c : col) c.clear();
        for (size_t i = 0; i < keys.size(); i++) {
            const int px = (int)std::round((keys[i].pt.x - minX) * invW), py = (int)std::round((keys[i].pt.y - minY) * invH);
            if (px < 0 || px >= FRAME_GRID_COLS || py < 0 || py >= FRAME_GRID_ROWS) continue;
            cell[px][py].push_back(i);
        }
    }
    std::vector<size_t> area(const std::vector<cv::KeyPoint>& keys, float x, float y, float r, int minLevel, int maxLevel) const {
        std::vector<size_t> v;
        const int nMinCellX = std::max(0, (int)std::floor((x - minX - r) * invW));
        if (nMinCellX >= FRAME_GRID_COLS) return v;
        const int nMaxCellX = std::min(FRAME_GRID_COLS - 1, (int)std::ceil((x - minX + r) * invW));
        if (nMaxCellX < 0) return v;
        const int nMinCellY = std::max(0, (int)std::floor((y - minY - r) * invH));
        if (nMinCellY >= FRAME_GRID_ROWS) return v;
        const int nMaxCellY = std::min(FRAME_GRID_ROWS - 1, (int)std::ceil((y - minY + r) * invH));
        if (nMaxCellY < 0) return v;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                for (size_t j : cell[ix][iy]) {
                    const cv::KeyPoint& kp = keys[j];
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    const float distx = kp.pt.x - x, disty = kp.pt.y - y;
                    if (std::fabs(distx) < r && std::fabs(disty) < r) v.push_back(j);
                }
        return v;
    }
};

class MapPoint {
public:
    // data (filled by the wrapper)
    cv::Mat mWorldPos, mNormalVector, mDescriptor;
    float mfMaxDistance = 0, mfMinDistance = 0;
    bool mbBad = false;
    int nObs = 1;
    std::map<KeyFrame*, size_t> mObservations;
    // fields ORBmatcher.cc reads / writes directly (include/MapPoint.h:86-92)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 0;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    // side effects recorded for the wrapper
    int fusedIdx = -1;
    MapPoint* replacedBy = nullptr;

    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    float GetMaxDistance() { return mfMaxDistance; }      // the two getters include/borb_matcher_adapters.hpp asks a maintainer to add
    float GetMinDistance() { return mfMinDistance; }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    void AddObservation(KeyFrame*, size_t idx) { fusedIdx = (int)idx; }   // recorded, not applied (see matchref_wrap.cpp)
    void Replace(MapPoint* pMP) { replacedBy = pMP; }
    template <class F> int PredictScale(const float& currentDist, F* pF) {
        const float ratio = mfMaxDistance / currentDist;
        int nScale = std::ceil(std::log(ratio) / pF->mfLogScaleFactor);
        if (nScale < 0) nScale = 0;
        else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
        return nScale;
    }
};

class Frame {
public:
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    cv::Mat mDescriptors, mTcw;
    std::vector<float> mvuRight, mvScaleFactors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    DBoW2::FeatureVector mFeatVec;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0, mfLogScaleFactor = 1;
    int mnScaleLevels = 0;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    FeatureGrid grid;
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const {
        return grid.area(mvKeysUn, x, y, r, minLevel, maxLevel);
    }
};

class KeyFrame {
public:
    int N = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors, Rcw, tcw, Ow;
    std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<MapPoint*> mvpMapPoints;
    DBoW2::FeatureVector mFeatVec;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mfLogScaleFactor = 1;
    int mnScaleLevels = 0;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    FeatureGrid grid;
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const { return grid.area(mvKeysUn, x, y, r, -1, -1); }
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
    vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints() {
        std::set<MapPoint*> s;
        for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p);
        return s;
    }
    bool fuseMode = false;       // Fuse wrappers: report no MapPoint in any slot, so every accepted point takes the AddObservation
                                 // branch (:961-965 / :1084-1088) and its bestIdx becomes visible to the wrapper
    MapPoint* GetMapPoint(const size_t& idx) { return fuseMode ? nullptr : mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint*, const size_t&) {}
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
};

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<pair<size_t, size_t> >& vMatchedPairs,
                               const bool bOnlyStereo);
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                     const cv::Mat& t12, const float th);
    int Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th = 3.0);
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint);

    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

protected:
    bool CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF);
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2
