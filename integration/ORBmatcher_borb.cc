// ORBmatcher_borb.cc — DROP-IN REPLACEMENT of the reference's src/ORBmatcher.cc: every ORBmatcher method keeps its signature
// (include/ORBmatcher.h:37-102) and forwards to libborb through include/borb_matcher_adapters.hpp.  What stays on the host
// is the handful of 3x3 cv::Mat products with which the reference prepares a pose before its search loops (camera centre,
// Sim3 scale division, epipole); the loops themselves run as CUDA kernels.  Build it in place of src/ORBmatcher.cc and link
// libborb.so.  tests/test_gpu_adapters.py compiles exactly this file against the oracle's plain-data
// Frame / KeyFrame / MapPoint stand-ins and checks every method against the reference's results on the GPU.
#include "ORBmatcher.h"

#include <climits>
#include <cstdint>

#include <opencv2/core/core.hpp>

#include "borb_matcher_adapters.hpp"

using namespace std;

namespace ORB_SLAM2 {

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

namespace {
// rows 0..2 of a 4x4 (or 3x4) float pose as 12 floats
void pose12(const cv::Mat& T, float* out) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) out[4 * r + c] = T.at<float>(r, c);
}
void pose12(const cv::Mat& R, const cv::Mat& t, float* out) {
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) out[4 * r + c] = R.at<float>(r, c);
        out[4 * r + 3] = t.at<float>(r);
    }
}
void vec3(const cv::Mat& v, float* out) { for (int k = 0; k < 3; k++) out[k] = v.at<float>(k); }
}  // namespace

int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th) {
    return borb::adapt::SearchByProjection(F, vpMapPoints, th, mfNNratio);
}

float ORBmatcher::RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }

bool ORBmatcher::CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF2) {
    // kept for callers outside the matcher (none in the reference); the triangulation search evaluates this test on the device
    const float a = kp1.pt.x * F12.at<float>(0, 0) + kp1.pt.y * F12.at<float>(1, 0) + F12.at<float>(2, 0);
    const float b = kp1.pt.x * F12.at<float>(0, 1) + kp1.pt.y * F12.at<float>(1, 1) + F12.at<float>(2, 1);
    const float c = kp1.pt.x * F12.at<float>(0, 2) + kp1.pt.y * F12.at<float>(1, 2) + F12.at<float>(2, 2);
    const float num = a * kp2.pt.x + b * kp2.pt.y + c, den = a * a + b * b;
    if (den == 0) return false;
    return num * num / den < 3.84 * pKF2->mvLevelSigma2[kp2.octave];
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) {
    return borb::adapt::SearchByBoW(pKF, F, vpMapPointMatches, mfNNratio, mbCheckOrientation);
}

int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th) {
    // divide the Sim3 scale out of the pose; camera centre in the world frame
    const cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    const cv::Mat Rcw = sRcw / scw;
    const cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    const cv::Mat Ow = -Rcw.t() * tcw;
    float T[12], O[3];
    pose12(Rcw, tcw, T); vec3(Ow, O);
    return borb::adapt::SearchByProjectionSim3(pKF, vpPoints, vpMatched, th, T, O);
}

int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12, int windowSize) {
    return borb::adapt::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize, mfNNratio, mbCheckOrientation);
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) {
    return borb::adapt::SearchByBoWKF(pKF1, pKF2, vpMatches12, mfNNratio, mbCheckOrientation);
}

int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vector<pair<size_t, size_t> >& vMatchedPairs,
                                       const bool bOnlyStereo) {
    // epipole: camera centre of pKF1 seen from pKF2
    const cv::Mat Cw = pKF1->GetCameraCenter();
    const cv::Mat C2 = pKF2->GetRotation() * Cw + pKF2->GetTranslation();
    const float invz = 1.0f / C2.at<float>(2);
    const float ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
    const float ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;
    float F[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) F[3 * r + c] = F12.at<float>(r, c);
    return borb::adapt::SearchForTriangulation(pKF1, pKF2, F, ex, ey, vMatchedPairs, bOnlyStereo, mbCheckOrientation);
}

int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th) {
    float T[12], O[3];
    pose12(pKF->GetRotation(), pKF->GetTranslation(), T);
    vec3(pKF->GetCameraCenter(), O);
    return borb::adapt::Fuse(pKF, vpMapPoints, th, T, O);
}

int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint) {
    const cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    const cv::Mat Rcw = sRcw / scw;
    const cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    const cv::Mat Ow = -Rcw.t() * tcw;
    float T[12], O[3];
    pose12(Rcw, tcw, T); vec3(Ow, O);
    return borb::adapt::FuseSim3(pKF, vpPoints, th, vpReplacePoint, T, O);
}

int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th) {
    // similarity transforms between the two cameras
    const cv::Mat sR12 = s12 * R12;
    const cv::Mat sR21 = (1.0 / s12) * R12.t();
    const cv::Mat t21 = -sR21 * t12;
    float T1[12], T2[12], S12[12], S21[12];
    pose12(pKF1->GetRotation(), pKF1->GetTranslation(), T1);
    pose12(pKF2->GetRotation(), pKF2->GetTranslation(), T2);
    pose12(sR12, t12, S12); pose12(sR21, t21, S21);
    return borb::adapt::SearchBySim3(pKF1, pKF2, vpMatches12, T1, T2, S12, S21, th);
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    // is the camera moving forward / backward along its axis by more than the baseline? (stereo / RGB-D only)
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat tlc = Rlw * twc + tlw;
    const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;
    return borb::adapt::SearchByProjectionLast(CurrentFrame, LastFrame, th, bForward, bBackward, mbCheckOrientation);
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;
    float T[12], O[3];
    pose12(CurrentFrame.mTcw, T); vec3(Ow, O);
    return borb::adapt::SearchByProjectionKF(CurrentFrame, pKF, sAlreadyFound, th, ORBdist, T, O, mbCheckOrientation);
}

void ORBmatcher::ComputeThreeMaxima(vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {
    // kept for source compatibility: the rotation-consistency culls run inside the kernels
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

// 256-bit Hamming distance of two descriptor rows: a host utility other reference code calls (MapPoint::ComputeDistinctiveDescriptors,
// src/MapPoint.cc:282; the batched GPU form is borb_distinctive_descriptors)
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    const uint64_t* pa = a.ptr<uint64_t>();
    const uint64_t* pb = b.ptr<uint64_t>();
    int dist = 0;
    for (int i = 0; i < 4; i++) dist += __builtin_popcountll(pa[i] ^ pb[i]);
    return dist;
}

}  // namespace ORB_SLAM2
