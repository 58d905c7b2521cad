// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/orb_prims.h / orb_port.h headers).
// Line-by-line restatement of Frame::ComputeStereoMatches (reference src/Frame.cc:466-640) on POD
// arrays.  The reference reads the member `mb` before it is initialised (:496 vs :114); the intended
// value mb = mbf/fx is an argument here.  Where the reference would index out of
// range (cv::Mat asserts / std::vector UB) this restatement skips the keypoint.
#include <algorithm>
#include <climits>
#include <cmath>
#include <utility>
#include <vector>

#include "orb_port.h"

extern "C" int orbport_stereo(const orbport_kp* kL, const uint8_t* dL, int nL, const orbport_kp* kR, const uint8_t* dR, int nR,
                              const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh, int nlevels,
                              const float* scale, const float* inv_scale, float bf, float b, float* uRight, float* depth,
                              int32_t* best_dist_dbg) {
    const int TH_HIGH = 100, TH_LOW = 50;
    for (int i = 0; i < nL; i++) { uRight[i] = -1.0f; depth[i] = -1.0f; if (best_dist_dbg) best_dist_dbg[i] = -1; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = lh[0];
    std::vector<std::vector<size_t>> vRowIndices(nRows);
    for (int iR = 0; iR < nR; iR++) {
        const float kpY = kR[iR].y;
        const float r = 2.0f * scale[kR[iR].octave];
        const int maxr = (int)std::ceil(kpY + r);
        const int minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);
    }
    const float minZ = b;
    const float minD = 0;
    const float maxD = bf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < nL; iL++) {
        const orbport_kp& kpL = kL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const size_t rowIdx = (size_t)vL;
        if (rowIdx >= (size_t)nRows) continue;
        const std::vector<size_t>& vCandidates = vRowIndices[rowIdx];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const orbport_kp& kpR = kR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orbport_hamming(dL + (size_t)iL * 32, dR + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = kR[bestIdxR].x;
            const float scaleFactor = inv_scale[kpL.octave];
            const float scaleduL = std::round(kpL.x * scaleFactor);
            const float scaledvL = std::round(kpL.y * scaleFactor);
            const float scaleduR0 = std::round(uR0 * scaleFactor);
            const int w = 5;
            const int W = lw[kpL.octave], H = lh[kpL.octave];
            const uint8_t* IL = pyrL[kpL.octave];
            const uint8_t* IR = pyrR[kpL.octave];
            const int cy = (int)scaledvL, cxL = (int)scaleduL, cxR = (int)scaleduR0;
            if (cy - w < 0 || cy + w + 1 > H || cxL - w < 0 || cxL + w + 1 > W) continue;   // cv::Mat range assert
            float ILp[11][11];
            for (int y = 0; y < 11; y++) for (int x = 0; x < 11; x++) ILp[y][x] = (float)IL[(size_t)(cy - w + y) * W + (cxL - w + x)];
            const float cLv = ILp[w][w];
            for (int y = 0; y < 11; y++) for (int x = 0; x < 11; x++) ILp[y][x] = ILp[y][x] - cLv * 1.0f;
            int bestDist2 = INT_MAX;
            int bestincR = 0;
            const int L = 5;
            std::vector<float> vDists(2 * L + 1);
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= W) continue;
            if (cxR - L - w < 0) continue;                                                       // cv::Mat range assert
            for (int incR = -L; incR <= +L; incR++) {
                float IRp[11][11];
                for (int y = 0; y < 11; y++) for (int x = 0; x < 11; x++) IRp[y][x] = (float)IR[(size_t)(cy - w + y) * W + (cxR + incR - w + x)];
                const float cRv = IRp[w][w];
                double acc = 0;   // cv::norm(NORM_L1) on CV_32F accumulates in double
                for (int y = 0; y < 11; y++) for (int x = 0; x < 11; x++) acc += std::fabs((double)(ILp[y][x] - (IRp[y][x] - cRv * 1.0f)));
                const float dist = (float)acc;
                if (dist < bestDist2) { bestDist2 = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scale[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                depth[iL] = bf / disparity;
                uRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDist2, iL));
                if (best_dist_dbg) best_dist_dbg[iL] = bestDist2;
            }
        }
    }
    if (vDistIdx.empty()) return 0;   // reference: UB (vDistIdx[0] on an empty vector, :627)
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        uRight[vDistIdx[i].second] = -1;
        depth[vDistIdx[i].second] = -1;
        kept--;
    }
    return kept;
}
