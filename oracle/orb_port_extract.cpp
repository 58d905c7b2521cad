// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/orb_prims.h / orb_port.h headers).
// Restatement of ORBextractor (reference src/ORBextractor.cc), independent of the verbatim build.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "orb_port.h"
#include "orb_prims.h"

using namespace orbprims;

namespace {

const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

constexpr int PATCH_SIZE = 31, HALF_PATCH = 15, EDGE = 19;

struct Cand { int x, y, score; };   // absolute level pixel coordinates

struct Extractor {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> perLevel, umax;
    // last-call state
    std::vector<int> lw, lh;
    std::vector<std::vector<uint8_t>> pyr, blur;
    std::vector<std::vector<Cand>> cands;
    std::vector<int> levelCount;
};

// ORBextractor.cc:410-470
void init_tables(Extractor& e) {
    const int L = e.nlevels;
    e.scale.assign(L, 1.f); e.sigma2.assign(L, 1.f); e.invScale.resize(L); e.invSigma2.resize(L);
    for (int i = 1; i < L; i++) {
        e.scale[i] = (float)(e.scale[i - 1] * e.scaleFactor);   // float * double -> double -> float (:421)
        e.sigma2[i] = e.scale[i] * e.scale[i];
    }
    for (int i = 0; i < L; i++) { e.invScale[i] = 1.0f / e.scale[i]; e.invSigma2[i] = 1.0f / e.sigma2[i]; }
    e.perLevel.resize(L);
    float factor = (float)(1.0f / e.scaleFactor);                // :436
    float nDesired = (float)(e.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)L)));  // :437
    int sum = 0;
    for (int l = 0; l < L - 1; l++) {
        e.perLevel[l] = cv_round(nDesired);
        sum += e.perLevel[l];
        nDesired *= factor;
    }
    e.perLevel[L - 1] = std::max(e.nfeatures - sum, 0);
    e.umax.assign(HALF_PATCH + 1, 0);
    int v, v0, vmax = cv_floor(HALF_PATCH * std::sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(HALF_PATCH * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH * HALF_PATCH;
    for (v = 0; v <= vmax; ++v) e.umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (e.umax[v0] == e.umax[v0 + 1]) ++v0;
        e.umax[v] = v0;
        ++v0;
    }
}

// Whole-level reformulation of the per-cell cv::FAST loop (ORBextractor.cc:784-829).
// score map S (0 where not a corner at tlow), cell-masked strict 3x3 NMS, per-cell ini/min threshold.
void level_candidates(const uint8_t* img, int W, int H, int iniTh, int minTh, std::vector<Cand>& out) {
    out.clear();
    const int minB = EDGE - 3, maxBX = W - EDGE + 3, maxBY = H - EDGE + 3;
    const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
    const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    const int x0 = EDGE, x1 = W - EDGE, y0 = EDGE, y1 = H - EDGE;   // detection domain [x0,x1) x [y0,y1)
    if (x1 <= x0 || y1 <= y0) return;
    iniTh = std::min(std::max(iniTh, 0), 255);
    minTh = std::min(std::max(minTh, 0), 255);
    const int tlow = std::min(iniTh, minTh);
    std::vector<uint8_t> S((size_t)W * H, 0);
    int pixel[25];
    fast_offsets16(pixel, W);
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            const uint8_t* p = img + (size_t)y * W + x;
            const int v = p[0];
            // 9 contiguous ring pixels all > v+t or all < v-t
            bool corner = false;
            for (int pol = 0; pol < 2 && !corner; pol++) {
                int count = 0;
                for (int k = 0; k < 25; k++) {
                    int q = p[pixel[k]];
                    bool hit = pol ? (q > v + tlow) : (q < v - tlow);
                    if (hit) { if (++count > 8) { corner = true; break; } } else count = 0;
                }
            }
            if (corner) S[(size_t)y * W + x] = (uint8_t)fast_corner_score16(p, pixel, tlow);
        }
    // per cell
    for (int ci = 0; ci < nRows; ci++) {
        const int cy0 = y0 + ci * hCell, cy1 = std::min(cy0 + hCell, y1);
        if (cy0 >= cy1) continue;
        for (int cj = 0; cj < nCols; cj++) {
            const int cx0 = x0 + cj * wCell, cx1 = std::min(cx0 + wCell, x1);
            if (cx0 >= cx1) continue;
            std::vector<Cand> kept;
            bool anyIni = false;
            for (int y = cy0; y < cy1; y++)
                for (int x = cx0; x < cx1; x++) {
                    const int s = S[(size_t)y * W + x];
                    if (s < tlow || s == 0) continue;
                    bool ismax = true;
                    for (int dy = -1; dy <= 1 && ismax; dy++)
                        for (int dx = -1; dx <= 1; dx++) {
                            if (!dx && !dy) continue;
                            const int qx = x + dx, qy = y + dy;
                            if (qx < cx0 || qx >= cx1 || qy < cy0 || qy >= cy1) continue;  // outside the cell's FAST domain == 0
                            if (!(s > S[(size_t)qy * W + qx])) { ismax = false; break; }
                        }
                    if (!ismax) continue;
                    kept.push_back(Cand{x, y, s});
                    if (s >= iniTh) anyIni = true;
                }
            const int t = anyIni ? iniTh : minTh;
            for (const Cand& c : kept) if (c.score >= t) out.push_back(c);
        }
    }
}

// DistributeOctTree (ORBextractor.cc:539-763) + DivideNode (:481-537), canonical tie-break.
// Coordinates of pts are RELATIVE to (minBorderX,minBorderY) like the reference's vToDistributeKeys.
struct QNode { int x0, x1, y0, y1; std::vector<int> pts; };

void divide(const QNode& n, const std::vector<Cand>& P, QNode c[4]) {
    const int halfX = (int)std::ceil((float)(n.x1 - n.x0) / 2), halfY = (int)std::ceil((float)(n.y1 - n.y0) / 2);
    const int mx = n.x0 + halfX, my = n.y0 + halfY;
    c[0] = QNode{n.x0, mx, n.y0, my, {}};
    c[1] = QNode{mx, n.x1, n.y0, my, {}};
    c[2] = QNode{n.x0, mx, my, n.y1, {}};
    c[3] = QNode{mx, n.x1, my, n.y1, {}};
    for (int i : n.pts) {
        const Cand& p = P[i];
        int q = (p.x < mx) ? ((p.y < my) ? 0 : 2) : ((p.y < my) ? 1 : 3);
        c[q].pts.push_back(i);
    }
}

std::vector<int> distribute(const std::vector<Cand>& P, int width, int height, int N) {
    std::vector<int> result;
    const int nIni = (int)std::round((float)width / height);
    if (nIni <= 0 || P.empty()) return result;   // reference: UB on nIni==0; empty input -> empty output
    const float hX = (float)width / nIni;
    std::vector<QNode> list(nIni);
    for (int i = 0; i < nIni; i++) list[i] = QNode{(int)(hX * (float)i), (int)(hX * (float)(i + 1)), 0, height, {}};
    for (int i = 0; i < (int)P.size(); i++) list[(size_t)((float)P[i].x / hX)].pts.push_back(i);
    {
        std::vector<QNode> tmp;
        for (auto& n : list) if (!n.pts.empty()) tmp.push_back(std::move(n));
        list.swap(tmp);
    }
    // One "commit": divide the nodes listed in `order` (indices into list, processing order), stopping
    // after the divide that makes size >= stopAt (stopAt<0: never).  New list = reversed(created) ++ undivided.
    auto commit = [&](const std::vector<int>& order, int stopAt, int& nToExpand) {
        std::vector<QNode> created;
        std::vector<char> divided(list.size(), 0);
        int size = (int)list.size();
        nToExpand = 0;
        for (int idx : order) {
            QNode c[4];
            divide(list[idx], P, c);
            divided[idx] = 1;
            size -= 1;
            for (int q = 0; q < 4; q++)
                if (!c[q].pts.empty()) {
                    if (c[q].pts.size() > 1) nToExpand++;
                    created.push_back(std::move(c[q]));
                    size++;
                }
            if (stopAt >= 0 && size >= stopAt) break;
        }
        std::vector<QNode> nl;
        nl.reserve(size);
        for (int i = (int)created.size() - 1; i >= 0; i--) nl.push_back(std::move(created[i]));
        for (size_t i = 0; i < list.size(); i++) if (!divided[i]) nl.push_back(std::move(list[i]));
        list.swap(nl);
    };
    bool finish = false;
    while (!finish) {
        const int prevSize = (int)list.size();
        std::vector<int> order;
        for (int i = 0; i < (int)list.size(); i++) if (list[i].pts.size() > 1) order.push_back(i);
        int nToExpand = 0;
        commit(order, -1, nToExpand);
        if ((int)list.size() >= N || (int)list.size() == prevSize) finish = true;
        else if ((int)list.size() + nToExpand * 3 > N) {
            while (!finish) {
                const int prev2 = (int)list.size();
                std::vector<int> ord;
                for (int i = 0; i < (int)list.size(); i++) if (list[i].pts.size() > 1) ord.push_back(i);
                // largest first; equal sizes: later-created first == smaller list position first
                std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return list[a].pts.size() > list[b].pts.size(); });
                int dummy;
                commit(ord, N, dummy);
                if ((int)list.size() >= N || (int)list.size() == prev2) finish = true;
            }
        }
    }
    for (const QNode& n : list) {
        int best = n.pts[0];
        for (size_t k = 1; k < n.pts.size(); k++) if (P[n.pts[k]].score > P[best].score) best = n.pts[k];
        result.push_back(best);
    }
    return result;
}

float ic_angle(const uint8_t* img, int W, int x, int y, const std::vector<int>& umax) {
    int m01 = 0, m10 = 0;
    const uint8_t* c = img + (size_t)y * W + x;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m10 += u * c[u];
    for (int v = 1; v <= HALF_PATCH; ++v) {
        int vsum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = c[u + v * W], vm = c[u - v * W];
            vsum += (vp - vm);
            m10 += u * (vp + vm);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

void descriptor(const uint8_t* img, int W, int x, int y, float angleDeg, uint8_t* desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = angleDeg * factorPI;
    const float a = cosf(angle), b = sinf(angle);
    const uint8_t* c = img + (size_t)y * W + x;
    const int8_t* p = kPattern;
    auto val = [&](int i) -> int {
        const float px = (float)p[2 * i], py = (float)p[2 * i + 1];
        const int yy = cv_round(px * b + py * a), xx = cv_round(px * a - py * b);
        return c[yy * W + xx];
    };
    for (int i = 0; i < 32; ++i, p += 32) {
        int v = 0;
        for (int k = 0; k < 8; k++) v |= (val(2 * k) < val(2 * k + 1)) << k;
        desc[i] = (uint8_t)v;
    }
}

}  // namespace

extern "C" {

void* orbport_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    Extractor* e = new Extractor();
    e->nfeatures = nfeatures; e->scaleFactor = scaleFactor; e->nlevels = nlevels; e->iniTh = iniThFAST; e->minTh = minThFAST;
    init_tables(*e);
    return e;
}
void orbport_destroy(void* h) { delete (Extractor*)h; }

void orbport_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* per_level, int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int l = 0; l < e->nlevels; l++) {
        scale[l] = e->scale[l]; inv_scale[l] = e->invScale[l]; sigma2[l] = e->sigma2[l]; inv_sigma2[l] = e->invSigma2[l];
        per_level[l] = e->perLevel[l];
    }
    for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}

int orbport_extract(void* h, const uint8_t* img, int w, int hgt, int stride, orbport_kp* kps, uint8_t* desc, int cap) {
    Extractor* e = (Extractor*)h;
    const int L = e->nlevels;
    e->lw.assign(L, 0); e->lh.assign(L, 0); e->pyr.assign(L, {}); e->blur.assign(L, {}); e->cands.assign(L, {});
    e->levelCount.assign(L, 0);
    if (!img || w <= 0 || hgt <= 0) return 0;
    // pyramid (:1107-1132): sizes from the ORIGINAL image; chain resize from previous level
    for (int l = 0; l < L; l++) {
        const float s = e->invScale[l];
        e->lw[l] = cv_round((float)w * s);
        e->lh[l] = cv_round((float)hgt * s);
        e->pyr[l].resize((size_t)e->lw[l] * e->lh[l]);
        if (l == 0) for (int y = 0; y < hgt; y++) std::memcpy(&e->pyr[0][(size_t)y * w], img + (size_t)y * stride, w);
        else resize_linear_u8(e->pyr[l - 1].data(), e->lw[l - 1], e->lh[l - 1], e->lw[l - 1], e->pyr[l].data(), e->lw[l], e->lh[l], e->lw[l]);
    }
    int total = 0;
    for (int l = 0; l < L; l++) {
        const int W = e->lw[l], H = e->lh[l];
        level_candidates(e->pyr[l].data(), W, H, e->iniTh, e->minTh, e->cands[l]);
        std::vector<Cand> rel(e->cands[l]);
        for (auto& c : rel) { c.x -= 16; c.y -= 16; }
        std::vector<int> sel = distribute(rel, W - 32, H - 32, e->perLevel[l]);
        e->levelCount[l] = (int)sel.size();
        if (sel.empty()) continue;
        e->blur[l].resize((size_t)W * H);
        gaussian_blur7_u8(e->pyr[l].data(), W, H, W, e->blur[l].data(), W);
        const int scaledPatch = (int)(PATCH_SIZE * e->scale[l]);
        for (int i : sel) {
            const Cand& c = e->cands[l][i];
            if (total < cap) {
                orbport_kp& k = kps[total];
                const float ang = ic_angle(e->pyr[l].data(), W, c.x, c.y, e->umax);
                descriptor(e->blur[l].data(), W, c.x, c.y, ang, desc + (size_t)total * 32);
                k.x = (float)c.x; k.y = (float)c.y;
                if (l != 0) { k.x *= e->scale[l]; k.y *= e->scale[l]; }
                k.size = (float)scaledPatch; k.angle = ang; k.response = (float)c.score; k.octave = l; k.class_id = -1;
            }
            total++;
        }
    }
    return total;
}

int orbport_level_size(void* h, int level, int* w, int* hgt) {
    Extractor* e = (Extractor*)h;
    if (level < 0 || level >= (int)e->lw.size()) return -1;
    *w = e->lw[level]; *hgt = e->lh[level];
    return 0;
}
const uint8_t* orbport_level_ptr(void* h, int level) { Extractor* e = (Extractor*)h; return e->pyr[level].empty() ? nullptr : e->pyr[level].data(); }
const uint8_t* orbport_blur_ptr(void* h, int level) { Extractor* e = (Extractor*)h; return e->blur[level].empty() ? nullptr : e->blur[level].data(); }
int orbport_candidates(void* h, int level, int32_t* xys, int cap) {
    Extractor* e = (Extractor*)h;
    const auto& c = e->cands[level];
    for (int i = 0; i < (int)c.size() && i < cap; i++) { xys[3 * i] = c[i].x; xys[3 * i + 1] = c[i].y; xys[3 * i + 2] = c[i].score; }
    return (int)c.size();
}
int orbport_level_count(void* h, int level) { return ((Extractor*)h)->levelCount[level]; }

int orbport_distribute(const int32_t* xys, int n, int width, int height, int N, int32_t* out_xys, int cap) {
    std::vector<Cand> P(n);
    for (int i = 0; i < n; i++) P[i] = Cand{xys[3 * i], xys[3 * i + 1], xys[3 * i + 2]};
    std::vector<int> sel = distribute(P, width, height, N);
    for (int i = 0; i < (int)sel.size() && i < cap; i++) {
        out_xys[3 * i] = P[sel[i]].x; out_xys[3 * i + 1] = P[sel[i]].y; out_xys[3 * i + 2] = P[sel[i]].score;
    }
    return (int)sel.size();
}

int orbport_hamming(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

}  // extern "C"
