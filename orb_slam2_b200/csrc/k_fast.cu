// FAST-9/16 score map + cell-local strict 3x3 NMS + per-cell ini/min threshold selection, for every
// level of every image of the batch in ONE launch.
//
// Replaces the per-cell cv::FAST loop of ORBextractor::ComputeKeyPointsOctTree (reference
// src/ORBextractor.cc:784-829; OpenCV features2d/fast.cpp FAST_t<16> + cornerScore<16>) using the
// whole-level reformulation of SURVEY §8(a3), verified identical to the per-cell loop in
// tests/test_oracle_extract.py:
//   S(p)    = max over the 16 nine-pixel arcs (both polarities) of min |I_p - I_q|, minus 1
//             (p is a corner at threshold t  <=>  S(p) >= t);
//   keep(p) = S(p) strictly greater than S(q) for the 8-neighbours q that lie in the SAME cell's
//             detection domain (neighbours outside count as 0);
//   a cell emits keep(p) with S>=iniTh if any exists, else keep(p) with S>=minTh.
// One CTA owns `cellsPerBlk` whole cells of one cell row, so NMS and the threshold decision are CTA-local.
//
// Pipeline inside a CTA (v2):
//   0. one elected thread issues a 3-D TMA tile load (cp.async.bulk.tensor, box 160 x (hCell+6) bytes at
//      ((x0-4)&~15, y0-3, image): the inner start coordinate must be 16-byte aligned) into shared memory
//      and everybody waits on its mbarrier;
//   1. packed quick-reject: a thread owns the 4 pixels of one ALIGNED 32-bit word of the tile and
//      slides down its rows; per ring position one VABSDIFF4 gives |I_q - I_p| for the 4 pixels and three
//      logic ops turn it into a per-byte ">t" flag; a FAST-9 arc contains at least one pixel of each of
//      the 8 antipodal ring pairs, so AND_j (f_j | f_{j+8}) == 0 rejects the pixel.  Even positions first;
//      odd positions only if something survives.  Survivors are appended to a shared-memory queue;
//   2. exact arc test + score for queued pixels only;  3. cell-local NMS for scored pixels only;
//   4. per-cell threshold decision and warp-aggregated append to the global candidate list.
// Output: unordered candidate list per (image, level) of packed (x,y,score); consumers break ties with
// the reference's emission order key (cell row, cell col, y, x), never with list position.
//
// Bound (target): HBM read of the level pixels, once — sum_l w_l*h_l bytes per image.  Measured: issue
// bound (see profiles/); the packed reject is what keeps the instruction count per pixel low.
#include <cuda.h>

#include <cstring>

#include "borb_internal.h"

namespace borb {

namespace {

constexpr int TP = 160;                 // TMA box width == smem tile pitch (bytes).  TMA needs a 16-byte aligned start
                                        // column, so the box starts at xs = (x0-4) & ~15 and domain px xx sits at column xx+off
constexpr int TROWS = 66;               // hCell <= 60, + 3 halo rows above and below
constexpr int SW = 128;                 // score-map pitch
constexpr int QCAP = 60 * 128;          // queue capacity >= every pixel of the largest tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Scalar 9-arc test (only used for thresholds > 127, outside the packed path): 0 none, 1 bright arc, 2 dark arc.
__device__ __forceinline__ int scalar_arc(const uint8_t* c, int t) {
    constexpr int P = TP;
    const int offs[16] = {3 * P, 3 * P + 1, 2 * P + 2, P + 3, 3, -P + 3, -2 * P + 2, -3 * P + 1,
                          -3 * P, -3 * P - 1, -2 * P - 2, -P - 3, -3, P - 3, 2 * P - 2, 3 * P - 1};
    const int v = c[0];
    unsigned B = 0, D = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int r = c[offs[k]];
        B |= (r > v + t ? 1u : 0u) << k;
        D |= (r < v - t ? 1u : 0u) << k;
    }
    auto arc9 = [](unsigned m) -> bool {
        unsigned x = m | (m << 16);
        unsigned a = x & (x >> 1);
        a &= a >> 2;
        a &= a >> 4;
        a &= x >> 8;
        return (a & 0xFFFFu) != 0;
    };
    return arc9(B) ? 1 : (arc9(D) ? 2 : 0);
}

// S(p) for a pixel already known to be a FAST-9 corner with the given polarity (dark: ring < centre).
// = max over the 16 nine-pixel arcs of the min signed difference, minus 1 (cornerScore<16>); the other
// polarity cannot have a 9-arc (9+9 > 16).  Sliding 9-window minimum as min3 of min3 on the circular ring.
__device__ __forceinline__ int corner_score(const uint8_t* c, bool dark) {
    constexpr int P = TP;
    const int v = c[0];
    int d[16];
    d[0] = c[3 * P];       d[1] = c[3 * P + 1];   d[2] = c[2 * P + 2];   d[3] = c[P + 3];
    d[4] = c[3];           d[5] = c[-P + 3];      d[6] = c[-2 * P + 2];  d[7] = c[-3 * P + 1];
    d[8] = c[-3 * P];      d[9] = c[-3 * P - 1];  d[10] = c[-2 * P - 2]; d[11] = c[-P - 3];
    d[12] = c[-3];         d[13] = c[P - 3];      d[14] = c[2 * P - 2];  d[15] = c[3 * P - 1];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = dark ? (v - d[k]) : (d[k] - v);
    int m3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m3[k] = min(min(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const int a = min(min(m3[k], m3[(k + 3) & 15]), m3[(k + 6) & 15]);
        const int b = min(min(m3[k + 1], m3[(k + 4) & 15]), m3[(k + 7) & 15]);
        best = max(max(best, a), b);
    }
    return best - 1;
}

// 4-byte window starting DX bytes after the start of W1 (W0|W1|W2 are three consecutive aligned words)
template <int DX>
__device__ __forceinline__ uint32_t win(uint32_t W0, uint32_t W1, uint32_t W2) {
    if (DX == 0) return W1;
    if (DX > 0) return __byte_perm(W1, W2, DX | ((DX + 1) << 4) | ((DX + 2) << 8) | ((DX + 3) << 12));
    constexpr int K = 4 + DX;
    return __byte_perm(W0, W1, K | ((K + 1) << 4) | ((K + 2) << 8) | ((K + 3) << 12));
}

// bit 7 of every byte of the result is set iff that byte of |q - v| exceeds t  (T1 = (t+1)*0x01010101, t <= 127)
__device__ __forceinline__ uint32_t gt_flag(uint32_t q, uint32_t v, uint32_t T1) {
    const uint32_t a = __vabsdiffu4(q, v);
    return ((a | 0x80808080u) - T1) | a;
}

constexpr uint32_t M7 = 0x7f7f7f7fu;
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t nb, uint32_t c) { return (a & nb) | (a & c) | (nb & c); }
// bit 7 of each byte: q > hi (unsigned bytes);  nhi7 = ~hi & M7 precomputed
__device__ __forceinline__ uint32_t gtu7(uint32_t q, uint32_t hi, uint32_t nhi7) { return maj3(q, ~hi, (q & M7) + nhi7); }
// bit 7 of each byte: lo > q;  lo7 = lo & M7 precomputed
__device__ __forceinline__ uint32_t ltu7(uint32_t q, uint32_t lo, uint32_t lo7) { return maj3(lo, ~q, lo7 + (~q & M7)); }

}  // namespace

struct TMaps { CUtensorMap m[BORB_MAX_LEVELS]; };

__global__ void __launch_bounds__(256) fast_kernel(const __grid_constant__ Geometry g, const __grid_constant__ TMaps tm,
                                                   uint32_t* __restrict__ cand, int* __restrict__ cand_cnt) {
    __shared__ __align__(128) uint8_t tile[TROWS * TP];
    __shared__ __align__(16) uint8_t score[60 * SW];
    __shared__ uint16_t queue[QCAP];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint16_t wqueue[60 * 32];    // words (row, lane) that survive the cheap reject
    __shared__ int qn, wqn;
    __shared__ int cellHasIni[128 / 30 + 1];
    __shared__ uint8_t cellOf[128];

    const int img = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && (int)blockIdx.x >= g.lv[l + 1].blkBase) l++;
    const LevelGeom& L = g.lv[l];
    const int local = blockIdx.x - L.blkBase;
    const int cellRow = local / L.blkCols, blkCol = local - cellRow * L.blkCols;
    const int cell0 = blkCol * L.cellsPerBlk;
    const int ncell = min(L.cellsPerBlk, L.nCols - cell0);
    const int x0 = EDGE + cell0 * L.wCell, x1 = min(x0 + ncell * L.wCell, L.w - EDGE);
    const int y0 = EDGE + cellRow * L.hCell, y1 = min(y0 + L.hCell, L.h - EDGE);
    if (x0 >= x1 || y0 >= y1) return;
    const int tw = x1 - x0, th = y1 - y0;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wrp = tid >> 5;

    // ---- 0. TMA: tile rows y0-3 .. y0+hCell+2, columns xs .. xs+159 of image `img`, level l
    const int xs = (x0 - 4) & ~15;          // 16-byte aligned box start (TMA requirement)
    const int off = x0 - xs;                // tile column of domain pixel xx = 0   (4..19)
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)TP * (uint32_t)(L.hCell + 6);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tm.m[l])), "r"(xs), "r"(y0 - 3), "r"(img), "r"(smem_u32(&bar))
            : "memory");
    }
    // overlap with the copy: clear the score map and the bookkeeping
    for (int i = tid; i < (th * SW) / 16; i += 256) reinterpret_cast<uint4*>(score)[i] = make_uint4(0, 0, 0, 0);
    if (tid < 128 / 30 + 1) cellHasIni[tid] = 0;
    if (tid < 128) cellOf[tid] = (uint8_t)(tid / L.wCell);
    if (tid == 0) { qn = 0; wqn = 0; }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FAST_TMA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra FAST_TMA_DONE;\n"
        "bra FAST_TMA_WAIT;\n"
        "FAST_TMA_DONE:\n"
        "}\n" ::"r"(smem_u32(&bar))
        : "memory");
    __syncthreads();

    const int tlow = min(g.ini_th, g.min_th);
    const bool packed_ok = tlow <= 127;

    // ---- 1a. packed reject on |diff| of the even ring positions; surviving WORDS (4 px) go to a word queue so
    //          that the exact test below runs with all lanes busy
    const uint32_t* T32 = reinterpret_cast<const uint32_t*>(tile);
    const int wbase = off >> 2;             // lane owns aligned tile word wbase+lane; byte b is domain px xx = 4*lane-(off&3)+b
    const uint32_t T1 = (uint32_t)(tlow + 1) * 0x01010101u;
    const uint32_t Tt = (uint32_t)tlow * 0x01010101u;
    {
        const int RG = (th + 7) >> 3;                 // rows per warp (8 warps)
        const int yBeg = wrp * RG, yEnd = min(th, yBeg + RG);
        const int xxb = 4 * lane - (off & 3);
        uint32_t vmask = 0;
#pragma unroll
        for (int b = 0; b < 4; b++)
            if (xxb + b >= 0 && xxb + b < tw) vmask |= 0x80u << (8 * b);
        // rolling window of 7 tile rows x 3 words; slot (j % 7) holds tile row (yy + j), j = 0..6 <=> dy = j-3
        uint32_t a0[7], a1[7], a2[7];
        if (yBeg < yEnd) {
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const uint32_t* rp = T32 + (yBeg + j) * (TP / 4) + wbase + lane - 1;
                a0[j] = rp[0]; a1[j] = rp[1]; a2[j] = rp[2];
            }
        }
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int yy = yBeg + it;
            if (yy < yEnd) {          // warp-uniform
                {
                    const uint32_t* rp = T32 + (yy + 6) * (TP / 4) + wbase + lane - 1;
                    a0[(it + 6) % 7] = rp[0]; a1[(it + 6) % 7] = rp[1]; a2[(it + 6) % 7] = rp[2];
                }
#define ROW(dy) a0[(it + (dy) + 3) % 7], a1[(it + (dy) + 3) % 7], a2[(it + (dy) + 3) % 7]
                bool keep = vmask != 0;
                if (packed_ok) {
                    const uint32_t v = a1[(it + 3) % 7];
                    uint32_t acc = gt_flag(win<0>(ROW(3)), v, T1) | gt_flag(win<0>(ROW(-3)), v, T1);      // pair (0,8)
                    acc &= gt_flag(win<2>(ROW(2)), v, T1) | gt_flag(win<-2>(ROW(-2)), v, T1);               // (2,10)
                    acc &= gt_flag(win<3>(ROW(0)), v, T1) | gt_flag(win<-3>(ROW(0)), v, T1);                // (4,12)
                    acc &= gt_flag(win<2>(ROW(-2)), v, T1) | gt_flag(win<-2>(ROW(2)), v, T1);               // (6,14)
                    keep = (acc & vmask) != 0;
                }
#undef ROW
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
                if (bal) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&wqn, __popc(bal));
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (keep) wqueue[base + __popc(bal & ((1u << lane) - 1))] = (uint16_t)((yy << 5) | lane);
                }
            }
        }
    }
    __syncthreads();

    // ---- 1b. exact, sign-aware 9-arc test on the queued words: per polarity the 16 per-position flags (bit 7 of
    //          each byte = that pixel's flag), antipodal early-outs, then "9 contiguous" as AND of three 3-runs
    {
        const int nw = wqn;
        for (int eb = 0; eb < nw; eb += 256) {
            const int e = eb + tid;
            uint32_t m = 0, mdark = 0;
            int yy = 0, xxb = 0;
            if (e < nw) {
                const int we = wqueue[e];
                yy = we >> 5;
                const int ln = we & 31;
                xxb = 4 * ln - (off & 3);
                uint32_t vmask = 0;
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (xxb + b >= 0 && xxb + b < tw) vmask |= 0x80u << (8 * b);
                if (packed_ok) {
                    uint32_t a0[7], a1[7], a2[7];
#pragma unroll
                    for (int j = 0; j < 7; j++) {
                        const uint32_t* rp = T32 + (yy + j) * (TP / 4) + wbase + ln - 1;
                        a0[j] = rp[0]; a1[j] = rp[1]; a2[j] = rp[2];
                    }
#define ROW(dy) a0[(dy) + 3], a1[(dy) + 3], a2[(dy) + 3]
                    const uint32_t v = a1[3];
                    const uint32_t hi = __vaddus4(v, Tt), lo = __vsubus4(v, Tt);
#pragma unroll
                    for (int pol = 0; pol < 2; pol++) {
                        uint32_t f[16];
                        const uint32_t k7 = pol ? (lo & M7) : (~hi & M7);
#define FLAG(q) (pol ? ltu7((q), lo, k7) : gtu7((q), hi, k7))
                        f[0] = FLAG(win<0>(ROW(3)));   f[8] = FLAG(win<0>(ROW(-3)));
                        f[2] = FLAG(win<2>(ROW(2)));   f[10] = FLAG(win<-2>(ROW(-2)));
                        f[4] = FLAG(win<3>(ROW(0)));   f[12] = FLAG(win<-3>(ROW(0)));
                        f[6] = FLAG(win<2>(ROW(-2)));  f[14] = FLAG(win<-2>(ROW(2)));
                        uint32_t ap = (f[0] | f[8]) & (f[2] | f[10]) & (f[4] | f[12]) & (f[6] | f[14]);
                        if (ap & vmask) {
                            f[1] = FLAG(win<1>(ROW(3)));   f[9] = FLAG(win<-1>(ROW(-3)));
                            f[3] = FLAG(win<3>(ROW(1)));   f[11] = FLAG(win<-3>(ROW(-1)));
                            f[5] = FLAG(win<3>(ROW(-1)));  f[13] = FLAG(win<-3>(ROW(1)));
                            f[7] = FLAG(win<1>(ROW(-3)));  f[15] = FLAG(win<-1>(ROW(3)));
                            ap &= (f[1] | f[9]) & (f[3] | f[11]) & (f[5] | f[13]) & (f[7] | f[15]);
                            if (ap & vmask) {
                                uint32_t p3[16];
#pragma unroll
                                for (int k = 0; k < 16; k++) p3[k] = f[k] & f[(k + 1) & 15] & f[(k + 2) & 15];
                                uint32_t any9 = 0;
#pragma unroll
                                for (int k = 0; k < 16; k++) any9 |= p3[k] & p3[(k + 3) & 15] & p3[(k + 6) & 15];
                                any9 &= vmask;
                                m |= any9;
                                if (pol) mdark = any9;
                            }
                        }
#undef FLAG
                    }
#undef ROW
                } else {
                    m = vmask;          // thresholds above 127 (never used by the reference configs): scalar test in phase 2
                }
            }
            // append this word's corner pixels to the corner queue (warp-aggregated)
            const unsigned any = __ballot_sync(0xFFFFFFFFu, m != 0);
            if (any) {
                const int c = __popc(m);
                int incl = c;
#pragma unroll
                for (int o2 = 1; o2 < 32; o2 <<= 1) {
                    const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o2);
                    if (lane >= o2) incl += t;
                }
                int base = 0;
                if (lane == 31) base = atomicAdd(&qn, incl);
                base = __shfl_sync(0xFFFFFFFFu, base, 31) + incl - c;
                uint32_t mm = m;
                while (mm) {
                    const int b = (__ffs(mm) - 1) >> 3;
                    mm &= mm - 1;
                    queue[base++] = (uint16_t)((((mdark >> (8 * b + 7)) & 1u) << 15) | (yy << 7) | (xxb + b));
                }
            }
        }
    }
    __syncthreads();
    const int nq = qn;

    // ---- 2. score of every corner (polarity known from the packed test)
    for (int e = tid; e < nq; e += 256) {
        const int q = queue[e];
        const int xx = q & 127, yy = (q >> 7) & 63;
        const uint8_t* c = &tile[(yy + 3) * TP + xx + off];
        bool dark = (q >> 15) != 0;
        bool corner = true;
        if (!packed_ok) {               // thresholds > 127: scalar 9-arc test decides corner-ness and polarity
            const int pol = scalar_arc(c, tlow);
            corner = pol != 0;
            dark = pol == 2;
        }
        if (corner) score[yy * SW + xx] = (uint8_t)corner_score(c, dark);
        else queue[e] = 0xFFFF;
    }
    __syncthreads();

    // ---- 3. cell-local strict NMS for the corners (0xFFFF = dropped)
    for (int e = tid; e < nq; e += 256) {
        const int q = queue[e];
        if (q == 0xFFFF) continue;
        const int xx = q & 127, yy = (q >> 7) & 63;
        const int s = score[yy * SW + xx];
        const int c = cellOf[xx];
        const int cx0 = c * L.wCell, cx1 = min(cx0 + L.wCell, tw);
        bool ismax = s > 0;
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                if (dx == 0 && dy == 0) continue;
                const int qx = xx + dx, qy = yy + dy;
                if (qx < cx0 || qx >= cx1 || qy < 0 || qy >= th) continue;
                if (!(s > (int)score[qy * SW + qx])) ismax = false;
            }
        if (ismax) {
            if (s >= g.ini_th) cellHasIni[c] = 1;
        } else
            queue[e] = 0xFFFF;
    }
    __syncthreads();

    // ---- 4. per-cell threshold + emit (warp-aggregated append)
    uint32_t* out = cand + (size_t)img * g.cand_image_stride + L.cand_off;
    int* cnt = cand_cnt + img * g.nlevels + l;
    for (int eb = 0; eb < nq; eb += 256) {
        const int e = eb + tid;
        int s = 0, xx = 0, yy = 0;
        if (e < nq) {
            const int q = queue[e];
            if (q != 0xFFFF) {
                xx = q & 127; yy = (q >> 7) & 63;
                s = score[yy * SW + xx];
                const int t = cellHasIni[cellOf[xx]] ? g.ini_th : g.min_th;
                if (s < t) s = 0;
            }
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, s > 0);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(cnt, __popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (s > 0) {
                const int pos = base + __popc(m & ((1u << lane) - 1));
                if (pos < L.cand_cap) out[pos] = pack_xys(x0 + xx, y0 + yy, s);
            }
        }
    }
}

// ---- host: tensor maps (one per level: 3-D {x, y, image} view of the pyramid buffer)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

borb_status build_fast_tmaps(const Geometry& g, const Workspace& ws, void* out_tmaps) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
            return BORB_ERR_CUDA;
        }
        encode = (EncodeTiledFn)fn;
    }
    TMaps* tm = reinterpret_cast<TMaps*>(out_tmaps);
    std::memset(tm, 0, sizeof(TMaps));
    for (int l = 0; l < g.nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        cuuint64_t dims[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)ws.max_images};
        cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.pyr_image_stride};
        cuuint32_t box[3] = {(cuuint32_t)TP, (cuuint32_t)(L.hCell + 6), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tm->m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ws.pyr + L.pyr_off, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled failed for level %d (CUresult %d)", l, (int)r);
            return BORB_ERR_CUDA;
        }
    }
    return BORB_OK;
}

size_t fast_tmaps_bytes() { return sizeof(TMaps); }

int launch_fast(const Geometry& g, const Workspace& ws, int n_images, cudaStream_t s) {
    dim3 grid(g.fast_blocks, n_images);
    fast_kernel<<<grid, 256, 0, s>>>(g, *reinterpret_cast<const TMaps*>(ws.fast_tmaps), ws.cand, ws.cand_cnt);
    return 1;
}

}  // namespace borb
