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
// Output: unordered candidate list per (image, level) of packed (x,y,score); consumers break ties with
// the reference's emission order key (cell row, cell col, y, x), never with list position.
//
// Bound (target): HBM read of the level pixels, once — sum_l w_l*h_l bytes per image.
#include "borb_internal.h"

namespace borb {

namespace {

constexpr int TILE_PITCH = FAST_TILE_W + 8;   // smem image-tile pitch (domain + 3 halo each side, padded)
constexpr int TILE_ROWS = 64 + 6;             // hCell < 61

// S(p) if p is a FAST-9 corner at threshold t, else 0.  c points at p inside the smem tile.
__device__ __forceinline__ int fast_score(const uint8_t* c, int t) {
    constexpr int P = TILE_PITCH;
    const int v = c[0];
    int r[16];
    r[0] = c[3 * P];       r[1] = c[3 * P + 1];   r[2] = c[2 * P + 2];   r[3] = c[P + 3];
    r[4] = c[3];           r[5] = c[-P + 3];      r[6] = c[-2 * P + 2];  r[7] = c[-3 * P + 1];
    r[8] = c[-3 * P];      r[9] = c[-3 * P - 1];  r[10] = c[-2 * P - 2]; r[11] = c[-P - 3];
    r[12] = c[-3];         r[13] = c[P - 3];      r[14] = c[2 * P - 2];  r[15] = c[3 * P - 1];
    const int hi = v + t, lo = v - t;
    unsigned B = 0, D = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        B |= (r[k] > hi ? 1u : 0u) << k;
        D |= (r[k] < lo ? 1u : 0u) << k;
    }
    // 9 contiguous set bits in the circular 16-bit mask
    auto arc9 = [](unsigned m) -> bool {
        unsigned x = m | (m << 16);
        unsigned a = x & (x >> 1);
        a &= a >> 2;
        a &= a >> 4;
        a &= x >> 8;
        return (a & 0xFFFFu) != 0;
    };
    const bool cb = arc9(B), cd = arc9(D);
    if (!cb && !cd) return 0;
    // score: max over arcs of min over the arc of the signed difference, for the polarity that fired
    // (both cannot fire: 9+9 > 16).  Sliding 9-window minimum on the circular sequence by doubling.
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = cd ? (v - r[k]) : (r[k] - v);
    int m2[16], m4[16], m8[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m2[k] = min(d[k], d[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) m8[k] = min(m4[k], m4[(k + 4) & 15]);
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) best = max(best, min(m8[k], d[(k + 8) & 15]));
    return best - 1;
}

}  // namespace

__global__ void __launch_bounds__(256) fast_kernel(const __grid_constant__ Geometry g, const uint8_t* __restrict__ pyr,
                                                   uint32_t* __restrict__ cand, int* __restrict__ cand_cnt) {
    __shared__ __align__(16) uint8_t tile[TILE_ROWS * TILE_PITCH];
    __shared__ uint8_t score[64 * FAST_TILE_W];
    __shared__ int cellHasIni[FAST_TILE_W / 30 + 1];

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
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const uint8_t* src = pyr + (size_t)img * g.pyr_image_stride + L.pyr_off;

    if (tid < FAST_TILE_W / 30 + 1) cellHasIni[tid] = 0;
    // stage the tile (domain + 3-px ring halo).  [19,W-19) x [19,H-19) keeps every read >= 16 px inside.
    {
        const int lw = tw + 6, lh = th + 6;
        for (int yy = threadIdx.y; yy < lh; yy += 8) {
            const uint8_t* row = src + (size_t)(y0 - 3 + yy) * L.pitch + (x0 - 3);
            for (int xx = threadIdx.x; xx < lw; xx += 32) tile[yy * TILE_PITCH + xx] = row[xx];
        }
    }
    __syncthreads();
    const int tlow = min(g.ini_th, g.min_th);
    for (int yy = threadIdx.y; yy < th; yy += 8)
        for (int xx = threadIdx.x; xx < tw; xx += 32)
            score[yy * FAST_TILE_W + xx] = (uint8_t)fast_score(&tile[(yy + 3) * TILE_PITCH + xx + 3], tlow);
    __syncthreads();
    // cell-local strict NMS; survivors overwrite the tile buffer (reused as "kept score" map)
    uint8_t* kept = tile;
    for (int yy = threadIdx.y; yy < th; yy += 8)
        for (int xx = threadIdx.x; xx < tw; xx += 32) {
            const int s = score[yy * FAST_TILE_W + xx];
            int k = 0;
            if (s > 0) {
                const int c = xx / L.wCell;
                const int cx0 = c * L.wCell, cx1 = min(cx0 + L.wCell, tw);
                bool ismax = true;
#pragma unroll
                for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                    for (int dx = -1; dx <= 1; dx++) {
                        if (dx == 0 && dy == 0) continue;
                        const int qx = xx + dx, qy = yy + dy;
                        if (qx < cx0 || qx >= cx1 || qy < 0 || qy >= th) continue;
                        if (!(s > (int)score[qy * FAST_TILE_W + qx])) ismax = false;
                    }
                if (ismax) {
                    k = s;
                    if (s >= g.ini_th) cellHasIni[c] = 1;
                }
            }
            kept[yy * FAST_TILE_W + xx] = (uint8_t)k;
        }
    __syncthreads();
    // emit (warp-aggregated append)
    uint32_t* out = cand + (size_t)img * g.cand_image_stride + L.cand_off;
    int* cnt = cand_cnt + img * g.nlevels + l;
    for (int yy = threadIdx.y; yy < th; yy += 8)
        for (int xb = 0; xb < tw; xb += 32) {
            const int xx = xb + threadIdx.x;
            int s = 0;
            if (xx < tw) {
                s = kept[yy * FAST_TILE_W + xx];
                if (s > 0) {
                    const int t = cellHasIni[xx / L.wCell] ? g.ini_th : g.min_th;
                    if (s < t) s = 0;
                }
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, s > 0);
            if (m) {
                int base = 0;
                if (threadIdx.x == 0) base = atomicAdd(cnt, __popc(m));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (s > 0) {
                    const int pos = base + __popc(m & ((1u << threadIdx.x) - 1));
                    if (pos < L.cand_cap) out[pos] = pack_xys(x0 + xx, y0 + yy, s);
                }
            }
        }
}

int launch_fast(const Geometry& g, const Workspace& ws, int n_images, cudaStream_t s) {
    dim3 block(32, 8), grid(g.fast_blocks, n_images);
    fast_kernel<<<grid, block, 0, s>>>(g, ws.pyr, ws.cand, ws.cand_cnt);
    return 1;
}

}  // namespace borb
