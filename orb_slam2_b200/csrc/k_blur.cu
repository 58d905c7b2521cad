// 7x7 sigma-2 Gaussian blur of every pyramid level (the image the rBRIEF tests sample).
//
// Replaces cv::GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) at reference
// src/ORBextractor.cc:1085-1086 with OpenCV >= 3.4 fixed-point semantics (pinned against cv2 4.13 in
// tests/test_oracle_prims.py): separable kernel q = [18 34 48 56 48 34 18]/256; row pass exact u16,
// column pass (sum + 2^15) >> 16; reflect-101 of the LEVEL itself at its borders.
//
// The right image edge needs no special case: every pyramid row ends in 8 bytes of reflect-101 padding.
// v3: no shared memory.  A warp owns a 120-px wide strip: lane k holds the aligned 32-bit word (4 px) at
// x = strip + 4(k-1), lanes 0 and 31 are halo only (their words feed the neighbours' windows by shuffle), and the
// warp slides down BT_R rows.  Per input row: one coalesced 128-byte load per warp, the row pass as two IDP.4A
// (dp4a) per pixel on PRMT-extracted byte windows (exact u16).  Row-pass results of consecutive rows are kept
// PACKED as u16 pairs (r, r+1), so the column pass is three IDP.2A (dp2a) + one IMAD per pixel:
//   s = 18 h[y-3] + 34 h[y-2] + 48 h[y-1] + 56 h[y] + 48 h[y+1] + 34 h[y+2] + 18 h[y+3] + 2^15,  out = s >> 16
// (s < 2^24, so the result is byte 2 of s: packed with PRMT, no clamp).  One launch covers all levels of all images.
//
// Bound: FMA/ALU issue (about 55 integer instructions per 4 output pixels); traffic = read + write of
// sum_l w_l*h_l bytes per image.
#include "borb_internal.h"

namespace borb {

namespace {
constexpr int BT_R = 16;                   // output rows per warp
constexpr int BT_WARPS = BLUR_TILE_H / BT_R;

__device__ __forceinline__ int reflect101(int p, int len) {
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return min(max(p, 0), len - 1);
}
}  // namespace

__global__ void __launch_bounds__(32 * BT_WARPS) blur_kernel(const __grid_constant__ Geometry g, const uint8_t* __restrict__ pyr,
                                                             uint8_t* __restrict__ blur) {
    const int img = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && (int)blockIdx.x >= g.blur_base[l + 1]) l++;
    const LevelGeom& L = g.lv[l];
    const int local = blockIdx.x - g.blur_base[l];
    const int tilesX = (L.w + BLUR_TILE_W - 1) / BLUR_TILE_W;
    const int ty = local / tilesX, tx = local - ty * tilesX;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int x = tx * BLUR_TILE_W + 4 * (lane - 1);     // first of this lane's 4 pixels (lane 0 / 31: halo words)
    const int yb = ty * BLUR_TILE_H + wrp * BT_R;          // first output row of this warp
    if (yb >= L.h) return;
    const uint8_t* src = pyr + (size_t)img * g.pyr_image_stride + L.pyr_off;
    uint8_t* dst = blur + (size_t)img * g.pyr_image_stride + L.pyr_off;
    const int W = L.w, H = L.h, pitch = L.pitch;
    const bool left_edge = (x == 0);
    const bool loads = x >= 0 && x < pitch;               // the word is inside the row buffer (pitch % 128 == 0, >= W + 8)
    const bool stores = lane >= 1 && lane <= 30 && x < W;
    const unsigned WLO = 18u | (34u << 8) | (48u << 16) | (56u << 24);
    const unsigned WHI = 48u | (34u << 8) | (18u << 16);

    // issue every row's load up front (independent, coalesced 128 B per warp) so that their latencies overlap
    uint32_t own[BT_R + 6];
    if (yb >= 3 && yb + BT_R + 2 < H) {                   // warp-uniform: no reflection, rows are consecutive
        const uint8_t* p = src + (size_t)(yb - 3) * pitch + x;
#pragma unroll
        for (int step = 0; step < BT_R + 6; step++, p += pitch) own[step] = loads ? *reinterpret_cast<const uint32_t*>(p) : 0u;
    } else {
#pragma unroll
        for (int step = 0; step < BT_R + 6; step++)
            own[step] = loads ? *reinterpret_cast<const uint32_t*>(src + (unsigned)(reflect101(yb + step - 3, H) * pitch) + x) : 0u;
    }

    uint32_t pk[6][4];     // pk[s % 6][i] = h[s][i] | h[s+1][i] << 16 for the last six input steps
    uint32_t hprev[4] = {0, 0, 0, 0};
#pragma unroll
    for (int step = 0; step < BT_R + 6; step++) {
        const uint32_t W1 = own[step];
        uint32_t W0 = __shfl_up_sync(0xFFFFFFFFu, W1, 1);
        const uint32_t W2 = __shfl_down_sync(0xFFFFFFFFu, W1, 1);
        // rows carry >= 8 bytes of reflect-101 padding after the last pixel (k_pyramid.cu), so the word after the last
        // pixel word is valid data for the right image edge too
        if (left_edge) W0 = __byte_perm(W1, W2, 0x1234);  // bytes -4..-1 = pixels 4,3,2,1 (reflect-101)
        uint32_t h[4];
        h[0] = __dp4a(__byte_perm(W0, W1, 0x4321), WLO, __dp4a(__byte_perm(W1, W2, 0x4321), WHI, 0u));
        h[1] = __dp4a(__byte_perm(W0, W1, 0x5432), WLO, __dp4a(__byte_perm(W1, W2, 0x5432), WHI, 0u));
        h[2] = __dp4a(__byte_perm(W0, W1, 0x6543), WLO, __dp4a(__byte_perm(W1, W2, 0x6543), WHI, 0u));
        h[3] = __dp4a(W1, WLO, __dp4a(W2, WHI, 0u));
        if (step >= 1) {
#pragma unroll
            for (int i = 0; i < 4; i++) pk[(step - 1) % 6][i] = __byte_perm(hprev[i], h[i], 0x5410);
        }
        if (step >= 6) {
            const int yout = yb + step - 6;
            if (yout < H && stores) {
                uint32_t s[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    // output row yout = input steps step-6 .. step: pairs (s-6,s-5) (s-4,s-3) (s-2,s-1) and h[step]
                    uint32_t acc = h[i] * 18u + 32768u;
                    acc = __dp2a_lo(pk[(step - 2) % 6][i], 48u | (34u << 8), acc);
                    acc = __dp2a_lo(pk[(step - 4) % 6][i], 48u | (56u << 8), acc);
                    s[i] = __dp2a_lo(pk[(step - 6) % 6][i], 18u | (34u << 8), acc);
                }
                const uint32_t out = __byte_perm(__byte_perm(s[0], s[1], 0x0062), __byte_perm(s[2], s[3], 0x0062), 0x5410);
                // rows are pitch-aligned and x % 4 == 0: one aligned store; bytes past W inside the pitch are scratch
                *reinterpret_cast<uint32_t*>(dst + (size_t)yout * pitch + x) = out;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) hprev[i] = h[i];
    }
}

int launch_blur(const Geometry& g, const Workspace& ws, int n_images, cudaStream_t s) {
    dim3 grid(g.blur_tiles, n_images);
    blur_kernel<<<grid, 32 * BT_WARPS, 0, s>>>(g, ws.pyr, ws.blur);
    return 1;
}

}  // namespace borb
