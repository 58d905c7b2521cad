// 7x7 sigma-2 Gaussian blur of every pyramid level (the image the rBRIEF tests sample).
//
// Replaces cv::GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) at reference
// src/ORBextractor.cc:1085-1086 with OpenCV >= 3.4 fixed-point semantics (pinned against cv2 4.13 in
// tests/test_oracle_prims.py): separable kernel q = [18 34 48 56 48 34 18]/256; row pass exact u16,
// column pass (sum + 2^15) >> 16; reflect-101 of the LEVEL itself at its borders.
//
// The right image edge needs no special case: every pyramid row ends in 8 bytes of reflect-101 padding.
// v2: no shared memory.  A warp owns a 128-px wide strip (one aligned 32-bit word = 4 px per lane) and
// slides down R rows: per input row one coalesced 128-byte load per warp, neighbour words by shuffle, the
// row pass as two IDP.4A (dp4a) per pixel on PRMT-extracted byte windows, a 7-row register window for the
// column pass, one aligned 32-bit store per 4 output pixels.  One launch covers all levels of all images.
//
// Bound: HBM/L2 streaming (read + write of sum_l w_l*h_l bytes per image).
#include "borb_internal.h"

namespace borb {

namespace {
constexpr int BT_W = 128;     // strip width (32 lanes x 4 px)
constexpr int BT_R = 8;       // output rows per warp
constexpr int BT_H = 64;      // rows per CTA (8 warps)

__device__ __forceinline__ int reflect101(int p, int len) {
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return min(max(p, 0), len - 1);
}
}  // namespace

__global__ void __launch_bounds__(256) blur_kernel(const __grid_constant__ Geometry g, const uint8_t* __restrict__ pyr,
                                                   uint8_t* __restrict__ blur) {
    const int img = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && (int)blockIdx.x >= g.blur_base[l + 1]) l++;
    const LevelGeom& L = g.lv[l];
    const int local = blockIdx.x - g.blur_base[l];
    const int tilesX = (L.w + BT_W - 1) / BT_W;
    const int ty = local / tilesX, tx = local - ty * tilesX;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int x = tx * BT_W + 4 * lane;            // first of this lane's 4 pixels
    const int yb = ty * BT_H + wrp * BT_R;          // first output row of this warp
    if (yb >= L.h) return;
    const uint8_t* src = pyr + (size_t)img * g.pyr_image_stride + L.pyr_off;
    uint8_t* dst = blur + (size_t)img * g.pyr_image_stride + L.pyr_off;
    const int W = L.w, H = L.h, pitch = L.pitch;
    const bool left_edge = (x == 0);
    const unsigned WLO = 18u | (34u << 8) | (48u << 16) | (56u << 24);
    const unsigned WHI = 48u | (34u << 8) | (18u << 16);

    // issue every row's load up front (independent, coalesced 128 B per warp) so that their latencies overlap
    uint32_t own[BT_R + 6];
#pragma unroll
    for (int step = 0; step < BT_R + 6; step++)
        own[step] = *reinterpret_cast<const uint32_t*>(src + (size_t)reflect101(yb + step - 3, H) * pitch + x);

    int win[7][4];    // row-pass results of the last 7 input rows (slot = input step % 7)
#pragma unroll
    for (int step = 0; step < BT_R + 6; step++) {
        const int yin = yb + step - 3;                    // input row feeding output rows yin-3 .. yin+3
        const int sy = reflect101(yin, H);
        const uint8_t* row = src + (size_t)sy * pitch;
        // the pitch is a multiple of 128 and x < pitch: the lane's own word is always inside the row buffer
        uint32_t W1 = own[step];
        uint32_t W0 = __shfl_up_sync(0xFFFFFFFFu, W1, 1);
        uint32_t W2 = __shfl_down_sync(0xFFFFFFFFu, W1, 1);
        if (lane == 0 && !left_edge) W0 = *reinterpret_cast<const uint32_t*>(row + x - 4);
        // rows carry >= 8 bytes of reflect-101 padding after the last pixel (k_pyramid.cu), so the word after the last
        // pixel word is valid data for the right image edge too
        if (lane == 31 && x + 4 < pitch) W2 = *reinterpret_cast<const uint32_t*>(row + x + 4);
        if (left_edge) W0 = __byte_perm(W1, W2, 0x1234);  // bytes -4..-1 = pixels 4,3,2,1 (reflect-101)
        int* r = win[step % 7];
        r[0] = __dp4a(__byte_perm(W0, W1, 0x4321), WLO, __dp4a(__byte_perm(W1, W2, 0x4321), WHI, 0u));
        r[1] = __dp4a(__byte_perm(W0, W1, 0x5432), WLO, __dp4a(__byte_perm(W1, W2, 0x5432), WHI, 0u));
        r[2] = __dp4a(__byte_perm(W0, W1, 0x6543), WLO, __dp4a(__byte_perm(W1, W2, 0x6543), WHI, 0u));
        r[3] = __dp4a(W1, WLO, __dp4a(W2, WHI, 0u));
        if (step >= 6) {
            const int yout = yb + step - 6;
            if (yout < H && x < W) {
                uint32_t out = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    // slots: output row yout uses input steps step-6 .. step
                    const int s = 18 * (win[(step - 6) % 7][i] + win[step % 7][i]) + 34 * (win[(step - 5) % 7][i] + win[(step - 1) % 7][i]) +
                                  48 * (win[(step - 4) % 7][i] + win[(step - 2) % 7][i]) + 56 * win[(step - 3) % 7][i];
                    out |= (uint32_t)min((unsigned)(s + 32768) >> 16, 255u) << (8 * i);
                }
                // rows are pitch-aligned and x % 4 == 0: one aligned store; bytes past W inside the pitch are scratch
                *reinterpret_cast<uint32_t*>(dst + (size_t)yout * pitch + x) = out;
            }
        }
    }
}

int launch_blur(const Geometry& g, const Workspace& ws, int n_images, cudaStream_t s) {
    dim3 grid(g.blur_tiles, n_images);
    blur_kernel<<<grid, 256, 0, s>>>(g, ws.pyr, ws.blur);
    return 1;
}

}  // namespace borb
