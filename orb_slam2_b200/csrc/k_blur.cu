// 7x7 sigma-2 Gaussian blur of every pyramid level (the image the rBRIEF tests sample).
//
// Replaces cv::GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) at reference
// src/ORBextractor.cc:1085-1086 with OpenCV >= 3.4 fixed-point semantics (pinned against cv2 4.13 in
// tests/test_oracle_prims.py): separable kernel q = [18 34 48 56 48 34 18]/256; row pass exact u16,
// column pass (sum + 2^15) >> 16; reflect-101 of the LEVEL itself at its borders.
// One launch covers all levels of all images; a CTA produces a 64x32 tile through shared memory.
//
// Bound: HBM/L2 streaming (read + write of sum_l w_l*h_l bytes per image).
#include "borb_internal.h"

namespace borb {

namespace {
constexpr int BT_W = 64, BT_H = 32;
__device__ __forceinline__ int reflect101(int p, int len) {
    // |p| excursions are <= 3 here; levels are >= 7 px wide/tall
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return min(max(p, 0), len - 1);   // clamp only matters for tile lanes beyond the image (results discarded)
}
}  // namespace

__global__ void __launch_bounds__(256) blur_kernel(const __grid_constant__ Geometry g, const uint8_t* __restrict__ pyr,
                                                   uint8_t* __restrict__ blur) {
    __shared__ uint8_t in[(BT_H + 6) * (BT_W + 8)];
    __shared__ uint16_t rowp[(BT_H + 6) * BT_W];
    const int img = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nlevels && (int)blockIdx.x >= g.blur_base[l + 1]) l++;
    const LevelGeom& L = g.lv[l];
    const int local = blockIdx.x - g.blur_base[l];
    const int tilesX = (L.w + BT_W - 1) / BT_W;
    const int ty = local / tilesX, tx = local - ty * tilesX;
    const int x0 = tx * BT_W, y0 = ty * BT_H;
    const int tid = threadIdx.x;
    const uint8_t* src = pyr + (size_t)img * g.pyr_image_stride + L.pyr_off;
    uint8_t* dst = blur + (size_t)img * g.pyr_image_stride + L.pyr_off;
    constexpr int IW = BT_W + 6, IP = BT_W + 8, IH = BT_H + 6;
    for (int i = tid; i < IW * IH; i += 256) {
        const int yy = i / IW, xx = i - yy * IW;
        const int sx = reflect101(x0 + xx - 3, L.w), sy = reflect101(y0 + yy - 3, L.h);
        in[yy * IP + xx] = src[(size_t)sy * L.pitch + sx];
    }
    __syncthreads();
    for (int i = tid; i < IH * BT_W; i += 256) {
        const int yy = i / BT_W, xx = i - yy * BT_W;
        const uint8_t* p = &in[yy * IP + xx];
        rowp[i] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    for (int i = tid; i < BT_H * BT_W; i += 256) {
        const int yy = i / BT_W, xx = i - yy * BT_W;
        const int gx = x0 + xx, gy = y0 + yy;
        if (gx < L.w && gy < L.h) {
            const uint16_t* p = &rowp[yy * BT_W + xx];
            const uint32_t s = 18u * (p[0] + p[6 * BT_W]) + 34u * (p[BT_W] + p[5 * BT_W]) + 48u * (p[2 * BT_W] + p[4 * BT_W]) + 56u * p[3 * BT_W];
            dst[(size_t)gy * L.pitch + gx] = (uint8_t)min((s + 32768u) >> 16, 255u);
        }
    }
}

int launch_blur(const Geometry& g, const Workspace& ws, int n_images, cudaStream_t s) {
    dim3 grid(g.blur_tiles, n_images);
    blur_kernel<<<grid, 256, 0, s>>>(g, ws.pyr, ws.blur);
    return 1;
}

}  // namespace borb
