// Scale pyramid: level l = fixed-point bilinear downscale of level l-1 (chain), all images of the batch
// per launch.  Replaces ORBextractor::ComputePyramid (reference src/ORBextractor.cc:1107-1132), i.e.
// cv::resize(INTER_LINEAR) on CV_8UC1: 11-bit coefficient tables, horizontal pass in int32, vertical
// pass (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.  The 19-px reflect border the reference
// adds (copyMakeBorder) is never read by any later stage and is not materialised.
//
// Bound: HBM/L2 streaming; algorithmic bytes per image = sum_l (w_l*h_l) written + the same read.
#include "borb_internal.h"

namespace borb {

__global__ void __launch_bounds__(256) pyr_resize_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ pyr_out,
                                                         const int16_t* __restrict__ tabs, LevelGeom src, LevelGeom dst,
                                                         unsigned image_stride) {
    const int img = blockIdx.z;
    const int dy = blockIdx.y * 8 + threadIdx.y;
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (dy >= dst.h || dx0 >= dst.w) return;
    const uint8_t* S = pyr + (size_t)img * image_stride + src.pyr_off;
    uint8_t* D = pyr_out + (size_t)img * image_stride + dst.pyr_off + (size_t)dy * dst.pitch;
    const int16_t* xt = tabs + (size_t)dst.xtab_off * 3;
    const int16_t* yt = tabs + (size_t)dst.ytab_off * 3;
    const int sy = yt[dy * 3], b0 = yt[dy * 3 + 1], b1 = yt[dy * 3 + 2];
    const int sy0 = min(max(sy, 0), src.h - 1), sy1 = min(max(sy + 1, 0), src.h - 1);
    const uint8_t* R0 = S + (size_t)sy0 * src.pitch;
    const uint8_t* R1 = S + (size_t)sy1 * src.pitch;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int dx = dx0 + i;
        if (dx < dst.w) {
            const int sx = xt[dx * 3], a0 = xt[dx * 3 + 1], a1 = xt[dx * 3 + 2];
            const int sx1 = min(sx + 1, src.w - 1);
            const int r0 = R0[sx] * a0 + R0[sx1] * a1;
            const int r1 = R1[sx] * a0 + R1[sx1] * a1;
            const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 0xFF) << (8 * i);
        }
    }
    // rows are pitch-aligned (pitch % 128 == 0) and dx0 % 4 == 0: one aligned 32-bit store; the
    // padding bytes past dst.w inside the pitch are scratch.
    *reinterpret_cast<uint32_t*>(D + dx0) = out;
}

// Level 0 from a tightly packed landing buffer (one big H2D copy) into the pitched pyramid layout.
// slot i of the landing buffer holds image `first + i*step` of the batch (stereo: left/right interleave).
__global__ void __launch_bounds__(256) repack_kernel(const uint8_t* __restrict__ stage, int src_stride, size_t src_image_bytes,
                                                     uint8_t* __restrict__ pyr, LevelGeom l0, unsigned image_stride) {
    const int img = blockIdx.z, y = blockIdx.y;
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x0 >= l0.w) return;
    const uint8_t* S = stage + (size_t)img * src_image_bytes + (size_t)y * src_stride + x0;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (x0 + i < l0.w) v |= (uint32_t)S[i] << (8 * i);
    *reinterpret_cast<uint32_t*>(pyr + (size_t)img * image_stride + l0.pyr_off + (size_t)y * l0.pitch + x0) = v;
}

int launch_repack(const Geometry& g, const Workspace& ws, const uint8_t* stage, int src_stride, size_t src_image_bytes,
                  int n_images, cudaStream_t s) {
    dim3 grid((g.lv[0].w + 1023) / 1024, g.lv[0].h, n_images);
    repack_kernel<<<grid, 256, 0, s>>>(stage, src_stride, src_image_bytes, ws.pyr, g.lv[0], g.pyr_image_stride);
    return 1;
}

int launch_pyramid(const Geometry& g, const Workspace& ws, int n_images, cudaStream_t s) {
    int launches = 0;
    for (int l = 1; l < g.nlevels; l++) {
        const LevelGeom& d = g.lv[l];
        dim3 block(32, 8), grid((d.w + 127) / 128, (d.h + 7) / 8, n_images);
        pyr_resize_kernel<<<grid, block, 0, s>>>(ws.pyr, ws.pyr, ws.tabs, g.lv[l - 1], d, g.pyr_image_stride);
        launches++;
    }
    return launches;
}

}  // namespace borb
