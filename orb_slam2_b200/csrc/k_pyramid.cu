// Scale pyramid: level l = fixed-point bilinear downscale of level l-1 (chain), all images of the batch
// per launch.  Replaces ORBextractor::ComputePyramid (reference src/ORBextractor.cc:1107-1132), i.e.
// cv::resize(INTER_LINEAR) on CV_8UC1: 11-bit coefficient tables, horizontal pass in int32, vertical
// pass (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.
//
// Layout: every level row carries >= 8 bytes of reflect-101 padding after its last pixel (the only part of the
// reference's 19-px copyMakeBorder that any later stage needs: the 7x7 blur reads up to 7 bytes past the row end
// as aligned words).  The padding columns are produced here as ordinary pixels through table entries that point at
// the reflected column.
//
// A thread makes 4 destination pixels: one packed table entry per pixel (source column + two 11-bit weights), the two
// source rows as three aligned 32-bit words each, byte pairs picked with PRMT and the horizontal pass as one
// IDP.2A (dp2a) per pixel and row.
//
// Bound: ALU/LSU issue (integer work per pixel), traffic sum_{l>=1} w_l*h_l bytes written + read per image.
#include "borb_internal.h"

namespace borb {

namespace {
__device__ __forceinline__ uint32_t pick2(uint32_t w0, uint32_t w1, uint32_t w2, int pos) {
    // bytes pos, pos+1 (pos in 0..9) of the 12-byte window w0|w1|w2, in the low half of the result
    const int k = pos >> 2, sh = pos & 3;
    const uint32_t lo = k == 0 ? w0 : (k == 1 ? w1 : w2);
    const uint32_t hi = k == 0 ? w1 : w2;
    return __byte_perm(lo, hi, sh | ((sh + 1) << 4));
}
}  // namespace

// xt/yt entries: {offset, c0, c1, 0} as 4 x int16.  A thread makes 4 px x PYR_ROWS rows (the x entries are reused).
constexpr int PYR_ROWS = 1;    // (4 rows per thread measured no faster: the launches are CTA-latency bound)
__global__ void __launch_bounds__(256) pyr_resize_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ pyr_out,
                                                         const int16_t* __restrict__ tabs, LevelGeom src, LevelGeom dst,
                                                         unsigned image_stride) {
    const int img = blockIdx.z;
    const int dyb = (blockIdx.y * 8 + threadIdx.y) * PYR_ROWS;
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int wpad = (dst.w + 8 + 3) & ~3;                  // pixels + reflect padding, whole words
    if (dyb >= dst.h || dx0 >= wpad) return;
    const uint8_t* S = pyr + (size_t)img * image_stride + src.pyr_off;
    uint8_t* D = pyr_out + (size_t)img * image_stride + dst.pyr_off;
    const uint2* xt = reinterpret_cast<const uint2*>(tabs + (size_t)dst.xtab_off * 4);
    const uint2* yt = reinterpret_cast<const uint2*>(tabs + (size_t)dst.ytab_off * 4);
    uint2 e[4];
#pragma unroll
    for (int i = 0; i < 4; i++) e[i] = xt[dx0 + i];
    const int base = (int)(e[0].x & 0xFFFF) & ~3;
    int lo = (int)(e[0].x & 0xFFFF), hi = lo;
#pragma unroll
    for (int i = 1; i < 4; i++) { const int s = (int)(e[i].x & 0xFFFF); lo = min(lo, s); hi = max(hi, s); }
    const bool windowed = lo >= base && hi - base <= 9;     // the 4 pixels read source columns inside one 12-byte aligned window
    // issue all row loads first (independent), then the arithmetic
    uint32_t ra[PYR_ROWS][3], rc[PYR_ROWS][3];
    int b0[PYR_ROWS], b1[PYR_ROWS], sy0[PYR_ROWS], sy1[PYR_ROWS];
#pragma unroll
    for (int r = 0; r < PYR_ROWS; r++) {
        const int dy = min(dyb + r, dst.h - 1);
        const uint2 ye = yt[dy];
        const int sy = (int)(short)(ye.x & 0xFFFF);
        b0[r] = (int)(ye.x >> 16); b1[r] = (int)(ye.y & 0xFFFF);
        sy0[r] = min(max(sy, 0), src.h - 1); sy1[r] = min(max(sy + 1, 0), src.h - 1);
        if (windowed) {
            const uint32_t* R0 = reinterpret_cast<const uint32_t*>(S + (size_t)sy0[r] * src.pitch + base);
            const uint32_t* R1 = reinterpret_cast<const uint32_t*>(S + (size_t)sy1[r] * src.pitch + base);
            ra[r][0] = R0[0]; ra[r][1] = R0[1]; ra[r][2] = R0[2];
            rc[r][0] = R1[0]; rc[r][1] = R1[1]; rc[r][2] = R1[2];
        }
    }
#pragma unroll
    for (int r = 0; r < PYR_ROWS; r++) {
        const int dy = dyb + r;
        if (dy >= dst.h) break;
        uint32_t out = 0;
        if (windowed) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int pos = (int)(e[i].x & 0xFFFF) - base;
                const uint32_t wts = (e[i].x >> 16) | (e[i].y << 16);           // c0 | c1 << 16
                const int r0 = (int)__dp2a_lo(wts, pick2(ra[r][0], ra[r][1], ra[r][2], pos), 0u);
                const int r1 = (int)__dp2a_lo(wts, pick2(rc[r][0], rc[r][1], rc[r][2], pos), 0u);
                const int v = (((b0[r] * (r0 >> 4)) >> 16) + ((b1[r] * (r1 >> 4)) >> 16) + 2) >> 2;
                out |= (uint32_t)(v & 0xFF) << (8 * i);
            }
        } else {
            // padding words that straddle the reflection point: source columns are not monotone; plain byte loads
            const uint8_t* R0 = S + (size_t)sy0[r] * src.pitch;
            const uint8_t* R1 = S + (size_t)sy1[r] * src.pitch;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int sx = (int)(e[i].x & 0xFFFF), x1 = min(sx + 1, src.w - 1);
                const int w0 = (int)(e[i].x >> 16), w1 = (int)(e[i].y & 0xFFFF);
                const int r0 = R0[sx] * w0 + R0[x1] * w1;
                const int r1 = R1[sx] * w0 + R1[x1] * w1;
                const int v = (((b0[r] * (r0 >> 4)) >> 16) + ((b1[r] * (r1 >> 4)) >> 16) + 2) >> 2;
                out |= (uint32_t)(v & 0xFF) << (8 * i);
            }
        }
        // rows are pitch-aligned (pitch % 128 == 0, pitch >= w + 8) and dx0 % 4 == 0: one aligned 32-bit store
        *reinterpret_cast<uint32_t*>(D + (size_t)dy * dst.pitch + dx0) = out;
    }
}

// Level 0 from a tightly packed landing buffer (one big H2D copy) into the pitched pyramid layout.
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

// reflect-101 padding of level 0 (columns w .. w+7 = columns w-2 .. w-9), after any kind of upload
__global__ void __launch_bounds__(256) pad_level0_kernel(uint8_t* __restrict__ pyr, LevelGeom l0, unsigned image_stride) {
    const int y = blockIdx.x * 256 + threadIdx.x, img = blockIdx.y;
    if (y >= l0.h) return;
    uint8_t* row = pyr + (size_t)img * image_stride + l0.pyr_off + (size_t)y * l0.pitch;
#pragma unroll
    for (int k = 0; k < 8; k++) row[l0.w + k] = row[max(l0.w - 2 - k, 0)];
}

int launch_repack(const Geometry& g, const Workspace& ws, const uint8_t* stage, int src_stride, size_t src_image_bytes,
                  int n_images, cudaStream_t s) {
    dim3 grid((g.lv[0].w + 1023) / 1024, g.lv[0].h, n_images);
    repack_kernel<<<grid, 256, 0, s>>>(stage, src_stride, src_image_bytes, ws.pyr, g.lv[0], g.pyr_image_stride);
    return 1;
}

int launch_pyramid(const Geometry& g, const Workspace& ws, int n_images, cudaStream_t s) {
    int launches = 0;
    pad_level0_kernel<<<dim3((g.lv[0].h + 255) / 256, n_images), 256, 0, s>>>(ws.pyr, g.lv[0], g.pyr_image_stride);
    launches++;
    for (int l = 1; l < g.nlevels; l++) {
        const LevelGeom& d = g.lv[l];
        dim3 block(32, 8), grid((d.w + 8 + 127) / 128, (d.h + 8 * PYR_ROWS - 1) / (8 * PYR_ROWS), n_images);
        pyr_resize_kernel<<<grid, block, 0, s>>>(ws.pyr, ws.pyr, ws.tabs, g.lv[l - 1], d, g.pyr_image_stride);
        launches++;
    }
    return launches;
}

}  // namespace borb
