// Standalone probe (dev tooling): TMA tile load variants, one per process (argv[1] = variant).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define WAIT(bar) asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory")
template <int RANK, bool GLOBAL_DESC>
__global__ void probe(const __grid_constant__ CUtensorMap one, const CUtensorMap* gdesc, int x, int y, int z, unsigned bytes, unsigned* out) {
    __shared__ __align__(128) unsigned char tile[66 * 256];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap* d = GLOBAL_DESC ? gdesc : &one;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(d)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(d)), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
    WAIT(bar);
    __syncthreads();
    unsigned s = 0;
    for (int i = threadIdx.x; i < (int)bytes; i += blockDim.x) s += tile[i] * (unsigned)(i % 251 + 1);
    atomicAdd(out, s);
}
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 1242, H = 375, P = 1280, N = 4;
    size_t stride = (size_t)P * H;
    std::vector<unsigned char> h(stride * N);
    for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)((i * 2654435761u) >> 13);
    unsigned char* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    // variants: rank, boxW, boxH, global-desc, l2promo
    struct V { int rank, bw, bh, gdesc, promo; } vs[] = {
        {3, 144, 38, 0, 2}, {2, 144, 38, 0, 2}, {2, 128, 32, 0, 2}, {3, 144, 38, 1, 2}, {3, 144, 38, 0, 0}, {2, 128, 32, 1, 0}, {2, 256, 32, 0, 0}, {2, 64, 32, 0, 0}, {2, 16, 8, 0, 0}};
    V v = vs[variant];
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[3] = {W, H, N}; cuuint64_t strides[2] = {P, stride}; cuuint32_t box[3] = {(cuuint32_t)v.bw, (cuuint32_t)v.bh, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, v.rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)v.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d rank %d box %dx%d gdesc %d promo %d: encode %d; ", variant, v.rank, v.bw, v.bh, v.gdesc, v.promo, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    CUtensorMap* gd; cudaMalloc(&gd, sizeof(tm)); cudaMemcpy(gd, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    unsigned* out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
    const int x = 16, y = 16, z = v.rank == 3 ? 2 : 0;
    unsigned bytes = v.bw * v.bh, want = 0;
    for (unsigned i = 0; i < bytes; i++) { int rr = i / v.bw, c = i % v.bw; int gx = x + c, gy = y + rr; unsigned val = (gx < W && gy < H) ? h[z * stride + (size_t)gy * P + gx] : 0; want += val * (unsigned)(i % 251 + 1); }
    if (v.rank == 3 && !v.gdesc) probe<3, false><<<1, 256>>>(tm, gd, x, y, z, bytes, out);
    if (v.rank == 3 && v.gdesc) probe<3, true><<<1, 256>>>(tm, gd, x, y, z, bytes, out);
    if (v.rank == 2 && !v.gdesc) probe<2, false><<<1, 256>>>(tm, gd, x, y, z, bytes, out);
    if (v.rank == 2 && v.gdesc) probe<2, true><<<1, 256>>>(tm, gd, x, y, z, bytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned got = 0; cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
    printf("%s got %u want %u %s\n", cudaGetErrorString(e), got, want, (e == cudaSuccess && got == want) ? "OK" : "FAIL");
    return 0;
}
