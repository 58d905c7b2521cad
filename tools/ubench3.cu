// dev tooling: does half2 min/max (HMNMX2) issue on the fma pipe, i.e. can it run beside VIMNMX3.S16x2 (alu pipe)?
// Same harness as ubench2.cu: X+Y reaching ~125 thread-ops/clk/SM = different pipes, ~63 = same pipe.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#define ITERS 4096
__device__ __forceinline__ unsigned h2u(__half2 h) { return *reinterpret_cast<unsigned*>(&h); }
__device__ __forceinline__ __half2 u2h(unsigned u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ unsigned opx(int OP, unsigned a, unsigned b, unsigned c) {
    switch (OP) {
        case 0: return __vimin3_s16x2(a, b, c);
        case 1: return (a & b) ^ c;                 // LOP3
        case 6: return a * b + c;                   // IMAD
        case 10: return h2u(__hmin2(u2h(a), u2h(b)));      // HMNMX2
        case 11: return h2u(__hmax2(__hmin2(u2h(a), u2h(b)), u2h(c)));   // 2 x HMNMX2
        case 12: return h2u(__hadd2(u2h(a), u2h(b)));      // HADD2
        case 13: return h2u(__hfma2(u2h(a), u2h(b), u2h(c)));   // HFMA2
        default: return a;
    }
}
template <int X, int Y>
__global__ void k(unsigned* out, unsigned seed) {
    unsigned a = (threadIdx.x & 1023) | 0x3c003c00u, b = a ^ 0x00010003u, c = a + 0x00020001u, d = b + 5u;
    unsigned e = a + 1, f = b + 2, g = c + 3, h = d + 4;
    a += seed & 1;
#pragma unroll 16
    for (int i = 0; i < ITERS; i++) {
        a = opx(X, a, b, c); e = opx(Y, e, f, g);
        b = opx(X, b, c, d); f = opx(Y, f, g, h);
        c = opx(X, c, d, a); g = opx(Y, g, h, e);
        d = opx(X, d, a, b); h = opx(Y, h, e, f);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
template <int X, int Y> void run(const char* name, double ops_per_iter) {
    unsigned* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<X, Y><<<148 * 8, 256>>>(out, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<X, Y><<<148 * 8, 256>>>(out, 2); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * (double)ITERS * ops_per_iter;
    printf("%-26s %.3f ms  ~%.1f thread-instr/clk/SM\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    run<10, 10>("hmnmx2 alone", 8); run<10, 1>("hmnmx2+lop3", 8); run<10, 6>("hmnmx2+imad", 8); run<10, 0>("hmnmx2+vimnmx3", 8);
    run<12, 1>("hadd2+lop3", 8); run<12, 6>("hadd2+imad", 8); run<13, 0>("hfma2+vimnmx3", 8); run<0, 0>("vimnmx3 alone", 8);
    run<11, 0>("2xhmnmx2+vimnmx3", 12);
    return 0;
}
