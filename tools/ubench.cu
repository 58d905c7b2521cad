// dev tooling: instruction throughput micro-benchmark (ops per clock per SM) for the packed-SIMD candidates.
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 4096
template <int OP>
__global__ void k(unsigned* out, unsigned seed) {
    unsigned a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u, c = a + 0x7f4a7c15u, d = b * 3u;
    unsigned e = a + 1, f = b + 2, g = c + 3, h = d + 4;
#pragma unroll 16
    for (int i = 0; i < ITERS; i++) {
        if (OP == 0) { a = __vimin3_s16x2(a, b, c); b = __vimin3_s16x2(b, c, d); c = __vimin3_s16x2(c, d, a); d = __vimin3_s16x2(d, a, b);
                       e = __vimin3_s16x2(e, f, g); f = __vimin3_s16x2(f, g, h); g = __vimin3_s16x2(g, h, e); h = __vimin3_s16x2(h, e, f); }
        if (OP == 1) { a = (a & b) ^ c; b = (b & c) ^ d; c = (c | d) ^ a; d = (d & a) ^ b; e = (e & f) ^ g; f = (f & g) ^ h; g = (g | h) ^ e; h = (h & e) ^ f; }
        if (OP == 2) { a = __vabsdiffu4(a, b); b = __vabsdiffu4(b, c); c = __vabsdiffu4(c, d); d = __vabsdiffu4(d, a); e = __vabsdiffu4(e, f); f = __vabsdiffu4(f, g); g = __vabsdiffu4(g, h); h = __vabsdiffu4(h, e); }
        if (OP == 3) { a = __byte_perm(a, b, 0x4321); b = __byte_perm(b, c, 0x5432); c = __byte_perm(c, d, 0x6543); d = __byte_perm(d, a, 0x4321); e = __byte_perm(e, f, 0x4321); f = __byte_perm(f, g, 0x5432); g = __byte_perm(g, h, 0x6543); h = __byte_perm(h, e, 0x4321); }
        if (OP == 4) { a = __dp4a(a, b, c); b = __dp4a(b, c, d); c = __dp4a(c, d, a); d = __dp4a(d, a, b); e = __dp4a(e, f, g); f = __dp4a(f, g, h); g = __dp4a(g, h, e); h = __dp4a(h, e, f); }
        if (OP == 5) { a = __vadd2(a, b); b = __vadd2(b, c); c = __vadd2(c, d); d = __vadd2(d, a); e = __vadd2(e, f); f = __vadd2(f, g); g = __vadd2(g, h); h = __vadd2(h, e); }
        if (OP == 6) { a = a + b; b = b + c; c = c + d; d = d + a; e = e + f; f = f + g; g = g + h; h = h + e; }
        if (OP == 7) { a = min((int)a, (int)b); b = max((int)b, (int)c); c = min((int)c, (int)d); d = max((int)d, (int)a); e = min((int)e, (int)f); f = max((int)f, (int)g); g = min((int)g, (int)h); h = max((int)h, (int)e); }
        if (OP == 8) { a = a * b + c; b = b * c + d; c = c * d + a; d = d * a + b; e = e * f + g; f = f * g + h; g = g * h + e; h = h * e + f; }
        if (OP == 9) { a = __vmaxs2(a, b); b = __vmins2(b, c); c = __vmaxs2(c, d); d = __vmins2(d, a); e = __vmaxs2(e, f); f = __vmins2(f, g); g = __vmaxs2(g, h); h = __vmins2(h, e); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
template <int OP> void run(const char* name) {
    unsigned* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 8, 256>>>(out, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<OP><<<148 * 8, 256>>>(out, 2); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * (double)ITERS * 8;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-18s %.3f ms  %.1f Gop/s  ~%.1f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
int main() {
    run<0>("vimin3_s16x2"); run<9>("vimnmx_s16x2"); run<1>("lop3"); run<2>("vabsdiff4"); run<3>("prmt"); run<4>("dp4a"); run<5>("vadd2"); run<6>("iadd"); run<7>("imnmx"); run<8>("imad");
    return 0;
}
