// dev tooling: which pipe does each packed op use?  Mix op X with LOP3 (alu pipe) and with IMAD (fma pipe):
// if X+Y reaches ~125 thread-ops/clk/SM the two sit on different pipes, ~63 means the same pipe.
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 4096
__device__ __forceinline__ unsigned opx(int OP, unsigned a, unsigned b, unsigned c) {
    switch (OP) {
        case 0: return __vimin3_s16x2(a, b, c);
        case 1: return (a & b) ^ c;                 // LOP3
        case 2: return __vabsdiffu4(a, b);
        case 3: return __byte_perm(a, b, 0x4321);
        case 4: return __dp4a(a, b, c);
        case 5: return __vadd2(a, b);
        case 6: return a * b + c;                   // IMAD
        case 7: return __vmaxs2(a, b);
        case 8: return __popc(a ^ b);               // POPC
        case 9: return (a >> 3) | (b << 5);         // SHF / LOP
        default: return a;
    }
}
template <int X, int Y>
__global__ void k(unsigned* out, unsigned seed) {
    unsigned a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u, c = a + 0x7f4a7c15u, d = b * 3u;
    unsigned e = a + 1, f = b + 2, g = c + 3, h = d + 4;
#pragma unroll 16
    for (int i = 0; i < ITERS; i++) {
        a = opx(X, a, b, c); e = opx(Y, e, f, g);
        b = opx(X, b, c, d); f = opx(Y, f, g, h);
        c = opx(X, c, d, a); g = opx(Y, g, h, e);
        d = opx(X, d, a, b); h = opx(Y, h, e, f);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
template <int X, int Y> void run(const char* name) {
    unsigned* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<X, Y><<<148 * 8, 256>>>(out, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<X, Y><<<148 * 8, 256>>>(out, 2); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * (double)ITERS * 8;
    printf("%-22s %.3f ms  ~%.1f thread-ops/clk/SM\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    run<1, 6>("lop3+imad"); run<0, 1>("vimnmx3+lop3"); run<0, 6>("vimnmx3+imad"); run<7, 6>("vimnmx+imad");
    run<2, 1>("vabsdiff4+lop3"); run<2, 6>("vabsdiff4+imad"); run<3, 1>("prmt+lop3"); run<3, 6>("prmt+imad");
    run<4, 1>("dp4a+lop3"); run<4, 6>("dp4a+imad"); run<5, 1>("vadd2+lop3"); run<5, 6>("vadd2+imad");
    run<8, 1>("popc+lop3"); run<8, 6>("popc+imad"); run<8, 8>("popc"); run<9, 6>("shf+imad"); run<9, 1>("shf+lop3");
    return 0;
}
