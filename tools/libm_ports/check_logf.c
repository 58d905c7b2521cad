// Exhaustive check of the glibc logf port used by orb_slam2_b200/csrc/k_match.cu (glibc_logf) against the host libm.
// build: gcc -O2 -ffp-contract=off -o check_logf check_logf.c -lm ; run time ~30 s; expected: mismatches nofma 0, fma 0
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static const struct { double invc, logc; } T[16] = {
  { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 },
  { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
  { 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2 },
  { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
  { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 },
  { 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3 },
  { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 },
  { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
  { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 },
  { 0x1p+0, 0x0p+0 },
  { 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 },
  { 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4 },
  { 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 },
  { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3 },
  { 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 },
  { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 },
};
static const double Ln2 = 0x1.62e42fefa39efp-1;
static const double A[3] = { -0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2 };
static inline uint32_t asuint(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float asfloat(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static float my_logf(float x, int use_fma) {
  uint32_t ix = asuint(x);
  if (ix == 0x3f800000) return 0;
  uint32_t tmp = ix - 0x3f330000;
  int i = (tmp >> (23 - 4)) % 16;
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0x1ffu << 23);
  double invc = T[i].invc, logc = T[i].logc;
  double z = (double)asfloat(iz);
  double r, y0, r2, y;
  if (use_fma) {
    r = fma(z, invc, -1.0);
    y0 = fma((double)k, Ln2, logc);
    r2 = r * r;
    y = fma(A[1], r, A[2]);
    y = fma(A[0], r2, y);
    y = fma(y, r2, (y0 + r));
  } else {
    r = z * invc - 1;
    y0 = logc + (double)k * Ln2;
    r2 = r * r;
    y = A[1] * r + A[2];
    y = A[0] * r2 + y;
    y = y * r2 + (y0 + r);
  }
  return (float)y;
}
int main(int argc, char** argv) {
  const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;   /* 1 = every positive normal float */
  long bad0 = 0, bad1 = 0, n = 0;
  for (uint64_t uu = 0x00800000u; uu < 0x7f800000u; uu += stride) {
    const uint32_t u = (uint32_t)uu;   // all positive normal floats
    float x = asfloat(u);
    float ref = logf(x);
    if (asuint(my_logf(x, 0)) != asuint(ref)) { if (bad0 < 3) printf("nofma mismatch %a: %a vs %a\n", x, my_logf(x,0), ref); bad0++; }
    if (asuint(my_logf(x, 1)) != asuint(ref)) { if (bad1 < 3) printf("fma mismatch %a: %a vs %a\n", x, my_logf(x,1), ref); bad1++; }
    n++;
  }
  printf("n=%ld mismatches: nofma %ld, fma %ld\n", n, bad0, bad1);
  return 0;
}
