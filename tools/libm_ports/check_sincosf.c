// Exhaustive check of the glibc sinf/cosf port used by orb_slam2_b200/csrc/k_describe.cu (glibc_sincosf) against the host libm,
// for every float in [0, 2*pi].  build: gcc -O2 -ffp-contract=off [-DUSEFMA -mfma] -o check_sincosf check_sincosf.c -lm ; ~1 min;
// expected: 0 mismatches for both functions, with and without -DUSEFMA.
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
typedef struct { double sign[4]; double hpi_inv, hpi, c0,c1,c2,c3,c4,s1,s2,s3; } sincos_t;
static const sincos_t T[2] = {
 {{1.0,-1.0,-1.0,1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
 {{1.0,-1.0,-1.0,1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
#ifdef USEFMA
#define MA(a,b,c) fma((a),(b),(c))
#else
#define MA(a,b,c) ((a)*(b)+(c))
#endif
static inline float poly(double x,double x2,const sincos_t*p,int n){
  if((n&1)==0){ double x3=x*x2; double s1=MA(x2,p->s3,p->s2); double x7=x3*x2; double s=MA(x3,p->s1,x); return (float)MA(x7,s1,s);}
  else { double x4=x2*x2; double c2=MA(x2,p->c4,p->c3); double c1=MA(x2,p->c1,p->c0); double x6=x4*x2; double c=MA(x4,p->c2,c1); return (float)MA(x6,c2,c);}
}
static inline uint32_t top12(float f){uint32_t u;memcpy(&u,&f,4);return (u>>20)&0x7ff;}
static float my_sc(float y,int iscos){
  double x=y; const sincos_t*p=&T[0]; int n;
  if(top12(y)<top12(0x1.921FB6p-1f)){ double x2=x*x; if(top12(y)<top12(0x1p-12f)) return iscos?1.0f:y; return poly(x,x2,p,iscos);}
  double r=x*p->hpi_inv; n=((int32_t)r+0x800000)>>24; x=MA(-(double)n,p->hpi,x);
  double s=p->sign[n&3]; if(n&2)p=&T[1];
  return poly(x*s,x*x,p,n^iscos);
}
int main(){
  uint32_t hi; float twopi=6.2831855f; memcpy(&hi,&twopi,4);
  long n=0, mc=0, ms=0;
  for(uint32_t u=0; u<=hi+100; u+=1){ float x; memcpy(&x,&u,4);
    n++; if(cosf(x)!=my_sc(x,1)) mc++; if(sinf(x)!=my_sc(x,0)) ms++; }
  printf("n=%ld cos mismatches=%ld sin mismatches=%ld\n",n,mc,ms);
}
