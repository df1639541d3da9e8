// TEST INFRASTRUCTURE ONLY -- exhaustive check of the glibc expf restatement that feature_math.cuh (expf_glibc) follows:
// every negative float down to -104 against the libm of this image.  use_fma = 1 is the sequence x86-64 glibc runs on
// FMA hardware (e_expf-fma: every multiply-add of the source contracted).  gcc -O2 -ffp-contract=off check_expf.c -lm
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static const uint64_t T[32] = {
0x3ff0000000000000, 0x3fefd9b0d3158574, 0x3fefb5586cf9890f, 0x3fef9301d0125b51,
0x3fef72b83c7d517b, 0x3fef54873168b9aa, 0x3fef387a6e756238, 0x3fef1e9df51fdee1,
0x3fef06fe0a31b715, 0x3feef1a7373aa9cb, 0x3feedea64c123422, 0x3feece086061892d,
0x3feebfdad5362a27, 0x3feeb42b569d4f82, 0x3feeab07dd485429, 0x3feea47eb03a5585,
0x3feea09e667f3bcd, 0x3fee9f75e8ec5f74, 0x3feea11473eb0187, 0x3feea589994cce13,
0x3feeace5422aa0db, 0x3feeb737b0cdc5e5, 0x3feec49182a3f090, 0x3feed503b23e255d,
0x3feee89f995ad3ad, 0x3feeff76f2fb5e47, 0x3fef199bdd85529c, 0x3fef3720dcef9069,
0x3fef5818dcfba487, 0x3fef7c97337b9b5f, 0x3fefa4afa2a490da, 0x3fefd0765b6e4540};
static inline uint64_t asu(double d){uint64_t u; memcpy(&u,&d,8); return u;}
static inline double asd(uint64_t u){double d; memcpy(&d,&u,8); return d;}
static float my_expf(float x, int use_fma){
  const double N = 32.0;
  const double InvLn2N = 0x1.71547652b82fep+0 * N, SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5/N/N/N, C1 = 0x1.ebfce50fac4f3p-3/N/N, C2 = 0x1.62e42ff0c52d6p-1/N;
  if (x < -0x1.9fe368p6f) return 0.0f;
  double xd = x, z = InvLn2N * xd;
  volatile double kdv = use_fma ? fma(InvLn2N, xd, SHIFT) : z + SHIFT; double kd = kdv;
  uint64_t ki = asu(kd); kd -= SHIFT;
  double r = use_fma ? fma(InvLn2N, xd, -kd) : z - kd;
  uint64_t t = T[ki % 32]; t += ki << (52 - 5);
  double s = asd(t), y, r2 = r*r;
  if (use_fma) { z = fma(C0, r, C1); y = fma(C2, r, 1.0); y = fma(z, r2, y); }
  else { volatile double a = C0*r; z = a + C1; volatile double b = C2*r; y = b + 1.0; volatile double c = z*r2; y = c + y; }
  y = y * s;
  return (float)y;
}
int main(){
  long bad0=0,bad1=0,n=0;
  // every float in [-104, -0.0]: bit patterns from 0x80000000 .. asuint(-104)
  float lim=-104.0f; uint32_t ulim; memcpy(&ulim,&lim,4);
  for (uint32_t u=0x80000000u; u<=ulim; ++u){ float x; memcpy(&x,&u,4);
    float g=expf(x); float a=my_expf(x,0), b=my_expf(x,1);
    uint32_t ug,ua,ub; memcpy(&ug,&g,4); memcpy(&ua,&a,4); memcpy(&ub,&b,4);
    if (ug!=ua) printf("x=%a (%.9g) glibc %a mine %a\n", x, x, g, a); bad0 += (ug!=ua); bad1 += (ug!=ub); ++n; }
  printf("checked %ld negative floats: mismatches no-fma %ld, fma %ld\n", n, bad0, bad1);
  return 0;
}
