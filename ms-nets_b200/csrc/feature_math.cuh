// feature_math.cuh -- channel normalisation and AML arithmetic shared by every
// kernel that emits feature planes (features.cu, ms_fused.cu, fte.cu).
#pragma once
#include "common.cuh"

namespace msn {

// ch0 clip(census,0,120)/120 ; ch1 (1+clip(ncc,-1,1))/2 ; ch2/ch3 clip(.,0,8192)/8192
// in IEEE fp32 exactly as NumPy evaluates them (cbmv_generator.py:283-287).
// The division by 120 is a true division (not a reciprocal multiply) so channel 0
// is bit-exact; /2 and /8192 are exact power-of-two scalings.  fill -> 1.0.
__device__ __forceinline__ float normalise_cost(float v, int matcher) {
  if (matcher == 0) return __fdiv_rn(fminf(fmaxf(v, 0.f), 120.f), 120.f);
  // One instruction each (FFMA.SAT / FMUL.SAT), bit-identical to the clip-then-scale forms:
  //  (1 + clip(v,-1,1)) * 0.5: for |v| <= 1 the sum 1 + v is rounded once and the halving is exact, so it
  //  equals round(0.5 v + 0.5) = fma(v, 0.5, 0.5); outside [-1,1] both forms give exactly 0 or 1.
  //  clip(v,0,8192) / 8192: the scaling by 2^-13 is exact, so clipping before or after it is the same.
  if (matcher == 1) return __saturatef(__fmaf_rn(v, 0.5f, 0.5f));
  return __saturatef(__fmul_rn(v, 1.0f / 8192.f));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// AML term expf(-(c-m)^2 / sigma) = 2^(-(c-m)^2 * log2(e)/sigma)   (featextract.cpp:444-452).
// Tolerance class: the reference calls glibc expf and sums sequentially; here the
// SFU ex2 (rel. error ~2^-22) is used and the sum is replayed in the same order.  Stated bound, checked in
// tests/: |AML - reference| <= 2e-6 on outputs in [0,1].
__host__ __device__ __forceinline__ float aml_scale(float sigma) { return 1.4426950408889634f / sigma; }
__device__ __forceinline__ float aml_e(float c, float m, float k) {
  const float t = c - m;
  return ex2_approx(-(t * t) * k);
}

}  // namespace msn
