// features.cu -- extract_features_left / extract_features_lr from four raw cost
// volumes (cbmv_generator.py:258-308, :84-254), generic in D and window sizes.
//
// Input: census, ncc, sobel(sadsob), sad(zsad) as [h][w][D] (D innermost, the layout
// get_costs hands over).  Output: [8 or 16][D][h][w] float32, W innermost -- what
// the reference reaches only after a float64 scratch, eight channel assignments
// and a strided transpose+cast (cbmv_generator.py:281,307-308).  One CTA owns 32
// consecutive pixels x all D of one matcher: coalesced D-rows in, transposed
// through shared memory, per-pixel min / AML denominator, coalesced W-rows out.
#include "common.cuh"
#include "feature_math.cuh"

namespace msn {

constexpr int kFeatPix = 32;
constexpr int kFeatWarps = 8;

// grid: (ceil(h*w/32), nviews*4); dynamic smem: D*33 floats.
__global__ void __launch_bounds__(kFeatWarps * 32)
features_from_costs_kernel(const float* __restrict__ c0, const float* __restrict__ c1,
                           const float* __restrict__ c2, const float* __restrict__ c3, int h, int w, int D,
                           float k0, float k1, float k2, float k3, float* __restrict__ out) {
  extern __shared__ float tile[];  // [D][33]
  __shared__ float s_min[kFeatPix];
  __shared__ float s_inv[kFeatPix];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.y & 3;          // matcher 0..3
  const bool right = (blockIdx.y >> 2);  // right-view channels 8..15
  const float* c = (m == 0) ? c0 : (m == 1) ? c1 : (m == 2) ? c2 : c3;
  const float k = (m == 0) ? k0 : (m == 1) ? k1 : (m == 2) ? k2 : k3;
  const long long n = (long long)h * w;
  const long long p0 = (long long)blockIdx.x * kFeatPix;
  const float first = c[0];

  // load + per-pixel minimum (warp-shuffle reduction over D)
  for (int pl = warp; pl < kFeatPix; pl += kFeatWarps) {
    const long long p = p0 + pl;
    float mn = kFill;
    if (p < n) {
      const int x = (int)(p % w);
      for (int d = lane; d < D; d += 32) {
        float v;
        if (!right)
          v = c[p * D + d];
        else
          v = (x < w - d) ? c[(p + d) * D + d] : first;  // featextract.cpp:151,159-161
        tile[d * 33 + pl] = v;
        mn = fminf(mn, v);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if (lane == 0) s_min[pl] = mn;
  }
  __syncthreads();
  const long long p = p0 + lane;
  const bool live = p < n;
  const float mn = s_min[lane];
  // normalised plane out, exponentials into the tile
  if (live) {
    float* o_norm = out + ((size_t)((right ? 8 : 0) + m) * D) * n + p;
    for (int d = warp; d < D; d += kFeatWarps) {
      const float v = tile[d * 33 + lane];
      st_stream(o_norm + (size_t)d * n, normalise_cost(v, m));
      tile[d * 33 + lane] = aml_e(v, mn, k);
    }
  }
  __syncthreads();
  // denominator in the reference's order: sequential fp32 over d (featextract.cpp:444-447)
  if (warp == 0) {
    float den = 0.f;
    if (live) {
#pragma unroll 8
      for (int d = 0; d < D; ++d) den = __fadd_rn(den, tile[d * 33 + lane]);
    }
    s_inv[lane] = aml_row_scale(den, mn != kFill, k);
  }
  __syncthreads();
  if (!live) return;
  const float inv = s_inv[lane];
  float* o_aml = out + ((size_t)((right ? 8 : 0) + 4 + m) * D) * n + p;
  for (int d = warp; d < D; d += kFeatWarps) st_stream(o_aml + (size_t)d * n, aml_apply(tile[d * 33 + lane], inv, k));
}

int launch_features_from_costs(const float* census, const float* ncc, const float* sobel, const float* sad,
                               int h, int w, int D, float cens_sigma, float ncc_sigma, float sad_sigma, int lr,
                               float* out, cudaStream_t s) {
  const long long n = (long long)h * w;
  if (n == 0 || D == 0) return 0;
  const size_t smem = (size_t)D * 33 * sizeof(float);
  MSN_REQUIRE(smem <= 200 * 1024, "features_from_costs: D=%d needs %zu B of shared memory (max 200 KiB)", D, smem);
  MSN_CUDA_OK(cudaFuncSetAttribute(features_from_costs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  dim3 grid(div_up(n, kFeatPix), lr ? 8 : 4);
  features_from_costs_kernel<<<grid, kFeatWarps * 32, smem, s>>>(
      census, ncc, sobel, sad, h, w, D, aml_scale(cens_sigma), aml_scale(ncc_sigma), aml_scale(sad_sigma),
      aml_scale(sad_sigma), out);
  MSN_LAUNCH_OK();
  return 0;
}

}  // namespace msn
