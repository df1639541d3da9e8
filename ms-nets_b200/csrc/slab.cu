// slab.cu -- three-phase feature assembly through global memory.
//
// Used (a) by the disparity-slab sharded path: every rank owns
// disparities [d_begin, d_begin+Dn) of the same pair and the AML minimum /
// denominator are all-reduced between the phases; (b) as the generic single-GPU
// path for window sizes or disparity counts the fused kernel (ms_fused.cu) is not
// specialised for.  One thread owns one cropped pixel and walks its Dn costs;
// a warp reads/writes 128 contiguous bytes of every (channel, d) row.
//
//   phase A : raw [Dn][H][W] x4  -> ch0-3 (normalised) , ch4-7 := raw (parked),
//             per-pixel slab minima                      (cbmv_generator.py:283-287)
//   phase B : partial AML denominators from the parked costs and the GLOBAL minima
//   phase C : ch4-7 := e / den in place                  (featextract.cpp:444-453)
// With lr != 0 the right-view channels 8-15 are produced the same way from
// c(y, x+d, d) (featextract.cpp:136-172).
#include "common.cuh"
#include "feature_math.cuh"

namespace msn {

struct RawViews {
  const float* v[4];      // census, ncc, sadsob, zsad : [Dn][H][W] at disparity d_begin
  const float* first[4];  // device address of c[0] per matcher (right view only)
};
struct Scales {
  float k[4];
};

__global__ void __launch_bounds__(256)
slab_phase_a_kernel(RawViews raw, int H, int W, int y0, int x0, int h, int w, int Dn, int d_begin, int lr,
                    float* __restrict__ out, float* __restrict__ mins) {
  const long long n = (long long)h * w;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int y = (int)(p / w), x = (int)(p % w);
  const size_t plane = (size_t)H * W;
  const size_t src = (size_t)(y + y0) * W + (x + x0);
  float firsts[4] = {0.f, 0.f, 0.f, 0.f};
  if (lr) {
#pragma unroll
    for (int m = 0; m < 4; ++m) firsts[m] = *raw.first[m];  // featextract.cpp:151
  }
  float mn[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) mn[i] = kFill;
  for (int dd = 0; dd < Dn; ++dd) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float v = raw.v[m][dd * plane + src];
      st_stream(out + ((size_t)m * Dn + dd) * n + p, normalise_cost(v, m));
      out[((size_t)(4 + m) * Dn + dd) * n + p] = v;
      mn[m] = fminf(mn[m], v);
    }
    if (lr) {
      const int d = d_begin + dd;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float v = (x < w - d) ? raw.v[m][dd * plane + src + d] : firsts[m];
        st_stream(out + ((size_t)(8 + m) * Dn + dd) * n + p, normalise_cost(v, m));
        out[((size_t)(12 + m) * Dn + dd) * n + p] = v;
        mn[4 + m] = fminf(mn[4 + m], v);
      }
    }
  }
  const int nm = lr ? 8 : 4;
  for (int i = 0; i < nm; ++i) mins[(size_t)i * n + p] = mn[i];
}

// Right-view channels from the PARKED left-view raw costs of a 16-channel tensor (phase A form:
// channels 4-7 = raw [Dn][h][w]): right(y,x,d) = left(y,x+d,d) for x < w-d, else c.flat[0]
// (featextract.cpp:136-172).  Writes channels 8-11 (normalised), 12-15 (raw, parked) and the
// right-view minima (planes 4-7 of mins).
__global__ void __launch_bounds__(256)
slab_right_view_kernel(float* __restrict__ out, int h, int w, int Dn, int d_begin, const float* __restrict__ first4,
                       float* __restrict__ mins) {
  const long long n = (long long)h * w;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int x = (int)(p % w);
  float firsts[4], mn[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    firsts[m] = first4 ? first4[m] : out[(size_t)(4 + m) * Dn * n];   // voxel (d=0, y=0, x=0) of the cropped volume
    mn[m] = kFill;
  }
  for (int dd = 0; dd < Dn; ++dd) {
    const int d = d_begin + dd;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float v = (x < w - d) ? out[((size_t)(4 + m) * Dn + dd) * n + p + d] : firsts[m];
      st_stream(out + ((size_t)(8 + m) * Dn + dd) * n + p, normalise_cost(v, m));
      out[((size_t)(12 + m) * Dn + dd) * n + p] = v;
      mn[m] = fminf(mn[m], v);
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) mins[(size_t)(4 + m) * n + p] = mn[m];
}

int launch_slab_right_view(float* out, int h, int w, int Dn, int d_begin, const float* d_first4, float* mins,
                           cudaStream_t s) {
  const long long n = (long long)h * w;
  if (n <= 0 || Dn <= 0) return 0;
  MSN_REQUIRE(d_first4 || d_begin == 0, "slab right view: d_begin > 0 needs the four c[0] values of slab 0");
  slab_right_view_kernel<<<div_up(n, 256), 256, 0, s>>>(out, h, w, Dn, d_begin, d_first4, mins);
  MSN_LAUNCH_OK();
  return 0;
}

// channel that parks / receives the AML of (view, matcher)
__device__ __forceinline__ int aml_channel(int i) { return (i >> 2) * 8 + 4 + (i & 3); }

__global__ void __launch_bounds__(256)
slab_phase_b_kernel(const float* __restrict__ out, const float* __restrict__ gmin, long long n, int Dn, int nm,
                    Scales sc, float* __restrict__ den) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;  // 0..nm-1
  if (p >= n) return;
  const float m = gmin[(size_t)i * n + p];
  const float k = sc.k[i & 3];
  const float* c = out + (size_t)aml_channel(i) * Dn * n + p;
  // sequential fp32 in d order, as the reference sums (featextract.cpp:444-447)
  float s = 0.f;
#pragma unroll 4
  for (int dd = 0; dd < Dn; ++dd) s = __fadd_rn(s, aml_e(c[(size_t)dd * n], m, k));
  den[(size_t)i * n + p] = s;
}

__global__ void __launch_bounds__(256)
slab_phase_c_kernel(float* __restrict__ out, const float* __restrict__ gmin, const float* __restrict__ gden,
                    long long n, int Dn, int nm, Scales sc) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (p >= n) return;
  const float m = gmin[(size_t)i * n + p];
  const float k = sc.k[i & 3];
  const float inv = aml_row_scale(gden[(size_t)i * n + p], m != kFill, k);
  float* c = out + (size_t)aml_channel(i) * Dn * n + p;
#pragma unroll 4
  for (int dd = 0; dd < Dn; ++dd) {
    const float v = c[(size_t)dd * n];
    st_stream(c + (size_t)dd * n, aml_apply(aml_e(v, m, k), inv, k));
  }
}

int launch_slab_phase_a(const float* census, const float* ncc, const float* sob, const float* sad, int H, int W,
                        int y0, int x0, int h, int w, int Dn, int d_begin, int lr, const float* d_first4,
                        float* out, float* mins, cudaStream_t s) {
  const long long n = (long long)h * w;
  if (n <= 0 || Dn <= 0) return 0;
  RawViews rv;
  rv.v[0] = census; rv.v[1] = ncc; rv.v[2] = sob; rv.v[3] = sad;
  for (int m = 0; m < 4; ++m) rv.first[m] = nullptr;
  if (lr) {
    // c.flat[0] of the cropped [h][w][D] volume is voxel (d=0, y0, x0): local when the
    // slab starts at 0, otherwise the caller passes the four values (device memory).
    if (d_first4) {
      for (int m = 0; m < 4; ++m) rv.first[m] = d_first4 + m;
    } else {
      MSN_REQUIRE(d_begin == 0, "slab phase A: lr with d_begin > 0 needs the four c[0] values of slab 0");
      for (int m = 0; m < 4; ++m) rv.first[m] = rv.v[m] + (size_t)y0 * W + x0;
    }
  }
  slab_phase_a_kernel<<<div_up(n, 256), 256, 0, s>>>(rv, H, W, y0, x0, h, w, Dn, d_begin, lr, out, mins);
  MSN_LAUNCH_OK();
  return 0;
}

static Scales make_scales(float cens_sigma, float ncc_sigma, float sad_sigma) {
  Scales sc;
  sc.k[0] = aml_scale(cens_sigma);
  sc.k[1] = aml_scale(ncc_sigma);
  sc.k[2] = aml_scale(sad_sigma);  // sobel channel uses sad_sigma (cbmv_generator.py:298)
  sc.k[3] = aml_scale(sad_sigma);
  return sc;
}

int launch_slab_phase_b(const float* out, const float* gmin, long long n, int Dn, int lr, float cens_sigma,
                        float ncc_sigma, float sad_sigma, float* den, cudaStream_t s) {
  if (n <= 0 || Dn <= 0) return 0;
  const int nm = lr ? 8 : 4;
  dim3 grid(div_up(n, 256), nm);
  slab_phase_b_kernel<<<grid, 256, 0, s>>>(out, gmin, n, Dn, nm, make_scales(cens_sigma, ncc_sigma, sad_sigma), den);
  MSN_LAUNCH_OK();
  return 0;
}

int launch_slab_phase_c(float* out, const float* gmin, const float* gden, long long n, int Dn, int lr,
                        float cens_sigma, float ncc_sigma, float sad_sigma, cudaStream_t s) {
  if (n <= 0 || Dn <= 0) return 0;
  const int nm = lr ? 8 : 4;
  dim3 grid(div_up(n, 256), nm);
  slab_phase_c_kernel<<<grid, 256, 0, s>>>(out, gmin, gden, n, Dn, nm, make_scales(cens_sigma, ncc_sigma, sad_sigma));
  MSN_LAUNCH_OK();
  return 0;
}

}  // namespace msn
