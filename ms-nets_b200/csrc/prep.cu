// prep.cu -- per-pixel statistics every matcher needs before the H x W x D sweep.
//
//   census_transform : wsize x wsize census bits, packed 32 per word
//                      (reference builds them as 16-bit SSE lanes, matchers.cpp:276-305)
//   window_mean      : ZSAD window means (matchers.cpp:469-486)
//   ncc_stats        : NCC window sum A and C = 1/sqrt(w^2*B - A^2) in fp64
//                      (matchers.cpp:125-150)
//   sobel            : horizontal 3x3 Sobel (matchers.cpp:515-554)
//
// All four are O(H*W) and touch ~0.5 MB images: they are launch-latency sized,
// not bandwidth sized.  The fused path (ms_fused.cu) does the same work for both
// images of every pair in one launch; these standalone kernels serve the
// drop-in per-function API.
#include "common.cuh"

namespace msn {

// One thread per pixel.  Bit k = a*wsize+b of pixel (y,x) is [I(y,x) < I(i+a, j+b)]
// with window origin (i,j) = (y-wc, x-wc); pixels whose window leaves the valid
// origin range get all-zero descriptors (never read by the cost kernel).
__global__ void census_transform_kernel(const uint8_t* __restrict__ img, int H, int W, int wsize, int nw,
                                        uint32_t* __restrict__ desc) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const Win win(wsize);
  uint32_t words[kMaxCensusWords];
#pragma unroll
  for (int k = 0; k < kMaxCensusWords; ++k) words[k] = 0u;
  if (win.row_ok(y, H) && win.col_ok(x, W)) {
    const int c = img[(size_t)y * W + x];
    const uint8_t* org = img + (size_t)(y - win.wc) * W + (x - win.wc);
    int bit = 0;
    for (int a = 0; a < wsize; ++a) {
      for (int b = 0; b < wsize; ++b, ++bit) {
        const uint32_t v = (c < (int)org[(size_t)a * W + b]) ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < kMaxCensusWords; ++k)
          if (k == (bit >> 5)) words[k] |= v << (bit & 31);
      }
    }
  }
  uint32_t* o = desc + ((size_t)y * W + x) * nw;
#pragma unroll
  for (int k = 0; k < kMaxCensusWords; ++k)
    if (k < nw) o[k] = words[k];
}

int launch_census_transform(const uint8_t* img, int H, int W, int wsize, uint32_t* desc, cudaStream_t s) {
  const int nw = (wsize * wsize + 31) / 32;
  dim3 block(128), grid(div_up(W, 128), H);
  census_transform_kernel<<<grid, block, 0, s>>>(img, H, W, wsize, nw, desc);
  MSN_LAUNCH_OK();
  return 0;
}

// mean = (float)(sum of taps) / (float)(w*w): the tap sum is an exact integer
// (<= 255*w*w < 2^24), the division is one IEEE fp32 division (matchers.cpp:482).
__global__ void window_mean_kernel(const uint8_t* __restrict__ img, int H, int W, int wsize,
                                   float* __restrict__ mean) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const Win win(wsize);
  float m = 0.f;
  if (win.row_ok(y, H) && win.col_ok(x, W)) {
    const uint8_t* org = img + (size_t)(y - win.wc) * W + (x - win.wc);
    int sum = 0;
    for (int a = 0; a < wsize; ++a)
      for (int b = 0; b < wsize; ++b) sum += org[(size_t)a * W + b];
    m = __fdiv_rn((float)sum, (float)(wsize * wsize));
  }
  mean[(size_t)y * W + x] = m;
}

int launch_window_mean(const uint8_t* img, int H, int W, int wsize, float* mean, cudaStream_t s) {
  dim3 block(128), grid(div_up(W, 128), H);
  window_mean_kernel<<<grid, block, 0, s>>>(img, H, W, wsize, mean);
  MSN_LAUNCH_OK();
  return 0;
}

// A = sum I, B = sum I^2 (exact integers); C = 1/sqrt(w^2*B - A*A) in fp64 with
// IEEE sqrt and division, exactly the reference's expression (matchers.cpp:146).
// A flat window gives sqrt(0) -> C = +inf, which the cost kernel maps to cost 1.
__global__ void ncc_stats_kernel(const uint8_t* __restrict__ img, int H, int W, int wsize,
                                 unsigned long long* __restrict__ A, double* __restrict__ C) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const Win win(wsize);
  unsigned long long a_sum = 0ull;
  double c = 0.0;
  if (win.row_ok(y, H) && win.col_ok(x, W)) {
    const uint8_t* org = img + (size_t)(y - win.wc) * W + (x - win.wc);
    unsigned long long b_sum = 0ull;
    for (int a = 0; a < wsize; ++a)
      for (int b = 0; b < wsize; ++b) {
        const unsigned v = org[(size_t)a * W + b];
        a_sum += v;
        b_sum += v * v;
      }
    const double var = __dsub_rn((double)((unsigned long long)(wsize * wsize) * b_sum),
                                 __dmul_rn((double)a_sum, (double)a_sum));
    c = __ddiv_rn(1.0, __dsqrt_rn(var));
  }
  A[(size_t)y * W + x] = a_sum;
  C[(size_t)y * W + x] = c;
}

int launch_ncc_stats(const uint8_t* img, int H, int W, int wsize, unsigned long long* A, double* C,
                     cudaStream_t s) {
  dim3 block(128), grid(div_up(W, 128), H);
  ncc_stats_kernel<<<grid, block, 0, s>>>(img, H, W, wsize, A, C);
  MSN_LAUNCH_OK();
  return 0;
}

// Gx = [-1 0 1; -2 0 2; -1 0 1] in integer arithmetic, written at (i+1, j+1)
// for i < H-3, j < W-3; zero elsewhere (matchers.cpp:527,538-547).
__global__ void sobel_kernel(const uint8_t* __restrict__ img, int H, int W, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  float v = 0.f;
  if (y >= 1 && y < H - 2 && x >= 1 && x < W - 2) {
    const uint8_t* p = img + (size_t)(y - 1) * W + (x - 1);
    const int g = ((int)p[2] - (int)p[0]) + 2 * ((int)p[W + 2] - (int)p[W]) +
                  ((int)p[2 * W + 2] - (int)p[2 * W]);
    v = (float)g;
  }
  out[(size_t)y * W + x] = v;
}

int launch_sobel(const uint8_t* img, int H, int W, float* out, cudaStream_t s) {
  dim3 block(128), grid(div_up(W, 128), H);
  sobel_kernel<<<grid, block, 0, s>>>(img, H, W, out);
  MSN_LAUNCH_OK();
  return 0;
}

__global__ void fill_kernel(float* __restrict__ p, size_t n, float v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st_stream(p + i, v);
}

int launch_fill(float* p, size_t n, float v, cudaStream_t s) {
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  fill_kernel<<<blocks, 256, 0, s>>>(p, n, v);
  MSN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------ anti-aliased rescale --
// down_sampling_input (cbmv_generator.py:465-482): uint8 -> float32 / 255 -> skimage.transform.rescale(
// anti_aliasing=True, order 1, mode='constant') -> * 255 -> uint8 (truncation).  skimage >= 0.19 runs
// scipy.ndimage.gaussian_filter (sigma = (factor - 1) / 2 per axis, zero boundary) and scipy.ndimage.zoom
// (order 1, grid mode, zero boundary), then clips to the input's range.  The truncation to uint8 makes the
// last bit of the float pipeline visible, so scipy's arithmetic is replayed operation by operation
// (oracle/ms_oracle.py rescale_antialiased_replay): every 1-D correlation is accumulated in fp64 in
// ni_filters.c's order -- centre tap, then the pairs from the farthest inwards, pair added before it is
// weighted -- and rounded to fp32 once per axis; the bilinear taps are (pixel * w_row) * w_col in fp64,
// rows outer.  Weights and zoom factors come from the host (NumPy evaluates them as scipy does).
constexpr int kRescaleMaxR = 16;
struct RescaleW {
  double w[2 * kRescaleMaxR + 1];
  int r;   // radius; < 0: no filtering along this axis
};

// pass 1: per-image min / max of the uint8 input and the vertical (axis 0) correlation -> tmp float [N][H][W]
__global__ void __launch_bounds__(256)
rescale_rows_kernel(const uint8_t* __restrict__ in, int H, int W, RescaleW wr, float* __restrict__ tmp,
                    int* __restrict__ minmax) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
  int lo = 255, hi = 0;
  if (x < W) {
    const uint8_t* img = in + (size_t)n * H * W;
    const int c = img[(size_t)y * W + x];
    lo = hi = c;
    float v = __fdiv_rn((float)c, 255.0f);
    if (wr.r >= 0) {
      double t = __dmul_rn((double)v, wr.w[wr.r]);
      for (int j = -wr.r; j < 0; ++j) {
        const int ya = y + j, yb = y - j;
        const double a = (ya >= 0) ? (double)__fdiv_rn((float)img[(size_t)ya * W + x], 255.0f) : 0.0;
        const double b = (yb < H) ? (double)__fdiv_rn((float)img[(size_t)yb * W + x], 255.0f) : 0.0;
        t = __dadd_rn(t, __dmul_rn(__dadd_rn(a, b), wr.w[wr.r + j]));
      }
      v = (float)t;
    }
    tmp[((size_t)n * H + y) * W + x] = v;
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(minmax + 2 * n, lo);
    atomicMax(minmax + 2 * n + 1, hi);
  }
}

// horizontal (axis 1) correlation of tmp at (y, x); zero outside the grid
__device__ __forceinline__ double rescale_col_tap(const float* __restrict__ row, int x, int W, const RescaleW& wc) {
  if (wc.r < 0) return (double)row[x];
  double t = __dmul_rn((double)row[x], wc.w[wc.r]);
  for (int j = -wc.r; j < 0; ++j) {
    const int xa = x + j, xb = x - j;
    const double a = (xa >= 0) ? (double)row[xa] : 0.0;
    const double b = (xb < W) ? (double)row[xb] : 0.0;
    t = __dadd_rn(t, __dmul_rn(__dadd_rn(a, b), wc.w[wc.r + j]));
  }
  return (double)(float)t;   // scipy stores the filtered image as float32
}

// pass 2: horizontal correlation at the four bilinear taps, zoom, clip, * 255, truncate
__global__ void __launch_bounds__(256)
rescale_zoom_kernel(const float* __restrict__ tmp, int H, int W, int oh, int ow, RescaleW wc, double zr, double zc,
                    const int* __restrict__ minmax, uint8_t* __restrict__ out) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y, n = blockIdx.z;
  if (ox >= ow) return;
  const double cr = __dsub_rn(__dmul_rn(__dadd_rn((double)oy, 0.5), zr), 0.5);
  const double cc = __dsub_rn(__dmul_rn(__dadd_rn((double)ox, 0.5), zc), 0.5);
  const double fr = floor(cr), fc = floor(cc);
  const double wr1 = __dsub_rn(cr, fr), wc1 = __dsub_rn(cc, fc);
  const double wrow[2] = {__dsub_rn(1.0, wr1), wr1}, wcol[2] = {__dsub_rn(1.0, wc1), wc1};
  const int r0 = (int)fr, c0 = (int)fc;
  const float* img = tmp + (size_t)n * H * W;
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int y = r0 + i, x = c0 + j;
      const double v = (y >= 0 && y < H && x >= 0 && x < W) ? rescale_col_tap(img + (size_t)y * W, x, W, wc) : 0.0;
      t = __dadd_rn(t, __dmul_rn(__dmul_rn(v, wrow[i]), wcol[j]));
    }
  float z = (float)t;
  const float lo = __fdiv_rn((float)minmax[2 * n], 255.0f), hi = __fdiv_rn((float)minmax[2 * n + 1], 255.0f);
  const bool keep_zero = (lo > 0.f) && (z == 0.f);        // skimage _clip_warp_output: cval outside the range survives
  z = fminf(fmaxf(z, lo), hi);
  if (keep_zero) z = 0.f;
  out[((size_t)n * oh + oy) * ow + ox] = (uint8_t)(int)__fmul_rn(z, 255.0f);   // astype(np.uint8): truncation
}

__global__ void rescale_init_minmax(int* mm, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { mm[2 * i] = 255; mm[2 * i + 1] = 0; }
}

int launch_rescale(const uint8_t* in, int N, int H, int W, int oh, int ow, const double* w_rows, int r_rows,
                   const double* w_cols, int r_cols, double zoom_rows, double zoom_cols, uint8_t* out, float* tmp,
                   int* minmax, cudaStream_t s) {
  MSN_REQUIRE(r_rows <= kRescaleMaxR && r_cols <= kRescaleMaxR, "rescale: filter radius above %d", kRescaleMaxR);
  MSN_REQUIRE(N <= 65535 && H <= 65535 && oh <= 65535, "rescale: image or batch too large");
  RescaleW wr, wc;
  memset(&wr, 0, sizeof(wr)); memset(&wc, 0, sizeof(wc));
  wr.r = r_rows; wc.r = r_cols;
  if (r_rows >= 0) memcpy(wr.w, w_rows, (2 * r_rows + 1) * sizeof(double));
  if (r_cols >= 0) memcpy(wc.w, w_cols, (2 * r_cols + 1) * sizeof(double));
  rescale_init_minmax<<<div_up(N, 128), 128, 0, s>>>(minmax, N);
  MSN_LAUNCH_OK();
  rescale_rows_kernel<<<dim3(div_up(W, 256), H, N), 256, 0, s>>>(in, H, W, wr, tmp, minmax);
  MSN_LAUNCH_OK();
  rescale_zoom_kernel<<<dim3(div_up(ow, 256), oh, N), 256, 0, s>>>(tmp, H, W, oh, ow, wc, zoom_rows, zoom_cols, minmax, out);
  MSN_LAUNCH_OK();
  return 0;
}

}  // namespace msn
