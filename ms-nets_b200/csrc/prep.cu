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

}  // namespace msn
