// regress.cu -- soft-argmin disparity regression and the 4D volume builders.
//
// soft-argmin replaces F.softmax(out,1) + disparityregression
// (gcnet_3dcnn.py:127-141): the reference makes >= 4 passes over the [N,D,H,W]
// volume and materialises arange(D).repeat(N,1,H,W); here one thread streams the D
// column of 1 or 4 neighbouring pixels once, with an online-softmax (running max,
// rescaled sums) so the volume is read exactly once: 4 B/voxel in, 4 B/pixel out.
// HBM-bound; 8 independent 128-bit loads are kept in flight per thread.
#include "common.cuh"
#include "feature_math.cuh"

namespace msn {

constexpr float kLog2e = 1.4426950408889634f;
constexpr int kSaChunk = 8;

template <int V>
struct Vec;
template <>
struct Vec<1> {
  using T = float;
  __device__ static void get(const T& v, float* f) { f[0] = v; }
};
template <>
struct Vec<4> {
  using T = float4;
  __device__ static void get(const T& v, float* f) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
};

// MODE 0: write disp; MODE 1: write partial (max, sum e, sum d*e) planes [N][3][HW];
// MODE 2: input already holds probabilities -> plain expectation sum_d d*p_d
// (the body of the reference's disparityregression, gcnet_3dcnn.py:136-139).
template <int V, int MODE>
__global__ void __launch_bounds__(256)
soft_argmin_kernel(const float* __restrict__ logits, int N, int D, long long HW, int d_begin,
                   float* __restrict__ out) {
  using VT = typename Vec<V>::T;
  const long long groups = HW / V;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (g >= groups) return;
  const VT* src = reinterpret_cast<const VT*>(logits + (size_t)n * D * HW) + g;
  float m[V], s[V], t[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { m[i] = -INFINITY; s[i] = 0.f; t[i] = 0.f; }
  for (int d0 = 0; d0 < D; d0 += kSaChunk) {
    VT raw[kSaChunk];
#pragma unroll
    for (int j = 0; j < kSaChunk; ++j)
      if (d0 + j < D) raw[j] = __ldcs(src + (size_t)(d0 + j) * groups);
    float x[kSaChunk][V];
#pragma unroll
    for (int j = 0; j < kSaChunk; ++j) Vec<V>::get(raw[j], x[j]);
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < V; ++i)
#pragma unroll
        for (int j = 0; j < kSaChunk; ++j)
          if (d0 + j < D) t[i] = fmaf(x[j][i], (float)(d_begin + d0 + j), t[i]);
      continue;
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float cm = m[i];
#pragma unroll
      for (int j = 0; j < kSaChunk; ++j)
        if (d0 + j < D) cm = fmaxf(cm, x[j][i]);
      const float sc = ex2_approx((m[i] - cm) * kLog2e);  // first chunk: 2^-inf = 0
      const float mlog = cm * kLog2e;
      float ss = s[i] * sc, tt = t[i] * sc;
#pragma unroll
      for (int j = 0; j < kSaChunk; ++j)
        if (d0 + j < D) {
          const float e = ex2_approx(fmaf(x[j][i], kLog2e, -mlog));
          ss += e;
          tt = fmaf(e, (float)(d_begin + d0 + j), tt);
        }
      m[i] = cm; s[i] = ss; t[i] = tt;
    }
  }
  if (MODE == 0 || MODE == 2) {
    float* o = out + (size_t)n * HW + g * V;
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = (MODE == 2) ? t[i] : t[i] / s[i];
  } else {
    float* o = out + (size_t)n * 3 * HW + g * V;
#pragma unroll
    for (int i = 0; i < V; ++i) { o[i] = m[i]; o[HW + i] = s[i]; o[2 * HW + i] = t[i]; }
  }
}

// mode: 0 soft-argmin, 1 slab partials, 2 expectation of given probabilities
int launch_soft_argmin(const float* logits, int N, int D, int H, int W, int d_begin, int mode,
                       float* out, cudaStream_t s) {
  MSN_REQUIRE(N >= 0 && D >= 1 && H >= 0 && W >= 0, "soft_argmin: bad shape N=%d D=%d H=%d W=%d", N, D, H, W);
  const long long HW = (long long)H * W;
  if (N == 0 || HW == 0) return 0;
  MSN_REQUIRE(N <= 65535, "soft_argmin: N=%d too large", N);
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  if (vec) {
    dim3 grid(div_up(HW / 4, 256), N);
    if (mode == 1) soft_argmin_kernel<4, 1><<<grid, 256, 0, s>>>(logits, N, D, HW, d_begin, out);
    else if (mode == 2) soft_argmin_kernel<4, 2><<<grid, 256, 0, s>>>(logits, N, D, HW, d_begin, out);
    else soft_argmin_kernel<4, 0><<<grid, 256, 0, s>>>(logits, N, D, HW, d_begin, out);
  } else {
    dim3 grid(div_up(HW, 256), N);
    if (mode == 1) soft_argmin_kernel<1, 1><<<grid, 256, 0, s>>>(logits, N, D, HW, d_begin, out);
    else if (mode == 2) soft_argmin_kernel<1, 2><<<grid, 256, 0, s>>>(logits, N, D, HW, d_begin, out);
    else soft_argmin_kernel<1, 0><<<grid, 256, 0, s>>>(logits, N, D, HW, d_begin, out);
  }
  MSN_LAUNCH_OK();
  return 0;
}

// Backward of soft-argmin for training (the reference trains through F.softmax + disparityregression,
// gcnet_3dcnn.py:127-141; psmnet_3dcnn.py:149-176):  d disp / d x_d = p_d * (d - disp), so
// grad_x[n][d][p] = grad_out[n][p] * p_d * (d - disp[n][p]).  Two sweeps over the column: (max, sum)
// online, then the gradient; thread = pixel, coalesced along W.
__global__ void __launch_bounds__(256)
soft_argmin_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ disp,
                       const float* __restrict__ gout, int D, long long HW, float* __restrict__ gin) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p >= HW) return;
  const float* x = logits + (size_t)n * D * HW + p;
  float m = -INFINITY, s = 0.f;
  for (int d = 0; d < D; ++d) {
    const float v = x[(size_t)d * HW];
    const float cm = fmaxf(m, v);
    s = s * ex2_approx((m - cm) * kLog2e) + ex2_approx((v - cm) * kLog2e);
    m = cm;
  }
  const float go = gout[(size_t)n * HW + p] / s, dv = disp[(size_t)n * HW + p];
  float* g = gin + (size_t)n * D * HW + p;
  for (int d = 0; d < D; ++d) {
    const float e = ex2_approx((x[(size_t)d * HW] - m) * kLog2e);
    g[(size_t)d * HW] = go * e * ((float)d - dv);
  }
}

int launch_soft_argmin_bwd(const float* logits, const float* disp, const float* gout, int N, int D, int H, int W,
                           float* gin, cudaStream_t s) {
  MSN_REQUIRE(N >= 0 && D >= 1 && H >= 0 && W >= 0, "soft_argmin_backward: bad shape");
  const long long HW = (long long)H * W;
  if (N == 0 || HW == 0) return 0;
  MSN_REQUIRE(N <= 65535, "soft_argmin_backward: N=%d too large", N);
  dim3 grid(div_up(HW, 256), N);
  soft_argmin_bwd_kernel<<<grid, 256, 0, s>>>(logits, disp, gout, D, HW, gin);
  MSN_LAUNCH_OK();
  return 0;
}

// parts: [P][N][3][HW] gathered slab partials -> disp [N][HW]
__global__ void soft_argmin_merge_kernel(const float* __restrict__ parts, int P, long long NHW3, long long HW,
                                         long long total, float* __restrict__ disp) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const long long n = q / HW, p = q % HW;
  const float* base = parts + n * 3 * HW + p;
  float M = -INFINITY;
  for (int r = 0; r < P; ++r) M = fmaxf(M, base[(size_t)r * NHW3]);
  float S = 0.f, T = 0.f;
  for (int r = 0; r < P; ++r) {
    const float* b = base + (size_t)r * NHW3;
    const float sc = ex2_approx((b[0] - M) * kLog2e);
    S = fmaf(b[HW], sc, S);
    T = fmaf(b[2 * HW], sc, T);
  }
  disp[q] = T / S;
}

int launch_soft_argmin_merge(const float* parts, int P, int N, int H, int W, float* disp, cudaStream_t s) {
  const long long HW = (long long)H * W, total = HW * N;
  if (total == 0) return 0;
  MSN_REQUIRE(P >= 1, "soft_argmin_merge: parts must be >= 1");
  soft_argmin_merge_kernel<<<div_up(total, 256), 256, 0, s>>>(parts, P, 3 * total, HW, total, disp);
  MSN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------ 4D volumes --
// concat: vol[n][c][d][y][x]   = x >= d ? fl[n][c][y][x]   : 0      (c <  C)
//         vol[n][C+c][d][y][x] = x >= d ? fr[n][c][y][x-d] : 0
// diff  : vol[n][c][d][y][x]   = x >= d ? fl[n][c][y][x] - fr[n][c][y][x-d] : 0
// Pure data movement, D-fold write amplification: HBM-write bound.  One thread owns 4
// consecutive x of one input channel row and walks ALL disparities: the left quad stays in
// registers, the right quad slides one column per disparity (one scalar load per step), and
// every step emits one (diff) or two (concat) 128-bit streaming stores.  blockIdx.y splits D
// into chunks so small inputs still fill the machine.
template <bool kDiff>
__global__ void __launch_bounds__(256)
shift_volume_kernel(const float* __restrict__ fl, const float* __restrict__ fr, int C, int H, int W, int D,
                    int d_chunk, float* __restrict__ vol) {
  const int Wq = (W + 3) >> 2;
  const long long row_groups = (long long)H * Wq;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= row_groups) return;
  const int nc = blockIdx.z;  // n * C + ci
  const int n = nc / C, ci = nc % C;
  const int y = (int)(g / Wq), x0 = (int)(g % Wq) * 4;
  const int d_lo = blockIdx.y * d_chunk, d_hi = min(D, d_lo + d_chunk);
  const float* lrow = fl + (((size_t)n * C + ci) * H + y) * W;
  const float* rrow = fr + (((size_t)n * C + ci) * H + y) * W;
  const int Cout = kDiff ? C : 2 * C;
  const size_t plane = (size_t)H * W;
  float* o_l = vol + ((((size_t)n * Cout + ci) * D + d_lo) * H + y) * W + x0;            // left copy / difference
  float* o_r = kDiff ? nullptr : o_l + (size_t)C * D * plane;                            // shifted right copy
  float l[4], r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x0 + i;
    l[i] = (x < W) ? lrow[x] : 0.f;
    const int xr = x - d_lo;
    r[i] = (x < W && xr >= 0) ? rrow[xr] : 0.f;
  }
  const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(vol) & 15) == 0;   // 128-bit stores: aligned rows
  for (int d = d_lo; d < d_hi; ++d) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool in = (x0 + i >= d) && (x0 + i < W);
      if (kDiff) a[i] = in ? __fsub_rn(l[i], r[i]) : 0.f;
      else {
        a[i] = in ? l[i] : 0.f;
        b[i] = in ? r[i] : 0.f;
      }
    }
    if (vec) {
      st_stream4(o_l, make_float4(a[0], a[1], a[2], a[3]));
      if (!kDiff) st_stream4(o_r, make_float4(b[0], b[1], b[2], b[3]));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (x0 + i < W) {
          st_stream(o_l + i, a[i]);
          if (!kDiff) st_stream(o_r + i, b[i]);
        }
    }
    o_l += plane;
    if (!kDiff) o_r += plane;
    // next disparity: the right quad moves one column to the left
    r[3] = r[2]; r[2] = r[1]; r[1] = r[0];
    const int xn = x0 - (d + 1);
    r[0] = (xn >= 0 && x0 < W) ? rrow[xn] : 0.f;
  }
}

int launch_shift_volume(const float* fl, const float* fr, int N, int C, int H, int W, int D, bool diff,
                        float* vol, cudaStream_t s) {
  MSN_REQUIRE(N >= 0 && C >= 1 && H >= 0 && W >= 0 && D >= 1, "volume: bad shape");
  if (N == 0 || H == 0 || W == 0) return 0;
  MSN_REQUIRE((long long)N * C <= 65535, "volume: N*C too large for one launch");
  const unsigned gx = div_up((long long)H * ((W + 3) / 4), 256);
  // enough CTAs for a few waves of 148 SMs x 8 CTAs; otherwise one thread walks all of D
  int chunks = 1;
  while (chunks < D && (long long)gx * N * C * chunks < 4 * 148 * 8) chunks *= 2;
  const int d_chunk = (D + chunks - 1) / chunks;
  dim3 grid(gx, (D + d_chunk - 1) / d_chunk, N * C);
  if (diff) shift_volume_kernel<true><<<grid, 256, 0, s>>>(fl, fr, C, H, W, D, d_chunk, vol);
  else shift_volume_kernel<false><<<grid, 256, 0, s>>>(fl, fr, C, H, W, D, d_chunk, vol);
  MSN_LAUNCH_OK();
  return 0;
}

}  // namespace msn
