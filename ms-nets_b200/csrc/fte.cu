// fte.cu -- libfeatextract drop-in kernels and the confidence pass.
//
//   swap_axes / swap_axes_back   featextract.cpp:49-105   (tiled transpose)
//   get_right_cost/get_left_cost featextract.cpp:136-172, :464-499
//   extract_likelihood (AML)     featextract.cpp:415-462  (warp per [D] row)
//   extract_ratio (PKRN)         featextract.cpp:320-356
//   WTA / second-min / peak-ratio / left-right check  (main_msnet.py:444-448 is
//   the only reference consumer: np.argmin over D; the rest has no reference code)
//
// Row kernels ([n][D], D innermost -- the reference's flattened [H][W][D]) use one
// warp per row with shuffle reductions over D; plane kernels ([D][n], the feature
// layout) use one thread per pixel, coalesced across the warp.
#include "common.cuh"
#include "feature_math.cuh"

namespace msn {

// ---------------------------------------------------------------- transposes --
// in[A][B] -> out[B][A] through a 32x33 shared tile; both sides coalesced.  Tiles are numbered along
// grid.x only (a-tile slowest), so neither extent is bound by the 65535 limit of grid.y / grid.z.
__global__ void transpose2d_kernel(const float* __restrict__ in, long long A, long long B, long long tiles_b,
                                   float* __restrict__ out) {
  __shared__ float tile[32][33];
  const long long tb = (long long)blockIdx.x % tiles_b, ta = (long long)blockIdx.x / tiles_b;
  const long long b0 = tb * 32, a0 = ta * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long a = a0 + r, b = b0 + threadIdx.x;
    if (a < A && b < B) tile[r][threadIdx.x] = in[a * B + b];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long b = b0 + r, a = a0 + threadIdx.x;
    if (a < A && b < B) st_stream(out + b * A + a, tile[threadIdx.x][r]);
  }
}

int launch_transpose2d(const float* in, long long A, long long B, float* out, cudaStream_t s) {
  if (A == 0 || B == 0) return 0;
  dim3 block(32, 8);
  const long long gx = (B + 31) / 32, gy = (A + 31) / 32;
  MSN_REQUIRE(gx * gy <= 2147483647LL, "transpose: %lld x %lld is too large for one launch", A, B);
  transpose2d_kernel<<<(unsigned)(gx * gy), block, 0, s>>>(in, A, B, gx, out);
  MSN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------- right / left re-index --
// right: out[y][x][d] = c[y][x+d][d] if x < W-d else c[0]   (featextract.cpp:151,159-161)
// left : out[y][x][d] = c[y][x-d][d] if x >= d  else c[0]   (featextract.cpp:479,487-489)
template <bool kRight>
__global__ void reindex_cost_kernel(const float* __restrict__ c, int H, int W, int D,
                                    float* __restrict__ out) {
  const long long total = (long long)H * W * D;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int d = (int)(idx % D);
  const long long p = idx / D;
  const int x = (int)(p % W);
  float v;
  if (kRight)
    v = (x < W - d) ? c[idx + (long long)d * D] : c[0];
  else
    v = (x >= d) ? c[idx - (long long)d * D] : c[0];
  st_stream(out + idx, v);
}

int launch_reindex_cost(const float* c, int H, int W, int D, bool right, float* out, cudaStream_t s) {
  const long long total = (long long)H * W * D;
  if (total == 0) return 0;
  if (right)
    reindex_cost_kernel<true><<<div_up(total, 256), 256, 0, s>>>(c, H, W, D, out);
  else
    reindex_cost_kernel<false><<<div_up(total, 256), 256, 0, s>>>(c, H, W, D, out);
  MSN_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------ warp reductions --
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// AML over rows of D costs (featextract.cpp:415-462).  The reference sums the
// denominator SEQUENTIALLY in fp32 (:444-447): terms below half an ulp of the running
// sum are dropped, which moves the result by up to D * 2^-24 relative.  To stay inside
// the 2e-6 tolerance the same order is replayed: a CTA stages 32 rows through shared
// memory ([D][33], coalesced both ways), all warps compute the exponentials, then one
// thread per row adds them in order d = 0..D-1.
constexpr int kAmlRows = 32;
constexpr int kAmlWarps = 8;
__global__ void __launch_bounds__(kAmlWarps * 32)
aml_rows_kernel(const float* __restrict__ cost, long long n, int D, float k, float* __restrict__ out) {
  extern __shared__ float tile[];  // [D][33]
  __shared__ float s_min[kAmlRows];
  __shared__ float s_inv[kAmlRows];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * kAmlRows;
  for (int r = warp; r < kAmlRows; r += kAmlWarps) {
    float m = kFill;
    if (r0 + r < n) {
      const float* c = cost + (r0 + r) * D;
      for (int d = lane; d < D; d += 32) {
        const float v = c[d];
        tile[d * 33 + r] = v;
        m = fminf(m, v);
      }
    }
    m = warp_min(m);
    if (lane == 0) s_min[r] = m;
  }
  __syncthreads();
  {
    const float m = s_min[lane];
    if (r0 + lane < n)
      for (int d = warp; d < D; d += kAmlWarps) tile[d * 33 + lane] = aml_e(tile[d * 33 + lane], m, k);
  }
  __syncthreads();
  if (warp == 0) {
    float den = 0.f;
    if (r0 + lane < n) {
#pragma unroll 8
      for (int d = 0; d < D; ++d) den = __fadd_rn(den, tile[d * 33 + lane]);
    }
    s_inv[lane] = aml_row_scale(den, s_min[lane] != kFill, k);
  }
  __syncthreads();
  for (int r = warp; r < kAmlRows; r += kAmlWarps) {
    if (r0 + r >= n) break;
    const float inv = s_inv[r];
    float* o = out + (r0 + r) * D;
    for (int d = lane; d < D; d += 32) st_stream(o + d, aml_apply(tile[d * 33 + r], inv, k));
  }
}

int launch_aml_rows(const float* cost, long long n, int D, float sigma, float* out, cudaStream_t s) {
  if (n == 0 || D == 0) return 0;
  const size_t smem = (size_t)D * 33 * sizeof(float);
  MSN_REQUIRE(smem <= 200 * 1024, "extract_likelihood: D=%d needs %zu B of shared memory (max 200 KiB)", D, smem);
  MSN_CUDA_OK(cudaFuncSetAttribute(aml_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  aml_rows_kernel<<<div_up(n, kAmlRows), kAmlWarps * 32, smem, s>>>(cost, n, D, aml_scale(sigma), out);
  MSN_LAUNCH_OK();
  return 0;
}

// out = (m + e) / (c + e) with IEEE fp32 add and division -> bit-exact
// (featextract.cpp:349); 0 for the whole row when m is fill.
__global__ void pkrn_rows_kernel(const float* __restrict__ cost, long long n, int D, float e,
                                 float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* c = cost + row * D;
  float* o = out + row * D;
  float m = kFill;
  for (int d = lane; d < D; d += 32) m = fminf(m, c[d]);
  m = warp_min(m);
  const bool dead = (m == kFill);
  const float num = __fadd_rn(m, e);
  for (int d = lane; d < D; d += 32)
    st_stream(o + d, dead ? 0.f : __fdiv_rn(num, __fadd_rn(c[d], e)));
}

int launch_pkrn_rows(const float* cost, long long n, int D, float e, float* out, cudaStream_t s) {
  if (n == 0 || D == 0) return 0;
  pkrn_rows_kernel<<<div_up(n, 8), 256, 0, s>>>(cost, n, D, e, out);
  MSN_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------- WTA --
// (min1, argmin, min2) with np.argmin tie-breaking (first minimal index wins);
// min2 is the second smallest entry counting duplicates.
struct Best {
  float m1, m2;
  int i1;
};
__device__ __forceinline__ Best best_init() { return Best{INFINITY, INFINITY, 0x7fffffff}; }
__device__ __forceinline__ void best_push(Best& b, float v, int i) {
  if (v < b.m1 || (v == b.m1 && i < b.i1)) {
    b.m2 = b.m1;
    b.m1 = v;
    b.i1 = i;
  } else {
    b.m2 = fminf(b.m2, v);
  }
}
__device__ __forceinline__ Best best_merge(const Best& a, const Best& b) {
  Best r;
  const bool a_first = (a.m1 < b.m1) || (a.m1 == b.m1 && a.i1 <= b.i1);
  if (a_first) {
    r.m1 = a.m1; r.i1 = a.i1; r.m2 = fminf(a.m2, b.m1);
  } else {
    r.m1 = b.m1; r.i1 = b.i1; r.m2 = fminf(b.m2, a.m1);
  }
  return r;
}
__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best other;
    other.m1 = __shfl_xor_sync(0xffffffffu, b.m1, o);
    other.m2 = __shfl_xor_sync(0xffffffffu, b.m2, o);
    other.i1 = __shfl_xor_sync(0xffffffffu, b.i1, o);
    b = best_merge(b, other);
  }
  return b;
}

// order-preserving map float -> uint32 (negative NCC costs sort below positive)
__device__ __forceinline__ uint32_t float_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// 64-bit merge key: (monotonic float bits << 32 | d) with the top bit flipped so
// that a SIGNED int64 all-reduce(min) across ranks orders like the unsigned key;
// equal costs resolve to the lowest disparity, as np.argmin does.
__device__ __forceinline__ long long pack_key(float m1, int d) {
  const unsigned long long u = ((unsigned long long)float_key(m1) << 32) | (uint32_t)d;
  return (long long)(u ^ 0x8000000000000000ull);
}

// layout 0: rows [n][D].  kLanes lanes share a row (32 / kLanes rows per warp): each lane reads
// 128-bit vectors when D is a multiple of 4, the (min, argmin, second-min) triples are merged
// with log2(kLanes) shuffle steps.  Index order decides ties, so the strided split is harmless.
template <int kLanes>
__global__ void __launch_bounds__(256)
wta_rows_kernel(const float* __restrict__ cost, long long n, int D, int d_begin, int32_t* __restrict__ amin,
                float* __restrict__ m1, float* __restrict__ m2, long long* __restrict__ keys) {
  constexpr int kRows = 32 / kLanes;
  const int lane = threadIdx.x & 31, sub = lane % kLanes;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kRows + lane / kLanes;
  Best b = best_init();
  if (row < n) {
    const float* c = cost + row * D;
    if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(cost) & 15) == 0)) {
      const float4* c4 = reinterpret_cast<const float4*>(c);
      const int D4 = D >> 2;
#pragma unroll 4
      for (int q = sub; q < D4; q += kLanes) {
        const float4 v = __ldg(c4 + q);
        best_push(b, v.x, 4 * q);
        best_push(b, v.y, 4 * q + 1);
        best_push(b, v.z, 4 * q + 2);
        best_push(b, v.w, 4 * q + 3);
      }
    } else {
      for (int d = sub; d < D; d += kLanes) best_push(b, c[d], d);
    }
  }
#pragma unroll
  for (int o = kLanes >> 1; o > 0; o >>= 1) {
    Best other;
    other.m1 = __shfl_xor_sync(0xffffffffu, b.m1, o);
    other.m2 = __shfl_xor_sync(0xffffffffu, b.m2, o);
    other.i1 = __shfl_xor_sync(0xffffffffu, b.i1, o);
    b = best_merge(b, other);
  }
  if (sub == 0 && row < n) {
    if (amin) amin[row] = b.i1;
    if (m1) m1[row] = b.m1;
    if (m2) m2[row] = b.m2;
    if (keys) keys[row] = pack_key(b.m1, d_begin + b.i1);
  }
}
// layout 1: planes [D][n]; a thread owns 4 consecutive pixels (128-bit loads, coalesced across
// the warp) and one of kSplit interleaved slices of D; the slices meet in shared memory.
constexpr int kWtaSplit = 4;
__global__ void __launch_bounds__(64 * kWtaSplit)
wta_planes_kernel(const float* __restrict__ cost, long long n, int D, int d_begin, int32_t* __restrict__ amin,
                  float* __restrict__ m1, float* __restrict__ m2, long long* __restrict__ keys) {
  __shared__ Best sb[kWtaSplit][64][4];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const long long p0 = ((long long)blockIdx.x * 64 + tx) * 4;
  const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(cost) & 15) == 0);
  Best b[4] = {best_init(), best_init(), best_init(), best_init()};
  if (p0 < n) {
    if (vec) {
#pragma unroll 4
      for (int d = ty; d < D; d += kWtaSplit) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(cost + (long long)d * n + p0));
        best_push(b[0], v.x, d);
        best_push(b[1], v.y, d);
        best_push(b[2], v.z, d);
        best_push(b[3], v.w, d);
      }
    } else {
      for (int d = ty; d < D; d += kWtaSplit)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (p0 + i < n) best_push(b[i], cost[(long long)d * n + p0 + i], d);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) sb[ty][tx][i] = b[i];
  __syncthreads();
  if (ty == 0 && p0 < n) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      Best r = b[i];
#pragma unroll
      for (int k = 1; k < kWtaSplit; ++k) r = best_merge(r, sb[k][tx][i]);
      const long long p = p0 + i;
      if (p < n) {
        if (amin) amin[p] = r.i1;
        if (m1) m1[p] = r.m1;
        if (m2) m2[p] = r.m2;
        if (keys) keys[p] = pack_key(r.m1, d_begin + r.i1);
      }
    }
  }
}

int launch_wta(const float* cost, long long n, int D, int layout, int d_begin, int32_t* amin, float* m1,
               float* m2, long long* keys, cudaStream_t s) {
  MSN_REQUIRE(D >= 1, "wta: D must be >= 1");
  MSN_REQUIRE(layout == 0 || layout == 1, "wta: layout must be 0 ([n][D]) or 1 ([D][n])");
  if (n == 0) return 0;
  if (layout == 0) {
    // two lanes per row (16 rows per warp) measured best at D = 192: 0.84 of the HBM peak against 0.82
    // with four lanes, 0.71 with eight, 0.58 with sixteen
    wta_rows_kernel<2><<<div_up(n, 8 * 16), 256, 0, s>>>(cost, n, D, d_begin, amin, m1, m2, keys);
  } else {
    wta_planes_kernel<<<div_up(n, 256), 64 * kWtaSplit, 0, s>>>(cost, n, D, d_begin, amin, m1, m2, keys);
  }
  MSN_LAUNCH_OK();
  return 0;
}

__global__ void wta_unpack_kernel(const long long* __restrict__ keys, long long n, int32_t* __restrict__ amin,
                                  float* __restrict__ m1) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const unsigned long long k = (unsigned long long)keys[p] ^ 0x8000000000000000ull;
  if (amin) amin[p] = (int32_t)(uint32_t)(k & 0xffffffffull);
  if (m1) m1[p] = key_float((uint32_t)(k >> 32));
}

// Merge of per-slab WTA triples ("all-gather per-GPU (min1, min2) pairs then local reduce"):
// parts ordered by ascending disparity range, each [n] (argmin with absolute disparity, smallest, second
// smallest).  Global smallest; on ties the FIRST part wins (np.argmin's lowest index); the second smallest is
// the second entry of the sorted multiset of all parts' pairs, duplicates counted.
__global__ void wta_merge_kernel(const int32_t* __restrict__ idx_p, const float* __restrict__ m1_p,
                                 const float* __restrict__ m2_p, int parts, long long n, int32_t* __restrict__ idx,
                                 float* __restrict__ m1, float* __restrict__ m2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = INFINITY, b = INFINITY;
  int k = 0;
  for (int p = 0; p < parts; ++p) {
    const float v1 = m1_p[(size_t)p * n + i], v2 = m2_p[(size_t)p * n + i];
    if (v1 < a) { b = a; a = v1; k = idx_p[(size_t)p * n + i]; }
    else if (v1 < b) b = v1;
    if (v2 < b) b = v2;      // v2 >= v1 >= a here
  }
  idx[i] = k; m1[i] = a; m2[i] = b;
}

int launch_wta_merge(const int32_t* idx_p, const float* m1_p, const float* m2_p, int parts, long long n, int32_t* idx,
                     float* m1, float* m2, cudaStream_t s) {
  if (n == 0) return 0;
  wta_merge_kernel<<<div_up(n, 256), 256, 0, s>>>(idx_p, m1_p, m2_p, parts, n, idx, m1, m2);
  MSN_LAUNCH_OK();
  return 0;
}

int launch_wta_unpack(const long long* keys, long long n, int32_t* amin, float* m1, cudaStream_t s) {
  if (n == 0) return 0;
  wta_unpack_kernel<<<div_up(n, 256), 256, 0, s>>>(keys, n, amin, m1);
  MSN_LAUNCH_OK();
  return 0;
}

__global__ void pkrn_conf_kernel(const float* __restrict__ m1, const float* __restrict__ m2, long long n,
                                 float e, float* __restrict__ conf) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float a = m1[p];
  conf[p] = (a == kFill) ? 0.f : __fdiv_rn(__fadd_rn(a, e), __fadd_rn(m2[p], e));
}

int launch_pkrn_conf(const float* m1, const float* m2, long long n, float e, float* conf, cudaStream_t s) {
  if (n == 0) return 0;
  pkrn_conf_kernel<<<div_up(n, 256), 256, 0, s>>>(m1, m2, n, e, conf);
  MSN_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------- left-right check --
// dL = argmin_d c[y][x][d]; dR = argmin_d (x+d < W ? c[y][x+d][d] : c[0]);
// one warp per pixel computes both with shuffle reductions.
__global__ void lrc_disp_kernel(const float* __restrict__ c, int H, int W, int D, int32_t* __restrict__ dl,
                                int32_t* __restrict__ dr) {
  const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= (long long)H * W) return;
  const int x = (int)(p % W);
  const float c0 = c[0];
  Best bl = best_init(), br = best_init();
  for (int d = lane; d < D; d += 32) {
    best_push(bl, c[p * D + d], d);
    best_push(br, (x < W - d) ? c[(p + d) * D + d] : c0, d);
  }
  bl = warp_best(bl);
  br = warp_best(br);
  if (lane == 0) {
    dl[p] = bl.i1;
    dr[p] = br.i1;
  }
}
__global__ void lrc_mask_kernel(const int32_t* __restrict__ dl, const int32_t* __restrict__ dr, int H, int W,
                                int thresh, uint8_t* __restrict__ mask) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)H * W) return;
  const int x = (int)(p % W);
  const int d = dl[p];
  uint8_t ok = 0;
  if (x - d >= 0) {
    const int diff = d - dr[p - d];
    ok = (diff <= thresh && -diff <= thresh) ? 1 : 0;
  }
  mask[p] = ok;
}

int launch_lrc(const float* c, int H, int W, int D, int thresh, int32_t* dl, int32_t* dr, uint8_t* mask,
               cudaStream_t s) {
  const long long n = (long long)H * W;
  if (n == 0) return 0;
  lrc_disp_kernel<<<div_up(n, 8), 256, 0, s>>>(c, H, W, D, dl, dr);
  MSN_LAUNCH_OK();
  lrc_mask_kernel<<<div_up(n, 256), 256, 0, s>>>(dl, dr, H, W, thresh, mask);
  MSN_LAUNCH_OK();
  return 0;
}

}  // namespace msn
