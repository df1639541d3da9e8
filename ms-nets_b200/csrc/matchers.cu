// matchers.cu -- standalone H x W x D cost kernels behind the libmatchers drop-in API.
//
// These serve the per-function boundary (mtc.census / nccNister / zsad) for any
// window size the reference accepts and keep the reference's OUTPUT LAYOUTS:
// census [H][W][D] (D innermost), the others [D][H][W].  The throughput path
// (default windows, [C][D][h][w] output, no intermediate volumes) is ms_fused.cu.
#include "common.cuh"

namespace msn {

// ------------------------------------------------------------------ census --
// cost(y,x,d) = popcount(desc_L(y,x) xor desc_R(y,x-d))   (matchers.cpp:311-342)
// One thread produces four consecutive disparities of one pixel and writes them
// with a single 128-bit streaming store along the innermost D axis; a warp
// covers 512 contiguous output bytes.
template <bool kVec4>
__global__ void census_cost_hwd_kernel(const uint32_t* __restrict__ dl, const uint32_t* __restrict__ dr,
                                       int H, int W, int D, int wsize, int nw, float* __restrict__ out) {
  const int dq = (D + 3) >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (idx >= (long long)W * dq) return;
  const int x = (int)(idx / dq);
  const int d0 = (int)(idx % dq) * 4;
  const Win win(wsize);
  const bool pix_ok = win.row_ok(y, H) && win.col_ok(x, W);
  float v[4] = {kFill, kFill, kFill, kFill};
  if (pix_ok) {
    uint32_t l[kMaxCensusWords];
    const uint32_t* lp = dl + ((size_t)y * W + x) * nw;
#pragma unroll
    for (int k = 0; k < kMaxCensusWords; ++k) l[k] = (k < nw) ? lp[k] : 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = d0 + i;
      if (d < D && (x - win.wc) >= d) {
        const uint32_t* rp = dr + ((size_t)y * W + (x - d)) * nw;
        int c = 0;
#pragma unroll
        for (int k = 0; k < kMaxCensusWords; ++k)
          if (k < nw) c += __popc(l[k] ^ rp[k]);
        v[i] = (float)c;
      }
    }
  }
  float* o = out + ((size_t)y * W + x) * D + d0;
  if (kVec4) {
    st_stream4(o, make_float4(v[0], v[1], v[2], v[3]));
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (d0 + i < D) st_stream(o + i, v[i]);
  }
}

int launch_census_cost_hwd(const uint32_t* dl, const uint32_t* dr, int H, int W, int D, int wsize,
                           float* out, cudaStream_t s) {
  if (D <= 0 || H <= 0 || W <= 0) return 0;
  const int nw = (wsize * wsize + 31) / 32;
  const int dq = (D + 3) / 4;
  dim3 block(256), grid(div_up((long long)W * dq, 256), H);
  if (D % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)   // 128-bit stores need a 16-byte aligned volume
    census_cost_hwd_kernel<true><<<grid, block, 0, s>>>(dl, dr, H, W, D, wsize, nw, out);
  else
    census_cost_hwd_kernel<false><<<grid, block, 0, s>>>(dl, dr, H, W, D, wsize, nw, out);
  MSN_LAUNCH_OK();
  return 0;
}

// Same cost in the [Dn][H][W] plane layout (disparities d_begin .. d_begin+Dn-1) for
// the generic / slab feature path; one thread per voxel, coalesced along W.
__global__ void census_cost_dhw_kernel(const uint32_t* __restrict__ dl, const uint32_t* __restrict__ dr,
                                       int H, int W, int d_begin, int wsize, int nw,
                                       float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, d = d_begin + blockIdx.z;
  if (x >= W) return;
  const Win win(wsize);
  float v = kFill;
  if (win.ok(y, x, d, H, W)) {
    const uint32_t* lp = dl + ((size_t)y * W + x) * nw;
    const uint32_t* rp = dr + ((size_t)y * W + (x - d)) * nw;
    int c = 0;
#pragma unroll
    for (int k = 0; k < kMaxCensusWords; ++k)
      if (k < nw) c += __popc(lp[k] ^ rp[k]);
    v = (float)c;
  }
  st_stream(out + ((size_t)blockIdx.z * H + y) * W + x, v);
}

int launch_census_cost_dhw(const uint32_t* dl, const uint32_t* dr, int H, int W, int d_begin, int Dn,
                           int wsize, float* out, cudaStream_t s) {
  if (Dn <= 0 || H <= 0 || W <= 0) return 0;
  const int nw = (wsize * wsize + 31) / 32;
  dim3 block(128), grid(div_up(W, 128), H, Dn);
  census_cost_dhw_kernel<<<grid, block, 0, s>>>(dl, dr, H, W, d_begin, wsize, nw, out);
  MSN_LAUNCH_OK();
  return 0;
}

// --------------------------------------------------------------------- ncc --
// cost = (float)( -(w^2*P - A_L*A_R) * C_L * C_R ), products left to right in
// fp64 (matchers.cpp:200-201); 1.0f when either C is not finite (:196,204).
// P = sum over the window of L*R(.-d) is an exact integer, so a direct window sum
// replaces the reference's per-disparity fp64 integral image bit for bit.
__global__ void ncc_cost_kernel(const uint8_t* __restrict__ L, const uint8_t* __restrict__ R,
                                const unsigned long long* __restrict__ Al,
                                const unsigned long long* __restrict__ Ar, const double* __restrict__ Cl,
                                const double* __restrict__ Cr, int H, int W, int d_begin, int wsize,
                                float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, d = d_begin + blockIdx.z;
  if (x >= W) return;
  const Win win(wsize);
  float v = kFill;
  if (win.ok(y, x, d, H, W)) {
    const uint8_t* lo = L + (size_t)(y - win.wc) * W + (x - win.wc);
    const uint8_t* ro = R + (size_t)(y - win.wc) * W + (x - win.wc - d);
    unsigned long long p = 0ull;
    for (int a = 0; a < wsize; ++a)
      for (int b = 0; b < wsize; ++b)
        p += (unsigned)lo[(size_t)a * W + b] * (unsigned)ro[(size_t)a * W + b];
    const size_t cl = (size_t)y * W + x, cr = (size_t)y * W + (x - d);
    const double c_l = Cl[cl], c_r = Cr[cr];
    if (isfinite(c_l) && isfinite(c_r)) {
      const double num = __dsub_rn(__dmul_rn((double)(wsize * wsize), (double)p), (double)(Al[cl] * Ar[cr]));
      v = (float)__dmul_rn(__dmul_rn(-num, c_l), c_r);
    } else {
      v = 1.0f;
    }
  }
  st_stream(out + ((size_t)blockIdx.z * H + y) * W + x, v);
}

int launch_ncc_cost(const uint8_t* L, const uint8_t* R, const unsigned long long* Al,
                    const unsigned long long* Ar, const double* Cl, const double* Cr, int H, int W,
                    int d_begin, int Dn, int wsize, float* out, cudaStream_t s) {
  if (Dn <= 0 || H <= 0 || W <= 0) return 0;
  dim3 block(128), grid(div_up(W, 128), H, Dn);
  ncc_cost_kernel<<<grid, block, 0, s>>>(L, R, Al, Ar, Cl, Cr, H, W, d_begin, wsize, out);
  MSN_LAUNCH_OK();
  return 0;
}

// -------------------------------------------------------------------- zsad --
// cost = sum over taps (row-major, sequential fp32) of |((L - mL) - R) + mR|
// (matchers.cpp:499-506).  Every add is an explicit round-to-nearest intrinsic so
// nothing is contracted or re-associated: the result is bit-identical.
__global__ void zsad_cost_kernel(const uint8_t* __restrict__ L, const uint8_t* __restrict__ R,
                                 const float* __restrict__ ml, const float* __restrict__ mr, int H, int W,
                                 int d_begin, int wsize, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, d = d_begin + blockIdx.z;
  if (x >= W) return;
  const Win win(wsize);
  float v = kFill;
  if (win.ok(y, x, d, H, W)) {
    const uint8_t* lo = L + (size_t)(y - win.wc) * W + (x - win.wc);
    const uint8_t* ro = R + (size_t)(y - win.wc) * W + (x - win.wc - d);
    const float mL = ml[(size_t)y * W + x], mR = mr[(size_t)y * W + (x - d)];
    float acc = 0.f;
    for (int a = 0; a < wsize; ++a)
      for (int b = 0; b < wsize; ++b) {
        float t = __fsub_rn((float)lo[(size_t)a * W + b], mL);
        t = __fsub_rn(t, (float)ro[(size_t)a * W + b]);
        t = __fadd_rn(t, mR);
        acc = __fadd_rn(acc, fabsf(t));
      }
    v = acc;
  }
  st_stream(out + ((size_t)blockIdx.z * H + y) * W + x, v);
}

int launch_zsad_cost(const uint8_t* L, const uint8_t* R, const float* ml, const float* mr, int H, int W,
                     int d_begin, int Dn, int wsize, float* out, cudaStream_t s) {
  if (Dn <= 0 || H <= 0 || W <= 0) return 0;
  dim3 block(128), grid(div_up(W, 128), H, Dn);
  zsad_cost_kernel<<<grid, block, 0, s>>>(L, R, ml, mr, H, W, d_begin, wsize, out);
  MSN_LAUNCH_OK();
  return 0;
}

}  // namespace msn
