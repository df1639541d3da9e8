// sadsob.cu -- SAD over Sobel images through an fp32 summed-area table whose
// ROUNDING IS PART OF THE RESULT (matchers.cpp:356-438).
//
// The reference builds, for every disparity, S = horizontal-prefix(vertical-
// prefix(|L(i,j) - R(i,j-d)|)) with sequential fp32 adds and evaluates the box as
// ((br - bl) - tr) + tl.  On a 560x980 pair the running sums pass 2^24, so S is
// NOT the exact integer sum: a bit-exact result needs the same
// add order.  Both prefix passes are therefore replayed as sequential chains --
// there are D*(W+1) independent vertical chains and D*(H+1) independent
// horizontal chains, plenty of parallelism -- organised so that every global
// access is coalesced:
//
//   1. vband kernel: one thread per (d, column) walks down the image and records
//      the vertical prefix at the first row of every 32-row band  -> Vb[d][band][j].
//   2. scan kernel: one warp per (d, band) sweeps left to right in 32x32 tiles.
//      Lanes first own COLUMNS (finish the vertical prefix inside the band from
//      Vb, coalesced image reads), the tile is transposed through shared memory,
//      lanes then own ROWS (continue the horizontal chain, carry in a register),
//      and finally own columns again to evaluate the boxes and store one
//      coalesced 128-byte row segment per output row.
//
// A band holds 32 table rows, i.e. 32 - wsize output rows (the box needs rows i
// and i + wsize), so bands overlap by wsize rows.
#include "common.cuh"

namespace msn {

constexpr int kSadTile = 32;
constexpr int kSadMaxW = 16;                       // wsize <= 16
constexpr int kSadVStride = kSadTile + 1;          // tV[32][33]
constexpr int kSadSStride = kSadTile + kSadMaxW + 1;  // tS[32][49]: halo(wsize) + 32 columns
constexpr int kSadWarps = 4;

__device__ __forceinline__ float absdiff_rn(float a, float b) { return fabsf(__fsub_rn(a, b)); }

// grid: (ceil(IW/128), Dn, N); thread = table column j in [0, W].
__global__ void sadsob_vband_kernel(const float* __restrict__ L, const float* __restrict__ R, int H, int W,
                                    int d_begin, int RB, int NB, size_t img_stride,
                                    float* __restrict__ Vb) {
  const int IW = W + 1;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int dd = blockIdx.y, d = d_begin + dd, n = blockIdx.z;
  if (j >= IW) return;
  const int jc = j - 1;
  const bool active = (jc >= d);  // implies jc >= 0; jc < W because j <= W
  const float* l = L + n * img_stride + jc;
  const float* r = R + n * img_stride + (jc - d);
  float* vb = Vb + (((size_t)n * gridDim.y + dd) * NB) * IW + j;
  float v = 0.f;
  int band = 0;
  // rows are taken 8 at a time so 16 independent loads are in flight per thread; the adds
  // stay strictly sequential (adding the 0.0f of an inactive column is exact)
  for (int row0 = 0; row0 < H; row0 += 8) {
    float av[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int row = row0 + q;
      av[q] = (active && row < H) ? absdiff_rn(__ldg(l + (size_t)row * W), __ldg(r + (size_t)row * W)) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int row = row0 + q;
      if (band < NB && row == band * RB) {
        vb[(size_t)band * IW] = v;
        ++band;
      }
      v = __fadd_rn(v, av[q]);
    }
  }
}

// grid: (ceil(Dn*NB / kSadWarps), 1, N); one warp per (dd, band).
// WS > 0: window size known at compile time (shared strides and box offsets become
// immediates); WS == 0: runtime wsize (any 1..16).
template <int WS>
__global__ void __launch_bounds__(kSadWarps * 32)
sadsob_scan_kernel(const float* __restrict__ L, const float* __restrict__ R, int H, int W, int Dn,
                   int d_begin, int wsize_rt, int NB, size_t img_stride,
                   const float* __restrict__ Vb, float* __restrict__ out, size_t out_stride, int out_pitch) {
  constexpr int SS = (WS > 0) ? (kSadTile + WS + ((WS & 1) ? 0 : 1)) : kSadSStride;  // odd stride
  __shared__ float sV[kSadWarps][kSadTile * kSadVStride];
  __shared__ float sS[kSadWarps][kSadTile * SS];
  const int wsize = (WS > 0) ? WS : wsize_rt;
  const int RB = kSadTile - wsize;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int job = blockIdx.x * kSadWarps + warp;
  if (job >= Dn * NB) return;  // whole warp exits together
  const int dd = job / NB, b = job % NB, n = blockIdx.z;
  const int d = d_begin + dd;
  const int IW = W + 1, wc = wsize / 2;
  const int i0 = b * RB;
  float* tV = sV[warp];
  float* tS = sS[warp];
  const float* Ln = L + n * img_stride;
  const float* Rn = R + n * img_stride;
  const float* vb = Vb + (((size_t)n * Dn + dd) * NB + b) * IW;
  float* o = out + n * out_stride + (size_t)dd * H * out_pitch;  // rows of out_pitch floats
  const int rows_live = min(kSadTile - 1, H - i0);      // image rows i0 .. i0+rows_live-1 exist
  const int rmax = min(RB, H - wsize - i0);             // origin rows produced by this band

  // table columns <= d are zero, so the sweep starts at the tile holding column d
  const int t0 = d / kSadTile, t1 = (W - 1) / kSadTile;
  float s = 0.f;  // horizontal carry of table row i0 + lane
  float* srow = tS + lane * SS;
#pragma unroll 1
  for (int k = 0; k < wsize; ++k) srow[kSadTile + k] = 0.f;
  for (int t = t0; t <= t1; ++t) {
    // (a) lanes own table columns: finish the vertical prefix inside the band.
    // All loads of the tile column are issued before the dependent add chain starts.
    const int j = t * kSadTile + lane;
    const int jc = j - 1;
    const bool active = (j < IW) && (jc >= d);
    float v = (j < IW) ? __ldg(vb + j) : 0.f;
    float av[kSadTile - 1];
    {
      const float* lp = Ln + (size_t)i0 * W + jc;
      const float* rp = lp + (Rn - Ln) - d;
#pragma unroll
      for (int r = 0; r < kSadTile - 1; ++r) {
        av[r] = (active && r < rows_live) ? absdiff_rn(__ldg(lp), __ldg(rp)) : 0.f;
        lp += W; rp += W;
      }
    }
    tV[lane] = v;
#pragma unroll
    for (int r = 1; r < kSadTile; ++r) {
      v = __fadd_rn(v, av[r - 1]);  // + 0.0f on inactive entries is exact
      tV[r * kSadVStride + lane] = v;
    }
    __syncwarp();
    // (b) lanes own table rows: slide the halo, continue the horizontal chain
#pragma unroll
    for (int k = 0; k < (WS > 0 ? WS : kSadMaxW); ++k)
      if (k < wsize) srow[k] = srow[kSadTile + k];
    {
      const float* vrow = tV + lane * kSadVStride;
      float* dst = srow + wsize;
#pragma unroll
      for (int c = 0; c < kSadTile; ++c) {
        s = __fadd_rn(s, vrow[c]);
        dst[c] = s;
      }
    }
    __syncwarp();
    // (c) lanes own origin columns again: boxes ((br - bl) - tr) + tl
    const int jo = t * kSadTile - wsize + lane;  // window origin column
    if (jo >= d && jo < W - wsize) {
      const float* top = tS + lane;
      const float* bot = top + wsize * SS;
      float* op = o + (size_t)(i0 + wc) * out_pitch + (jo + wc);
#pragma unroll 9
      for (int r = 0; r < rmax; ++r) {
        const float val = __fadd_rn(__fsub_rn(__fsub_rn(bot[wsize], bot[0]), top[wsize]), top[0]);
        st_stream(op, val);
        top += SS; bot += SS; op += out_pitch;
      }
    }
    __syncwarp();
  }
}

static inline void sadsob_geom(int H, int wsize, int* RB, int* NB) {
  *RB = kSadTile - wsize;
  const int rows = H - wsize;  // origin rows i in [0, H - wsize)
  *NB = rows > 0 ? (rows + *RB - 1) / *RB : 0;
}

size_t sadsob_workspace_bytes_n(int N, int H, int W, int Dn, int wsize) {
  int RB, NB;
  sadsob_geom(H, wsize, &RB, &NB);
  return (size_t)N * Dn * (NB > 0 ? NB : 1) * (W + 1) * sizeof(float);
}
size_t sadsob_workspace_bytes(int H, int W, int D, int wsize) {
  return sadsob_workspace_bytes_n(1, H, W, D, wsize);
}

// N pairs; L/R are float [N][H][W]; out is [N][Dn][H][out_pitch] (out_pitch >= W; the caller may
// pre-offset `out` by a few columns) with stride out_stride floats per pair.
int launch_sadsob_n(const float* L, const float* R, int N, int H, int W, int Dn, int d_begin, int wsize,
                    float* out, size_t out_stride, int out_pitch, bool write_fill, void* workspace,
                    cudaStream_t s) {
  MSN_REQUIRE(wsize >= 1 && wsize <= kSadMaxW, "sadsob: wsize %d unsupported (1..%d)", wsize, kSadMaxW);
  if (write_fill) {
    for (int n = 0; n < N; ++n)
      if (launch_fill(out + n * out_stride, (size_t)Dn * H * out_pitch, kFill, s)) return 1;
  }
  int RB, NB;
  sadsob_geom(H, wsize, &RB, &NB);
  if (NB <= 0 || W - wsize <= 0 || Dn <= 0) return 0;
  float* Vb = static_cast<float*>(workspace);
  dim3 g1(div_up(W + 1, 128), Dn, N);
  sadsob_vband_kernel<<<g1, 128, 0, s>>>(L, R, H, W, d_begin, RB, NB, (size_t)H * W, Vb);
  MSN_LAUNCH_OK();
  dim3 g2(div_up((long long)Dn * NB, kSadWarps), 1, N);
  if (wsize == 5)  // the reference's default sobelw (cbmv_generator.py:440)
    sadsob_scan_kernel<5><<<g2, kSadWarps * 32, 0, s>>>(L, R, H, W, Dn, d_begin, wsize, NB, (size_t)H * W, Vb, out,
                                                        out_stride, out_pitch);
  else
    sadsob_scan_kernel<0><<<g2, kSadWarps * 32, 0, s>>>(L, R, H, W, Dn, d_begin, wsize, NB, (size_t)H * W, Vb, out,
                                                        out_stride, out_pitch);
  MSN_LAUNCH_OK();
  return 0;
}

int launch_sadsob(const float* L, const float* R, int H, int W, int D, int d_begin, int wsize, float* out,
                  bool write_fill, void* workspace, cudaStream_t s) {
  return launch_sadsob_n(L, R, 1, H, W, D, d_begin, wsize, out, (size_t)D * H * W, W, write_fill, workspace, s);
}

}  // namespace msn
