// ms_fused.cu -- the throughput path: MS feature volume [N][8][D][h][w] straight from
// the uint8 pair, no intermediate cost volumes except the sadsob scratch.
//
// Replaces, in one pass, get_costs (census + nccNister + zsad + sobel/sadsob +
// 3x swap_axes + border crop, cbmv_generator.py:27-79) and extract_features_left
// (clip/normalise, 4x extract_likelihood, float64 scratch, transpose + cast,
// cbmv_generator.py:258-308) for the default windows 11/3/5/5.
//
// Launch sequence (all on one stream):
//   1. ms_prep_kernel    per padded pixel of every image: census code (4x u32), ZSAD
//                        window mean, NCC window sum A and fp64 C = 1/sqrt(9B - A^2),
//                        float copy of the pixel, Sobel response.
//   2. sadsob vband+scan (sadsob.cu) -> raw SAD-of-Sobel volume [N][D][H][W]; its fp32
//                        summed-area table needs whole-row sequential scans, so it
//                        cannot live inside an x-tile.
//   3. ms_fused_kernel   one CTA = one output row y x 32 pixels x ALL D:
//        stage    right-image row data (census codes, stats, 5 float rows) and the
//                 tile's SAD-of-Sobel costs stream into shared memory with cp.async.
//        phase 1  8 warps split D; lane = pixel.  Per (pixel, d): census popcount,
//                 NCC (9 fp32 products of exact integers, fp64 scaling), ZSAD (75
//                 ordered fp32 adds over a register-resident 5x5 window that slides
//                 with d); raw costs are parked in shared memory (13 B/voxel: three
//                 floats + census byte); per-pixel minima.
//        phase 2  warp-specialised, both halves only READ the parked costs:
//                 warps 0-3: one thread per (pixel, matcher) adds the AML denominator in
//                 d order (the reference's sequential fp32 sum, featextract.cpp:444-447);
//                 warps 4-7: thread = (pixel quad, d): channels 0-3 normalised and stored
//                 as 128-bit row segments.
//        phase 3  channels 4-7 = exp(-(c-m)^2/sigma) / den, 128-bit row segments.
//
// Bounding resource: HBM writes (32 B per voxel) co-limited by issue slots -- ZSAD
// alone is 75 dependent-order FADDs per voxel (DESIGN.md has the arithmetic).
#include "ms_fused.cuh"

#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "feature_math.cuh"

namespace msn {

namespace {

constexpr int kWarps = 8;
constexpr int kTile = 32;     // pixels per CTA
constexpr int kPadT = 2;      // padded rows above/below (ZSAD halo)
constexpr int kPadR = 40;     // padded columns to the right (tile overhang + halo)
constexpr int kCensW = 11, kNccW = 3, kSadW = 5;
constexpr int kMaxFusedD = 448;

struct __align__(16) RStat {
  float mean;  // ZSAD window mean (matchers.cpp:482)
  float A;     // NCC window sum, exact integer <= 2295 (matchers.cpp:140)
  double C;    // 1/sqrt(9*B - A*A) in fp64 (matchers.cpp:146)
};

struct FusedGeom {
  int N, H, W, h, w, D, bh, bwl;
  int Hp, Wp, padL;
  int sxo, Ws;   // SAD-of-Sobel scratch: column offset and row pitch (tile starts land on 16 B)
  __host__ __device__ size_t img_px() const { return (size_t)Hp * Wp; }
};

FusedGeom make_geom(int N, int H, int W, const msn_ms_params* p) {
  FusedGeom g;
  g.N = N; g.H = H; g.W = W; g.D = p->ndisp;
  g.bh = p->board_h; g.bwl = p->board_w_left;
  g.h = H - 2 * p->board_h;
  g.w = W - p->board_w_left - p->board_w_right;
  g.padL = (g.D + 1 + 40 + 8 + 7) & ~7;  // D-1 columns of disparity + dummy-step slack (kSlack) + halo/alignment
  g.Hp = H + 2 * kPadT;
  g.Wp = (W + g.padL + kPadR + 3) & ~3;
  g.sxo = (4 - (g.bwl & 3)) & 3;            // x0 + bwl + sxo is a multiple of 4 for every tile
  g.Ws = sadsob_fast_pitch(W + g.sxo + kTile);  // compile-time pitch of the scan kernels (0: too wide)
  return g;
}

struct FusedWs {
  uint4* desc[2];
  RStat* stat[2];
  float* fimg[2];
  float* sob[2];
  float* sadsob;   // [N][D][H][W]
  void* sad_ws;
  size_t total;
  void carve(char* base, const FusedGeom& g) {
    size_t off = 0;
    auto take = [&](size_t bytes) { char* q = base ? base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return q; };
    const size_t np = (size_t)g.N * g.img_px(), n = (size_t)g.N * g.H * g.W;
    for (int i = 0; i < 2; ++i) desc[i] = (uint4*)take(np * sizeof(uint4));
    for (int i = 0; i < 2; ++i) stat[i] = (RStat*)take(np * sizeof(RStat));
    for (int i = 0; i < 2; ++i) fimg[i] = (float*)take(np * sizeof(float));
    for (int i = 0; i < 2; ++i) sob[i] = (float*)take((size_t)g.N * (g.H + kSadRowPad) * g.Ws * sizeof(float));  // zero padded
    sadsob = (float*)take((size_t)g.N * g.D * g.H * g.Ws * sizeof(float) + 256);
    sad_ws = take(sadsob_workspace_bytes_n(g.N, g.H, g.W, g.D, kSadW));
    total = off;
  }
};

// ------------------------------------------------------------------ prep --
// grid (ceil(Wp/128), Hp, 2N); blockIdx.z = 2*n + side.
__global__ void __launch_bounds__(128)
ms_prep_kernel(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right, FusedGeom g,
               uint4* __restrict__ descL, uint4* __restrict__ descR, RStat* __restrict__ statL,
               RStat* __restrict__ statR, float* __restrict__ fL, float* __restrict__ fR,
               float* __restrict__ sobL, float* __restrict__ sobR) {
  const int xp = blockIdx.x * blockDim.x + threadIdx.x;
  const int yp = blockIdx.y;
  const int n = blockIdx.z >> 1, side = blockIdx.z & 1;
  if (xp >= g.Wp) return;
  const int H = g.H, W = g.W;
  const uint8_t* img = (side ? right : left) + (size_t)n * H * W;
  const int X = xp - g.padL, Y = yp - kPadT;
  const bool inside = (X >= 0 && X < W && Y >= 0 && Y < H);
  const size_t po = (size_t)n * g.img_px() + (size_t)yp * g.Wp + xp;

  uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
  RStat st;
  st.mean = 0.f; st.A = 0.f; st.C = 0.0;
  float pix = 0.f;
  if (inside) {
    const int c = img[(size_t)Y * W + X];
    pix = (float)c;
    // census 11x11: bit k = a*11+b set iff centre < tap (matchers.cpp:290-297)
    if (Y >= 5 && Y < H - 6 && X >= 5 && X < W - 6) {
      const uint8_t* org = img + (size_t)(Y - 5) * W + (X - 5);
      uint32_t words[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int a = 0; a < kCensW; ++a) {
#pragma unroll
        for (int b = 0; b < kCensW; ++b) {
          const int bit = a * kCensW + b;
          const uint32_t v = (c < (int)__ldg(org + (size_t)a * W + b)) ? 1u : 0u;
          words[bit >> 5] |= v << (bit & 31);
        }
      }
      w0 = words[0]; w1 = words[1]; w2 = words[2]; w3 = words[3];
    }
    // ZSAD 5x5 mean: exact integer sum, one IEEE division (matchers.cpp:472-485)
    if (Y >= 2 && Y < H - 3 && X >= 2 && X < W - 3) {
      const uint8_t* org = img + (size_t)(Y - 2) * W + (X - 2);
      int sum = 0;
#pragma unroll
      for (int a = 0; a < kSadW; ++a)
#pragma unroll
        for (int b = 0; b < kSadW; ++b) sum += __ldg(org + (size_t)a * W + b);
      st.mean = __fdiv_rn((float)sum, 25.0f);
    }
    // NCC 3x3: A, and C in fp64 exactly as matchers.cpp:146
    if (Y >= 1 && Y < H - 2 && X >= 1 && X < W - 2) {
      const uint8_t* org = img + (size_t)(Y - 1) * W + (X - 1);
      unsigned a_sum = 0, b_sum = 0;
#pragma unroll
      for (int a = 0; a < kNccW; ++a)
#pragma unroll
        for (int b = 0; b < kNccW; ++b) {
          const unsigned v = __ldg(org + (size_t)a * W + b);
          a_sum += v;
          b_sum += v * v;
        }
      st.A = (float)a_sum;
      const double var = __dsub_rn((double)(9u * b_sum), __dmul_rn((double)a_sum, (double)a_sum));
      st.C = __ddiv_rn(1.0, __dsqrt_rn(var));
    }
    // Sobel Gx (matchers.cpp:538-547); unpadded array feeds the sadsob scan
    float sv = 0.f;
    if (Y >= 1 && Y < H - 2 && X >= 1 && X < W - 2) {
      const uint8_t* p = img + (size_t)(Y - 1) * W + (X - 1);
      sv = (float)(((int)p[2] - (int)p[0]) + 2 * ((int)p[W + 2] - (int)p[W]) +
                   ((int)p[2 * W + 2] - (int)p[2 * W]));
    }
    (side ? sobR : sobL)[(size_t)n * (H + kSadRowPad) * g.Ws + (size_t)Y * g.Ws + X] = sv;
  }
  (side ? descR : descL)[po] = make_uint4(w0, w1, w2, w3);
  (side ? statR : statL)[po] = st;
  (side ? fR : fL)[po] = pix;
}

// ----------------------------------------------------------------- fused --
struct FusedArgs {
  FusedGeom g;
  const uint4 *descL, *descR;
  const RStat *statL, *statR;
  const float *fL, *fR;
  const float* sadsob;  // [N][D][H][W] (+ slack)
  float* out;           // [N][8][D][h][w]
  float k_cen, k_ncc, k_sad;
  int DC;               // disparity steps per warp (multiple of 5)
  int tiles_x;
  int num_tiles;
  float stagger_ns;     // persistent kernel: CTA start offsets are drawn from [0, stagger_ns)
};

// Shared-memory layout for disparity counts up to DMAX.  Row strides are compile-time
// so every shared access in the hot loop is "pointer + immediate".
//   SLACK : right-image entries below index 0 reached by dummy steps (d >= D); 8*DC - D of them
//   NBUF  : staging buffers (1: one tile per CTA; 2: persistent CTA, next tile prefetched)
constexpr int kSlack = 40;  // enough for every D (see make_geom: padL covers it)
template <int DMAX, int SLACK = kSlack, int NBUF = 1>
struct Lay {
  static constexpr int kSl = SLACK;
  static constexpr int RW = (DMAX + kTile - 1 + SLACK + 3) & ~3;   // desc / stat entries
  static constexpr int RWF = RW + 8;                               // float row: halo 2+2, align shift <= 3
  static constexpr int LF = 40;                                    // left float row: 32 + halo 4, shift <= 3
  static constexpr int DS = DMAX + 1;                              // parked planes + 1 scratch plane
  // one staging buffer: right-image row data of a tile (+ the 32 left pixels' data when the
  // persistent kernel prefetches them through TMA as well)
  static constexpr size_t st_desc = 0;
  static constexpr size_t st_stat = st_desc + (size_t)RW * 16;
  static constexpr size_t st_rf = st_stat + (size_t)RW * 16;
  static constexpr size_t st_ldesc = st_rf + (size_t)5 * RWF * 4;
  static constexpr size_t st_lstat = st_ldesc + (NBUF > 1 ? kTile * 16 : 0);
  static constexpr size_t st_lf = st_lstat + (NBUF > 1 ? kTile * 16 : 0);
  static constexpr size_t st_bytes = st_lf + (NBUF > 1 ? 5 * LF * 4 : 0);
  static constexpr size_t off_stage = 0;
  static constexpr size_t off_red = off_stage + NBUF * st_bytes;              // [8][4][32]
  static constexpr size_t off_min = off_red + (size_t)kWarps * 4 * 32 * 4;   // [4][32]
  static constexpr size_t off_inv = off_min + 4 * 32 * 4;                    // [4][32]
  static constexpr size_t off_lut = off_inv + 4 * 32 * 4;                    // [128]
  static constexpr size_t off_par = (off_lut + 128 * 4 + 127) & ~(size_t)127;  // [3][DS][32] ncc, sadsob, zsad (128 B aligned: TMA destination)
  static constexpr size_t off_cen = off_par + (size_t)3 * DS * 32 * 4;       // [DS][32] bytes
  static constexpr size_t off_bar = (off_cen + (size_t)DS * 32 + 15) & ~(size_t)15;  // mbarriers (8 B each)
  static constexpr size_t bytes = off_bar + 32;
};

__device__ __forceinline__ float int_to_float_small(int c) {  // exact for 0 <= c < 2^23
  return __uint_as_float(0x4B000000u | (uint32_t)c) - 8388608.0f;
}
// c / 120 correctly rounded for integer c in [0,120] (exhaustively checked in
// tests/test_host_math.py): reciprocal multiply + one FMA residual correction.
__device__ __forceinline__ float div120_exact(float cf) {
  const float r = 1.0f / 120.0f;
  float q = __fmul_rn(cf, r);
  const float rem = __fmaf_rn(-120.0f, q, cf);
  return __fmaf_rn(rem, r, q);
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
// ---- TMA / bulk-copy helpers (cp.async.bulk*, completion through an mbarrier) ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MSN_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MSN_DONE_%=;\n"
      "bra MSN_WAIT_%=;\n"
      "MSN_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, both sides 16 B aligned)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 3-D tiled tensor copy global -> shared through a CUtensorMap (out-of-range elements read as 0)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct TileId {
  int n, y, x0;
};
__device__ __forceinline__ TileId decode_tile(int tile, const FusedArgs& a) {
  TileId t;
  const int xt = tile % a.tiles_x;
  tile /= a.tiles_x;
  t.y = tile % a.g.h;
  t.n = tile / a.g.h;
  t.x0 = xt * kTile;
  return t;
}

// Asynchronously copies the right-image row data of `t` into one staging buffer:
// census codes and stats of the D+31(+slack) columns the tile can touch and the five
// float rows of the ZSAD/NCC windows (16-byte cp.async; the float rows start at a
// 4-float aligned column, the 0..3 float shift is returned through *shift).
template <class L>
__device__ __forceinline__ void stage_right(const FusedArgs& a, const TileId& t, unsigned char* buf) {
  const FusedGeom& g = a.g;
  const int D = g.D;
  const int RWn = D + kTile - 1 + L::kSl;
  const int XbaseP = t.x0 + g.bwl - (D - 1) - L::kSl + g.padL;
  const int Yp = t.y + g.bh + kPadT;
  const size_t img_off = (size_t)t.n * g.img_px();
  const uint4* gd = a.descR + img_off + (size_t)Yp * g.Wp + XbaseP;
  const uint4* gs = reinterpret_cast<const uint4*>(a.statR + img_off + (size_t)Yp * g.Wp + XbaseP);
  uint4* s_desc = reinterpret_cast<uint4*>(buf + L::st_desc);
  uint4* s_stat = reinterpret_cast<uint4*>(buf + L::st_stat);
  float* s_rf = reinterpret_cast<float*>(buf + L::st_rf);
  for (int i = threadIdx.x; i < RWn; i += kWarps * 32) {
    cp_async16(s_desc + i, gd + i);
    cp_async16(s_stat + i, gs + i);
  }
  const int fstart = (XbaseP - 2) & ~3;                 // aligned first float column
  const int nvec = (RWn + 4 + 3 + 3) >> 2;              // 16-byte groups per row (covers any shift)
  for (int i = threadIdx.x; i < 5 * nvec; i += kWarps * 32) {
    const int r = i / nvec, v = i - r * nvec;
    cp_async16(s_rf + r * L::RWF + 4 * v, a.fR + img_off + (size_t)(Yp - 2 + r) * g.Wp + fstart + 4 * v);
  }
}

// Same data through the TMA engine: seven 1-D bulk copies (plus seven tiny ones for the 32
// left pixels when with_left), issued by a single thread, completion counted in bytes on `bar`.
template <class L>
__device__ __forceinline__ void stage_rows_tma(const FusedArgs& a, const TileId& t, unsigned char* buf,
                                               unsigned long long* bar, bool with_left) {
  const FusedGeom& g = a.g;
  const int D = g.D;
  const int RWn = D + kTile - 1 + L::kSl;
  const int XbaseP = t.x0 + g.bwl - (D - 1) - L::kSl + g.padL;
  const int Yp = t.y + g.bh + kPadT;
  const size_t img_off = (size_t)t.n * g.img_px();
  const int fstart = (XbaseP - 2) & ~3;
  const int nvec = (RWn + 4 + 3 + 3) >> 2;
  const unsigned row_bytes = (unsigned)RWn * 16u, frow_bytes = (unsigned)nvec * 16u;
  const int XpL = t.x0 + g.bwl + g.padL;                 // padded column of the tile's first left pixel
  const int lstart = (XpL - 2) & ~3;
  mbar_expect_tx(bar, 2u * row_bytes + 5u * frow_bytes + (with_left ? (2u * kTile * 16u + 5u * L::LF * 4u) : 0u));
  bulk_load(buf + L::st_desc, a.descR + img_off + (size_t)Yp * g.Wp + XbaseP, row_bytes, bar);
  bulk_load(buf + L::st_stat, a.statR + img_off + (size_t)Yp * g.Wp + XbaseP, row_bytes, bar);
  float* s_rf = reinterpret_cast<float*>(buf + L::st_rf);
#pragma unroll
  for (int r = 0; r < 5; ++r)
    bulk_load(s_rf + r * L::RWF, a.fR + img_off + (size_t)(Yp - 2 + r) * g.Wp + fstart, frow_bytes, bar);
  if (with_left) {
    bulk_load(buf + L::st_ldesc, a.descL + img_off + (size_t)Yp * g.Wp + XpL, kTile * 16u, bar);
    bulk_load(buf + L::st_lstat, a.statL + img_off + (size_t)Yp * g.Wp + XpL, kTile * 16u, bar);
    float* s_lf = reinterpret_cast<float*>(buf + L::st_lf);
#pragma unroll
    for (int r = 0; r < 5; ++r)
      bulk_load(s_lf + r * L::LF, a.fL + img_off + (size_t)(Yp - 2 + r) * g.Wp + lstart, L::LF * 4u, bar);
  }
}

// The tile's D x 32 SAD-of-Sobel costs: ONE 3-D tensor copy straight into parking plane 1.
// They come from DRAM and are only needed after phase 1, hence their own barrier.
__device__ __forceinline__ void stage_sad_tma(const FusedArgs& a, const CUtensorMap* sad_map, const TileId& t,
                                              float* park_plane1, unsigned long long* bar_sad) {
  const FusedGeom& g = a.g;
  mbar_expect_tx(bar_sad, (unsigned)g.D * kTile * 4u);
  tma_load_3d(park_plane1, sad_map, t.x0 + g.bwl + g.sxo, t.y + g.bh, t.n * g.D, bar_sad);  // inner coordinate % 4 == 0
}

// The lane's own left-image data: census code, stats, 5x5 float window.
struct LeftRegs {
  uint4 desc;
  uint4 stat;
  float px[5][5];
};
__device__ __forceinline__ void load_left(const FusedArgs& a, const TileId& t, int lane, LeftRegs& lr) {
  const FusedGeom& g = a.g;
  const int Yp = t.y + g.bh + kPadT;
  const int Xp = t.x0 + lane + g.bwl + g.padL;
  const size_t img_off = (size_t)t.n * g.img_px();
  lr.desc = __ldg(a.descL + img_off + (size_t)Yp * g.Wp + Xp);
  lr.stat = __ldg(reinterpret_cast<const uint4*>(a.statL + img_off + (size_t)Yp * g.Wp + Xp));
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const float* gf = a.fL + img_off + (size_t)(Yp - 2 + r) * g.Wp + (Xp - 2);
#pragma unroll
    for (int c = 0; c < 5; ++c) lr.px[r][c] = __ldg(gf + c);
  }
}

// Left data from the persistent kernel's staging buffer (delivered by TMA).
template <class L>
__device__ __forceinline__ void load_left_smem(const FusedArgs& a, const TileId& t, const unsigned char* buf, int lane,
                                               LeftRegs& lr) {
  const int shift = (t.x0 + a.g.bwl + a.g.padL - 2) & 3;
  lr.desc = reinterpret_cast<const uint4*>(buf + L::st_ldesc)[lane];
  lr.stat = reinterpret_cast<const uint4*>(buf + L::st_lstat)[lane];
  const float* s_lf = reinterpret_cast<const float*>(buf + L::st_lf) + shift + lane;
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) lr.px[r][c] = s_lf[r * L::LF + c];
}

// Everything a tile does once its right-image rows are staged in `buf` (and every thread has
// passed the barrier that made them visible): phases 1-3 described at the top of the file.
// `lr` holds the lane's left-image data; the SAD-of-Sobel tile is awaited on bar_sad (TMA).
template <class L, bool kTma>
__device__ __forceinline__ void tile_compute(const FusedArgs& a, unsigned char* smem_raw, const TileId& t,
                                             const unsigned char* buf, const LeftRegs& lr,
                                             unsigned long long* bar_sad, unsigned sad_parity) {
  const FusedGeom& g = a.g;
  const int D = g.D;
  float* s_red = reinterpret_cast<float*>(smem_raw + L::off_red);    // [8][4][32]
  float* s_min = reinterpret_cast<float*>(smem_raw + L::off_min);    // [4][32]
  float* s_inv = reinterpret_cast<float*>(smem_raw + L::off_inv);    // [4][32]
  float* s_lut = reinterpret_cast<float*>(smem_raw + L::off_lut);    // [128]
  float* s_par = reinterpret_cast<float*>(smem_raw + L::off_par);    // [3][DS][32]
  uint8_t* s_cen = smem_raw + L::off_cen;                            // [DS][32]
  constexpr int PS = L::DS * 32;                                     // floats per parked matcher
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = g.H, W = g.W;
  const int d_lo = warp * a.DC;
  const int d_end = min(D, d_lo + a.DC);  // real disparities of this warp: [d_lo, d_end)
  const size_t plane = (size_t)g.h * g.w;
  const size_t chan = plane * D;
  // 128-bit stores need 16-byte aligned rows
  const bool vec_ok = ((g.w & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
  // AML-phase mapping: thread = (pixel quad q, disparity lane dl); d = dl, dl+32, ...
  const int q4 = (threadIdx.x & 7) * 4;
  const int dl = threadIdx.x >> 3;
  const int X = t.x0 + lane + g.bwl;      // bordered image column of this lane
  const int Y = t.y + g.bh;               // bordered image row
  const uint4 ld = lr.desc;
  const RStat ls = *reinterpret_cast<const RStat*>(&lr.stat);
  float at[5][5];   // (L - mL), hoisted over all d   (matchers.cpp:503)
  float l3[3][3];   // centre 3x3 of L as float for NCC
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      at[r][c] = __fsub_rn(lr.px[r][c], ls.mean);
      if (r >= 1 && r <= 3 && c >= 1 && c <= 3) l3[r - 1][c - 1] = lr.px[r][c];
    }
  {
    // validity: cost(y,x,d) exists iff the window origin is inside and x - wc >= d (and d < D)
    const int dmax_cen = min(D - 1, (Y >= 5 && Y < H - 6 && X >= 5 && X < W - 6) ? X - 5 : -1);
    const int dmax_ncc = min(D - 1, (Y >= 1 && Y < H - 2 && X >= 1 && X < W - 2) ? X - 1 : -1);
    const int dmax_sad = min(D - 1, (Y >= 2 && Y < H - 3 && X >= 2 && X < W - 3) ? X - 2 : -1);

    // ---- phase 1: raw costs into the parking planes, per-pixel minima ----------------
    {
      const uint4* s_desc = reinterpret_cast<const uint4*>(buf + L::st_desc);
      const uint4* s_stat = reinterpret_cast<const uint4*>(buf + L::st_stat);
      const float* s_rf = reinterpret_cast<const float*>(buf + L::st_rf);
      // shared index of right column X - d is ir = lane + L::kSl + (D-1) - d; falls by one per step
      const int XbaseP = t.x0 + g.bwl - (D - 1) - L::kSl + g.padL;
      const int shift = (XbaseP - 2) & 3;
      const int ir0 = lane + L::kSl + (D - 1) - d_lo;
      const float* rfp = s_rf + shift + ir0;
      const uint4* dscp = s_desc + ir0;
      const uint4* sttp = s_stat + ir0;
      float rw[5][5];  // sliding 5x5 right window; logical column c lives in rw[.][(c - s) mod 5]
#pragma unroll
      for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) rw[r][c] = rfp[r * L::RWF + c];
      int min_cen = 255;
      float min_ncc = kFill, min_sob = kFill, min_sad = kFill;

      for (int base = 0; base < a.DC; base += 5) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int d = d_lo + base + k;
          const int ds = min(d, D);  // dummy steps (d >= D) park into the scratch plane
          float* park = s_par + ds * 32 + lane;
          const uint4 rd = *dscp;
          const uint4 rs_raw = *sttp;
          const RStat rs = *reinterpret_cast<const RStat*>(&rs_raw);

          // census: Hamming distance of the packed codes (matchers.cpp:323-337)
          const int cen = __popc(ld.x ^ rd.x) + __popc(ld.y ^ rd.y) + __popc(ld.z ^ rd.z) + __popc(ld.w ^ rd.w);
          const int cen_b = (d <= dmax_cen) ? cen : 255;

          // NCC: P exact in fp32 (< 2^24); scaling in fp64 left to right (matchers.cpp:200-201)
          float P = 0.f;
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) P = __fmaf_rn(l3[r][c], rw[r + 1][(c + 1 - k + 5) % 5], P);
          const float num = __fmaf_rn(9.0f, P, -__fmul_rn(ls.A, rs.A));
          const double tt = __dmul_rn(__dmul_rn(-(double)num, ls.C), rs.C);
          float ncc = (float)tt;
          ncc = (fabsf(ncc) <= 3.0e38f) ? ncc : 1.0f;  // either C was inf (flat window), :196,204
          ncc = (d <= dmax_ncc) ? ncc : kFill;

          // ZSAD: 25 taps row-major, ((L - mL) - R) + mR, sequential fp32 (matchers.cpp:499-506)
          float z = 0.f;
#pragma unroll
          for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) {
              const float u = __fadd_rn(__fsub_rn(at[r][c], rw[r][(c - k + 5) % 5]), rs.mean);
              z = __fadd_rn(z, fabsf(u));
            }
          z = (d <= dmax_sad) ? z : kFill;

          s_cen[ds * 32 + lane] = (uint8_t)cen_b;
          park[0] = ncc;
          park[2 * PS] = z;
          min_cen = min(min_cen, cen_b);
          min_ncc = fminf(min_ncc, ncc);
          min_sad = fminf(min_sad, z);
          // slide the window: next step's new left column
          --rfp; --dscp; --sttp;
#pragma unroll
          for (int r = 0; r < 5; ++r) rw[r][(4 - k) % 5] = rfp[r * L::RWF];
        }
      }
      // SAD-of-Sobel costs of this warp's disparities (delivered by TMA / cp.async while the loop
      // above ran): replace what lies outside the valid region by fill, take the minimum
      if (kTma) mbar_wait(bar_sad, sad_parity);
      {
        float* sp = s_par + PS + d_lo * 32 + lane;
#pragma unroll 4
        for (int d = d_lo; d < d_end; ++d, sp += 32) {
          float v = *sp;
          if (d > dmax_sad) {
            v = kFill;
            *sp = v;
          }
          min_sob = fminf(min_sob, v);
        }
      }
      s_red[(warp * 4 + 0) * 32 + lane] = (min_cen == 255) ? kFill : (float)min_cen;
      s_red[(warp * 4 + 1) * 32 + lane] = min_ncc;
      s_red[(warp * 4 + 2) * 32 + lane] = min_sob;
      s_red[(warp * 4 + 3) * 32 + lane] = min_sad;
    }
    __syncthreads();
    if (threadIdx.x < 128) {  // minima across the 8 warps
      float v = kFill;
#pragma unroll
      for (int wv = 0; wv < kWarps; ++wv) v = fminf(v, s_red[wv * 128 + threadIdx.x]);
      s_min[threadIdx.x] = v;
    }
    __syncthreads();

    // ---- phase 2 (warp-specialised, the two halves run concurrently and never write the
    //      parking planes, so there is no hazard between them):
    //      warps 0-3  one thread per (pixel, matcher): AML denominator, exponentials
    //                 evaluated on the fly and added in the reference's order -- sequential
    //                 fp32 over d (featextract.cpp:444-447);
    //      warps 4-7  thread = (pixel quad, d): channels 0-3 (cbmv_generator.py:283-287)
    //                 normalised and stored as 128-bit row segments. -----------------------
    float* orow = a.out + (size_t)t.n * 8 * chan + (size_t)t.y * g.w + (t.x0 + q4);
    const int nlive = min(4, g.w - (t.x0 + q4));  // live pixels of this quad (<= 0: none)
    if (warp < 4) {
      const float mm = s_min[warp * 32 + lane];
      float den = 0.f;
      const int Dfull = D & ~7;   // groups of 8 without guards, then a guarded tail
      if (warp == 0) {
        const int mc = (mm == kFill) ? 0 : (int)mm;
        const uint8_t* c = s_cen + lane;
        for (int d0 = 0; d0 < Dfull; d0 += 8, c += 8 * 32) {
          float ev[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) ev[j] = s_lut[min((int)c[j * 32] - mc, 127)];
#pragma unroll
          for (int j = 0; j < 8; ++j) den = __fadd_rn(den, ev[j]);
        }
        for (int d = Dfull; d < D; ++d, c += 32) den = __fadd_rn(den, s_lut[min((int)c[0] - mc, 127)]);
      } else {
        const float kq = (warp == 1) ? a.k_ncc : a.k_sad;
        const float* e = s_par + (warp - 1) * PS + lane;
        for (int d0 = 0; d0 < Dfull; d0 += 8, e += 8 * 32) {
          float ev[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) ev[j] = aml_e(e[j * 32], mm, kq);
#pragma unroll
          for (int j = 0; j < 8; ++j) den = __fadd_rn(den, ev[j]);
        }
        for (int d = Dfull; d < D; ++d, e += 32) den = __fadd_rn(den, aml_e(e[0], mm, kq));
      }
      s_inv[warp * 32 + lane] = (mm == kFill) ? 0.f : 1.0f / den;
    } else {
      for (int d = dl - 16; d < D; d += 16) {   // dl in [16,32) for warps 4-7
        const float* e0 = s_par + d * 32 + q4;
        const uchar4 cb = *reinterpret_cast<const uchar4*>(s_cen + d * 32 + q4);
        float4 c0;
        c0.x = (cb.x == 255) ? 1.0f : div120_exact(int_to_float_small(cb.x));
        c0.y = (cb.y == 255) ? 1.0f : div120_exact(int_to_float_small(cb.y));
        c0.z = (cb.z == 255) ? 1.0f : div120_exact(int_to_float_small(cb.z));
        c0.w = (cb.w == 255) ? 1.0f : div120_exact(int_to_float_small(cb.w));
        const float4 v1 = *reinterpret_cast<const float4*>(e0);
        const float4 v2 = *reinterpret_cast<const float4*>(e0 + PS);
        const float4 v3 = *reinterpret_cast<const float4*>(e0 + 2 * PS);
        const float4 c1 = make_float4(normalise_cost(v1.x, 1), normalise_cost(v1.y, 1), normalise_cost(v1.z, 1),
                                      normalise_cost(v1.w, 1));
        const float4 c2 = make_float4(normalise_cost(v2.x, 2), normalise_cost(v2.y, 2), normalise_cost(v2.z, 2),
                                      normalise_cost(v2.w, 2));
        const float4 c3 = make_float4(normalise_cost(v3.x, 3), normalise_cost(v3.y, 3), normalise_cost(v3.z, 3),
                                      normalise_cost(v3.w, 3));
        float* o = orow + (size_t)d * plane;
        if (vec_ok && nlive == 4) {
          st_stream4(o, c0);
          st_stream4(o + chan, c1);
          st_stream4(o + 2 * chan, c2);
          st_stream4(o + 3 * chan, c3);
        } else {
          const float cc[4][4] = {{c0.x, c0.y, c0.z, c0.w}, {c1.x, c1.y, c1.z, c1.w}, {c2.x, c2.y, c2.z, c2.w},
                                  {c3.x, c3.y, c3.z, c3.w}};
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < nlive) st_stream(o + ch * chan + i, cc[ch][i]);
        }
      }
    }
    __syncthreads();

    // ---- phase 3: channels 4-7 = exp(-(c-m)^2/sigma) / den, 128-bit row segments --------
    {
      const float4 m_cen4 = *reinterpret_cast<const float4*>(s_min + q4);
      const float4 m_ncc4 = *reinterpret_cast<const float4*>(s_min + 32 + q4);
      const float4 m_sob4 = *reinterpret_cast<const float4*>(s_min + 64 + q4);
      const float4 m_sad4 = *reinterpret_cast<const float4*>(s_min + 96 + q4);
      const float4 i0 = *reinterpret_cast<const float4*>(s_inv + q4);
      const float4 i1 = *reinterpret_cast<const float4*>(s_inv + 32 + q4);
      const float4 i2 = *reinterpret_cast<const float4*>(s_inv + 64 + q4);
      const float4 i3 = *reinterpret_cast<const float4*>(s_inv + 96 + q4);
      const int mcx = (m_cen4.x == kFill) ? 0 : (int)m_cen4.x, mcy = (m_cen4.y == kFill) ? 0 : (int)m_cen4.y;
      const int mcz = (m_cen4.z == kFill) ? 0 : (int)m_cen4.z, mcw = (m_cen4.w == kFill) ? 0 : (int)m_cen4.w;
      float* arow = orow + 4 * chan;
      for (int d = dl; d < D; d += 32) {
        const float* e0 = s_par + d * 32 + q4;
        const uchar4 cb = *reinterpret_cast<const uchar4*>(s_cen + d * 32 + q4);
        float4 a0;
        a0.x = s_lut[min((int)cb.x - mcx, 127)] * i0.x;
        a0.y = s_lut[min((int)cb.y - mcy, 127)] * i0.y;
        a0.z = s_lut[min((int)cb.z - mcz, 127)] * i0.z;
        a0.w = s_lut[min((int)cb.w - mcw, 127)] * i0.w;
        const float4 v1 = *reinterpret_cast<const float4*>(e0);
        const float4 v2 = *reinterpret_cast<const float4*>(e0 + PS);
        const float4 v3 = *reinterpret_cast<const float4*>(e0 + 2 * PS);
        const float4 a1 = make_float4(aml_e(v1.x, m_ncc4.x, a.k_ncc) * i1.x, aml_e(v1.y, m_ncc4.y, a.k_ncc) * i1.y,
                                      aml_e(v1.z, m_ncc4.z, a.k_ncc) * i1.z, aml_e(v1.w, m_ncc4.w, a.k_ncc) * i1.w);
        const float4 a2 = make_float4(aml_e(v2.x, m_sob4.x, a.k_sad) * i2.x, aml_e(v2.y, m_sob4.y, a.k_sad) * i2.y,
                                      aml_e(v2.z, m_sob4.z, a.k_sad) * i2.z, aml_e(v2.w, m_sob4.w, a.k_sad) * i2.w);
        const float4 a3 = make_float4(aml_e(v3.x, m_sad4.x, a.k_sad) * i3.x, aml_e(v3.y, m_sad4.y, a.k_sad) * i3.y,
                                      aml_e(v3.z, m_sad4.z, a.k_sad) * i3.z, aml_e(v3.w, m_sad4.w, a.k_sad) * i3.w);
        float* o = arow + (size_t)d * plane;
        if (vec_ok && nlive == 4) {
          st_stream4(o, a0);
          st_stream4(o + chan, a1);
          st_stream4(o + 2 * chan, a2);
          st_stream4(o + 3 * chan, a3);
        } else {
          const float cc[4][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w}, {a2.x, a2.y, a2.z, a2.w},
                                  {a3.x, a3.y, a3.z, a3.w}};
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < nlive) st_stream(o + ch * chan + i, cc[ch][i]);
        }
      }
    }

  }
}

// One CTA per tile.
// kTma: right-image rows and the SAD-of-Sobel tile arrive through cp.async.bulk /
// cp.async.bulk.tensor (D <= 256: box limit); otherwise through LDGSTS.  The tensor copy's
// inner coordinate must be a multiple of 16 bytes (measured: anything else raises an
// illegal-instruction fault), which is why the scratch is stored with column offset sxo.
template <int DMAX, bool kTma>
__global__ void __launch_bounds__(kWarps * 32, 2)
ms_fused_kernel(const FusedArgs a, const __grid_constant__ CUtensorMap sad_map) {
  using L = Lay<DMAX>;
  extern __shared__ __align__(128) unsigned char smem_raw[];  // TMA destinations need 128 B
  const FusedGeom& g = a.g;
  const int D = g.D;
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem_raw + L::off_bar);  // [0] rows, [1] sadsob tile
  float* s_lut = reinterpret_cast<float*>(smem_raw + L::off_lut);
  float* s_par = reinterpret_cast<float*>(smem_raw + L::off_par);
  constexpr int PS = L::DS * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TileId t = decode_tile(blockIdx.x, a);
  unsigned char* buf = smem_raw + L::off_stage;
  if (kTma) {
    if (threadIdx.x == 0) {
      mbar_init(&s_bar[0], 1);
      mbar_init(&s_bar[1], 1);
      stage_sad_tma(a, &sad_map, t, s_par + PS, &s_bar[1]);
      stage_rows_tma<L>(a, t, buf, &s_bar[0], false);
    }
  } else {
    // sadsob costs of this lane's own disparities: async global -> parked plane 1
    const int d_lo = warp * a.DC, d_end = min(D, d_lo + a.DC);
    const size_t splane = (size_t)g.H * g.Ws;
    const float* src = a.sadsob + ((size_t)t.n * D * g.H + (t.y + g.bh)) * g.Ws + (t.x0 + lane + g.bwl + g.sxo) +
                       (size_t)d_lo * splane;
    float* dst = s_par + PS + d_lo * 32 + lane;
    for (int d = d_lo; d < d_end; ++d, src += splane, dst += 32) cp_async4(dst, src);
    stage_right<L>(a, t, buf);
  }
  if (threadIdx.x < 128) {
    const int kk = threadIdx.x;
    s_lut[kk] = (kk <= 120) ? ex2_approx(-(float)(kk * kk) * a.k_cen) : 0.f;
  }
  LeftRegs lr;
  load_left(a, t, lane, lr);
  __syncthreads();                       // barrier init + LUT visible to everyone
  if (kTma) mbar_wait(&s_bar[0], 0);
  else {
    cp_async_wait_all();
    __syncthreads();
  }
  tile_compute<L, kTma>(a, smem_raw, t, buf, lr, &s_bar[1], 0);
}

// Persistent variant (TMA only): gridDim.x resident CTAs walk the tiles with stride gridDim.x;
// while a tile is in phase 1 the TMA engine already fetches the next tile's right-image rows
// and left-pixel data into the other staging buffer.  Selected with MSNETS_FUSED_PERSISTENT=1.
template <int DMAX, int SLACK>
__global__ void __launch_bounds__(kWarps * 32, 2)
ms_fused_persistent_kernel(const FusedArgs a, const __grid_constant__ CUtensorMap sad_map) {
  using L = Lay<DMAX, SLACK, 2>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem_raw + L::off_bar);  // [0],[1] rows per buffer, [2] sadsob
  float* s_lut = reinterpret_cast<float*>(smem_raw + L::off_lut);
  float* s_par = reinterpret_cast<float*>(smem_raw + L::off_par);
  constexpr int PS = L::DS * 32;
  const int lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  if (tile >= a.num_tiles) return;
  TileId t = decode_tile(tile, a);
  if (a.stagger_ns > 0.f) {
    // desynchronise the resident CTAs: pseudo-random start offset within one tile period
    const unsigned frac = (blockIdx.x * 2654435761u) >> 26;            // 0..63
    unsigned ns = (unsigned)(a.stagger_ns * (frac / 64.0f));
    while (ns > 0) {
      const unsigned step = ns > 50000u ? 50000u : ns;
      __nanosleep(step);
      ns -= step;
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_init(&s_bar[2], 1);
    stage_rows_tma<L>(a, t, smem_raw + L::off_stage, &s_bar[0], true);
  }
  if (threadIdx.x < 128) {
    const int kk = threadIdx.x;
    s_lut[kk] = (kk <= 120) ? ex2_approx(-(float)(kk * kk) * a.k_cen) : 0.f;
  }
  __syncthreads();
  for (int it = 0; tile < a.num_tiles; ++it, tile += gridDim.x) {
    const int b = it & 1;
    unsigned char* buf = smem_raw + L::off_stage + (size_t)b * L::st_bytes;
    const int next = tile + gridDim.x;
    TileId tn = t;
    if (threadIdx.x == 0) {
      // this tile's SAD-of-Sobel costs (the parking planes are free: end-of-tile barrier below)
      stage_sad_tma(a, &sad_map, t, s_par + PS, &s_bar[2]);
      // next tile's rows into the other buffer (last read during the previous tile's phase 1)
      if (next < a.num_tiles) {
        tn = decode_tile(next, a);
        stage_rows_tma<L>(a, tn, smem_raw + L::off_stage + (size_t)(b ^ 1) * L::st_bytes, &s_bar[b ^ 1], true);
      }
    }
    mbar_wait(&s_bar[b], (it >> 1) & 1);
    LeftRegs lr;
    load_left_smem<L>(a, t, buf, lane, lr);
    tile_compute<L, true>(a, smem_raw, t, buf, lr, &s_bar[2], it & 1);
    __syncthreads();  // parking planes, minima and staging buffer b are free again
    if (next < a.num_tiles) t = decode_tile(next, a);
  }
}

// Optional per-kernel timing (msn_profile_enable): CUDA events recorded on the launch
// stream between the kernels of the fused sequence, read back by msn_profile_read.
struct ProfRec { cudaEvent_t ev[4]; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;

}  // namespace

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry-point lookup (no libcuda link)
static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (PFN_encodeTiled)p;
  }();
  return fn;
}
static bool persistent_enabled() {
  const char* e = getenv("MSNETS_FUSED_PERSISTENT");
  return e && e[0] == '1';
}
static bool tma_disabled() {
  const char* e = getenv("MSNETS_NO_TMA");
  return e && e[0] == '1';
}

int profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  return 0;
}

int profile_read(double* prep_ms, double* sadsob_ms, double* fused_ms, int* calls) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double t[3] = {0, 0, 0};
  for (ProfRec& r : g_prof) {
    MSN_CUDA_OK(cudaEventSynchronize(r.ev[3]));
    for (int i = 0; i < 3; ++i) {
      float ms = 0.f;
      MSN_CUDA_OK(cudaEventElapsedTime(&ms, r.ev[i], r.ev[i + 1]));
      t[i] += ms;
    }
    for (int i = 0; i < 4; ++i) cudaEventDestroy(r.ev[i]);
  }
  if (prep_ms) *prep_ms = t[0];
  if (sadsob_ms) *sadsob_ms = t[1];
  if (fused_ms) *fused_ms = t[2];
  if (calls) *calls = (int)g_prof.size();
  g_prof.clear();
  return 0;
}

bool fused_supported(const msn_ms_params* p, int Dn) {
  return p->censw == kCensW && p->nccw == kNccW && p->sadw == kSadW && p->sobelw == kSadW && p->lr == 0 &&
         Dn == p->ndisp && p->ndisp <= kMaxFusedD;  // (image width is checked at launch)
}

size_t fused_workspace_bytes(int N, int H, int W, int Dn, const msn_ms_params* p) {
  (void)Dn;
  FusedGeom g = make_geom(N, H, W, p);
  FusedWs ws;
  ws.carve(nullptr, g);
  return ws.total;
}

int launch_ms_fused(const uint8_t* d_left, const uint8_t* d_right, int N, int H, int W, const msn_ms_params* p,
                    float* d_out, char* workspace, cudaStream_t s) {
  if (N == 0) return 0;
  FusedGeom g = make_geom(N, H, W, p);
  FusedWs ws;
  ws.carve(workspace, g);
  MSN_REQUIRE(2 * N <= 65535 && g.Hp <= 65535, "ms_features: batch or image too large for one launch");

  bool prof;
  ProfRec rec;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof = g_prof_on;
  }
  if (prof) {
    for (int i = 0; i < 4; ++i) MSN_CUDA_OK(cudaEventCreate(&rec.ev[i]));
    MSN_CUDA_OK(cudaEventRecord(rec.ev[0], s));
  }
  // zero the row padding of the Sobel images (the scan reads up to 31 rows past H)
  for (int i = 0; i < 2; ++i)
    MSN_CUDA_OK(cudaMemsetAsync(ws.sob[i], 0, (size_t)N * (H + kSadRowPad) * g.Ws * sizeof(float), s));
  dim3 pgrid(div_up(g.Wp, 128), g.Hp, 2 * N);
  ms_prep_kernel<<<pgrid, 128, 0, s>>>(d_left, d_right, g, ws.desc[0], ws.desc[1], ws.stat[0], ws.stat[1],
                                       ws.fimg[0], ws.fimg[1], ws.sob[0], ws.sob[1]);
  MSN_LAUNCH_OK();
  if (prof) MSN_CUDA_OK(cudaEventRecord(rec.ev[1], s));
  if (launch_sadsob5_padded(ws.sob[0], ws.sob[1], N, H, W, g.D, 0, ws.sadsob + g.sxo, ws.sad_ws, s)) return 1;
  if (prof) MSN_CUDA_OK(cudaEventRecord(rec.ev[2], s));

  FusedArgs a;
  a.g = g;
  a.descL = ws.desc[0]; a.descR = ws.desc[1];
  a.statL = ws.stat[0]; a.statR = ws.stat[1];
  a.fL = ws.fimg[0]; a.fR = ws.fimg[1];
  a.sadsob = ws.sadsob;
  a.out = d_out;
  a.k_cen = aml_scale(p->cens_sigma);
  a.k_ncc = aml_scale(p->ncc_sigma);
  a.k_sad = aml_scale(p->sad_sigma);
  a.DC = 5 * (((g.D + kWarps - 1) / kWarps + 4) / 5);
  a.tiles_x = (g.w + kTile - 1) / kTile;
  const long long tiles = (long long)N * g.h * a.tiles_x;
  MSN_REQUIRE(tiles <= 2147483647LL, "ms_features: too many tiles for one launch");
  a.num_tiles = (int)tiles;
  {
    const char* e = getenv("MSNETS_FUSED_STAGGER_NS");
    a.stagger_ns = e ? (float)atof(e) : 0.f;
  }
  // TMA path: 3-D tensor map over the SAD-of-Sobel scratch [N*D][H][W], box 32 x 1 x D
  CUtensorMap sad_map;
  memset(&sad_map, 0, sizeof(sad_map));
  bool use_tma = g.D <= 256 && !tma_disabled();
  if (use_tma) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
      use_tma = false;
    } else {
      const cuuint64_t gdim[3] = {(cuuint64_t)g.Ws, (cuuint64_t)H, (cuuint64_t)N * g.D};
      const cuuint64_t gstr[2] = {(cuuint64_t)g.Ws * 4, (cuuint64_t)H * g.Ws * 4};
      const cuuint32_t box[3] = {(cuuint32_t)kTile, 1u, (cuuint32_t)g.D};
      const cuuint32_t estr[3] = {1u, 1u, 1u};
      const CUresult rc = enc(&sad_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ws.sadsob, gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) use_tma = false;
    }
  }
  // experimental persistent variant (MSNETS_FUSED_PERSISTENT=1): needs TMA and at most 8 dummy steps
  if (use_tma && persistent_enabled() && 8 * a.DC - g.D <= 8 && g.D <= 192) {
    int dev = 0, sms = 0;
    MSN_CUDA_OK(cudaGetDevice(&dev));
    MSN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
#define MSN_PERSIST_CASE(DMAX)                                                                        \
  if (g.D <= DMAX) {                                                                                  \
    const size_t smem = Lay<DMAX, 8, 2>::bytes;                                                       \
    MSN_CUDA_OK(cudaFuncSetAttribute(ms_fused_persistent_kernel<DMAX, 8>,                             \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    int per_sm = 0;                                                                                   \
    MSN_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ms_fused_persistent_kernel<DMAX, 8>, \
                                                              kWarps * 32, smem));                    \
    MSN_REQUIRE(per_sm >= 1, "ms_features: persistent kernel does not fit");                          \
    const long long grid = (long long)per_sm * sms < tiles ? (long long)per_sm * sms : tiles;         \
    ms_fused_persistent_kernel<DMAX, 8><<<(unsigned)grid, kWarps * 32, smem, s>>>(a, sad_map);        \
  } else
    MSN_PERSIST_CASE(64)
    MSN_PERSIST_CASE(128)
    MSN_PERSIST_CASE(192) {}
#undef MSN_PERSIST_CASE
    MSN_LAUNCH_OK();
    if (prof) {
      MSN_CUDA_OK(cudaEventRecord(rec.ev[3], s));
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_prof.push_back(rec);
    }
    return 0;
  }
#define MSN_FUSED_LAUNCH(DMAX, TMA)                                                                   \
  {                                                                                                   \
    const size_t smem = Lay<DMAX>::bytes;                                                             \
    MSN_CUDA_OK(cudaFuncSetAttribute(ms_fused_kernel<DMAX, TMA>,                                      \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    ms_fused_kernel<DMAX, TMA><<<(unsigned)tiles, kWarps * 32, smem, s>>>(a, sad_map);                \
  }
#define MSN_FUSED_CASE(DMAX)                                                                          \
  if (g.D <= DMAX) {                                                                                  \
    if (use_tma && DMAX <= 256) MSN_FUSED_LAUNCH(DMAX <= 256 ? DMAX : 256, true)                      \
    else MSN_FUSED_LAUNCH(DMAX, false)                                                                \
  } else
  MSN_FUSED_CASE(64)
  MSN_FUSED_CASE(128)
  MSN_FUSED_CASE(192)
  MSN_FUSED_CASE(256)
  MSN_FUSED_CASE(384)
  MSN_FUSED_CASE(448) { return fail("ms_features: D=%d exceeds the fused kernel's limit", g.D); }
#undef MSN_FUSED_CASE
#undef MSN_FUSED_LAUNCH
  MSN_LAUNCH_OK();
  if (prof) {
    MSN_CUDA_OK(cudaEventRecord(rec.ev[3], s));
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(rec);
  }
  return 0;
}

}  // namespace msn
