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
//                        float copy of the pixel, Sobel response.  The right image's planes
//                        are written DE-INTERLEAVED (even columns, then odd columns of a row).
//   2. sadsob scan       (sadsob.cu) -> raw SAD-of-Sobel volume [N][D][H][W]; its fp32
//                        summed-area table needs whole-row sequential scans, so it
//                        cannot live inside an x-tile.
//   3. ms_fused_kernel   one CTA = one output row y x 32 pixels x ALL D (ms_fused_tile.cuh).
//
// Bounding resources, in the order they bite (ncu, profiles/): the L1 / shared-memory data pipe
// (one 128-byte wavefront per cycle per SM), then HBM writes (32 B per voxel); the FP32 pipe
// (ZSAD: 75 ordered adds per voxel, issued as packed FADD2) is third.
#include "ms_fused.cuh"

#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "feature_math.cuh"

namespace msn {

namespace {

constexpr int kTile = 32;     // pixels per tile: one output row segment of 128 bytes
constexpr int kPadT = 2;      // padded rows above/below (ZSAD halo)
constexpr int kPadR = 48;     // padded columns to the right (tile overhang + halo + copy granules)
constexpr int kCensW = 11, kNccW = 3, kSadW = 5;
constexpr int kMaxFusedD = 384;   // more disparities: slabs of kFusedSlabD (capi.cu)

struct __align__(16) RStat {
  float mean;  // ZSAD window mean (matchers.cpp:482)
  float A;     // NCC window sum, exact integer <= 2295 (matchers.cpp:140)
  double C;    // 1/sqrt(9*B - A*A) in fp64 (matchers.cpp:146)
};

struct FusedGeom {
  int N, H, W, h, w, bh, bwl;
  int D;    // disparities handled by this launch: [d0, d0 + D)
  int d0;   // first disparity (> 0 only for a disparity slab)
  int Hp, Wp, padL;
  int sxo, Ws;   // SAD-of-Sobel scratch: column offset and row pitch (tile starts land on 16 B)
  __host__ __device__ size_t img_px() const { return (size_t)Hp * Wp; }
};

FusedGeom make_geom(int N, int H, int W, const msn_ms_params* p) {
  FusedGeom g;
  g.N = N; g.H = H; g.W = W;
  g.d0 = p->d_count > 0 ? p->d_begin : 0;
  g.D = p->d_count > 0 ? p->d_count : p->ndisp;
  g.bh = p->board_h; g.bwl = p->board_w_left;
  g.h = H - 2 * p->board_h;
  g.w = W - p->board_w_left - p->board_w_right;
  g.padL = (g.d0 + g.D + 32 + 7) & ~7;      // the leftmost staged column (see stage_geo) stays >= 0
  g.Hp = H + 2 * kPadT;
  g.Wp = (W + g.padL + kPadR + 7) & ~7;     // multiple of 8: both halves of a de-interleaved row start on 16 B
  g.sxo = (4 - (g.bwl & 3)) & 3;            // x0 + bwl + sxo is a multiple of 4 for every tile
  g.Ws = sadsob_fast_pitch(W + g.sxo + kTile);  // compile-time pitch of the scan kernels (0: too wide)
  return g;
}

struct FusedWs {
  uint4* desc[2];   // census codes; [1] (right image) de-interleaved like every right-image plane
  RStat* statL;
  float* aR;        // right image: NCC window sum A
  double* cR;       // right image: NCC C
  float* meanR;     // right image: ZSAD window mean
  float* fimg[2];
  float* sob[2];
  float* sadsob;    // [N][D][H][Ws]
  void* sad_ws;
  size_t total;
  void carve(char* base, const FusedGeom& g) {
    size_t off = 0;
    auto take = [&](size_t bytes) { char* q = base ? base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return q; };
    const size_t np = (size_t)g.N * g.img_px() + 64;   // (+ slack: a half-row copy may run a granule past the last row)
    for (int i = 0; i < 2; ++i) desc[i] = (uint4*)take(np * sizeof(uint4));
    statL = (RStat*)take(np * sizeof(RStat));
    aR = (float*)take(np * sizeof(float));
    cR = (double*)take(np * sizeof(double));
    meanR = (float*)take(np * sizeof(float));
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
               float* __restrict__ aR, double* __restrict__ cR, float* __restrict__ meanR,
               float* __restrict__ fL, float* __restrict__ fR, float* __restrict__ sobL, float* __restrict__ sobR) {
  const int xp = blockIdx.x * blockDim.x + threadIdx.x;
  const int yp = blockIdx.y;
  const int n = blockIdx.z >> 1, side = blockIdx.z & 1;
  if (xp >= g.Wp) return;
  const int H = g.H, W = g.W;
  const uint8_t* img = (side ? right : left) + (size_t)n * H * W;
  const int X = xp - g.padL, Y = yp - kPadT;
  const bool inside = (X >= 0 && X < W && Y >= 0 && Y < H);
  const size_t rowo = (size_t)n * g.img_px() + (size_t)yp * g.Wp;

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
  if (side) {
    // right image: even columns first, then odd columns (ms_fused_tile.cuh: lanes own pixel pairs)
    const size_t po = rowo + (size_t)(xp >> 1) + (size_t)(xp & 1) * (g.Wp >> 1);
    descR[po] = make_uint4(w0, w1, w2, w3);
    aR[po] = st.A;
    cR[po] = st.C;
    meanR[po] = st.mean;
    fR[po] = pix;
  } else {
    const size_t po = rowo + xp;
    descL[po] = make_uint4(w0, w1, w2, w3);
    statL[po] = st;
    fL[po] = pix;
  }
}

// ----------------------------------------------------------------- fused --
struct FusedArgs {
  FusedGeom g;
  const uint4 *descL, *descR;
  const RStat* statL;
  const float* aR;
  const double* cR;
  const float* meanR;
  const float *fL, *fR;
  const float* sadsob;  // [N][D][H][Ws] (+ slack)
  float* out;           // [N][8][D][h][w]
  float* mins;          // slab phase A only: [N][mins_planes][h][w], planes 0-3 = per-pixel minima of this launch's disparities
  int out_channels;     // channel count of the output tensor (pair stride): 8, or 16 when the caller adds the right view
  int mins_planes;      // 4, or 8 when the caller adds the right view
  int out_D, out_d0;    // slab phase A: disparity count of the output tensor and where this launch's slab sits in it
  int mins_accumulate;  // slab phase A: fold into the minima already in `mins` (a later slab of the same volume)
  float k_cen, k_ncc, k_sad;
  float neg_zero;       // -0.0f (x + -0.0f == x exactly): an operand the compiler cannot fold
  int DC;               // disparity steps per d-group
  int tiles_x;
  int xflags;           // experiment switches (MSNETS_X)
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- TMA / bulk-copy helpers (cp.async.bulk*, completion through an mbarrier) ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Waits for the phase with the given parity.  try_wait suspends the warp in hardware (up to the
// hint, in ns) instead of polling, so waiting warps leave the issue slots to the working ones.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const uint32_t addr = smem_u32(bar);
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(1000000u)
        : "memory");
  } while (!done);
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

// ---- packed fp32x2 arithmetic (Blackwell FADD2: two IEEE round-to-nearest adds per issue slot;
//      operand B may be a scalar register broadcast to both halves, |x| is an operand modifier)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 abs2(f32x2 v) {
  float lo, hi;
  upk2(v, lo, hi);
  return pk2(fabsf(lo), fabsf(hi));
}

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

// The tile's D x 32 SAD-of-Sobel costs: ONE 3-D tensor copy straight into parking plane 1.
// They come from DRAM and are only needed after phase 1, hence their own barrier.
__device__ __forceinline__ void stage_sad_tma(const FusedArgs& a, const CUtensorMap* sad_map, const TileId& t,
                                              float* park_plane1, unsigned long long* bar_sad) {
  const FusedGeom& g = a.g;
  mbar_expect_tx(bar_sad, (unsigned)g.D * kTile * 4u);
  tma_load_3d(park_plane1, sad_map, t.x0 + g.bwl + g.sxo, t.y + g.bh, t.n * g.D, bar_sad);  // inner coordinate % 4 == 0
}


#include "ms_fused_tile.cuh"

// ---- one CTA per tile ------------------------------------------------------------------------
// kTma: right-image rows and the SAD-of-Sobel tile arrive through cp.async.bulk /
// cp.async.bulk.tensor (D <= 256: box limit); otherwise through LDGSTS.  The tensor copy's
// inner coordinate must be a multiple of 16 bytes (measured: anything else raises an
// illegal-instruction fault), which is why the scratch is stored with column offset sxo.
// kSlabA: stop after phase 1 and emit what slab.cu's phase A emits (tile_slab_a).
template <int DMAX, bool kTma, bool kSlabA>
__global__ void __launch_bounds__(256, 2)
ms_fused_kernel(const FusedArgs a, const __grid_constant__ CUtensorMap sad_map) {
  using L = Lay3<DMAX>;
  constexpr int NT = 256;
  constexpr int PS = L::PS;
  extern __shared__ __align__(128) unsigned char smem_raw[];  // TMA destinations need 128 B
  const FusedGeom& g = a.g;
  const int D = g.D;
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem_raw + L::off_bar);  // [0] rows, [1] sadsob tile
  float* s_red = reinterpret_cast<float*>(smem_raw + L::off_red);    // [16][4][32]
  float* s_min = reinterpret_cast<float*>(smem_raw + L::off_min);
  float* s_inv = reinterpret_cast<float*>(smem_raw + L::off_inv);
  float* s_par = reinterpret_cast<float*>(smem_raw + L::off_par);    // [3][DS][32]
  uint8_t* s_cen = smem_raw + L::off_par + L::pk_cen;                // [DS][32]
  const int tid = threadIdx.x;
  const TileId t = decode_tile(blockIdx.x, a);
  const StageGeo sg = stage_geo(a, t);
  if (kTma) {
    if (tid == 0) {
      mbar_init(&s_bar[0], 1);
      mbar_init(&s_bar[1], 1);
      mbar_init_fence();
      // rows first: phase 1 waits for them; the SAD-of-Sobel box is only needed after phase 1
      stage_rows<L, true, NT>(a, t, sg, smem_raw, &s_bar[0]);
      stage_sad_tma(a, &sad_map, t, s_par + PS, &s_bar[1]);
    }
  } else {
    // the tile's SAD-of-Sobel costs: async global -> parked plane 1
    const size_t splane = (size_t)g.H * g.Ws;
    const float* src = a.sadsob + ((size_t)t.n * D * g.H + (t.y + g.bh)) * g.Ws + (t.x0 + g.bwl + g.sxo);
    for (int i = tid; i < D * kTile; i += NT) cp_async4(s_par + PS + i, src + (size_t)(i >> 5) * splane + (i & 31));
    stage_rows<L, false, NT>(a, t, sg, smem_raw, nullptr);
  }

  P1Ctx c;
  c.pr = tid & 15;
  const int grp = tid >> 4;
  Left2 lr;
  load_left2(a, t, c.pr, lr);
  c.dA0 = grp * a.DC;
  c.nsteps = a.DC;
  c.cx0 = t.x0 + 2 * c.pr + g.bwl + g.padL - (g.d0 + c.dA0);
  // largest local disparity with a cost, per pixel and matcher (-1 - d0 or less: none):
  // cost(y,x,d) exists iff the window origin is inside the image and x - wc >= d (and d < D)
  int dmaxCN[4], dmaxZA, dmaxZB;   // CN: census A, census B, ncc A, ncc B
  {
    const int H = g.H, W = g.W;
    const int XA = t.x0 + 2 * c.pr + g.bwl, XB = XA + 1, Y = t.y + g.bh;
    const bool yc = (Y >= 5 && Y < H - 6), yn = (Y >= 1 && Y < H - 2), yz = (Y >= 2 && Y < H - 3);
    dmaxCN[0] = min(D - 1, ((yc && XA >= 5 && XA < W - 6) ? XA - 5 : -1) - g.d0);
    dmaxCN[1] = min(D - 1, ((yc && XB >= 5 && XB < W - 6) ? XB - 5 : -1) - g.d0);
    dmaxCN[2] = min(D - 1, ((yn && XA >= 1 && XA < W - 2) ? XA - 1 : -1) - g.d0);
    dmaxCN[3] = min(D - 1, ((yn && XB >= 1 && XB < W - 2) ? XB - 1 : -1) - g.d0);
    dmaxZA = min(D - 1, ((yz && XA >= 2 && XA < W - 3) ? XA - 2 : -1) - g.d0);
    dmaxZB = min(D - 1, ((yz && XB >= 2 && XB < W - 3) ? XB - 2 : -1) - g.d0);
  }
  if (!kTma) cp_async_wait_all();
  __syncthreads();                       // barrier init (and LDGSTS data) visible to everyone
  if (kTma) mbar_wait(&s_bar[0], 0);

  P1Min mn;
  mn.cenA = 255; mn.cenB = 255;
  mn.nccA = kFill; mn.nccB = kFill; mn.sadA = kFill; mn.sadB = kFill;
  if (grp == 0)
    p1_extra_b0<L>(a, t, smem_raw, sg, s_par, s_cen, lr, c.pr, c.cx0 + 1, dmaxCN[1], dmaxCN[3], dmaxZB, mn);
  p1_zsad<L>(a, smem_raw, sg, s_par, lr, c, dmaxZA, dmaxZB, mn);
  p1_census_ncc<L>(a, t, smem_raw, sg, s_par, s_cen, c, dmaxCN, mn);
  {
    float* r0 = s_red + grp * 4 * kTile + 2 * c.pr;
    r0[0] = (mn.cenA == 255) ? kFill : (float)mn.cenA;
    r0[1] = (mn.cenB == 255) ? kFill : (float)mn.cenB;
    r0[kTile] = mn.nccA;
    r0[kTile + 1] = mn.nccB;
    r0[3 * kTile] = mn.sadA;
    r0[3 * kTile + 1] = mn.sadB;
  }
  if (kTma) mbar_wait(&s_bar[1], 0);   // (LDGSTS: landed before the first barrier)
  {
    const int Xl = t.x0 + g.bwl, Yr = t.y + g.bh;   // every pixel of the tile has a SAD-of-Sobel cost at every disparity?
    const bool sob_all = (Xl - 2 >= g.d0 + D - 1) && (Xl + kTile - 1 < g.W - 3) && (Yr >= 2) && (Yr < g.H - 3);
    sob_finish<L>(a, t, s_par, s_red, tid, sob_all);
  }
  __syncthreads();
  if (kSlabA) tile_slab_a<L>(a, t, tid, s_par, s_cen, s_red);
  else if (a.xflags & 2) tile_back_half_otf<L>(a, t, tid, s_par, s_cen, s_red, s_min, s_inv);
  else tile_back_half<L>(a, t, tid, s_par, s_cen, reinterpret_cast<float*>(smem_raw + L::off_cene), s_red, s_min, s_inv);
}

// Optional per-kernel timing (msn_profile_enable): CUDA events recorded on the launch
// stream between the kernels of the fused sequence, read back by msn_profile_read.
struct ProfRec { cudaEvent_t ev[4]; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
constexpr size_t kProfMaxPending = 4096;   // records kept when nobody calls msn_profile_read

// cuTensorMapEncodeTiled results, keyed by what they describe (a launch sequence repeats the same map)
struct MapKey {
  const void* base;
  int Ws, H, ND, D;
  bool operator==(const MapKey& o) const { return base == o.base && Ws == o.Ws && H == o.H && ND == o.ND && D == o.D; }
};
std::mutex g_map_mu;
std::vector<std::pair<MapKey, CUtensorMap>> g_maps;

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
static bool tma_disabled() {
  const char* e = getenv("MSNETS_NO_TMA");
  return e && e[0] == '1';
}

// 3-D tensor map over the SAD-of-Sobel scratch [N*D][H][Ws], box 32 x 1 x D (cached)
static bool sad_tensor_map(const FusedGeom& g, int H, int N, const float* base, CUtensorMap* out) {
  const MapKey key{base, g.Ws, H, N * g.D, g.D};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    for (auto& kv : g_maps)
      if (kv.first == key) {
        *out = kv.second;
        return true;
      }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)g.Ws, (cuuint64_t)H, (cuuint64_t)N * g.D};
  const cuuint64_t gstr[2] = {(cuuint64_t)g.Ws * 4, (cuuint64_t)H * g.Ws * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kTile, 1u, (cuuint32_t)g.D};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  if (enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_maps.size() >= 64) g_maps.erase(g_maps.begin());
  g_maps.emplace_back(key, *out);
  return true;
}

int profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  if (!g_prof_on) {   // switching off drops whatever was never read
    for (ProfRec& r : g_prof)
      for (int i = 0; i < 4; ++i) cudaEventDestroy(r.ev[i]);
    g_prof.clear();
  }
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
  return p->censw == kCensW && p->nccw == kNccW && p->sadw == kSadW && p->sobelw == kSadW &&
         Dn <= kMaxFusedD;  // (image width is checked at launch; p->lr: the caller adds the right view)
}

size_t fused_workspace_bytes(int N, int H, int W, int Dn, const msn_ms_params* p) {
  (void)Dn;
  FusedGeom g = make_geom(N, H, W, p);
  FusedWs ws;
  ws.carve(nullptr, g);
  return ws.total;
}

// One launch of an instantiation; the dynamic shared-memory opt-in is set once per instantiation
// and device.
template <int DMAX, bool kTma, bool kSlabA>
static int launch_inst(const FusedArgs& a, const CUtensorMap& map, long long tiles, cudaStream_t s) {
  auto kern = ms_fused_kernel<DMAX, kTma, kSlabA>;
  constexpr size_t smem = Lay3<DMAX>::bytes;
  static std::mutex mu;
  static unsigned long long done_mask = 0;   // bit per device ordinal (< 64)
  int dev = 0;
  MSN_CUDA_OK(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 64 || !((done_mask >> dev) & 1ull)) {
      MSN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (dev < 64) done_mask |= 1ull << dev;
    }
  }
  kern<<<(unsigned)tiles, 256, smem, s>>>(a, map);
  return 0;
}

// d_mins == nullptr: the whole feature volume (p describes all disparities).  Otherwise phase A of
// the slab path for disparities [p->d_begin, p->d_begin + p->d_count): the output tensor holds
// out_D disparities per channel and the slab starts at out_d0 in it (a rank's own slab tensor:
// out_D = d_count, out_d0 = 0; slabs of one big volume on one GPU: out_D = ndisp, out_d0 = d_begin,
// accumulate != 0 from the second slab on).
int launch_ms_fused(const uint8_t* d_left, const uint8_t* d_right, int N, int H, int W, const msn_ms_params* p,
                    float* d_out, float* d_mins, char* workspace, cudaStream_t s, int out_D, int out_d0,
                    int accumulate) {
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
  ms_prep_kernel<<<pgrid, 128, 0, s>>>(d_left, d_right, g, ws.desc[0], ws.desc[1], ws.statL, ws.aR, ws.cR, ws.meanR,
                                       ws.fimg[0], ws.fimg[1], ws.sob[0], ws.sob[1]);
  MSN_LAUNCH_OK();
  if (prof) MSN_CUDA_OK(cudaEventRecord(rec.ev[1], s));
  if (launch_sadsob5_padded(ws.sob[0], ws.sob[1], N, H, W, g.D, g.d0, ws.sadsob + g.sxo, ws.sad_ws, s)) return 1;
  if (prof) MSN_CUDA_OK(cudaEventRecord(rec.ev[2], s));

  FusedArgs a;
  a.g = g;
  a.descL = ws.desc[0]; a.descR = ws.desc[1];
  a.statL = ws.statL;
  a.aR = ws.aR; a.cR = ws.cR; a.meanR = ws.meanR;
  a.fL = ws.fimg[0]; a.fR = ws.fimg[1];
  a.sadsob = ws.sadsob;
  a.out = d_out;
  a.mins = d_mins;
  a.out_channels = p->lr ? 16 : 8;
  a.mins_planes = p->lr ? 8 : 4;
  a.out_D = out_D > 0 ? out_D : g.D;
  a.out_d0 = out_d0;
  a.mins_accumulate = accumulate;
  a.k_cen = aml_scale(p->cens_sigma);
  a.k_ncc = aml_scale(p->ncc_sigma);
  a.k_sad = aml_scale(p->sad_sigma);
  a.neg_zero = -0.0f;
  { const char* e = getenv("MSNETS_X"); a.xflags = e ? atoi(e) : 0; }
  a.tiles_x = (g.w + kTile - 1) / kTile;
  a.DC = (g.D + kG2 - 1) / kG2;
  const long long tiles = (long long)N * g.h * a.tiles_x;
  MSN_REQUIRE(tiles <= 2147483647LL, "ms_features: too many tiles for one launch");
  CUtensorMap sad_map;
  memset(&sad_map, 0, sizeof(sad_map));
  bool use_tma = g.D <= 256 && !tma_disabled();
  if (use_tma) use_tma = sad_tensor_map(g, H, N, ws.sadsob, &sad_map);
#define MSN_FUSED_LAUNCH(DMAX, TMA)                                                \
  {                                                                                \
    if (d_mins) { if (launch_inst<DMAX, TMA, true>(a, sad_map, tiles, s)) return 1; } \
    else { if (launch_inst<DMAX, TMA, false>(a, sad_map, tiles, s)) return 1; }    \
  }
#define MSN_FUSED_CASE(DMAX)                                                       \
  if (g.D <= DMAX) {                                                               \
    if (use_tma && DMAX <= 256) MSN_FUSED_LAUNCH(DMAX <= 256 ? DMAX : 256, true)   \
    else MSN_FUSED_LAUNCH(DMAX, false)                                             \
  } else
  MSN_FUSED_CASE(64)
  MSN_FUSED_CASE(128)
  MSN_FUSED_CASE(192)
  MSN_FUSED_CASE(256)
  MSN_FUSED_CASE(384) { return fail("ms_features: D=%d exceeds the fused kernel's limit", g.D); }
#undef MSN_FUSED_CASE
#undef MSN_FUSED_LAUNCH
  MSN_LAUNCH_OK();
  if (prof) {
    MSN_CUDA_OK(cudaEventRecord(rec.ev[3], s));
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof.size() >= kProfMaxPending) {   // nobody reads: recycle the oldest record instead of growing
      for (int i = 0; i < 4; ++i) cudaEventDestroy(g_prof.front().ev[i]);
      g_prof.erase(g_prof.begin());
    }
    g_prof.push_back(rec);
  }
  return 0;
}

}  // namespace msn
