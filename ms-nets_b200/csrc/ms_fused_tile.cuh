// ms_fused_tile.cuh -- the tile body of ms_fused_kernel (included inside ms_fused.cu's anonymous
// namespace, after the TMA / packed-fp32 helpers and FusedArgs).
//
// Tile = (pair n, output row y, 32 consecutive x) x all D; one CTA of 256 threads per tile.
//
// What bounds this kernel (ncu, profiles/r2*): two resources at once -- the L1 / shared-memory data
// pipe (one 128-byte wavefront per cycle per SM; the round-1 kernel ran it at 78 % of peak) and
// instruction issue with only 16 resident warps per SM.  Hence: no lookup tables and no bank
// conflicts in the phase-1 loads (de-interleaved staging) for the former, packed fp32 arithmetic,
// exponentials evaluated once and no work on voxels without a cost for the latter.
//
//   phase 1   thread = (pixel PAIR (2k, 2k+1), d-group g of 16).  Per step the thread evaluates
//             two voxels on a DIAGONAL of the volume: A = (x, d) and B = (x+1, d+1).  Both read
//             the SAME right-image column x-d, hence the same right census code, the same right
//             NCC/ZSAD statistics and the same 5x5 right window -- so every ZSAD tap is one packed
//             FADD2 with the right pixel broadcast to both halves (75 FADD2 per voxel pair, each
//             half an IEEE fp32 add in the reference's order, matchers.cpp:499-506), every NCC
//             product one packed FFMA2, and the shared-memory loads are paid once per pair.
//             Neighbouring lanes own neighbouring PAIRS, i.e. right columns two apart: the staged
//             right-image rows are stored DE-INTERLEAVED (even columns, then odd columns -- the
//             prep kernel writes them that way and the TMA copies fetch the two halves), so at any
//             step the 16 lanes of a d-group read 16 consecutive words: conflict-free.
//             Two loops so that neither outgrows the 128-register budget of 2 CTAs/SM:
//               loop Z   ZSAD over a register-resident right window that slides one column per
//                        step (6 physical columns: the next column loads while this one computes)
//               loop CN  census (xor + popc) and NCC (9 FFMA2 of exact integers, fp64 scaling),
//                        six steps side by side for instruction-level parallelism
//             Raw costs are parked in shared memory [d][32] (ncc, zsad floats; census byte); the
//             tile's SAD-of-Sobel costs arrive by TMA straight into their parking plane.
//             A thread walks its disparities in blocks of 6 steps; a block in which every voxel has
//             its costs runs a body without validity selects, a block past the valid range only
//             parks fill (no arithmetic at all), the rest a generic body -- chosen per warp.
//   back half every AML exponential is evaluated ONCE:
//     pass E  thread = (pixel quad, d): channels 0-3 normalised and stored (128-bit rows), the
//             exponentials exp(-(c-m)^2/sigma) written over the parked costs (census: own plane
//             in the dead staging area);
//     pass S  one thread per (pixel, matcher): the AML denominator, added sequentially in d
//             order as the reference does (featextract.cpp:444-447) -- loads and adds only;
//     pass N  thread = (pixel quad, d): channels 4-7 = e * (1/den), 128-bit rows.
#pragma once

constexpr int kG2 = 16;       // d-groups per tile (16 pixel pairs x 16 groups = 256 threads)

// Staged right-image data of a tile: columns [cl, ch] (padded coordinates) of the tile's row,
//   cl = Xt - d0 - 16*DC - 2,  ch = Xt - d0 + 33,  Xt = x0 + board_w_left + padL,
// every array as two halves (even columns / odd columns).  Column c lives in half (c & 1) at
// element (c >> 1) - pstart, pstart = (cl >> 1) rounded down to the array's 16-byte granule.
// The staging area and the per-group minima are dead once the back half starts: the same bytes
// then hold the census AML exponentials as a fourth float plane [D][32].
template <int DMAX>
struct Lay3 {
  static constexpr int HE = DMAX / 2 + 28;                       // positions per half before alignment slack
  static constexpr int capD = HE + 2;                            // uint4  (granule 1)
  static constexpr int capC = (HE + 4 + 1) & ~1;                 // double (granule 2)
  static constexpr int capF = (HE + 8 + 3) & ~3;                 // float  (granule 4)
  static constexpr size_t st_desc = 0;                                     // [2][capD] uint4 census codes
  static constexpr size_t st_c = st_desc + (size_t)2 * capD * 16;          // [2][capC] double NCC C
  static constexpr size_t st_a = st_c + (size_t)2 * capC * 8;              // [2][capF] float NCC A
  static constexpr size_t st_mean = st_a + (size_t)2 * capF * 4;           // [2][capF] float ZSAD mean
  static constexpr size_t st_rf = st_mean + (size_t)2 * capF * 4;          // [5][2][capF] float pixel rows
  static constexpr size_t st_bytes = (st_rf + (size_t)10 * capF * 4 + 127) & ~(size_t)127;
  static constexpr int DS = DMAX + 1;                            // parked rows; row D is scratch for dummy steps
  static constexpr int PS = DS * kTile;                          // floats per parked matcher
  static constexpr size_t off_red = st_bytes;                    // [kG2][4][32] per-group minima
  static constexpr size_t red_end = off_red + (size_t)kG2 * 4 * kTile * 4;
  static constexpr size_t off_cene = 0;                          // back half: [DMAX][32] census exponentials (aliases the above)
  static constexpr size_t cene_end = (size_t)DMAX * kTile * 4;
  static constexpr size_t off_min = ((red_end > cene_end ? red_end : cene_end) + 127) & ~(size_t)127;   // [4][32]
  static constexpr size_t off_inv = off_min + 4 * kTile * 4;     // [4][32]
  static constexpr size_t off_par = (off_inv + 4 * kTile * 4 + 127) & ~(size_t)127;   // [3][DS][32] floats; plane 1 is a TMA destination
  static constexpr size_t pk_cen = (size_t)3 * PS * 4;           // then [DS][32] census bytes
  static constexpr size_t off_bar = off_par + ((pk_cen + (size_t)DS * kTile + 127) & ~(size_t)127);
  static constexpr size_t bytes = off_bar + 32;
};

__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// ---- staging ---------------------------------------------------------------------------------
struct StageGeo {
  int pD, pC, pF;   // first staged position of the uint4 / double / float arrays
  int nD, nC, nF;   // elements copied per half
};
__device__ __forceinline__ StageGeo stage_geo(const FusedArgs& a, const TileId& t) {
  const FusedGeom& g = a.g;
  const int Xt = t.x0 + g.bwl + g.padL;
  const int cl = Xt - g.d0 - kG2 * a.DC - 2, ch = Xt - g.d0 + 33;
  StageGeo s;
  const int pl = cl >> 1, ph = ch >> 1;
  s.pD = pl;
  s.pC = pl & ~1;
  s.pF = pl & ~3;
  s.nD = ph - s.pD + 1;
  s.nC = (ph - s.pC + 2) & ~1;
  s.nF = (ph - s.pF + 4) & ~3;
  return s;
}

// kTma: one thread issues the 18 bulk copies (completion counted in bytes on `bar`); otherwise all
// threads copy with LDGSTS.
template <class L, bool kTma, int NT>
__device__ __forceinline__ void stage_rows(const FusedArgs& a, const TileId& t, const StageGeo& s, unsigned char* buf,
                                           unsigned long long* bar) {
  const FusedGeom& g = a.g;
  const int Yp = t.y + g.bh + kPadT;
  const size_t img_off = (size_t)t.n * g.img_px();
  const size_t row = img_off + (size_t)Yp * g.Wp;
  const int H2 = g.Wp >> 1;
  auto copy = [&](void* dst, const void* src, unsigned bytes) {
    if (kTma) {
      bulk_load(dst, src, bytes, bar);
    } else {
      for (unsigned i = threadIdx.x * 16u; i < bytes; i += NT * 16u)
        cp_async16(static_cast<unsigned char*>(dst) + i, static_cast<const unsigned char*>(src) + i);
    }
  };
  if (kTma) mbar_expect_tx(bar, 2u * (unsigned)s.nD * 16u + 2u * (unsigned)s.nC * 8u + 14u * (unsigned)s.nF * 4u);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    copy(buf + L::st_desc + (size_t)p * L::capD * 16, a.descR + row + (size_t)p * H2 + s.pD, (unsigned)s.nD * 16u);
    copy(buf + L::st_c + (size_t)p * L::capC * 8, a.cR + row + (size_t)p * H2 + s.pC, (unsigned)s.nC * 8u);
    copy(buf + L::st_a + (size_t)p * L::capF * 4, a.aR + row + (size_t)p * H2 + s.pF, (unsigned)s.nF * 4u);
    copy(buf + L::st_mean + (size_t)p * L::capF * 4, a.meanR + row + (size_t)p * H2 + s.pF, (unsigned)s.nF * 4u);
#pragma unroll
    for (int r = 0; r < 5; ++r)
      copy(buf + L::st_rf + (size_t)(2 * r + p) * L::capF * 4,
           a.fR + img_off + (size_t)(Yp - 2 + r) * g.Wp + (size_t)p * H2 + s.pF, (unsigned)s.nF * 4u);
  }
}

// address of (padded) column c in a two-half staged array of T; `cap` elements per half
template <class T>
__device__ __forceinline__ const T* colptr(const unsigned char* base, int cap, int pstart, int c) {
  return reinterpret_cast<const T*>(base) + (c & 1) * cap + ((c >> 1) - pstart);
}

// ---- left-image data of a pixel pair: straight from global memory (before the staging wait) ---
struct Left2 {
  float meanA, meanB;   // ZSAD window means
  float px[5][6];       // rows y-2..y+2, columns xA-2 .. xA+3
};
__device__ __forceinline__ void load_left2(const FusedArgs& a, const TileId& t, int pr, Left2& lr) {
  const FusedGeom& g = a.g;
  const int Yp = t.y + g.bh + kPadT;
  const int Xp = t.x0 + 2 * pr + g.bwl + g.padL;
  const size_t img_off = (size_t)t.n * g.img_px();
  const RStat* sp = a.statL + img_off + (size_t)Yp * g.Wp + Xp;
  lr.meanA = __ldg(&sp[0].mean);
  lr.meanB = __ldg(&sp[1].mean);
  const float* gf = a.fL + img_off + (size_t)(Yp - 2) * g.Wp + (Xp - 2);
#pragma unroll
  for (int r = 0; r < 5; ++r) {
#pragma unroll
    for (int c = 0; c < 6; ++c) lr.px[r][c] = __ldg(gf + c);
    gf += g.Wp;
  }
}

// per-thread description of its share of a tile: pixel pair pr = tile pixels 2*pr (voxel A) and
// 2*pr+1 (voxel B); nsteps steps from local disparity dA0 (voxel B: dA0 + 1); cx0 = padded
// right-image column x_A - d_A at step 0 (falls by one per step).
struct P1Ctx {
  int pr, dA0, nsteps, cx0;
};

struct P1Min {
  int cenA, cenB;
  float nccA, nccB, sadA, sadB;
};

__device__ __forceinline__ int popc128(const uint4& a, const uint4& b) {
  return __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
}

// NCC of one voxel from its exact numerator: fl32( (-(num) * C_L) * C_R ) in fp64, +1 when either
// C is not finite (matchers.cpp:196-204)
__device__ __forceinline__ float ncc_scale(float num, double cl, double cr) {
  const float v = (float)__dmul_rn(__dmul_rn(-(double)num, cl), cr);
  return (fabsf(v) <= 3.0e38f) ? v : 1.0f;
}

// In a block of 6 steps starting at column cb, step sI reads column cb - sI: even steps through the
// pointer of cb, odd steps through the pointer of cb - 1 (the other half), both with immediates.
#define MSN_STEP_PTR(p0, p1, sI) (((sI) & 1) ? (p1) - ((sI) >> 1) : (p0) - ((sI) >> 1))

// A thread walks its disparities in BLOCKS of 6 steps (the period of the register-resident sliding
// window).  Costs exist for d <= dmax (a per-pixel bound: the window must fit left of x - d), so a
// block is one of: all 12 voxels have costs -> the clean body (no selects, one basic block); none has
// (monotone in d: nothing after it has either) -> only fill is parked; mixed / partial -> the generic
// body with validity selects.  The choice is made per WARP (ballot), so there is no divergence.
enum { kBlkClean = 0, kBlkGeneric = 1, kBlkFill = 2 };
__device__ __forceinline__ int block_kind(int dA, int steps_left, int dmaxA_tight, int dmaxB_tight, int dmaxA_loose,
                                          int dmaxB_loose) {
  const bool clean = (steps_left >= 6) && (dA + 5 <= dmaxA_tight) && (dA + 6 <= dmaxB_tight);
  const bool none = (dA > dmaxA_loose) && (dA + 1 > dmaxB_loose);
  if (__all_sync(0xffffffffu, clean)) return kBlkClean;
  if (__all_sync(0xffffffffu, none)) return kBlkFill;
  return kBlkGeneric;
}

// ---- loop Z: ZSAD ---------------------------------------------------------------------------
// 25 taps row-major, ((L - mL) - R) + mR, sequential fp32 per voxel (matchers.cpp:499-506); the
// two voxels of the pair occupy the two halves of every packed operation.
template <class L>
struct ZState {
  const float *q0p, *q1p, *m0p, *m1p;
  float wv[5][6];   // sliding right window, 5 logical columns + the one being loaded for the next step
  float* pz;        // zsad plane, column of voxel A
  int dA;
};

template <class L, bool kClean>
__device__ __forceinline__ void z_block(ZState<L>& s, int steps_left, int D, const f32x2 (&ap)[5][5], int dmaxA,
                                        int dmaxB, P1Min& mn) {
#pragma unroll
  for (int sI = 0; sI < 6; ++sI) {
    if (!kClean && sI > 0 && sI >= steps_left) break;
#define WV(r, j) s.wv[r][((j) + 12 - sI) % 6]
    const int dA = s.dA + sI, dB = dA + 1;
    // next step's new left column into the spare slot (logical column -1 of this step)
    {
      const float* q = MSN_STEP_PTR(s.q0p, s.q1p, sI);
#pragma unroll
      for (int r = 0; r < 5; ++r) WV(r, 5) = q[r * 2 * L::capF];
    }
    const float mR = *MSN_STEP_PTR(s.m0p, s.m1p, sI);
    const f32x2 mR2 = pk2(mR, mR);
    f32x2 acc = pk2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float w = WV(r, j);
        acc = add2(acc, abs2(add2(sub2(ap[r][j], pk2(w, w)), mR2)));
      }
    float zA, zB;
    upk2(acc, zA, zB);
    int rowA = dA, rowB = dB;
    if (!kClean) {
      zA = (dA <= dmaxA) ? zA : kFill;
      zB = (dB <= dmaxB) ? zB : kFill;
      rowA = min(dA, D);     // dummy steps (d >= D) park into the scratch row
      rowB = min(dB, D);
    }
    s.pz[rowA * kTile] = zA;
    s.pz[rowB * kTile + 1] = zB;
    mn.sadA = fminf(mn.sadA, zA);
    mn.sadB = fminf(mn.sadB, zB);
#undef WV
  }
}

template <class L>
__device__ __forceinline__ void p1_zsad(const FusedArgs& a, const unsigned char* stage, const StageGeo& sg,
                                        float* s_par, const Left2& lr, const P1Ctx& c, int dmaxA, int dmaxB,
                                        P1Min& mn) {
  const int D = a.g.D;
  f32x2 ap[5][5];   // (L_A[tap] - mL_A, L_B[tap] - mL_B), hoisted over all d
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int j = 0; j < 5; ++j)
      ap[r][j] = pk2(__fsub_rn(lr.px[r][j], lr.meanA), __fsub_rn(lr.px[r][j + 1], lr.meanB));
  ZState<L> s;
  const int cx = c.cx0;
  const unsigned char* rfb = stage + L::st_rf;
  // the column loaded at step sI for step sI+1: cx - sI - 3
  s.q0p = colptr<float>(rfb, L::capF, sg.pF, cx - 3);
  s.q1p = colptr<float>(rfb, L::capF, sg.pF, cx - 4);
  s.m0p = colptr<float>(stage + L::st_mean, L::capF, sg.pF, cx);
  s.m1p = colptr<float>(stage + L::st_mean, L::capF, sg.pF, cx - 1);
  // initial window, columns cx-2 .. cx+2: (cx-4)+2, (cx-3)+2, (cx-4)+4, (cx-3)+4, (cx-4)+6
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const float* q = (j & 1) ? s.q0p + 1 + (j >> 1) : s.q1p + 1 + (j >> 1);
#pragma unroll
    for (int r = 0; r < 5; ++r) s.wv[r][j] = q[r * 2 * L::capF];
  }
  s.pz = s_par + 2 * L::PS + 2 * c.pr;
  s.dA = c.dA0;

  for (int base = 0; base < c.nsteps; base += 6) {
    const int kind = block_kind(s.dA, c.nsteps - base, dmaxA, dmaxB, dmaxA, dmaxB);
    if (kind == kBlkClean) {
      z_block<L, true>(s, 6, D, ap, dmaxA, dmaxB, mn);
    } else if (kind == kBlkGeneric) {
      z_block<L, false>(s, c.nsteps - base, D, ap, dmaxA, dmaxB, mn);
    } else {
#pragma unroll
      for (int sI = 0; sI < 6; ++sI) {
        if (sI >= c.nsteps - base) break;
        s.pz[min(s.dA + sI, D) * kTile] = kFill;
        s.pz[min(s.dA + sI + 1, D) * kTile + 1] = kFill;
      }
    }
    s.dA += 6;
    s.q0p -= 3; s.q1p -= 3; s.m0p -= 3; s.m1p -= 3;
  }
}

// ---- loop CN: census + NCC ------------------------------------------------------------------
// Per voxel NCC is one long dependent chain (9 products -> fp64 scaling -> fp32) and census a burst
// of XU work, so a block evaluates its six steps side by side: all right-image data of the block is
// loaded up front (8 window columns x 3 rows, 6 codes, 6 A, 6 C; no state carried between blocks)
// and the six steps are independent instruction streams for the scheduler to interleave.
// dmax: census A, census B, ncc A, ncc B
template <class L, bool kClean>
__device__ __forceinline__ void cn_block(const unsigned char* stage, const StageGeo& sg, int cx, int dA0,
                                         int steps_left, int D, const uint4& ldA, const uint4& ldB,
                                         const f32x2 (&l3)[3][3], f32x2 lA2, double lCA, double lCB,
                                         const int (&dmax)[4], float* pn, uint8_t* pc, P1Min& mn) {
  // columns cx-6 .. cx+1 of pixel rows y-1 .. y+1 (rows 1..3 of the staged five)
  const unsigned char* rfb = stage + L::st_rf + (size_t)2 * L::capF * 4;
  const float* e0 = colptr<float>(rfb, L::capF, sg.pF, cx - 6);   // cx-6, cx-4, cx-2, cx
  const float* e1 = colptr<float>(rfb, L::capF, sg.pF, cx - 5);   // cx-5, cx-3, cx-1, cx+1
  float w[3][8];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      w[r][2 * k] = e0[r * 2 * L::capF + k];
      w[r][2 * k + 1] = e1[r * 2 * L::capF + k];
    }
  const float* a0p = colptr<float>(stage + L::st_a, L::capF, sg.pF, cx);
  const float* a1p = colptr<float>(stage + L::st_a, L::capF, sg.pF, cx - 1);
  const double* c0p = colptr<double>(stage + L::st_c, L::capC, sg.pC, cx);
  const double* c1p = colptr<double>(stage + L::st_c, L::capC, sg.pC, cx - 1);
  const uint4* d0p = colptr<uint4>(stage + L::st_desc, L::capD, sg.pD, cx);
  const uint4* d1p = colptr<uint4>(stage + L::st_desc, L::capD, sg.pD, cx - 1);
  const f32x2 nine2 = pk2(9.0f, 9.0f);
#pragma unroll
  for (int sI = 0; sI < 6; ++sI) {
    if (!kClean && sI > 0 && sI >= steps_left) break;
    const int dA = dA0 + sI, dB = dA + 1;
    const uint4 rd = *MSN_STEP_PTR(d0p, d1p, sI);
    const double rC = *MSN_STEP_PTR(c0p, c1p, sI);
    const float rA = *MSN_STEP_PTR(a0p, a1p, sI);
    // census: Hamming distance of the packed codes (matchers.cpp:323-337)
    int cenA = popc128(ldA, rd);
    int cenB = popc128(ldB, rd);
    // NCC: P exact in fp32 (< 2^24); scaling in fp64 left to right (matchers.cpp:200-201)
    f32x2 P = pk2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float wj = w[r][5 - sI + j];   // column cx - sI - 1 + j
        P = fma2(l3[r][j], pk2(wj, wj), P);
      }
    const float nra = -rA;
    const f32x2 num2 = fma2(nine2, P, mul2(lA2, pk2(nra, nra)));   // 9P - A_L*A_R, exact
    float numA, numB;
    upk2(num2, numA, numB);
    float nccA = ncc_scale(numA, lCA, rC);
    float nccB = ncc_scale(numB, lCB, rC);
    int rowA = dA, rowB = dB;
    if (!kClean) {
      cenA = (dA <= dmax[0]) ? cenA : 255;
      cenB = (dB <= dmax[1]) ? cenB : 255;
      nccA = (dA <= dmax[2]) ? nccA : kFill;
      nccB = (dB <= dmax[3]) ? nccB : kFill;
      rowA = min(dA, D);
      rowB = min(dB, D);
    }
    pc[rowA * kTile] = (uint8_t)cenA;
    pc[rowB * kTile + 1] = (uint8_t)cenB;
    pn[rowA * kTile] = nccA;
    pn[rowB * kTile + 1] = nccB;
    mn.cenA = min(mn.cenA, cenA);
    mn.cenB = min(mn.cenB, cenB);
    mn.nccA = fminf(mn.nccA, nccA);
    mn.nccB = fminf(mn.nccB, nccB);
  }
}

template <class L>
__device__ __forceinline__ void p1_census_ncc(const FusedArgs& a, const TileId& t, const unsigned char* stage,
                                              const StageGeo& sg, float* s_par, uint8_t* s_cen, const P1Ctx& c,
                                              const int (&dmax)[4], P1Min& mn) {
  const FusedGeom& g = a.g;
  const int D = g.D;
  // the pair's left census codes, 3x4 patch and NCC statistics (read here, after loop Z, rather than
  // kept live through it: they come from L1/L2)
  uint4 ldA, ldB;
  f32x2 l3[3][3];
  f32x2 lA2;
  double lCA, lCB;
  {
    const int Yp = t.y + g.bh + kPadT;
    const int Xp = t.x0 + 2 * c.pr + g.bwl + g.padL;
    const size_t img_off = (size_t)t.n * g.img_px();
    const uint4* dp = a.descL + img_off + (size_t)Yp * g.Wp + Xp;
    ldA = __ldg(dp);
    ldB = __ldg(dp + 1);
    const float* gf = a.fL + img_off + (size_t)(Yp - 1) * g.Wp + (Xp - 1);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __ldg(gf + k);
      // + (-0.0) is exact; the run-time operand keeps each pair in registers of its own (ptxas would
      // otherwise rebuild the overlapping pairs with MOVs at every use)
#pragma unroll
      for (int j = 0; j < 3; ++j) l3[r][j] = add2(pk2(v[j], v[j + 1]), pk2(a.neg_zero, a.neg_zero));
      gf += g.Wp;
    }
    const RStat* sp = a.statL + img_off + (size_t)Yp * g.Wp + Xp;
    lA2 = pk2(__ldg(&sp[0].A), __ldg(&sp[1].A));
    lCA = __ldg(&sp[0].C);
    lCB = __ldg(&sp[1].C);
  }
  float* pn = s_par + 2 * c.pr;     // ncc plane, column of voxel A
  uint8_t* pc = s_cen + 2 * c.pr;   // census bytes, column of voxel A
  int dA = c.dA0, cx = c.cx0;
  for (int base = 0; base < c.nsteps; base += 6, dA += 6, cx -= 6) {
    const int kind = block_kind(dA, c.nsteps - base, dmax[0], dmax[1], dmax[2], dmax[3]);
    if (kind == kBlkClean) {
      cn_block<L, true>(stage, sg, cx, dA, 6, D, ldA, ldB, l3, lA2, lCA, lCB, dmax, pn, pc, mn);
    } else if (kind == kBlkGeneric) {
      cn_block<L, false>(stage, sg, cx, dA, c.nsteps - base, D, ldA, ldB, l3, lA2, lCA, lCB, dmax, pn, pc, mn);
    } else {
#pragma unroll
      for (int sI = 0; sI < 6; ++sI) {
        if (sI >= c.nsteps - base) break;
        const int rowA = min(dA + sI, D), rowB = min(dA + sI + 1, D);
        pc[rowA * kTile] = 255;
        pc[rowB * kTile + 1] = 255;
        pn[rowA * kTile] = kFill;
        pn[rowB * kTile + 1] = kFill;
      }
    }
  }
}

// Group 0's extra voxel: B = (odd pixel, d = 0) alone (its diagonal partner would be d = -1).
// cxB: padded right column of B at d = 0.  dmaxB*: validity bounds of pixel B (cost iff 0 <= dmax).
template <class L>
__device__ __forceinline__ void p1_extra_b0(const FusedArgs& a, const TileId& t, const unsigned char* stage,
                                            const StageGeo& sg, float* s_par, uint8_t* s_cen, const Left2& lr, int pr,
                                            int cxB, int dmaxB_cen, int dmaxB_ncc, int dmaxB_sad, P1Min& mn) {
  const FusedGeom& g = a.g;
  const uint4 lsB_raw = __ldg(reinterpret_cast<const uint4*>(
      a.statL + (size_t)t.n * g.img_px() + (size_t)(t.y + g.bh + kPadT) * g.Wp + (t.x0 + 2 * pr + 1 + g.bwl + g.padL)));
  const RStat lsB = *reinterpret_cast<const RStat*>(&lsB_raw);
  const uint4 ldB = __ldg(a.descL + (size_t)t.n * g.img_px() + (size_t)(t.y + g.bh + kPadT) * g.Wp +
                          (t.x0 + 2 * pr + 1 + g.bwl + g.padL));
  const uint4 rd = *colptr<uint4>(stage + L::st_desc, L::capD, sg.pD, cxB);
  const double rC = *colptr<double>(stage + L::st_c, L::capC, sg.pC, cxB);
  const float rA = *colptr<float>(stage + L::st_a, L::capF, sg.pF, cxB);
  const float mR = *colptr<float>(stage + L::st_mean, L::capF, sg.pF, cxB);
  int cen = popc128(ldB, rd);
  float P = 0.f, z = 0.f;
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float w = colptr<float>(stage + L::st_rf, L::capF, sg.pF, cxB - 2 + j)[r * 2 * L::capF];
      if (r >= 1 && r <= 3 && j >= 1 && j <= 3) P = __fmaf_rn(lr.px[r][j + 1], w, P);
      z = __fadd_rn(z, fabsf(__fadd_rn(__fsub_rn(__fsub_rn(lr.px[r][j + 1], lr.meanB), w), mR)));
    }
  const float num = __fmaf_rn(9.0f, P, -__fmul_rn(lsB.A, rA));
  float ncc = ncc_scale(num, lsB.C, rC);
  cen = (dmaxB_cen >= 0) ? cen : 255;
  ncc = (dmaxB_ncc >= 0) ? ncc : kFill;
  z = (dmaxB_sad >= 0) ? z : kFill;
  s_cen[2 * pr + 1] = (uint8_t)cen;
  s_par[2 * pr + 1] = ncc;
  s_par[2 * L::PS + 2 * pr + 1] = z;
  mn.cenB = min(mn.cenB, cen);
  mn.nccB = fminf(mn.nccB, ncc);
  mn.sadB = fminf(mn.sadB, z);
}

// ---- the SAD-of-Sobel plane (arrived by TMA / LDGSTS): validity and per-pixel minima ---------
// thread = (pixel, part of 8): disparities outside the valid region become fill.
template <class L>
__device__ __forceinline__ void sob_finish(const FusedArgs& a, const TileId& t, float* s_par, float* s_red, int tid,
                                           bool all_valid) {
  const FusedGeom& g = a.g;
  const int D = g.D;
  const int px = tid & 31, part = tid >> 5;
  const int per = (D + 7) >> 3;
  const int d_lo = part * per, d_end = min(D, d_lo + per);
  float* sp = s_par + L::PS + d_lo * kTile + px;
  float m = kFill;
  if (all_valid) {
#pragma unroll 4
    for (int d = d_lo; d < d_end; ++d, sp += kTile) m = fminf(m, *sp);
  } else {
    const int X = t.x0 + px + g.bwl, Y = t.y + g.bh;
    const int dmax = min(D - 1, ((Y >= 2 && Y < g.H - 3 && X >= 2 && X < g.W - 3) ? X - 2 : -1) - g.d0);
#pragma unroll 4
    for (int d = d_lo; d < d_end; ++d, sp += kTile) {
      float v = *sp;
      if (d > dmax) {
        v = kFill;
        *sp = v;
      }
      m = fminf(m, v);
    }
  }
  // planes of s_red: [group][matcher][32]; the SAD-of-Sobel minima use groups 0-7, the rest is fill
  s_red[(part * 4 + 2) * kTile + px] = m;
  s_red[((part + 8) * 4 + 2) * kTile + px] = kFill;
}

// ---- back half ------------------------------------------------------------------------------
// Channel 0 = clip(census, 0, 120) / 120 as a TRUE division (cbmv_generator.py:283): q = k*r refined
// by two FMAs is the correctly rounded quotient for every k in 0..120 (checked exhaustively in
// tests/test_host_math.py); a parked 255 (no cost) clips to 120 -> 1.0.  No table: shared-memory
// wavefronts are the scarcer resource here.
__device__ __forceinline__ float census_ch0(unsigned k) {
  const float kf = (float)min(k, 120u);
  const float r = 1.0f / 120.0f;
  const float q = __fmul_rn(kf, r);
  const float rem = __fmaf_rn(-q, 120.0f, kf);
  return __fmaf_rn(rem, r, q);
}
// AML exponential of a parked census byte: exp(-(k - m)^2 / sigma); no cost (255) -> 0
__device__ __forceinline__ float census_e(unsigned k, int mc, float kc) {
  const int t = (int)k - mc;
  const float e = ex2_approx(-(float)(t * t) * kc);
  return (k == 255u) ? 0.f : e;
}

// Stores four channel planes' 4-pixel row segments: 128-bit streaming stores when the rows are
// 16 B aligned and the quad is fully inside the image (kVec), guarded scalars otherwise.
template <bool kVec>
__device__ __forceinline__ void store_quads(float* o, size_t chan, int nlive, const float4& c0, const float4& c1,
                                            const float4& c2, const float4& c3) {
  if (kVec) {
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

// Pass E for thread = (pixel quad q4, disparities dl, dl+32, ...): channels 0-3 normalised
// (cbmv_generator.py:283-287) and stored as 128-bit row segments; every AML exponential
// exp(-(c-m)^2/sigma) evaluated ONCE -- the three float planes are overwritten in place, the census
// exponentials go to their own plane s_ce (the dead staging area).
template <bool kVec>
__device__ __forceinline__ void pass_e(float* s_par, const uint8_t* s_cen, float* s_ce, const float* s_min, int PS,
                                       int q4, int dl, int D, float* orow, size_t plane, size_t chan, int nlive,
                                       float k0, float k1, float k2) {
  const float4 m0 = *reinterpret_cast<const float4*>(s_min + q4);
  const float4 m1 = *reinterpret_cast<const float4*>(s_min + kTile + q4);
  const float4 m2 = *reinterpret_cast<const float4*>(s_min + 2 * kTile + q4);
  const float4 m3 = *reinterpret_cast<const float4*>(s_min + 3 * kTile + q4);
  const int mcx = (m0.x == kFill) ? 0 : (int)m0.x, mcy = (m0.y == kFill) ? 0 : (int)m0.y;
  const int mcz = (m0.z == kFill) ? 0 : (int)m0.z, mcw = (m0.w == kFill) ? 0 : (int)m0.w;
#pragma unroll 1
  for (int d = dl; d < D; d += 32) {
    float* e0 = s_par + d * kTile + q4;
    const uchar4 cb = *reinterpret_cast<const uchar4*>(s_cen + d * kTile + q4);
    const float4 v1 = *reinterpret_cast<const float4*>(e0);
    const float4 v2 = *reinterpret_cast<const float4*>(e0 + PS);
    const float4 v3 = *reinterpret_cast<const float4*>(e0 + 2 * PS);
    const float4 c0 = make_float4(census_ch0(cb.x), census_ch0(cb.y), census_ch0(cb.z), census_ch0(cb.w));
    const float4 c1 = make_float4(normalise_cost(v1.x, 1), normalise_cost(v1.y, 1), normalise_cost(v1.z, 1),
                                  normalise_cost(v1.w, 1));
    const float4 c2 = make_float4(normalise_cost(v2.x, 2), normalise_cost(v2.y, 2), normalise_cost(v2.z, 2),
                                  normalise_cost(v2.w, 2));
    const float4 c3 = make_float4(normalise_cost(v3.x, 3), normalise_cost(v3.y, 3), normalise_cost(v3.z, 3),
                                  normalise_cost(v3.w, 3));
    store_quads<kVec>(orow + (size_t)d * plane, chan, nlive, c0, c1, c2, c3);
    *reinterpret_cast<float4*>(s_ce + d * kTile + q4) =
        make_float4(census_e(cb.x, mcx, k0), census_e(cb.y, mcy, k0), census_e(cb.z, mcz, k0), census_e(cb.w, mcw, k0));
    *reinterpret_cast<float4*>(e0) =
        make_float4(aml_e(v1.x, m1.x, k1), aml_e(v1.y, m1.y, k1), aml_e(v1.z, m1.z, k1), aml_e(v1.w, m1.w, k1));
    *reinterpret_cast<float4*>(e0 + PS) =
        make_float4(aml_e(v2.x, m2.x, k2), aml_e(v2.y, m2.y, k2), aml_e(v2.z, m2.z, k2), aml_e(v2.w, m2.w, k2));
    *reinterpret_cast<float4*>(e0 + 2 * PS) =
        make_float4(aml_e(v3.x, m3.x, k2), aml_e(v3.y, m3.y, k2), aml_e(v3.z, m3.z, k2), aml_e(v3.w, m3.w, k2));
  }
}

// Pass N: channels 4-7 = e * (1/den), 128-bit row segments.
template <bool kVec>
__device__ __forceinline__ void pass_n(const float* s_par, const float* s_ce, const float* s_inv, int PS, int q4,
                                       int dl, int D, float* arow, size_t plane, size_t chan, int nlive) {
  const float4 i0 = *reinterpret_cast<const float4*>(s_inv + q4);
  const float4 i1 = *reinterpret_cast<const float4*>(s_inv + kTile + q4);
  const float4 i2 = *reinterpret_cast<const float4*>(s_inv + 2 * kTile + q4);
  const float4 i3 = *reinterpret_cast<const float4*>(s_inv + 3 * kTile + q4);
#pragma unroll 2
  for (int d = dl; d < D; d += 32) {
    const float* e0 = s_par + d * kTile + q4;
    const float4 v0 = *reinterpret_cast<const float4*>(s_ce + d * kTile + q4);
    const float4 v1 = *reinterpret_cast<const float4*>(e0);
    const float4 v2 = *reinterpret_cast<const float4*>(e0 + PS);
    const float4 v3 = *reinterpret_cast<const float4*>(e0 + 2 * PS);
    const float4 a0 = make_float4(v0.x * i0.x, v0.y * i0.y, v0.z * i0.z, v0.w * i0.w);
    const float4 a1 = make_float4(v1.x * i1.x, v1.y * i1.y, v1.z * i1.z, v1.w * i1.w);
    const float4 a2 = make_float4(v2.x * i2.x, v2.y * i2.y, v2.z * i2.z, v2.w * i2.w);
    const float4 a3 = make_float4(v3.x * i3.x, v3.y * i3.y, v3.z * i3.z, v3.w * i3.w);
    store_quads<kVec>(arow + (size_t)d * plane, chan, nlive, a0, a1, a2, a3);
  }
}

// Pass S for one thread = (pixel, matcher): den = sum over d of e, added SEQUENTIALLY in d order as the
// reference does (featextract.cpp:444-447; a tree sum is measurably outside the 2e-6 bound).  Only
// loads and adds; the loads run a batch of 16 ahead of the dependent adds.
__device__ __forceinline__ float chain_sum(const float* e, int D) {
  constexpr int B = 16;
  float den = 0.f;
  float bufA[B], bufB[B];
  const int nb = D / (2 * B);   // double batches
  if (nb > 0) {
#pragma unroll
    for (int j = 0; j < B; ++j) bufA[j] = e[j * kTile];
  }
  for (int b = 0; b < nb; ++b) {
    const float* eb = e + (size_t)(2 * b + 1) * B * kTile;
#pragma unroll
    for (int j = 0; j < B; ++j) bufB[j] = eb[j * kTile];
#pragma unroll
    for (int j = 0; j < B; ++j) den = __fadd_rn(den, bufA[j]);
    if (b + 1 < nb) {
#pragma unroll
      for (int j = 0; j < B; ++j) bufA[j] = eb[(B + j) * kTile];
    }
#pragma unroll
    for (int j = 0; j < B; ++j) den = __fadd_rn(den, bufB[j]);
  }
  for (int d = nb * 2 * B; d < D; ++d) den = __fadd_rn(den, e[d * kTile]);
  return den;
}

template <class L>
__device__ __forceinline__ void tile_back_half(const FusedArgs& a, const TileId& t, int tid, float* s_par,
                                               const uint8_t* s_cen, float* s_ce, const float* s_red, float* s_min,
                                               float* s_inv) {
  constexpr int PS = L::PS;
  const FusedGeom& g = a.g;
  const int D = g.D;
  const size_t plane = (size_t)g.h * g.w;
  const size_t chan = plane * D;
  if (tid < 4 * kTile) {  // minima across the d-groups
    float v = kFill;
#pragma unroll
    for (int gq = 0; gq < kG2; ++gq) v = fminf(v, s_red[gq * 4 * kTile + tid]);
    s_min[tid] = v;
  }
  __syncthreads();        // s_min visible; staging area and s_red are dead from here on (-> s_ce)
  const int warp = tid >> 5, lane = tid & 31;
  const int q4 = (tid & 7) * 4;
  const int dl = tid >> 3;
  // 128-bit stores need 16-byte aligned rows
  const bool vec_ok = ((g.w & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
  float* orow = a.out + (size_t)t.n * a.out_channels * chan + (size_t)t.y * g.w + (t.x0 + q4);
  const int nlive = min(4, g.w - (t.x0 + q4));  // live pixels of this quad (<= 0: none)
  const bool vec = vec_ok && nlive == 4;
  if (vec) pass_e<true>(s_par, s_cen, s_ce, s_min, PS, q4, dl, D, orow, plane, chan, nlive, a.k_cen, a.k_ncc, a.k_sad);
  else pass_e<false>(s_par, s_cen, s_ce, s_min, PS, q4, dl, D, orow, plane, chan, nlive, a.k_cen, a.k_ncc, a.k_sad);
  __syncthreads();
  if (warp < 4) {
    const float mm = s_min[warp * kTile + lane];
    const float* e = (warp == 0 ? s_ce : s_par + (warp - 1) * PS) + lane;
    const float den = chain_sum(e, D);
    s_inv[warp * kTile + lane] = (mm == kFill) ? 0.f : 1.0f / den;
  }
  __syncthreads();
  if (vec) pass_n<true>(s_par, s_ce, s_inv, PS, q4, dl, D, orow + 4 * chan, plane, chan, nlive);
  else pass_n<false>(s_par, s_ce, s_inv, PS, q4, dl, D, orow + 4 * chan, plane, chan, nlive);
}

// ---- alternative back half: exponentials recomputed on the fly (no write-back) ----------------
// Channels 0-3 (cbmv_generator.py:283-287) for thread = (pixel quad q4, disparities d0, d0+16, ... < d1)
template <bool kVec>
__device__ __forceinline__ void store_ch03(const float* s_par, const uint8_t* s_cen, int PS, int q4, int d0, int d1,
                                           float* orow, size_t plane, size_t chan, int nlive) {
#pragma unroll 1
  for (int d = d0; d < d1; d += 16) {
    const float* e0 = s_par + d * kTile + q4;
    const uchar4 cb = *reinterpret_cast<const uchar4*>(s_cen + d * kTile + q4);
    const float4 v1 = *reinterpret_cast<const float4*>(e0);
    const float4 v2 = *reinterpret_cast<const float4*>(e0 + PS);
    const float4 v3 = *reinterpret_cast<const float4*>(e0 + 2 * PS);
    const float4 c0 = make_float4(census_ch0(cb.x), census_ch0(cb.y), census_ch0(cb.z), census_ch0(cb.w));
    const float4 c1 = make_float4(normalise_cost(v1.x, 1), normalise_cost(v1.y, 1), normalise_cost(v1.z, 1),
                                  normalise_cost(v1.w, 1));
    const float4 c2 = make_float4(normalise_cost(v2.x, 2), normalise_cost(v2.y, 2), normalise_cost(v2.z, 2),
                                  normalise_cost(v2.w, 2));
    const float4 c3 = make_float4(normalise_cost(v3.x, 3), normalise_cost(v3.y, 3), normalise_cost(v3.z, 3),
                                  normalise_cost(v3.w, 3));
    store_quads<kVec>(orow + (size_t)d * plane, chan, nlive, c0, c1, c2, c3);
  }
}

// Channels 4-7 = exp(-(c-m)^2/sigma) / den for thread = (pixel quad q4, disparities dl, dl+32, ...),
// exponentials recomputed from the parked costs, 128-bit row segments.
template <bool kVec>
__device__ __forceinline__ void phase3_quads(const float* s_par, const uint8_t* s_cen, const float* s_min,
                                             const float* s_inv, int PS, int q4, int dl, int D, float* arow,
                                             size_t plane, size_t chan, int nlive, float k0, float k1, float k2) {
  const float4 m_cen4 = *reinterpret_cast<const float4*>(s_min + q4);
  const float4 m1 = *reinterpret_cast<const float4*>(s_min + kTile + q4);
  const float4 m2 = *reinterpret_cast<const float4*>(s_min + 2 * kTile + q4);
  const float4 m3 = *reinterpret_cast<const float4*>(s_min + 3 * kTile + q4);
  const float4 i0 = *reinterpret_cast<const float4*>(s_inv + q4);
  const float4 i1 = *reinterpret_cast<const float4*>(s_inv + kTile + q4);
  const float4 i2 = *reinterpret_cast<const float4*>(s_inv + 2 * kTile + q4);
  const float4 i3 = *reinterpret_cast<const float4*>(s_inv + 3 * kTile + q4);
  const int mcx = (m_cen4.x == kFill) ? 0 : (int)m_cen4.x, mcy = (m_cen4.y == kFill) ? 0 : (int)m_cen4.y;
  const int mcz = (m_cen4.z == kFill) ? 0 : (int)m_cen4.z, mcw = (m_cen4.w == kFill) ? 0 : (int)m_cen4.w;
#pragma unroll 1
  for (int d = dl; d < D; d += 32) {
    const float* e0 = s_par + d * kTile + q4;
    const uchar4 cb = *reinterpret_cast<const uchar4*>(s_cen + d * kTile + q4);
    const float4 v1 = *reinterpret_cast<const float4*>(e0);
    const float4 v2 = *reinterpret_cast<const float4*>(e0 + PS);
    const float4 v3 = *reinterpret_cast<const float4*>(e0 + 2 * PS);
    const float4 a0 = make_float4(census_e(cb.x, mcx, k0) * i0.x, census_e(cb.y, mcy, k0) * i0.y,
                                  census_e(cb.z, mcz, k0) * i0.z, census_e(cb.w, mcw, k0) * i0.w);
    const float4 a1 = make_float4(aml_e(v1.x, m1.x, k1) * i1.x, aml_e(v1.y, m1.y, k1) * i1.y,
                                  aml_e(v1.z, m1.z, k1) * i1.z, aml_e(v1.w, m1.w, k1) * i1.w);
    const float4 a2 = make_float4(aml_e(v2.x, m2.x, k2) * i2.x, aml_e(v2.y, m2.y, k2) * i2.y,
                                  aml_e(v2.z, m2.z, k2) * i2.z, aml_e(v2.w, m2.w, k2) * i2.w);
    const float4 a3 = make_float4(aml_e(v3.x, m3.x, k2) * i3.x, aml_e(v3.y, m3.y, k2) * i3.y,
                                  aml_e(v3.z, m3.z, k2) * i3.z, aml_e(v3.w, m3.w, k2) * i3.w);
    store_quads<kVec>(arow + (size_t)d * plane, chan, nlive, a0, a1, a2, a3);
  }
}

template <class L>
__device__ __forceinline__ void tile_back_half_otf(const FusedArgs& a, const TileId& t, int tid, const float* s_par,
                                               const uint8_t* s_cen, const float* s_red, float* s_min,
                                               float* s_inv) {
  constexpr int PS = L::PS;
  const FusedGeom& g = a.g;
  const int D = g.D;
  const size_t plane = (size_t)g.h * g.w;
  const size_t chan = plane * D;
  if (tid < 4 * kTile) {  // minima across the d-groups
    float v = kFill;
#pragma unroll
    for (int gq = 0; gq < kG2; ++gq) v = fminf(v, s_red[gq * 4 * kTile + tid]);
    s_min[tid] = v;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int q4 = (tid & 7) * 4;
  // 128-bit stores need 16-byte aligned rows
  const bool vec_ok = ((g.w & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
  float* orow = a.out + (size_t)t.n * a.out_channels * chan + (size_t)t.y * g.w + (t.x0 + q4);
  const int nlive = min(4, g.w - (t.x0 + q4));  // live pixels of this quad (<= 0: none)
  const bool vec = vec_ok && nlive == 4;
  if (warp < 4) {
    // the reference's sequential fp32 sum over d (featextract.cpp:444-447), exponentials on the fly
    const float mm = s_min[warp * kTile + lane];
    float den = 0.f;
    const int Dfull = D & ~7;   // groups of 8 without guards, then a guarded tail
    if (warp == 0) {
      const int mc = (mm == kFill) ? 0 : (int)mm;
      const uint8_t* c = s_cen + lane;
      for (int d0 = 0; d0 < Dfull; d0 += 8, c += 8 * kTile) {
        float ev[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ev[j] = census_e(c[j * kTile], mc, a.k_cen);
#pragma unroll
        for (int j = 0; j < 8; ++j) den = __fadd_rn(den, ev[j]);
      }
      for (int d = Dfull; d < D; ++d, c += kTile) den = __fadd_rn(den, census_e(c[0], mc, a.k_cen));
    } else {
      const float kq = (warp == 1) ? a.k_ncc : a.k_sad;
      const float* e = s_par + (warp - 1) * PS + lane;
      for (int d0 = 0; d0 < Dfull; d0 += 8, e += 8 * kTile) {
        float ev[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ev[j] = aml_e(e[j * kTile], mm, kq);
#pragma unroll
        for (int j = 0; j < 8; ++j) den = __fadd_rn(den, ev[j]);
      }
      for (int d = Dfull; d < D; ++d, e += kTile) den = __fadd_rn(den, aml_e(e[0], mm, kq));
    }
    s_inv[warp * kTile + lane] = (mm == kFill) ? 0.f : 1.0f / den;
  } else {
    // channels 0-3: thread = (pixel quad, d), 16 disparities per sweep of the four warps
    if (vec) store_ch03<true>(s_par, s_cen, PS, q4, (tid >> 3) - 16, D, orow, plane, chan, nlive);
    else store_ch03<false>(s_par, s_cen, PS, q4, (tid >> 3) - 16, D, orow, plane, chan, nlive);
  }
  __syncthreads();
  const int dl = tid >> 3;
  if (vec) phase3_quads<true>(s_par, s_cen, s_min, s_inv, PS, q4, dl, D, orow + 4 * chan, plane, chan, nlive, a.k_cen, a.k_ncc, a.k_sad);
  else phase3_quads<false>(s_par, s_cen, s_min, s_inv, PS, q4, dl, D, orow + 4 * chan, plane, chan, nlive, a.k_cen, a.k_ncc, a.k_sad);
}

// Back half of a tile for DISPARITY-SLAB SHARDING (phase A of slab.cu): the AML
// minimum and denominator need the other ranks' disparities, so the tile only stores channels
// 0-3, parks the RAW costs in channels 4-7 (census as float; fill where there is no cost) and
// writes the slab's per-pixel minima; slab_phase_b/c finish the job after the all-reduces.
template <class L>
__device__ __forceinline__ void tile_slab_a(const FusedArgs& a, const TileId& t, int tid, const float* s_par,
                                            const uint8_t* s_cen, const float* s_red) {
  constexpr int PS = L::PS;
  const FusedGeom& g = a.g;
  const int D = g.D;
  const size_t plane = (size_t)g.h * g.w;
  const size_t chan = plane * a.out_D;
  if (tid < 4 * kTile) {  // minima across the d-groups -> global
    float v = kFill;
#pragma unroll
    for (int gq = 0; gq < kG2; ++gq) v = fminf(v, s_red[gq * 4 * kTile + tid]);
    const int m = tid / kTile, x = t.x0 + tid % kTile;
    if (x < g.w) {
      float* mp = a.mins + (((size_t)t.n * a.mins_planes + m) * g.h + t.y) * g.w + x;
      *mp = a.mins_accumulate ? fminf(*mp, v) : v;
    }
  }
  const int q4 = (tid & 7) * 4;
  const int dl = tid >> 3;
  const bool vec_ok = ((g.w & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
  float* orow = a.out + (size_t)t.n * a.out_channels * chan + (size_t)t.y * g.w + (t.x0 + q4);
  const int nlive = min(4, g.w - (t.x0 + q4));
  const bool vec = vec_ok && nlive == 4;
#pragma unroll 2
  for (int d = dl; d < D; d += 32) {
    const float* e0 = s_par + d * kTile + q4;
    const uchar4 cb = *reinterpret_cast<const uchar4*>(s_cen + d * kTile + q4);
    const float4 v1 = *reinterpret_cast<const float4*>(e0);
    const float4 v2 = *reinterpret_cast<const float4*>(e0 + PS);
    const float4 v3 = *reinterpret_cast<const float4*>(e0 + 2 * PS);
    const float4 v0 = make_float4(cb.x == 255 ? kFill : (float)cb.x, cb.y == 255 ? kFill : (float)cb.y,
                                  cb.z == 255 ? kFill : (float)cb.z, cb.w == 255 ? kFill : (float)cb.w);
    const float4 c0 = make_float4(census_ch0(cb.x), census_ch0(cb.y), census_ch0(cb.z), census_ch0(cb.w));
    const float4 c1 = make_float4(normalise_cost(v1.x, 1), normalise_cost(v1.y, 1), normalise_cost(v1.z, 1),
                                  normalise_cost(v1.w, 1));
    const float4 c2 = make_float4(normalise_cost(v2.x, 2), normalise_cost(v2.y, 2), normalise_cost(v2.z, 2),
                                  normalise_cost(v2.w, 2));
    const float4 c3 = make_float4(normalise_cost(v3.x, 3), normalise_cost(v3.y, 3), normalise_cost(v3.z, 3),
                                  normalise_cost(v3.w, 3));
    float* o = orow + (size_t)(a.out_d0 + d) * plane;
    if (vec) {
      store_quads<true>(o, chan, nlive, c0, c1, c2, c3);
      // parked raw costs are read again by slab_phase_b/c: plain (cached) stores
      *reinterpret_cast<float4*>(o + 4 * chan) = v0;
      *reinterpret_cast<float4*>(o + 5 * chan) = v1;
      *reinterpret_cast<float4*>(o + 6 * chan) = v2;
      *reinterpret_cast<float4*>(o + 7 * chan) = v3;
    } else {
      store_quads<false>(o, chan, nlive, c0, c1, c2, c3);
      const float rr[4][4] = {{v0.x, v0.y, v0.z, v0.w}, {v1.x, v1.y, v1.z, v1.w}, {v2.x, v2.y, v2.z, v2.w},
                              {v3.x, v3.y, v3.z, v3.w}};
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nlive) o[(4 + ch) * chan + i] = rr[ch][i];
    }
  }
}
