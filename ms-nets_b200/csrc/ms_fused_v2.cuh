// ms_fused_v2.cuh -- the tile body of ms_fused_kernel (included inside ms_fused.cu's anonymous
// namespace, after the TMA / packed-fp32 helpers and FusedArgs).
//
// Tile = (pair n, output row y, 32 consecutive x) x all D; one CTA of 256 threads per tile.
//
//   phase 1   thread = (pixel PAIR (2k, 2k+1), d-group g of 16).  Per step the thread evaluates
//             two voxels on a DIAGONAL of the volume: A = (x, d) and B = (x+1, d+1).  Both read
//             the SAME right-image column x-d, hence the same right census code, the same right
//             NCC/ZSAD statistics and the same 5x5 right window -- so every ZSAD tap is one packed
//             FADD2 with the right pixel broadcast to both halves (75 FADD2 per voxel pair, each
//             half an IEEE fp32 add in the reference's order, matchers.cpp:499-506), every NCC
//             product one packed FFMA2, and the shared-memory loads are paid once per pair.
//             Two loops so that neither outgrows the 128-register budget of 2 CTAs/SM:
//               loop CN  census (xor + popc) and NCC (9 FFMA2 of exact integers, fp64 scaling)
//               loop Z   ZSAD over a register-resident right window that slides one column per
//                        step (6 physical columns: the next column loads while this one computes)
//             Raw costs are parked in shared memory [d][32] (ncc, zsad floats; census byte); the
//             tile's SAD-of-Sobel costs arrive by TMA straight into their parking plane.
//   back half every AML exponential is evaluated ONCE:
//     pass E  thread = (pixel quad, d): channels 0-3 normalised and stored (128-bit rows), the
//             three float planes overwritten in place by exp(-(c-m)^2/sigma);
//     pass S  one thread per (pixel, matcher): the AML denominator, added sequentially in d
//             order as the reference does (featextract.cpp:444-447) -- loads and adds only;
//     pass N  thread = (pixel quad, d): channels 4-7 = e * (1/den), 128-bit rows.
#pragma once

constexpr int kG2 = 16;       // d-groups per tile (16 pixel pairs x 16 groups = 256 threads)
constexpr int kSlack2 = 24;   // right-image columns left of x-(D-1) that dummy steps (d >= D) may read

template <int DMAX>
struct Lay2 {
  static constexpr int SL = kSlack2;
  static constexpr int RW = (DMAX + kTile + SL + 3) & ~3;                // desc / stat entries: columns x-d, d in [-1, D-1+SL]
  static constexpr int RWF = RW + 12;                                    // float row: halo 2+2, alignment shift <= 3, 16-byte granules
  static constexpr size_t st_desc = 0;                                   // [RW] uint4 census codes
  static constexpr size_t st_stat = st_desc + (size_t)RW * 16;           // [RW] RStat
  static constexpr size_t st_rf = st_stat + (size_t)RW * 16;             // [5][RWF] float pixel rows
  static constexpr size_t st_mean = st_rf + (size_t)5 * RWF * 4;         // [RWF] ZSAD window means
  static constexpr size_t st_bytes = (st_mean + (size_t)RWF * 4 + 127) & ~(size_t)127;
  static constexpr int DS = DMAX + 1;                                    // parked rows; row DMAX.. is scratch for dummy steps
  static constexpr int PS = DS * kTile;                                  // floats per parked matcher
  static constexpr size_t off_red = st_bytes;                            // [kG2][4][32] per-group minima
  static constexpr size_t off_min = off_red + (size_t)kG2 * 4 * kTile * 4;   // [4][32]
  static constexpr size_t off_inv = off_min + 4 * kTile * 4;             // [4][32]
  static constexpr size_t off_lut = off_inv + 4 * kTile * 4;             // [256] census AML exponentials (0 from 121 on)
  static constexpr size_t off_lutn = off_lut + 256 * 4;                  // [256] census byte -> channel 0
  static constexpr size_t off_par = (off_lutn + 256 * 4 + 127) & ~(size_t)127;   // [3][DS][32] floats; plane 1 is a TMA destination
  static constexpr size_t pk_cen = (size_t)3 * PS * 4;                   // then [DS][32] census bytes
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

// ---- staging (right-image row data of the tile) ---------------------------------------------
// Entry i of the desc/stat rows is bordered column XbaseP + i (padded coordinates); the column of
// pixel px at local disparity step d is ir = px + SL + (D-1) - d, d in [-1, D-1+SL].
template <class L>
__device__ __forceinline__ int stage2_xbase(const FusedArgs& a, const TileId& t) {
  return t.x0 + a.g.bwl - (a.g.d0 + a.g.D - 1) - L::SL + a.g.padL;
}

template <class L>
__device__ __forceinline__ void stage2_rows_tma(const FusedArgs& a, const TileId& t, unsigned char* buf,
                                                unsigned long long* bar) {
  const FusedGeom& g = a.g;
  const int RWn = g.D + kTile + L::SL;
  const int XbaseP = stage2_xbase<L>(a, t);
  const int Yp = t.y + g.bh + kPadT;
  const size_t img_off = (size_t)t.n * g.img_px();
  const int fstart = (XbaseP - 2) & ~3;
  const unsigned row_bytes = (unsigned)RWn * 16u;
  const unsigned frow_bytes = (unsigned)((RWn + 4 + 3 + 3) >> 2) * 16u;
  const int mstart = XbaseP & ~3;
  const unsigned mrow_bytes = (unsigned)((RWn + 3 + 3) >> 2) * 16u;
  mbar_expect_tx(bar, 2u * row_bytes + 5u * frow_bytes + mrow_bytes);
  bulk_load(buf + L::st_desc, a.descR + img_off + (size_t)Yp * g.Wp + XbaseP, row_bytes, bar);
  bulk_load(buf + L::st_stat, a.statR + img_off + (size_t)Yp * g.Wp + XbaseP, row_bytes, bar);
  float* s_rf = reinterpret_cast<float*>(buf + L::st_rf);
#pragma unroll
  for (int r = 0; r < 5; ++r)
    bulk_load(s_rf + r * L::RWF, a.fR + img_off + (size_t)(Yp - 2 + r) * g.Wp + fstart, frow_bytes, bar);
  bulk_load(buf + L::st_mean, a.meanR0 + img_off + (size_t)Yp * g.Wp + mstart, mrow_bytes, bar);
}

template <class L, int NT>
__device__ __forceinline__ void stage2_rows_ldgsts(const FusedArgs& a, const TileId& t, unsigned char* buf) {
  const FusedGeom& g = a.g;
  const int RWn = g.D + kTile + L::SL;
  const int XbaseP = stage2_xbase<L>(a, t);
  const int Yp = t.y + g.bh + kPadT;
  const size_t img_off = (size_t)t.n * g.img_px();
  const uint4* gd = a.descR + img_off + (size_t)Yp * g.Wp + XbaseP;
  const uint4* gs = reinterpret_cast<const uint4*>(a.statR + img_off + (size_t)Yp * g.Wp + XbaseP);
  uint4* s_desc = reinterpret_cast<uint4*>(buf + L::st_desc);
  uint4* s_stat = reinterpret_cast<uint4*>(buf + L::st_stat);
  for (int i = threadIdx.x; i < RWn; i += NT) {
    cp_async16(s_desc + i, gd + i);
    cp_async16(s_stat + i, gs + i);
  }
  float* s_rf = reinterpret_cast<float*>(buf + L::st_rf);
  const int fstart = (XbaseP - 2) & ~3;
  const int nvec = (RWn + 4 + 3 + 3) >> 2;
  for (int i = threadIdx.x; i < 5 * nvec; i += NT) {
    const int r = i / nvec, v = i - r * nvec;
    cp_async16(s_rf + r * L::RWF + 4 * v, a.fR + img_off + (size_t)(Yp - 2 + r) * g.Wp + fstart + 4 * v);
  }
  float* s_mean = reinterpret_cast<float*>(buf + L::st_mean);
  const int mstart = XbaseP & ~3;
  const int nvm = (RWn + 3 + 3) >> 2;
  for (int i = threadIdx.x; i < nvm; i += NT)
    cp_async16(s_mean + 4 * i, a.meanR0 + img_off + (size_t)Yp * g.Wp + mstart + 4 * i);
}

// ---- left-image data of a pixel pair: straight from global memory (before the staging wait) ---
struct Left2 {
  uint4 descA, descB;
  uint4 statA, statB;   // RStat bits
  float px[5][6];       // rows y-2..y+2, columns xA-2 .. xA+3
};
__device__ __forceinline__ void load_left2(const FusedArgs& a, const TileId& t, int pr, Left2& lr) {
  const FusedGeom& g = a.g;
  const int Yp = t.y + g.bh + kPadT;
  const int Xp = t.x0 + 2 * pr + g.bwl + g.padL;
  const size_t img_off = (size_t)t.n * g.img_px();
  const uint4* dp = a.descL + img_off + (size_t)Yp * g.Wp + Xp;
  const uint4* sp = reinterpret_cast<const uint4*>(a.statL + img_off + (size_t)Yp * g.Wp + Xp);
  lr.descA = __ldg(dp);
  lr.descB = __ldg(dp + 1);
  lr.statA = __ldg(sp);
  lr.statB = __ldg(sp + 1);
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const float* gf = a.fL + img_off + (size_t)(Yp - 2 + r) * g.Wp + (Xp - 2);
#pragma unroll
    for (int c = 0; c < 6; ++c) lr.px[r][c] = __ldg(gf + c);
  }
}

// per-thread description of its share of a tile
struct P1Ctx {
  int pr;          // pixel pair: tile pixels 2*pr (voxel A) and 2*pr+1 (voxel B)
  int dA0;         // local disparity of voxel A at step 0 (group 0 starts at -1: its B covers d = 0)
  int nsteps;      // steps of this thread
  int ir0;         // staged column index at step 0
  int shift;       // alignment shift of the float rows
  int mofs;        // alignment shift of the mean row
  int dmaxA[3], dmaxB[3];   // largest local d with a cost: census, ncc, zsad (-1: none)
  bool lastB_dummy;         // the thread's last step has dB == D (fast path only: last group)
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

// ---- loop CN: census + NCC ------------------------------------------------------------------
// kFast: interior tile whose d-groups cover D exactly and whose step count is a multiple of 6 --
// no validity selects, no row clamps, one basic block per 6 steps.
template <class L, bool kFast>
__device__ __forceinline__ void p1_census_ncc(const FusedArgs& a, const unsigned char* stage, float* s_par,
                                              uint8_t* s_cen, const Left2& lr, const P1Ctx& c, P1Min& mn) {
  const int D = a.g.D;
  const RStat lsA = *reinterpret_cast<const RStat*>(&lr.statA);
  const RStat lsB = *reinterpret_cast<const RStat*>(&lr.statB);
  f32x2 l3[3][3];   // (A, B) left pixels of the 3x3 NCC windows
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      // + (-0.0) is exact; the run-time operand keeps each pair in registers of its own (ptxas would
      // otherwise rebuild the overlapping pairs from the left window with MOVs at every step)
      l3[r][j] = add2(pk2(lr.px[r + 1][1 + j], lr.px[r + 1][2 + j]), pk2(a.neg_zero, a.neg_zero));
    }
  const f32x2 lA2 = pk2(lsA.A, lsB.A);
  const f32x2 nine2 = pk2(9.0f, 9.0f);

  const uint4* dscp = reinterpret_cast<const uint4*>(stage + L::st_desc) + c.ir0;
  const uint4* sttp = reinterpret_cast<const uint4*>(stage + L::st_stat) + c.ir0;
  // float column of right pixel (x - d - 1): s_rf index shift + ir + 1
  const float* rfp = reinterpret_cast<const float*>(stage + L::st_rf) + L::RWF + c.shift + c.ir0 + 1;
  float w3[3][3];   // sliding 3x3 right window; logical column j lives in w3[.][(j + ROT) % 3]
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int j = 0; j < 3; ++j) w3[r][j] = rfp[r * L::RWF + j];
  float* pn = s_par + 2 * c.pr;              // ncc plane
  uint8_t* pc = s_cen + 2 * c.pr;
  int dA = c.dA0;

  for (int base = 0; base < c.nsteps; base += 6) {
#pragma unroll
    for (int sI = 0; sI < 6; ++sI) {
      if (!kFast && sI > 0 && base + sI >= c.nsteps) break;
#define W3(r, j) w3[r][((j) + 12 - sI) % 3]
      const int dB = dA + 1;
      const uint4 rd = dscp[-sI];
      const uint4 rs_raw = sttp[-sI];
      const RStat rs = *reinterpret_cast<const RStat*>(&rs_raw);
      // census: Hamming distance of the packed codes (matchers.cpp:323-337)
      int cenA = popc128(lr.descA, rd);
      int cenB = popc128(lr.descB, rd);
      // NCC: P exact in fp32 (< 2^24); scaling in fp64 left to right (matchers.cpp:200-201)
      f32x2 P = pk2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float w = W3(r, j);
          P = fma2(l3[r][j], pk2(w, w), P);
        }
      const float nra = -rs.A;
      const f32x2 num2 = fma2(nine2, P, mul2(lA2, pk2(nra, nra)));   // 9P - A_L*A_R, exact
      float numA, numB;
      upk2(num2, numA, numB);
      float nccA = ncc_scale(numA, lsA.C, rs.C);
      float nccB = ncc_scale(numB, lsB.C, rs.C);
      int rowA = dA, rowB = dB;
      if (!kFast) {
        cenA = (dA >= 0 && dA <= c.dmaxA[0]) ? cenA : 255;
        cenB = (dB <= c.dmaxB[0]) ? cenB : 255;
        nccA = (dA >= 0 && dA <= c.dmaxA[1]) ? nccA : kFill;
        nccB = (dB <= c.dmaxB[1]) ? nccB : kFill;
        rowA = (dA >= 0 && dA < D) ? dA : D;     // dummy steps park into the scratch row
        rowB = (dB < D) ? dB : D;
      }
      pc[rowA * kTile] = (uint8_t)cenA;
      pc[rowB * kTile + 1] = (uint8_t)cenB;
      pn[rowA * kTile] = nccA;
      pn[rowB * kTile + 1] = nccB;
      mn.cenA = min(mn.cenA, cenA);
      mn.nccA = fminf(mn.nccA, nccA);
      if (!(kFast && sI == 5 && c.lastB_dummy && base + 6 >= c.nsteps)) {
        mn.cenB = min(mn.cenB, cenB);
        mn.nccB = fminf(mn.nccB, nccB);
      }
      // slide the window one column left: the next step's logical column 0
#pragma unroll
      for (int r = 0; r < 3; ++r) w3[r][(0 + 12 - (sI + 1)) % 3] = rfp[r * L::RWF - 1 - sI];
      dA += 1;
#undef W3
    }
    dscp -= 6; sttp -= 6; rfp -= 6;
  }
}

// ---- loop Z: ZSAD ---------------------------------------------------------------------------
// 25 taps row-major, ((L - mL) - R) + mR, sequential fp32 per voxel (matchers.cpp:499-506); the
// two voxels of the pair occupy the two halves of every packed operation.
template <class L, bool kFast>
__device__ __forceinline__ void p1_zsad(const FusedArgs& a, const unsigned char* stage, float* s_par,
                                        const Left2& lr, const P1Ctx& c, P1Min& mn) {
  const int D = a.g.D;
  const float mLA = reinterpret_cast<const RStat*>(&lr.statA)->mean;
  const float mLB = reinterpret_cast<const RStat*>(&lr.statB)->mean;
  f32x2 ap[5][5];   // (L_A[tap] - mL_A, L_B[tap] - mL_B), hoisted over all d
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int j = 0; j < 5; ++j) ap[r][j] = pk2(__fsub_rn(lr.px[r][j], mLA), __fsub_rn(lr.px[r][j + 1], mLB));
  // float column (x - d - 2) of the step: s_rf index shift + ir
  const float* rfp = reinterpret_cast<const float*>(stage + L::st_rf) + c.shift + c.ir0;
  const float* mnp = reinterpret_cast<const float*>(stage + L::st_mean) + c.mofs + c.ir0;
  float wv[5][6];   // sliding right window, 5 logical columns + the one being loaded for the next step
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int j = 0; j < 5; ++j) wv[r][j] = rfp[r * L::RWF + j];
  float* pz = s_par + 2 * L::PS + 2 * c.pr;   // zsad plane
  int dA = c.dA0;

  for (int base = 0; base < c.nsteps; base += 6) {
#pragma unroll
    for (int sI = 0; sI < 6; ++sI) {
      if (!kFast && sI > 0 && base + sI >= c.nsteps) break;
#define WV(r, j) wv[r][((j) + 12 - sI) % 6]
      const int dB = dA + 1;
      // next step's new left column into the spare slot (logical column -1 of this step)
#pragma unroll
      for (int r = 0; r < 5; ++r) WV(r, 5) = rfp[r * L::RWF - 1 - sI];
      const float mR = mnp[-sI];
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
      if (!kFast) {
        zA = (dA >= 0 && dA <= c.dmaxA[2]) ? zA : kFill;
        zB = (dB <= c.dmaxB[2]) ? zB : kFill;
        rowA = (dA >= 0 && dA < D) ? dA : D;
        rowB = (dB < D) ? dB : D;
      }
      pz[rowA * kTile] = zA;
      pz[rowB * kTile + 1] = zB;
      mn.sadA = fminf(mn.sadA, zA);
      if (!(kFast && sI == 5 && c.lastB_dummy && base + 6 >= c.nsteps)) mn.sadB = fminf(mn.sadB, zB);
      dA += 1;
#undef WV
    }
    rfp -= 6; mnp -= 6;
  }
}

// Group 0's extra step (fast path): voxel B = (odd pixel, d = 0) alone; A would be d = -1.
template <class L>
__device__ __forceinline__ void p1_extra_b0(const FusedArgs& a, const unsigned char* stage, float* s_par,
                                            uint8_t* s_cen, const Left2& lr, int pr, int ir, int shift, int mofs,
                                            P1Min& mn) {
  const RStat lsB = *reinterpret_cast<const RStat*>(&lr.statB);
  const uint4 rd = reinterpret_cast<const uint4*>(stage + L::st_desc)[ir];
  const uint4 rs_raw = reinterpret_cast<const uint4*>(stage + L::st_stat)[ir];
  const RStat rs = *reinterpret_cast<const RStat*>(&rs_raw);
  const float* rf = reinterpret_cast<const float*>(stage + L::st_rf) + shift + ir;   // column x - d - 2
  const float mR = reinterpret_cast<const float*>(stage + L::st_mean)[mofs + ir];
  const int cen = popc128(lr.descB, rd);
  float P = 0.f, z = 0.f;
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float w = rf[r * L::RWF + j];
      if (r >= 1 && r <= 3 && j >= 1 && j <= 3) P = __fmaf_rn(lr.px[r][j + 1], w, P);
      z = __fadd_rn(z, fabsf(__fadd_rn(__fsub_rn(__fsub_rn(lr.px[r][j + 1], lsB.mean), w), mR)));
    }
  const float num = __fmaf_rn(9.0f, P, -__fmul_rn(lsB.A, rs.A));
  const float ncc = ncc_scale(num, lsB.C, rs.C);
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
// pass E for thread = (pixel quad q4, disparities dl, dl+32, ...): channels 0-3 stored, the float
// planes overwritten by their AML exponentials.
template <bool kVec>
__device__ __forceinline__ void pass_e(float* s_par, const uint8_t* s_cen, const float* s_lutn, const float* s_min,
                                       int PS, int q4, int dl, int D, float* orow, size_t plane, size_t chan,
                                       int nlive, float k1, float k2) {
  const float4 m1 = *reinterpret_cast<const float4*>(s_min + kTile + q4);
  const float4 m2 = *reinterpret_cast<const float4*>(s_min + 2 * kTile + q4);
  const float4 m3 = *reinterpret_cast<const float4*>(s_min + 3 * kTile + q4);
#pragma unroll 1
  for (int d = dl; d < D; d += 32) {
    float* e0 = s_par + d * kTile + q4;
    const uchar4 cb = *reinterpret_cast<const uchar4*>(s_cen + d * kTile + q4);
    const float4 v1 = *reinterpret_cast<const float4*>(e0);
    const float4 v2 = *reinterpret_cast<const float4*>(e0 + PS);
    const float4 v3 = *reinterpret_cast<const float4*>(e0 + 2 * PS);
    const float4 c0 = make_float4(s_lutn[cb.x], s_lutn[cb.y], s_lutn[cb.z], s_lutn[cb.w]);
    const float4 c1 = make_float4(normalise_cost(v1.x, 1), normalise_cost(v1.y, 1), normalise_cost(v1.z, 1),
                                  normalise_cost(v1.w, 1));
    const float4 c2 = make_float4(normalise_cost(v2.x, 2), normalise_cost(v2.y, 2), normalise_cost(v2.z, 2),
                                  normalise_cost(v2.w, 2));
    const float4 c3 = make_float4(normalise_cost(v3.x, 3), normalise_cost(v3.y, 3), normalise_cost(v3.z, 3),
                                  normalise_cost(v3.w, 3));
    store_quads<kVec>(orow + (size_t)d * plane, chan, nlive, c0, c1, c2, c3);
    *reinterpret_cast<float4*>(e0) =
        make_float4(aml_e(v1.x, m1.x, k1), aml_e(v1.y, m1.y, k1), aml_e(v1.z, m1.z, k1), aml_e(v1.w, m1.w, k1));
    *reinterpret_cast<float4*>(e0 + PS) =
        make_float4(aml_e(v2.x, m2.x, k2), aml_e(v2.y, m2.y, k2), aml_e(v2.z, m2.z, k2), aml_e(v2.w, m2.w, k2));
    *reinterpret_cast<float4*>(e0 + 2 * PS) =
        make_float4(aml_e(v3.x, m3.x, k2), aml_e(v3.y, m3.y, k2), aml_e(v3.z, m3.z, k2), aml_e(v3.w, m3.w, k2));
  }
}

// pass N: channels 4-7 = e / den
template <bool kVec>
__device__ __forceinline__ void pass_n(const float* s_par, const uint8_t* s_cen, const float* s_lut,
                                       const float* s_min, const float* s_inv, int PS, int q4, int dl, int D,
                                       float* arow, size_t plane, size_t chan, int nlive) {
  const float4 m_cen4 = *reinterpret_cast<const float4*>(s_min + q4);
  const float4 i0 = *reinterpret_cast<const float4*>(s_inv + q4);
  const float4 i1 = *reinterpret_cast<const float4*>(s_inv + kTile + q4);
  const float4 i2 = *reinterpret_cast<const float4*>(s_inv + 2 * kTile + q4);
  const float4 i3 = *reinterpret_cast<const float4*>(s_inv + 3 * kTile + q4);
  // census exponentials come from the table: entry (byte - min), >= 121 -> 0 (no cost)
  const float* l0 = s_lut - ((m_cen4.x == kFill) ? 0 : (int)m_cen4.x);
  const float* l1 = s_lut - ((m_cen4.y == kFill) ? 0 : (int)m_cen4.y);
  const float* l2 = s_lut - ((m_cen4.z == kFill) ? 0 : (int)m_cen4.z);
  const float* l3 = s_lut - ((m_cen4.w == kFill) ? 0 : (int)m_cen4.w);
#pragma unroll 1
  for (int d = dl; d < D; d += 32) {
    const float* e0 = s_par + d * kTile + q4;
    const uchar4 cb = *reinterpret_cast<const uchar4*>(s_cen + d * kTile + q4);
    const float4 v1 = *reinterpret_cast<const float4*>(e0);
    const float4 v2 = *reinterpret_cast<const float4*>(e0 + PS);
    const float4 v3 = *reinterpret_cast<const float4*>(e0 + 2 * PS);
    const float4 a0 = make_float4(l0[cb.x] * i0.x, l1[cb.y] * i0.y, l2[cb.z] * i0.z, l3[cb.w] * i0.w);
    const float4 a1 = make_float4(v1.x * i1.x, v1.y * i1.y, v1.z * i1.z, v1.w * i1.w);
    const float4 a2 = make_float4(v2.x * i2.x, v2.y * i2.y, v2.z * i2.z, v2.w * i2.w);
    const float4 a3 = make_float4(v3.x * i3.x, v3.y * i3.y, v3.z * i3.z, v3.w * i3.w);
    store_quads<kVec>(arow + (size_t)d * plane, chan, nlive, a0, a1, a2, a3);
  }
}

template <class L>
__device__ __forceinline__ void tile_back_half2(const FusedArgs& a, const TileId& t, int tid, float* s_par,
                                                const uint8_t* s_cen, const float* s_red, float* s_min, float* s_inv,
                                                const float* s_lut, const float* s_lutn) {
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
  const int dl = tid >> 3;
  const bool vec_ok = ((g.w & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
  float* orow = a.out + (size_t)t.n * a.out_channels * chan + (size_t)t.y * g.w + (t.x0 + q4);
  const int nlive = min(4, g.w - (t.x0 + q4));  // live pixels of this quad (<= 0: none)
  const bool vec = vec_ok && nlive == 4;
  if (vec) pass_e<true>(s_par, s_cen, s_lutn, s_min, PS, q4, dl, D, orow, plane, chan, nlive, a.k_ncc, a.k_sad);
  else pass_e<false>(s_par, s_cen, s_lutn, s_min, PS, q4, dl, D, orow, plane, chan, nlive, a.k_ncc, a.k_sad);
  __syncthreads();
  if (warp < 4) {
    // pass S: the reference's sequential fp32 sum over d (featextract.cpp:444-447)
    const float mm = s_min[warp * kTile + lane];
    float den = 0.f;
    const int Dfull = D & ~7;
    if (warp == 0) {
      const float* lt = s_lut - ((mm == kFill) ? 0 : (int)mm);
      const uint8_t* c = s_cen + lane;
      for (int d0 = 0; d0 < Dfull; d0 += 8, c += 8 * kTile) {
        float ev[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ev[j] = lt[c[j * kTile]];
#pragma unroll
        for (int j = 0; j < 8; ++j) den = __fadd_rn(den, ev[j]);
      }
      for (int d = Dfull; d < D; ++d, c += kTile) den = __fadd_rn(den, lt[c[0]]);
    } else {
      const float* e = s_par + (warp - 1) * PS + lane;
      for (int d0 = 0; d0 < Dfull; d0 += 8, e += 8 * kTile) {
        float ev[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ev[j] = e[j * kTile];
#pragma unroll
        for (int j = 0; j < 8; ++j) den = __fadd_rn(den, ev[j]);
      }
      for (int d = Dfull; d < D; ++d, e += kTile) den = __fadd_rn(den, e[0]);
    }
    s_inv[warp * kTile + lane] = (mm == kFill) ? 0.f : 1.0f / den;
  }
  __syncthreads();
  if (vec) pass_n<true>(s_par, s_cen, s_lut, s_min, s_inv, PS, q4, dl, D, orow + 4 * chan, plane, chan, nlive);
  else pass_n<false>(s_par, s_cen, s_lut, s_min, s_inv, PS, q4, dl, D, orow + 4 * chan, plane, chan, nlive);
}
