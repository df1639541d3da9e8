/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the MS-Nets matching-space hot path.
 *
 * A plain-C restatement of what the reference's two Boost.Python libraries
 * compute (reference: /root/reference/src/cpp/matchers/matchers.cpp and
 * /root/reference/src/cpp/featextract/featextract.cpp) plus the NumPy glue of
 * src/dataloader/cbmv_generator.py and the soft-argmin of
 * src/models/gcnet_3dcnn.py.  Every function cites the reference lines it
 * follows.  It is NOT a copy: integer-exact quantities (census, NCC sums) are
 * computed by direct window sums instead of the reference's SSE / integral
 * images; floating-point quantities replay the reference's operation ORDER so
 * that results are bit-identical (pinned against oracle/_ref, the unmodified
 * reference compiled here, by tests/test_oracle_vs_ref.py and the committed
 * fixtures under tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link/call this file.  The product path (ms-nets_b200/) never does.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -fPIC -shared
 *        (see oracle/build_oracle.py).  No -ffast-math: rounding order matters.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* matchers.cpp:65,251,377,462 -- std::fill_n(res, n, RAND_MAX): the int
 * 2147483647 converted to float is 2147483648.0f. */
#define ORC_FILL 2147483648.0f

static void fill_f32(float* p, size_t n, float v) {
  for (size_t i = 0; i < n; ++i) p[i] = v;
}

/* ---------------------------------------------------------------- census --
 * matchers.cpp:232-353.  Census bit k of pixel (y,x) is [centre < tap_k] over
 * the wsize x wsize window, taps row-major (:285-297); cost is the number of
 * differing bits between L(y,x) and R(y,x-d) (:323-337; the zero-padded lanes
 * 121..127 never differ).  Written only for window origins i < H-wsize,
 * j < W-wsize and d <= j (:315-319).  Output layout [H][W][D], D innermost
 * (:244,337). */
void orc_census(const uint8_t* L, const uint8_t* R, int H, int W, int D, int wsize,
                float* out) {
  const int wc = wsize / 2;
  fill_f32(out, (size_t)H * W * D, ORC_FILL);
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < H - wsize; ++i) {
    for (int j = 0; j < W - wsize; ++j) {
      const int dend = (D < j + 1) ? D : j + 1;
      const int cl = L[(size_t)(i + wc) * W + (j + wc)];
      for (int d = 0; d < dend; ++d) {
        const int jr = j - d;
        const int cr = R[(size_t)(i + wc) * W + (jr + wc)];
        int diff = 0;
        for (int a = 0; a < wsize; ++a) {
          const uint8_t* lrow = L + (size_t)(i + a) * W + j;
          const uint8_t* rrow = R + (size_t)(i + a) * W + jr;
          for (int b = 0; b < wsize; ++b) diff += ((cl < lrow[b]) != (cr < rrow[b]));
        }
        out[((size_t)(i + wc) * W + (j + wc)) * D + d] = (float)diff;
      }
    }
  }
}

/* ------------------------------------------------------------- nccNister --
 * matchers.cpp:47-228.  A = sum I, B = sum I^2 over the window (exact integers,
 * :125-143), C = 1/sqrt(w^2*B - A*A) in double (:146-147); per disparity
 * P = sum L*R(.-d) (exact, :155-194); cost = (float)( -(w^2*P - A_L*A_R)*C_L*C_R )
 * evaluated left to right in double (:200-201), or 1.0f when either C is not
 * finite (:196,204).  Window origins i < H-wsize, d <= j < W-wsize (:186,190).
 * Output layout [D][H][W] (:58). */
void orc_ncc(const uint8_t* L, const uint8_t* R, int H, int W, int D, int wsize,
             float* out) {
  const int wc = wsize / 2;
  const int sq = wsize * wsize;
  fill_f32(out, (size_t)H * W * D, ORC_FILL);
  uint64_t* Al = (uint64_t*)calloc((size_t)H * W, sizeof(uint64_t));
  uint64_t* Ar = (uint64_t*)calloc((size_t)H * W, sizeof(uint64_t));
  double* Cl = (double*)calloc((size_t)H * W, sizeof(double));
  double* Cr = (double*)calloc((size_t)H * W, sizeof(double));
#pragma omp parallel for
  for (int i = 0; i < H - wsize; ++i) {
    for (int j = 0; j < W - wsize; ++j) {
      uint64_t al = 0, ar = 0, bl = 0, br = 0;
      for (int a = 0; a < wsize; ++a)
        for (int b = 0; b < wsize; ++b) {
          const uint64_t l = L[(size_t)(i + a) * W + j + b];
          const uint64_t r = R[(size_t)(i + a) * W + j + b];
          al += l; ar += r; bl += l * l; br += r * r;
        }
      const size_t c = (size_t)(i + wc) * W + (j + wc);
      Al[c] = al; Ar[c] = ar;
      Cl[c] = 1 / sqrt((double)((uint64_t)sq * bl) - (double)al * (double)al);
      Cr[c] = 1 / sqrt((double)((uint64_t)sq * br) - (double)ar * (double)ar);
    }
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (int d = 0; d < D; ++d) {
    for (int i = 0; i < H - wsize; ++i) {
      for (int j = d; j < W - wsize; ++j) {
        uint64_t p = 0;
        for (int a = 0; a < wsize; ++a)
          for (int b = 0; b < wsize; ++b)
            p += (uint64_t)L[(size_t)(i + a) * W + j + b] *
                 (uint64_t)R[(size_t)(i + a) * W + j + b - d];
        const size_t cl = (size_t)(i + wc) * W + (j + wc);
        const size_t cr = (size_t)(i + wc) * W + (j - d + wc);
        float v;
        if (isfinite(Cl[cl]) && isfinite(Cr[cr])) {
          const double num = (double)sq * (double)p - (double)(Al[cl] * Ar[cr]);
          const double t = -num * Cl[cl] * Cr[cr];
          v = (float)t;
        } else {
          v = 1.0f;
        }
        out[(size_t)d * H * W + cl] = v;
      }
    }
  }
  free(Al); free(Ar); free(Cl); free(Cr);
}

/* ------------------------------------------------------------------ zsad --
 * matchers.cpp:442-512.  Window means: float sum of the taps then one fp32
 * division by w^2 (:472-485).  Cost: fp32 accumulation, taps row-major, each
 * term |((L - mL) - R) + mR| evaluated left to right in fp32 (:499-506) with
 * mL at the left centre and mR at the right centre (x-d).  Window origins
 * i < H-wsize, d <= j < W-wsize.  Output [D][H][W]. */
void orc_zsad(const uint8_t* L, const uint8_t* R, int H, int W, int D, int wsize,
              float* out) {
  const int wc = wsize / 2;
  const int sq = wsize * wsize;
  fill_f32(out, (size_t)H * W * D, ORC_FILL);
  float* ml = (float*)calloc((size_t)H * W, sizeof(float));
  float* mr = (float*)calloc((size_t)H * W, sizeof(float));
#pragma omp parallel for
  for (int i = 0; i < H - wsize; ++i)
    for (int j = 0; j < W - wsize; ++j) {
      float sl = 0.f, sr = 0.f;
      for (int a = 0; a < wsize; ++a)
        for (int b = 0; b < wsize; ++b) {
          sl += (float)L[(size_t)(i + a) * W + j + b];
          sr += (float)R[(size_t)(i + a) * W + j + b];
        }
      ml[(size_t)(i + wc) * W + j + wc] = sl / (float)sq;
      mr[(size_t)(i + wc) * W + j + wc] = sr / (float)sq;
    }
#pragma omp parallel for schedule(dynamic, 1)
  for (int d = 0; d < D; ++d)
    for (int i = 0; i < H - wsize; ++i)
      for (int j = d; j < W - wsize; ++j) {
        const float mL = ml[(size_t)(i + wc) * W + j + wc];
        const float mR = mr[(size_t)(i + wc) * W + j - d + wc];
        volatile float acc = 0.f; /* volatile: forbid any re-association */
        for (int a = 0; a < wsize; ++a)
          for (int b = 0; b < wsize; ++b) {
            float t = (float)L[(size_t)(i + a) * W + j + b] - mL;
            t = t - (float)R[(size_t)(i + a) * W + j - d + b];
            t = t + mR;
            acc = acc + fabsf(t);
          }
        out[(size_t)d * H * W + (size_t)(i + wc) * W + j + wc] = acc;
      }
  free(ml); free(mr);
}

/* ----------------------------------------------------------------- sobel --
 * matchers.cpp:515-554.  Horizontal 3x3 Sobel in integer arithmetic, written
 * at (i+1,j+1) for i < H-3, j < W-3 (:538-547); zero elsewhere (:527). */
void orc_sobel(const uint8_t* img, int H, int W, float* out) {
  fill_f32(out, (size_t)H * W, 0.f);
  for (int i = 0; i < H - 3; ++i)
    for (int j = 0; j < W - 3; ++j) {
      const uint8_t* p = img + (size_t)i * W + j;
      const int g = (p[2] - p[0]) + 2 * (p[W + 2] - p[W]) + (p[2 * W + 2] - p[2 * W]);
      out[(size_t)(i + 1) * W + j + 1] = (float)g;
    }
}

/* ---------------------------------------------------------------- sadsob --
 * matchers.cpp:356-438.  Per disparity an fp32 summed-area table of
 * |L(i,j) - R(i,j-d)| (zero for j < d, :388-394) is built by a vertical prefix
 * pass then a horizontal prefix pass (:396-411), both sequential fp32 adds, and
 * the box is ((br - bl) - tr) + tl (:421-423).  The rounding of the running
 * sums is part of the reference's result, so the scan order is replayed
 * exactly.  Window origins i < H-wsize, d <= j < W-wsize.  Output [D][H][W]. */
void orc_sadsob(const float* L, const float* R, int H, int W, int D, int wsize,
                float* out) {
  const int wc = wsize / 2;
  const int IH = H + 1, IW = W + 1;
  fill_f32(out, (size_t)H * W * D, ORC_FILL);
#pragma omp parallel
  {
    float* S = (float*)malloc((size_t)IH * IW * sizeof(float));
#pragma omp for schedule(dynamic, 1)
    for (int d = 0; d < D; ++d) {
      memset(S, 0, (size_t)IH * IW * sizeof(float));
      for (int i = 0; i < H; ++i)
        for (int j = d; j < W; ++j)
          S[(size_t)(i + 1) * IW + j + 1] = fabsf(L[(size_t)i * W + j] - R[(size_t)i * W + j - d]);
      for (int i = 1; i < IH; ++i)
        for (int j = 0; j < IW; ++j) {
          volatile float t = S[(size_t)i * IW + j] + S[(size_t)(i - 1) * IW + j];
          S[(size_t)i * IW + j] = t;
        }
      for (int i = 0; i < IH; ++i)
        for (int j = 1; j < IW; ++j) {
          volatile float t = S[(size_t)i * IW + j] + S[(size_t)i * IW + j - 1];
          S[(size_t)i * IW + j] = t;
        }
      for (int i = 0; i < H - wsize; ++i)
        for (int j = d; j < W - wsize; ++j) {
          volatile float t = S[(size_t)(i + wsize) * IW + j + wsize] - S[(size_t)(i + wsize) * IW + j];
          t = t - S[(size_t)i * IW + j + wsize];
          t = t + S[(size_t)i * IW + j];
          out[(size_t)d * H * W + (size_t)(i + wc) * W + j + wc] = t;
        }
    }
    free(S);
  }
}

/* ------------------------------------------------------------ axis moves --
 * featextract.cpp:49-76 (swap_axes: [D][H][W] -> [H][W][D]) and :78-105. */
void orc_swap_axes(const float* in, int D, int H, int W, float* out) {
#pragma omp parallel for
  for (long p = 0; p < (long)H * W; ++p)
    for (int d = 0; d < D; ++d) out[(size_t)p * D + d] = in[(size_t)d * H * W + p];
}
void orc_swap_axes_back(const float* in, int H, int W, int D, float* out) {
#pragma omp parallel for
  for (long p = 0; p < (long)H * W; ++p)
    for (int d = 0; d < D; ++d) out[(size_t)d * H * W + p] = in[(size_t)p * D + d];
}

/* featextract.cpp:136-172.  res[y][x][d] = c[y][x+d][d] for x < W-d; everything
 * else holds c[0][0][0] (:151 -- the first element of whatever array was
 * passed in, a quirk that is part of the behaviour). */
void orc_get_right_cost(const float* c, int H, int W, int D, float* out) {
  fill_f32(out, (size_t)H * W * D, c[0]);
#pragma omp parallel for
  for (int y = 0; y < H; ++y)
    for (int d = 0; d < D; ++d)
      for (int x = 0; x < W - d; ++x)
        out[((size_t)y * W + x) * D + d] = c[((size_t)y * W + x + d) * D + d];
}
/* featextract.cpp:464-499.  res[y][x][d] = c[y][x-d][d] for x >= d; else c[0]. */
void orc_get_left_cost(const float* c, int H, int W, int D, float* out) {
  fill_f32(out, (size_t)H * W * D, c[0]);
#pragma omp parallel for
  for (int y = 0; y < H; ++y)
    for (int d = 0; d < D; ++d)
      for (int x = d; x < W; ++x)
        out[((size_t)y * W + x) * D + d] = c[((size_t)y * W + x - d) * D + d];
}

/* ------------------------------------------------------------------- AML --
 * featextract.cpp:415-462 (the 2-argument extract_likelihood).  Per row of D
 * costs: m = min (strict <, start value fill, :435-442); den = sequential fp32
 * sum of expf(-(c-m)^2/sigma) (:444-447); out = expf(-((c-m)^2/sigma))/den, or
 * 0 for every entry when m == fill (:449-453). */
void orc_aml(const float* cost, long n, int D, float sigma, float* out) {
#pragma omp parallel for
  for (long r = 0; r < n; ++r) {
    const float* c = cost + (size_t)r * D;
    float* o = out + (size_t)r * D;
    float m = ORC_FILL;
    for (int k = 0; k < D; ++k)
      if (c[k] < m) m = c[k];
    float den = 0.f;
    for (int k = 0; k < D; ++k) {
      const float num = c[k] - m;
      den += expf(-(num * num) / sigma);
    }
    for (int k = 0; k < D; ++k) {
      const float t = c[k] - m;
      o[k] = (m == ORC_FILL) ? 0.0f : expf(-((t * t) / sigma)) / den;
    }
  }
}

/* featextract.cpp:320-356 (the 2-argument extract_ratio, PKRN-style):
 * out = (m + e) / (c + e), or 0 for the whole row when m == fill. */
void orc_pkrn(const float* cost, long n, int D, float e, float* out) {
#pragma omp parallel for
  for (long r = 0; r < n; ++r) {
    const float* c = cost + (size_t)r * D;
    float* o = out + (size_t)r * D;
    float m = ORC_FILL;
    for (int k = 0; k < D; ++k)
      if (c[k] < m) m = c[k];
    for (int k = 0; k < D; ++k) o[k] = (m == ORC_FILL) ? 0.0f : (m + e) / (c[k] + e);
  }
}

/* ------------------------------------------------------------ soft-argmin --
 * gcnet_3dcnn.py:127-141: prob = softmax over D of logits[N][D][H][W];
 * disp = sum_d d * prob_d.  Accumulated in double here so the oracle itself is
 * not the dominant error; the parity tolerance (1e-3 px) is in the tests. */
void orc_soft_argmin(const float* logits, int N, int D, int H, int W, float* disp) {
  const size_t HW = (size_t)H * W;
#pragma omp parallel for
  for (long q = 0; q < (long)N * (long)HW; ++q) {
    const size_t n = (size_t)q / HW, p = (size_t)q % HW;
    const float* x = logits + n * D * HW + p;
    float mx = x[0];
    for (int d = 1; d < D; ++d)
      if (x[d * HW] > mx) mx = x[d * HW];
    double den = 0.0, num = 0.0;
    for (int d = 0; d < D; ++d) {
      const double e = exp((double)x[d * HW] - (double)mx);
      den += e;
      num += e * d;
    }
    disp[q] = (float)(num / den);
  }
}
