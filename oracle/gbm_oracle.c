/*
 * Oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of the SpectralMC batch-generation
 * path, used (a) as a second, independent checker of the NumPy restatement in oracle/gbm.py and
 * (b) as the multi-core CPU baseline timed by bench.py.  Never linked into or called from the
 * product (spectralmc_b200/).
 *
 * Each function cites the reference lines it follows (/root/reference/src/spectralmc/...).
 * The normal stream is the NEW counter-based stream specified in oracle/philox.py (the
 * reference's CuPy XORWOW draws, async_normals.py:214-215, are third-party and unpinned); this file
 * states the DEFAULT stream (Philox4x32-10) only — the opt-in seven-round stream is stated in philox.py.
 *
 * Build: make -C oracle   ->  oracle/_build/libgbm_oracle.so   (plain C, no OpenMP: callers
 * parallelise over path ranges with threads — ctypes releases the GIL — see oracle/cport.py)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u
#define F64_STREAM_BIT 0x80000000u

/* Philox4x32-10 (Salmon et al., SC'11); checked against the Random123 KAT vectors in tests */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define F32_REFINE_BIT 0x80000000u
#define F32_SHORT_BIT 0x40000000u /* row-group word of the short-matrix layout (float32, rows <= 3): see oracle/philox.py */
/* adjacent columns sharing one block */
static int short_group(int dtype, int64_t rows) { return (dtype == 0 && rows >= 1 && rows <= 3) ? (int)(6 / rows) : 1; }
static double uniform_21(uint32_t field) { return ((double)field + 0.5) * 0x1p-21; }
/* float64 stream: 43-bit radius field w0 << 11 | w1 >> 21, 21-bit angle field w1 & 0x1fffff of a pair's two words */
static double uniform_43(uint32_t w0, uint32_t w1) { return ((double)(((uint64_t)w0 << 11) | (w1 >> 21)) + 0.5) * 0x1p-43; }
static void box_muller(double u1, double u2, double* even, double* odd) {
  const double r = sqrt(-2.0 * log(u1)), theta = 2.0 * M_PI * (u2 - 0.5);
  *even = r * cos(theta);
  *odd = r * sin(theta);
}

/* one block of normals for column `col`, row group q: 6 values (float32 stream: 21-bit fields,
 * three pairs, rare refinement block) or 4 values (float64 stream: two pairs).  See oracle/philox.py. */
static void normals_block(int dtype, uint32_t col, uint32_t q, uint64_t seed, uint64_t k, double z[6]) {
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t ctr[4] = {col, q, (uint32_t)k, (uint32_t)(k >> 32) & 0x7fffffffu}, x[4];
  if (dtype == 0) {
    oracle_philox4x32_10(ctr, key, x);
    const uint32_t radius[3] = {x[0] >> 11, x[1] >> 11, x[2] >> 11};
    const uint32_t angle[3] = {((x[0] & 0x7ffu) << 10) | (x[3] >> 22), ((x[1] & 0x7ffu) << 10) | ((x[3] >> 12) & 0x3ffu),
                               ((x[2] & 0x7ffu) << 10) | ((x[3] >> 2) & 0x3ffu)};
    uint32_t y[4] = {0, 0, 0, 0};
    if (radius[0] == 0 || radius[1] == 0 || radius[2] == 0) {
      uint32_t c2[4] = {col, q | F32_REFINE_BIT, ctr[2], ctr[3]};
      oracle_philox4x32_10(c2, key, y);
    }
    for (int p = 0; p < 3; ++p) {
      const double u1 = radius[p] ? uniform_21(radius[p]) : ((double)(y[p] >> 9) + 0.5) * 0x1p-44;
      box_muller(u1, uniform_21(angle[p]), &z[2 * p], &z[2 * p + 1]);
    }
    for (int i = 0; i < 6; ++i) z[i] = (double)(float)z[i];
  } else {
    ctr[3] |= F64_STREAM_BIT;
    oracle_philox4x32_10(ctr, key, x);
    box_muller(uniform_43(x[0], x[1]), uniform_21(x[1] & 0x1fffffu), &z[0], &z[1]);
    box_muller(uniform_43(x[2], x[3]), uniform_21(x[3] & 0x1fffffu), &z[2], &z[3]);
  }
}

/* K1: the matrix_index-th (rows, cols) matrix; out is float (dtype 0) or double (dtype 1) */
void oracle_normals(void* out, int64_t rows, int64_t cols, int dtype, uint64_t seed, uint64_t matrix_index) {
  const int G = short_group(dtype, rows);
  if (G > 1) { /* short layout: element (i, j) = normal (j % G) * rows + i of block j / G */
    for (int64_t j = 0; j < cols; ++j) {
      double z[6];
      normals_block(0, (uint32_t)(j / G), F32_SHORT_BIT, seed, matrix_index, z);
      for (int64_t i = 0; i < rows; ++i) ((float*)out)[i * cols + j] = (float)z[(j % G) * rows + i];
    }
    return;
  }
  const int per = dtype == 0 ? 6 : 4;
  const int64_t nq = (rows + per - 1) / per;
  for (int64_t j = 0; j < cols; ++j) {
    for (int64_t q = 0; q < nq; ++q) {
      double z[6];
      normals_block(dtype, (uint32_t)j, (uint32_t)q, seed, matrix_index, z);
      for (int i = 0; i < per && q * per + i < rows; ++i) {
        if (dtype == 0) ((float*)out)[(q * per + i) * cols + j] = (float)z[i];
        else ((double*)out)[(q * per + i) * cols + j] = z[i];
      }
    }
  }
}

/* K2: the path kernel, gbm.py:241-257.  float64 arithmetic for both storage dtypes, narrowing
 * on store (compiled Numba typing). */
void oracle_paths_inplace(void* io, int64_t rows, int64_t cols, int dtype, double dt, double X0, double r,
                          double d, double v, int log_flag) {
  const double sqrt_dt = sqrt(dt); /* gbm.py:243 */
  for (int64_t j = 0; j < cols; ++j) {
    double X = X0; /* gbm.py:244 */
    const double drift = log_flag ? r - d - 0.5 * v * v : r - d; /* gbm.py:246,252 */
    for (int64_t i = 0; i < rows; ++i) {
      const double z = dtype == 0 ? (double)((float*)io)[i * cols + j] : ((double*)io)[i * cols + j];
      const double dW = z * sqrt_dt; /* gbm.py:248,254 */
      if (log_flag) X *= exp(drift * dt + v * dW); /* gbm.py:249 */
      else X = fabs(X + (drift * X * dt + v * X * dW)); /* gbm.py:255-256 */
      if (dtype == 0) ((float*)io)[i * cols + j] = (float)X; else ((double*)io)[i * cols + j] = X;
    }
  }
}

/* Terminal prices of global paths [path_begin, path_end) of one contract, streaming (no T x P
 * matrix): Philox normals -> K2 (gbm.py:241-257).  contract = X0,K,T,r,d,v.  Values are narrowed
 * to the engine dtype as the reference's store does and widened again.  Returns their sum. */
double oracle_terminal_range(const double* contract, int64_t T, int dtype, int log_flag, uint64_t seed,
                             uint64_t matrix_index, int64_t path_begin, int64_t path_end, double* terminal_out) {
  const double X0 = contract[0], Tm = contract[2], r = contract[3], d = contract[4], v = contract[5];
  const double dt = Tm / (double)T, sqrt_dt = sqrt(dt); /* gbm.py:411,243 */
  const double drift = log_flag ? r - d - 0.5 * v * v : r - d;
  const int per = dtype == 0 ? 6 : 4;
  const int G = short_group(dtype, T);
  double tsum = 0.0;
  for (int64_t j = path_begin; j < path_end; ++j) {
    double X = X0;
    for (int64_t q = 0; q * per < T; ++q) {
      double z[6];
      int first = 0;
      if (G > 1) { /* short layout: the path's T normals sit at (j % G) * T of block j / G */
        normals_block(0, (uint32_t)(j / G), F32_SHORT_BIT, seed, matrix_index, z);
        first = (int)(j % G) * (int)T;
      } else {
        normals_block(dtype, (uint32_t)j, (uint32_t)q, seed, matrix_index, z);
      }
      for (int i = 0; i < per && q * per + i < T; ++i) {
        const double dW = z[first + i] * sqrt_dt;
        if (log_flag) X *= exp(drift * dt + v * dW);
        else X = fabs(X + (drift * X * dt + v * X * dW));
      }
    }
    const double xt = dtype == 0 ? (double)(float)X : X;
    terminal_out[j - path_begin] = xt;
    tsum += xt;
  }
  return tsum;
}

/* [NORMALIZE of the terminal row, gbm.py:437-438] -> put payoff (gbm.py:473) -> CF =
 * FFT_n(mean_b mat) (gbm_trainer.py:814-817 by linearity; DFT evaluated directly in float64).
 * terminal = all P = N*B terminal prices; cf_out = N interleaved (re, im) doubles.  Engine-dtype
 * arithmetic (float32 when dtype == 0) is reproduced for forwards/df/payoff.
 * Returns the mean put price. */
double oracle_cf_from_terminal(const double* contract, int64_t N, int64_t B, int dtype, int normalize,
                               const double* terminal, double terminal_sum, double* cf_out) {
  const double X0 = contract[0], K = contract[1], Tm = contract[2], r = contract[3], d = contract[4];
  const int64_t P = N * B;
  /* times[-1] = T; forwards[-1], df[-1] in the engine dtype (gbm.py:429-431) */
  double fwd, df, Kd = K;
  if (dtype == 0) {
    const float tl = (float)Tm;
    fwd = (double)((float)X0 * expf((float)(r - d) * tl));
    df = (double)expf((float)(-r) * tl);
    Kd = (double)(float)K;
  } else {
    fwd = X0 * exp((r - d) * Tm);
    df = exp(-r * Tm);
  }
  double scale = 1.0;
  if (normalize) {
    const double mean = terminal_sum / (double)P;
    scale = dtype == 0 ? (double)((float)fwd / (float)mean) : fwd / mean;
  }
  double* colsum = (double*)calloc((size_t)N, sizeof(double));
  double psum = 0.0;
  for (int64_t b = 0; b < B; ++b)
    for (int64_t n = 0; n < N; ++n) {
      const double x = terminal[b * N + n];
      double put;
      if (dtype == 0) {
        const float xs = normalize ? (float)x * (float)scale : (float)x;
        const float diff = (float)Kd - xs;
        put = (double)((float)df * (diff > 0.f ? diff : 0.f));
      } else {
        const double xs = normalize ? x * scale : x;
        put = df * fmax(Kd - xs, 0.0);
      }
      colsum[n] += put;
      psum += put;
    }
  for (int64_t k = 0; k < N; ++k) {
    double re = 0.0, im = 0.0;
    for (int64_t n = 0; n < N; ++n) {
      const double ang = -2.0 * M_PI * (double)((k * n) % N) / (double)N;
      re += colsum[n] * cos(ang);
      im += colsum[n] * sin(ang);
    }
    cf_out[2 * k] = re / (double)B;
    cf_out[2 * k + 1] = im / (double)B;
  }
  free(colsum);
  return psum / (double)P;
}

