// SMC_CF_ROW_FFT — the literal formulation of the CF estimate: one radix-2 FFT per
// network_size row, fused with the batch-mean reduction, no HBM round trip:
//     cp.mean(cp.fft.fft(mat, axis=1), axis=0)      (gbm_trainer.py:409-412, 814-817)
//
// One warp owns a row at a time.  Lane l holds elements i = e*32 + l (e < E = N/32), so the row is
// read with fully coalesced 128-byte requests.  Decimation-in-frequency, natural order in /
// bit-reversed order out:
//   * stages with span >= 32 pair elements held by the SAME lane  -> register butterflies,
//     twiddles from a shared-memory table;
//   * the last five stages (span 16..1) pair lane l with lane l ^ span -> __shfl_xor_sync
//     butterflies, twiddles kept in registers (they depend on the lane only).
// The bit-reversed spectrum of every row is accumulated in float64 registers; the permutation
// is undone once, when the batch mean is written.  Reductions are fixed-order (warp -> CTA ->
// grid), so the result is bit-reproducible.
//
// By linearity the same estimate is FFT_n(mean_b mat) (SMC_CF_MEAN_THEN_FFT, smc_cf.cu), which
// does B-fold less arithmetic and stays on the HBM roofline; this kernel exists because the
// reference computes it row by row, and to measure the two against each other.
#include <algorithm>

#include "smc_device.cuh"
#include "smc_internal.h"

namespace smc {

constexpr int RF_BLOCK = 256;
constexpr int RF_WARPS = RF_BLOCK / 32;
constexpr int RF_MAX_CTAS = 1024;

template <typename Real>
struct Cx {
  Real re, im;
};

template <typename Real>
__device__ __forceinline__ Real shfl_xor(Real v, int mask) {
  return __shfl_xor_sync(0xffffffffu, v, mask);
}

// STORE_ROWS = true writes every row's spectrum to `spectra` [batches, N, 2] in natural order
// instead of accumulating the batch mean (smc_fft_rows, the ComputeFFT operator).
template <typename Real, int E, bool STORE_ROWS = false>
__global__ void __launch_bounds__(RF_BLOCK)
    rowfft_mean_kernel(const Real* __restrict__ mat, int64_t batches, int64_t rows_per_cta,
                       double* __restrict__ partial /* [ctas, N, 2] bit-reversed order */,
                       Real* __restrict__ spectra = nullptr) {
  constexpr int N = 32 * E;
  __shared__ Real tw_re[N / 2], tw_im[N / 2];
  __shared__ double red[RF_WARPS][32 * 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < N / 2; j += RF_BLOCK) {
    double s, c;
    sincospi(-2.0 * static_cast<double>(j) / static_cast<double>(N), &s, &c);  // exp(-2 pi i j / N)
    tw_re[j] = static_cast<Real>(c);
    tw_im[j] = static_cast<Real>(s);
  }
  // twiddles of the shuffle stages: span h in {16,8,4,2,1}, index (lane mod h) * N / (2h)
  Real sw_re[5], sw_im[5];
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    const int h = 16 >> t;
    double s, c;
    sincospi(-static_cast<double>(lane & (h - 1)) / static_cast<double>(h), &s, &c);
    sw_re[t] = static_cast<Real>(c);
    sw_im[t] = static_cast<Real>(s);
  }
  __syncthreads();

  double acc_re[E], acc_im[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc_re[e] = acc_im[e] = 0.0;

  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
  const int64_t row1 = min(row0 + rows_per_cta, batches);
  for (int64_t row = row0 + warp; row < row1; row += RF_WARPS) {
    Real xr[E], xi[E];
    const Real* src = mat + row * N + lane;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      xr[e] = __ldcs(src + 32 * e);
      xi[e] = Real(0);
    }
    // register stages: span = 32 * he, he = E/2 .. 1
#pragma unroll
    for (int he = E / 2; he >= 1; he >>= 1) {
      const int stride = E / (2 * he);  // N / (2 * span)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((e & he) == 0) {
          const int k = ((e & (he - 1)) * 32 + lane) * stride;
          const Real wr = tw_re[k], wi = tw_im[k];
          const Real ar = xr[e], ai = xi[e], br = xr[e + he], bi = xi[e + he];
          xr[e] = ar + br;
          xi[e] = ai + bi;
          const Real dr = ar - br, di = ai - bi;
          xr[e + he] = dr * wr - di * wi;
          xi[e + he] = dr * wi + di * wr;
        }
      }
    }
    // shuffle stages: span 16 .. 1, partner lane ^ span
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const int h = 16 >> t;
      const bool upper = (lane & h) != 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const Real pr = shfl_xor(xr[e], h), pi = shfl_xor(xi[e], h);
        if (!upper) {
          xr[e] += pr;
          xi[e] += pi;
        } else {
          const Real dr = pr - xr[e], di = pi - xi[e];  // a - b with a from the lower lane
          xr[e] = dr * sw_re[t] - di * sw_im[t];
          xi[e] = dr * sw_im[t] + di * sw_re[t];
        }
      }
    }
    if (STORE_ROWS) {
      constexpr int LOG2N = 5 + (E == 1 ? 0 : E == 2 ? 1 : E == 4 ? 2 : E == 8 ? 3 : 4);
      Real* dst = spectra + row * N * 2;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const unsigned k = __brev(static_cast<unsigned>(e * 32 + lane)) >> (32 - LOG2N);  // position holds X[bitrev]
        dst[2 * k] = xr[e];
        dst[2 * k + 1] = xi[e];
      }
      continue;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      acc_re[e] += static_cast<double>(xr[e]);
      acc_im[e] += static_cast<double>(xi[e]);
    }
  }
  if (STORE_ROWS) return;
  // CTA fold over warps (fixed order), one element slot e at a time
  double* out = partial + static_cast<int64_t>(blockIdx.x) * N * 2;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    red[warp][2 * lane] = acc_re[e];
    red[warp][2 * lane + 1] = acc_im[e];
    __syncthreads();
    if (warp == 0) {
      double sr = 0.0, si = 0.0;
      for (int w = 0; w < RF_WARPS; ++w) {
        sr += red[w][2 * lane];
        si += red[w][2 * lane + 1];
      }
      out[2 * (e * 32 + lane)] = sr;
      out[2 * (e * 32 + lane) + 1] = si;
    }
    __syncthreads();
  }
}

// fold the CTA partials, scale by 1/B, undo the bit reversal, narrow
template <typename Real>
__global__ void __launch_bounds__(RF_BLOCK)
    rowfft_finalize_kernel(const double* __restrict__ partial, int64_t ctas, int n, int log2n, double scale,
                           Real* __restrict__ out) {
  for (int j = threadIdx.x; j < n; j += RF_BLOCK) {
    double sr = 0.0, si = 0.0;
    for (int64_t c = 0; c < ctas; ++c) {
      sr += partial[(c * n + j) * 2];
      si += partial[(c * n + j) * 2 + 1];
    }
    const unsigned k = __brev(static_cast<unsigned>(j)) >> (32 - log2n);
    out[2 * k] = static_cast<Real>(sr * scale);
    out[2 * k + 1] = static_cast<Real>(si * scale);
  }
}

// any N <= 8192: one CTA per row, table-driven DFT in float64 shared memory (O(N^2) per row)
template <typename Real>
__global__ void __launch_bounds__(RF_BLOCK)
    dft_rows_kernel(const Real* __restrict__ mat, int64_t n, Real* __restrict__ spectra) {
  extern __shared__ double dsm[];
  double* x = dsm;
  double* twr = dsm + n;
  double* twi = twr + n;
  const int64_t row = blockIdx.x;
  for (int64_t j = threadIdx.x; j < n; j += RF_BLOCK) {
    double s, c;
    sincospi(-2.0 * static_cast<double>(j) / static_cast<double>(n), &s, &c);
    twr[j] = c;
    twi[j] = s;
    x[j] = static_cast<double>(mat[row * n + j]);
  }
  __syncthreads();
  for (int64_t k = threadIdx.x; k < n; k += RF_BLOCK) {
    double re = 0.0, im = 0.0;
    int64_t m = 0;
    for (int64_t j = 0; j < n; ++j) {
      re = fma(x[j], twr[m], re);
      im = fma(x[j], twi[m], im);
      m += k;
      if (m >= n) m -= n;
    }
    spectra[(row * n + k) * 2] = static_cast<Real>(re);
    spectra[(row * n + k) * 2 + 1] = static_cast<Real>(im);
  }
}

struct RowFftPlan {
  int64_t ctas, rows_per_cta;
};

static RowFftPlan rowfft_plan(int64_t batches) {
  RowFftPlan p;
  p.rows_per_cta = (batches + RF_MAX_CTAS - 1) / RF_MAX_CTAS;
  p.rows_per_cta = (p.rows_per_cta + RF_WARPS - 1) / RF_WARPS * RF_WARPS;
  p.ctas = (batches + p.rows_per_cta - 1) / p.rows_per_cta;
  return p;
}

bool rowfft_supported(int64_t n) { return n >= 32 && n <= 512 && (n & (n - 1)) == 0; }

size_t rowfft_workspace_bytes(int64_t batches, int64_t n) {
  return align_up(static_cast<size_t>(rowfft_plan(batches).ctas) * n * 2 * sizeof(double)) + 256;
}

template <typename Real>
static int rowfft_launch(const Real* mat, int64_t batches, int64_t n, Real* out, double* partial, cudaStream_t st) {
  const RowFftPlan p = rowfft_plan(batches);
  const unsigned grid = static_cast<unsigned>(p.ctas);
  switch (n) {
    case 32: rowfft_mean_kernel<Real, 1><<<grid, RF_BLOCK, 0, st>>>(mat, batches, p.rows_per_cta, partial); break;
    case 64: rowfft_mean_kernel<Real, 2><<<grid, RF_BLOCK, 0, st>>>(mat, batches, p.rows_per_cta, partial); break;
    case 128: rowfft_mean_kernel<Real, 4><<<grid, RF_BLOCK, 0, st>>>(mat, batches, p.rows_per_cta, partial); break;
    case 256: rowfft_mean_kernel<Real, 8><<<grid, RF_BLOCK, 0, st>>>(mat, batches, p.rows_per_cta, partial); break;
    case 512: rowfft_mean_kernel<Real, 16><<<grid, RF_BLOCK, 0, st>>>(mat, batches, p.rows_per_cta, partial); break;
    default: return set_error(SMC_EUNSUPPORTED, "SMC_CF_ROW_FFT needs a power-of-two network_size in [32, 512]");
  }
  SMC_LAUNCH_OK("rowfft_mean_kernel");
  int log2n = 0;
  while ((int64_t(1) << log2n) < n) ++log2n;
  rowfft_finalize_kernel<Real><<<1, RF_BLOCK, 0, st>>>(partial, p.ctas, static_cast<int>(n), log2n,
                                                      1.0 / static_cast<double>(batches), out);
  SMC_LAUNCH_OK("rowfft_finalize_kernel");
  return SMC_OK;
}

template <typename Real>
static int fft_rows_launch(const Real* mat, int64_t batches, int64_t n, Real* out, cudaStream_t st) {
  if (rowfft_supported(n)) {
    const RowFftPlan p = rowfft_plan(batches);
    const unsigned grid = static_cast<unsigned>(p.ctas);
    switch (n) {
      case 32: rowfft_mean_kernel<Real, 1, true><<<grid, RF_BLOCK, 0, st>>>(mat, batches, p.rows_per_cta, nullptr, out); break;
      case 64: rowfft_mean_kernel<Real, 2, true><<<grid, RF_BLOCK, 0, st>>>(mat, batches, p.rows_per_cta, nullptr, out); break;
      case 128: rowfft_mean_kernel<Real, 4, true><<<grid, RF_BLOCK, 0, st>>>(mat, batches, p.rows_per_cta, nullptr, out); break;
      case 256: rowfft_mean_kernel<Real, 8, true><<<grid, RF_BLOCK, 0, st>>>(mat, batches, p.rows_per_cta, nullptr, out); break;
      default: rowfft_mean_kernel<Real, 16, true><<<grid, RF_BLOCK, 0, st>>>(mat, batches, p.rows_per_cta, nullptr, out); break;
    }
    SMC_LAUNCH_OK("rowfft kernel (store)");
    return SMC_OK;
  }
  if (n > 8192) return set_error(SMC_EUNSUPPORTED, "smc_fft_rows: network_size %lld > 8192 is not supported", (long long)n);
  if (batches > 0x7fffffffLL) return set_error(SMC_EINVAL, "smc_fft_rows: too many rows");
  const size_t smem = static_cast<size_t>(n) * 24;
  if (smem > 48 * 1024)
    SMC_CUDA_OK(cudaFuncSetAttribute(dft_rows_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dft_rows_kernel<Real><<<static_cast<unsigned>(batches), RF_BLOCK, smem, st>>>(mat, n, out);
  SMC_LAUNCH_OK("dft_rows_kernel");
  return SMC_OK;
}

int fft_rows(const void* mat, int64_t batches, int64_t n, int dtype, void* out, cudaStream_t st) {
  if (dtype == SMC_F32) return fft_rows_launch<float>(static_cast<const float*>(mat), batches, n, static_cast<float*>(out), st);
  return fft_rows_launch<double>(static_cast<const double*>(mat), batches, n, static_cast<double*>(out), st);
}

int rowfft_mean(const void* mat, int64_t batches, int64_t n, int dtype, void* out, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  if (!rowfft_supported(n))
    return set_error(SMC_EUNSUPPORTED,
                     "smc_cf_fft_mean: SMC_CF_ROW_FFT needs a power-of-two network_size in [32, 512] (got %lld); "
                     "use SMC_CF_MEAN_THEN_FFT",
                     (long long)n);
  if (ws_bytes < rowfft_workspace_bytes(batches, n) - 256)
    return set_error(SMC_EWORKSPACE, "smc_cf_fft_mean: workspace too small");
  if (dtype == SMC_F32)
    return rowfft_launch<float>(static_cast<const float*>(mat), batches, n, static_cast<float*>(out),
                                static_cast<double*>(ws), st);
  return rowfft_launch<double>(static_cast<const double*>(mat), batches, n, static_cast<double*>(out),
                               static_cast<double*>(ws), st);
}

}  // namespace smc
