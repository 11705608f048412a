// Library plumbing and the small memory-bound kernels around the path:
//   payoff (gbm.py:467-474), host-price means (gbm.py:494-498), all-rows forward normalisation
//   (gbm.py:437-438) and the pipe-calibration microbenchmarks used by bench.py.
// All reductions are fixed-order (chunking depends on the problem size only).
#include <algorithm>
#include <cstring>
#include <string>

#include "smc_device.cuh"
#include "smc_internal.h"

namespace smc {

static thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

void clear_error() { g_last_error.clear(); }

int sm_count() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
      cached = n;
    else
      return 0;
  }
  return cached;
}

constexpr int BLK = 256;
constexpr int64_t CHUNK = 4096;  // elements per CTA in the streaming reductions

// ---- payoff -------------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(BLK)
    payoff_kernel(const Real* __restrict__ terminal, int64_t n, Real K, Real df, Real* __restrict__ put,
                  Real* __restrict__ call) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * BLK + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * BLK) {
    const Real x = terminal[i];
    const Real a = K - x, b = x - K;
    if (put) put[i] = df * (a > Real(0) ? a : Real(0));    // gbm.py:473
    if (call) call[i] = df * (b > Real(0) ? b : Real(0));  // gbm.py:474
  }
}

// ---- means of up to three vectors ---------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(BLK)
    means3_partial_kernel(const Real* __restrict__ a, const Real* __restrict__ b, const Real* __restrict__ c,
                          int64_t n, double* __restrict__ partial /* [chunks, 3] */) {
  __shared__ double sm[32];
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * CHUNK;
  const int64_t i1 = min(i0 + CHUNK, n);
  double sa = 0.0, sb = 0.0, sc = 0.0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += BLK) {
    if (a) sa += static_cast<double>(a[i]);
    if (b) sb += static_cast<double>(b[i]);
    if (c) sc += static_cast<double>(c[i]);
  }
  const double ta = block_sum(sa, sm);
  const double tb = block_sum(sb, sm);
  const double tc = block_sum(sc, sm);
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x + 0] = ta;
    partial[3 * blockIdx.x + 1] = tb;
    partial[3 * blockIdx.x + 2] = tc;
  }
}

__global__ void __launch_bounds__(BLK)
    means3_final_kernel(const double* __restrict__ partial, int64_t chunks, double inv_n, double* __restrict__ out3) {
  __shared__ double sm[32];
  double s[3] = {0.0, 0.0, 0.0};
  for (int64_t i = threadIdx.x; i < chunks; i += BLK)
    for (int k = 0; k < 3; ++k) s[k] += partial[3 * i + k];
  for (int k = 0; k < 3; ++k) {
    const double t = block_sum(s[k], sm);
    if (threadIdx.x == 0) out3[k] = t * inv_n;
  }
}

// ---- all-rows forward normalisation -----------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(BLK)
    row_partial_kernel(const Real* __restrict__ sims, int64_t cols, int64_t chunks, double* __restrict__ partial) {
  __shared__ double sm[32];
  const int64_t row = blockIdx.x / chunks, chunk = blockIdx.x - row * chunks;
  const int64_t j0 = chunk * CHUNK, j1 = min(j0 + CHUNK, cols);
  const Real* src = sims + row * cols;
  double s = 0.0;
  for (int64_t j = j0 + threadIdx.x; j < j1; j += BLK) s += static_cast<double>(src[j]);
  const double t = block_sum(s, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

template <typename Real>
__global__ void __launch_bounds__(BLK)
    row_ratio_kernel(const double* __restrict__ partial, int64_t chunks, int64_t cols,
                     const Real* __restrict__ forwards, Real* __restrict__ ratio) {
  __shared__ double sm[32];
  const int64_t row = blockIdx.x;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < chunks; i += BLK) s += partial[row * chunks + i];
  const double t = block_sum(s, sm);
  if (threadIdx.x == 0) {
    const Real mean = static_cast<Real>(t / static_cast<double>(cols));  // gbm.py:437
    ratio[row] = forwards[row] / mean;                                    // gbm.py:438
  }
}

template <typename Real>
__global__ void __launch_bounds__(BLK)
    row_scale_kernel(Real* __restrict__ sims, int64_t cols, int64_t chunks, const Real* __restrict__ ratio) {
  const int64_t row = blockIdx.x / chunks, chunk = blockIdx.x - row * chunks;
  const int64_t j0 = chunk * CHUNK, j1 = min(j0 + CHUNK, cols);
  Real* dst = sims + row * cols;
  const Real f = ratio[row];
  for (int64_t j = j0 + threadIdx.x; j < j1; j += BLK) dst[j] *= f;
}

// ---- pipe calibration ---------------------------------------------------------------------
__global__ void __launch_bounds__(BLK) calib_ffma_kernel(int64_t iters, float* sink) {
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 1.0f + 1e-3f * (threadIdx.x + k);
  const float m = 0.999999f, c = 1e-7f;
  for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 123.456f) sink[0] = s;
}

__global__ void __launch_bounds__(BLK) calib_mufu_kernel(int64_t iters, float* sink) {
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 1e-3f * (threadIdx.x + k);
  for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = mufu_ex2(a[k]) - 1.0f;  // 1 MUFU + 1 FADD
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 123.456f) sink[0] = s;
}

__global__ void __launch_bounds__(BLK) calib_dfma_kernel(int64_t iters, float* sink) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-3 * (threadIdx.x + k);
  const double m = 0.999999, c = 1e-7;
  for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 123.456) sink[0] = static_cast<float>(s);
}

__global__ void __launch_bounds__(BLK) calib_philox_kernel(int64_t iters, PhiloxKeys key, float* sink) {
  uint32_t acc = 0;
  const uint32_t col = blockIdx.x * BLK + threadIdx.x;
  for (int64_t i = 0; i < iters; ++i) {
    uint32_t x[4];
    philox4x32_10(col, static_cast<uint32_t>(i), 7u, 0u, key, x);
    acc ^= x[0] ^ x[1] ^ x[2] ^ x[3];
  }
  if (acc == 0x12345678u) sink[0] = 1.0f;
}

}  // namespace smc

using namespace smc;

extern "C" int smc_version(void) { return SMC_VERSION; }

extern "C" const char* smc_last_error(void) { return g_last_error.c_str(); }

extern "C" int smc_device_info(int* sms, int* major, int* minor) {
  clear_error();
  int dev = 0;
  SMC_CUDA_OK(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  SMC_CUDA_OK(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  SMC_CUDA_OK(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  SMC_CUDA_OK(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = a;
  if (major) *major = b;
  if (minor) *minor = c;
  return SMC_OK;
}

extern "C" int smc_payoff(const void* terminal, int64_t n, int dtype, double K, double df, void* put, void* call,
                          void* stream) {
  clear_error();
  SMC_REQUIRE(terminal != nullptr, "smc_payoff: terminal is NULL");
  SMC_REQUIRE(n > 0, "smc_payoff: n must be > 0");
  SMC_REQUIRE(dtype == SMC_F32 || dtype == SMC_F64, "smc_payoff: invalid dtype %d", dtype);
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n + BLK - 1) / BLK, 1 << 20));
  if (dtype == SMC_F32)
    payoff_kernel<float><<<grid, BLK, 0, as_stream(stream)>>>(static_cast<const float*>(terminal), n,
                                                              static_cast<float>(K), static_cast<float>(df),
                                                              static_cast<float*>(put), static_cast<float*>(call));
  else
    payoff_kernel<double><<<grid, BLK, 0, as_stream(stream)>>>(static_cast<const double*>(terminal), n, K, df,
                                                               static_cast<double*>(put), static_cast<double*>(call));
  SMC_LAUNCH_OK("payoff_kernel");
  return SMC_OK;
}

extern "C" size_t smc_means3_workspace_bytes(int64_t n) {
  if (n <= 0) return 0;
  return align_up(static_cast<size_t>((n + CHUNK - 1) / CHUNK) * 3 * sizeof(double));
}

extern "C" int smc_means3(const void* a, const void* b, const void* c, int64_t n, int dtype, double* out3, void* ws,
                          size_t ws_bytes, void* stream) {
  clear_error();
  SMC_REQUIRE(out3 && ws, "smc_means3: NULL pointer");
  SMC_REQUIRE(n > 0, "smc_means3: n must be > 0");
  SMC_REQUIRE(dtype == SMC_F32 || dtype == SMC_F64, "smc_means3: invalid dtype %d", dtype);
  if (ws_bytes < smc_means3_workspace_bytes(n)) return set_error(SMC_EWORKSPACE, "smc_means3: workspace too small");
  const int64_t chunks = (n + CHUNK - 1) / CHUNK;
  SMC_REQUIRE(chunks <= 0x7fffffffLL, "smc_means3: n too large");
  double* partial = static_cast<double*>(ws);
  cudaStream_t st = as_stream(stream);
  if (dtype == SMC_F32)
    means3_partial_kernel<float><<<static_cast<unsigned>(chunks), BLK, 0, st>>>(
        static_cast<const float*>(a), static_cast<const float*>(b), static_cast<const float*>(c), n, partial);
  else
    means3_partial_kernel<double><<<static_cast<unsigned>(chunks), BLK, 0, st>>>(
        static_cast<const double*>(a), static_cast<const double*>(b), static_cast<const double*>(c), n, partial);
  SMC_LAUNCH_OK("means3_partial_kernel");
  means3_final_kernel<<<1, BLK, 0, st>>>(partial, chunks, 1.0 / static_cast<double>(n), out3);
  SMC_LAUNCH_OK("means3_final_kernel");
  return SMC_OK;
}

extern "C" size_t smc_normalize_rows_workspace_bytes(int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  const int64_t chunks = (cols + CHUNK - 1) / CHUNK;
  return align_up(static_cast<size_t>(rows) * chunks * sizeof(double)) + align_up(static_cast<size_t>(rows) * 8);
}

extern "C" int smc_normalize_rows(void* sims, int64_t rows, int64_t cols, int dtype, const void* forwards, void* ws,
                                  size_t ws_bytes, void* stream) {
  clear_error();
  SMC_REQUIRE(sims && forwards && ws, "smc_normalize_rows: NULL pointer");
  SMC_REQUIRE(rows > 0 && cols > 0, "smc_normalize_rows: invalid shape (%lld, %lld)", (long long)rows, (long long)cols);
  SMC_REQUIRE(dtype == SMC_F32 || dtype == SMC_F64, "smc_normalize_rows: invalid dtype %d", dtype);
  if (ws_bytes < smc_normalize_rows_workspace_bytes(rows, cols))
    return set_error(SMC_EWORKSPACE, "smc_normalize_rows: workspace too small");
  const int64_t chunks = (cols + CHUNK - 1) / CHUNK;
  SMC_REQUIRE(rows * chunks <= 0x7fffffffLL, "smc_normalize_rows: matrix too large for one launch");
  double* partial = static_cast<double*>(ws);
  void* ratio = static_cast<char*>(ws) + align_up(static_cast<size_t>(rows) * chunks * sizeof(double));
  cudaStream_t st = as_stream(stream);
  const unsigned grid = static_cast<unsigned>(rows * chunks);
  if (dtype == SMC_F32) {
    row_partial_kernel<float><<<grid, BLK, 0, st>>>(static_cast<const float*>(sims), cols, chunks, partial);
    row_ratio_kernel<float><<<static_cast<unsigned>(rows), BLK, 0, st>>>(partial, chunks, cols,
                                                                         static_cast<const float*>(forwards),
                                                                         static_cast<float*>(ratio));
    row_scale_kernel<float><<<grid, BLK, 0, st>>>(static_cast<float*>(sims), cols, chunks, static_cast<const float*>(ratio));
  } else {
    row_partial_kernel<double><<<grid, BLK, 0, st>>>(static_cast<const double*>(sims), cols, chunks, partial);
    row_ratio_kernel<double><<<static_cast<unsigned>(rows), BLK, 0, st>>>(partial, chunks, cols,
                                                                          static_cast<const double*>(forwards),
                                                                          static_cast<double*>(ratio));
    row_scale_kernel<double><<<grid, BLK, 0, st>>>(static_cast<double*>(sims), cols, chunks, static_cast<const double*>(ratio));
  }
  SMC_LAUNCH_OK("normalize_rows kernels");
  return SMC_OK;
}

extern "C" int smc_pipe_calibrate(int kind, int64_t iters, double* ops, float* sink, void* stream) {
  clear_error();
  SMC_REQUIRE(kind >= 0 && kind <= 3, "smc_pipe_calibrate: invalid kind %d", kind);
  SMC_REQUIRE(iters > 0 && sink != nullptr && ops != nullptr, "smc_pipe_calibrate: bad argument");
  const int sms = sm_count();
  SMC_REQUIRE(sms > 0, "smc_pipe_calibrate: no CUDA device");
  const unsigned grid = static_cast<unsigned>(sms) * 8u;
  const double threads = static_cast<double>(grid) * BLK;
  cudaStream_t st = as_stream(stream);
  if (kind == 0) {
    calib_ffma_kernel<<<grid, BLK, 0, st>>>(iters, sink);
    *ops = threads * static_cast<double>(iters) * 8.0;  // FFMA lane-ops
  } else if (kind == 1) {
    calib_mufu_kernel<<<grid, BLK, 0, st>>>(iters, sink);
    *ops = threads * static_cast<double>(iters) * 8.0;  // MUFU lane-ops
  } else if (kind == 3) {
    calib_dfma_kernel<<<grid, BLK, 0, st>>>(iters, sink);
    *ops = threads * static_cast<double>(iters) * 8.0;  // DFMA lane-ops
  } else {
    calib_philox_kernel<<<grid, BLK, 0, st>>>(iters, make_philox_keys(42), sink);
    *ops = threads * static_cast<double>(iters);  // Philox blocks
  }
  SMC_LAUNCH_OK("calibration kernel");
  return SMC_OK;
}
