// Statistical diagnostics of the float32 normal stream, computed on the device without materialising a
// matrix: the stream is this library's own specification (oracle/philox.py — the reference's CuPy XORWOW bits,
// /root/reference/src/spectralmc/async_normals.py:214-215, are third-party and unpinned), so the library also
// ships the means to audit it at sample sizes no HBM matrix reaches (2^33 draws and more in a second).
//   smc_diag_stream_fields_f32   histogram of the 21-bit radius and angle fields (chi-square over all 2^21 cells),
//                                tail counts of the resulting normals beyond 4 / 5 / 5.5 / 6 sigma (refined entries
//                                included) and their first four power sums
//   smc_diag_stream_lags_f32     sums of products z[i, j] z[i - lag, j] for lags 1..6 along a path (within and across
//                                the 6-row blocks) and z[i, j] z[i, j + 1] across adjacent columns
// Test infrastructure in the sense that only tests/test_gpu_stream_battery.py calls it; it reuses the very device
// functions the path kernels draw their normals with (smc_device.cuh), which is the point.
#include "smc_internal.h"  // first: fixes SMC_NS / the Philox round count of this build of the file
#include "smc_device.cuh"

namespace SMC_NS {
using namespace ::smc;  // shared helpers (smc_internal.h); everything that draws normals lives in SMC_NS

constexpr int DIAG_BLOCK = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// block b of the audit <-> counter (column = b mod cols, row group = b div cols) of matrix (k_lo, k_hi)
__global__ void __launch_bounds__(DIAG_BLOCK)
    diag_fields_kernel(PhiloxKeys key, uint32_t k_lo, uint32_t k_hi, unsigned long long n_blocks, uint32_t cols,
                       unsigned* __restrict__ radius_hist, unsigned* __restrict__ angle_hist,
                       unsigned long long* __restrict__ tails /* [4] */, double* __restrict__ power_sums /* [4] */) {
  unsigned long long t4 = 0, t5 = 0, t55 = 0, t6 = 0;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * DIAG_BLOCK;
  for (unsigned long long b = static_cast<unsigned long long>(blockIdx.x) * DIAG_BLOCK + threadIdx.x; b < n_blocks; b += stride) {
    const uint32_t col = static_cast<uint32_t>(b % cols), q = static_cast<uint32_t>(b / cols);
    uint32_t x[4];
    philox4x32_10(col, q, k_lo, k_hi, key, x);
    // the six 21-bit fields, exactly as oracle/philox.py cuts them
    atomicAdd(radius_hist + (x[0] >> 11), 1u);
    atomicAdd(radius_hist + (x[1] >> 11), 1u);
    atomicAdd(radius_hist + (x[2] >> 11), 1u);
    atomicAdd(angle_hist + (((x[0] & 0x7ffu) << 10) | (x[3] >> 22)), 1u);
    atomicAdd(angle_hist + (((x[1] & 0x7ffu) << 10) | ((x[3] >> 12) & 0x3ffu)), 1u);
    atomicAdd(angle_hist + (((x[2] & 0x7ffu) << 10) | ((x[3] >> 2) & 0x3ffu)), 1u);
    float z[6];
    normals6_f32(col, q, k_lo, k_hi, key, z);  // the stream as specified (refinement of zero radius fields included)
    float loc1 = 0.f, loc2 = 0.f, loc3 = 0.f, loc4 = 0.f;
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const float a = fabsf(z[u]);
      t4 += a > 4.0f;
      t5 += a > 5.0f;
      t55 += a > 5.5f;
      t6 += a > 6.0f;
      const float z2 = z[u] * z[u];
      loc1 += z[u];
      loc2 += z2;
      loc3 += z2 * z[u];
      loc4 += z2 * z2;
    }
    s1 += loc1;
    s2 += loc2;
    s3 += loc3;
    s4 += loc4;
  }
  t4 = warp_sum(t4); t5 = warp_sum(t5); t55 = warp_sum(t55); t6 = warp_sum(t6);
  s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3); s4 = warp_sum(s4);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(tails + 0, t4); atomicAdd(tails + 1, t5); atomicAdd(tails + 2, t55); atomicAdd(tails + 3, t6);
    atomicAdd(power_sums + 0, s1); atomicAdd(power_sums + 1, s2); atomicAdd(power_sums + 2, s3); atomicAdd(power_sums + 3, s4);
  }
}

// thread = column; walks `rows` (a multiple of 6) rows.  sums[lag - 1] += z[i] z[i - lag] for lag 1..6 (i >= lag),
// sums[6] += z[i, j] z[i, j + 1] for columns j, j + 1 in one warp.
__global__ void __launch_bounds__(DIAG_BLOCK)
    diag_lags_kernel(PhiloxKeys key, uint32_t k_lo, uint32_t k_hi, uint32_t cols, uint32_t row_groups, double* __restrict__ sums /* [7] */) {
  const uint32_t col = blockIdx.x * DIAG_BLOCK + threadIdx.x;
  const bool live = col < cols;
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  float h[6] = {0, 0, 0, 0, 0, 0};  // h[k] = z[i - 1 - k]
  for (uint32_t q = 0; q < row_groups; ++q) {
    float z[6] = {0, 0, 0, 0, 0, 0};
    if (live) normals6_f32(col, q, k_lo, k_hi, key, z);
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const float right = __shfl_down_sync(0xffffffffu, z[u], 1);
      if (live && (threadIdx.x & 31) != 31 && col + 1 < cols) s[6] += static_cast<double>(z[u] * right);
      const uint32_t i = q * 6 + u;
#pragma unroll
      for (int k = 0; k < 6; ++k)
        if (i > static_cast<uint32_t>(k)) s[k] += static_cast<double>(z[u] * h[k]);
#pragma unroll
      for (int k = 5; k > 0; --k) h[k] = h[k - 1];
      h[0] = z[u];
    }
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const double t = warp_sum(live ? s[k] : 0.0);
    if ((threadIdx.x & 31) == 0) atomicAdd(sums + k, t);
  }
}

// the launches, in this build's namespace (Philox4x32-10 in the plain build, -7 under SMC_STREAM_P7)
static int launch_fields(uint64_t seed, uint64_t matrix_index, uint64_t n_blocks, uint32_t cols, uint32_t* radius_hist,
                         uint32_t* angle_hist, uint64_t* tails4, double* power_sums4, void* stream) {
  const int sms = sm_count();
  SMC_REQUIRE(sms > 0, "smc_diag_stream_fields_f32: no CUDA device");
  diag_fields_kernel<<<static_cast<unsigned>(sms) * 8u, DIAG_BLOCK, 0, as_stream(stream)>>>(
      make_philox_keys(seed), static_cast<uint32_t>(matrix_index), static_cast<uint32_t>(matrix_index >> 32), n_blocks, cols,
      radius_hist, angle_hist, reinterpret_cast<unsigned long long*>(tails4), power_sums4);
  SMC_LAUNCH_OK("diag_fields_kernel");
  return SMC_OK;
}

static int launch_lags(uint64_t seed, uint64_t matrix_index, uint32_t cols, uint32_t rows, double* sums7, void* stream) {
  diag_lags_kernel<<<(cols + DIAG_BLOCK - 1) / DIAG_BLOCK, DIAG_BLOCK, 0, as_stream(stream)>>>(
      make_philox_keys(seed), static_cast<uint32_t>(matrix_index), static_cast<uint32_t>(matrix_index >> 32), cols, rows / 6, sums7);
  SMC_LAUNCH_OK("diag_lags_kernel");
  return SMC_OK;
}

}  // namespace SMC_NS

using namespace SMC_NS;

#ifdef SMC_STREAM_P7
extern "C" int smc_p7_diag_stream_fields_f32(uint64_t seed, uint64_t matrix_index, uint64_t n_blocks, uint32_t cols,
                                             uint32_t* radius_hist, uint32_t* angle_hist, uint64_t* tails4, double* power_sums4,
                                             void* stream) {
  return launch_fields(seed, matrix_index, n_blocks, cols, radius_hist, angle_hist, tails4, power_sums4, stream);
}
extern "C" int smc_p7_diag_stream_lags_f32(uint64_t seed, uint64_t matrix_index, uint32_t cols, uint32_t rows, double* sums7,
                                           void* stream) {
  return launch_lags(seed, matrix_index, cols, rows, sums7, stream);
}
#else
extern "C" int smc_diag_stream_fields_f32(uint64_t seed, uint64_t matrix_index, uint64_t n_blocks, uint32_t cols,
                                          uint32_t* radius_hist, uint32_t* angle_hist, uint64_t* tails4, double* power_sums4,
                                          int stream_version, void* stream) {
  clear_error();
  SMC_REQUIRE(radius_hist && angle_hist && tails4 && power_sums4, "smc_diag_stream_fields_f32: NULL pointer");
  SMC_REQUIRE(n_blocks > 0 && cols > 0 && n_blocks / cols < 0x40000000ull, "smc_diag_stream_fields_f32: bad shape");
  SMC_REQUIRE((matrix_index >> 63) == 0, "smc_diag_stream_fields_f32: matrix_index must be < 2^63");
  SMC_REQUIRE(stream_version == SMC_STREAM_PHILOX10 || stream_version == SMC_STREAM_PHILOX7, "smc_diag_stream_fields_f32: invalid stream_version %d", stream_version);
  if (stream_version == SMC_STREAM_PHILOX7)
    return smc_p7_diag_stream_fields_f32(seed, matrix_index, n_blocks, cols, radius_hist, angle_hist, tails4, power_sums4, stream);
  return launch_fields(seed, matrix_index, n_blocks, cols, radius_hist, angle_hist, tails4, power_sums4, stream);
}

extern "C" int smc_diag_stream_lags_f32(uint64_t seed, uint64_t matrix_index, uint32_t cols, uint32_t rows, double* sums7,
                                        int stream_version, void* stream) {
  clear_error();
  SMC_REQUIRE(sums7 != nullptr, "smc_diag_stream_lags_f32: NULL pointer");
  SMC_REQUIRE(cols > 0 && rows >= 12 && rows % 6 == 0, "smc_diag_stream_lags_f32: rows must be a multiple of 6, at least 12");
  SMC_REQUIRE((matrix_index >> 63) == 0, "smc_diag_stream_lags_f32: matrix_index must be < 2^63");
  SMC_REQUIRE(stream_version == SMC_STREAM_PHILOX10 || stream_version == SMC_STREAM_PHILOX7, "smc_diag_stream_lags_f32: invalid stream_version %d", stream_version);
  if (stream_version == SMC_STREAM_PHILOX7) return smc_p7_diag_stream_lags_f32(seed, matrix_index, cols, rows, sums7, stream);
  return launch_lags(seed, matrix_index, cols, rows, sums7, stream);
}
#endif
