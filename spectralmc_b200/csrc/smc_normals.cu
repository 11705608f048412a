// K1 — materialised normal supply: smc_philox_normals.
// Replaces cp.random.default_rng(seed).standard_normal((rows, cols), dtype)
// (/root/reference/src/spectralmc/async_normals.py:214-215).
//
// HBM-write bound: sizeof(real) bytes per normal.  Each thread owns VEC adjacent columns and a
// run of row groups; it evaluates one Philox block per (column, row group), transposes the
// results in registers and issues one 16-byte store per row, so a warp writes 512 contiguous
// bytes per row (coalesced, vectorised).
#include <algorithm>
#include <cstdlib>

#include "smc_device.cuh"
#include "smc_internal.h"

namespace smc {

constexpr int NORMALS_BLOCK = 256;
constexpr int GROUPS_PER_THREAD = 4;  // row groups walked by one thread (24 rows f32 / 8 rows f64)

template <int VEC>
__global__ void __launch_bounds__(NORMALS_BLOCK, 4)
    philox_normals_f32_kernel(float* __restrict__ out, int64_t rows, int64_t cols, PhiloxKeys key,
                              uint32_t k_lo, uint32_t k_hi) {
  const int64_t col0 = (static_cast<int64_t>(blockIdx.x) * NORMALS_BLOCK + threadIdx.x) * VEC;
  if (col0 >= cols) return;
  const int64_t nq = (rows + 5) / 6;
  for (int64_t q0 = static_cast<int64_t>(blockIdx.y) * GROUPS_PER_THREAD; q0 < nq;
       q0 += static_cast<int64_t>(gridDim.y) * GROUPS_PER_THREAD) {
#pragma unroll 1
    for (int g = 0; g < GROUPS_PER_THREAD; ++g) {
      const int64_t q = q0 + g;
      if (q >= nq) break;
      float z[VEC][6];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (VEC == 1 || col0 + v < cols)
          normals6_f32(static_cast<uint32_t>(col0 + v), static_cast<uint32_t>(q), k_lo, k_hi, key, z[v]);
      }
#pragma unroll
      for (int rr = 0; rr < 6; ++rr) {
        const int64_t row = 6 * q + rr;
        if (row < rows) {
          float* dst = out + row * cols + col0;
          if (VEC == 4) {
            __stcs(reinterpret_cast<float4*>(dst), make_float4(z[0][rr], z[1][rr], z[2][rr], z[3][rr]));
          } else {
            __stcs(dst, z[0][rr]);
          }
        }
      }
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(NORMALS_BLOCK)
    philox_normals_f64_kernel(double* __restrict__ out, int64_t rows, int64_t cols, PhiloxKeys key,
                              uint32_t k_lo, uint32_t k_hi) {
  const int64_t col0 = (static_cast<int64_t>(blockIdx.x) * NORMALS_BLOCK + threadIdx.x) * VEC;
  if (col0 >= cols) return;
  const int64_t nq = (rows + 1) >> 1;
  for (int64_t q0 = static_cast<int64_t>(blockIdx.y) * GROUPS_PER_THREAD; q0 < nq;
       q0 += static_cast<int64_t>(gridDim.y) * GROUPS_PER_THREAD) {
#pragma unroll
    for (int g = 0; g < GROUPS_PER_THREAD; ++g) {
      const int64_t q = q0 + g;
      if (q >= nq) break;
      double z[VEC][2];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (VEC == 1 || col0 + v < cols)
          normals2_f64(static_cast<uint32_t>(col0 + v), static_cast<uint32_t>(q), k_lo, k_hi, key, z[v]);
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int64_t row = 2 * q + rr;
        if (row < rows) {
          double* dst = out + row * cols + col0;
          if (VEC == 2) {
            __stcs(reinterpret_cast<double2*>(dst), make_double2(z[0][rr], z[1][rr]));
          } else {
            __stcs(dst, z[0][rr]);
          }
        }
      }
    }
  }
}

}  // namespace smc

using namespace smc;

extern "C" int smc_philox_normals(void* out, int64_t rows, int64_t cols, int dtype, uint64_t seed,
                                  uint64_t matrix_index, void* stream) {
  clear_error();
  SMC_REQUIRE(rows > 0 && cols > 0, "smc_philox_normals: invalid shape (%lld, %lld)", (long long)rows,
              (long long)cols);
  SMC_REQUIRE(out != nullptr, "smc_philox_normals: out is NULL");
  SMC_REQUIRE(dtype == SMC_F32 || dtype == SMC_F64, "smc_philox_normals: invalid dtype %d", dtype);
  SMC_REQUIRE(cols <= 0xffffffffLL, "smc_philox_normals: cols %lld exceeds the 32-bit path counter",
              (long long)cols);
  SMC_REQUIRE((matrix_index >> 63) == 0, "smc_philox_normals: matrix_index must be < 2^63");
  const PhiloxKeys key = make_philox_keys(seed);
  const uint32_t k_lo = static_cast<uint32_t>(matrix_index);
  const uint32_t k_hi = static_cast<uint32_t>(matrix_index >> 32);
  cudaStream_t st = as_stream(stream);
  const bool aligned16 = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  const char* force_scalar = std::getenv("SMC_NORMALS_SCALAR");  // tuning knob
  const bool allow_vec = !(force_scalar && force_scalar[0] == '1');
  if (dtype == SMC_F32) {
    const bool vec = allow_vec && aligned16 && (cols % 4 == 0);
    const int v = vec ? 4 : 1;
    const int64_t nq = (rows + 5) / 6;
    dim3 grid(static_cast<unsigned>((cols / v + (cols % v != 0) + NORMALS_BLOCK - 1) / NORMALS_BLOCK),
              static_cast<unsigned>(std::min<int64_t>((nq + GROUPS_PER_THREAD - 1) / GROUPS_PER_THREAD, 65535)));
    if (vec)
      philox_normals_f32_kernel<4><<<grid, NORMALS_BLOCK, 0, st>>>(static_cast<float*>(out), rows, cols, key, k_lo, k_hi);
    else
      philox_normals_f32_kernel<1><<<grid, NORMALS_BLOCK, 0, st>>>(static_cast<float*>(out), rows, cols, key, k_lo, k_hi);
  } else {
    const bool vec = allow_vec && aligned16 && (cols % 2 == 0);
    const int v = vec ? 2 : 1;
    const int64_t nq = (rows + 1) / 2;
    dim3 grid(static_cast<unsigned>((cols / v + (cols % v != 0) + NORMALS_BLOCK - 1) / NORMALS_BLOCK),
              static_cast<unsigned>(std::min<int64_t>((nq + GROUPS_PER_THREAD - 1) / GROUPS_PER_THREAD, 65535)));
    if (vec)
      philox_normals_f64_kernel<2><<<grid, NORMALS_BLOCK, 0, st>>>(static_cast<double*>(out), rows, cols, key, k_lo, k_hi);
    else
      philox_normals_f64_kernel<1><<<grid, NORMALS_BLOCK, 0, st>>>(static_cast<double*>(out), rows, cols, key, k_lo, k_hi);
  }
  SMC_LAUNCH_OK("philox_normals_kernel");
  return SMC_OK;
}
