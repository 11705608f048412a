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

#include "smc_internal.h"  // first: fixes SMC_NS / the Philox round count of this build of the file
#include "smc_device.cuh"

namespace SMC_NS {
using namespace ::smc;  // shared helpers (smc_internal.h)

constexpr int NORMALS_BLOCK = 256;
constexpr int GROUPS_PER_THREAD = 4;  // row groups walked by one thread (24 rows f32 / 16 rows f64)

template <int VEC>
struct StoreVec;
template <>
struct StoreVec<1> {
  static __device__ __forceinline__ void st(float* p, const float (&z)[1][6], int r) { __stcs(p, z[0][r]); }
};
template <>
struct StoreVec<2> {
  static __device__ __forceinline__ void st(float* p, const float (&z)[2][6], int r) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(z[0][r], z[1][r]));
  }
};
template <>
struct StoreVec<4> {
  static __device__ __forceinline__ void st(float* p, const float (&z)[4][6], int r) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(z[0][r], z[1][r], z[2][r], z[3][r]));
  }
};

// grid.x covers column groups, grid.y covers runs of `groups_per_cta` row groups.  Full 6-row
// groups take the unguarded fast path (pointer bumped by `cols` per row); only the last group of a
// matrix whose row count is not a multiple of 6 checks rows.
#ifndef SMC_NORMALS_MIN_CTAS
#define SMC_NORMALS_MIN_CTAS 3  // measured best of {3,4,5,6} x run lengths: 5.9 TB/s (profiles/r1_codegen_variant_matrix.txt)
#endif
#ifndef SMC_NORMALS_MAX_GROUPS
#define SMC_NORMALS_MAX_GROUPS 42
#endif
template <int VEC>
__global__ void __launch_bounds__(NORMALS_BLOCK, SMC_NORMALS_MIN_CTAS)
    philox_normals_f32_kernel(float* __restrict__ out, int64_t rows, int64_t cols, PhiloxKeys key,
                              uint32_t k_lo, uint32_t k_hi, int groups_per_cta) {
  const int64_t col0 = (static_cast<int64_t>(blockIdx.x) * NORMALS_BLOCK + threadIdx.x) * VEC;
  if (col0 >= cols) return;
  const int64_t nq = (rows + 5) / 6, full = rows / 6;
  const int64_t q_begin = static_cast<int64_t>(blockIdx.y) * groups_per_cta;
  const int64_t q_end = min(q_begin + groups_per_cta, nq);
  float* dst = out + 6 * q_begin * cols + col0;
  const uint32_t c = static_cast<uint32_t>(col0);
  int64_t q = q_begin;
  for (; q < min(q_end, full); ++q) {
    float z[VEC][6];
#pragma unroll
    for (int v = 0; v < VEC; ++v) normals6_f32(c + v, static_cast<uint32_t>(q), k_lo, k_hi, key, z[v]);
#pragma unroll
    for (int rr = 0; rr < 6; ++rr) {
      StoreVec<VEC>::st(dst, z, rr);
      dst += cols;
    }
  }
  if (q < q_end) {  // ragged last group
    float z[VEC][6];
#pragma unroll
    for (int v = 0; v < VEC; ++v) normals6_f32(c + v, static_cast<uint32_t>(q), k_lo, k_hi, key, z[v]);
#pragma unroll
    for (int rr = 0; rr < 5; ++rr) {
      if (6 * q + rr < rows) StoreVec<VEC>::st(dst, z, rr);
      dst += cols;
    }
  }
}

// short float32 matrices (ROWS <= 3): one thread per block of G = 6 / ROWS adjacent columns
template <int ROWS>
__global__ void __launch_bounds__(NORMALS_BLOCK)
    philox_normals_f32_short_kernel(float* __restrict__ out, int64_t cols, PhiloxKeys key, uint32_t k_lo, uint32_t k_hi) {
  constexpr int G = 6 / ROWS;
  const int64_t g = static_cast<int64_t>(blockIdx.x) * NORMALS_BLOCK + threadIdx.x;
  const int64_t col0 = g * G;
  if (col0 >= cols) return;
  float z[6];
  normals6_f32(static_cast<uint32_t>(g), F32_SHORT_BIT, k_lo, k_hi, key, z);
#pragma unroll
  for (int u = 0; u < G; ++u) {
    if (col0 + u < cols) {
#pragma unroll
      for (int i = 0; i < ROWS; ++i) __stcs(out + i * cols + col0 + u, z[u * ROWS + i]);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(NORMALS_BLOCK)
    philox_normals_f64_kernel(double* __restrict__ out, int64_t rows, int64_t cols, PhiloxKeys key,
                              uint32_t k_lo, uint32_t k_hi) {
  const int64_t col0 = (static_cast<int64_t>(blockIdx.x) * NORMALS_BLOCK + threadIdx.x) * VEC;
  if (col0 >= cols) return;
  const int64_t nq = (rows + 3) >> 2;  // one block = four rows (two pairs) of a column
  for (int64_t q0 = static_cast<int64_t>(blockIdx.y) * GROUPS_PER_THREAD; q0 < nq;
       q0 += static_cast<int64_t>(gridDim.y) * GROUPS_PER_THREAD) {
#pragma unroll
    for (int g = 0; g < GROUPS_PER_THREAD; ++g) {
      const int64_t q = q0 + g;
      if (q >= nq) break;
      double z[VEC][4];
      const int pairs = 4 * q + 2 < rows ? 2 : 1;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (VEC == 1 || col0 + v < cols)
          normals4_f64(static_cast<uint32_t>(col0 + v), static_cast<uint32_t>(q), k_lo, k_hi, key, z[v], pairs);
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int64_t row = 4 * q + rr;
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

// argument-checked launch, in this build's namespace (Philox4x32-10 in the plain build, -7 under SMC_STREAM_P7)
static int launch_normals(void* out, int64_t rows, int64_t cols, int dtype, uint64_t seed, uint64_t matrix_index, void* stream) {
  const PhiloxKeys key = make_philox_keys(seed);
  const uint32_t k_lo = static_cast<uint32_t>(matrix_index);
  const uint32_t k_hi = static_cast<uint32_t>(matrix_index >> 32);
  cudaStream_t st = as_stream(stream);
  const bool aligned16 = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  // tuning knob: SMC_NORMALS_VEC=1|2|4 caps the float32 store width
  const char* vec_env = std::getenv("SMC_NORMALS_VEC");
  const int vec_cap = vec_env ? std::atoi(vec_env) : 4;
  const bool allow_vec = vec_cap > 1;
  if (dtype == SMC_F32 && rows <= 3) {  // short layout (oracle/philox.py)
    const int64_t groups = (cols + 6 / rows - 1) / (6 / rows);
    const unsigned grid = static_cast<unsigned>((groups + NORMALS_BLOCK - 1) / NORMALS_BLOCK);
    float* o = static_cast<float*>(out);
    if (rows == 1) philox_normals_f32_short_kernel<1><<<grid, NORMALS_BLOCK, 0, st>>>(o, cols, key, k_lo, k_hi);
    else if (rows == 2) philox_normals_f32_short_kernel<2><<<grid, NORMALS_BLOCK, 0, st>>>(o, cols, key, k_lo, k_hi);
    else philox_normals_f32_short_kernel<3><<<grid, NORMALS_BLOCK, 0, st>>>(o, cols, key, k_lo, k_hi);
  } else if (dtype == SMC_F32) {
    const bool a8 = (reinterpret_cast<uintptr_t>(out) & 7u) == 0;
    int v = 1;
    if (vec_cap >= 4 && aligned16 && cols % 4 == 0) v = 4;
    else if (vec_cap >= 2 && a8 && cols % 2 == 0) v = 2;
    const int64_t nq = (rows + 5) / 6;
    // enough CTAs along y to fill the machine for narrow matrices, long runs for wide ones
    const int64_t col_ctas = (cols / v + NORMALS_BLOCK - 1) / NORMALS_BLOCK;
    int groups_per_cta = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(nq, (col_ctas * nq + 4735) / 4736)));
    groups_per_cta = std::min(groups_per_cta, SMC_NORMALS_MAX_GROUPS);
    const int64_t gy = (nq + groups_per_cta - 1) / groups_per_cta;
    SMC_REQUIRE(gy <= 65535, "smc_philox_normals: too many rows (%lld) for one launch", (long long)rows);
    dim3 grid(static_cast<unsigned>(col_ctas), static_cast<unsigned>(gy));
    if (v == 4)
      philox_normals_f32_kernel<4><<<grid, NORMALS_BLOCK, 0, st>>>(static_cast<float*>(out), rows, cols, key, k_lo, k_hi, groups_per_cta);
    else if (v == 2)
      philox_normals_f32_kernel<2><<<grid, NORMALS_BLOCK, 0, st>>>(static_cast<float*>(out), rows, cols, key, k_lo, k_hi, groups_per_cta);
    else
      philox_normals_f32_kernel<1><<<grid, NORMALS_BLOCK, 0, st>>>(static_cast<float*>(out), rows, cols, key, k_lo, k_hi, groups_per_cta);
  } else {
    const bool vec = allow_vec && aligned16 && (cols % 2 == 0);
    const int v = vec ? 2 : 1;
    const int64_t nq = (rows + 3) / 4;
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

}  // namespace SMC_NS

using namespace SMC_NS;

#ifdef SMC_STREAM_P7
extern "C" int smc_p7_philox_normals(void* out, int64_t rows, int64_t cols, int dtype, uint64_t seed, uint64_t matrix_index,
                                     void* stream) {
  return launch_normals(out, rows, cols, dtype, seed, matrix_index, stream);
}
#else
extern "C" int smc_philox_normals_v(void* out, int64_t rows, int64_t cols, int dtype, uint64_t seed,
                                    uint64_t matrix_index, int stream_version, void* stream) {
  clear_error();
  SMC_REQUIRE(rows > 0 && cols > 0, "smc_philox_normals: invalid shape (%lld, %lld)", (long long)rows,
              (long long)cols);
  SMC_REQUIRE(out != nullptr, "smc_philox_normals: out is NULL");
  SMC_REQUIRE(dtype == SMC_F32 || dtype == SMC_F64, "smc_philox_normals: invalid dtype %d", dtype);
  SMC_REQUIRE(cols <= 0xffffffffLL, "smc_philox_normals: cols %lld exceeds the 32-bit path counter",
              (long long)cols);
  SMC_REQUIRE((matrix_index >> 63) == 0, "smc_philox_normals: matrix_index must be < 2^63");
  SMC_REQUIRE(stream_version == SMC_STREAM_PHILOX10 || stream_version == SMC_STREAM_PHILOX7, "smc_philox_normals: invalid stream_version %d", stream_version);
  if (stream_version == SMC_STREAM_PHILOX7) return smc_p7_philox_normals(out, rows, cols, dtype, seed, matrix_index, stream);
  return launch_normals(out, rows, cols, dtype, seed, matrix_index, stream);
}

extern "C" int smc_philox_normals(void* out, int64_t rows, int64_t cols, int dtype, uint64_t seed,
                                  uint64_t matrix_index, void* stream) {
  return smc_philox_normals_v(out, rows, cols, dtype, seed, matrix_index, SMC_STREAM_PHILOX10, stream);
}
#endif
