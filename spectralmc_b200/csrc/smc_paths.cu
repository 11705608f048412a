// K2 — the GBM/Black-Scholes path kernel on MATERIALISED normals.
// Replaces the Numba kernel SimulateBlackScholes
// (/root/reference/src/spectralmc/gbm.py:224-257; launch at gbm.py:413-426).
//
// HBM bound: the in-place form reads and writes sizeof(real) per path-step; the terminal-only
// form reads sizeof(real) per path-step.  One thread owns VEC adjacent columns (16-byte
// accesses) and walks the rows in groups of UNROLL: all loads of a group are issued before the
// first dependent arithmetic, so each thread keeps UNROLL x 16 bytes in flight.
//
// Numerics follow the reference's compiled data flow: float64 arithmetic for both storage
// dtypes, narrowing only on store.  For float32 storage the log-Euler running
// value is carried as a float64 log2-return s2 = sum_j (a2 + b2 z_j); each stored value is
// X0 * 2^s2 with the fraction exponentiated by one MUFU.EX2 (relative error 2^-22, i.e. inside
// the float32 store rounding the reference applies anyway), which keeps the kernel on the HBM
// roofline instead of the FP64 pipe.  float64 storage reproduces the reference's running
// product X *= exp(drift dt + v dW) with libdevice exp.
#include <algorithm>

#include "smc_device.cuh"
#include "smc_internal.h"

namespace smc {

#ifndef SMC_PATH_UNROLL
#define SMC_PATH_UNROLL 8
#endif
#ifndef SMC_PATH_STREAM_STORE
#define SMC_PATH_STREAM_STORE 1  // st.global.cs: +1 % on the in-place kernel (6.1 vs 6.05 TB/s)
#endif
constexpr int PATH_UNROLL = SMC_PATH_UNROLL;

struct PathConsts {
  double X0;
  double sqrt_dt;   // gbm.py:243
  double drift_dt;  // (r - d - v^2/2) dt   (log-Euler)  |  (r - d) dt (simple Euler)
  double v;         // gbm.py:249/255
  double a2, b2;    // log2(e) * drift_dt, log2(e) * v * sqrt_dt
};

inline PathConsts make_path_consts(double dt, double X0, double r, double d, double v, int scheme) {
  PathConsts c;
  c.X0 = X0;
  c.sqrt_dt = std::sqrt(dt);
  const double drift = scheme == SMC_LOG_EULER ? (r - d - 0.5 * v * v) : (r - d);
  c.drift_dt = drift * dt;
  c.v = v;
  const double log2e = 1.4426950408889634074;
  c.a2 = c.drift_dt * log2e;
  c.b2 = v * c.sqrt_dt * log2e;
  return c;
}

// X0 * 2^s2 narrowed to float32: integer part through the exponent field, fraction via MUFU.EX2
__device__ __forceinline__ float exp2_scaled_f32(double s2, float X0f) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: adds round-to-nearest-integer
  const double t = s2 + magic;
  int n = __double2loint(t);
  const float frac = static_cast<float>(s2 - (t - magic));  // in [-0.5, 0.5]
  n = max(-252, min(252, n));
  const int n1 = n >> 1, n2 = n - n1;
  const float e = X0f * mufu_ex2(frac);
  return e * __int_as_float((n1 + 127) << 23) * __int_as_float((n2 + 127) << 23);
}

template <typename Real, int VEC>
struct Pack;
template <>
struct Pack<float, 4> {
  using type = float4;
};
template <>
struct Pack<float, 1> {
  using type = float;
};
template <>
struct Pack<double, 2> {
  using type = double2;
};
template <>
struct Pack<double, 1> {
  using type = double;
};

// __launch_bounds__(1024): the CTA size is the caller's threads_per_block (32..1024, gbm.py:71)
template <typename Real, int VEC, int SCHEME, bool STORE_PATHS>
__global__ void __launch_bounds__(1024) gbm_paths_kernel(Real* __restrict__ io, const Real* __restrict__ normals,
                                 Real* __restrict__ terminal, int64_t rows, int64_t cols,
                                 PathConsts k) {
  using P = typename Pack<Real, VEC>::type;
  const int64_t col0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * VEC;
  if (col0 >= cols) return;  // gbm.py:242
  const Real* src = STORE_PATHS ? io : normals;

  constexpr bool F32 = sizeof(Real) == 4;
  constexpr bool LOG = SCHEME == SMC_LOG_EULER;
  // running state per owned column: log2-return (float32 log-Euler, and float64 terminal-only
  // log-Euler in natural-log units) or the price itself
  double state[VEC];
  constexpr bool LOGSUM = LOG && (F32 || !STORE_PATHS);
#pragma unroll
  for (int v = 0; v < VEC; ++v) state[v] = LOGSUM ? 0.0 : k.X0;
  const float X0f = static_cast<float>(k.X0);

  for (int64_t i0 = 0; i0 < rows; i0 += PATH_UNROLL) {
    P buf[PATH_UNROLL];
#pragma unroll
    for (int u = 0; u < PATH_UNROLL; ++u)
      if (i0 + u < rows) buf[u] = __ldcs(reinterpret_cast<const P*>(src + (i0 + u) * cols + col0));
#pragma unroll
    for (int u = 0; u < PATH_UNROLL; ++u) {
      if (i0 + u < rows) {
        Real* zv = reinterpret_cast<Real*>(&buf[u]);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const double z = static_cast<double>(zv[v]);
          Real outv = Real(0);
          if (LOG) {
            if (F32) {
              state[v] = fma(k.b2, z, state[v]) + k.a2;
              if (STORE_PATHS) outv = static_cast<Real>(exp2_scaled_f32(state[v], X0f));
            } else if (STORE_PATHS) {
              const double dW = z * k.sqrt_dt;                 // gbm.py:248
              state[v] *= exp(fma(k.v, dW, k.drift_dt));       // gbm.py:249
              outv = static_cast<Real>(state[v]);
            } else {
              state[v] += fma(k.v, z * k.sqrt_dt, k.drift_dt);
            }
          } else {
            const double dW = z * k.sqrt_dt;                   // gbm.py:254
            const double X = state[v];
            state[v] = fabs(X + (k.drift_dt * X + k.v * X * dW));  // gbm.py:255-256
            outv = static_cast<Real>(state[v]);
          }
          if (STORE_PATHS) zv[v] = outv;
        }
        if (STORE_PATHS) {
          if (SMC_PATH_STREAM_STORE) __stcs(reinterpret_cast<P*>(io + (i0 + u) * cols + col0), buf[u]);
          else *reinterpret_cast<P*>(io + (i0 + u) * cols + col0) = buf[u];
        }
      }
    }
  }
  if (!STORE_PATHS) {
    P t;
    Real* tv = reinterpret_cast<Real*>(&t);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (LOGSUM)
        tv[v] = F32 ? static_cast<Real>(exp2_scaled_f32(state[v], X0f))
                    : static_cast<Real>(k.X0 * exp(state[v]));
      else
        tv[v] = static_cast<Real>(state[v]);
    }
    *reinterpret_cast<P*>(terminal + col0) = t;
  }
}

template <typename Real, int VEC, bool STORE_PATHS>
static void launch_paths(Real* io, const Real* normals, Real* terminal, int64_t rows, int64_t cols,
                         const PathConsts& k, int scheme, int tpb, cudaStream_t st) {
  const int64_t threads = (cols + VEC - 1) / VEC;
  const unsigned grid = static_cast<unsigned>((threads + tpb - 1) / tpb);
  if (scheme == SMC_LOG_EULER)
    gbm_paths_kernel<Real, VEC, SMC_LOG_EULER, STORE_PATHS><<<grid, tpb, 0, st>>>(io, normals, terminal, rows, cols, k);
  else
    gbm_paths_kernel<Real, VEC, SMC_SIMPLE_EULER, STORE_PATHS><<<grid, tpb, 0, st>>>(io, normals, terminal, rows, cols, k);
}

template <bool STORE_PATHS>
static int dispatch_paths(void* io, const void* normals, void* terminal, int64_t rows, int64_t cols,
                          int dtype, const PathConsts& k, int scheme, int tpb, cudaStream_t st) {
  const void* base = STORE_PATHS ? io : normals;
  bool aligned16 = (reinterpret_cast<uintptr_t>(base) & 15u) == 0;
  if (!STORE_PATHS) aligned16 = aligned16 && (reinterpret_cast<uintptr_t>(terminal) & 15u) == 0;
  if (dtype == SMC_F32) {
    if (aligned16 && cols % 4 == 0)
      launch_paths<float, 4, STORE_PATHS>(static_cast<float*>(io), static_cast<const float*>(normals),
                                          static_cast<float*>(terminal), rows, cols, k, scheme, tpb, st);
    else
      launch_paths<float, 1, STORE_PATHS>(static_cast<float*>(io), static_cast<const float*>(normals),
                                          static_cast<float*>(terminal), rows, cols, k, scheme, tpb, st);
  } else {
    if (aligned16 && cols % 2 == 0)
      launch_paths<double, 2, STORE_PATHS>(static_cast<double*>(io), static_cast<const double*>(normals),
                                           static_cast<double*>(terminal), rows, cols, k, scheme, tpb, st);
    else
      launch_paths<double, 1, STORE_PATHS>(static_cast<double*>(io), static_cast<const double*>(normals),
                                           static_cast<double*>(terminal), rows, cols, k, scheme, tpb, st);
  }
  SMC_LAUNCH_OK("gbm_paths_kernel");
  return SMC_OK;
}

}  // namespace smc

using namespace smc;

static int check_common(const char* fn, int64_t rows, int64_t cols, int dtype, int scheme) {
  SMC_REQUIRE(rows > 0 && cols > 0, "%s: invalid shape (%lld, %lld)", fn, (long long)rows, (long long)cols);
  SMC_REQUIRE(dtype == SMC_F32 || dtype == SMC_F64, "%s: invalid dtype %d", fn, dtype);
  SMC_REQUIRE(scheme == SMC_LOG_EULER || scheme == SMC_SIMPLE_EULER, "%s: invalid scheme %d", fn, scheme);
  return SMC_OK;
}

extern "C" int smc_gbm_paths_inplace(void* io, int64_t rows, int64_t cols, int dtype, double dt, double X0,
                                     double r, double d, double v, int scheme, int threads_per_block,
                                     void* stream) {
  clear_error();
  SMC_REQUIRE(io != nullptr, "smc_gbm_paths_inplace: io is NULL");
  if (int e = check_common("smc_gbm_paths_inplace", rows, cols, dtype, scheme)) return e;
  const int t = threads_per_block;
  SMC_REQUIRE(t == 32 || t == 64 || t == 128 || t == 256 || t == 512 || t == 1024,
              "smc_gbm_paths_inplace: threads_per_block %d not in {32,64,128,256,512,1024}", t);
  SMC_REQUIRE(dt >= 0.0, "smc_gbm_paths_inplace: dt must be >= 0");
  const PathConsts k = make_path_consts(dt, X0, r, d, v, scheme);
  return dispatch_paths<true>(io, nullptr, nullptr, rows, cols, dtype, k, scheme, t, as_stream(stream));
}

extern "C" int smc_gbm_terminal_from_normals(const void* normals, int64_t rows, int64_t cols, int dtype,
                                             double dt, double X0, double r, double d, double v, int scheme,
                                             void* terminal, void* stream) {
  clear_error();
  SMC_REQUIRE(normals != nullptr && terminal != nullptr, "smc_gbm_terminal_from_normals: NULL pointer");
  if (int e = check_common("smc_gbm_terminal_from_normals", rows, cols, dtype, scheme)) return e;
  SMC_REQUIRE(dt >= 0.0, "smc_gbm_terminal_from_normals: dt must be >= 0");
  const PathConsts k = make_path_consts(dt, X0, r, d, v, scheme);
  return dispatch_paths<false>(nullptr, normals, terminal, rows, cols, dtype, k, scheme, 256, as_stream(stream));
}
