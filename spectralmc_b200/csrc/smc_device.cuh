// Device-side building blocks: Philox4x32-10, exact uniforms, Box–Muller (MUFU for float32,
// libdevice for float64) and the GBM step algebra.  The stream these functions define is
// specified normatively in oracle/philox.py; it replaces the CuPy XORWOW draws of
// /root/reference/src/spectralmc/async_normals.py:214-215.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace smc {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;
constexpr uint32_t F64_STREAM_BIT = 0x80000000u;

// Round keys are a function of the seed only: computed once on the host and passed by value
// in the kernel parameters, so every use is a constant-bank operand (no per-thread IADDs).
struct PhiloxKeys {
  uint32_t k0[10];
  uint32_t k1[10];
};

inline PhiloxKeys make_philox_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = static_cast<uint32_t>(seed), b = static_cast<uint32_t>(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    k.k0[r] = a;
    k.k1[r] = b;
    a += PHILOX_W0;
    b += PHILOX_W1;
  }
  return k;
}

// One Philox4x32-10 block.  Counter layout used by every caller: (c0, c1, c2, c3) =
// (path column, row group q, matrix_index lo, matrix_index hi | dtype bit).  c1 is the only
// word that changes inside a path's time loop; it sits in an XOR slot so that both first-round
// products and one second-round product are loop-invariant (17 IMAD.WIDE per block, not 20).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKeys& key, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(PHILOX_M0) * c0;  // IMAD.WIDE.U32
    const uint64_t p1 = static_cast<uint64_t>(PHILOX_M1) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ key.k0[r];  // LOP3
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ key.k1[r];
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// ---- MUFU wrappers (explicit PTX so the SASS shows MUFU.LG2/SQRT/SIN/COS/EX2) ----------
__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_sin(float x) {
  float y;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_cos(float x) {
  float y;
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- float32 block: 4 words -> 2 Box–Muller pairs -> 4 normals ---------------------------
// radius uniform u1 = (m + 0.5) 2^-23 with m = x >> 9 (one LEA.HI builds the float in [1, 2));
// the m == 0 bin (probability 2^-23) is refined with the 9 low bits to (j + 0.5) 2^-32, so the
// tail reaches 6.76 sigma.  Every step is exact in float32 and none is an integer->float
// conversion (I2F would land on the XU pipe the MUFUs need).  The refinement sits behind a
// single rarely-taken branch per block so the hot loop does not issue its instructions.
static __device__ __noinline__ float refine_radius_uniform(uint32_t x, float u) {
  if (x < 512u) {
    const float t = __uint_as_float((x << 14) + 0x3f800000u);  // 1 + j 2^-9, j = x & 0x1ff
    u = (t - 0x1.ff8p-1f) * 0x1p-23f;                          // (j + 0.5) 2^-9 * 2^-23
  }
  return u;
}

__device__ __forceinline__ void box_muller_f32(float u1, uint32_t xb, float& z_even, float& z_odd) {
  // w = u2 - 0.5 - 2^-24 exactly; theta = 2 pi (u2 - 0.5) in (-pi, pi)
  const float w = __uint_as_float((xb >> 9) + 0x3f800000u) - 1.5f;
  const float theta = fmaf(w, 6.28318530717958648f, 3.74507028e-07f /* 2 pi 2^-24 */);
  const float r = mufu_sqrt(-1.38629436111989062f /* -2 ln 2 */ * mufu_lg2(u1));
  z_even = r * mufu_cos(theta);
  z_odd = r * mufu_sin(theta);
}

// 4 normals for rows 4q .. 4q+3 of column `col` of matrix (k_lo, k_hi)
__device__ __forceinline__ void normals4_f32(uint32_t col, uint32_t q, uint32_t k_lo, uint32_t k_hi,
                                             const PhiloxKeys& key, float (&z)[4]) {
  uint32_t x[4];
  philox4x32_10(col, q, k_lo, k_hi, key, x);
  // (1 + m 2^-23) - (1 - 2^-24) = (m + 0.5) 2^-23
  float ua = __uint_as_float((x[0] >> 9) + 0x3f800000u) - 0x1.fffffep-1f;
  float uc = __uint_as_float((x[2] >> 9) + 0x3f800000u) - 0x1.fffffep-1f;
  if (__builtin_expect(min(x[0], x[2]) < 512u, 0)) {
    ua = refine_radius_uniform(x[0], ua);
    uc = refine_radius_uniform(x[2], uc);
  }
  box_muller_f32(ua, x[1], z[0], z[1]);
  box_muller_f32(uc, x[3], z[2], z[3]);
}

// ---- float64 block: 4 words -> 1 pair --------------------------------------------------
__device__ __forceinline__ void normals2_f64(uint32_t col, uint32_t q, uint32_t k_lo, uint32_t k_hi,
                                             const PhiloxKeys& key, double (&z)[2]) {
  uint32_t x[4];
  philox4x32_10(col, q, k_lo, k_hi | F64_STREAM_BIT, key, x);
  // u = (m + 0.5) 2^-52, m = low 52 bits of (hi:lo); exact
  const double u1 =
      __hiloint2double(static_cast<int>((x[0] & 0x000fffffu) | 0x3ff00000u), static_cast<int>(x[1])) -
      0x1.fffffffffffffp-1;
  const double w =
      __hiloint2double(static_cast<int>((x[2] & 0x000fffffu) | 0x3ff00000u), static_cast<int>(x[3])) -
      1.5;  // u2 - 0.5 - 2^-53
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * w + 0x1p-52, &s, &c);  // angle / pi = 2 (u2 - 0.5), exact
  z[0] = r * c;
  z[1] = r * s;
}

// ---- per-contract constants --------------------------------------------------------------
// Derived in float64 from the six contract scalars (the reference passes them to the kernel as
// float64, gbm.py:419-425) and rounded once where a float32 kernel consumes them.
struct ContractRow {
  double X0, K, T, r, d, v;
};

__device__ __forceinline__ ContractRow load_contract(const double* __restrict__ contracts, int64_t c) {
  const double* p = contracts + 6 * c;
  ContractRow k;
  k.X0 = __ldg(p + 0);
  k.K = __ldg(p + 1);
  k.T = __ldg(p + 2);
  k.r = __ldg(p + 3);
  k.d = __ldg(p + 4);
  k.v = __ldg(p + 5);
  return k;
}

// fixed-order block sum of one double per thread (deterministic: shuffle tree + smem, no atomics).
// `scratch` must hold >= 32 doubles.  Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nwarps = (blockDim.x + 31) >> 5;
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < nwarps; ++w) t += scratch[w];
  return t;
}

}  // namespace smc
