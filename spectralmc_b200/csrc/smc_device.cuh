// Device-side building blocks: Philox4x32-10, exact uniforms, Box–Muller (MUFU for float32,
// libdevice for float64) and the GBM step algebra.  The stream these functions define is
// specified normatively in oracle/philox.py; it replaces the CuPy XORWOW draws of
// /root/reference/src/spectralmc/async_normals.py:214-215.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#ifndef SMC_NS
#define SMC_NS smc
#endif

namespace SMC_NS {

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
#ifndef SMC_PHILOX_ROUNDS
#define SMC_PHILOX_ROUNDS 10  // 7 in the -DSMC_STREAM_P7 build of the stream-drawing translation units (smc_internal.h)
#endif
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKeys& key, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < SMC_PHILOX_ROUNDS; ++r) {
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

// ---- float32 block: 4 words (128 bits) -> 3 Box–Muller pairs -> 6 normals ------------------
// Each uniform gets 21 bits: the radius fields are the top 21 bits of x0, x1, x2; the angle
// fields are the low 11 bits of the same word followed by a 10-bit slice of x3 (126 bits used).
// u = (F + 0.5) 2^-21 is built as a float in [1, 2) with one shift/funnel-shift + one LOP3 and an
// exact FADD — no integer->float conversion (I2F would land on the XU pipe the MUFUs need).
// Six normals per block instead of four cuts the Philox (IMAD.WIDE) work per normal by a third;
// the kernel is then bounded by the XU pipe and the issue port rather than by the FMA-heavy pipe.
// A zero radius field (probability 2^-21) is refined with 23 fresh bits from a second block whose
// counter has the top bit of the row-group word set, extending the tail to 7.9 sigma; it sits
// behind one rarely-taken branch per block.
constexpr uint32_t F32_REFINE_BIT = 0x80000000u;
// Short matrices (float32, rows <= 3; the reference's own tests run ONE timestep): G = 6 / rows adjacent
// columns share one block — counter (column / G, F32_SHORT_BIT, k lo, k hi) — and element (i, j) is
// normal (j % G) * rows + i of it, so a block's six normals are all consumed (oracle/philox.py).
constexpr uint32_t F32_SHORT_BIT = 0x40000000u;

#ifndef SMC_LOP3_REG
#define SMC_LOP3_REG 1
#endif
__device__ __forceinline__ float unit_float_21(uint32_t field_in_bits_22_2) {
#if SMC_LOP3_REG
  // (v & mask) | one as ONE LOP3: the instruction takes a single immediate, so the exponent word has to sit in a register —
  // ptxas never puts it there by itself (it emits AND-immediate + OR-immediate: 24 of the 76 ALU instructions per 12 normals).
  uint32_t one, d;
  asm("mov.b32 %0, 0x3f800000;" : "=r"(one));
  asm("lop3.b32 %0, %1, 0x007ffffc, %2, 0xEA;" : "=r"(d) : "r"(field_in_bits_22_2), "r"(one));
  return __uint_as_float(d);  // 1 + F 2^-21
#else
  return __uint_as_float((field_in_bits_22_2 & 0x007ffffcu) | 0x3f800000u);  // 1 + F 2^-21
#endif
}

#ifndef SMC_RADIUS_LEA
#define SMC_RADIUS_LEA 1
#endif
// The 21-bit radius field (top 21 bits of a word) as a float in [1, 2) with the field in the LOW mantissa bits: a shift and an
// add of the exponent word, which is ONE instruction (LEA.HI) — the field at mantissa bits 22..2 (unit_float_21) takes a shift
// and a LOP3.  The value is 1 + R 2^-23, a quarter of the scale: the callers form the uniform as 4 f - (4 - 2^-22) = (R + 0.5) 2^-21
// in one exact FFMA where they had one exact FADD, so every value downstream is bit-identical to the two-instruction form.
// (Carrying the factor through the logarithm instead — lg2(u / 4) + 2 — loses MUFU.LG2's absolute-error regime near u = 1 and
// cancels for small radii: tried, caught by tests/test_gpu_fused.py.)
__device__ __forceinline__ float radius_float_top21(uint32_t word) {
#if SMC_RADIUS_LEA
  return __uint_as_float((word >> 11) + 0x3f800000u);
#else
  return unit_float_21(word >> 9);
#endif
}

__device__ __forceinline__ void box_muller_f32(float u1, float angle_unit, float& z_even, float& z_odd) {
  // angle_unit = 1 + A 2^-21;  u2 - 0.5 = angle_unit - 1.5 + 2^-22;  theta = 2 pi (u2 - 0.5) in ONE
  // FFMA: 2 pi angle_unit - (3 pi - 2 pi 2^-22).  The product is < 4 pi, so theta carries an absolute
  // rounding error <= 4.8e-7 rad — the size of MUFU.SIN/COS's own error (2^-20.9) and unbiased.
  // Saving the separate exact subtraction is worth 2.2 % on the fused kernel (1.377 vs 1.408 ms at c2).
  const float theta = fmaf(angle_unit, 6.28318530717958648f, -9.424776462741265f);
  const float r = mufu_sqrt(-1.38629436111989062f /* -2 ln 2 */ * mufu_lg2(u1));
  z_even = r * mufu_cos(theta);
  z_odd = r * mufu_sin(theta);
}

#ifndef SMC_SHL_ALU
#define SMC_SHL_ALU 0
#endif
// x << n; SMC_SHL_ALU = 1 asks for the funnel-shift form (ALU pipe) instead of ptxas' IMAD.SHL (FMA-heavy pipe, where Philox lives)
template <int N>
__device__ __forceinline__ uint32_t shl_bits(uint32_t x) {
#if SMC_SHL_ALU
  uint32_t d;
  asm("shf.l.clamp.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(0u), "r"(x), "n"(N));
  return d;
#else
  return x << N;
#endif
}

// the rare path: radius field of pair `p` was zero -> u = (m + 0.5) 2^-44, m = top 23 bits of
// word p of the refinement block.  Everything is passed by value (registers): taking the address
// of the kernel-parameter key block would force a local-memory copy in the caller.
static __device__ __noinline__ float3 refine_radius_uniforms(uint32_t col, uint32_t q, uint32_t k_lo, uint32_t k_hi,
                                                             uint32_t seed_lo, uint32_t seed_hi, uint32_t x0, uint32_t x1,
                                                             uint32_t x2, float u0, float u1, float u2) {
  PhiloxKeys key;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    key.k0[r] = seed_lo + static_cast<uint32_t>(r) * PHILOX_W0;
    key.k1[r] = seed_hi + static_cast<uint32_t>(r) * PHILOX_W1;
  }
  uint32_t y[4];
  philox4x32_10(col, q | F32_REFINE_BIT, k_lo, k_hi, key, y);
  if ((x0 >> 11) == 0u) u0 = (__uint_as_float((y[0] >> 9) + 0x3f800000u) - 0x1.fffffep-1f) * 0x1p-21f;
  if ((x1 >> 11) == 0u) u1 = (__uint_as_float((y[1] >> 9) + 0x3f800000u) - 0x1.fffffep-1f) * 0x1p-21f;
  if ((x2 >> 11) == 0u) u2 = (__uint_as_float((y[2] >> 9) + 0x3f800000u) - 0x1.fffffep-1f) * 0x1p-21f;
  return make_float3(u0, u1, u2);
}

// 6 normals for rows 6q .. 6q+5 of column `col` of matrix (k_lo, k_hi).
// REFINE = true is the stream as specified (one rarely-taken branch per block).
// REFINE = false draws the coarse radius uniform for every pair and only folds the three radius
// words into `min_word`; a caller that consumes a whole path can test `min_word < 2048` ONCE after
// its time loop and, in the 6e-5 of paths where some block needed the refinement, redo the path
// with REFINE = true.  The result is identical to the specified stream; the hot loop loses the
// compare/branch/reconvergence instructions (5 of 98 issue slots per block).
// NPAIRS < 3 evaluates only the first NPAIRS Box–Muller pairs (rows 6q .. 6q+2*NPAIRS-1): the
// ragged last block of a path whose length is not a multiple of 6, and in particular the whole
// path when timesteps <= 4 (the reference's own tests run timesteps = 1).
template <bool REFINE, int NPAIRS = 3>
__device__ __forceinline__ void normals6_f32_impl(uint32_t col, uint32_t q, uint32_t k_lo, uint32_t k_hi,
                                                  const PhiloxKeys& key, float (&z)[6], uint32_t& min_word) {
  uint32_t x[4];
  philox4x32_10(col, q, k_lo, k_hi, key, x);
  float u[3];
#pragma unroll
#if SMC_RADIUS_LEA
  for (int p = 0; p < 3; ++p) u[p] = fmaf(radius_float_top21(x[p]), 4.0f, -0x1.fffffep+1f);  // 4 (1 + R 2^-23) - (4 - 2^-22) = (R + 0.5) 2^-21, exact
#else
  for (int p = 0; p < 3; ++p) u[p] = unit_float_21(x[p] >> 9) - 0x1.fffff8p-1f;  // (R + 0.5) 2^-21
#endif
  if (REFINE) {
    if (__builtin_expect(min(min(x[0], x[1]), x[2]) < 2048u, 0)) {
      const float3 f = refine_radius_uniforms(col, q, k_lo, k_hi, key.k0[0], key.k1[0], x[0], x[1], x[2], u[0], u[1], u[2]);
      u[0] = f.x;
      u[1] = f.y;
      u[2] = f.z;
    }
  } else {
    min_word = min(min(min_word, x[0]), min(x[1], x[2]));  // unused pairs may only cause a needless exact redo
  }
#ifndef SMC_BM_ORDER
#define SMC_BM_ORDER 2  // evaluation order of the pairs: a pure scheduling hint for ptxas (results identical)
#endif
#if SMC_BM_ORDER == 0
  box_muller_f32(u[0], unit_float_21(__funnelshift_l(x[3], x[0], 12)), z[0], z[1]);
  if (NPAIRS >= 2) box_muller_f32(u[1], unit_float_21(__funnelshift_l(shl_bits<10>(x[3]), x[1], 12)), z[2], z[3]);
  if (NPAIRS >= 3) box_muller_f32(u[2], unit_float_21(__funnelshift_l(shl_bits<20>(x[3]), x[2], 12)), z[4], z[5]);
#elif SMC_BM_ORDER == 1
  if (NPAIRS >= 3) box_muller_f32(u[2], unit_float_21(__funnelshift_l(shl_bits<20>(x[3]), x[2], 12)), z[4], z[5]);
  if (NPAIRS >= 2) box_muller_f32(u[1], unit_float_21(__funnelshift_l(shl_bits<10>(x[3]), x[1], 12)), z[2], z[3]);
  box_muller_f32(u[0], unit_float_21(__funnelshift_l(x[3], x[0], 12)), z[0], z[1]);
#else
  const float a0 = unit_float_21(__funnelshift_l(x[3], x[0], 12));
  const float a1 = unit_float_21(__funnelshift_l(shl_bits<10>(x[3]), x[1], 12));
  const float a2 = unit_float_21(__funnelshift_l(shl_bits<20>(x[3]), x[2], 12));
  box_muller_f32(u[0], a0, z[0], z[1]);
  if (NPAIRS >= 2) box_muller_f32(u[1], a1, z[2], z[3]);
  if (NPAIRS >= 3) box_muller_f32(u[2], a2, z[4], z[5]);
#endif
}

// ---- packed FP32 (sm_100 FADD2 / FMUL2 / FFMA2: two float lanes per issue slot) -------------
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack_f32x2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ float2 add_f32x2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack_f32x2(a.x, a.y)), "l"(pack_f32x2(b.x, b.y)));
  return unpack_f32x2(d);
}
__device__ __forceinline__ float2 mul_f32x2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack_f32x2(a.x, a.y)), "l"(pack_f32x2(b.x, b.y)));
  return unpack_f32x2(d);
}
__device__ __forceinline__ float2 fma_f32x2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(pack_f32x2(a.x, a.y)), "l"(pack_f32x2(b.x, b.y)), "l"(pack_f32x2(c.x, c.y)));
  return unpack_f32x2(d);
}

// Two Box–Muller pairs at once, accumulated: acc += (z_even + z_odd) / sqrt(2) of pair A in lane x and of pair B in
// lane y.  Only the SUM of a pair's two normals enters the log-Euler path, and
//   r cos(theta) + r sin(theta) = sqrt(2) r sin(theta + pi / 4),
// so ONE MUFU.SIN replaces MUFU.COS + MUFU.SIN — three MUFUs per pair instead of four on the XU pipe that bounds the
// kernel — and the pi / 4 rides in the constant of the FFMA that forms the angle.  The caller multiplies the path's
// accumulated sum by sqrt(2) once.  (Same draws as box_muller_f32; the MUFU error per pair is that of one sine.)
__device__ __forceinline__ float2 box_muller_sum_f32x2(float2 radius_unit, float2 angle_unit, float2 acc) {
#if SMC_RADIUS_LEA
  // radius_unit = 1 + R 2^-23 (radius_float_top21): u = 4 f - (4 - 2^-22) = (R + 0.5) 2^-21, exact
  const float2 u = fma_f32x2(radius_unit, make_float2(4.0f, 4.0f), make_float2(-0x1.fffffep+1f, -0x1.fffffep+1f));
#else
  const float2 u = add_f32x2(radius_unit, make_float2(-0x1.fffff8p-1f, -0x1.fffff8p-1f));
#endif
  // theta + pi / 4 = 2 pi angle_unit - (3 pi - 2 pi 2^-22) + pi / 4
  const float2 theta = fma_f32x2(angle_unit, make_float2(6.28318530717958648f, 6.28318530717958648f),
                                 make_float2(-8.639378299343817f, -8.639378299343817f));
  const float2 m = mul_f32x2(make_float2(mufu_lg2(u.x), mufu_lg2(u.y)),
                             make_float2(-1.38629436111989062f, -1.38629436111989062f));
  const float2 r = make_float2(mufu_sqrt(m.x), mufu_sqrt(m.y));
  return fma_f32x2(r, make_float2(mufu_sin(theta.x), mufu_sin(theta.y)), acc);
}

// Sum of the 12 normals of row groups q and q + 1 (coarse radius uniforms, like REFINE = false), divided by sqrt(2),
// added to acc.x + acc.y.  Three packed Box–Muller evaluations over the six pairs.
__device__ __forceinline__ float2 normals12_sum_f32x2(uint32_t col, uint32_t q, uint32_t k_lo, uint32_t k_hi,
                                                      const PhiloxKeys& key, float2 acc, uint32_t& min_word) {
  uint32_t x[4], y[4];
  philox4x32_10(col, q, k_lo, k_hi, key, x);
  philox4x32_10(col, q + 1u, k_lo, k_hi, key, y);
  min_word = min(min(min_word, min(min(x[0], x[1]), x[2])), min(min(y[0], y[1]), y[2]));  // three 3-input minima
  const float ax0 = unit_float_21(__funnelshift_l(x[3], x[0], 12));
  const float ax1 = unit_float_21(__funnelshift_l(shl_bits<10>(x[3]), x[1], 12));
  const float ax2 = unit_float_21(__funnelshift_l(shl_bits<20>(x[3]), x[2], 12));
  const float ay0 = unit_float_21(__funnelshift_l(y[3], y[0], 12));
  const float ay1 = unit_float_21(__funnelshift_l(shl_bits<10>(y[3]), y[1], 12));
  const float ay2 = unit_float_21(__funnelshift_l(shl_bits<20>(y[3]), y[2], 12));
  acc = box_muller_sum_f32x2(make_float2(radius_float_top21(x[0]), radius_float_top21(y[0])), make_float2(ax0, ay0), acc);
  acc = box_muller_sum_f32x2(make_float2(radius_float_top21(x[1]), radius_float_top21(y[1])), make_float2(ax1, ay1), acc);
  return box_muller_sum_f32x2(make_float2(radius_float_top21(x[2]), radius_float_top21(y[2])), make_float2(ax2, ay2), acc);
}

__device__ __forceinline__ void normals6_f32(uint32_t col, uint32_t q, uint32_t k_lo, uint32_t k_hi,
                                             const PhiloxKeys& key, float (&z)[6]) {
  uint32_t unused = 0;
  normals6_f32_impl<true>(col, q, k_lo, k_hi, key, z, unused);
}

// ---- float64 block: 4 words -> 2 pairs (oracle/philox.py) -------------------------------
// One Philox block feeds TWO Box-Muller pairs, 64 bits each.  The words (w0, w1) of a pair:
//   radius field R = w0 << 11 | w1 >> 21 (43 bits), angle field A = w1 & 0x1fffff (21 bits).
// Both uniforms are put together as the mantissa of a double in [1, 2), so no integer -> float conversion runs:
//   1 + (R + 0.5) 2^-43 : high word 0x3ff00000 | R >> 23, low word (R << 9 | 0x100) mod 2^32
//   1 + (A + 0.5) 2^-21 : high word 0x3ff00000 | A >> 1,  low word (A & 1) << 31 | 0x40000000
//
// Logarithm, square root and sine / cosine are specialised to the arguments Box-Muller produces (u in (0, 1), a normal
// number; -2 ln u in (0, 62); |angle / pi| <= 1) and written for the FP64 pipe, which bounds these kernels (DFMA issues
// every 2 cycles per warp, profiles/r2_pipe_overlap_f64_microbench.txt): every FP64 instruction that is not needed for
// 1e-13 is gone, and the polynomial coefficients live in __constant__ memory so each DFMA takes its coefficient as a
// constant-bank operand (libdevice's general-purpose versions made ptxas rebuild ~40 64-bit literals per iteration and
// carry a range test, a branch and an out-of-line slow path per division / square root).  The polynomials are the
// classical minimax sets of Sun's FDLIBM (k_sin.c, k_cos.c, e_log.c); tests/test_gpu_normals.py holds the device
// output to 1e-13 of the float64 oracle.  43 FP64 arithmetic instructions per pair (59 with libdevice-style last-bit
// corrections, 51 before the radius was computed as one chain):
//   radius (24)  x = -2 ln u = -2 k ln 2 + s2 (2 + R(s2^2)),  s2 = -2 f / (2 + f) = f / -(1 + f / 2),  u = 2^k (1 + f):
//                the quotient is MUFU.RCP64H (2^-22) + one cubic Newton step (2^-66) without residual correction, the
//                factor -2 rides on the divisor, FDLIBM's Lg_n are pre-divided by 4^n (exact) because s2^2 = 4 s^2, and
//                2 + R is one Estrin pair.  r = sqrt(x) = g0 (1 + e / 2 + 3 e^2 / 8), g0 = x y0, e = 1 - g0 y0 from
//                MUFU.RSQ64H (2^-22; truncation 5 e^3 / 16 < 2^-67), again without residual correction.
//   angle (17)   v = 2 angle / pi = 4 (1 + u2) - 6 exactly; n = rint(v), x = (v - n) pi / 2 (one-word pi / 2: absolute
//                error 2^-55), sin x = x + x z S(z), cos x = 1 + z (-1/2 + z C(z)), quadrant n mod 4.
static __constant__ double kSinCoef[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03,
                                          -1.98412698298579493134e-04, 2.75573137070700676789e-06,
                                          -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double kCosCoef[6] = {4.16666666666666019037e-02,  -1.38888888888741095749e-03,
                                          2.48015872894767294178e-05,  -2.75573143513906633035e-07,
                                          2.08757232129817482790e-09,  -1.13596475577881948265e-11};
// Lg_n / 4^n, n = 1..7 (e_log.c's Lg1..Lg7; the power-of-two scaling is exact)
static __constant__ double kLogCoef[7] = {6.666666666666735130e-01 / 4.0,     3.999999999940941908e-01 / 16.0,
                                          2.857142874366239149e-01 / 64.0,    2.222219843214978396e-01 / 256.0,
                                          1.818357216161805012e-01 / 1024.0,  1.531383769920937332e-01 / 4096.0,
                                          1.479819860511658591e-01 / 16384.0};
static __constant__ double kF64Misc[5] = {-2.0 * 6.93147180559945286227e-01 /* -2 ln 2 */, 1.57079632679489655800e+00 /* pi / 2 */,
                                          4.0, -6.0 /* v = 4 (1 + u2) - 6 */, 0.375};

// r = sqrt(-2 ln u) for u = d - 1, d in [1, 2) with a non-zero mantissa (so u is a normal number in (0, 1))
__device__ __forceinline__ double radius_f64(double d) {
  const double u = d - 1.0;  // exact, and normalises the mantissa
  // u = 2^k m, m in [sqrt(2) / 2, sqrt(2)): shift the high word so that the exponent field steps at sqrt(2) (branch-free)
  const int hx = __double2hiint(u) + (0x3ff00000 - 0x3fe6a09e);
  const double kd = static_cast<double>((hx >> 20) - 1023);  // I2F.F64; the 2^52 magic-number form measured 2.6 % slower
  const double f = __hiloint2double((hx & 0x000fffff) + 0x3fe6a09e, __double2loint(u)) - 1.0;
  const double dn = fma(f, -0.5, -1.0);  // -(2 + f) / 2
  double rc;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(dn));
  double e = fma(-dn, rc, 1.0);
  e = fma(e, e, e);                      // e + e^2
  rc = fma(rc, e, rc);                   // rc (1 + e + e^2): relative error 2^-66
  const double s2 = f * rc;              // -2 s
  const double z = s2 * s2, w = z * z;
  const double even = fma(w, fma(w, fma(w, kLogCoef[5], kLogCoef[3]), kLogCoef[1]), 2.0);   // 2 + Lg2 z^2 + Lg4 z^4 + Lg6 z^6
  const double odd = fma(w, fma(w, fma(w, kLogCoef[6], kLogCoef[4]), kLogCoef[2]), kLogCoef[0]);
  const double x = fma(kd, kF64Misc[0], s2 * fma(z, odd, even));  // -2 ln u > 0
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double g = x * y;
  const double e2 = fma(-g, y, 1.0);                       // 1 - x y^2
  return fma(g * e2, fma(e2, kF64Misc[4], 0.5), g);       // g (1 + e / 2 + 3 e^2 / 8)
}

// (s0, c0) = (sin, cos)(pi t - n pi / 2), n = rint(2 t), for t = 2 d - 3, d = 1 + u2 in [1, 2); returns n
__device__ __forceinline__ int sincos_reduced_f64(double d, double& s0, double& c0) {
  const double v = fma(d, kF64Misc[2], kF64Misc[3]);  // 2 t in [-2, 2), exact
  const double n = rint(v);
  const double x = (v - n) * kF64Misc[1];             // |x| <= pi / 4
  const double z = x * x;
  const double ps = fma(z, fma(z, fma(z, fma(z, fma(z, kSinCoef[5], kSinCoef[4]), kSinCoef[3]), kSinCoef[2]), kSinCoef[1]), kSinCoef[0]);
  const double pc = fma(z, fma(z, fma(z, fma(z, fma(z, kCosCoef[5], kCosCoef[4]), kCosCoef[3]), kCosCoef[2]), kCosCoef[1]), kCosCoef[0]);
  s0 = fma(x * z, ps, x);
  c0 = fma(z, fma(z, pc, -0.5), 1.0);
  return static_cast<int>(n);
}

// (v & mask) | bits as ONE LOP3 with `bits` in a register (see unit_float_21)
template <uint32_t MASK, uint32_t BITS>
__device__ __forceinline__ uint32_t and_or_bits(uint32_t v) {
#if SMC_LOP3_REG
  uint32_t bits, d;
  asm("mov.b32 %0, %1;" : "=r"(bits) : "n"(BITS));
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(v), "n"(MASK), "r"(bits));
  return d;
#else
  return (v & MASK) | BITS;
#endif
}
__device__ __forceinline__ double f64_radius(uint32_t w0, uint32_t w1) {
  const uint32_t lo = and_or_bits<0xfffffe00u, 0x100u>(__funnelshift_l(w1, w0, 20));
  return radius_f64(__hiloint2double(static_cast<int>((w0 >> 12) | 0x3ff00000u), static_cast<int>(lo)));
}
__device__ __forceinline__ double f64_angle(uint32_t w1) {  // 1 + u2
  return __hiloint2double(static_cast<int>(and_or_bits<0x000fffffu, 0x3ff00000u>(w1 >> 1)), static_cast<int>((w1 << 31) | 0x40000000u));
}

// (sin(pi y) / y - pi) / y^2 on |y| <= 1/2, degree 7 in y^2 (interpolated at the Chebyshev nodes with 50-digit arithmetic;
// the double-rounded set reproduces sin(pi y) to 9e-17 absolute, the rounding of pi included)
static __constant__ double kSinPiCoef[9] = {3.14159265358979311600e+00,  -5.16771278004997025590e+00, 2.55016403987734019410e+00,
                                            -5.99264529320340910701e-01, 8.21458865966949863813e-02,  -7.37043071891425437964e-03,
                                            4.66300869507401522483e-04,  -2.19061871420551991800e-05, 7.72556449995128730662e-07};
static __constant__ double kSumMisc[2] = {2.0, -2.75};  // angle / pi + 1/4 = 2 (1 + u2) - 2.75

// acc + (the four normals of normals4_f64) / sqrt(2) — the same draws, for the log-Euler sum, where only the sum of a
// pair's two normals matters:  r cos(theta) + r sin(theta) = sqrt(2) r sin(theta + pi / 4),  ONE odd polynomial on half
// a period instead of a sine and a cosine on a quarter (13 instead of 19 FP64 instructions per pair; the caller
// multiplies the accumulated sum by sqrt(2) once per path).  With t = theta / pi + 1/4 in [-3/4, 5/4), n = rint(t) in
// {-1, 0, 1} and y = t - n:  sin(pi t) = (-1)^n sin(pi y),  sin(pi y) = y (pi + y^2 Q(y^2)).
__device__ __forceinline__ double normals4_sum_f64(uint32_t col, uint32_t q, uint32_t k_lo, uint32_t k_hi,
                                                   const PhiloxKeys& key, double acc) {
  uint32_t x[4];
  philox4x32_10(col, q, k_lo, k_hi | F64_STREAM_BIT, key, x);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const double r = f64_radius(x[2 * p], x[2 * p + 1]);
    const double t = fma(f64_angle(x[2 * p + 1]), kSumMisc[0], kSumMisc[1]);  // exact
#ifndef SMC_F64_MAGIC_RINT
#define SMC_F64_MAGIC_RINT 1  // t + 1.5 * 2^52 - 1.5 * 2^52 (two DADD) instead of FRND.F64 + F2I.F64: c4s 42.09 -> 41.38 ms
#endif
#if SMC_F64_MAGIC_RINT
    const double tmp = t + 0x1.8p52;                                          // integer part of t lands in the low mantissa bits
    const double n = tmp - 0x1.8p52;
#else
    const double n = rint(t);
#endif
    const double y = t - n;                                                   // exact, |y| <= 1/2
    const double w = y * y;
    double poly = kSinPiCoef[8];
#pragma unroll
    for (int c = 7; c >= 0; --c) poly = fma(w, poly, kSinPiCoef[c]);
    const double sn = y * poly;                                               // sin(pi y)
#if SMC_F64_MAGIC_RINT
    const uint32_t flip = static_cast<uint32_t>(__double2loint(tmp)) << 31;   // n odd
#else
    const uint32_t flip = static_cast<uint32_t>(static_cast<int>(n)) << 31;   // n odd
#endif
    acc = fma(__hiloint2double(__double2hiint(r) ^ static_cast<int>(flip), __double2loint(r)), sn, acc);
  }
  return acc;
}

// z[0..3] = rows 4q .. 4q + 3 of column `col`; `pairs` (1 or 2) = how many of the two pairs the caller needs
__device__ __forceinline__ void normals4_f64(uint32_t col, uint32_t q, uint32_t k_lo, uint32_t k_hi,
                                             const PhiloxKeys& key, double (&z)[4], int pairs = 2) {
  uint32_t x[4];
  philox4x32_10(col, q, k_lo, k_hi | F64_STREAM_BIT, key, x);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    if (p < pairs) {
      const double r = f64_radius(x[2 * p], x[2 * p + 1]);
      double s0, c0;
      const int n = sincos_reduced_f64(f64_angle(x[2 * p + 1]), s0, c0) & 3;
      const double a = (n & 1) ? c0 : s0, b = (n & 1) ? s0 : c0;
      z[2 * p] = r * (((n + 1) & 2) ? -b : b);  // cos: n = 0 -> c, 1 -> -s, 2 -> -c, 3 -> s
      z[2 * p + 1] = r * ((n & 2) ? -a : a);    // sin: n = 0 -> s, 1 -> c, 2 -> -s, 3 -> -c
    } else {
      z[2 * p] = z[2 * p + 1] = 0.0;
    }
  }
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

}  // namespace SMC_NS
