// Tuning microbenchmark (not part of the product): times variants of the fused inner loop
// (Philox4x32-10 -> Box-Muller -> log-return sum) to separate scheduling effects from pipe limits.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../include -I../../spectralmc_b200/csrc \
//        -o fused_variants fused_variants.cu && ./fused_variants
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "smc_device.cuh"

using namespace smc;

constexpr int T = 252;

__device__ __forceinline__ uint32_t mulhi_ptx(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t mullo_ptx(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

template <int ROUNDS>
__device__ __forceinline__ void philox_split(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& key,
                                             uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t h0 = mulhi_ptx(PHILOX_M0, c0), l0 = mullo_ptx(PHILOX_M0, c0);
    const uint32_t h1 = mulhi_ptx(PHILOX_M1, c2), l1 = mullo_ptx(PHILOX_M1, c2);
    const uint32_t n0 = h1 ^ c1 ^ key.k0[r];
    const uint32_t n2 = h0 ^ c3 ^ key.k1[r];
    c1 = l1; c3 = l0; c0 = n0; c2 = n2;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <int ROUNDS>
__device__ __forceinline__ void philox_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& key,
                                         uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(PHILOX_M0) * c0;
    const uint64_t p1 = static_cast<uint64_t>(PHILOX_M1) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ key.k0[r];
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ key.k1[r];
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// MODE 0 full; 1 no refinement branch; 2 philox only (xor-sum); 3 box-muller only (cheap counter hash)
template <int ROUNDS, int MODE>
__device__ __forceinline__ float block_sum4(uint32_t col, uint32_t q, const PhiloxKeys& key) {
  uint32_t x[4];
  if (MODE == 3) {
    x[0] = col * 2654435761u + q * 40503u; x[1] = x[0] ^ (q * 2246822519u); x[2] = x[1] + col; x[3] = x[2] ^ 0x9e3779b9u;
  } else if (MODE >= 4) {
    philox_split<ROUNDS>(col, q, 7u, 0u, key, x);
  } else {
    philox_r<ROUNDS>(col, q, 7u, 0u, key, x);
  }
  if (MODE == 2 || MODE == 5) return __uint_as_float(((x[0] ^ x[1] ^ x[2] ^ x[3]) >> 9) | 0x3f800000u);
  float ua = __uint_as_float((x[0] >> 9) + 0x3f800000u) - 0x1.fffffep-1f;
  float uc = __uint_as_float((x[2] >> 9) + 0x3f800000u) - 0x1.fffffep-1f;
  if (MODE == 0 || MODE == 4) {
    if (__builtin_expect(min(x[0], x[2]) < 512u, 0)) {
      ua = ua * 0.5f;
      uc = uc * 0.5f;
    }
  }
  float z0, z1, z2, z3;
  box_muller_f32(ua, __uint_as_float((x[1] >> 9) + 0x3f800000u), z0, z1);
  box_muller_f32(uc, __uint_as_float((x[3] >> 9) + 0x3f800000u), z2, z3);
  return (z0 + z1) + (z2 + z3);
}

// 21-bit uniforms: one Philox block -> 3 Box-Muller pairs -> 6 normals
__device__ __forceinline__ float u21(uint32_t top_aligned) {  // field in bits 31..11
  return __uint_as_float(((top_aligned >> 9) & 0x007ffffcu) | 0x3f800000u);
}
template <int MODE>
__device__ __forceinline__ float block_sum6(uint32_t col, uint32_t q, const PhiloxKeys& key) {
  uint32_t x[4];
  philox_r<10>(col, q, 7u, 0u, key, x);
  const uint32_t f0 = x[0], f1 = __funnelshift_l(x[1], x[0], 21), f2 = x[1] << 10;
  const uint32_t f3 = x[2], f4 = __funnelshift_l(x[3], x[2], 21), f5 = x[3] << 10;
  float ua = u21(f0) - 0x1.fffffcp-1f, ub = u21(f2) - 0x1.fffffcp-1f, uc = u21(f4) - 0x1.fffffcp-1f;
  if (MODE == 0) {
    if (__builtin_expect(min(min(f0, f2), f4) < 2048u, 0)) {
      ua *= 0.5f; ub *= 0.5f; uc *= 0.5f;
    }
  }
  float z0, z1, z2, z3, z4, z5;
  {
    const float w = u21(f1) - 1.5f; const float th = fmaf(w, 6.28318530717958648f, 1.4980281e-06f);
    const float r = mufu_sqrt(-1.38629436111989062f * mufu_lg2(ua)); z0 = r * mufu_cos(th); z1 = r * mufu_sin(th);
  }
  {
    const float w = u21(f3) - 1.5f; const float th = fmaf(w, 6.28318530717958648f, 1.4980281e-06f);
    const float r = mufu_sqrt(-1.38629436111989062f * mufu_lg2(ub)); z2 = r * mufu_cos(th); z3 = r * mufu_sin(th);
  }
  {
    const float w = u21(f5) - 1.5f; const float th = fmaf(w, 6.28318530717958648f, 1.4980281e-06f);
    const float r = mufu_sqrt(-1.38629436111989062f * mufu_lg2(uc)); z4 = r * mufu_cos(th); z5 = r * mufu_sin(th);
  }
  return ((z0 + z1) + (z2 + z3)) + (z4 + z5);
}

template <int UNROLL, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) posthoc6_kernel(float* out, int64_t paths_per_thread, PhiloxKeys key) {
  const uint32_t tid = blockIdx.x * BLOCK + threadIdx.x;
  const uint32_t nthreads = gridDim.x * BLOCK;
  float total = 0.f;
  for (int64_t p = 0; p < paths_per_thread; ++p) {
    float s = 0.f;
    uint32_t mw = 0xffffffffu;
    const uint32_t col = tid + static_cast<uint32_t>(p) * nthreads;
#pragma unroll UNROLL
    for (uint32_t q = 0; q < T / 6; ++q) {
      float z[6];
      normals6_f32_impl<false>(col, q, 7u, 0u, key, z, mw);
      s += ((z[0] + z[1]) + (z[2] + z[3])) + (z[4] + z[5]);
    }
    if (mw < 2048u) {
      s = 0.f;
#pragma unroll 1
      for (uint32_t q = 0; q < T / 6; ++q) {
        float z[6];
        normals6_f32(col, q, 7u, 0u, key, z);
        s += ((z[0] + z[1]) + (z[2] + z[3])) + (z[4] + z[5]);
      }
    }
    total += mufu_ex2(s * 0.01f);
  }
  out[tid] = total;
}

template <int UNROLL, int BLOCK, int MINB>
void runh(const char* name) {
  const int64_t total_paths = 8388608;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, posthoc6_kernel<UNROLL, BLOCK, MINB>, BLOCK, 0);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, posthoc6_kernel<UNROLL, BLOCK, MINB>);
  const int grid = static_cast<int>(total_paths / (2 * BLOCK));
  const int64_t ppt = 2;
  float* out;
  cudaMalloc(&out, sizeof(float) * static_cast<size_t>(grid) * BLOCK);
  const PhiloxKeys key = make_philox_keys(7);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(a);
    posthoc6_kernel<UNROLL, BLOCK, MINB><<<grid, BLOCK>>>(out, ppt, key);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  const double steps = static_cast<double>(grid) * BLOCK * ppt * T;
  printf("%-44s regs=%3d occ=%2d grid=%6d ppt=%3lld  %.3f ms  %.3e path-steps/s\n", name, fa.numRegs, occ, grid, (long long)ppt, best, steps / (best * 1e-3));
  cudaFree(out);
}

template <int UNROLL, int PATHS, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) product6_kernel(float* out, int64_t paths_per_thread, PhiloxKeys key) {
  const uint32_t tid = blockIdx.x * BLOCK + threadIdx.x;
  const uint32_t nthreads = gridDim.x * BLOCK;
  float total = 0.f;
  for (int64_t p = 0; p < paths_per_thread; p += PATHS) {
    float s[PATHS];
    uint32_t col[PATHS];
#pragma unroll
    for (int k = 0; k < PATHS; ++k) { s[k] = 0.f; col[k] = tid + static_cast<uint32_t>(p + k) * nthreads; }
#pragma unroll UNROLL
    for (uint32_t q = 0; q < T / 6; ++q) {
#pragma unroll
      for (int k = 0; k < PATHS; ++k) {
        float z[6];
        normals6_f32(col[k], q, 7u, 0u, key, z);
        s[k] += ((z[0] + z[1]) + (z[2] + z[3])) + (z[4] + z[5]);
      }
    }
#pragma unroll
    for (int k = 0; k < PATHS; ++k) total += mufu_ex2(s[k] * 0.01f);
  }
  out[tid] = total;
}

template <int UNROLL, int PATHS, int BLOCK, int MINB>
void runp(const char* name) {
  const int64_t total_paths = 8388608;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, product6_kernel<UNROLL, PATHS, BLOCK, MINB>, BLOCK, 0);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, product6_kernel<UNROLL, PATHS, BLOCK, MINB>);
  const int grid = static_cast<int>(total_paths / (2 * BLOCK));
  const int64_t ppt = 2;
  float* out;
  cudaMalloc(&out, sizeof(float) * static_cast<size_t>(grid) * BLOCK);
  const PhiloxKeys key = make_philox_keys(7);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(a);
    product6_kernel<UNROLL, PATHS, BLOCK, MINB><<<grid, BLOCK>>>(out, ppt, key);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  const double steps = static_cast<double>(grid) * BLOCK * ppt * T;
  printf("%-44s regs=%3d occ=%2d grid=%6d ppt=%3lld  %.3f ms  %.3e path-steps/s\n", name, fa.numRegs, occ, grid, (long long)ppt, best, steps / (best * 1e-3));
  cudaFree(out);
}

template <int MODE, int UNROLL, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) variant6_kernel(float* out, int64_t paths_per_thread, PhiloxKeys key) {
  const uint32_t tid = blockIdx.x * BLOCK + threadIdx.x;
  const uint32_t nthreads = gridDim.x * BLOCK;
  float total = 0.f;
  for (int64_t p = 0; p < paths_per_thread; ++p) {
    float s = 0.f;
    const uint32_t col = tid + static_cast<uint32_t>(p) * nthreads;
#pragma unroll UNROLL
    for (uint32_t q = 0; q < T / 6; ++q) s += block_sum6<MODE>(col, q, key);
    total += mufu_ex2(s * 0.01f);
  }
  out[tid] = total;
}

template <int MODE, int UNROLL, int BLOCK, int MINB>
void run6(const char* name) {
  const int64_t total_paths = 8388608;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, variant6_kernel<MODE, UNROLL, BLOCK, MINB>, BLOCK, 0);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, variant6_kernel<MODE, UNROLL, BLOCK, MINB>);
  const int grid = static_cast<int>(total_paths / (2 * BLOCK));
  const int64_t ppt = 2;
  float* out;
  cudaMalloc(&out, sizeof(float) * static_cast<size_t>(grid) * BLOCK);
  const PhiloxKeys key = make_philox_keys(7);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(a);
    variant6_kernel<MODE, UNROLL, BLOCK, MINB><<<grid, BLOCK>>>(out, ppt, key);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  const double steps = static_cast<double>(grid) * BLOCK * ppt * T;
  printf("%-44s regs=%3d occ=%2d grid=%6d ppt=%3lld  %.3f ms  %.3e path-steps/s\n", name, fa.numRegs, occ, grid, (long long)ppt, best, steps / (best * 1e-3));
  cudaFree(out);
}

template <int ROUNDS, int MODE, int UNROLL, int PATHS, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) variant_kernel(float* out, int64_t paths_per_thread, PhiloxKeys key) {
  const uint32_t tid = blockIdx.x * BLOCK + threadIdx.x;
  const uint32_t nthreads = gridDim.x * BLOCK;
  float total = 0.f;
  for (int64_t p = 0; p < paths_per_thread; p += PATHS) {
    float s[PATHS];
    uint32_t col[PATHS];
#pragma unroll
    for (int k = 0; k < PATHS; ++k) { s[k] = 0.f; col[k] = tid + static_cast<uint32_t>(p + k) * nthreads; }
#pragma unroll UNROLL
    for (uint32_t q = 0; q < T / 4; ++q) {
#pragma unroll
      for (int k = 0; k < PATHS; ++k) s[k] += block_sum4<ROUNDS, MODE>(col[k], q, key);
    }
#pragma unroll
    for (int k = 0; k < PATHS; ++k) total += mufu_ex2(s[k] * 0.01f);
  }
  out[tid] = total;
}

template <int ROUNDS, int MODE, int UNROLL, int PATHS, int BLOCK, int MINB>
void run(const char* name, int sms) {
  const int64_t total_paths = 8388608;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, variant_kernel<ROUNDS, MODE, UNROLL, PATHS, BLOCK, MINB>, BLOCK, 0);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, variant_kernel<ROUNDS, MODE, UNROLL, PATHS, BLOCK, MINB>);
  // same decomposition as the product: 16384 CTAs x 256 threads x 2 paths (scaled for other block sizes)
  const int grid = static_cast<int>(total_paths / (2 * BLOCK) / (PATHS > 2 ? PATHS / 2 : 1));
  const int64_t ppt = total_paths / (static_cast<int64_t>(grid) * BLOCK);
  float* out;
  cudaMalloc(&out, sizeof(float) * static_cast<size_t>(grid) * BLOCK);
  const PhiloxKeys key = make_philox_keys(7);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(a);
    variant_kernel<ROUNDS, MODE, UNROLL, PATHS, BLOCK, MINB><<<grid, BLOCK>>>(out, ppt, key);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  const double steps = static_cast<double>(grid) * BLOCK * ppt * T;
  printf("%-44s regs=%3d occ=%2d grid=%6d ppt=%3lld  %.3f ms  %.3e path-steps/s  %s\n", name, fa.numRegs, occ, grid,
         (long long)ppt, best, steps / (best * 1e-3), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs=%d\n", sms);
  //   ROUNDS MODE UNROLL PATHS BLOCK MINB
  run<10, 0, 2, 1, 256, 1>("base: unroll2 1path b256", sms);
  run<10, 0, 1, 1, 256, 1>("unroll1", sms);
  run<10, 0, 4, 1, 256, 1>("unroll4", sms);
  run<10, 0, 2, 1, 256, 8>("unroll2 minb8 (<=32 regs)", sms);
  run<10, 0, 1, 1, 256, 8>("unroll1 minb8", sms);
  run<10, 0, 1, 2, 256, 1>("2 paths interleaved, unroll1", sms);
  run<10, 0, 2, 2, 256, 1>("2 paths interleaved, unroll2", sms);
  run<10, 0, 1, 4, 256, 1>("4 paths interleaved, unroll1", sms);
  run<10, 0, 2, 1, 128, 1>("block128 unroll2", sms);
  run<10, 0, 2, 1, 512, 1>("block512 unroll2", sms);
  run<10, 1, 2, 1, 256, 1>("no refinement branch", sms);
  run<10, 2, 2, 1, 256, 1>("philox only", sms);
  run<10, 3, 2, 1, 256, 1>("box-muller only", sms);
  runh<1, 256, 1>("post-hoc refinement, unroll1");
  runh<2, 256, 1>("post-hoc refinement, unroll2");
  runh<1, 256, 5>("post-hoc refinement, unroll1 minb5");
  runh<1, 256, 6>("post-hoc refinement, unroll1 minb6");
  runh<1, 128, 1>("post-hoc refinement, block128");
  runp<1, 1, 256, 1>("product normals6, 1 path, unroll1");
  runp<2, 1, 256, 1>("product normals6, 1 path, unroll2");
  runp<1, 2, 256, 1>("product normals6, 2 paths interleaved");
  runp<1, 1, 128, 1>("product normals6, block128");
  runp<1, 1, 256, 6>("product normals6, minb6 (<=40 regs)");
  run6<0, 1, 256, 1>("6 normals/block (21-bit), unroll1");
  run6<0, 2, 256, 1>("6 normals/block (21-bit), unroll2");
  run6<1, 2, 256, 1>("6 normals/block, no refinement");
  run<7, 0, 2, 1, 256, 1>("philox-7 full", sms);
  run<7, 2, 2, 1, 256, 1>("philox-7 only", sms);
  return 0;
}
