// Pairwise pipe-overlap microbenchmark for the fused float32 kernel's instruction mix (B200):
// how many SM cycles one warp-iteration costs when MUFU, IMAD.WIDE.U32, LOP3 and FFMA streams are
// issued alone and together, at the fused kernel's occupancy (40 warps / SM).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_overlap pipe_overlap.cu && ./pipe_overlap
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NM, int NI, int NL, int NF, int ND = 0>
__global__ void __launch_bounds__(256) mix(int iters, float* sink, long long* cycles) {
  float m[8];
  uint32_t a[8], b[8], l[8];
  float f[8];
  double d[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    m[k] = 0.5f + 0.01f * (threadIdx.x + k);
    a[k] = threadIdx.x * 2654435761u + k;
    b[k] = a[k] ^ 0x9e3779b9u;
    l[k] = a[k] * 3u + k;
    f[k] = 1.0f + k;
    d[k] = 1.0 + k;
  }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NM; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[k & 7]));
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      uint64_t p = static_cast<uint64_t>(a[k & 7]) * 0xD2511F53u;
      a[k & 7] = static_cast<uint32_t>(p >> 32) ^ static_cast<uint32_t>(p);  // + 1 LOP3 per product
    }
#pragma unroll
    for (int k = 0; k < NL; ++k) l[k & 7] = (l[k & 7] ^ b[(k + 1) & 7]) & (b[k & 7] | 0x55u + it);
#pragma unroll
    for (int k = 0; k < NF; ++k) f[k & 7] = fmaf(f[k & 7], 1.0000001f, 0.25f);
#pragma unroll
    for (int k = 0; k < ND; ++k) d[k & 7] = fma(d[k & 7], 1.0000001, 0.25);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += m[k] + f[k] + static_cast<float>(a[k] ^ l[k]) + static_cast<float>(d[k]);
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int NM, int NI, int NL, int NF, int ND = 0>
void run(const char* name, float* sink, long long* cyc) {
  const int iters = 20000;
  mix<NM, NI, NL, NF, ND><<<148 * 5, 256>>>(iters, sink, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  mix<NM, NI, NL, NF, ND><<<148 * 5, 256>>>(iters, sink, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  // 40 warps per SM = 10 per SMSP: cycles per warp-iteration of one SMSP
  printf("%-28s MUFU %2d IMAD.WIDE %2d LOP3 %2d(+%d) FFMA %2d DFMA %2d : %7.1f cycles / warp-iteration / SMSP\n", name, NM, NI, NL, NI, NF, ND,
         double(c) / iters / 10.0);
  printf("%-28s   by events at 1965 MHz: %7.1f cycles\n", "", ms * 1.965e6 / iters / 10.0);
}

int main() {
  float* sink;
  long long* cyc;
  cudaMalloc(&sink, 16);
  cudaMalloc(&cyc, 16);
  run<8, 0, 0, 0>("MUFU", sink, cyc);
  run<0, 11, 0, 0>("IMAD.WIDE(+LOP3)", sink, cyc);
  run<0, 0, 16, 0>("LOP3", sink, cyc);
  run<0, 0, 0, 16>("FFMA", sink, cyc);
  run<8, 11, 0, 0>("MUFU+IMAD", sink, cyc);
  run<8, 0, 16, 0>("MUFU+LOP3", sink, cyc);
  run<8, 0, 0, 16>("MUFU+FFMA", sink, cyc);
  run<0, 11, 16, 0>("IMAD+LOP3", sink, cyc);
  run<0, 11, 0, 16>("IMAD+FFMA", sink, cyc);
  run<8, 11, 16, 0>("MUFU+IMAD+LOP3", sink, cyc);
  run<8, 11, 16, 8>("MUFU+IMAD+LOP3+FFMA (kernel mix)", sink, cyc);
  run<0, 0, 0, 0, 16>("DFMA", sink, cyc);
  run<0, 0, 16, 0, 16>("DFMA+LOP3", sink, cyc);
  run<0, 8, 0, 0, 16>("DFMA+IMAD", sink, cyc);
  run<0, 0, 0, 16, 16>("DFMA+FFMA", sink, cyc);
  run<8, 0, 0, 0, 16>("DFMA+MUFU", sink, cyc);
  run<0, 0, 32, 0, 0>("LOP3 x32", sink, cyc);
  run<0, 0, 16, 16, 0>("LOP3+FFMA", sink, cyc);
  return 0;
}
