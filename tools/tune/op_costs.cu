// Issue-cost table for the fused float32 kernel's instruction forms (B200): SM cycles per warp
// instruction per SM sub-partition when 16 independent instances of ONE instruction form are issued
// per loop iteration by 10 warps per sub-partition (throughput, not latency).  CUDA events, 1965 MHz.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define OP_LOOP(NAME, DECL, BODY, SINK)                                                         \
  __global__ void __launch_bounds__(256) k_##NAME(int iters, unsigned* sink, unsigned kparam) { \
    DECL;                                                                                        \
    _Pragma("unroll 1") for (int it = 0; it < iters; ++it) {                                    \
      _Pragma("unroll") for (int k = 0; k < 16; ++k) { BODY; }                                  \
    }                                                                                            \
    unsigned s = 0;                                                                              \
    _Pragma("unroll") for (int k = 0; k < 16; ++k) s ^= SINK;                                   \
    if (s == 0x12345u) sink[0] = s;                                                              \
  }

__device__ __forceinline__ void op_wide(uint32_t& a, uint32_t& b) {
  unsigned long long p;
  asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p) : "r"(a));
  asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(p));
}
__device__ __forceinline__ void op_ffma2(float& a0, float& a1, float b0, float b1) {
  unsigned long long x;
  unsigned long long y;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a0), "f"(a1));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b0), "f"(b1));
  asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x) : "l"(y));
  asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(x));
}
__device__ __forceinline__ void op_i2f(uint32_t& a) {
  float f;
  asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(a));
  a = __float_as_uint(f);
}

#define UDECL uint32_t a[16], b[16]; for (int k = 0; k < 16; ++k) { a[k] = threadIdx.x * 2654435761u + k; b[k] = a[k] * 40503u + 7u; }
#define FDECL float a[16], b[16]; for (int k = 0; k < 16; ++k) { a[k] = 1.0f + 0.001f * (threadIdx.x + k); b[k] = 0.5f + 0.002f * k; }

OP_LOOP(lop3_rrr, UDECL, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b[k]), "r"(b[(k + 1) & 15])), a[k])
OP_LOOP(lop3_rrc, UDECL, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b[k]), "r"(kparam)), a[k])
OP_LOOP(lop3_rii, UDECL, asm volatile("lop3.b32 %0, %0, 0x007ffffc, 0x3f800000, 0xea;" : "+r"(a[k])), a[k])
OP_LOOP(shf_r, UDECL, asm volatile("shr.u32 %0, %0, 1;" : "+r"(a[k])); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[k]) : "r"(a[k])), b[k])
OP_LOOP(iadd, UDECL, asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b[k])), a[k])
OP_LOOP(shf_funnel, UDECL, asm volatile("shf.l.wrap.b32 %0, %0, %1, 12;" : "+r"(a[k]) : "r"(b[k])), a[k])
OP_LOOP(imad_wide, UDECL, op_wide(a[k], b[k]), (a[k] ^ b[k]))
OP_LOOP(imad_hi, UDECL, asm volatile("mad.hi.u32 %0, %0, 0x00800001, 0x3f800000;" : "+r"(a[k])), a[k])
OP_LOOP(imad_lo, UDECL, asm volatile("mad.lo.u32 %0, %0, 0x9E3779B9, %1;" : "+r"(a[k]) : "r"(b[k])), a[k])
OP_LOOP(imad_shl, UDECL, asm volatile("shl.b32 %0, %0, 10;" : "+r"(a[k])); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[k]) : "r"(a[k])), b[k])
OP_LOOP(vimnmx, UDECL, asm volatile("min.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b[k])), a[k])
OP_LOOP(prmt, UDECL, asm volatile("prmt.b32 %0, %0, %1, 0x3215;" : "+r"(a[k]) : "r"(b[k])), a[k])
OP_LOOP(ffma_rrr, FDECL, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(b[(k + 1) & 15])), __float_as_uint(a[k]))
OP_LOOP(ffma_rri, FDECL, asm volatile("fma.rn.f32 %0, %0, %1, 0f3E800000;" : "+f"(a[k]) : "f"(b[k])), __float_as_uint(a[k]))
OP_LOOP(ffma_rii, FDECL, asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f3E800000;" : "+f"(a[k])), __float_as_uint(a[k]))
OP_LOOP(fadd_ri, FDECL, asm volatile("add.rn.f32 %0, %0, 0f3E800000;" : "+f"(a[k])), __float_as_uint(a[k]))
OP_LOOP(fmul_rz, FDECL, asm volatile("mul.rz.f32 %0, %0, 0f3F7FFFFF;" : "+f"(a[k])), __float_as_uint(a[k]))
OP_LOOP(ffma2, FDECL, if (k < 8) op_ffma2(a[k], a[k + 8], b[k], b[k + 8]), __float_as_uint(a[k]))
OP_LOOP(mufu_ex2, FDECL, asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k])), __float_as_uint(a[k]))
OP_LOOP(i2f, UDECL, op_i2f(a[k]), a[k])

template <typename K>
void run(const char* name, K kern, unsigned* sink, double per_iter) {
  const int iters = 20000;
  kern<<<148 * 5, 256>>>(iters, sink, 0x9e3779b9u);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<148 * 5, 256>>>(iters, sink, 0x9e3779b9u);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-12s %6.2f cycles per warp-instruction per SMSP (%.0f per iteration; loop overhead ~3 instr not subtracted)\n", name,
         ms * 1.965e6 / iters / 10.0 / per_iter, per_iter);
}

int main() {
  unsigned* sink;
  cudaMalloc(&sink, 16);
  run("lop3 r,r,r", k_lop3_rrr, sink, 16);
  run("lop3 r,r,c[]", k_lop3_rrc, sink, 16);
  run("lop3 r,i,i", k_lop3_rii, sink, 16);
  run("shr+iadd", k_shf_r, sink, 32);
  run("iadd", k_iadd, sink, 16);
  run("shf funnel", k_shf_funnel, sink, 16);
  run("imad.wide", k_imad_wide, sink, 16);
  run("imad.hi", k_imad_hi, sink, 16);
  run("imad.lo", k_imad_lo, sink, 16);
  run("shl+iadd", k_imad_shl, sink, 32);
  run("min.u32", k_vimnmx, sink, 16);
  run("prmt", k_prmt, sink, 16);
  run("ffma r,r,r", k_ffma_rrr, sink, 16);
  run("ffma r,r,i", k_ffma_rri, sink, 16);
  run("ffma r,i,i", k_ffma_rii, sink, 16);
  run("fadd r,i", k_fadd_ri, sink, 16);
  run("fmul.rz r,i", k_fmul_rz, sink, 16);
  run("ffma2 (x8)", k_ffma2, sink, 8);
  run("mufu.ex2", k_mufu_ex2, sink, 16);
  run("i2f.u32", k_i2f, sink, 16);
  return 0;
}
