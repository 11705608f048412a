// Internal helpers shared by the translation units of libspectralmc_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "spectralmc_b200.h"

// ---- stream variants ------------------------------------------------------------------------------
// The translation units that draw normals (smc_normals.cu, smc_cf.cu, smc_diag.cu) are compiled TWICE:
// once as they are (Philox4x32-10, namespace smc, the exported smc_* entry points) and once with
// -DSMC_STREAM_P7 (Philox4x32-7, namespace smc_p7, three internal entry points smc_p7_*) — the round count has
// to be a compile-time constant for the round keys to stay constant-bank operands, and a second namespace
// keeps the two sets of kernels apart at link time.  The exported entry points dispatch on
// `stream_version` (SMC_STREAM_PHILOX10 / SMC_STREAM_PHILOX7).  smc_device.cuh follows SMC_NS.
#ifdef SMC_STREAM_P7
#define SMC_NS smc_p7
#define SMC_PHILOX_ROUNDS 7
#else
#define SMC_NS smc
#endif

#define SMC_INTERNAL __attribute__((visibility("hidden")))
extern "C" {
SMC_INTERNAL int smc_p7_philox_normals(void* out, int64_t rows, int64_t cols, int dtype, uint64_t seed, uint64_t matrix_index, void* stream);
SMC_INTERNAL int smc_p7_cf_fused(const smc_fused_args* args, const smc_p2p_group* group, void* cf_out, void* workspace, size_t workspace_bytes,
                    void* stream);
SMC_INTERNAL int smc_p7_fused_terminal(const smc_fused_args* args, void* terminal, double* terminal_sum, void* workspace,
                          size_t workspace_bytes, void* stream);
SMC_INTERNAL int smc_p7_diag_stream_fields_f32(uint64_t seed, uint64_t matrix_index, uint64_t n_blocks, uint32_t cols, uint32_t* radius_hist,
                                  uint32_t* angle_hist, uint64_t* tails4, double* power_sums4, void* stream);
SMC_INTERNAL int smc_p7_diag_stream_lags_f32(uint64_t seed, uint64_t matrix_index, uint32_t cols, uint32_t rows, double* sums7, void* stream);
}

namespace smc {

// thread-local message behind smc_last_error()
int set_error(int code, const char* fmt, ...);
void clear_error();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

inline size_t real_size(int dtype) { return dtype == SMC_F64 ? 8 : 4; }

// number of SMs of the current device, cached
int sm_count();

// SMC_CF_ROW_FFT (smc_rowfft.cu)
bool rowfft_supported(int64_t n);
size_t rowfft_workspace_bytes(int64_t batches, int64_t n);
int fft_rows(const void* mat, int64_t batches, int64_t n, int dtype, void* out, cudaStream_t st);
int rowfft_mean(const void* mat, int64_t batches, int64_t n, int dtype, void* out, void* ws, size_t ws_bytes,
                cudaStream_t st);

}  // namespace smc

#define SMC_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return ::smc::set_error(SMC_EINVAL, __VA_ARGS__); \
  } while (0)

#define SMC_CUDA_OK(expr)                                                                 \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::smc::set_error(SMC_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                        \
  } while (0)

// after a kernel launch: catches bad launch configurations immediately; execution faults are
// sticky and surface at the caller's next synchronisation (or the next call into the library)
#define SMC_LAUNCH_OK(name)                                                                   \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess)                                                                    \
      return ::smc::set_error(SMC_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)
