// Internal helpers shared by the translation units of libspectralmc_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "spectralmc_b200.h"

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
