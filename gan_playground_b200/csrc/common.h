// Host-side plumbing shared by all translation units of libgpb200: error reporting, launch counting.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/gpb200.h"

namespace gp {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace gp

#define GP_CHECK_CUDA(expr)                                                                              \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return gp::set_error(GP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                                    \
  } while (0)

#define GP_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return gp::set_error(GP_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Check the launch itself (configuration errors); asynchronous faults surface at the next sync in the host.
#define GP_CHECK_LAUNCH()                  \
  do {                                     \
    GP_CHECK_CUDA(cudaPeekAtLastError());  \
    gp::count_launch();                    \
  } while (0)
