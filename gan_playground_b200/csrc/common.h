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

// Programmatic dependent launch (PDL). Every kernel of this library starts with pdl_sync(): it lets the NEXT launch on the
// stream become resident while this grid is still running (griddepcontrol.launch_dependents) and then waits until every
// grid this one depends on has completed and flushed its memory (griddepcontrol.wait) before it touches global memory.
// Launched with the programmatic-stream-serialization attribute (launch_pdl, GP_PDL=1) a kernel's launch latency and
// prologue overlap the tail of its predecessor - in eager mode and, as programmatic edges, inside captured CUDA graphs;
// without the attribute both instructions are no-ops. OFF by default: measured on the B200 (profiles/r02_pdl_ab.txt) the
// graph-replayed steps gain nothing (cfg2 15.30 vs 14.88 ms, batch 128 3.03 vs 3.03, cfg3 3.92 vs 3.85, cfg4 17.5 vs
// 17.7): kernel-to-kernel latency inside a graph is already below what the parked CTAs cost. Ordering stays transitive because EVERY kernel waits before it
// exits. conv_gemm_kernel places the wait after its prologue (barrier init, TMEM allocation, descriptor prefetch).
bool pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_launch_dependents();
  pdl_wait();
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_attrs(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
}

// kernel<<<grid, block, smem, st>>>(args...) with the PDL attribute when enabled; errors surface through
// GP_CHECK_LAUNCH (cudaPeekAtLastError) exactly as for the chevron syntax
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  (void)launch_attrs(kernel, grid, block, smem, st, pdl_enabled(), 1u, static_cast<Args&&>(args)...);
}
// the same without the attribute (kernels that spin on other GPUs)
template <typename... KArgs, typename... Args>
inline void launch_plain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  (void)launch_attrs(kernel, grid, block, smem, st, false, 1u, static_cast<Args&&>(args)...);
}
#endif

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
