// Shared host-side helpers for the C-ABI library (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

namespace dm {

// BatchNorm partial sums go to [group][kBnSlots][2][c] slot scratches (dm_elem.cu consumers, dm_gemm.cu epilogue)
constexpr int kBnSlots = 8;

// Thread-local last-error string returned by dm_last_error().
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

// Programmatic dependent launch (PDL): every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and starts with pdl_sync(): `griddepcontrol.wait` (the previous
// kernel of the stream has completed and its writes are visible) followed by `griddepcontrol.launch_dependents` (the
// NEXT kernel of the stream may be scheduled now: its blocks do their prologue -- barrier init, TMEM allocation,
// tensor-map prefetch, index arithmetic -- and then sit in their own wait while this kernel runs).  Nothing before
// pdl_sync() may touch global memory.  Every kernel executes the wait, so completion is transitive along the stream.
// DM_PDL=0 launches without the attribute (the instructions are then no-ops).
bool pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
#endif

#define DM_REQUIRE(cond, ...)                       \
  do {                                              \
    if (!(cond)) return dm::set_error(-1, __VA_ARGS__); \
  } while (0)

}  // namespace dm
