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

#define DM_REQUIRE(cond, ...)                       \
  do {                                              \
    if (!(cond)) return dm::set_error(-1, __VA_ARGS__); \
  } while (0)

}  // namespace dm
