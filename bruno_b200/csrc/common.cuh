// Shared host-side plumbing for the C-ABI: error codes, last-error string, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bruno_sm100.h"

namespace bruno {

void set_last_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return BRUNO_ECUDA;
  }
  return BRUNO_OK;
}

#define BRUNO_REQUIRE(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      bruno::set_last_error(__VA_ARGS__); \
      return BRUNO_EINVAL;            \
    }                                 \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace bruno
