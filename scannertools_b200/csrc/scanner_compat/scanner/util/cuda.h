// Scanner-API look-alike: CUDA error helpers (optical_flow_kernel_gpu.cpp:97,
// histogram_kernel_gpu.cpp:68).
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CUDA_PROTECT(s) s

#define CU_CHECK(ans) \
  { ::scanner::cu_assert((ans), __FILE__, __LINE__); }

namespace scanner {
inline void cu_assert(cudaError_t code, const char* file, int line) {
  if (code != cudaSuccess) {
    fprintf(stderr, "CUDA error: %s (%s:%d)\n", cudaGetErrorString(code), file, line);
    abort();   // the reference aborts the worker on CUDA errors; no exception crosses the boundary
  }
}
}  // namespace scanner
