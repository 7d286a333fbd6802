// Common runtime glue for the scannertools_b200 CUDA sources.
// Product builds: nvcc, sm_100a.  (STB_CPU_EMU is defined only by tests/cuda_emu, which
// compiles the same kernel source against a CPU emulation of the CUDA execution model.)
#pragma once
#ifdef STB_CPU_EMU_BUILD
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <utility>

#include "stb.h"

#ifndef STB_CPU_EMU
#include <atomic>
namespace stb { extern std::atomic<long long> g_launches; }
template <class... KArgs, class... Args>
static inline void stb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  kernel<<<grid, block, smem, s>>>(args...);
  stb::g_launches.fetch_add(1, std::memory_order_relaxed);
}
#define STB_DYN_SMEM(T, name)                                      \
  extern __shared__ __align__(16) unsigned char stb_dyn_smem_[];   \
  T* name = reinterpret_cast<T*>(stb_dyn_smem_)
#else
#define STB_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(cuda_emu::dyn_smem())
#endif

namespace stb {

constexpr int kMaxPtrBatch = 64;
// device pointers passed by value in kernel parameter space (no table upload, no sync)
template <class T>
struct PtrBatch {
  T* p[kMaxPtrBatch];
};

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define STB_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t stb_e_ = (expr);                              \
    if (stb_e_ != cudaSuccess) return stb::cuda_fail(stb_e_, #expr); \
  } while (0)

#define STB_CHECK_LAUNCH(what)                                \
  do {                                                        \
    cudaError_t stb_e_ = cudaGetLastError();                  \
    if (stb_e_ != cudaSuccess) return stb::cuda_fail(stb_e_, what); \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// The per-frame streaming kernels (histograms) run on a flattened 1-D grid of exactly one
// resident wave: `total` blocks are split over n frames as evenly as possible -- the first
// rem = total % n frames get base_blocks + 1 blocks, the others base_blocks = total / n -- so a
// batch fills every block slot whatever the frame count.  This maps a block index to
// (frame, part of that frame, number of parts of that frame).
__host__ __device__ inline void flat_grid_decode(unsigned b, unsigned base_blocks, unsigned rem, unsigned& frame,
                                                 unsigned& part, unsigned& nparts) {
  const unsigned big = rem * (base_blocks + 1u);
  if (b < big) {
    frame = b / (base_blocks + 1u);
    part = b - frame * (base_blocks + 1u);
    nparts = base_blocks + 1u;
  } else {
    const unsigned bb = b - big;
    frame = rem + bb / base_blocks;
    part = bb - (bb / base_blocks) * base_blocks;
    nparts = base_blocks;
  }
}
int num_sms();
int current_device();
int flow_hist_device(const float* const* d_flow, int n, unsigned long long npx, int32_t* d_out, cudaStream_t s,
                     bool zero_out);

}  // namespace stb
