// Developer probe (not a product path): cost of shared-memory reductions on sm_100a as a function of the
// number of ACTIVE lanes per instruction -- decides whether merging same-bin byte pairs (a second, sparsely
// populated RED for the lanes whose two bins differ) can beat one full RED per byte in hist_rgb16_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_probe atoms_probe.cu && ./atoms_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

// mode 0: every lane, lane-private word (bank == lane), bin varies per iteration
// mode 1: only lanes with (lane % stride == 0) active
// mode 2: full RED followed by a RED with 1/stride of the lanes (the pair-merge pattern)
// mode 3: non-atomic LDS + IADD + STS on the lane-private word
// mode 4: full RED adding 2 (value in a register)
__global__ void probe(int mode, int stride, unsigned long long* cycles, unsigned* sink) {
  extern __shared__ unsigned sh[];
  for (int i = threadIdx.x; i < 48 * 32 * (blockDim.x / 32); i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sh) + (warp * 48 * 32 + lane) * 4;
  unsigned x = threadIdx.x * 2654435761u + 12345u;
  const bool act = (lane % stride) == 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < kIters; ++i) {
    x = x * 1664525u + 1013904223u;
    const unsigned a0 = base + ((x >> 9) & 0x780u);
    const unsigned a1 = base + 2048 + ((x >> 17) & 0x780u);
    if (mode == 0) {
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a0) : "memory");
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a1) : "memory");
    } else if (mode == 1) {
      if (act) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a0) : "memory");
      if (act) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a1) : "memory");
    } else if (mode == 2) {
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a0) : "memory");
      if (act) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a1) : "memory");
    } else if (mode == 3) {
      unsigned v0, v1;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v0) : "r"(a0) : "memory");
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v1) : "r"(a1) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(a0), "r"(v0 + 1) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(a1), "r"(v1 + 1) : "memory");
    } else {
      const unsigned two = 1 + (x >> 31);
      asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a0), "r"(two) : "memory");
      asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a1), "r"(two) : "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  unsigned s = 0;
  for (int i = threadIdx.x; i < 48 * 32 * (blockDim.x / 32); i += blockDim.x) s += sh[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
}

int main() {
  unsigned long long* d_c; unsigned* d_s;
  cudaMalloc(&d_c, 8 * 1024); cudaMalloc(&d_s, 4 * 1024 * 1024);
  const int threads = 384, blocks_per_sm = 3;
  const int smem = 48 * 32 * (threads / 32) * 4;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  struct { int mode, stride; const char* what; } cases[] = {
      {0, 1, "2 full REDs / iter (lane-private)"}, {1, 2, "2 REDs, 16 lanes active"}, {1, 4, "2 REDs, 8 lanes active"},
      {1, 8, "2 REDs, 4 lanes active"}, {1, 32, "2 REDs, 1 lane active"}, {2, 4, "1 full + 1 RED with 8 lanes"},
      {2, 8, "1 full + 1 RED with 4 lanes"}, {2, 32, "1 full + 1 RED with 1 lane"}, {3, 1, "2 x (LDS + STS), non-atomic"},
      {4, 1, "2 full REDs with a register operand"}};
  for (auto& c : cases) {
    probe<<<sms * blocks_per_sm, threads, smem>>>(c.mode, c.stride, d_c, d_s);
    probe<<<sms * blocks_per_sm, threads, smem>>>(c.mode, c.stride, d_c, d_s);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[8]; cudaMemcpy(h, d_c, sizeof h, cudaMemcpyDeviceToHost);
    // 3 blocks x 12 warps per SM share the pipe: cycles per (warp-level) shared op per SM
    const double ops = 2.0 * kIters * (threads / 32) * blocks_per_sm;
    printf("%-38s %8llu cycles/block  %.2f cycles per warp-level op per SM\n", c.what, h[0], (double)h[0] / ops);
  }
  return 0;
}
