// TEST INFRASTRUCTURE ONLY -- a minimal CPU emulator of the CUDA execution model, used by
// tests/ (-m "not gpu") to execute the *same kernel source* as the product on a machine
// without a GPU, so indexing/halo/tiling logic is checked against the oracle before any GPU
// time is spent.  It is never built into, linked with, or loadable by the product library
// (scannertools_b200/_lib.py only ever opens libscannertools_b200.so, which is nvcc-built).
//
// Model: one pool of blockDim threads (pthreads); blocks of the grid run one after another;
// __syncthreads() is a pthread barrier; warp collectives rendezvous per 32-thread warp.
#pragma once
#include <pthread.h>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define STB_CPU_EMU 1

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

namespace cuda_emu {
struct BlockCtx {
  pthread_barrier_t block_bar;
  std::vector<pthread_barrier_t> warp_bar;
  std::vector<uint64_t> xchg;  // [nthreads] exchange slots for warp collectives
  unsigned char* dyn_smem = nullptr;
};
extern thread_local BlockCtx* g_ctx;
extern thread_local unsigned g_linear_tid;
}  // namespace cuda_emu

extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern thread_local dim3 blockDim;
extern thread_local dim3 gridDim;

static inline void __syncthreads() { pthread_barrier_wait(&cuda_emu::g_ctx->block_bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  pthread_barrier_wait(&cuda_emu::g_ctx->warp_bar[cuda_emu::g_linear_tid / 32]);
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

namespace cuda_emu {
static inline uint64_t warp_xchg_read(uint64_t mine, unsigned src_lane) {
  BlockCtx* c = g_ctx;
  unsigned w = g_linear_tid / 32;
  c->xchg[g_linear_tid] = mine;
  pthread_barrier_wait(&c->warp_bar[w]);
  uint64_t v = c->xchg[w * 32 + (src_lane & 31)];
  pthread_barrier_wait(&c->warp_bar[w]);
  return v;
}
template <class T> static inline uint64_t to_bits(T v) { uint64_t b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
}  // namespace cuda_emu

template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
  return cuda_emu::from_bits<T>(cuda_emu::warp_xchg_read(cuda_emu::to_bits(v), (unsigned)src));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
  return cuda_emu::from_bits<T>(cuda_emu::warp_xchg_read(cuda_emu::to_bits(v), (cuda_emu::g_linear_tid & 31) ^ (unsigned)m));
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  unsigned lane = cuda_emu::g_linear_tid & 31;
  unsigned src = lane >= d ? lane - d : lane;
  return cuda_emu::from_bits<T>(cuda_emu::warp_xchg_read(cuda_emu::to_bits(v), src));
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  unsigned lane = cuda_emu::g_linear_tid & 31;
  unsigned src = lane + d < 32 ? lane + d : lane;
  return cuda_emu::from_bits<T>(cuda_emu::warp_xchg_read(cuda_emu::to_bits(v), src));
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned r = 0;
  cuda_emu::BlockCtx* c = cuda_emu::g_ctx;
  unsigned w = cuda_emu::g_linear_tid / 32;
  c->xchg[cuda_emu::g_linear_tid] = pred ? 1 : 0;
  pthread_barrier_wait(&c->warp_bar[w]);
  for (unsigned l = 0; l < 32; ++l) r |= (unsigned)(c->xchg[w * 32 + l] & 1) << l;
  pthread_barrier_wait(&c->warp_bar[w]);
  return r;
}
static inline unsigned __match_any_sync(unsigned, unsigned v) {
  unsigned r = 0;
  cuda_emu::BlockCtx* c = cuda_emu::g_ctx;
  unsigned w = cuda_emu::g_linear_tid / 32;
  c->xchg[cuda_emu::g_linear_tid] = v;
  pthread_barrier_wait(&c->warp_bar[w]);
  for (unsigned l = 0; l < 32; ++l) r |= (unsigned)(c->xchg[w * 32 + l] == v) << l;
  pthread_barrier_wait(&c->warp_bar[w]);
  return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }

static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline int __float2int_rd(float a) { return (int)std::floor(a); }
static inline int __float2int_rn(float a) { return (int)std::nearbyint(a); }
static inline int __float_as_int(float a) { int r; std::memcpy(&r, &a, 4); return r; }
static inline float __int_as_float(int a) { float r; std::memcpy(&r, &a, 4); return r; }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline int __double2int_rd(double a) { return (int)std::floor(a); }
static inline int __double2int_ru(double a) { return (int)std::ceil(a); }
static inline int __double2int_rn(double a) { return (int)std::nearbyint(a); }
static inline unsigned __vsub4(unsigned a, unsigned b) {
  unsigned r = 0;
  for (int k = 0; k < 4; ++k) r |= (((a >> (8 * k)) - (b >> (8 * k))) & 0xffu) << (8 * k);
  return r;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }

// ---- runtime API subset ------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline const char* cudaGetErrorString(cudaError_t e) { return e ? "cuda_emu error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, void*) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeAsync(void* p, void*) { return cudaFree(p); }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
#define cudaStreamNonBlocking 1u
#define cudaEventDisableTiming 2u
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8

namespace cuda_emu {
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
static inline unsigned char* dyn_smem() { return g_ctx->dyn_smem; }
}  // namespace cuda_emu

template <class... KArgs, class... Args>
static inline void stb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t, Args... args) {
  cuda_emu::run_grid(grid, block, smem, [&]() { kernel(args...); });
}
