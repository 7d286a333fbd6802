// Developer probe: minimal 3-D TMA tile load with the same descriptor / PTX as iter15_tma_kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int BW = 64, BH = 46;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int x0, int y0, int z, int* status) {
  __shared__ __align__(128) float tile[BW * BH];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(tile)),
                 "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x0), "r"(y0), "r"(z), "r"(smem_u32(&bar))
                 : "memory");
  }
  unsigned done = 0;
  for (unsigned spin = 0; spin < (1u << 22) && !done; ++spin) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  if (threadIdx.x == 0) *status = done ? 1 : -1;
  if (done)
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}

int main() {
  const int w = 240, h = 30, planes = 10;
  std::vector<float> host((size_t)w * h * planes);
  for (size_t i = 0; i < host.size(); ++i) host[i] = (float)i;
  float *d, *dout; int* dstat;
  cudaMalloc(&d, host.size() * 4); cudaMalloc(&dout, BW * BH * 4); cudaMalloc(&dstat, 4);
  cudaMemcpy(d, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  printf("entry point: %d %d %p\n", (int)e, (int)q, sym);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Fn enc = (Fn)sym;
  for (int boxh : {46, 30}) {
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4};
    cuuint32_t box[3] = {BW, (cuuint32_t)boxh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("boxh %d encode result %d\n", boxh, (int)r);
    if (r != CUDA_SUCCESS) continue;
    if (boxh != BH) continue;
    const int xs[] = {8, -8, 8, -8, 200, 232, 236, 8};
    const int ys[] = {0, 0, -7, -7, 0, 3, -7, 20};
    for (int t = 0; t < 8; ++t) {
      int x0 = xs[t], y0 = ys[t], z = 3;
      probe<<<1, 128>>>(m, dout, x0, y0, z, dstat);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("probe x0=%d y0=%d: sync=%d (%s)\n", x0, y0, (int)e, cudaGetErrorString(e)); cudaDeviceReset(); return 0; }
      int st = 0; cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost);
      std::vector<float> o(BW * BH);
      cudaMemcpy(o.data(), dout, BW * BH * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r2 = 0; r2 < BH; ++r2) for (int c = 0; c < BW; ++c) {
        int x = x0 + c, y = y0 + r2;
        float exp = (x >= 0 && x < w && y >= 0 && y < h) ? host[((size_t)z * h + y) * w + x] : 0.f;
        if (o[r2 * BW + c] != exp) ++bad;
      }
      printf("probe x0=%d y0=%d: sync=%d (%s) status=%d mismatches=%d\n", x0, y0, (int)e, cudaGetErrorString(e), st, bad);
    }
  }
  return 0;
}
