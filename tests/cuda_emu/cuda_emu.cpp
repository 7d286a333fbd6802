// TEST INFRASTRUCTURE ONLY -- see cuda_emu.h.
#include "cuda_emu.h"

thread_local uint3 threadIdx;
thread_local uint3 blockIdx;
thread_local dim3 blockDim;
thread_local dim3 gridDim;

namespace cuda_emu {
thread_local BlockCtx* g_ctx = nullptr;
thread_local unsigned g_linear_tid = 0;

namespace {
struct Job {
  BlockCtx* ctx;
  dim3 grid, block;
  const std::function<void()>* body;
  unsigned tid;
};

void* worker(void* arg) {
  Job* j = (Job*)arg;
  g_ctx = j->ctx;
  g_linear_tid = j->tid;
  blockDim = j->block;
  gridDim = j->grid;
  threadIdx.x = j->tid % j->block.x;
  threadIdx.y = (j->tid / j->block.x) % j->block.y;
  threadIdx.z = j->tid / (j->block.x * j->block.y);
  for (unsigned bz = 0; bz < j->grid.z; ++bz)
    for (unsigned by = 0; by < j->grid.y; ++by)
      for (unsigned bx = 0; bx < j->grid.x; ++bx) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        (*j->body)();
        pthread_barrier_wait(&j->ctx->block_bar);  // block boundary: smem reuse is safe
      }
  return nullptr;
}
}  // namespace

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  unsigned nt = block.x * block.y * block.z;
  if (nt == 0 || grid.x * grid.y * grid.z == 0) return;
  BlockCtx ctx;
  pthread_barrier_init(&ctx.block_bar, nullptr, nt);
  unsigned nwarps = (nt + 31) / 32;
  ctx.warp_bar.resize(nwarps);
  for (unsigned w = 0; w < nwarps; ++w) {
    unsigned cnt = (w + 1) * 32 <= nt ? 32 : nt - w * 32;
    pthread_barrier_init(&ctx.warp_bar[w], nullptr, cnt);
  }
  ctx.xchg.assign(nwarps * 32, 0);
  std::vector<unsigned char> dyn(smem + 64);
  ctx.dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
  std::vector<Job> jobs(nt);
  std::vector<pthread_t> th(nt);
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 256 * 1024);
  for (unsigned t = 0; t < nt; ++t) {
    jobs[t] = Job{&ctx, grid, block, &body, t};
    if (pthread_create(&th[t], &attr, worker, &jobs[t]) != 0) { std::perror("cuda_emu pthread_create"); std::abort(); }
  }
  for (unsigned t = 0; t < nt; ++t) pthread_join(th[t], nullptr);
  pthread_attr_destroy(&attr);
  for (unsigned w = 0; w < nwarps; ++w) pthread_barrier_destroy(&ctx.warp_bar[w]);
  pthread_barrier_destroy(&ctx.block_bar);
}
}  // namespace cuda_emu
