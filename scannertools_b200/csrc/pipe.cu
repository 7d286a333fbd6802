// Host-buffer entry points (include/stb.h "stb_pipe_*"): the end-to-end path when this library
// is driven with HOST frames instead of the device frames Scanner's engine hands its GPU
// kernels.  Owns two device frame slots, a copy stream and a compute stream; the H2D copy of
// batch c+1 overlaps the kernels of batch c; per-frame results (192 B / 512 B) are copied
// back once per call; flow frames (when requested) stream back per batch on a third stream.
// Pass page-locked host memory for the copies to be truly asynchronous.
#include "stb_rt.h"

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace stb;

struct stb_pipe {
  int W, H, B, want_flow, device;
  size_t frame_bytes;
  cudaStream_t s_copy, s_comp, s_out;
  cudaEvent_t copied[2], comp_done[2], out_done[2];
  cudaEvent_t call_done[2];   // completion of the two most recent asynchronous calls
  long long calls;            // asynchronous calls submitted so far
  uint8_t* d_frames[2];   // [B+1] frames each
  float* d_flow[2];       // [B] flow frames each (only when flow frames are returned)
  int32_t* d_res;         // per-frame results for the whole call (hist or flow-hist)
  int32_t* d_S;
  size_t res_cap_frames;
  stb_farneback* fb;
};

namespace {

int ensure_results(stb_pipe* p, size_t n) {
  if (n <= p->res_cap_frames) return STB_OK;
  if (p->d_res) cudaFree(p->d_res);
  if (p->d_S) cudaFree(p->d_S);
  p->d_res = nullptr; p->d_S = nullptr; p->res_cap_frames = 0;
  size_t cap = n < 1024 ? 1024 : n;
  cudaError_t e = cudaMalloc((void**)&p->d_res, cap * STB_FLOWHIST_INTS * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_S, cap * sizeof(int32_t));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("stb_pipe: result buffer allocation failed: %s", cudaGetErrorString(e));
    return STB_ERR_ALLOC;
  }
  p->res_cap_frames = cap;
  return STB_OK;
}

}  // namespace

extern "C" {

int stb_pipe_create(int width, int height, int max_batch, int want_flow, stb_pipe** out) {
  if (!out) { set_error("stb_pipe_create: out is NULL"); return STB_ERR_INVALID; }
  *out = nullptr;
  if (width <= 0 || height <= 0 || max_batch <= 0 || max_batch > 4096) {
    set_error("stb_pipe_create: invalid argument (%dx%d, max_batch=%d)", width, height, max_batch);
    return STB_ERR_INVALID;
  }
  stb_pipe* p = new (std::nothrow) stb_pipe();
  if (!p) { set_error("stb_pipe_create: out of host memory"); return STB_ERR_ALLOC; }
  std::memset(p, 0, sizeof(*p));
  p->W = width; p->H = height; p->B = max_batch; p->want_flow = want_flow;
  p->device = current_device();
  p->frame_bytes = (size_t)3 * width * height;
  // keep each frame 16-byte aligned inside the slot so the vector paths are taken
  const size_t stride = (p->frame_bytes + 15) & ~(size_t)15;
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc((void**)&p->d_frames[i], stride * (size_t)(max_batch + 1));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_copy, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_comp, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&p->copied[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->comp_done[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->out_done[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->call_done[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    int rc = cuda_fail(e, "stb_pipe_create");
    stb_pipe_destroy(p);
    return rc;
  }
  if (want_flow) {
    int rc = stb_farneback_create(width, height, max_batch, nullptr, &p->fb);
    if (rc) { stb_pipe_destroy(p); return rc; }
  }
  *out = p;
  return STB_OK;
}

int stb_pipe_destroy(stb_pipe* p) {
  if (!p) return STB_OK;
  if (p->s_comp) cudaStreamSynchronize(p->s_comp);
  if (p->s_copy) cudaStreamSynchronize(p->s_copy);
  if (p->s_out) cudaStreamSynchronize(p->s_out);
  if (p->fb) stb_farneback_destroy(p->fb);
  for (int i = 0; i < 2; ++i) {
    if (p->d_frames[i]) cudaFree(p->d_frames[i]);
    if (p->d_flow[i]) cudaFree(p->d_flow[i]);
    if (p->copied[i]) cudaEventDestroy(p->copied[i]);
    if (p->comp_done[i]) cudaEventDestroy(p->comp_done[i]);
    if (p->out_done[i]) cudaEventDestroy(p->out_done[i]);
    if (p->call_done[i]) cudaEventDestroy(p->call_done[i]);
  }
  if (p->d_res) cudaFree(p->d_res);
  if (p->d_S) cudaFree(p->d_S);
  if (p->s_copy) cudaStreamDestroy(p->s_copy);
  if (p->s_comp) cudaStreamDestroy(p->s_comp);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
  return STB_OK;
}

int stb_pipe_hist(stb_pipe* p, const uint8_t* h_frames, int n, int32_t* h_hist, int32_t* h_S) {
  if (!p || n < 0 || (n > 0 && (!h_frames || !h_hist))) { set_error("stb_pipe_hist: invalid argument"); return STB_ERR_INVALID; }
  if (n == 0) return STB_OK;
  int rc = ensure_results(p, (size_t)n);
  if (rc) return rc;
  const size_t fbytes = p->frame_bytes;
  const size_t stride = (fbytes + 15) & ~(size_t)15;
  const int B = p->B + 1;  // the slot holds B+1 frames; the histogram path can use all of them
  int c = 0;
  for (int f0 = 0; f0 < n; f0 += B, ++c) {
    const int m = n - f0 < B ? n - f0 : B;
    const int slot = c & 1;
    STB_CUDA(cudaStreamWaitEvent(p->s_copy, p->comp_done[slot], 0));
    if (stride == fbytes) {
      STB_CUDA(cudaMemcpyAsync(p->d_frames[slot], h_frames + (size_t)f0 * fbytes, (size_t)m * fbytes, cudaMemcpyHostToDevice, p->s_copy));
    } else {
      for (int i = 0; i < m; ++i)
        STB_CUDA(cudaMemcpyAsync(p->d_frames[slot] + (size_t)i * stride, h_frames + (size_t)(f0 + i) * fbytes, fbytes,
                                 cudaMemcpyHostToDevice, p->s_copy));
    }
    STB_CUDA(cudaEventRecord(p->copied[slot], p->s_copy));
    STB_CUDA(cudaStreamWaitEvent(p->s_comp, p->copied[slot], 0));
    rc = stb_hist_rgb16_strided(p->d_frames[slot], stride, m, p->W, p->H, p->d_res + (size_t)f0 * STB_HIST_INTS, p->s_comp);
    if (rc) return rc;
    STB_CUDA(cudaEventRecord(p->comp_done[slot], p->s_comp));
  }
  if (h_S) {
    rc = stb_shot_scores(p->d_res, n, nullptr, p->d_S, p->s_comp);
    if (rc) return rc;
    STB_CUDA(cudaMemcpyAsync(h_S, p->d_S, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, p->s_comp));
  }
  STB_CUDA(cudaMemcpyAsync(h_hist, p->d_res, (size_t)n * STB_HIST_INTS * sizeof(int32_t), cudaMemcpyDeviceToHost, p->s_comp));
  STB_CUDA(cudaStreamSynchronize(p->s_comp));
  return STB_OK;
}

int stb_pipe_flow_async(stb_pipe* p, const uint8_t* h_frames, int n, float* h_flow, int32_t* h_flow_hist, int* ticket) {
  if (ticket) *ticket = -1;
  if (!p || n < 0 || (n > 0 && !h_frames)) { set_error("stb_pipe_flow: invalid argument"); return STB_ERR_INVALID; }
  if (!p->fb) { set_error("stb_pipe_flow: pipe was created with want_flow = 0"); return STB_ERR_INVALID; }
  if (n == 0) return STB_OK;
  if (!h_flow && !h_flow_hist) { set_error("stb_pipe_flow: no output requested"); return STB_ERR_INVALID; }
  // is the previous asynchronous call still running?  (then its kernels hide this call's first upload)
  const bool idle = p->calls == 0 || cudaEventQuery(p->call_done[(p->calls - 1) & 1]) == cudaSuccess;
  (void)cudaGetLastError();
  if ((size_t)n > p->res_cap_frames) {   // growing the result buffers must not race with pending work
    STB_CUDA(cudaStreamSynchronize(p->s_comp));
    STB_CUDA(cudaStreamSynchronize(p->s_out));
  }
  int rc = ensure_results(p, (size_t)n);
  if (rc) return rc;
  const size_t fbytes = p->frame_bytes;
  const size_t stride = (fbytes + 15) & ~(size_t)15;
  const size_t flow_bytes = (size_t)p->W * p->H * 2 * sizeof(float);
  if (h_flow) {
    for (int i = 0; i < 2; ++i)
      if (!p->d_flow[i]) {
        cudaError_t e = cudaMalloc((void**)&p->d_flow[i], flow_bytes * (size_t)p->B);
        if (e != cudaSuccess) { (void)cudaGetLastError(); set_error("stb_pipe_flow: flow slot allocation failed"); return STB_ERR_ALLOC; }
      }
  }
  const int B = p->B;
  std::vector<const uint8_t*> fr((size_t)B + 1);
  std::vector<float*> fl((size_t)B);
  int c = 0, prev_m = 0;
  for (int p0 = 0; p0 < n; ++c) {
    // The first upload of a call cannot overlap any compute, and batch c+1's upload only hides
    // behind batch c's kernels if batch c is not much smaller: start at an eighth of a batch and
    // double (B/8, B/4, B/2, B, B, ...), so the exposed copy time is one small upload.
    int cap = B;
    if (idle && n > B && c < 3) {
      cap = B >> (3 - c);
      if (cap < 1) cap = 1;
    }
    const int m = n - p0 < cap ? n - p0 : cap;
    const int slot = c & 1;
    STB_CUDA(cudaStreamWaitEvent(p->s_copy, p->comp_done[slot], 0));
    int first = 0;
    if (c > 0) {
      // frame p0 is the last frame of the previous batch: device-to-device, not over PCIe again
      STB_CUDA(cudaMemcpyAsync(p->d_frames[slot], p->d_frames[slot ^ 1] + (size_t)prev_m * stride, fbytes, cudaMemcpyDeviceToDevice, p->s_copy));
      first = 1;
    }
    if (stride == fbytes) {
      STB_CUDA(cudaMemcpyAsync(p->d_frames[slot] + (size_t)first * stride, h_frames + (size_t)(p0 + first) * fbytes,
                               (size_t)(m + 1 - first) * fbytes, cudaMemcpyHostToDevice, p->s_copy));
    } else {
      for (int i = first; i <= m; ++i)
        STB_CUDA(cudaMemcpyAsync(p->d_frames[slot] + (size_t)i * stride, h_frames + (size_t)(p0 + i) * fbytes, fbytes,
                                 cudaMemcpyHostToDevice, p->s_copy));
    }
    STB_CUDA(cudaEventRecord(p->copied[slot], p->s_copy));
    STB_CUDA(cudaStreamWaitEvent(p->s_comp, p->copied[slot], 0));
    for (int i = 0; i <= m; ++i) fr[i] = p->d_frames[slot] + (size_t)i * stride;
    if (h_flow) {
      STB_CUDA(cudaStreamWaitEvent(p->s_comp, p->out_done[slot], 0));  // flow slot drained by the D2H of batch c-2
      for (int i = 0; i < m; ++i) fl[i] = p->d_flow[slot] + (size_t)i * (flow_bytes / sizeof(float));
    }
    if (h_flow_hist)
      rc = stb_farneback_run_hist(p->fb, fr.data(), m, h_flow ? fl.data() : nullptr, p->d_res + (size_t)p0 * STB_FLOWHIST_INTS, p->s_comp);
    else
      rc = stb_farneback_run(p->fb, fr.data(), m, fl.data(), p->s_comp);
    if (rc) return rc;
    STB_CUDA(cudaEventRecord(p->comp_done[slot], p->s_comp));
    if (h_flow) {
      STB_CUDA(cudaStreamWaitEvent(p->s_out, p->comp_done[slot], 0));
      STB_CUDA(cudaMemcpyAsync(h_flow + (size_t)p0 * (flow_bytes / sizeof(float)), p->d_flow[slot], (size_t)m * flow_bytes,
                               cudaMemcpyDeviceToHost, p->s_out));
      STB_CUDA(cudaEventRecord(p->out_done[slot], p->s_out));
    }
    prev_m = m;   // the next batch's frame 0 is this batch's frame m
    p0 += m;
  }
  if (h_flow_hist)
    STB_CUDA(cudaMemcpyAsync(h_flow_hist, p->d_res, (size_t)n * STB_FLOWHIST_INTS * sizeof(int32_t), cudaMemcpyDeviceToHost, p->s_comp));
  if (h_flow) STB_CUDA(cudaStreamWaitEvent(p->s_comp, p->out_done[(c - 1) & 1], 0));   // join the flow read-back stream
  const int t = (int)(p->calls & 1);
  STB_CUDA(cudaEventRecord(p->call_done[t], p->s_comp));
  p->calls++;
  if (ticket) *ticket = t;
  return STB_OK;
}

int stb_pipe_wait(stb_pipe* p, int ticket) {
  if (!p || ticket < 0 || ticket > 1) { set_error("stb_pipe_wait: invalid argument"); return STB_ERR_INVALID; }
  STB_CUDA(cudaEventSynchronize(p->call_done[ticket]));
  return STB_OK;
}

int stb_pipe_flow(stb_pipe* p, const uint8_t* h_frames, int n, float* h_flow, int32_t* h_flow_hist) {
  int ticket = -1;
  int rc = stb_pipe_flow_async(p, h_frames, n, h_flow, h_flow_hist, &ticket);
  if (rc || ticket < 0) return rc;
  return stb_pipe_wait(p, ticket);
}

int stb_host_alloc(size_t bytes, int write_combined, void** out) {
  if (!out || bytes == 0) { set_error("stb_host_alloc: invalid argument"); return STB_ERR_INVALID; }
  *out = nullptr;
#ifdef STB_CPU_EMU
  *out = std::calloc(1, bytes);
  return *out ? STB_OK : STB_ERR_ALLOC;
#else
  cudaError_t e = cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    *out = nullptr;
    set_error("stb_host_alloc: cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return ((int)e == 100 || (int)e == 35) ? STB_ERR_NO_DEVICE : STB_ERR_ALLOC;
  }
  return STB_OK;
#endif
}

int stb_host_free(void* p) {
  if (!p) return STB_OK;
#ifdef STB_CPU_EMU
  std::free(p);
  return STB_OK;
#else
  STB_CUDA(cudaFreeHost(p));
  return STB_OK;
#endif
}

}  // extern "C"
