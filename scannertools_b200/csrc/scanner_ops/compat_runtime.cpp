// Implementation of the compat shim's allocators (scanner_compat/scanner/util/memory.h) and a
// small C harness that drives the registered kernels through the Scanner-style interface the
// way the engine's evaluator does (build Elements, call execute(), read the output columns).
// Built ONLY when the real Scanner headers/libraries are absent; tests call the harness via
// ctypes to check the C++ host side end to end.
#include <cstring>
#include <memory>

#include "scanner/api/kernel.h"
#include "scanner/api/op.h"
#include "scanner/util/cuda.h"
#include "scanner/util/memory.h"
#include "stb.h"

namespace scanner {

u8* new_buffer(DeviceHandle device, size_t size) {
  u8* p = nullptr;
  if (device.type == DeviceType::GPU) {
    CU_CHECK(cudaSetDevice(device.id));
    CU_CHECK(cudaMalloc((void**)&p, size ? size : 1));
  } else {
    p = static_cast<u8*>(malloc(size ? size : 1));
  }
  return p;
}

void delete_buffer(DeviceHandle device, u8* buffer) {
  if (!buffer) return;
  if (device.type == DeviceType::GPU) {
    cudaSetDevice(device.id);
    cudaFree(buffer);
  } else {
    free(buffer);
  }
}

u8* new_block_buffer(DeviceHandle device, size_t size, i32 /*refs*/) { return new_buffer(device, size); }

Frame* new_frame(DeviceHandle device, FrameInfo info) { return new Frame(info, new_buffer(device, info.size())); }

std::vector<Frame*> new_frames(DeviceHandle device, FrameInfo info, i32 num) {
  // one block for the whole batch, like Scanner's block allocator
  std::vector<Frame*> frames;
  const size_t stride = (info.size() + 255) & ~(size_t)255;
  u8* block = new_buffer(device, stride * (size_t)(num > 0 ? num : 1));
  for (i32 i = 0; i < num; ++i) frames.push_back(new Frame(info, block + (size_t)i * stride));
  return frames;
}

void memcpy_buffer(u8* dst, DeviceHandle dst_device, const u8* src, DeviceHandle src_device, size_t size) {
  if (dst_device.type == DeviceType::CPU && src_device.type == DeviceType::CPU) { memcpy(dst, src, size); return; }
  CU_CHECK(cudaMemcpy(dst, src, size, cudaMemcpyDefault));
}

}  // namespace scanner

using namespace scanner;

namespace {

BaseKernel* make_kernel(const char* op, int device_id) {
  const KernelInfo* k = Registry::get().find_kernel(op, DeviceType::GPU);
  if (!k) return nullptr;
  KernelConfig cfg;
  cfg.devices.push_back(DeviceHandle{DeviceType::GPU, device_id});
  return k->factory(cfg);
}

struct DeviceFrames {
  DeviceHandle dev;
  std::vector<Frame*> frames;
  u8* block = nullptr;
  DeviceFrames(DeviceHandle d, const u8* host, int n, FrameInfo info) : dev(d) {
    const size_t stride = (info.size() + 255) & ~(size_t)255;
    block = new_buffer(dev, stride * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
      CU_CHECK(cudaMemcpy(block + (size_t)i * stride, host + (size_t)i * info.size(), info.size(), cudaMemcpyHostToDevice));
      frames.push_back(new Frame(info, block + (size_t)i * stride));
    }
  }
  ~DeviceFrames() {
    for (Frame* f : frames) delete f;
    delete_buffer(dev, block);
  }
};

}  // namespace

extern "C" {

#define SHIM_API __attribute__((visibility("default")))

// number of ops / kernels the static initialisers registered, and a name check
SHIM_API int stb_shim_registered(const char* op, int* is_batched, int* stencil_lo, int* stencil_hi) {
  const auto& reg = Registry::get();
  auto it = reg.ops.find(op);
  const KernelInfo* k = reg.find_kernel(op, DeviceType::GPU);
  if (it == reg.ops.end() || !k) return 0;
  if (is_batched) *is_batched = k->batched ? 1 : 0;
  if (stencil_lo) *stencil_lo = it->second.stencil.empty() ? 0 : it->second.stencil.front();
  if (stencil_hi) *stencil_hi = it->second.stencil.empty() ? 0 : it->second.stencil.back();
  return 1;
}

// Histogram: n host RGB frames -> n*48 int32 (through HistogramKernelGPU::execute)
SHIM_API int stb_shim_histogram(const uint8_t* h_frames, int n, int w, int h, int32_t* h_out, int device_id) {
  std::unique_ptr<BaseKernel> base(make_kernel("Histogram", device_id));
  auto* k = dynamic_cast<BatchedKernel*>(base.get());
  if (!k) return -1;
  DeviceHandle dev{DeviceType::GPU, device_id};
  DeviceFrames in(dev, h_frames, n, FrameInfo(h, w, 3, FrameType::U8));
  BatchedElements input(1), output(1);
  for (Frame* f : in.frames) input[0].push_back(Element(f));
  k->execute(input, output);
  if ((int)output[0].size() != n) return -2;
  for (int i = 0; i < n; ++i) {
    if (output[0][i].size != 192) return -3;
    CU_CHECK(cudaMemcpy(h_out + (size_t)i * 48, output[0][i].buffer, 192, cudaMemcpyDeviceToHost));
  }
  delete_buffer(dev, output[0][0].buffer);   // the block buffer starts at element 0
  return 0;
}

// OpticalFlow: n+1 host RGB frames -> n flow frames, via the {0,1} stencil batch layout
SHIM_API int stb_shim_optical_flow(const uint8_t* h_frames, int n, int w, int h, float* h_flow, int device_id) {
  std::unique_ptr<BaseKernel> base(make_kernel("OpticalFlow", device_id));
  auto* k = dynamic_cast<StenciledBatchedKernel*>(base.get());
  if (!k) return -1;
  DeviceHandle dev{DeviceType::GPU, device_id};
  DeviceFrames in(dev, h_frames, n + 1, FrameInfo(h, w, 3, FrameType::U8));
  StenciledBatchedElements input(1);
  BatchedElements output(1);
  for (int i = 0; i < n; ++i) input[0].push_back(Elements{Element(in.frames[i]), Element(in.frames[i + 1])});
  k->execute(input, output);
  if ((int)output[0].size() != n) return -2;
  const size_t fbytes = (size_t)w * h * 2 * sizeof(float);
  for (int i = 0; i < n; ++i) {
    Frame* f = output[0][i].as_frame();
    if (f->as_frame_info() != FrameInfo(h, w, 2, FrameType::F32)) return -3;
    CU_CHECK(cudaMemcpy(h_flow + (size_t)i * (fbytes / sizeof(float)), f->data, fbytes, cudaMemcpyDeviceToHost));
  }
  u8* block = output[0][0].as_frame()->data;
  for (int i = 0; i < n; ++i) delete output[0][i].as_frame();
  delete_buffer(dev, block);
  return 0;
}

// FlowHistogram: n host flow frames -> n*128 int32
SHIM_API int stb_shim_flow_histogram(const float* h_flow, int n, int w, int h, int32_t* h_out, int device_id) {
  std::unique_ptr<BaseKernel> base(make_kernel("FlowHistogram", device_id));
  auto* k = dynamic_cast<BatchedKernel*>(base.get());
  if (!k) return -1;
  DeviceHandle dev{DeviceType::GPU, device_id};
  DeviceFrames in(dev, reinterpret_cast<const uint8_t*>(h_flow), n, FrameInfo(h, w, 2, FrameType::F32));
  BatchedElements input(1), output(1);
  for (Frame* f : in.frames) input[0].push_back(Element(f));
  k->execute(input, output);
  if ((int)output[0].size() != n) return -2;
  for (int i = 0; i < n; ++i) {
    if (output[0][i].size != 512) return -3;
    CU_CHECK(cudaMemcpy(h_out + (size_t)i * 128, output[0][i].buffer, 512, cudaMemcpyDeviceToHost));
  }
  delete_buffer(dev, output[0][0].buffer);
  return 0;
}

// FrameDifference: (prev, cur) host frames -> one host frame
SHIM_API int stb_shim_frame_difference(const uint8_t* h_prev, const uint8_t* h_cur, int w, int h, int c, uint8_t* h_out,
                                       int device_id) {
  std::unique_ptr<BaseKernel> base(make_kernel("FrameDifference", device_id));
  auto* k = dynamic_cast<StenciledKernel*>(base.get());
  if (!k) return -1;
  Result r;
  k->validate(&r);
  if (!r.success()) return -4;
  DeviceHandle dev{DeviceType::GPU, device_id};
  FrameInfo info(h, w, c, FrameType::U8);
  DeviceFrames a(dev, h_prev, 1, info), b(dev, h_cur, 1, info);
  StenciledElements input(1);
  Elements output(1);                 // the engine pre-sizes one Element per output column
  input[0].push_back(Element(a.frames[0]));
  input[0].push_back(Element(b.frames[0]));
  k->execute(input, output);
  if (output.size() != 1 || output[0].is_null()) return -2;
  Frame* f = output[0].as_frame();
  CU_CHECK(cudaMemcpy(h_out, f->data, info.size(), cudaMemcpyDeviceToHost));
  delete_buffer(dev, f->data);
  delete f;
  return 0;
}

}  // extern "C"
