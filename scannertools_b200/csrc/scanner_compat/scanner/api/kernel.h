// Scanner-API look-alike: kernel base classes (see util/common.h).  Signatures follow the
// reference's overrides: histogram_kernel_gpu.cpp:31-32, optical_flow_kernel_cpu.cpp:27-28,
// optical_flow_kernel_gpu.cpp:45-46, blur_kernel_cpu.cpp:44, resize_kernel.cpp:28.
#pragma once
#include "scanner/api/frame.h"

namespace scanner {

struct KernelConfig {
  std::vector<DeviceHandle> devices;
  std::vector<std::string> input_columns;
  std::vector<std::string> output_columns;
  std::vector<u8> args;
  i32 node_id = 0;
};

class BaseKernel {
 public:
  explicit BaseKernel(const KernelConfig&) {}
  virtual ~BaseKernel() {}
  virtual void validate(Result* result) { result->set_success(true); }
  virtual void reset() {}
};

class BatchedKernel : public BaseKernel {
 public:
  explicit BatchedKernel(const KernelConfig& c) : BaseKernel(c) {}
  virtual void execute(const BatchedElements& input_columns, BatchedElements& output_columns) = 0;
};

class StenciledKernel : public BaseKernel {
 public:
  explicit StenciledKernel(const KernelConfig& c) : BaseKernel(c) {}
  virtual void execute(const StenciledElements& input_columns, Elements& output_columns) = 0;
};

class StenciledBatchedKernel : public BaseKernel {
 public:
  explicit StenciledBatchedKernel(const KernelConfig& c) : BaseKernel(c) {}
  virtual void execute(const StenciledBatchedElements& input_columns, BatchedElements& output_columns) = 0;
};

// optical_flow_kernel_cpu.cpp:19-25,30: check_frame() refreshes frame_info_ and calls
// new_frame_info() when the incoming FrameInfo changes.
class VideoKernel {
 public:
  virtual ~VideoKernel() {}

 protected:
  void check_frame(const DeviceHandle& /*device*/, const Element& element) {
    const FrameInfo info = element.as_const_frame()->as_frame_info();
    if (!have_info_ || info != frame_info_) {
      frame_info_ = info;
      have_info_ = true;
      new_frame_info();
    }
  }
  virtual void new_frame_info() {}
  FrameInfo frame_info_;

 private:
  bool have_info_ = false;
};

}  // namespace scanner
