// OpticalFlow op, GPU kernel -- drop-in for the reference's OpticalFlowKernelGPU
// (scannertools_cpp/imgproc/optical_flow_kernel_gpu.cpp:12-112): StenciledBatchedKernel +
// VideoKernel, batch of B stencil pairs -> B device flow frames (H x W x 2 F32) allocated with
// new_frames and handed over with insert_frame.  cv::cuda::cvtColor + FarnebackOpticalFlow::calc
// (:66-89) are replaced by one stb_farneback_run over the B+1 unique frames.
// Direction follows the CPU kernel (optical_flow_kernel_cpu.cpp:41, flow from stencil[0] to
// stencil[1]); the reference GPU kernel's reversed argument order (:82-87) is a defect
//  and is not reproduced.  No process-global state is touched (the
// reference toggles the global cv::cuda buffer pool, :19-20,32-33).
#include <vector>

#include "scanner/api/kernel.h"
#include "scanner/api/op.h"
#include "scanner/util/cuda.h"
#include "scanner/util/memory.h"
#include "stb_check.h"

namespace scanner {

class OpticalFlowKernelGPU : public StenciledBatchedKernel, public VideoKernel {
 public:
  OpticalFlowKernelGPU(const KernelConfig& config)
    : StenciledBatchedKernel(config), device_(config.devices[0]), handle_(nullptr), capacity_(0) {
    set_device();
    CU_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  }

  ~OpticalFlowKernelGPU() {
    set_device();
    stb_farneback_destroy(handle_);
    cudaStreamDestroy(stream_);
  }

  void new_frame_info() override {
    set_device();
    stb_farneback_destroy(handle_);   // scratch is sized per FrameInfo
    handle_ = nullptr;
    capacity_ = 0;
  }

  void reset() override { set_device(); }   // no cross-batch state (the reference clears initial_frame_, :40-43)

  void execute(const StenciledBatchedElements& input_columns, BatchedElements& output_columns) override {
    set_device();
    auto& frame_col = input_columns[0];
    check_frame(device_, frame_col[0][0]);

    const i32 input_count = (i32)frame_col.size();
    frames_.clear();
    for (i32 i = 0; i < input_count; ++i) frames_.push_back(frame_col[i][0].as_const_frame()->data);
    frames_.push_back(frame_col.back()[1].as_const_frame()->data);   // B+1 unique frames (:52-57)

    if (input_count > capacity_) {
      stb_farneback_destroy(handle_);
      handle_ = nullptr;
      capacity_ = input_count < 16 ? 16 : input_count;
      STB_CHECK(stb_farneback_create(frame_info_.width(), frame_info_.height(), capacity_, nullptr, &handle_));
    }

    FrameInfo out_frame_info(frame_info_.height(), frame_info_.width(), 2, FrameType::F32);
    std::vector<Frame*> output_frames = new_frames(device_, out_frame_info, input_count);
    flows_.resize(input_count);
    for (i32 i = 0; i < input_count; ++i) flows_[i] = reinterpret_cast<float*>(output_frames[i]->data);

    STB_CHECK(stb_farneback_run(handle_, frames_.data(), input_count, flows_.data(), stream_));
    for (i32 i = 0; i < input_count; ++i) insert_frame(output_columns[0], output_frames[i]);
    CU_CHECK(cudaStreamSynchronize(stream_));
  }

 private:
  void set_device() { CU_CHECK(cudaSetDevice(device_.id)); }

  DeviceHandle device_;
  cudaStream_t stream_;
  stb_farneback* handle_;
  i32 capacity_;
  std::vector<const uint8_t*> frames_;
  std::vector<float*> flows_;
};

#ifndef STB_SKIP_OP_DECLARATIONS
REGISTER_OP(OpticalFlow).frame_input("frame").frame_output("flow").stencil({0, 1});
#endif

REGISTER_KERNEL(OpticalFlow, OpticalFlowKernelGPU).device(DeviceType::GPU).batch().num_devices(1);
}  // namespace scanner
