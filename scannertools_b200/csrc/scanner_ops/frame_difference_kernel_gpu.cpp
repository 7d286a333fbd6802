// FrameDifference op, GPU kernel.  The reference file
// (scannertools_cpp/imgproc/frame_difference_kernel_cpu.cpp:25-81) is dead code that does not
// compile (constructor named BlurKernel :27, missing ';' :61, stray ';' in the registration
// :78-79) and indexes without x (:59); this implements its evident intent -- stencil {-1, 0},
// out = frame[t] - frame[t-1] per byte with u8 wrap-around -- behind the same op name.
#include "scanner/api/kernel.h"
#include "scanner/api/op.h"
#include "scanner/util/cuda.h"
#include "scanner/util/memory.h"
#include "stb_check.h"

namespace scanner {

class FrameDifferenceKernelGPU : public StenciledKernel, public VideoKernel {
 public:
  FrameDifferenceKernelGPU(const KernelConfig& config) : StenciledKernel(config), device_(config.devices[0]) {
    valid_.set_success(true);
    CU_CHECK(cudaSetDevice(device_.id));
    CU_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  }

  ~FrameDifferenceKernelGPU() {
    cudaSetDevice(device_.id);
    cudaStreamDestroy(stream_);
  }

  void validate(Result* result) override { result->CopyFrom(valid_); }

  void execute(const StenciledElements& input_columns, Elements& output_columns) override {
    auto& frame_col = input_columns[0];
    CU_CHECK(cudaSetDevice(device_.id));
    check_frame(device_, frame_col[0]);

    const Frame* secondary = frame_col[0].as_const_frame();   // t-1
    const Frame* primary = frame_col[1].as_const_frame();     // t
    FrameInfo info = primary->as_frame_info();
    if (secondary->as_frame_info() != info || primary->type != FrameType::U8) {
      // the subtraction is defined on bytes (frame_difference_kernel_cpu.cpp:51-61 reads u8)
      STB_FATAL("FrameDifference (B200): both frames must be U8 with one FrameInfo");
    }
    Frame* output_frame = new_frame(device_, info);
    STB_CHECK(stb_frame_diff(secondary->data, primary->data, output_frame->data, info.size(), stream_));
    insert_frame(output_columns[0], output_frame);   // one pre-sized slot per output column (frame_difference_kernel_cpu.cpp:64)
    CU_CHECK(cudaStreamSynchronize(stream_));
  }

 private:
  DeviceHandle device_;
  cudaStream_t stream_;
  Result valid_;
};

REGISTER_OP(FrameDifference).frame_input("frame").frame_output("frame").stencil({-1, 0});

REGISTER_KERNEL(FrameDifference, FrameDifferenceKernelGPU).device(DeviceType::GPU).num_devices(1);
}  // namespace scanner
