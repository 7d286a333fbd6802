// FlowHistogram op, GPU kernel.  The reference only ships a CPU kernel
// (scannertools/old/cpp_ops/flow_histogram_kernel_cpu.cpp:12-67); this is its device-side
// counterpart with the same batched interface and the same 512-byte output element
// (int32[2][64]: magnitude bins then angle bins), so `FlowHistogram(flow=..., device=GPU)` can
// consume OpticalFlow's device frames without a round trip through the host.
#include <vector>

#include "scanner/api/kernel.h"
#include "scanner/api/op.h"
#include "scanner/util/cuda.h"
#include "scanner/util/memory.h"
#include "stb_check.h"

namespace scanner {
namespace {
const i32 BINS = 64;   // flow_histogram_kernel_cpu.cpp:9
}

class FlowHistogramKernelGPU : public BatchedKernel, public VideoKernel {
 public:
  FlowHistogramKernelGPU(const KernelConfig& config) : BatchedKernel(config), device_(config.devices[0]) {
    CU_CHECK(cudaSetDevice(device_.id));
    CU_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  }

  ~FlowHistogramKernelGPU() {
    cudaSetDevice(device_.id);
    cudaStreamDestroy(stream_);
  }

  void execute(const BatchedElements& input_columns, BatchedElements& output_columns) override {
    auto& frame_col = input_columns[0];
    CU_CHECK(cudaSetDevice(device_.id));
    check_frame(device_, frame_col[0]);

    const size_t hist_size = BINS * 2 * sizeof(int);
    const i32 input_count = (i32)num_rows(frame_col);
    u8* output_block = new_block_buffer_size(device_, hist_size, input_count);

    flows_.resize(input_count);
    for (i32 i = 0; i < input_count; ++i) flows_[i] = reinterpret_cast<const float*>(frame_col[i].as_const_frame()->data);
    STB_CHECK(stb_flow_hist(flows_.data(), input_count, frame_info_.width(), frame_info_.height(),
                            reinterpret_cast<int32_t*>(output_block), stream_));
    for (i32 i = 0; i < input_count; ++i) insert_element(output_columns[0], output_block + i * hist_size, hist_size);
    CU_CHECK(cudaStreamSynchronize(stream_));
  }

 private:
  DeviceHandle device_;
  cudaStream_t stream_;
  std::vector<const float*> flows_;
};

#ifndef STB_SKIP_OP_DECLARATIONS
REGISTER_OP(FlowHistogram).frame_input("flow").output("histogram");
#endif

REGISTER_KERNEL(FlowHistogram, FlowHistogramKernelGPU).device(DeviceType::GPU).batch().num_devices(1);
}  // namespace scanner
