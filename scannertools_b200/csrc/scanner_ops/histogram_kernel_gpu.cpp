// Histogram op, GPU kernel -- drop-in for the reference's HistogramKernelGPU
// (scannertools_cpp/imgproc/histogram_kernel_gpu.cpp:12-82): same class shape, same
// REGISTER_KERNEL line, same output (one 192-byte element per frame carved out of one device
// block buffer).  Where the reference calls cvc::split + 3 x cvc::histEven per frame on the
// default stream (:49-57), this issues ONE stb_hist_rgb16 launch for the whole batch.
// The op declaration (REGISTER_OP(Histogram)...) stays in histogram_kernel_cpu.cpp:52 of the
// reference build; it is repeated here only so this library is self-contained without it.
#include <vector>

#include "scanner/api/kernel.h"
#include "scanner/api/op.h"
#include "scanner/util/cuda.h"
#include "scanner/util/memory.h"
#include "stb_check.h"

namespace scanner {
namespace {
const i32 BINS = 16;   // histogram_kernel_cpu.cpp:8
}

class HistogramKernelGPU : public BatchedKernel, public VideoKernel {
 public:
  HistogramKernelGPU(const KernelConfig& config) : BatchedKernel(config), device_(config.devices[0]) {
    set_device();
    CU_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  }

  ~HistogramKernelGPU() {
    set_device();
    cudaStreamDestroy(stream_);
  }

  void new_frame_info() override { set_device(); }

  void execute(const BatchedElements& input_columns, BatchedElements& output_columns) override {
    auto& frame_col = input_columns[0];
    set_device();
    check_frame(device_, frame_col[0]);

    const size_t hist_size = BINS * 3 * sizeof(int);
    const i32 input_count = (i32)num_rows(frame_col);
    u8* output_block = new_block_buffer(device_, hist_size * input_count, input_count);

    frames_.resize(input_count);
    for (i32 i = 0; i < input_count; ++i) frames_[i] = frame_col[i].as_const_frame()->data;
    STB_CHECK(stb_hist_rgb16(frames_.data(), input_count, frame_info_.width(), frame_info_.height(),
                             reinterpret_cast<int32_t*>(output_block), stream_));
    for (i32 i = 0; i < input_count; ++i) insert_element(output_columns[0], output_block + i * hist_size, hist_size);
    // the engine consumes the column after execute() returns (the reference also syncs, :62-64)
    CU_CHECK(cudaStreamSynchronize(stream_));
  }

 private:
  void set_device() { CUDA_PROTECT({ CU_CHECK(cudaSetDevice(device_.id)); }); }

  DeviceHandle device_;
  cudaStream_t stream_;
  std::vector<const uint8_t*> frames_;
};

#ifndef STB_SKIP_OP_DECLARATIONS
REGISTER_OP(Histogram).frame_input("frame").output("histogram", ColumnType::Bytes, "Histogram");
#endif

REGISTER_KERNEL(Histogram, HistogramKernelGPU).device(DeviceType::GPU).batch().num_devices(1);
}  // namespace scanner
