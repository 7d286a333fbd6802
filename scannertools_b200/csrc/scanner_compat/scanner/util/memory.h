// Scanner-API look-alike: output allocation and hand-over (see util/common.h).
// new_block_buffer / new_block_buffer_size / new_frame / new_frames / insert_frame /
// insert_element as called at histogram_kernel_gpu.cpp:40-41,59, histogram_kernel_cpu.cpp:23,44,
// optical_flow_kernel_cpu.cpp:34,42, optical_flow_kernel_gpu.cpp:63-64,88.
// In real Scanner these are ref-counted pool allocations owned by the engine; here they are
// plain cudaMalloc / malloc blocks released by the test harness (delete_element).
#pragma once
#include "scanner/api/frame.h"

namespace scanner {

u8* new_buffer(DeviceHandle device, size_t size);
void delete_buffer(DeviceHandle device, u8* buffer);
// one allocation shared by `refs` elements (ref counting is the engine's business)
u8* new_block_buffer(DeviceHandle device, size_t size, i32 refs);
inline u8* new_block_buffer_size(DeviceHandle device, size_t element_size, i32 refs) {
  return new_block_buffer(device, element_size * (size_t)refs, refs);
}
Frame* new_frame(DeviceHandle device, FrameInfo info);
std::vector<Frame*> new_frames(DeviceHandle device, FrameInfo info, i32 num);

// batched kernels append to a column (histogram_kernel_cpu.cpp:44, resize_kernel.cpp:83) ...
inline void insert_frame(Elements& column, Frame* frame) { column.push_back(Element(frame)); }
inline void insert_element(Elements& column, u8* buffer, size_t size) { column.push_back(Element(buffer, size)); }
// ... non-batched kernels fill the ONE pre-sized slot of each output column: `output_columns` is an
// Elements with one Element per column (blur_kernel_cpu.cpp:79, optical_flow_kernel_cpu.cpp:42,
// frame_difference_kernel_cpu.cpp:64)
inline void insert_frame(Element& element, Frame* frame) { element = Element(frame); }
inline void insert_element(Element& element, u8* buffer, size_t size) { element = Element(buffer, size); }

void memcpy_buffer(u8* dst, DeviceHandle dst_device, const u8* src, DeviceHandle src_device, size_t size);

}  // namespace scanner
