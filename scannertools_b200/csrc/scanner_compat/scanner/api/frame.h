// Scanner-API look-alike: Frame / FrameInfo / Element containers (see util/common.h).
// Members as used at optical_flow_kernel_cpu.cpp:32-34, histogram_kernel_cpu.cpp:16-18,
// frame_difference_kernel_cpu.cpp:45-54.
#pragma once
#include "scanner/util/common.h"

namespace scanner {

enum class FrameType { U8 = 0, F32 = 1, F64 = 2 };

inline size_t size_of_frame_type(FrameType t) { return t == FrameType::U8 ? 1 : (t == FrameType::F32 ? 4 : 8); }

class FrameInfo {
 public:
  FrameInfo() = default;
  FrameInfo(int height, int width, int channels, FrameType type)
    : shape{height, width, channels}, type(type) {}
  int height() const { return shape[0]; }
  int width() const { return shape[1]; }
  int channels() const { return shape[2]; }
  size_t size() const { return (size_t)shape[0] * shape[1] * shape[2] * size_of_frame_type(type); }
  bool operator==(const FrameInfo& o) const {
    return shape[0] == o.shape[0] && shape[1] == o.shape[1] && shape[2] == o.shape[2] && type == o.type;
  }
  bool operator!=(const FrameInfo& o) const { return !(*this == o); }
  int shape[3] = {0, 0, 0};
  FrameType type = FrameType::U8;
};

// frames are packed row-major HWC with no row pitch (blur_kernel_cpu.cpp:70)
class Frame {
 public:
  Frame(FrameInfo info, u8* buffer) : data(buffer), type(info.type), info_(info) {}
  FrameInfo as_frame_info() const { return info_; }
  int height() const { return info_.height(); }
  int width() const { return info_.width(); }
  int channels() const { return info_.channels(); }
  size_t size() const { return info_.size(); }
  u8* data;
  FrameType type;   // resize_kernel.cpp:64 reads frame->type

 private:
  FrameInfo info_;
};

struct Element {
  Element() = default;
  Element(u8* buf, size_t sz) : buffer(buf), size(sz), is_frame(false) {}
  explicit Element(Frame* frame) : buffer(reinterpret_cast<u8*>(frame)), size(sizeof(Frame)), is_frame(true) {}
  const Frame* as_const_frame() const { return reinterpret_cast<const Frame*>(buffer); }
  Frame* as_frame() const { return reinterpret_cast<Frame*>(buffer); }
  bool is_null() const { return buffer == nullptr; }
  u8* buffer = nullptr;
  size_t size = 0;
  bool is_frame = false;
};

using Elements = std::vector<Element>;
using BatchedElements = std::vector<Elements>;                       // [column][row]
using StenciledElements = std::vector<Elements>;                     // [column][stencil position]
using StenciledBatchedElements = std::vector<std::vector<Elements>>; // [column][row][stencil position]

inline size_t num_rows(const Elements& column) { return column.size(); }

}  // namespace scanner
