// Scanner-API look-alike (compat shim) -- ONLY used when the real Scanner headers are absent.
// It declares exactly the members the hot-path kernels of scannertools use (SURVEY.md §8b,
// collected from the reference's call sites; the real headers live in the Scanner engine,
// which is not part of /root/reference).  When building against real Scanner, put its include
// directory first and this directory is never consulted.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace scanner {

using u8 = uint8_t;
using i32 = int32_t;
using i64 = int64_t;
using u64 = uint64_t;
using f32 = float;

enum class DeviceType { CPU = 0, GPU = 1 };

struct DeviceHandle {
  DeviceType type;
  i32 id;
  bool operator==(const DeviceHandle& o) const { return type == o.type && id == o.id; }
  bool operator!=(const DeviceHandle& o) const { return !(*this == o); }
};

static const DeviceHandle CPU_DEVICE = {DeviceType::CPU, 0};

// reference usage: RESULT_ERROR(&valid_, "...") + validate(Result*) (blur_kernel_cpu.cpp:29-33,44)
class Result {
 public:
  bool success() const { return success_; }
  const std::string& msg() const { return msg_; }
  void set_success(bool s) { success_ = s; }
  void set_msg(const std::string& m) { msg_ = m; }
  void CopyFrom(const Result& o) { *this = o; }

 private:
  bool success_ = true;
  std::string msg_;
};

#define RESULT_ERROR(result__, ...)                           \
  do {                                                        \
    char scanner_buf__[512];                                  \
    snprintf(scanner_buf__, sizeof(scanner_buf__), __VA_ARGS__); \
    (result__)->set_success(false);                           \
    (result__)->set_msg(scanner_buf__);                       \
  } while (0)

}  // namespace scanner
