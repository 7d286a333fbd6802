// Scanner-API look-alike: static-initialiser op / kernel registration (see util/common.h).
// REGISTER_OP(Name).frame_input(..).output(..).stencil({..}) and
// REGISTER_KERNEL(Name, Class).device(..).batch().num_devices(1) build entries in a process-wide
// registry, as Scanner's macros do (histogram_kernel_cpu.cpp:52-57, optical_flow_kernel_cpu.cpp:51-58).
#pragma once
#include <functional>
#include <map>
#include <memory>

#include "scanner/api/kernel.h"

namespace scanner {

enum class ColumnType { Other = 0, Video = 1, Bytes = 2 };

struct OpInfo {
  std::string name;
  std::vector<std::pair<std::string, ColumnType>> inputs, outputs;
  std::vector<std::string> output_type_names;
  std::vector<i32> stencil;
};

struct KernelInfo {
  std::string op_name;
  DeviceType device = DeviceType::CPU;
  bool batched = false;
  i32 num_devices = 1;
  std::function<BaseKernel*(const KernelConfig&)> factory;
};

class Registry {
 public:
  static Registry& get() { static Registry r; return r; }
  std::map<std::string, OpInfo> ops;
  std::vector<KernelInfo> kernels;
  const KernelInfo* find_kernel(const std::string& op, DeviceType dev) const {
    for (const auto& k : kernels) if (k.op_name == op && k.device == dev) return &k;
    return nullptr;
  }
};

class OpBuilder {
 public:
  explicit OpBuilder(const std::string& name) { info_.name = name; }
  OpBuilder& frame_input(const std::string& n) { info_.inputs.push_back({n, ColumnType::Video}); return commit(); }
  OpBuilder& input(const std::string& n, ColumnType t = ColumnType::Bytes) { info_.inputs.push_back({n, t}); return commit(); }
  OpBuilder& frame_output(const std::string& n) { info_.outputs.push_back({n, ColumnType::Video}); info_.output_type_names.push_back(""); return commit(); }
  OpBuilder& output(const std::string& n, ColumnType t = ColumnType::Bytes, const std::string& type_name = "") {
    info_.outputs.push_back({n, t}); info_.output_type_names.push_back(type_name); return commit();
  }
  OpBuilder& stencil(const std::vector<i32>& s) { info_.stencil = s; return commit(); }

 private:
  OpBuilder& commit() { Registry::get().ops[info_.name] = info_; return *this; }
  OpInfo info_;
};

class KernelBuilder {
 public:
  KernelBuilder(const std::string& op, std::function<BaseKernel*(const KernelConfig&)> f) {
    idx_ = Registry::get().kernels.size();
    KernelInfo k; k.op_name = op; k.factory = std::move(f);
    Registry::get().kernels.push_back(k);
  }
  KernelBuilder& device(DeviceType d) { me().device = d; return *this; }
  KernelBuilder& batch(i32 = 1) { me().batched = true; return *this; }
  KernelBuilder& num_devices(i32 n) { me().num_devices = n; return *this; }

 private:
  KernelInfo& me() { return Registry::get().kernels[idx_]; }
  size_t idx_;
};

#define SCANNER_CAT_(a, b) a##b
#define SCANNER_CAT(a, b) SCANNER_CAT_(a, b)
#define REGISTER_OP(name__) static ::scanner::OpBuilder SCANNER_CAT(scanner_op_reg_, __COUNTER__) = ::scanner::OpBuilder(#name__)
#define REGISTER_KERNEL(name__, kernel__)                                                         \
  static ::scanner::KernelBuilder SCANNER_CAT(scanner_kernel_reg_, __COUNTER__) = ::scanner::KernelBuilder( \
      #name__, [](const ::scanner::KernelConfig& config) -> ::scanner::BaseKernel* { return new kernel__(config); })

}  // namespace scanner
