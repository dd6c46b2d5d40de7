// multi_gpu_utils.cpp -- run_multi_gpu_jit_host over wdb_multi_project_filter_host.
// Reference: src/multi_gpu_utils.cpp:5-63 (sequential per-device upload/compile/launch/download).
#include "multi_gpu_utils.hpp"

#include <stdexcept>

#include "warpcore.h"

namespace {
const void *host_ptr(const HostColumn &c) {
  switch (c.type) {
  case DataType::Int32: return std::get<std::vector<int32_t>>(c.data).data();
  case DataType::Int64: return std::get<std::vector<int64_t>>(c.data).data();
  case DataType::Float32: return std::get<std::vector<float>>(c.data).data();
  case DataType::Float64: return std::get<std::vector<double>>(c.data).data();
  case DataType::String: return nullptr;
  }
  return nullptr;
}
}  // namespace

std::vector<float> run_multi_gpu_jit_host(const HostTable &host, const std::string &expr_cuda, const std::string &cond_cuda) {
  const int64_t n = host.num_rows();
  std::vector<wdb_col_t> cols;
  for (const auto &c : host.columns) cols.push_back(wdb_col_t{c.name.c_str(), static_cast<int>(c.type), host_ptr(c), n});
  std::vector<float> result(static_cast<size_t>(n));
  int64_t count = 0;
  // ndev = 0: every visible device, like cudaGetDeviceCount in the reference (:8-9).  Rows failing
  // the condition come back as 0.0f (the reference leaves them uninitialised: SURVEY F4).
  if (wdb_multi_project_filter_host(0, nullptr, cols.data(), static_cast<int>(cols.size()), expr_cuda.c_str(), cond_cuda.c_str(),
                                    result.data(), n, WDB_DENSE_ZERO, &count))
    throw std::runtime_error(wdb_last_error());
  return result;
}
