// multi_gpu_utils.cpp -- run_multi_gpu_jit_host over wdb_multi_project_filter_host.
// Reference: src/multi_gpu_utils.cpp:5-63 (sequential per-device upload/compile/launch/download).
#include "multi_gpu_utils.hpp"

#include <algorithm>
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
std::vector<wdb_col_t> host_cols(const HostTable &host) {
  std::vector<wdb_col_t> cols;
  for (const auto &c : host.columns) cols.push_back(wdb_col_t{c.name.c_str(), static_cast<int>(c.type), host_ptr(c), host.num_rows()});
  return cols;
}
}  // namespace

MultiGpuGroups run_multi_gpu_group_host(const HostTable &host, const std::string &val_cuda, const std::string &key_cuda,
                                        const std::string &cond_cuda, AggregationType agg, bool descending) {
  const int64_t n = host.num_rows();
  const std::vector<wdb_col_t> cols = host_cols(host);
  MultiGpuGroups out;
  int64_t cap = std::min<int64_t>(std::max<int64_t>(n, 1), 1 << 20), groups = 0;
  for (;;) {   // the number of groups is not known in advance: grow the output once if it did not fit
    out.keys.resize(static_cast<size_t>(cap));
    out.vals.resize(static_cast<size_t>(cap));
    if (!wdb_multi_group_agg_host(0, nullptr, cols.data(), static_cast<int>(cols.size()), val_cuda.c_str(), key_cuda.c_str(), cond_cuda.c_str(),
                                  static_cast<int>(agg), descending ? WDB_ORDER_KEY_DESC : WDB_ORDER_KEY_ASC, n, 0, out.keys.data(), out.vals.data(),
                                  cap, &groups))
      break;
    const std::string msg = wdb_last_error();
    if (msg.find("exceed the output capacity") == std::string::npos || cap >= n) throw std::runtime_error(msg);
    cap = n;
  }
  out.keys.resize(static_cast<size_t>(groups));
  out.vals.resize(static_cast<size_t>(groups));
  return out;
}

std::vector<float> run_multi_gpu_topk_host(const HostTable &host, const std::string &key_cuda, const std::string &val_cuda,
                                           const std::string &cond_cuda, bool descending, int limit, int offset) {
  const std::vector<wdb_col_t> cols = host_cols(host);
  std::vector<float> out(static_cast<size_t>(std::max(limit, 1)));
  int64_t m = 0;
  if (wdb_multi_topk_host(0, nullptr, cols.data(), static_cast<int>(cols.size()), key_cuda.c_str(), val_cuda.c_str(), cond_cuda.c_str(), descending ? 1 : 0,
                          limit, offset, host.num_rows(), out.data(), &m))
    throw std::runtime_error(wdb_last_error());
  out.resize(static_cast<size_t>(m));
  return out;
}

std::vector<float> run_multi_gpu_compact_host(const HostTable &host, const std::string &expr_cuda, const std::string &cond_cuda) {
  const int64_t n = host.num_rows();
  const std::vector<wdb_col_t> cols = host_cols(host);
  std::vector<float> out(static_cast<size_t>(n));
  int64_t count = 0;
  if (wdb_multi_project_filter_host(0, nullptr, cols.data(), static_cast<int>(cols.size()), expr_cuda.c_str(), cond_cuda.c_str(), out.data(), n,
                                    WDB_COMPACT, &count))
    throw std::runtime_error(wdb_last_error());
  out.resize(static_cast<size_t>(count));
  return out;
}

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
