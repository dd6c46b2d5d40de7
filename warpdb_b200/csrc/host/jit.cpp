// jit.cpp -- reference-signature shims over the C ABI (see jit.hpp).
#include "jit.hpp"

#include <cuda_runtime.h>

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "warpcore.h"

namespace {
[[noreturn]] void raise_last() { throw std::runtime_error(wdb_last_error()); }

// The reference re-reads ./custom.cu from the working directory on every call and prepends it to
// the kernel (src/jit.cpp:65-73); a missing file means "no UDFs".
void refresh_udf_source() {
  std::ifstream in("custom.cu");
  std::string src;
  if (in) {
    std::stringstream ss;
    ss << in.rdbuf();
    src = ss.str();
  }
  wdb_set_udf_source(src.c_str());
}

std::vector<wdb_col_t> describe(const Table &table) {
  std::vector<wdb_col_t> cols;
  cols.reserve(table.columns.size());
  for (const auto &c : table.columns) cols.push_back(wdb_col_t{c.name.c_str(), static_cast<int>(c.type), c.device_ptr, table.num_rows});
  return cols;
}
}  // namespace

void jit_compile_and_launch(const std::string &expr_code, const std::string &condition_code, const Table &table,
                            float *d_output, int device_id) {
  refresh_udf_source();
  const std::vector<wdb_col_t> cols = describe(table);
  int64_t count = 0;  // passing a host count makes the call synchronous like cuCtxSynchronize (src/jit.cpp:171)
  if (wdb_project_filter(device_id, nullptr, cols.data(), static_cast<int>(cols.size()), expr_code.c_str(), condition_code.c_str(),
                         d_output, table.num_rows, WDB_DENSE, nullptr, &count))
    raise_last();
}

void jit_group_sum(const std::string &val_expr_code, const std::string &key_expr_code, float *d_price, int *d_quantity,
                   float *d_out_vals, int *d_out_keys, int *d_count, int N, int device_id) {
  refresh_udf_source();
  // the reference kernel hard-wires the two parameters (float* price, int* quantity): src/jit.cpp:194
  const wdb_col_t cols[2] = {{"price", WDB_FLOAT32, d_price, N}, {"quantity", WDB_INT32, d_quantity, N}};
  int64_t groups = 0;
  if (wdb_group_agg(device_id, nullptr, cols, 2, val_expr_code.c_str(), key_expr_code.c_str(), "", WDB_SUM, WDB_ORDER_FIRST, N, 0,
                    d_out_keys, d_out_vals, N, &groups))
    raise_last();
  const int g = static_cast<int>(groups);
  if (cudaMemcpy(d_count, &g, sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) throw std::runtime_error("CUDA error: count copy failed");
}

void jit_sort_pairs(int *d_keys, float *d_vals, int count, bool ascending, int device_id) {
  if (wdb_sort_pairs(device_id, nullptr, d_keys, d_vals, count, ascending ? 1 : 0)) raise_last();
  if (cudaDeviceSynchronize() != cudaSuccess) throw std::runtime_error("CUDA error: sort failed");
}

void jit_sort_float(float *d_vals, int count, bool ascending, int device_id) {
  if (wdb_sort_float(device_id, nullptr, d_vals, count, ascending ? 1 : 0)) raise_last();
  if (cudaDeviceSynchronize() != cudaSuccess) throw std::runtime_error("CUDA error: sort failed");
}
