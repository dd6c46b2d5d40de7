// csv_loader.cpp -- see csv_loader.hpp.  Errors follow the reference's messages where tests or
// callers can observe them ("Unable to open file", "Empty CSV file", "Schema size does not match
// column count": src/csv_loader.cpp:49-68).
#include "csv_loader.hpp"

#include "warpcore.h"

#include <cuda_runtime.h>

#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

int HostTable::num_rows() const {
  if (columns.empty()) return 0;
  return std::visit([](const auto &v) { return static_cast<int>(v.size()); }, columns.front().data);
}
const HostColumn *HostTable::get_column(const std::string &name) const {
  for (const auto &c : columns)
    if (c.name == name) return &c;
  return nullptr;
}

std::vector<std::string> split_csv_header(const std::string &header_line) {
  std::vector<std::string> names;
  std::stringstream ss(header_line);
  std::string field;
  while (std::getline(ss, field, ',')) {
    while (!field.empty() && (field.back() == '\r' || field.back() == '\n')) field.pop_back();
    names.push_back(field);
  }
  return names;
}

namespace {
ColumnData empty_column(DataType t) {
  switch (t) {
  case DataType::Int32: return std::vector<int32_t>();
  case DataType::Int64: return std::vector<int64_t>();
  case DataType::Float32: return std::vector<float>();
  case DataType::Float64: return std::vector<double>();
  case DataType::String: return std::vector<std::string>();
  }
  return std::vector<float>();
}
void append_field(HostColumn &col, const std::string &text) {
  switch (col.type) {
  case DataType::Int32: std::get<std::vector<int32_t>>(col.data).push_back(std::stoi(text)); break;
  case DataType::Int64: std::get<std::vector<int64_t>>(col.data).push_back(std::stoll(text)); break;
  case DataType::Float32: std::get<std::vector<float>>(col.data).push_back(std::stof(text)); break;
  case DataType::Float64: std::get<std::vector<double>>(col.data).push_back(std::stod(text)); break;
  case DataType::String: std::get<std::vector<std::string>>(col.data).push_back(text); break;
  }
}
void append_row(HostTable &t, const std::string &line) {
  std::stringstream ss(line);
  std::string field;
  for (auto &col : t.columns) {
    if (!std::getline(ss, field, ',')) field.clear();
    append_field(col, field);
  }
}
HostTable make_host_table(const std::vector<std::string> &names, const std::vector<DataType> &types) {
  HostTable t;
  for (size_t i = 0; i < names.size(); ++i) t.columns.push_back(HostColumn{names[i], types[i], empty_column(types[i])});
  return t;
}
size_t element_size(DataType t) {
  switch (t) {
  case DataType::Int32: case DataType::Float32: return 4;
  case DataType::Int64: case DataType::Float64: return 8;
  case DataType::String: return 0;
  }
  return 0;
}
const void *host_data(const HostColumn &c) {
  switch (c.type) {
  case DataType::Int32: return std::get<std::vector<int32_t>>(c.data).data();
  case DataType::Int64: return std::get<std::vector<int64_t>>(c.data).data();
  case DataType::Float32: return std::get<std::vector<float>>(c.data).data();
  case DataType::Float64: return std::get<std::vector<double>>(c.data).data();
  case DataType::String: return nullptr;
  }
  return nullptr;
}
void cuda_or_throw(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " (" + what + ")");
}
}  // namespace

// HostTable::price / ::quantity (legacy view, see csv_loader.hpp)
static void fill_legacy_view(HostTable &t) {
  auto to_vec = [](const HostColumn *c, auto &out) {
    if (!c) return;
    std::visit([&](const auto &v) {
      using E = typename std::decay_t<decltype(v)>::value_type;
      if constexpr (!std::is_same_v<E, std::string>) {
        out.clear();
        for (const auto &x : v) out.push_back(static_cast<typename std::decay_t<decltype(out)>::value_type>(x));
      }
    }, c->data);
  };
  to_vec(t.get_column("price"), t.price);
  to_vec(t.get_column("quantity"), t.quantity);
}

HostTable load_csv_to_host(const std::string &filepath, const std::vector<DataType> &schema) {
  std::ifstream file(filepath);
  if (!file.is_open()) {
    std::cerr << "Failed to open file: " << filepath << std::endl;
    throw std::runtime_error("Unable to open file");
  }
  std::string header;
  if (!std::getline(file, header)) throw std::runtime_error("Empty CSV file");
  const std::vector<std::string> names = split_csv_header(header);
  std::vector<DataType> types = schema;
  if (!types.empty() && types.size() != names.size()) throw std::runtime_error("Schema size does not match column count");
  if (types.empty()) types.assign(names.size(), DataType::Float32);   // reference default: src/csv_loader.cpp:68
  HostTable t = make_host_table(names, types);
  std::string line;
  while (std::getline(file, line)) {
    if (line.empty()) continue;
    append_row(t, line);
  }
  fill_legacy_view(t);
  return t;
}

HostTable load_csv_chunk(std::istream &stream, int max_rows, bool &finished, const std::vector<std::string> &names) {
  HostTable t = make_host_table(names, std::vector<DataType>(names.size(), DataType::Float32));
  int count = 0;
  std::string line;
  while (count < max_rows && std::getline(stream, line)) {
    if (line.empty()) continue;
    append_row(t, line);
    ++count;
  }
  finished = !stream.good();
  return t;
}

Table upload_to_gpu(const HostTable &host) { return upload_to_gpu(host, nullptr); }

Table upload_to_gpu(const HostTable &host, std::vector<ColumnIngestStats> *stats) {
  Table table;
  table.num_rows = host.num_rows();
  for (const auto &hc : host.columns) {
    void *d = nullptr;
    const size_t bytes = element_size(hc.type) * static_cast<size_t>(table.num_rows);
    ColumnIngestStats st;
    st.name = hc.name;
    if (hc.type != DataType::String) {   // string columns stay on the host (src/csv_loader.cpp:151-155)
      cuda_or_throw(cudaMalloc(&d, bytes ? bytes : 4), "cudaMalloc column");
      const wdb_col_t col{hc.name.c_str(), static_cast<int>(hc.type), host_data(hc), table.num_rows};
      wdb_zonemap_t *zm = nullptr;
      const bool want = stats != nullptr && table.num_rows > 0;
      if (wdb_upload_column(0, nullptr, &col, d, want ? 0 : -1, want ? &zm : nullptr, want ? &st.min : nullptr, want ? &st.max : nullptr)) {
        cudaFree(d);
        throw std::runtime_error(wdb_last_error());
      }
      st.numeric = want;
      st.zonemap = zm;
    }
    if (stats) stats->push_back(st);
    table.columns.push_back(ColumnDesc{hc.name, hc.type, d, table.num_rows});
  }
  cuda_or_throw(cudaStreamSynchronize(nullptr), "upload");
  return table;
}

Table load_csv_to_gpu(const std::string &filepath, const std::vector<DataType> &schema) {
  return upload_to_gpu(load_csv_to_host(filepath, schema));
}

void free_table(Table &table) {
  for (auto &c : table.columns)
    if (c.device_ptr) { cudaFree(c.device_ptr); c.device_ptr = nullptr; }
}

// NDJSON: one object per line with "price" (number) and "quantity" (integer) members
// (reference: src/json_loader.cpp:16-53 searches for the same two member names)
namespace {
bool member_number(const std::string &line, const char *name, double *out) {
  const std::string key = std::string("\"") + name + "\"";
  size_t p = line.find(key);
  if (p == std::string::npos) return false;
  p = line.find(':', p + key.size());
  if (p == std::string::npos) return false;
  try { *out = std::stod(line.substr(p + 1)); } catch (const std::exception &) { return false; }
  return true;
}
}  // namespace

HostTable load_json_to_host(const std::string &filepath) {
  std::ifstream file(filepath);
  if (!file.is_open()) throw std::runtime_error("Unable to open file");
  HostTable t = make_host_table({"price", "quantity"}, {DataType::Float32, DataType::Int32});
  std::string line;
  while (std::getline(file, line)) {
    double p = 0, q = 0;
    if (!member_number(line, "price", &p) || !member_number(line, "quantity", &q)) continue;
    std::get<std::vector<float>>(t.columns[0].data).push_back(static_cast<float>(p));
    std::get<std::vector<int32_t>>(t.columns[1].data).push_back(static_cast<int32_t>(q));
  }
  return t;
}
Table load_json_to_gpu(const std::string &filepath) { return upload_to_gpu(load_json_to_host(filepath)); }
