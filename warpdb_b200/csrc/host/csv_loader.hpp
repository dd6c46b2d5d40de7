// csv_loader.hpp -- boundary data model (Table / ColumnDesc / HostTable) and the minimal CSV / NDJSON
// ingest needed to construct a WarpDB from the reference's fixtures.  Same names and fields as the
// reference's include/csv_loader.hpp:13-87 and include/json_loader.hpp; ingest itself is host-bound
// and outside the hot path.
#pragma once
#include <cstdint>
#include <iosfwd>
#include <string>
#include <variant>
#include <vector>

enum class DataType { Int32, Int64, Float32, Float64, String };

struct ColumnDesc {
  std::string name;
  DataType type;
  void *device_ptr;
  int length;
};

struct ColumnStatsFloat { float min = 0.0f; float max = 0.0f; int null_count = 0; };
struct ColumnStatsInt { int min = 0; int max = 0; int null_count = 0; };
struct TableStats {
  ColumnStatsFloat price;
  ColumnStatsInt quantity;
};

struct Table {
  std::vector<ColumnDesc> columns;
  int num_rows = 0;
  template <typename T> T *get_column_ptr(const std::string &name) const {
    for (const auto &c : columns)
      if (c.name == name) return static_cast<T *>(c.device_ptr);
    return nullptr;
  }
};

using ColumnData = std::variant<std::vector<int32_t>, std::vector<int64_t>, std::vector<float>, std::vector<double>,
                                std::vector<std::string>>;
struct HostColumn {
  std::string name;
  DataType type;
  ColumnData data;
};
struct HostTable {
  std::vector<HostColumn> columns;
  int num_rows() const;
  const HostColumn *get_column(const std::string &name) const;
  // Legacy two-column view that the reference's own tests/sql_features_test.cpp still uses
  // (`h.price[i]`, `h.quantity[i]`); filled by load_csv_to_host / load_json_to_host when columns of
  // these names exist.  Not read by anything in this library.
  std::vector<float> price;
  std::vector<int> quantity;
};

// What the ingest pass learns about a column while it uploads it (the statistics the reference's
// TableStats were meant to hold, per column instead of hard-wired to price / quantity): the exact
// min / max and the zone map (an opaque wdb_zonemap_t*, owned by whoever receives this struct).
struct ColumnIngestStats {
  std::string name;
  bool numeric = false;
  double min = 0.0, max = 0.0;
  void *zonemap = nullptr;
};

HostTable load_csv_to_host(const std::string &filepath, const std::vector<DataType> &schema = {});
// src/csv_loader.cpp:126-161.  Columns go up in pinned, asynchronous chunks; with `stats` the zone map
// and min / max of every numeric column are built on the device in the same pass (wdb_upload_column).
Table upload_to_gpu(const HostTable &table);
Table upload_to_gpu(const HostTable &table, std::vector<ColumnIngestStats> *stats);
Table load_csv_to_gpu(const std::string &filepath, const std::vector<DataType> &schema = {});
// At most max_rows rows from an open stream whose header line has already been consumed by the
// caller.  `names` are the column names of that header (the reference re-reads a "header" from
// every chunk and so loses the first row of each chunk: src/csv_loader.cpp:186-199).
HostTable load_csv_chunk(std::istream &stream, int max_rows, bool &finished, const std::vector<std::string> &names);
std::vector<std::string> split_csv_header(const std::string &header_line);

HostTable load_json_to_host(const std::string &filepath);   // NDJSON with "price" and "quantity"
Table load_json_to_gpu(const std::string &filepath);

void free_table(Table &table);
