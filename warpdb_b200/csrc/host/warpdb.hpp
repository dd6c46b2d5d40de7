// warpdb.hpp -- the WarpDB facade with the reference's public signatures (include/warpdb.hpp:11-48),
// running on the B200-native core (include/warpcore.h).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "arrow_utils.hpp"
#include "csv_loader.hpp"
#include "expression.hpp"
#include "jit.hpp"

class WarpDB {
public:
  // .csv (optional explicit schema; default every column Float32) or .json (NDJSON price/quantity)
  explicit WarpDB(const std::string &filepath, const std::vector<DataType> &schema = {});
  // adopt columns that already live on the device (benchmarks, sharded tables); not in the reference
  explicit WarpDB(Table device_table, HostTable host_table = {});
  ~WarpDB();
  WarpDB(const WarpDB &) = delete;
  WarpDB &operator=(const WarpDB &) = delete;

  // "<expr> [WHERE <cond>]" -> one float per table row; rows failing cond yield 0.0f
  std::vector<float> query(const std::string &expr);
  // SELECT [DISTINCT] <expr|AGG(expr)> FROM t [JOIN u ON t.a = u.b]... [WHERE c] [GROUP BY k] [HAVING h]
  //        [ORDER BY e [ASC|DESC]] [LIMIT n] [OFFSET m]
  std::vector<float> query_sql(const std::string &sql);
  // Make another table joinable under `name` (`... FROM t JOIN name ON t.k = name.k`).  Not in the
  // reference: its WarpDB holds one table, parses JOIN clauses (src/expression.cpp:375-401) and never
  // executes them; a JOIN naming a table that was not attached joins this table with itself
  // ("Currently JOIN loads the same table for demonstration purposes", include/warpdb.hpp:22).
  void attach(const std::string &name, const std::string &filepath, const std::vector<DataType> &schema = {});
  void attach(const std::string &name, Table device_table);
  // rows the last JOIN query produced before WHERE (-1: the last query_sql had no JOIN)
  long long last_join_rows() const { return last_join_rows_; }
  // same as query() over every visible GPU (row-range shards of the host copy of the table)
  std::vector<float> query_multi_gpu(const std::string &expr);
  // query_sql() over every visible GPU: GROUP BY aggregates and ORDER BY ... LIMIT run on row-range
  // shards and are merged GPU to GPU inside the core (NCCL over NVLink); plain SELECT compacts per
  // shard.  HAVING / DISTINCT are not offered on this path.  Not in the reference (src/warpdb.cpp:508-542
  // shards query() only).
  std::vector<float> query_sql_multi_gpu(const std::string &sql);
  // stream a CSV in chunks of rows_per_chunk rows through all GPUs
  static std::vector<float> query_multi_gpu_csv(const std::string &csv_path, const std::string &expr, int rows_per_chunk = 1000000);
  // what the last query_multi_gpu_csv call of this thread spent where: parse_ms (reading + parsing chunks, on
  // the helper thread), gpu_ms (upload + kernels + download of the chunks), wall_ms; parse_ms + gpu_ms > wall_ms
  // is the overlap
  struct CsvStreamStats { double parse_ms = 0, gpu_ms = 0, wall_ms = 0; int chunks = 0; };
  static CsvStreamStats last_csv_stream_stats();
  // query() exported through the Arrow C data interface
  void query_arrow(const std::string &expr, ArrowArray *out_array, ArrowSchema *out_schema, bool use_shared_memory = false);

  // query() whose result stays in HBM, exported through the Arrow C *device* data interface
  // (ARROW_DEVICE_CUDA): no D2H copy, buffers[1] is a device pointer owned by the array.
  // SURVEY section 8(f) item 2; not in the reference (src/arrow_utils.cpp:37-94 copies to the host).
  void query_arrow_device(const std::string &expr, ArrowDeviceArray *out_array, ArrowSchema *out_schema);

  int num_rows() const { return table_.num_rows; }
  const Table &table() const { return table_; }

  // zone maps (per-4096-row min/max) are built lazily per column and drive pruning of
  // `col <op> const` terms of WHERE clauses; disable to run every query unpruned
  void set_zone_pruning(bool on) { zone_pruning_ = on; }
  long long last_zones_live() const { return last_zones_live_; }
  long long last_zones_total() const { return last_zones_total_; }

private:
  struct PruneTerm { std::string column; int op; double value; };
  std::vector<PruneTerm> prune_terms(const ASTNode *cond) const;
  void *zonemap_for(const std::string &column);   // wdb_zonemap_t*, nullptr if not prunable
  std::vector<struct wdb_prune> prune_preds(const ASTNode *cond_ast);
  std::vector<float> run_sql(QueryAST &ast);        // everything after parsing, on table_
  std::vector<float> run_sql_join(QueryAST &ast);   // inner equi-joins first, then run_sql on the joined columns
  int filter_project(const std::string &expr, const std::string &cond, const ASTNode *cond_ast, float *d_out, int mode,
                     long long *count);

  Table table_;
  HostTable host_table_;
  bool owns_device_ = true;
  bool zone_pruning_ = true;
  std::map<std::string, void *> zonemaps_;
  std::map<std::string, std::pair<long long, long long>> key_ranges_;   // min/max of integer GROUP BY columns
  long long last_zones_live_ = -1, last_zones_total_ = -1;
  std::map<std::string, std::unique_ptr<WarpDB>> attached_;
  long long last_join_rows_ = -1;
};
