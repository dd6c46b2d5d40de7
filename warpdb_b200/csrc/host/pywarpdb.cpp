// pywarpdb.cpp -- Python bindings with the reference's surface (bindings/python/pywarpdb.cpp:7-38:
// WarpDB(path), .query, .query_multi_gpu, static .query_multi_gpu_csv, .query_arrow -> 2 capsules)
// plus what the reference leaves unbound (query_sql, the schema constructor argument) and a few
// front-end helpers used by the parity tests.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <sstream>

#include "optimizer.hpp"
#include "warpdb.hpp"

namespace py = pybind11;

namespace {
const char *token_type_name(TokenType t) {
  switch (t) {
  case TokenType::Identifier: return "Identifier";
  case TokenType::Number: return "Number";
  case TokenType::Operator: return "Operator";
  case TokenType::Keyword: return "Keyword";
  case TokenType::End: return "End";
  }
  return "?";
}
std::string item_text(const ASTNode *n) {
  if (auto w = dynamic_cast<const WindowFunctionNode *>(n)) return "WIN" + std::to_string(static_cast<int>(w->agg)) + "(" + w->expr->to_cuda_expr() + ")";
  if (auto a = dynamic_cast<const AggregationNode *>(n)) return "AGG" + std::to_string(static_cast<int>(a->agg)) + "(" + a->expr->to_cuda_expr() + ")";
  return n->to_cuda_expr();
}
// same one-line format as oracle/ref_front_main.cpp and orc_query_summary
std::string query_summary(const std::string &sql, bool extended) {
  const QueryAST q = extended ? parse_query_extended(tokenize(sql)) : parse_query(tokenize(sql));
  std::ostringstream o;
  o << "select=[";
  for (size_t i = 0; i < q.select_list.size(); ++i) o << (i ? ";" : "") << item_text(q.select_list[i].get());
  o << "] from=" << q.from_table << " joins=[";
  for (size_t i = 0; i < q.joins.size(); ++i) o << (i ? ";" : "") << q.joins[i].table << ":" << item_text(q.joins[i].condition.get());
  o << "] where=" << (q.where ? item_text(q.where->get()) : "-") << " group=";
  if (q.group_by) {
    o << "[";
    for (size_t i = 0; i < q.group_by->keys.size(); ++i) o << (i ? ";" : "") << item_text(q.group_by->keys[i].get());
    o << "]";
  } else o << "-";
  o << " having=" << (q.having ? item_text(q.having->get()) : "-") << " order=";
  if (q.order_by) o << item_text(q.order_by->expr.get()) << (q.order_by->ascending ? ":ASC" : ":DESC"); else o << "-";
  o << " limit=" << (q.limit ? std::to_string(q.limit->count) : "-") << " offset=" << (q.offset ? std::to_string(q.offset->count) : "-")
    << " distinct=" << (q.distinct ? 1 : 0);
  return o.str();
}
}  // namespace

PYBIND11_MODULE(pywarpdb, m) {
  m.doc() = "WarpDB on the B200-native execution core";
  py::enum_<DataType>(m, "DataType")
      .value("Int32", DataType::Int32).value("Int64", DataType::Int64).value("Float32", DataType::Float32)
      .value("Float64", DataType::Float64).value("String", DataType::String);

  py::class_<WarpDB>(m, "WarpDB")
      .def(py::init<const std::string &>())
      .def(py::init<const std::string &, const std::vector<DataType> &>(), py::arg("filepath"), py::arg("schema"))
      .def("query", &WarpDB::query)
      .def("query_sql", &WarpDB::query_sql, py::arg("sql"))
      .def("query_multi_gpu", &WarpDB::query_multi_gpu, py::arg("expr"),
           "Execute expression using all available GPUs on the current table.")
      .def("query_sql_multi_gpu", &WarpDB::query_sql_multi_gpu, py::arg("sql"),
           "query_sql over all GPUs: partial aggregates / top-k candidates are merged GPU to GPU (NCCL).")
      .def_static("query_multi_gpu_csv", &WarpDB::query_multi_gpu_csv, py::arg("csv_path"), py::arg("expr"),
                  py::arg("rows_per_chunk") = 1000000, "Stream a CSV file in chunks across all GPUs and return results.")
      .def_static("last_csv_stream_stats",
                  []() {
                    const WarpDB::CsvStreamStats st = WarpDB::last_csv_stream_stats();
                    py::dict d;
                    d["parse_ms"] = st.parse_ms;
                    d["gpu_ms"] = st.gpu_ms;
                    d["wall_ms"] = st.wall_ms;
                    d["chunks"] = st.chunks;
                    return d;
                  },
                  "Where the last query_multi_gpu_csv call spent its time; parse_ms + gpu_ms > wall_ms is the overlap.")
      .def("query_arrow",
           [](WarpDB &db, const std::string &expr, bool shared_memory) {
             auto *arr = new ArrowArray();
             auto *schema = new ArrowSchema();
             db.query_arrow(expr, arr, schema, shared_memory);
             py::capsule array_capsule(arr, "arrow_array", [](PyObject *cap) {
               auto *a = static_cast<ArrowArray *>(PyCapsule_GetPointer(cap, "arrow_array"));
               if (a && a->release) a->release(a);
               delete a;
             });
             py::capsule schema_capsule(schema, "arrow_schema", [](PyObject *cap) {
               auto *s = static_cast<ArrowSchema *>(PyCapsule_GetPointer(cap, "arrow_schema"));
               if (s && s->release) s->release(s);
               delete s;
             });
             return py::make_tuple(array_capsule, schema_capsule);
           },
           py::arg("expr"), py::arg("shared_memory") = false,
           "Return result as Arrow C Data Interface capsules (ArrowArray, ArrowSchema).")
      .def("attach",
           [](WarpDB &db, const std::string &name, const std::string &filepath, const std::vector<DataType> &schema) { db.attach(name, filepath, schema); },
           py::arg("name"), py::arg("filepath"), py::arg("schema") = std::vector<DataType>{},
           "Load another table and make it joinable by name: ... FROM t JOIN name ON t.k = name.k")
      .def("last_join_rows", &WarpDB::last_join_rows)
      .def("num_rows", &WarpDB::num_rows)
      .def("set_zone_pruning", &WarpDB::set_zone_pruning)
      .def("last_zones_live", &WarpDB::last_zones_live)
      .def("last_zones_total", &WarpDB::last_zones_total);

  // front-end helpers (parity tests against tests/golden/frontend.json)
  m.def("expr_to_cuda", [](const std::string &text) { return parse_expression(tokenize(text))->to_cuda_expr(); });
  m.def("tokenize_dump", [](const std::string &text) {
    std::string out;
    for (const auto &t : tokenize(text))
      out += std::string(token_type_name(t.type)) + ":" + t.value + ":" + std::to_string(t.line) + ":" + std::to_string(t.column) + "\n";
    return out;
  });
  m.def("query_summary", &query_summary, py::arg("sql"), py::arg("extended") = false);
  m.def("analyze_condition", [](const std::string &cond, const std::vector<std::tuple<std::string, double, double>> &ranges) {
    std::vector<ColumnRange> r;
    for (const auto &t : ranges) r.push_back(ColumnRange{std::get<0>(t), std::get<1>(t), std::get<2>(t)});
    ASTNodePtr c = parse_expression(tokenize(cond));
    bool at = false, af = false;
    analyze_condition(c.get(), r, at, af);
    return py::make_tuple(at, af);
  });
  m.def("execute_query_optimized", [](WarpDB &db, const std::string &expr, const std::string &where) {
    Table t = db.table();
    execute_query_optimized(expr, where, t);
  });
}
