// optimizer.cpp -- statistics-driven pruning in front of the fused kernel.
// Reference: src/optimizer.cpp:21-61 (its analyze_condition, :13-17, ignores its inputs and
// TableStats is never filled in).  Here column min/max come from a device reduction
// (wdb_column_minmax) and decide conditions of the shape `col <op> const` joined by AND / OR.
#include "optimizer.hpp"

#include <cuda_runtime.h>

#include <iostream>
#include <stdexcept>
#include <vector>

#include "jit.hpp"
#include "warpcore.h"

namespace {
enum class Tri { False, True, Unknown };

Tri tri_and(Tri a, Tri b) {
  if (a == Tri::False || b == Tri::False) return Tri::False;
  if (a == Tri::True && b == Tri::True) return Tri::True;
  return Tri::Unknown;
}
Tri tri_or(Tri a, Tri b) {
  if (a == Tri::True || b == Tri::True) return Tri::True;
  if (a == Tri::False && b == Tri::False) return Tri::False;
  return Tri::Unknown;
}
const ColumnRange *find_range(const std::vector<ColumnRange> &ranges, const std::string &name) {
  for (const auto &r : ranges)
    if (r.name == name) return &r;
  return nullptr;
}
// value range [lo,hi] of a column compared with constant c (compared as floats, like the kernel)
Tri decide(const std::string &op, double lo, double hi, double c) {
  if (op == ">") return lo > c ? Tri::True : (hi <= c ? Tri::False : Tri::Unknown);
  if (op == ">=") return lo >= c ? Tri::True : (hi < c ? Tri::False : Tri::Unknown);
  if (op == "<") return hi < c ? Tri::True : (lo >= c ? Tri::False : Tri::Unknown);
  if (op == "<=") return hi <= c ? Tri::True : (lo > c ? Tri::False : Tri::Unknown);
  if (op == "==") return (lo == c && hi == c) ? Tri::True : ((c < lo || c > hi) ? Tri::False : Tri::Unknown);
  if (op == "!=") return (c < lo || c > hi) ? Tri::True : ((lo == c && hi == c) ? Tri::False : Tri::Unknown);
  return Tri::Unknown;
}
std::string mirrored(const std::string &op) {
  if (op == ">") return "<";
  if (op == "<") return ">";
  if (op == ">=") return "<=";
  if (op == "<=") return ">=";
  return op;
}
Tri eval_tri(const ASTNode *n, const std::vector<ColumnRange> &ranges) {
  const auto *b = dynamic_cast<const BinaryOpNode *>(n);
  if (!b) return Tri::Unknown;
  if (b->op == "&&") return tri_and(eval_tri(b->left.get(), ranges), eval_tri(b->right.get(), ranges));
  if (b->op == "||") return tri_or(eval_tri(b->left.get(), ranges), eval_tri(b->right.get(), ranges));
  const auto *lv = dynamic_cast<const VariableNode *>(b->left.get());
  const auto *rv = dynamic_cast<const VariableNode *>(b->right.get());
  const auto *lc = dynamic_cast<const ConstantNode *>(b->left.get());
  const auto *rc = dynamic_cast<const ConstantNode *>(b->right.get());
  if (lv && rc) {
    if (const ColumnRange *r = find_range(ranges, lv->name)) return decide(b->op, r->min, r->max, std::stof(rc->value));
  } else if (lc && rv) {
    if (const ColumnRange *r = find_range(ranges, rv->name)) return decide(mirrored(b->op), r->min, r->max, std::stof(lc->value));
  }
  return Tri::Unknown;
}
}  // namespace

void analyze_condition(const ASTNode *cond, const std::vector<ColumnRange> &ranges, bool &always_true, bool &always_false) {
  const Tri t = cond ? eval_tri(cond, ranges) : Tri::True;
  always_true = t == Tri::True;
  always_false = t == Tri::False;
}

std::vector<ColumnRange> compute_column_ranges(const Table &table, int device_id) {
  std::vector<ColumnRange> out;
  for (const auto &c : table.columns) {
    if (c.type == DataType::String || !c.device_ptr || table.num_rows == 0) continue;
    const wdb_col_t col{c.name.c_str(), static_cast<int>(c.type), c.device_ptr, table.num_rows};
    double lo = 0, hi = 0;
    if (wdb_column_minmax(device_id, nullptr, &col, &lo, &hi)) throw std::runtime_error(wdb_last_error());
    out.push_back(ColumnRange{c.name, lo, hi});
  }
  return out;
}

void execute_query_optimized(const std::string &expr_part, const std::string &where_part, Table &table) {
  ASTNodePtr expr_ast = parse_expression(tokenize(expr_part));
  ASTNodePtr cond_ast;
  if (!where_part.empty()) cond_ast = parse_expression(tokenize(where_part));

  bool always_true = false, always_false = false;
  if (cond_ast) analyze_condition(cond_ast.get(), compute_column_ranges(table), always_true, always_false);
  if (always_false) {
    std::cout << "[Optimizer] Filter eliminates all rows.\n";   // src/optimizer.cpp:38-41
    return;
  }
  const std::string expr_cuda = expr_ast->to_cuda_expr();
  std::string cond_cuda;
  if (cond_ast && !always_true) cond_cuda = cond_ast->to_cuda_expr();

  float *d_output = nullptr;
  const size_t n = static_cast<size_t>(table.num_rows);
  if (cudaMalloc(&d_output, sizeof(float) * (n ? n : 1)) != cudaSuccess) throw std::runtime_error("CUDA error: out of memory");
  cudaMemset(d_output, 0, sizeof(float) * (n ? n : 1));
  try {
    jit_compile_and_launch(expr_cuda, cond_cuda, table, d_output);
  } catch (...) {
    cudaFree(d_output);
    throw;
  }
  std::vector<float> h(n);
  cudaMemcpy(h.data(), d_output, sizeof(float) * n, cudaMemcpyDeviceToHost);
  cudaFree(d_output);
  for (size_t i = 0; i < n; ++i) std::cout << "Result[" << i << "] = " << h[i] << "\n";   // src/optimizer.cpp:56-58
}
