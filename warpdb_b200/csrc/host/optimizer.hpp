// optimizer.hpp -- reference: include/optimizer.hpp:7-8
#pragma once
#include <memory>
#include <string>

#include "csv_loader.hpp"
#include "expression.hpp"

void execute_query_optimized(const std::string &expr_part, const std::string &where_part, Table &table);

// What the reference's analyze_condition stub (src/optimizer.cpp:13-17) was meant to do: decide from
// column min/max whether a condition made of `col <op> const` comparisons joined by AND/OR is
// always true or always false.  ranges: name -> [min,max].
struct ColumnRange { std::string name; double min, max; };
void analyze_condition(const ASTNode *cond, const std::vector<ColumnRange> &ranges, bool &always_true, bool &always_false);
std::vector<ColumnRange> compute_column_ranges(const Table &table, int device_id = 0);
