// expression.hpp -- front end of the B200 WarpDB core: tokens, AST and SQL clauses.
//
// API-compatible with the reference's include/expression.hpp (same type and member names, same
// to_cuda_expr() strings, same error messages -- pinned by tests/golden/frontend.json, which was
// generated from the reference's own parser), written from scratch: the parser is a re-entrant
// cursor object instead of file-level statics (reference: src/expression.cpp:122-127), and code
// generation lives out of line in expression.cpp.
#pragma once
#include <memory>
#include <optional>
#include <string>
#include <vector>

// ---- tokens (reference: include/expression.hpp:7-16) -----------------------------------------
enum class TokenType { Identifier, Number, Operator, Keyword, End };

struct Token {
  TokenType type;
  std::string value;
  int line = 1;
  int column = 1;
};

std::vector<Token> tokenize(const std::string &input);

// ---- expression tree (reference: include/expression.hpp:18-97,112-121) -------------------------
enum class ASTNodeType { Constant, Variable, BinaryOp, FunctionCall, Aggregation };
enum class AggregationType { Sum, Avg, Count, Min, Max };

struct ASTNode {
  virtual ~ASTNode();
  // CUDA-C text of the node: `col[idx]` for columns, float literals with an `f` suffix
  virtual std::string to_cuda_expr() const = 0;
  virtual ASTNodeType type() const = 0;
};
using ASTNodePtr = std::unique_ptr<ASTNode>;

struct ConstantNode : ASTNode {
  std::string value;
  explicit ConstantNode(const std::string &val);
  std::string to_cuda_expr() const override;
  ASTNodeType type() const override;
};

struct VariableNode : ASTNode {
  std::string name;
  explicit VariableNode(const std::string &n);
  std::string to_cuda_expr() const override;
  ASTNodeType type() const override;
};

struct BinaryOpNode : ASTNode {
  std::string op;
  ASTNodePtr left;
  ASTNodePtr right;
  BinaryOpNode(std::string o, ASTNodePtr l, ASTNodePtr r);
  std::string to_cuda_expr() const override;
  ASTNodeType type() const override;
};

struct FunctionCallNode : ASTNode {
  std::string name;
  std::vector<ASTNodePtr> args;
  FunctionCallNode(std::string n, std::vector<ASTNodePtr> a);
  std::string to_cuda_expr() const override;
  ASTNodeType type() const override;
};

struct AggregationNode : ASTNode {
  AggregationType agg;
  ASTNodePtr expr;
  AggregationNode(AggregationType a, ASTNodePtr e);
  std::string to_cuda_expr() const override;   // the aggregated expression
  ASTNodeType type() const override;
  std::string agg_kernel() const;              // "sum" | "avg" | "count" | "min" | "max"
};

struct OrderByClause {
  ASTNodePtr expr;
  bool ascending;
};
struct LimitClause { int count; };
struct OffsetClause { int count; };

struct WindowFunctionNode : ASTNode {
  AggregationType agg;
  ASTNodePtr expr;
  std::vector<ASTNodePtr> partition_by;
  std::optional<OrderByClause> order_by;
  WindowFunctionNode(AggregationType a, ASTNodePtr e);
  std::string to_cuda_expr() const override;   // "<window>"
  ASTNodeType type() const override;
};

// ---- expression entry points (reference: include/expression.hpp:83-85) -------------------------
ASTNodePtr parse_expression(const std::vector<Token> &tokens);
ASTNodePtr parse_logical_and(const std::vector<Token> &tokens);
ASTNodePtr parse_logical_or(const std::vector<Token> &tokens);

// ---- SELECT statement (reference: include/expression.hpp:123-145) ------------------------------
struct JoinClause {
  std::string table;
  ASTNodePtr condition;
};
struct GroupByClause { std::vector<ASTNodePtr> keys; };

struct QueryAST {
  std::vector<ASTNodePtr> select_list;
  std::string from_table;
  std::vector<JoinClause> joins;
  std::optional<ASTNodePtr> where;
  std::optional<GroupByClause> group_by;
  std::optional<ASTNodePtr> having;
  std::optional<OrderByClause> order_by;
  std::optional<LimitClause> limit;
  std::optional<OffsetClause> offset;
  bool distinct = false;
};

// Strict: accepts exactly what the reference's parse_query accepts (src/expression.cpp:270-531).
QueryAST parse_query(const std::vector<Token> &tokens);
// Superset used by WarpDB::query_sql: additionally accepts what the reference's own tests write but
// its grammar rejects (tests/sql_features_test.cpp:33,36; tests/having_distinct_test.cpp:7):
// aggregates inside HAVING, OFFSET before LIMIT, and ORDER BY <expr> without ASC/DESC followed by
// LIMIT/OFFSET, and a HAVING clause that ends the statement.
QueryAST parse_query_extended(const std::vector<Token> &tokens);
