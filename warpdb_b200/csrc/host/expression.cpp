// expression.cpp -- tokenizer, precedence-climbing expression parser and SELECT clause parser.
// Behaviour (accepted language, AST shape, generated CUDA text, exception messages) follows the
// reference's src/expression.cpp and include/expression.hpp; see expression.hpp.
#include "expression.hpp"

#include <array>
#include <cctype>
#include <stdexcept>
#include <utility>

// ------------------------------------------------------------------------------------------------
// nodes
// ------------------------------------------------------------------------------------------------
ASTNode::~ASTNode() = default;

ConstantNode::ConstantNode(const std::string &val) : value(val) {}
ASTNodeType ConstantNode::type() const { return ASTNodeType::Constant; }
std::string ConstantNode::to_cuda_expr() const {
  // every literal becomes a float literal: "10" -> "10.0f", "0.9" -> "0.9f" (expression.hpp:32-38)
  const bool has_point = value.find('.') != std::string::npos;
  return has_point ? value + "f" : value + ".0f";
}

VariableNode::VariableNode(const std::string &n) : name(n) {}
ASTNodeType VariableNode::type() const { return ASTNodeType::Variable; }
std::string VariableNode::to_cuda_expr() const { return name + "[idx]"; }

BinaryOpNode::BinaryOpNode(std::string o, ASTNodePtr l, ASTNodePtr r) : op(std::move(o)), left(std::move(l)), right(std::move(r)) {}
ASTNodeType BinaryOpNode::type() const { return ASTNodeType::BinaryOp; }
std::string BinaryOpNode::to_cuda_expr() const {
  std::string s = "(";
  s += left->to_cuda_expr();
  s += ' ';
  s += op;
  s += ' ';
  s += right->to_cuda_expr();
  s += ')';
  return s;
}

FunctionCallNode::FunctionCallNode(std::string n, std::vector<ASTNodePtr> a) : name(std::move(n)), args(std::move(a)) {}
ASTNodeType FunctionCallNode::type() const { return ASTNodeType::FunctionCall; }
std::string FunctionCallNode::to_cuda_expr() const {
  std::string s = name + "(";
  const char *sep = "";
  for (const auto &a : args) {
    s += sep;
    s += a->to_cuda_expr();
    sep = ", ";
  }
  return s + ")";
}

AggregationNode::AggregationNode(AggregationType a, ASTNodePtr e) : agg(a), expr(std::move(e)) {}
ASTNodeType AggregationNode::type() const { return ASTNodeType::Aggregation; }
std::string AggregationNode::to_cuda_expr() const { return expr->to_cuda_expr(); }
std::string AggregationNode::agg_kernel() const {
  static const std::array<const char *, 5> names = {"sum", "avg", "count", "min", "max"};
  const auto i = static_cast<size_t>(agg);
  return i < names.size() ? names[i] : "";
}

WindowFunctionNode::WindowFunctionNode(AggregationType a, ASTNodePtr e) : agg(a), expr(std::move(e)) {}
ASTNodeType WindowFunctionNode::type() const { return ASTNodeType::Aggregation; }
std::string WindowFunctionNode::to_cuda_expr() const { return "<window>"; }

// ------------------------------------------------------------------------------------------------
// tokenizer
// ------------------------------------------------------------------------------------------------
namespace {

const char *type_label(TokenType t) {
  switch (t) {
  case TokenType::Identifier: return "Identifier";
  case TokenType::Number: return "Number";
  case TokenType::Operator: return "Operator";
  case TokenType::Keyword: return "Keyword";
  case TokenType::End: return "End";
  }
  return "Unknown";
}

bool is_sql_keyword(const std::string &upper) {
  static const char *const words[] = {"SELECT", "FROM", "WHERE", "JOIN", "ON", "GROUP", "BY", "ORDER", "ASC", "DESC", "LIMIT",
                                      "OFFSET", "SUM", "AVG", "COUNT", "MIN", "MAX", "OVER", "PARTITION", "AND", "OR", "HAVING",
                                      "DISTINCT"};
  for (const char *w : words)
    if (upper == w) return true;
  return false;
}

// Character cursor that tracks 1-based line/column like the reference's error messages expect.
class Scanner {
public:
  explicit Scanner(const std::string &text) : s_(text) {}
  bool done() const { return i_ >= s_.size(); }
  char cur() const { return s_[i_]; }
  bool has_next() const { return i_ + 1 < s_.size(); }
  char next() const { return s_[i_ + 1]; }
  int line() const { return line_; }
  int column() const { return col_; }
  char take() {
    const char c = s_[i_++];
    if (c == '\n') { ++line_; col_ = 1; } else ++col_;
    return c;
  }
private:
  const std::string &s_;
  size_t i_ = 0;
  int line_ = 1, col_ = 1;
};

bool ident_start(char c) { return std::isalpha(static_cast<unsigned char>(c)) || c == '_'; }
bool ident_part(char c) { return std::isalnum(static_cast<unsigned char>(c)) || c == '_' || c == '.'; }
bool digit(char c) { return std::isdigit(static_cast<unsigned char>(c)) != 0; }

}  // namespace

std::vector<Token> tokenize(const std::string &input) {
  std::vector<Token> out;
  Scanner sc(input);
  while (!sc.done()) {
    const char c = sc.cur();
    if (std::isspace(static_cast<unsigned char>(c))) { sc.take(); continue; }
    Token t{TokenType::End, "", sc.line(), sc.column()};
    if (ident_start(c)) {
      while (!sc.done() && ident_part(sc.cur())) t.value += sc.take();
      std::string upper = t.value;
      for (auto &ch : upper) ch = static_cast<char>(std::toupper(static_cast<unsigned char>(ch)));
      if (is_sql_keyword(upper)) { t.type = TokenType::Keyword; t.value = upper; }
      else t.type = TokenType::Identifier;
    } else if (digit(c) || (c == '.' && sc.has_next() && digit(sc.next()))) {
      // digits with at most one decimal point; no sign, no exponent
      bool seen_point = false;
      while (!sc.done() && (digit(sc.cur()) || (sc.cur() == '.' && !seen_point))) {
        seen_point = seen_point || sc.cur() == '.';
        t.value += sc.take();
      }
      t.type = TokenType::Number;
    } else if (c == '>' || c == '<' || c == '=' || c == '!') {
      t.value += sc.take();
      if (!sc.done() && sc.cur() == '=') t.value += sc.take();
      t.type = TokenType::Operator;
    } else if (std::string("+-*/(),.").find(c) != std::string::npos) {
      t.value += sc.take();
      t.type = TokenType::Operator;
    } else {
      throw std::runtime_error("Unknown character '" + std::string(1, c) + "' at line " + std::to_string(sc.line()) +
                               " column " + std::to_string(sc.column()));
    }
    out.push_back(std::move(t));
  }
  out.push_back(Token{TokenType::End, "", sc.line(), sc.column()});
  return out;
}

// ------------------------------------------------------------------------------------------------
// expression parser: or > and > comparison > additive > multiplicative > factor
// (parentheses and call arguments restart at the additive level, as in the reference)
// ------------------------------------------------------------------------------------------------
namespace {

bool is_op(const Token &t, const char *v) { return t.type == TokenType::Operator && t.value == v; }
bool is_kw(const Token &t, const char *v) { return t.type == TokenType::Keyword && t.value == v; }

bool aggregate_of(const std::string &kw, AggregationType *out) {
  static const std::pair<const char *, AggregationType> table[] = {{"SUM", AggregationType::Sum}, {"AVG", AggregationType::Avg},
                                                                    {"COUNT", AggregationType::Count}, {"MIN", AggregationType::Min},
                                                                    {"MAX", AggregationType::Max}};
  for (const auto &e : table)
    if (kw == e.first) { *out = e.second; return true; }
  return false;
}

class ExprParser {
public:
  // [first,last) is the token window; an End token is implied after it
  ExprParser(const Token *first, const Token *last, bool allow_aggregates)
      : cur_(first), end_(last), aggregates_(allow_aggregates) {}

  ASTNodePtr parse_all(int top_level) {
    ASTNodePtr n = top_level == 0 ? disjunction() : conjunction();
    if (peek().type != TokenType::End) throw std::runtime_error("Unexpected tokens remaining: " + peek().value);
    return n;
  }

private:
  const Token *cur_, *end_;
  bool aggregates_;
  Token sentinel_{TokenType::End, "", 0, 0};

  const Token &peek() const { return cur_ < end_ ? *cur_ : sentinel_; }
  bool accept(const char *op) {
    if (!is_op(peek(), op)) return false;
    ++cur_;
    return true;
  }
  // tries each operator of a level in order; returns the matched text or nullptr
  template <size_t N> const char *accept_any(const char *const (&ops)[N]) {
    for (const char *o : ops)
      if (accept(o)) return o;
    return nullptr;
  }
  template <size_t N> ASTNodePtr left_assoc(ASTNodePtr (ExprParser::*operand)(), const char *const (&ops)[N]) {
    ASTNodePtr lhs = (this->*operand)();
    while (const char *o = accept_any(ops)) {
      ASTNodePtr rhs = (this->*operand)();
      lhs = std::make_unique<BinaryOpNode>(o, std::move(lhs), std::move(rhs));
    }
    return lhs;
  }
  ASTNodePtr keyword_chain(ASTNodePtr (ExprParser::*operand)(), const char *keyword, const char *emitted) {
    ASTNodePtr lhs = (this->*operand)();
    while (is_kw(peek(), keyword)) {
      ++cur_;
      ASTNodePtr rhs = (this->*operand)();
      lhs = std::make_unique<BinaryOpNode>(emitted, std::move(lhs), std::move(rhs));
    }
    return lhs;
  }

  ASTNodePtr disjunction() { return keyword_chain(&ExprParser::conjunction, "OR", "||"); }
  ASTNodePtr conjunction() { return keyword_chain(&ExprParser::comparison, "AND", "&&"); }
  ASTNodePtr comparison() {
    static const char *const ops[] = {">", "<", ">=", "<=", "==", "!=", "="};
    return left_assoc(&ExprParser::additive, ops);
  }
  ASTNodePtr additive() {
    static const char *const ops[] = {"+", "-"};
    return left_assoc(&ExprParser::multiplicative, ops);
  }
  ASTNodePtr multiplicative() {
    static const char *const ops[] = {"*", "/"};
    return left_assoc(&ExprParser::factor, ops);
  }
  ASTNodePtr factor() {
    const Token &t = peek();
    if (t.type == TokenType::Number) {
      ++cur_;
      return std::make_unique<ConstantNode>(t.value);
    }
    if (t.type == TokenType::Identifier) {
      ++cur_;
      if (!accept("(")) return std::make_unique<VariableNode>(t.value);
      std::vector<ASTNodePtr> args;
      if (!accept(")")) {
        do args.push_back(additive()); while (accept(","));
        if (!accept(")")) throw std::runtime_error("Expected ')' after arguments");
      }
      return std::make_unique<FunctionCallNode>(t.value, std::move(args));
    }
    AggregationType agg;
    if (aggregates_ && t.type == TokenType::Keyword && aggregate_of(t.value, &agg)) {   // HAVING extension
      ++cur_;
      if (!accept("(")) throw std::runtime_error("Invalid syntax for " + t.value + " aggregation");
      ASTNodePtr inner = additive();
      if (!accept(")")) throw std::runtime_error("Expected ')'");
      return std::make_unique<AggregationNode>(agg, std::move(inner));
    }
    if (accept("(")) {
      ASTNodePtr inner = additive();
      if (!accept(")")) throw std::runtime_error("Expected ')'");
      return inner;
    }
    throw std::runtime_error(std::string("Unexpected token (") + type_label(t.type) + ": " + t.value + ")");
  }
};

// the public entry points receive a vector that normally ends with an End token
ASTNodePtr parse_vector(const std::vector<Token> &tokens, int top_level) {
  const Token *b = tokens.data(), *e = b + tokens.size();
  for (const Token *p = b; p < e; ++p)
    if (p->type == TokenType::End) { e = p; break; }
  return ExprParser(b, e, false).parse_all(top_level);
}

ASTNodePtr parse_window(const Token *b, const Token *e, bool aggregates = false) {
  for (const Token *p = b; p < e; ++p)
    if (p->type == TokenType::End) { e = p; break; }
  return ExprParser(b, e, aggregates).parse_all(0);
}

}  // namespace

ASTNodePtr parse_expression(const std::vector<Token> &tokens) { return parse_vector(tokens, 0); }
ASTNodePtr parse_logical_or(const std::vector<Token> &tokens) { return parse_vector(tokens, 0); }
ASTNodePtr parse_logical_and(const std::vector<Token> &tokens) { return parse_vector(tokens, 1); }

// ------------------------------------------------------------------------------------------------
// SELECT statement
// ------------------------------------------------------------------------------------------------
namespace {

class QueryParser {
public:
  QueryParser(const std::vector<Token> &tokens, bool extended) : t_(tokens), ext_(extended) {
    size_ = t_.size();
    end_ = size_;
    if (end_ > 0 && t_[end_ - 1].type == TokenType::End) --end_;
  }

  QueryAST run() {
    QueryAST q;
    expect_keyword("SELECT");
    if (pos_ < size_ && is_kw(t_[pos_], "DISTINCT")) { q.distinct = true; ++pos_; }
    select_list(q);
    expect_keyword("FROM");
    if (pos_ >= size_ || t_[pos_].type != TokenType::Identifier) throw located("Expected table name after FROM");
    q.from_table = t_[pos_++].value;
    joins(q);
    if (pos_ < end_ && is_kw(t_[pos_], "WHERE")) {
      ++pos_;
      const size_t start = pos_;
      skip_until({"GROUP", "ORDER", "HAVING", "LIMIT"}, end_);
      q.where = parse_window(&t_[0] + start, &t_[0] + pos_);
    }
    if (pos_ < end_ && is_kw(t_[pos_], "GROUP")) group_by(q);
    having(q);
    if (pos_ < size_ && is_kw(t_[pos_], "ORDER")) order_by(q);
    limit_offset(q);
    if (pos_ != end_) throw std::runtime_error("Unexpected token in query near: " + (pos_ < size_ ? t_[pos_].value : std::string()));
    return q;
  }

private:
  const std::vector<Token> &t_;
  bool ext_;
  size_t size_ = 0, end_ = 0, pos_ = 0;

  std::runtime_error located(const std::string &what) const {
    const Token &ref = pos_ < size_ ? t_[pos_] : t_.back();
    return std::runtime_error(what + " at line " + std::to_string(ref.line) + " column " + std::to_string(ref.column));
  }
  void expect_keyword(const char *kw) {
    if (pos_ >= size_ || !is_kw(t_[pos_], kw)) throw located(std::string("Expected keyword '") + kw + "'");
    ++pos_;
  }
  bool at_any_keyword(std::initializer_list<const char *> kws) const {
    if (t_[pos_].type != TokenType::Keyword) return false;
    for (const char *k : kws)
      if (t_[pos_].value == k) return true;
    return false;
  }
  void skip_until(std::initializer_list<const char *> kws, size_t limit) {
    while (pos_ < limit && !at_any_keyword(kws)) ++pos_;
  }

  ASTNodePtr select_item(size_t b, size_t e) {
    AggregationType agg;
    if (e > b && t_[b].type == TokenType::Keyword && aggregate_of(t_[b].value, &agg)) {
      size_t over = e;
      for (size_t i = b; i < e; ++i)
        if (is_kw(t_[i], "OVER")) { over = i; break; }
      const bool call_shape = over - b > 1 && is_op(t_[b + 1], "(") && is_op(t_[over - 1], ")");
      if (!call_shape) throw std::runtime_error("Invalid syntax for " + t_[b].value + " aggregation");
      ASTNodePtr inner = parse_window(&t_[0] + b + 2, &t_[0] + over - 1);
      if (over < e) return std::make_unique<WindowFunctionNode>(agg, std::move(inner));
      return std::make_unique<AggregationNode>(agg, std::move(inner));
    }
    return parse_window(&t_[0] + b, &t_[0] + e);
  }

  void select_list(QueryAST &q) {
    while (pos_ < end_ && !is_kw(t_[pos_], "FROM")) {
      const size_t start = pos_;
      int depth = 0;
      for (; pos_ < end_; ++pos_) {
        if (is_op(t_[pos_], "(")) ++depth;
        if (is_op(t_[pos_], ")")) --depth;
        if (depth == 0 && (is_op(t_[pos_], ",") || is_kw(t_[pos_], "FROM"))) break;
      }
      q.select_list.push_back(select_item(start, pos_));
      if (pos_ < end_ && is_op(t_[pos_], ",")) ++pos_;
    }
  }

  void joins(QueryAST &q) {
    while (pos_ < size_ && is_kw(t_[pos_], "JOIN")) {
      ++pos_;
      if (pos_ >= size_ || t_[pos_].type != TokenType::Identifier) throw located("Expected table name after JOIN");
      JoinClause jc;
      jc.table = t_[pos_++].value;
      expect_keyword("ON");
      const size_t start = pos_;
      skip_until({"WHERE", "GROUP", "ORDER", "HAVING", "JOIN", "LIMIT"}, end_);
      jc.condition = parse_window(&t_[0] + start, &t_[0] + pos_);
      q.joins.push_back(std::move(jc));
    }
  }

  void group_by(QueryAST &q) {
    ++pos_;
    expect_keyword("BY");
    GroupByClause gb;
    while (pos_ < end_) {
      const size_t start = pos_;
      while (pos_ < end_ && !is_op(t_[pos_], ",") && !at_any_keyword({"ORDER", "HAVING"}) &&
             !(ext_ && at_any_keyword({"LIMIT", "OFFSET"})))
        ++pos_;
      gb.keys.push_back(parse_window(&t_[0] + start, &t_[0] + pos_));
      if (pos_ < end_ && is_op(t_[pos_], ",")) ++pos_;
      if (pos_ < size_ && (at_any_keyword({"ORDER", "HAVING"}) || (ext_ && at_any_keyword({"LIMIT", "OFFSET"})))) break;
    }
    q.group_by = std::move(gb);
  }

  void having(QueryAST &q) {
    // the reference scans twice (second time also stopping at OFFSET) and lets the scan run over the
    // End token, after which its trailing-token check reads past the vector; strict mode reports
    // that as a trailing-token error with an empty lexeme, extended mode stops at the End token
    for (int pass = 0; pass < 2; ++pass) {
      if (!(pos_ < size_ && is_kw(t_[pos_], "HAVING"))) continue;
      ++pos_;
      const size_t start = pos_;
      const size_t limit = ext_ ? end_ : size_;
      while (pos_ < limit && !(at_any_keyword({"ORDER", "LIMIT"}) || ((pass == 1 || ext_) && at_any_keyword({"OFFSET"})))) ++pos_;
      q.having = parse_window(&t_[0] + start, &t_[0] + std::min(pos_, end_), ext_);
    }
  }

  void order_by(QueryAST &q) {
    ++pos_;
    expect_keyword("BY");
    const size_t start = pos_;
    while (pos_ < end_ && !at_any_keyword({"ASC", "DESC"}) && !(ext_ && at_any_keyword({"LIMIT", "OFFSET"}))) ++pos_;
    OrderByClause ob;
    ob.expr = parse_window(&t_[0] + start, &t_[0] + pos_);
    ob.ascending = true;
    if (pos_ < end_ && at_any_keyword({"ASC", "DESC"})) {
      ob.ascending = t_[pos_].value == "ASC";
      ++pos_;
    }
    q.order_by = std::move(ob);
  }

  void limit_clause(QueryAST &q) {
    ++pos_;
    if (pos_ >= size_ || t_[pos_].type != TokenType::Number) throw located("Expected numeric value after LIMIT");
    q.limit = LimitClause{std::stoi(t_[pos_].value)};
    ++pos_;
  }
  void offset_clause(QueryAST &q) {
    ++pos_;
    if (pos_ >= size_ || t_[pos_].type != TokenType::Number) throw std::runtime_error("Expected numeric value after OFFSET");
    q.offset = OffsetClause{std::stoi(t_[pos_].value)};
    ++pos_;
  }
  void limit_offset(QueryAST &q) {
    if (ext_ && pos_ < end_ && is_kw(t_[pos_], "OFFSET")) {     // OFFSET n LIMIT m (tests/sql_features_test.cpp:33)
      offset_clause(q);
      if (pos_ < end_ && is_kw(t_[pos_], "LIMIT")) limit_clause(q);
      return;
    }
    if (pos_ < end_ && is_kw(t_[pos_], "LIMIT")) limit_clause(q);
    if (pos_ < size_ && is_kw(t_[pos_], "OFFSET")) offset_clause(q);
  }
};

}  // namespace

QueryAST parse_query(const std::vector<Token> &tokens) { return QueryParser(tokens, false).run(); }
QueryAST parse_query_extended(const std::vector<Token> &tokens) { return QueryParser(tokens, true).run(); }
