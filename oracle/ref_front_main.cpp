// ref_front_main.cpp -- driver around the REFERENCE's own front end (src/expression.cpp,
// include/expression.hpp), used only to generate tests/golden/frontend.json in the build
// container.  Test infrastructure; the reference sources are compiled from where they lie
// under /root/reference (see oracle/Makefile), never copied into this repository.
//
// stdin: one request per line, "<K> <text>" with K in {T,E,Q}; "\n" in text is an escaped newline.
// stdout: one line per request: "OK <payload>" or "ERR <message>"; newlines escaped as "\n".
#include "expression.hpp"
#include <iostream>
#include <sstream>
#include <string>

static std::string unescape(const std::string &s) {
  std::string o;
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '\\' && i + 1 < s.size() && s[i + 1] == 'n') { o += '\n'; ++i; }
    else o += s[i];
  }
  return o;
}
static std::string escape(const std::string &s) {
  std::string o;
  for (char c : s) { if (c == '\n') o += "\\n"; else o += c; }
  return o;
}
static const char *tname(TokenType t) {
  switch (t) {
  case TokenType::Identifier: return "Identifier";
  case TokenType::Number: return "Number";
  case TokenType::Operator: return "Operator";
  case TokenType::Keyword: return "Keyword";
  case TokenType::End: return "End";
  }
  return "?";
}
static std::string item(const ASTNode *n) {
  if (auto w = dynamic_cast<const WindowFunctionNode *>(n))
    return "WIN" + std::to_string(static_cast<int>(w->agg)) + "(" + w->expr->to_cuda_expr() + ")";
  if (auto a = dynamic_cast<const AggregationNode *>(n))
    return "AGG" + std::to_string(static_cast<int>(a->agg)) + "(" + a->expr->to_cuda_expr() + ")";
  return n->to_cuda_expr();
}
static std::string summary(const QueryAST &q) {
  std::ostringstream o;
  o << "select=[";
  for (size_t i = 0; i < q.select_list.size(); ++i) { if (i) o << ";"; o << item(q.select_list[i].get()); }
  o << "] from=" << q.from_table << " joins=[";
  for (size_t i = 0; i < q.joins.size(); ++i) { if (i) o << ";"; o << q.joins[i].table << ":" << item(q.joins[i].condition.get()); }
  o << "] where=";
  if (q.where) o << item(q.where->get()); else o << "-";
  o << " group=";
  if (q.group_by) {
    o << "[";
    for (size_t i = 0; i < q.group_by->keys.size(); ++i) { if (i) o << ";"; o << item(q.group_by->keys[i].get()); }
    o << "]";
  } else o << "-";
  o << " having=";
  if (q.having) o << item(q.having->get()); else o << "-";
  o << " order=";
  if (q.order_by) o << item(q.order_by->expr.get()) << (q.order_by->ascending ? ":ASC" : ":DESC"); else o << "-";
  o << " limit=";
  if (q.limit) o << q.limit->count; else o << "-";
  o << " offset=";
  if (q.offset) o << q.offset->count; else o << "-";
  o << " distinct=" << (q.distinct ? 1 : 0);
  return o.str();
}
int main() {
  std::string line;
  while (std::getline(std::cin, line)) {
    if (line.size() < 2) continue;
    char k = line[0];
    std::string text = unescape(line.substr(2));
    try {
      if (k == 'T') {
        std::string out;
        for (const auto &t : tokenize(text))
          out += std::string(tname(t.type)) + ":" + t.value + ":" + std::to_string(t.line) + ":" + std::to_string(t.column) + "\n";
        std::cout << "OK " << escape(out) << "\n";
      } else if (k == 'E') {
        auto ast = parse_expression(tokenize(text));
        std::cout << "OK " << escape(ast->to_cuda_expr()) << "\n";
      } else if (k == 'Q') {
        QueryAST q = parse_query(tokenize(text));
        std::cout << "OK " << escape(summary(q)) << "\n";
      }
    } catch (const std::exception &e) {
      std::cout << "ERR " << escape(e.what()) << "\n";
    }
  }
  return 0;
}
