/*
 * wdb_oracle.c -- CPU ORACLE (test infrastructure, not product code).  See wdb_oracle.h for
 * the list of reference files each part restates and for the parity-pinning status.
 *
 * Plain C11, no dependencies beyond libc/libm/pthreads.
 */
#define _GNU_SOURCE
#include "wdb_oracle.h"
#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * small helpers
 * ---------------------------------------------------------------------------------------- */
static void set_err(char *err, size_t errlen, const char *fmt, ...) {
  if (!err || !errlen) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, errlen, fmt, ap);
  va_end(ap);
}
static char *xstrndup(const char *s, size_t n) {
  char *r = (char *)malloc(n + 1);
  memcpy(r, s, n);
  r[n] = 0;
  return r;
}
static char *xstrdup(const char *s) { return xstrndup(s, strlen(s)); }

typedef struct { char *p; size_t len, cap; int fixed; } sbuf;
static void sb_put(sbuf *b, const char *s) {
  size_t n = strlen(s);
  if (b->fixed) { /* counting writer into a caller buffer */
    if (b->p && b->len < b->cap) {
      size_t room = b->cap - 1 - b->len < n ? b->cap - 1 - b->len : n;
      if (b->cap > 0 && b->len < b->cap - 1) memcpy(b->p + b->len, s, room);
    }
    b->len += n;
    return;
  }
  if (b->len + n + 1 > b->cap) {
    b->cap = (b->len + n + 1) * 2;
    b->p = (char *)realloc(b->p, b->cap);
  }
  memcpy(b->p + b->len, s, n + 1);
  b->len += n;
}
static size_t sb_finish_fixed(sbuf *b) {
  if (b->p && b->cap) b->p[b->len < b->cap ? b->len : b->cap - 1] = 0;
  return b->len;
}

/* ------------------------------------------------------------------------------------------
 * tokenizer  (reference: src/expression.cpp:22-120)
 * ---------------------------------------------------------------------------------------- */
enum { T_IDENT, T_NUMBER, T_OP, T_KEYWORD, T_END };
static const char *tok_type_name(int t) { /* expression.cpp:6-20 */
  switch (t) {
  case T_IDENT: return "Identifier";
  case T_NUMBER: return "Number";
  case T_OP: return "Operator";
  case T_KEYWORD: return "Keyword";
  case T_END: return "End";
  }
  return "Unknown";
}
typedef struct { int type; char *value; int line, col; } tok_t;
typedef struct { tok_t *t; int n, cap; } toks_t;

static void toks_push(toks_t *ts, int type, const char *v, size_t vlen, int line, int col) {
  if (ts->n == ts->cap) {
    ts->cap = ts->cap ? ts->cap * 2 : 32;
    ts->t = (tok_t *)realloc(ts->t, sizeof(tok_t) * ts->cap);
  }
  ts->t[ts->n].type = type;
  ts->t[ts->n].value = xstrndup(v, vlen);
  ts->t[ts->n].line = line;
  ts->t[ts->n].col = col;
  ts->n++;
}
static void toks_free(toks_t *ts) {
  for (int i = 0; i < ts->n; i++) free(ts->t[i].value);
  free(ts->t);
  ts->t = NULL;
  ts->n = ts->cap = 0;
}
static int is_keyword(const char *upper) { /* expression.cpp:58-62 */
  static const char *kw[] = {"SELECT", "FROM", "WHERE", "JOIN", "ON", "GROUP", "BY", "ORDER",
                             "ASC", "DESC", "LIMIT", "OFFSET", "SUM", "AVG", "COUNT", "MIN",
                             "MAX", "OVER", "PARTITION", "AND", "OR", "HAVING", "DISTINCT", NULL};
  for (int i = 0; kw[i]; i++)
    if (!strcmp(kw[i], upper)) return 1;
  return 0;
}
static int tokenize(const char *in, toks_t *out, char *err, size_t errlen) {
  size_t i = 0, n = strlen(in);
  int line = 1, col = 1;
  memset(out, 0, sizeof(*out));
#define ADV(c) do { if ((c) == '\n') { line++; col = 1; } else col++; } while (0)
  while (i < n) {
    unsigned char c = (unsigned char)in[i];
    if (isspace(c)) { ADV(c); i++; continue; }
    if (isalpha(c) || c == '_') {
      int sl = line, sc = col;
      size_t s = i;
      while (i < n && (isalnum((unsigned char)in[i]) || in[i] == '_' || in[i] == '.')) { ADV(in[i]); i++; }
      char *id = xstrndup(in + s, i - s), *up = xstrdup(id);
      for (char *p = up; *p; p++) *p = (char)toupper((unsigned char)*p);
      if (is_keyword(up)) toks_push(out, T_KEYWORD, up, strlen(up), sl, sc);
      else toks_push(out, T_IDENT, id, strlen(id), sl, sc);
      free(id); free(up);
    } else if (isdigit(c) || (c == '.' && i + 1 < n && isdigit((unsigned char)in[i + 1]))) {
      int sl = line, sc = col, has_dot = 0;
      size_t s = i;
      while (i < n && (isdigit((unsigned char)in[i]) || (!has_dot && in[i] == '.'))) {
        if (in[i] == '.') has_dot = 1;
        ADV(in[i]); i++;
      }
      toks_push(out, T_NUMBER, in + s, i - s, sl, sc);
    } else if (c == '>' || c == '<' || c == '=' || c == '!') {
      int sl = line, sc = col;
      size_t s = i;
      if (i + 1 < n && in[i + 1] == '=') { ADV(in[i]); i++; }
      ADV(in[i]); i++;
      toks_push(out, T_OP, in + s, i - s, sl, sc);
    } else if (strchr("+-*/()<>,.", c)) {
      int sl = line, sc = col;
      ADV(c); i++;
      toks_push(out, T_OP, in + i - 1, 1, sl, sc);
    } else {
      set_err(err, errlen, "Unknown character '%c' at line %d column %d", c, line, col);
      toks_free(out);
      return 1;
    }
  }
#undef ADV
  toks_push(out, T_END, "", 0, line, col);
  return 0;
}

int orc_tokenize_dump(const char *text, char *out, size_t outlen, char *err, size_t errlen) {
  toks_t ts;
  if (tokenize(text, &ts, err, errlen)) return 1;
  sbuf b = {out, 0, outlen, 1};
  char tmp[64];
  for (int i = 0; i < ts.n; i++) {
    sb_put(&b, tok_type_name(ts.t[i].type));
    sb_put(&b, ":");
    sb_put(&b, ts.t[i].value);
    snprintf(tmp, sizeof tmp, ":%d:%d\n", ts.t[i].line, ts.t[i].col);
    sb_put(&b, tmp);
  }
  sb_finish_fixed(&b);
  toks_free(&ts);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * AST + expression parser  (reference: src/expression.cpp:122-268, include/expression.hpp)
 * ---------------------------------------------------------------------------------------- */
enum { N_CONST, N_VAR, N_BIN, N_CALL, N_AGG, N_WINDOW };
enum { OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_GT, OP_LT, OP_GE, OP_LE, OP_EQ, OP_NE, OP_AND, OP_OR, OP_ASSIGN, OP_BAD };
enum { FN_DISCOUNT, FN_SQRTF, FN_FABSF, FN_FLOORF, FN_CEILF, FN_TRUNCF, FN_FMINF, FN_FMAXF, FN_UNKNOWN };

struct orc_node {
  int type;
  char *text;            /* constant text / variable name / operator / function name */
  int agg;               /* N_AGG, N_WINDOW */
  struct orc_node *l, *r; /* N_BIN; N_AGG/N_WINDOW use l */
  struct orc_node **args; /* N_CALL */
  int nargs;
  /* bind results (evaluation) */
  int op, fn, col, kind, fuse;
  float cf;
};

static orc_node *mk(int type, const char *text) {
  orc_node *n = (orc_node *)calloc(1, sizeof(orc_node));
  n->type = type;
  n->text = xstrdup(text);
  n->col = -1;
  return n;
}
void orc_free(orc_node *n) {
  if (!n) return;
  orc_free(n->l);
  orc_free(n->r);
  for (int i = 0; i < n->nargs; i++) orc_free(n->args[i]);
  free(n->args);
  free(n->text);
  free(n);
}

typedef struct {
  const tok_t *t; /* token window; must end with a T_END token */
  int n, cur;
  int ext_agg;    /* accept SUM(...) etc. as a factor (HAVING extension) */
  char *err; size_t errlen; int failed;
} parser;

static const tok_t *pk(parser *p) { return &p->t[p->cur < p->n ? p->cur : p->n - 1]; }
static int match_op(parser *p, const char *op) {
  const tok_t *t = pk(p);
  if (t->type == T_OP && !strcmp(t->value, op)) { p->cur++; return 1; }
  return 0;
}
static void pfail(parser *p, const char *fmt, ...) {
  if (p->failed) return;
  p->failed = 1;
  if (!p->err || !p->errlen) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(p->err, p->errlen, fmt, ap);
  va_end(ap);
}
static orc_node *p_add(parser *p);
static orc_node *p_or(parser *p);

static int agg_from_kw(const char *kw) {
  if (!strcmp(kw, "SUM")) return ORC_SUM;
  if (!strcmp(kw, "AVG")) return ORC_AVG;
  if (!strcmp(kw, "COUNT")) return ORC_COUNT;
  if (!strcmp(kw, "MIN")) return ORC_MIN;
  if (!strcmp(kw, "MAX")) return ORC_MAX;
  return -1;
}

static orc_node *p_factor(parser *p) { /* expression.cpp:205-235 */
  if (p->failed) return NULL;
  const tok_t *t = pk(p);
  if (t->type == T_NUMBER) { p->cur++; return mk(N_CONST, t->value); }
  if (t->type == T_IDENT) {
    p->cur++;
    if (match_op(p, "(")) {
      orc_node *c = mk(N_CALL, t->value);
      if (!match_op(p, ")")) {
        do {
          orc_node *a = p_add(p);
          if (p->failed) { orc_free(a); orc_free(c); return NULL; }
          c->args = (orc_node **)realloc(c->args, sizeof(orc_node *) * (c->nargs + 1));
          c->args[c->nargs++] = a;
        } while (match_op(p, ","));
        if (!match_op(p, ")")) { pfail(p, "Expected ')' after arguments"); orc_free(c); return NULL; }
      }
      return c;
    }
    return mk(N_VAR, t->value);
  }
  if (p->ext_agg && t->type == T_KEYWORD && agg_from_kw(t->value) >= 0) {
    /* extension (not in the reference grammar): AGG "(" add ")" inside HAVING */
    int a = agg_from_kw(t->value);
    p->cur++;
    if (!match_op(p, "(")) { pfail(p, "Invalid syntax for %s aggregation", t->value); return NULL; }
    orc_node *inner = p_add(p);
    if (p->failed) { orc_free(inner); return NULL; }
    if (!match_op(p, ")")) { pfail(p, "Expected ')'"); orc_free(inner); return NULL; }
    orc_node *n = mk(N_AGG, t->value);
    n->agg = a;
    n->l = inner;
    return n;
  }
  if (match_op(p, "(")) {
    orc_node *n = p_add(p);
    if (p->failed) { orc_free(n); return NULL; }
    if (!match_op(p, ")")) { pfail(p, "Expected ')'"); orc_free(n); return NULL; }
    return n;
  }
  pfail(p, "Unexpected token (%s: %s)", tok_type_name(t->type), t->value);
  return NULL;
}
static orc_node *mkbin(const char *op, orc_node *l, orc_node *r) {
  orc_node *n = mk(N_BIN, op);
  n->l = l;
  n->r = r;
  return n;
}
static orc_node *p_term(parser *p) { /* :193-202 */
  orc_node *n = p_factor(p);
  while (!p->failed) {
    const char *op = match_op(p, "*") ? "*" : match_op(p, "/") ? "/" : NULL;
    if (!op) break;
    orc_node *r = p_factor(p);
    n = mkbin(op, n, r);
  }
  return n;
}
static orc_node *p_add(parser *p) { /* :144-153 */
  orc_node *n = p_term(p);
  while (!p->failed) {
    const char *op = match_op(p, "+") ? "+" : match_op(p, "-") ? "-" : NULL;
    if (!op) break;
    orc_node *r = p_term(p);
    n = mkbin(op, n, r);
  }
  return n;
}
static orc_node *p_cmp(parser *p) { /* :156-166 */
  orc_node *n = p_add(p);
  static const char *ops[] = {">", "<", ">=", "<=", "==", "!=", "=", NULL};
  while (!p->failed) {
    const char *op = NULL;
    for (int i = 0; ops[i]; i++)
      if (match_op(p, ops[i])) { op = ops[i]; break; }
    if (!op) break;
    orc_node *r = p_add(p);
    n = mkbin(op, n, r);
  }
  return n;
}
static int is_kw(const tok_t *t, const char *kw) { return t->type == T_KEYWORD && !strcmp(t->value, kw); }
static orc_node *p_and(parser *p) { /* :169-178 */
  orc_node *n = p_cmp(p);
  while (!p->failed && is_kw(pk(p), "AND")) {
    p->cur++;
    orc_node *r = p_cmp(p);
    n = mkbin("&&", n, r);
  }
  return n;
}
static orc_node *p_or(parser *p) { /* :181-190 */
  orc_node *n = p_and(p);
  while (!p->failed && is_kw(pk(p), "OR")) {
    p->cur++;
    orc_node *r = p_and(p);
    n = mkbin("||", n, r);
  }
  return n;
}
/* parse_expression over a token window (a T_END is implied at t[n]) : :238-248 */
static orc_node *parse_tokens(const tok_t *t, int n, int ext_agg, char *err, size_t errlen) {
  tok_t *w = (tok_t *)malloc(sizeof(tok_t) * (n + 1));
  memcpy(w, t, sizeof(tok_t) * n);
  w[n].type = T_END; w[n].value = (char *)""; w[n].line = 0; w[n].col = 0;
  parser p = {w, n + 1, 0, ext_agg, err, errlen, 0};
  orc_node *node = p_or(&p);
  if (!p.failed && pk(&p)->type != T_END) pfail(&p, "Unexpected tokens remaining: %s", pk(&p)->value);
  if (p.failed) { orc_free(node); node = NULL; }
  free(w);
  return node;
}
orc_node *orc_parse_expression(const char *text, char *err, size_t errlen) {
  toks_t ts;
  if (tokenize(text, &ts, err, errlen)) return NULL;
  orc_node *n = parse_tokens(ts.t, ts.n - 1, 0, err, errlen); /* drop the T_END; window re-adds it */
  toks_free(&ts);
  return n;
}

static void cuda_expr(const orc_node *n, sbuf *b) { /* include/expression.hpp:32-38,45,56-59,69-78,93,119 */
  switch (n->type) {
  case N_CONST:
    sb_put(b, n->text);
    sb_put(b, strchr(n->text, '.') ? "f" : ".0f");
    break;
  case N_VAR: sb_put(b, n->text); sb_put(b, "[idx]"); break;
  case N_BIN:
    sb_put(b, "("); cuda_expr(n->l, b); sb_put(b, " "); sb_put(b, n->text); sb_put(b, " ");
    cuda_expr(n->r, b); sb_put(b, ")");
    break;
  case N_CALL:
    sb_put(b, n->text); sb_put(b, "(");
    for (int i = 0; i < n->nargs; i++) { if (i) sb_put(b, ", "); cuda_expr(n->args[i], b); }
    sb_put(b, ")");
    break;
  case N_AGG: cuda_expr(n->l, b); break;
  case N_WINDOW: sb_put(b, "<window>"); break;
  }
}
size_t orc_to_cuda_expr(const orc_node *n, char *out, size_t outlen) {
  sbuf b = {out, 0, outlen, 1};
  cuda_expr(n, &b);
  return sb_finish_fixed(&b);
}

/* ------------------------------------------------------------------------------------------
 * parse_query  (reference: src/expression.cpp:270-531, with the `}` the reference lacks after
 * :522 -- SURVEY F1/F2).  The clause-scanning quirks are kept: see SURVEY Appendix A.
 * ---------------------------------------------------------------------------------------- */
void orc_free_query(orc_query_t *q) {
  if (!q) return;
  for (int i = 0; i < q->n_select; i++) orc_free(q->select_list[i]);
  free(q->select_list);
  free(q->from_table);
  for (int i = 0; i < q->n_joins; i++) { free(q->join_tables[i]); orc_free(q->join_conds[i]); }
  free(q->join_tables); free(q->join_conds);
  orc_free(q->where);
  for (int i = 0; i < q->n_group; i++) orc_free(q->group_keys[i]);
  free(q->group_keys);
  orc_free(q->having);
  orc_free(q->order_expr);
  free(q);
}
static int is_opv(const tok_t *t, const char *v) { return t->type == T_OP && !strcmp(t->value, v); }

static orc_node *parse_select_item(const tok_t *it, int n, char *err, size_t errlen) { /* :296-337 */
  if (n > 0 && it[0].type == T_KEYWORD && agg_from_kw(it[0].value) >= 0) {
    int over = n;
    for (int i = 0; i < n; i++)
      if (is_kw(&it[i], "OVER")) { over = i; break; }
    int has_paren = over > 1 && is_opv(&it[1], "(") && is_opv(&it[over - 1], ")");
    if (!has_paren) { set_err(err, errlen, "Invalid syntax for %s aggregation", it[0].value); return NULL; }
    orc_node *inner = parse_tokens(it + 2, over - 3 < 0 ? 0 : over - 3, 0, err, errlen);
    if (!inner) return NULL;
    orc_node *a = mk(over < n ? N_WINDOW : N_AGG, it[0].value);
    a->agg = agg_from_kw(it[0].value);
    a->l = inner;
    return a;
  }
  return parse_tokens(it, n, 0, err, errlen);
}

orc_query_t *orc_parse_query(const char *sql, int ext, char *err, size_t errlen) {
  toks_t ts;
  if (tokenize(sql, &ts, err, errlen)) return NULL;
  const tok_t *t = ts.t;
  int size = ts.n, end = size, pos = 0;
  if (end > 0 && t[end - 1].type == T_END) --end;
  orc_query_t *q = (orc_query_t *)calloc(1, sizeof(*q));
  q->order_asc = 1;
#define FAIL(...) do { set_err(err, errlen, __VA_ARGS__); goto fail; } while (0)
#define LC_L (pos < size ? t[pos].line : t[size - 1].line)
#define LC_C (pos < size ? t[pos].col : t[size - 1].col)
#define EXPECT_KW(kw) do { if (pos >= size || !is_kw(&t[pos], kw)) { \
    FAIL("Expected keyword '%s' at line %d column %d", kw, LC_L, LC_C); } \
    pos++; } while (0)
  EXPECT_KW("SELECT");
  if (pos < size && is_kw(&t[pos], "DISTINCT")) { q->distinct = 1; pos++; }
  while (pos < end) { /* :339-361 */
    if (is_kw(&t[pos], "FROM")) break;
    int start = pos, depth = 0;
    while (pos < end) {
      if (is_opv(&t[pos], "(")) depth++;
      if (is_opv(&t[pos], ")")) depth--;
      if (depth == 0 && (is_opv(&t[pos], ",") || is_kw(&t[pos], "FROM"))) break;
      pos++;
    }
    orc_node *item = parse_select_item(t + start, pos - start, err, errlen);
    if (!item) goto fail;
    q->select_list = (orc_node **)realloc(q->select_list, sizeof(orc_node *) * (q->n_select + 1));
    q->select_list[q->n_select++] = item;
    if (pos < end && is_opv(&t[pos], ",")) pos++;
  }
  EXPECT_KW("FROM");
  if (pos >= size || t[pos].type != T_IDENT)
    FAIL("Expected table name after FROM at line %d column %d", LC_L, LC_C);
  q->from_table = xstrdup(t[pos++].value);
  while (pos < size && is_kw(&t[pos], "JOIN")) { /* :375-401 */
    pos++;
    if (pos >= size || t[pos].type != T_IDENT)
      FAIL("Expected table name after JOIN at line %d column %d", LC_L, LC_C);
    char *jt = xstrdup(t[pos++].value);
    if (pos >= size || !is_kw(&t[pos], "ON")) {
      free(jt);
      FAIL("Expected keyword 'ON' at line %d column %d", LC_L, LC_C);
    }
    pos++;
    int start = pos;
    while (pos < end && !(t[pos].type == T_KEYWORD &&
                          (!strcmp(t[pos].value, "WHERE") || !strcmp(t[pos].value, "GROUP") ||
                           !strcmp(t[pos].value, "ORDER") || !strcmp(t[pos].value, "HAVING") ||
                           !strcmp(t[pos].value, "JOIN") || !strcmp(t[pos].value, "LIMIT"))))
      pos++;
    orc_node *c = parse_tokens(t + start, pos - start, 0, err, errlen);
    if (!c) { free(jt); goto fail; }
    q->join_tables = (char **)realloc(q->join_tables, sizeof(char *) * (q->n_joins + 1));
    q->join_conds = (orc_node **)realloc(q->join_conds, sizeof(orc_node *) * (q->n_joins + 1));
    q->join_tables[q->n_joins] = jt;
    q->join_conds[q->n_joins++] = c;
  }
  if (pos < end && is_kw(&t[pos], "WHERE")) { /* :403-415 */
    pos++;
    int start = pos;
    while (pos < end && !(t[pos].type == T_KEYWORD &&
                          (!strcmp(t[pos].value, "GROUP") || !strcmp(t[pos].value, "ORDER") ||
                           !strcmp(t[pos].value, "HAVING") || !strcmp(t[pos].value, "LIMIT"))))
      pos++;
    q->where = parse_tokens(t + start, pos - start, 0, err, errlen);
    if (!q->where) goto fail;
  }
  if (pos < end && is_kw(&t[pos], "GROUP")) { /* :417-443 */
    pos++;
    EXPECT_KW("BY");
    q->has_group = 1;
    while (pos < end) {
      int start = pos;
      while (pos < end && !is_opv(&t[pos], ",") && !(is_kw(&t[pos], "ORDER") || is_kw(&t[pos], "HAVING"))) pos++;
      orc_node *k = parse_tokens(t + start, pos - start, 0, err, errlen);
      if (!k) goto fail;
      q->group_keys = (orc_node **)realloc(q->group_keys, sizeof(orc_node *) * (q->n_group + 1));
      q->group_keys[q->n_group++] = k;
      if (pos < end && is_opv(&t[pos], ",")) pos++;
      if (pos < size && (is_kw(&t[pos], "ORDER") || is_kw(&t[pos], "HAVING"))) break;
    }
  }
  for (int blk = 0; blk < 2; blk++) { /* the two HAVING blocks :446-472 */
    if (pos < size && is_kw(&t[pos], "HAVING")) {
      pos++;
      int start = pos;
      while (pos < size && !(t[pos].type == T_KEYWORD &&
                             (!strcmp(t[pos].value, "ORDER") || !strcmp(t[pos].value, "LIMIT") ||
                              (blk == 1 && !strcmp(t[pos].value, "OFFSET")))))
        pos++;
      int stop = pos > end ? end : pos; /* the reference's window may swallow the End token */
      orc_free(q->having);
      q->having = parse_tokens(t + start, stop - start, ext, err, errlen);
      if (!q->having) goto fail;
    }
  }
  if (pos < size && is_kw(&t[pos], "ORDER")) { /* :474-495 */
    pos++;
    EXPECT_KW("BY");
    int start = pos;
    while (pos < end && !(is_kw(&t[pos], "ASC") || is_kw(&t[pos], "DESC"))) pos++;
    q->order_expr = parse_tokens(t + start, pos - start, 0, err, errlen);
    if (!q->order_expr) goto fail;
    q->has_order = 1;
    q->order_asc = 1;
    if (pos < end && (is_kw(&t[pos], "ASC") || is_kw(&t[pos], "DESC"))) {
      q->order_asc = !strcmp(t[pos].value, "ASC");
      pos++;
    }
  }
  if (pos < end && is_kw(&t[pos], "LIMIT")) { /* :497-512 */
    pos++;
    if (pos >= size || t[pos].type != T_NUMBER)
      FAIL("Expected numeric value after LIMIT at line %d column %d", LC_L, LC_C);
    q->has_limit = 1;
    q->limit = atoi(t[pos].value);
    pos++;
  }
  if (pos < size && is_kw(&t[pos], "OFFSET")) { /* :515-523 */
    pos++;
    if (pos >= size || t[pos].type != T_NUMBER) FAIL("Expected numeric value after OFFSET");
    q->has_offset = 1;
    q->offset = atoi(t[pos].value);
    pos++;
  }
  if (pos != end) /* :524-528 (pos==size reads past the vector in the reference: UB; we print "") */
    FAIL("Unexpected token in query near: %s", pos < size ? t[pos].value : "");
  toks_free(&ts);
  return q;
fail:
  toks_free(&ts);
  orc_free_query(q);
  return NULL;
#undef FAIL
#undef LC_L
#undef LC_C
#undef EXPECT_KW
}

static void summary_item(const orc_node *n, sbuf *b) {
  char tmp[16];
  if (n->type == N_AGG || n->type == N_WINDOW) {
    snprintf(tmp, sizeof tmp, "%s%d(", n->type == N_AGG ? "AGG" : "WIN", n->agg);
    sb_put(b, tmp);
    cuda_expr(n->l, b);
    sb_put(b, ")");
  } else
    cuda_expr(n, b);
}
/* select=[a;b] from=t joins=[t2:(c)] where=.. group=[..] having=.. order=..:ASC limit=n offset=n distinct=d */
size_t orc_query_summary(const orc_query_t *q, char *out, size_t outlen) {
  sbuf b = {out, 0, outlen, 1};
  char tmp[64];
  sb_put(&b, "select=[");
  for (int i = 0; i < q->n_select; i++) { if (i) sb_put(&b, ";"); summary_item(q->select_list[i], &b); }
  sb_put(&b, "] from="); sb_put(&b, q->from_table ? q->from_table : "");
  sb_put(&b, " joins=[");
  for (int i = 0; i < q->n_joins; i++) {
    if (i) sb_put(&b, ";");
    sb_put(&b, q->join_tables[i]); sb_put(&b, ":"); summary_item(q->join_conds[i], &b);
  }
  sb_put(&b, "] where=");
  if (q->where) summary_item(q->where, &b); else sb_put(&b, "-");
  sb_put(&b, " group=");
  if (q->has_group) {
    sb_put(&b, "[");
    for (int i = 0; i < q->n_group; i++) { if (i) sb_put(&b, ";"); summary_item(q->group_keys[i], &b); }
    sb_put(&b, "]");
  } else sb_put(&b, "-");
  sb_put(&b, " having=");
  if (q->having) summary_item(q->having, &b); else sb_put(&b, "-");
  sb_put(&b, " order=");
  if (q->has_order) { summary_item(q->order_expr, &b); sb_put(&b, q->order_asc ? ":ASC" : ":DESC"); } else sb_put(&b, "-");
  if (q->has_limit) snprintf(tmp, sizeof tmp, " limit=%d", q->limit); else snprintf(tmp, sizeof tmp, " limit=-");
  sb_put(&b, tmp);
  if (q->has_offset) snprintf(tmp, sizeof tmp, " offset=%d", q->offset); else snprintf(tmp, sizeof tmp, " offset=-");
  sb_put(&b, tmp);
  snprintf(tmp, sizeof tmp, " distinct=%d", q->distinct);
  sb_put(&b, tmp);
  return sb_finish_fixed(&b);
}

/* ------------------------------------------------------------------------------------------
 * validation (src/warpdb.cpp:19-44)
 * ---------------------------------------------------------------------------------------- */
static int find_col(const orc_col_t *cols, int ncols, const char *name) {
  for (int i = 0; i < ncols; i++)
    if (!strcmp(cols[i].name, name)) return i;
  return -1;
}
int orc_validate(const orc_node *n, const orc_col_t *cols, int ncols, char *err, size_t errlen) {
  if (!n) return 0;
  if (n->type == N_VAR && find_col(cols, ncols, n->text) < 0) {
    set_err(err, errlen, "Unknown column: %s", n->text);
    return 1;
  }
  if (orc_validate(n->l, cols, ncols, err, errlen)) return 1;
  if (orc_validate(n->r, cols, ncols, err, errlen)) return 1;
  for (int i = 0; i < n->nargs; i++)
    if (orc_validate(n->args[i], cols, ncols, err, errlen)) return 1;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * typed evaluation.  The reference's GPU path pastes the expression into CUDA C where
 * `col[idx]` has the column's C type (src/jit.cpp:31-45,75-79) and every literal is a float
 * (include/expression.hpp:32-38): C++ usual arithmetic conversions apply.  Kinds are ordered
 * like DataType so the common type of a binary op is max(kind).
 * ---------------------------------------------------------------------------------------- */
typedef struct { int k; union { int32_t i; int64_t l; float f; double d; } u; } val_t;

static int op_from_text(const char *s) {
  static const char *names[] = {"+", "-", "*", "/", ">", "<", ">=", "<=", "==", "!=", "&&", "||", "="};
  for (int i = 0; i < 13; i++)
    if (!strcmp(s, names[i])) return i;
  return OP_BAD;
}
static int fn_from_text(const char *s) {
  static const char *names[] = {"discount", "sqrtf", "fabsf", "floorf", "ceilf", "truncf", "fminf", "fmaxf"};
  for (int i = 0; i < 8; i++)
    if (!strcmp(s, names[i])) return i;
  return FN_UNKNOWN;
}
static int is_float_mul(const orc_node *n, int kind) {
  if (n->type == N_BIN && n->op == OP_MUL && n->kind == kind) return 1;
  if (n->type == N_CALL && n->fn == FN_DISCOUNT && kind == ORC_FLOAT32) return 1; /* custom.cu:1-3 inlines to a*b */
  return 0;
}
/* resolves columns/ops and computes the static C++ type of every node */
static int bind(orc_node *n, const orc_col_t *cols, int ncols, int contract, char *err, size_t errlen) {
  if (!n) return 0;
  switch (n->type) {
  case N_CONST:
    n->cf = strtof(n->text, NULL); /* literal with f suffix == std::stof (warpdb.cpp:130) */
    n->kind = ORC_FLOAT32;
    return 0;
  case N_VAR:
    n->col = find_col(cols, ncols, n->text);
    if (n->col < 0) { set_err(err, errlen, "Unknown column: %s", n->text); return 1; }
    if (cols[n->col].dtype > ORC_FLOAT64) { set_err(err, errlen, "oracle: column %s has no numeric type", n->text); return 1; }
    n->kind = cols[n->col].dtype;
    return 0;
  case N_BIN: {
    if (bind(n->l, cols, ncols, contract, err, errlen) || bind(n->r, cols, ncols, contract, err, errlen)) return 1;
    n->op = op_from_text(n->text);
    if (n->op == OP_BAD) { set_err(err, errlen, "oracle: unsupported operator %s", n->text); return 1; }
    int common = n->l->kind > n->r->kind ? n->l->kind : n->r->kind;
    if (n->op <= OP_DIV) n->kind = common;
    else if (n->op == OP_ASSIGN) {
      if (n->l->type != N_VAR) { set_err(err, errlen, "oracle: assignment to a non-column"); return 1; }
      n->kind = n->l->kind;
    } else n->kind = ORC_INT32; /* bool promotes to int */
    n->fuse = 0;
    if (contract && (n->op == OP_ADD || n->op == OP_SUB) && n->kind >= ORC_FLOAT32) {
      if (is_float_mul(n->l, n->kind)) n->fuse = 1;       /* (a*b) +- c -> fma(a,b,+-c) */
      else if (is_float_mul(n->r, n->kind)) n->fuse = 2;  /* c +- (a*b) -> fma(+-a,b,c) */
    }
    return 0;
  }
  case N_CALL:
    for (int i = 0; i < n->nargs; i++)
      if (bind(n->args[i], cols, ncols, contract, err, errlen)) return 1;
    n->fn = fn_from_text(n->text);
    if (n->fn == FN_UNKNOWN) { set_err(err, errlen, "oracle: unknown function %s", n->text); return 1; }
    {
      int want = (n->fn == FN_DISCOUNT || n->fn == FN_FMINF || n->fn == FN_FMAXF) ? 2 : 1;
      if (n->nargs != want) { set_err(err, errlen, "oracle: %s takes %d arguments", n->text, want); return 1; }
    }
    n->kind = ORC_FLOAT32;
    return 0;
  case N_AGG:
    if (bind(n->l, cols, ncols, contract, err, errlen)) return 1;
    n->kind = n->l->kind;
    return 0;
  default:
    set_err(err, errlen, "oracle: window functions are not evaluated");
    return 1;
  }
}

static inline double v_f64(val_t v) { return v.k == ORC_INT32 ? (double)v.u.i : v.k == ORC_INT64 ? (double)v.u.l : v.k == ORC_FLOAT32 ? (double)v.u.f : v.u.d; }
static inline float v_f32(val_t v) { return v.k == ORC_INT32 ? (float)v.u.i : v.k == ORC_INT64 ? (float)v.u.l : v.k == ORC_FLOAT32 ? v.u.f : (float)v.u.d; }
static inline int64_t v_i64(val_t v) { return v.k == ORC_INT32 ? (int64_t)v.u.i : v.u.l; } /* ints only */
static inline int v_true(val_t v) { return v.k == ORC_INT32 ? v.u.i != 0 : v.k == ORC_INT64 ? v.u.l != 0 : v.k == ORC_FLOAT32 ? v.u.f != 0.0f : v.u.d != 0.0; }
static inline val_t mk_i(int32_t i) { val_t v; v.k = ORC_INT32; v.u.i = i; return v; }
static inline val_t mk_f(float f) { val_t v; v.k = ORC_FLOAT32; v.u.f = f; return v; }
static inline val_t mk_d(double d) { val_t v; v.k = ORC_FLOAT64; v.u.d = d; return v; }
static inline val_t mk_l(int64_t l) { val_t v; v.k = ORC_INT64; v.u.l = l; return v; }
/* CUDA float->int conversions saturate and map NaN to 0 (cvt.rzi.s32.f32 / .f64) */
static inline int32_t sat_i32_from_f64(double d) {
  if (d != d) return 0;
  if (d >= 2147483647.0) return INT32_MAX;
  if (d <= -2147483648.0) return INT32_MIN;
  return (int32_t)d;
}
static inline int32_t v_i32_cast(val_t v) {
  switch (v.k) {
  case ORC_INT32: return v.u.i;
  case ORC_INT64: return (int32_t)(uint32_t)(uint64_t)v.u.l;
  case ORC_FLOAT32: return sat_i32_from_f64((double)v.u.f);
  default: return sat_i32_from_f64(v.u.d);
  }
}
static inline val_t v_cast(val_t v, int kind) {
  switch (kind) {
  case ORC_INT32: return mk_i(v_i32_cast(v));
  case ORC_INT64:
    if (v.k <= ORC_INT64) return mk_l(v_i64(v));
    { double d = v_f64(v); if (d != d) return mk_l(0); if (d >= 9223372036854775807.0) return mk_l(INT64_MAX); if (d <= -9223372036854775808.0) return mk_l(INT64_MIN); return mk_l((int64_t)d); }
  case ORC_FLOAT32: return mk_f(v_f32(v));
  default: return mk_d(v_f64(v));
  }
}

typedef struct { const orc_col_t *cols; int64_t row; } ectx;

static val_t eval(const orc_node *n, const ectx *c);

static inline val_t load_col(const orc_col_t *col, int64_t row) {
  switch (col->dtype) {
  case ORC_INT32: return mk_i(((const int32_t *)col->data)[row]);
  case ORC_INT64: return mk_l(((const int64_t *)col->data)[row]);
  case ORC_FLOAT32: return mk_f(((const float *)col->data)[row]);
  default: return mk_d(((const double *)col->data)[row]);
  }
}
/* operands of a float multiply (a BinaryOp '*' or the discount() UDF), converted to `kind` */
static void mul_operands(const orc_node *m, const ectx *c, int kind, val_t *a, val_t *b) {
  const orc_node *l = m->type == N_CALL ? m->args[0] : m->l;
  const orc_node *r = m->type == N_CALL ? m->args[1] : m->r;
  *a = v_cast(eval(l, c), kind);
  *b = v_cast(eval(r, c), kind);
}
static val_t eval_bin(const orc_node *n, const ectx *c) {
  int op = n->op;
  if (op == OP_AND) return mk_i(v_true(eval(n->l, c)) && v_true(eval(n->r, c)));
  if (op == OP_OR) return mk_i(v_true(eval(n->l, c)) || v_true(eval(n->r, c)));
  if (op == OP_ASSIGN) return v_cast(eval(n->r, c), n->kind);
  if (n->fuse) { /* NVRTC default --fmad=true contracts a float multiply feeding an add/sub */
    val_t a, b, z;
    if (n->fuse == 1) {
      mul_operands(n->l, c, n->kind, &a, &b);
      z = v_cast(eval(n->r, c), n->kind);
      if (n->kind == ORC_FLOAT32) return mk_f(fmaf(a.u.f, b.u.f, op == OP_ADD ? z.u.f : -z.u.f));
      return mk_d(fma(a.u.d, b.u.d, op == OP_ADD ? z.u.d : -z.u.d));
    }
    z = v_cast(eval(n->l, c), n->kind);
    mul_operands(n->r, c, n->kind, &a, &b);
    if (n->kind == ORC_FLOAT32) return mk_f(fmaf(op == OP_ADD ? a.u.f : -a.u.f, b.u.f, z.u.f));
    return mk_d(fma(op == OP_ADD ? a.u.d : -a.u.d, b.u.d, z.u.d));
  }
  val_t l = eval(n->l, c), r = eval(n->r, c);
  int k = l.k > r.k ? l.k : r.k;
  if (k == ORC_FLOAT32) {
    float a = v_f32(l), b = v_f32(r);
    switch (op) {
    case OP_ADD: return mk_f(a + b);
    case OP_SUB: return mk_f(a - b);
    case OP_MUL: return mk_f(a * b);
    case OP_DIV: return mk_f(a / b);
    case OP_GT: return mk_i(a > b);
    case OP_LT: return mk_i(a < b);
    case OP_GE: return mk_i(a >= b);
    case OP_LE: return mk_i(a <= b);
    case OP_EQ: return mk_i(a == b);
    default: return mk_i(a != b);
    }
  }
  if (k == ORC_FLOAT64) {
    double a = v_f64(l), b = v_f64(r);
    switch (op) {
    case OP_ADD: return mk_d(a + b);
    case OP_SUB: return mk_d(a - b);
    case OP_MUL: return mk_d(a * b);
    case OP_DIV: return mk_d(a / b);
    case OP_GT: return mk_i(a > b);
    case OP_LT: return mk_i(a < b);
    case OP_GE: return mk_i(a >= b);
    case OP_LE: return mk_i(a <= b);
    case OP_EQ: return mk_i(a == b);
    default: return mk_i(a != b);
    }
  }
  {
    int64_t a = v_i64(l), b = v_i64(r), res = 0;
    switch (op) {
    case OP_ADD: res = (int64_t)((uint64_t)a + (uint64_t)b); break;
    case OP_SUB: res = (int64_t)((uint64_t)a - (uint64_t)b); break;
    case OP_MUL: res = (int64_t)((uint64_t)a * (uint64_t)b); break;
    case OP_DIV: res = (b == 0) ? -1 : (b == -1 ? (int64_t)(0 - (uint64_t)a) : a / b); break; /* UB in C++; never tested */
    case OP_GT: return mk_i(a > b);
    case OP_LT: return mk_i(a < b);
    case OP_GE: return mk_i(a >= b);
    case OP_LE: return mk_i(a <= b);
    case OP_EQ: return mk_i(a == b);
    default: return mk_i(a != b);
    }
    if (k == ORC_INT32) return mk_i((int32_t)(uint32_t)(uint64_t)res);
    return mk_l(res);
  }
}
static val_t eval(const orc_node *n, const ectx *c) {
  switch (n->type) {
  case N_CONST: return mk_f(n->cf);
  case N_VAR: return load_col(&c->cols[n->col], c->row);
  case N_BIN: return eval_bin(n, c);
  case N_AGG: return eval(n->l, c);
  case N_CALL: {
    float a = v_f32(eval(n->args[0], c));
    float b = n->nargs > 1 ? v_f32(eval(n->args[1], c)) : 0.0f;
    switch (n->fn) {
    case FN_DISCOUNT: return mk_f(a * b); /* custom.cu:1-3 */
    case FN_SQRTF: return mk_f(sqrtf(a));
    case FN_FABSF: return mk_f(fabsf(a));
    case FN_FLOORF: return mk_f(floorf(a));
    case FN_CEILF: return mk_f(ceilf(a));
    case FN_TRUNCF: return mk_f(truncf(a));
    case FN_FMINF: return mk_f(fminf(a, b));
    default: return mk_f(fmaxf(a, b));
    }
  }
  }
  return mk_f(0.0f);
}

/* ------------------------------------------------------------------------------------------
 * operators
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const orc_node *expr, *cond;
  const orc_col_t *cols;
  int64_t r0, r1, base;
  float *out;
  uint8_t *mask;
} pf_job;
static void *pf_run(void *arg) {
  pf_job *j = (pf_job *)arg;
  ectx c = {j->cols, 0};
  for (int64_t i = j->r0; i < j->r1; i++) {
    c.row = i;
    int keep = j->cond ? v_true(eval(j->cond, &c)) : 1;
    if (j->mask) j->mask[i - j->base] = (uint8_t)keep;
    if (keep) j->out[i - j->base] = v_f32(eval(j->expr, &c));
  }
  return NULL;
}
static orc_node *ast_clone(const orc_node *n) {
  if (!n) return NULL;
  orc_node *c = mk(n->type, n->text);
  c->agg = n->agg;
  c->l = ast_clone(n->l);
  c->r = ast_clone(n->r);
  c->nargs = n->nargs;
  if (n->nargs) {
    c->args = (orc_node **)malloc(sizeof(orc_node *) * n->nargs);
    for (int i = 0; i < n->nargs; i++) c->args[i] = ast_clone(n->args[i]);
  }
  return c;
}
/* bound private copies so the caller's AST stays const and re-bindable */
static int bind_pair(const orc_node *a, const orc_node *b, orc_node **oa, orc_node **ob, const orc_col_t *cols,
                     int ncols, int contract, char *err, size_t errlen) {
  *oa = ast_clone(a);
  *ob = ast_clone(b);
  if ((*oa && bind(*oa, cols, ncols, contract, err, errlen)) || (*ob && bind(*ob, cols, ncols, contract, err, errlen))) {
    orc_free(*oa); orc_free(*ob);
    *oa = *ob = NULL;
    return 1;
  }
  return 0;
}

int orc_project_filter(const orc_node *expr, const orc_node *cond, const orc_col_t *cols, int ncols,
                       int64_t row0, int64_t row1, float *out, uint8_t *mask, int contract,
                       int nthreads, char *err, size_t errlen) {
  orc_node *e, *c;
  if (bind_pair(expr, cond, &e, &c, cols, ncols, contract, err, errlen)) return 1;
  if (nthreads < 1) nthreads = 1;
  int64_t n = row1 - row0;
  if (n < 4096) nthreads = 1;
  pf_job *jobs = (pf_job *)calloc(nthreads, sizeof(pf_job));
  pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
  int64_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    int64_t a = row0 + t * chunk, b = a + chunk > row1 ? row1 : a + chunk;
    if (a > row1) a = row1;
    pf_job j = {e, c, cols, a, b, row0, out, mask};
    jobs[t] = j;
    if (nthreads == 1) pf_run(&jobs[t]);
    else pthread_create(&th[t], NULL, pf_run, &jobs[t]);
  }
  if (nthreads > 1)
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(jobs); free(th);
  orc_free(e); orc_free(c);
  return 0;
}

int orc_filter_compact(const orc_node *expr, const orc_node *cond, const orc_col_t *cols, int ncols,
                       int64_t row0, int64_t row1, float *out, int64_t *out_count, int contract,
                       char *err, size_t errlen) {
  orc_node *e, *c;
  if (bind_pair(expr, cond, &e, &c, cols, ncols, contract, err, errlen)) return 1;
  ectx x = {cols, 0};
  int64_t k = 0;
  for (int64_t i = row0; i < row1; i++) {
    x.row = i;
    if (c && !v_true(eval(c, &x))) continue;
    out[k++] = v_f32(eval(e, &x));
  }
  *out_count = k;
  orc_free(e); orc_free(c);
  return 0;
}

/* open-addressing int32 -> AggData map (warpdb.cpp:379-384 uses std::map<int,AggData>) */
typedef struct { int32_t key; int used; int64_t first; double sum, count, mn, mx; } agg_slot;
typedef struct { agg_slot *s; int64_t cap, n; } agg_map;
static uint64_t hash32(int32_t k) { uint64_t x = (uint32_t)k; x *= 0x9E3779B97F4A7C15ull; return x ^ (x >> 29); }
static agg_slot *agg_find(agg_map *m, int32_t key) {
  if ((m->n + 1) * 2 > m->cap) {
    int64_t ncap = m->cap ? m->cap * 2 : 1024;
    agg_slot *ns = (agg_slot *)calloc(ncap, sizeof(agg_slot));
    for (int64_t i = 0; i < m->cap; i++)
      if (m->s[i].used) {
        uint64_t h = hash32(m->s[i].key) & (ncap - 1);
        while (ns[h].used) h = (h + 1) & (ncap - 1);
        ns[h] = m->s[i];
      }
    free(m->s);
    m->s = ns;
    m->cap = ncap;
  }
  uint64_t h = hash32(key) & (m->cap - 1);
  while (m->s[h].used && m->s[h].key != key) h = (h + 1) & (m->cap - 1);
  if (!m->s[h].used) {
    m->s[h].used = 1;
    m->s[h].key = key;
    m->s[h].first = m->n++;
  }
  return &m->s[h];
}
static int cmp_slot_first(const void *a, const void *b) {
  const agg_slot *x = (const agg_slot *)a, *y = (const agg_slot *)b;
  return x->first < y->first ? -1 : x->first > y->first;
}
static int cmp_slot_asc(const void *a, const void *b) {
  const agg_slot *x = (const agg_slot *)a, *y = (const agg_slot *)b;
  return x->key < y->key ? -1 : x->key > y->key;
}
static int cmp_slot_desc(const void *a, const void *b) { return cmp_slot_asc(b, a); }

static float agg_result(const agg_slot *g, int agg) { /* warpdb.cpp:429-435 */
  switch (agg) {
  case ORC_SUM: return (float)g->sum;
  case ORC_AVG: return (float)(g->sum / g->count);
  case ORC_COUNT: return (float)g->count;
  case ORC_MIN: return (float)g->mn;
  default: return (float)g->mx;
  }
}
/* builds the group table; caller frees *out_slots */
static int group_build(const orc_node *val, const orc_node *key, const orc_node *cond, int agg,
                       const orc_col_t *cols, int ncols, int64_t row0, int64_t row1, int contract,
                       agg_slot **out_slots, int64_t *out_n, char *err, size_t errlen) {
  orc_node *v, *k, *c, *dummy;
  if (bind_pair(val, key, &v, &k, cols, ncols, contract, err, errlen)) return 1;
  if (bind_pair(cond, NULL, &c, &dummy, cols, ncols, contract, err, errlen)) { orc_free(v); orc_free(k); return 1; }
  agg_map m = {0, 0, 0};
  ectx x = {cols, 0};
  for (int64_t i = row0; i < row1; i++) {
    x.row = i;
    if (c && !v_true(eval(c, &x))) continue;
    int32_t kk = v_i32_cast(eval(k, &x));             /* static_cast<int>(...) warpdb.cpp:374 / jit.cpp:200 */
    float fv = 0.0f;
    if (agg != ORC_COUNT) fv = v_f32(eval(v, &x));     /* warpdb.cpp:375-378 */
    agg_slot *g = agg_find(&m, kk);
    if (g->count == 0.0) { g->mn = g->mx = (double)fv; }
    g->sum += (double)fv;
    g->count += 1.0;
    if ((double)fv < g->mn) g->mn = (double)fv;
    if ((double)fv > g->mx) g->mx = (double)fv;
  }
  agg_slot *dense = (agg_slot *)malloc(sizeof(agg_slot) * (m.n ? m.n : 1));
  int64_t j = 0;
  for (int64_t i = 0; i < m.cap; i++)
    if (m.s[i].used) dense[j++] = m.s[i];
  free(m.s);
  *out_slots = dense;
  *out_n = j;
  orc_free(v); orc_free(k); orc_free(c);
  return 0;
}
int orc_group_agg(const orc_node *val, const orc_node *key, const orc_node *cond, int agg, int order,
                  const orc_col_t *cols, int ncols, int64_t row0, int64_t row1, int32_t *out_keys,
                  float *out_vals, double *out_sums, int64_t *out_counts, int64_t cap,
                  int64_t *out_groups, int contract, char *err, size_t errlen) {
  agg_slot *g;
  int64_t n;
  if (group_build(val, key, cond, agg, cols, ncols, row0, row1, contract, &g, &n, err, errlen)) return 1;
  qsort(g, n, sizeof(agg_slot), order == ORC_ORDER_FIRST ? cmp_slot_first : order == ORC_ORDER_KEY_ASC ? cmp_slot_asc : cmp_slot_desc);
  if (n > cap) { free(g); set_err(err, errlen, "oracle: %lld groups exceed capacity %lld", (long long)n, (long long)cap); return 1; }
  for (int64_t i = 0; i < n; i++) {
    if (out_keys) out_keys[i] = g[i].key;
    if (out_vals) out_vals[i] = agg_result(&g[i], agg);
    if (out_sums) out_sums[i] = g[i].sum;
    if (out_counts) out_counts[i] = (int64_t)g[i].count;
  }
  *out_groups = n;
  free(g);
  return 0;
}

/* stable merge sort on (key,value) float pairs: equal to the reference's bubble sorts */
typedef struct { float k, v; } kv_t;
static void kv_msort(kv_t *a, kv_t *tmp, int64_t n, int asc) {
  if (n < 2) return;
  int64_t h = n / 2;
  kv_msort(a, tmp, h, asc);
  kv_msort(a + h, tmp, n - h, asc);
  int64_t i = 0, j = h, o = 0;
  while (i < h && j < n) {
    /* take right only when strictly out of order (bubble sort swaps on > / <) */
    int take_right = asc ? (a[i].k > a[j].k) : (a[i].k < a[j].k);
    tmp[o++] = take_right ? a[j++] : a[i++];
  }
  while (i < h) tmp[o++] = a[i++];
  while (j < n) tmp[o++] = a[j++];
  memcpy(a, tmp, sizeof(kv_t) * n);
}
void orc_sort_float(float *vals, int64_t n, int ascending) {
  if (n < 2) return;
  kv_t *a = (kv_t *)malloc(sizeof(kv_t) * n), *t = (kv_t *)malloc(sizeof(kv_t) * n);
  for (int64_t i = 0; i < n; i++) a[i].k = a[i].v = vals[i];
  kv_msort(a, t, n, ascending);
  for (int64_t i = 0; i < n; i++) vals[i] = a[i].v;
  free(a); free(t);
}
typedef struct { int32_t k; float v; } iv_t;
static void iv_msort(iv_t *a, iv_t *tmp, int64_t n, int asc) {
  if (n < 2) return;
  int64_t h = n / 2;
  iv_msort(a, tmp, h, asc);
  iv_msort(a + h, tmp, n - h, asc);
  int64_t i = 0, j = h, o = 0;
  while (i < h && j < n) {
    int take_right = asc ? (a[i].k > a[j].k) : (a[i].k < a[j].k);
    tmp[o++] = take_right ? a[j++] : a[i++];
  }
  while (i < h) tmp[o++] = a[i++];
  while (j < n) tmp[o++] = a[j++];
  memcpy(a, tmp, sizeof(iv_t) * n);
}
void orc_sort_pairs(int32_t *keys, float *vals, int64_t n, int ascending) {
  if (n < 2) return;
  iv_t *a = (iv_t *)malloc(sizeof(iv_t) * n), *t = (iv_t *)malloc(sizeof(iv_t) * n);
  for (int64_t i = 0; i < n; i++) { a[i].k = keys[i]; a[i].v = vals[i]; }
  iv_msort(a, t, n, ascending);
  for (int64_t i = 0; i < n; i++) { keys[i] = a[i].k; vals[i] = a[i].v; }
  free(a); free(t);
}

int orc_topk(const orc_node *expr, const orc_node *cond, const orc_col_t *cols, int ncols,
             int64_t row0, int64_t row1, int descending, int64_t k, int64_t offset, float *out,
             int64_t *out_n, int contract, char *err, size_t errlen) {
  int64_t n = row1 - row0, m = 0;
  float *all = (float *)malloc(sizeof(float) * (n ? n : 1));
  if (orc_filter_compact(expr, cond, cols, ncols, row0, row1, all, &m, contract, err, errlen)) { free(all); return 1; }
  orc_sort_float(all, m, !descending); /* jit_sort_float then truncate: warpdb.cpp:453-455,483-495 */
  int64_t o = 0;
  for (int64_t i = offset; i < m && o < k; i++) out[o++] = all[i];
  *out_n = o;
  free(all);
  return 0;
}

/* ---- inner equi-join (definition: nested loop, probe-major) ------------------------------- */
typedef struct { int64_t k, r; } kr_t;
static void kr_msort(kr_t *a, kr_t *tmp, int64_t n) {   /* stable: equal keys keep their row order */
  if (n < 2) return;
  int64_t h = n / 2;
  kr_msort(a, tmp, h);
  kr_msort(a + h, tmp, n - h);
  int64_t i = 0, j = h, o = 0;
  while (i < h && j < n) tmp[o++] = (a[j].k < a[i].k) ? a[j++] : a[i++];
  while (i < h) tmp[o++] = a[i++];
  while (j < n) tmp[o++] = a[j++];
  memcpy(a, tmp, sizeof(kr_t) * n);
}
int64_t orc_join_pairs(const int64_t *probe, int64_t n, const int64_t *build, int64_t m, int indexed,
                       int64_t *out_probe, int64_t *out_build, int64_t cap) {
  int64_t o = 0;
  if (!indexed) {
    for (int64_t i = 0; i < n; i++)
      for (int64_t j = 0; j < m; j++)
        if (probe[i] == build[j]) {
          if (out_probe && o < cap) out_probe[o] = i;
          if (out_build && o < cap) out_build[o] = j;
          o++;
        }
    return o;
  }
  kr_t *a = (kr_t *)malloc(sizeof(kr_t) * (m ? m : 1)), *t = (kr_t *)malloc(sizeof(kr_t) * (m ? m : 1));
  for (int64_t j = 0; j < m; j++) { a[j].k = build[j]; a[j].r = j; }
  kr_msort(a, t, m);
  for (int64_t i = 0; i < n; i++) {
    int64_t lo = 0, hi = m;
    while (lo < hi) { int64_t mid = lo + (hi - lo) / 2; if (a[mid].k < probe[i]) lo = mid + 1; else hi = mid; }
    for (int64_t q = lo; q < m && a[q].k == probe[i]; q++) {
      if (out_probe && o < cap) out_probe[o] = i;
      if (out_build && o < cap) out_build[o] = a[q].r;
      o++;
    }
  }
  free(a); free(t);
  return o;
}

void orc_shard_range(int64_t n, int ndev, int dev, int64_t *start, int64_t *end) { /* multi_gpu_utils.cpp:24-31 */
  int64_t chunk = (n + ndev - 1) / ndev;
  int64_t s = (int64_t)dev * chunk, e = s + chunk < n ? s + chunk : n;
  if (s > n) s = n;
  if (e < s) e = s;
  *start = s;
  *end = e;
}

/* ------------------------------------------------------------------------------------------
 * WarpDB::query  (src/warpdb.cpp:199-257)
 * ---------------------------------------------------------------------------------------- */
static int split_where(const char *q, char **expr_part, char **where_part) { /* :204-213 */
  size_t n = strlen(q);
  char *up = xstrdup(q);
  for (size_t i = 0; i < n; i++) up[i] = (char)toupper((unsigned char)up[i]);
  char *w = strstr(up, "WHERE");
  if (w) {
    size_t pos = (size_t)(w - up);
    *expr_part = xstrndup(q, pos);
    *where_part = xstrdup(q + pos + 5);
  } else {
    *expr_part = xstrdup(q);
    *where_part = xstrdup("");
  }
  free(up);
  return 0;
}
int orc_query(const char *query, const orc_col_t *cols, int ncols, int64_t nrows, float *out,
              uint8_t *mask, int contract, int nthreads, char *err, size_t errlen) {
  char sub[512];
  if (!query || !*query) { set_err(err, errlen, "Empty query expression"); return 1; }
  char *ep, *wp;
  split_where(query, &ep, &wp);
  orc_node *e = orc_parse_expression(ep, sub, sizeof sub), *c = NULL;
  int rc = 1;
  if (!e) { set_err(err, errlen, "Failed to parse expression: %s", sub); goto done; }
  if (orc_validate(e, cols, ncols, err, errlen)) goto done;
  if (*wp) {
    c = orc_parse_expression(wp, sub, sizeof sub);
    if (!c) { set_err(err, errlen, "Failed to parse WHERE clause: %s", sub); goto done; }
    if (orc_validate(c, cols, ncols, sub, sizeof sub)) { set_err(err, errlen, "Failed to parse WHERE clause: %s", sub); goto done; }
  }
  rc = orc_project_filter(e, c, cols, ncols, 0, nrows, out, mask, contract, nthreads, err, errlen);
done:
  orc_free(e); orc_free(c);
  free(ep); free(wp);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * WarpDB::query_sql, host semantics "B" (src/warpdb.cpp:297-498 as disentangled in SURVEY F5
 * and Appendix D)
 * ---------------------------------------------------------------------------------------- */
static float eval_having(const orc_node *n, const agg_slot *g) { /* warpdb.cpp:387-417: float arithmetic */
  if (!n) return 1.0f;
  if (n->type == N_CONST) return strtof(n->text, NULL);
  if (n->type == N_BIN) {
    float l = eval_having(n->l, g), r = eval_having(n->r, g);
    switch (op_from_text(n->text)) {
    case OP_ADD: return l + r;
    case OP_SUB: return l - r;
    case OP_MUL: return l * r;
    case OP_DIV: return l / r;
    case OP_GT: return l > r;
    case OP_LT: return l < r;
    case OP_GE: return l >= r;
    case OP_LE: return l <= r;
    case OP_EQ: return l == r;
    case OP_NE: return l != r;
    case OP_AND: return (l != 0.0f) && (r != 0.0f);   /* not in the reference's eval (returns 0) */
    case OP_OR: return (l != 0.0f) || (r != 0.0f);
    default: return 0.0f;
    }
  }
  if (n->type == N_AGG) {
    switch (n->agg) {
    case ORC_SUM: return (float)g->sum;
    case ORC_AVG: return (float)(g->sum / g->count);
    case ORC_COUNT: return (float)g->count;
    case ORC_MIN: return (float)g->mn;
    default: return (float)g->mx;
    }
  }
  return 0.0f;
}
static int cmp_f_asc(const void *a, const void *b) {
  float x = *(const float *)a, y = *(const float *)b;
  return x < y ? -1 : x > y;
}
int orc_query_sql(const char *sql, const orc_col_t *cols, int ncols, int64_t nrows, float *out,
                  int64_t cap, int64_t *out_n, int contract, char *err, size_t errlen) {
  char sub[512];
  orc_query_t *q = orc_parse_query(sql, 1, sub, sizeof sub);
  if (!q) { set_err(err, errlen, "Failed to parse SQL: %s", sub); return 1; }
  int rc = 1;
  float *res = NULL, *keys = NULL;
  int64_t n = 0;
#define VCTX(node, ctx) do { if ((node) && orc_validate((node), cols, ncols, sub, sizeof sub)) { \
    set_err(err, errlen, "%s: %s", ctx, sub); goto done; } } while (0)
  for (int i = 0; i < q->n_select; i++) VCTX(q->select_list[i], "SELECT clause");
  for (int i = 0; i < q->n_joins; i++) VCTX(q->join_conds[i], "JOIN condition");
  VCTX(q->where, "WHERE clause");
  for (int i = 0; i < q->n_group; i++) VCTX(q->group_keys[i], "GROUP BY");
  VCTX(q->order_expr, "ORDER BY");
#undef VCTX
  if (q->n_select < 1) { set_err(err, errlen, "Empty select list"); goto done; }
  if (q->has_group) {
    const orc_node *s0 = q->select_list[0];
    if (s0->type != N_AGG) { set_err(err, errlen, "Only aggregation queries supported with GROUP BY"); goto done; }
    if (q->n_group < 1) { set_err(err, errlen, "GROUP BY needs a key"); goto done; }
    agg_slot *g;
    int64_t ng;
    if (group_build(s0->l, q->group_keys[0], q->where, s0->agg, cols, ncols, 0, nrows, contract, &g, &ng, err, errlen)) goto done;
    /* std::map order = key ascending; ORDER BY sorts by key with the asc flag (warpdb.cpp:370-371) */
    qsort(g, ng, sizeof(agg_slot), (q->has_order && !q->order_asc) ? cmp_slot_desc : cmp_slot_asc);
    res = (float *)malloc(sizeof(float) * (ng ? ng : 1));
    for (int64_t i = 0; i < ng; i++) {
      if (eval_having(q->having, &g[i]) == 0.0f) continue;
      res[n++] = agg_result(&g[i], s0->agg);
    }
    free(g);
    if (q->distinct) {
      qsort(res, n, sizeof(float), cmp_f_asc);
      int64_t u = 0;
      for (int64_t i = 0; i < n; i++)
        if (u == 0 || res[u - 1] != res[i]) res[u++] = res[i];
      n = u;
    }
  } else {
    res = (float *)malloc(sizeof(float) * (nrows ? nrows : 1));
    if (orc_filter_compact(q->select_list[0], q->where, cols, ncols, 0, nrows, res, &n, contract, err, errlen)) goto done;
    char a[1024], b[1024];
    int same_expr = 0;
    if (q->has_order) {
      orc_to_cuda_expr(q->select_list[0], a, sizeof a);
      orc_to_cuda_expr(q->order_expr, b, sizeof b);
      same_expr = !strcmp(a, b);
    }
    if (q->distinct) { /* warpdb.cpp:463-468 */
      qsort(res, n, sizeof(float), cmp_f_asc);
      int64_t u = 0;
      for (int64_t i = 0; i < n; i++)
        if (u == 0 || res[u - 1] != res[i]) res[u++] = res[i];
      n = u;
      if (q->has_order && !same_expr) { set_err(err, errlen, "DISTINCT with ORDER BY on a different expression is not supported"); goto done; }
    }
    if (q->has_order) { /* warpdb.cpp:470-476 (keyed sort), :453-455 (same expr -> jit_sort_float) */
      if (same_expr) orc_sort_float(res, n, q->order_asc);
      else {
        keys = (float *)malloc(sizeof(float) * (n ? n : 1));
        int64_t nk = 0;
        if (orc_filter_compact(q->order_expr, q->where, cols, ncols, 0, nrows, keys, &nk, contract, err, errlen)) goto done;
        kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (n ? n : 1)), *tmp = (kv_t *)malloc(sizeof(kv_t) * (n ? n : 1));
        for (int64_t i = 0; i < n; i++) { kv[i].k = keys[i]; kv[i].v = res[i]; }
        kv_msort(kv, tmp, n, q->order_asc);
        for (int64_t i = 0; i < n; i++) res[i] = kv[i].v;
        free(kv); free(tmp);
      }
    }
  }
  { /* OFFSET then LIMIT: warpdb.cpp:485-495 */
    int64_t off = q->has_offset ? q->offset : 0;
    if (off > n) off = n;
    int64_t m = n - off;
    if (q->has_limit && q->limit < m) m = q->limit;
    if (m > cap) { set_err(err, errlen, "oracle: result of %lld rows exceeds capacity", (long long)m); goto done; }
    memcpy(out, res + off, sizeof(float) * m);
    *out_n = m;
  }
  rc = 0;
done:
  free(res); free(keys);
  orc_free_query(q);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * synthetic columns: splitmix64 finaliser over (seed, row); identical on the GPU
 * (warpdb_b200/csrc/synth.cuh).  Not part of the reference (it has no generator).
 * ---------------------------------------------------------------------------------------- */
uint64_t orc_mix64(uint64_t seed, uint64_t row) {
  uint64_t z = seed + (row + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
void orc_synth_f32(float *out, int64_t n, uint64_t seed, float lo, float hi, int64_t row0) {
  float span = hi - lo;
  for (int64_t i = 0; i < n; i++) {
    float u = (float)(orc_mix64(seed, (uint64_t)(row0 + i)) >> 40) * 0x1p-24f; /* [0,1) exactly representable */
    out[i] = fmaf(u, span, lo);
  }
}
void orc_synth_i32(int32_t *out, int64_t n, uint64_t seed, int32_t lo, int32_t hi_excl, int64_t row0) {
  uint64_t range = (uint64_t)((int64_t)hi_excl - (int64_t)lo);
  for (int64_t i = 0; i < n; i++) {
    uint64_t h = orc_mix64(seed, (uint64_t)(row0 + i)) >> 32;
    out[i] = (int32_t)((int64_t)lo + (int64_t)((h * range) >> 32));
  }
}
