#!/usr/bin/env python
"""Generate tests/golden/frontend.json by running the REFERENCE's own front end
(/root/reference/src/expression.cpp, built by oracle/Makefile into oracle/_ref/ref_front)
over the inputs below.  Run in the build container only (the GPU box has no /root/reference):

    make -C oracle _ref/ref_front && python tests/golden/make_frontend_golden.py

Every entry is {"kind": T|E|Q, "text": ..., "ok": bool, "out": payload-or-error-message}.
Inputs come from the reference's tests (tests/test_expression.cpp, precedence_tests.cpp,
tokenizer_tests.cpp, expression_tests.cpp, parsing_error_tests.cpp, parse_query_error_test.cpp,
tokenize_error_test.cpp, query_parser_test.cpp, sql_features_test.cpp, having_distinct_test.cpp),
BASELINE.json's configs and the grammar quirks listed in SURVEY.md Appendix A.
"""
import json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_front")

EXPRS = [
    "price > 10", "quantity <= 5", "discount(price, 0.9)", "price > 10 AND quantity < 5",
    "price > 10 OR quantity < 5", "price + quantity * 2", "(price + quantity) * 2", "1 2",
    "(price + 5", "price", "price + 1", "invalid@", "price * quantity", "price * quantity * 1.08",
    "price * 0.9", "price * discount", "price > 20", "price >= 20.5", "price != quantity",
    "price == 3", "price = 3", "a.b + c_d", ".5 * price", "5. * price", "1.2.3", "price / quantity - 2",
    "price - quantity - 1", "price / 2 / 3", "f()", "f(a, b, c)", "f(g(a), 2)", "f(a", "f(a,", "f(,)",
    "-price", "price * -1", "1e3", "(price > 1) * 2", "f(price > 1)", "price > 1 > 0",
    "price > 10 AND quantity < 5 OR price < 2", "price > 10 OR quantity < 5 AND price < 2",
    "a AND b AND c", "a OR b OR c", "sqrtf(price) + 1", "((price))", "()", "", "   ", "price +",
    "* price", "price ! quantity", "price !", "SUM(price)", "select", "price > 10 and quantity < 5",
    "price > 10 Or quantity < 5", "PRICE", "price\n+\n1", "x y", "1 + 2 * 3 - 4 / 5",
    "price * (quantity + 1) * 1.08", "fminf(price, 50) * fmaxf(quantity, 2)", "price , 1", "price . 1",
    "price)", "a<b", "a<=b", "a>=b", "a<>b", "a==b", "a=b", "a!=b", "10", "10.", "007", "0.90",
]
TOKENS = [
    "price > 10", "(price + 5) * quantity", "price > 10 AND quantity < 5", "price & 5", "price # 1\n",
    "a.b >= .5", "x\n  y", "SELECT sum(price) FROM t", "a!=b", "a!b", "1.2.3", "a = = b", "tab\tsep",
    "select Select SELECT", "group by order by asc desc limit offset having distinct over partition",
    "price $", "\n\n  @", "a ; b", "[", "a_1 _b B2.c.d", "3.14.15.9", "<= >= == != = < > !",
]
QUERIES = [
    "SELECT SUM(price), quantity FROM sales JOIN items ON sales.id = items.id WHERE price > 10 GROUP BY quantity ORDER BY price DESC LIMIT 5",
    "SELECT price FROM test EXTRA", "SELECT price", "SELECT foo FROM test",
    "SELECT SUM(price) FROM test GROUP BY quantity ORDER BY quantity ASC",
    "SELECT price FROM test ORDER BY price DESC LIMIT 2",
    "SELECT price FROM test ORDER BY price DESC OFFSET 1 LIMIT 2",
    "SELECT price FROM test ORDER BY price DESC LIMIT 2 OFFSET 1",
    "SELECT SUM(price) FROM test GROUP BY quantity HAVING SUM(price) > 15 ORDER BY quantity ASC",
    "SELECT SUM(price) FROM test GROUP BY quantity HAVING 1 > 0 ORDER BY quantity ASC",
    "SELECT SUM(price) FROM test GROUP BY quantity HAVING 1 > 0 LIMIT 3",
    "SELECT DISTINCT quantity FROM test ORDER BY quantity DESC",
    "SELECT SUM(price) FROM t GROUP BY quantity",
    "SELECT price FROM t ORDER BY price DESC LIMIT 5",
    "SELECT discount(price, 0.9) FROM t ORDER BY discount(price, 0.9) DESC LIMIT 5",
    "SELECT price * 0.9 FROM t WHERE price > 20",
    "SELECT price FROM t ORDER BY price LIMIT 5", "SELECT price FROM t ORDER BY price",
    "SELECT price FROM t ORDER BY price ASC", "SELECT SUM(price) FROM t GROUP BY quantity LIMIT 3",
    "SELECT SUM(price) FROM t GROUP BY quantity, price ORDER BY quantity DESC",
    "SELECT AVG(price) FROM t GROUP BY quantity", "SELECT COUNT(price) FROM t GROUP BY quantity",
    "SELECT MIN(price) FROM t GROUP BY quantity", "SELECT MAX(price * 2) FROM t GROUP BY quantity",
    "SELECT SUM(price) OVER PARTITION BY quantity FROM t", "SELECT SUM price FROM t",
    "SELECT SUM(price) * 2 FROM t", "SELECT FROM t", "SELECT price, FROM t", "SELECT price FROM",
    "SELECT price FROM 5", "SELECT price FROM t JOIN", "SELECT price FROM t JOIN u", "SELECT price FROM t JOIN u ON",
    "SELECT price FROM t JOIN u ON t.a = u.a JOIN v ON u.b = v.b WHERE price > 1",
    "SELECT price FROM t WHERE", "SELECT price FROM t WHERE price > 1 AND quantity < 3 LIMIT 10",
    "SELECT price FROM t LIMIT", "SELECT price FROM t LIMIT x", "SELECT price FROM t LIMIT 2.5",
    "SELECT price FROM t LIMIT 3 OFFSET", "SELECT price FROM t LIMIT 3 OFFSET x", "SELECT price FROM t OFFSET 2",
    "SELECT price FROM t GROUP quantity", "SELECT price FROM t ORDER price", "SELECT price FROM t GROUP BY",
    "select price from t where price > 3", "SELECT DISTINCT price FROM t", "price * 2",
    "SELECT f(price, quantity), price FROM t", "SELECT (price + 1) * 2, quantity FROM t WHERE quantity = 2",
    "SELECT price FROM t WHERE price > 1 GROUP BY quantity HAVING 2 > 1 ORDER BY quantity DESC LIMIT 4 OFFSET 1",
    "SELECT price\nFROM t\nWHERE price > 1\nLIMIT", "SELECT price FROM t WHERE price # 1", "",
    # JOIN statements the product executes (tests/test_gpu_join.py); the reference parses them and stops there
    "SELECT price * rate FROM sales JOIN items ON sales.item = items.id",
    "SELECT sales.price * items.rate FROM sales JOIN items ON item = id WHERE items.rate > 0.5 AND sales.quantity < 5",
    "SELECT price FROM sales JOIN items ON items.id == sales.item WHERE cat == 3",
    "SELECT SUM(price * rate) FROM sales JOIN items ON sales.item = items.id GROUP BY cat",
    "SELECT AVG(price) FROM sales JOIN items ON sales.item = items.id WHERE quantity > 2 GROUP BY cat ORDER BY cat DESC",
    "SELECT COUNT(price) FROM sales JOIN items ON sales.item = items.id GROUP BY cat, quantity",
    "SELECT price * rate FROM sales JOIN items ON sales.item = items.id ORDER BY price * rate DESC LIMIT 7",
    "SELECT DISTINCT cat FROM sales JOIN items ON sales.item = items.id ORDER BY cat ASC",
    "SELECT price FROM sales JOIN items ON sales.item = items.id WHERE price > 99 LIMIT 5 OFFSET 2",
    "SELECT price * rate * boost FROM sales JOIN items ON sales.item = items.id JOIN cats ON items.cat = cats.cid WHERE quantity > 1",
    "SELECT sales.price - other.price FROM sales JOIN other ON sales.item = other.item WHERE sales.quantity > other.quantity",
    "SELECT price FROM sales JOIN items ON sales.item > items.id", "SELECT price FROM sales JOIN ON a = b",
    "SELECT price FROM sales JOIN items ON sales.item = items.id JOIN", "SELECT price FROM sales JOIN items sales.item = items.id",
]

def esc(s):
    return s.replace("\n", "\\n")

def main():
    if not os.path.exists(TOOL):
        sys.exit("build oracle/_ref/ref_front first: make -C oracle _ref/ref_front")
    reqs = [("E", e) for e in EXPRS] + [("T", t) for t in TOKENS] + [("Q", q) for q in QUERIES]
    inp = "".join(f"{k} {esc(t)}\n" for k, t in reqs)
    # a request whose text is empty still needs the two-char prefix
    out = subprocess.run([TOOL], input=inp, capture_output=True, text=True, check=True).stdout.split("\n")
    entries = []
    for (k, t), line in zip(reqs, out):
        ok = line.startswith("OK ") or line == "OK"
        payload = line[3:] if ok else line[4:]
        entries.append({"kind": k, "text": t, "ok": ok, "out": payload.replace("\\n", "\n")})
    assert len(entries) == len(reqs), (len(entries), len(reqs))
    with open(os.path.join(HERE, "frontend.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_frontend_golden.py",
                   "source": "reference src/expression.cpp (+ missing '}' after :522) via oracle/_ref/ref_front",
                   "entries": entries}, f, indent=1)
    print(f"wrote {len(entries)} entries")

if __name__ == "__main__":
    main()
