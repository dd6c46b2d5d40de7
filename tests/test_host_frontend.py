"""CPU: the product's C++ front end (warpdb_b200/csrc/host/expression.cpp, through the pywarpdb
module) against the golden vectors generated from the reference's own parser, and against the
reference tests that exercise the front end (tests/test_expression.cpp, precedence_tests.cpp,
tokenizer_tests.cpp, expression_tests.cpp, parsing_error_tests.cpp, parse_query_error_test.cpp,
tokenize_error_test.cpp, query_parser_test.cpp, identifier_validation_test.cpp)."""
import pytest

from warpdb_b200 import build as wbuild


@pytest.fixture(scope="module")
def pw():
    wbuild.build_host()
    from warpdb_b200 import pywarpdb
    return pywarpdb


def test_expressions_match_reference(pw, golden_frontend):
    n = 0
    for e in golden_frontend:
        if e["kind"] != "E":
            continue
        n += 1
        if e["ok"]:
            assert pw.expr_to_cuda(e["text"]) == e["out"], e["text"]
        else:
            with pytest.raises(RuntimeError) as ei:
                pw.expr_to_cuda(e["text"])
            assert str(ei.value) == e["out"], e["text"]
    assert n > 60


def test_tokens_match_reference(pw, golden_frontend):
    for e in golden_frontend:
        if e["kind"] != "T":
            continue
        if e["ok"]:
            assert pw.tokenize_dump(e["text"]) == e["out"], e["text"]
        else:
            with pytest.raises(RuntimeError) as ei:
                pw.tokenize_dump(e["text"])
            assert str(ei.value) == e["out"], e["text"]


def test_strict_queries_match_reference(pw, golden_frontend):
    n = 0
    for e in golden_frontend:
        if e["kind"] != "Q":
            continue
        n += 1
        if e["ok"]:
            assert pw.query_summary(e["text"]) == e["out"], e["text"]
        else:
            with pytest.raises(RuntimeError) as ei:
                pw.query_summary(e["text"])
            assert str(ei.value) == e["out"], e["text"]
    assert n > 40


def test_extended_grammar_is_a_superset(pw, golden_frontend):
    # whatever the reference accepts parses identically in extended mode
    for e in golden_frontend:
        if e["kind"] == "Q" and e["ok"]:
            assert pw.query_summary(e["text"], True) == e["out"], e["text"]
    # and the statements the reference's own tests use but its grammar rejects now parse
    s = pw.query_summary("SELECT price FROM test ORDER BY price DESC OFFSET 1 LIMIT 2", True)      # sql_features_test.cpp:33
    assert "order=price[idx]:DESC limit=2 offset=1" in s
    s = pw.query_summary("SELECT SUM(price) FROM test GROUP BY quantity HAVING SUM(price) > 15 ORDER BY quantity ASC", True)  # :36
    assert "having=(price[idx] > 15.0f) order=quantity[idx]:ASC" in s
    s = pw.query_summary("SELECT SUM(price) FROM test GROUP BY quantity HAVING COUNT(price) > 1", True)   # having_distinct_test.cpp:7
    assert "having=(price[idx] > 1.0f) order=-" in s
    s = pw.query_summary("SELECT price FROM t ORDER BY price LIMIT 5", True)
    assert "order=price[idx]:ASC limit=5" in s
    s = pw.query_summary("SELECT SUM(price) FROM t GROUP BY quantity LIMIT 3", True)
    assert "group=[quantity[idx]]" in s and "limit=3" in s


def test_reference_parser_tests(pw):
    # tests/test_expression.cpp
    assert pw.expr_to_cuda("price > 10") == "(price[idx] > 10.0f)"
    assert pw.expr_to_cuda("quantity <= 5") == "(quantity[idx] <= 5.0f)"
    assert pw.expr_to_cuda("discount(price, 0.9)") == "discount(price[idx], 0.9f)"
    assert pw.expr_to_cuda("price > 10 AND quantity < 5") == "((price[idx] > 10.0f) && (quantity[idx] < 5.0f))"
    assert pw.expr_to_cuda("price > 10 OR quantity < 5") == "((price[idx] > 10.0f) || (quantity[idx] < 5.0f))"
    # tests/precedence_tests.cpp
    assert pw.expr_to_cuda("price + quantity * 2") == "(price[idx] + (quantity[idx] * 2.0f))"
    assert pw.expr_to_cuda("(price + quantity) * 2") == "((price[idx] + quantity[idx]) * 2.0f)"
    # tests/tokenizer_tests.cpp
    toks = pw.tokenize_dump("price > 10").strip().split("\n")
    assert [t.split(":")[0] for t in toks] == ["Identifier", "Operator", "Number", "End"]
    toks = pw.tokenize_dump("(price + 5) * quantity").strip().split("\n")
    assert [t.split(":")[0] for t in toks] == ["Operator", "Identifier", "Operator", "Number", "Operator", "Operator", "Identifier", "End"]
    assert "Keyword:AND" in pw.tokenize_dump("price > 10 AND quantity < 5") and "Keyword:OR" not in pw.tokenize_dump("price > 10 AND quantity < 5")
    # tests/expression_tests.cpp, parsing_error_tests.cpp, tokenize_error_test.cpp, parse_query_error_test.cpp
    with pytest.raises(RuntimeError, match="Unexpected token"):
        pw.expr_to_cuda("1 2")
    with pytest.raises(RuntimeError, match="Unknown character"):
        pw.tokenize_dump("price & 5")
    with pytest.raises(RuntimeError, match="Unexpected token"):
        pw.query_summary("SELECT price FROM test EXTRA")
    with pytest.raises(RuntimeError, match=r"Expected '\)'"):
        pw.expr_to_cuda("(price + 5")
    with pytest.raises(RuntimeError, match="line 1.*column|column.*line 1"):
        pw.tokenize_dump("price # 1\n")
    with pytest.raises(RuntimeError, match="line.*column"):
        pw.query_summary("SELECT price")
    # tests/query_parser_test.cpp
    s = pw.query_summary("SELECT SUM(price), quantity FROM sales JOIN items ON sales.id = items.id WHERE price > 10 GROUP BY quantity ORDER BY price DESC LIMIT 5")
    assert "select=[AGG0(price[idx]);quantity[idx]]" in s and "joins=[items:" in s and "where=(price[idx] > 10.0f)" in s
    assert "group=[quantity[idx]]" in s and "order=price[idx]:DESC" in s and "limit=5" in s


def test_optimizer_analyze_condition(pw):
    """What the reference's analyze_condition stub (src/optimizer.cpp:13-17) was meant to decide."""
    r = [("price", 10.5, 30.0), ("quantity", 2.0, 5.0)]
    assert pw.analyze_condition("price > 10", r) == (True, False)
    assert pw.analyze_condition("price > 30", r) == (False, True)
    assert pw.analyze_condition("price > 20", r) == (False, False)
    assert pw.analyze_condition("price >= 10.5 AND quantity < 6", r) == (True, False)
    assert pw.analyze_condition("price > 20 AND quantity > 5", r) == (False, True)
    assert pw.analyze_condition("price > 20 OR quantity <= 5", r) == (True, False)
    assert pw.analyze_condition("100 < price", r) == (False, True)
    assert pw.analyze_condition("price != 7", r) == (True, False)
    assert pw.analyze_condition("price * 2 > 1", r) == (False, False)      # not a col-vs-const comparison: unknown
    assert pw.analyze_condition("other > 1", r) == (False, False)
