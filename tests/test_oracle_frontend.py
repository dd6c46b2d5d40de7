"""CPU: the oracle's front end against golden vectors produced by the reference's own
src/expression.cpp (tests/golden/frontend.json, generator: tests/golden/make_frontend_golden.py)."""
import pytest

from oracle import pyoracle as orc


def test_golden_has_reference_test_inputs(golden_frontend):
    texts = {(e["kind"], e["text"]) for e in golden_frontend}
    # tests/test_expression.cpp:10-33, precedence_tests.cpp:9,16, expression_tests.cpp:7
    for t in ["price > 10", "quantity <= 5", "discount(price, 0.9)", "price > 10 AND quantity < 5",
              "price > 10 OR quantity < 5", "price + quantity * 2", "(price + quantity) * 2", "1 2"]:
        assert ("E", t) in texts
    assert len(golden_frontend) > 120


def test_expressions_match_reference(golden_frontend):
    n = 0
    for e in golden_frontend:
        if e["kind"] != "E":
            continue
        n += 1
        if e["ok"]:
            assert orc.Expr(e["text"]).cuda() == e["out"], e["text"]
        else:
            with pytest.raises(orc.OracleError) as ei:
                orc.Expr(e["text"])
            assert str(ei.value) == e["out"], e["text"]
    assert n > 60


def test_tokens_match_reference(golden_frontend):
    for e in golden_frontend:
        if e["kind"] != "T":
            continue
        if e["ok"]:
            assert orc.tokenize_dump(e["text"]) == e["out"], e["text"]
        else:
            with pytest.raises(orc.OracleError) as ei:
                orc.tokenize_dump(e["text"])
            assert str(ei.value) == e["out"], e["text"]


def test_queries_match_reference(golden_frontend):
    n = 0
    for e in golden_frontend:
        if e["kind"] != "Q":
            continue
        n += 1
        if e["ok"]:
            assert orc.query_summary(e["text"]) == e["out"], e["text"]
        else:
            with pytest.raises(orc.OracleError) as ei:
                orc.query_summary(e["text"])
            assert str(ei.value) == e["out"], e["text"]
    assert n > 40


def test_having_aggregate_extension():
    # tests/sql_features_test.cpp:36 intends this; the reference grammar rejects it (golden: error).
    s = orc.query_summary("SELECT SUM(price) FROM test GROUP BY quantity HAVING SUM(price) > 15 ORDER BY quantity ASC", ext=True)
    assert "having=(price[idx] > 15.0f)" in s  # to_cuda_expr of an aggregation is its inner expression
