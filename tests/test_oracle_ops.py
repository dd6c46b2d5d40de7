"""CPU: oracle operators against the known-answer vectors of the reference's fixtures
(reference tests jit_arch_test.cpp:23-31, extended_types_test.cpp:8-10,
sql_features_test.cpp:11-37, having_distinct_test.cpp:7-12) and basic properties."""
import numpy as np
import pytest

from oracle import pyoracle as orc


def hexf(a):
    return [format(int(x), "08x") for x in np.asarray(a, np.float32).view(np.uint32)]


def test_kat_projection(fixtures):
    t = fixtures["test"]
    out, mask = orc.query("price * quantity WHERE price > 10", t)
    assert mask.tolist() == [1, 1, 1, 1]
    assert hexf(out) == ["41fc0000", "42a00000", "41f40000", "43160000"]
    out, _ = orc.query("price * quantity * 1.08", t)
    assert hexf(out) == ["4208147b", "42accccd", "4203c290", "43220000"]
    out, _ = orc.query("price * 0.9", t)
    assert hexf(out) == ["41173333", "41900000", "415b9999", "41d80000"]
    out2, _ = orc.query("discount(price, 0.9)", t)
    assert hexf(out2) == hexf(out)
    out, _ = orc.query("price + 1", t)
    assert out.tolist() == [11.5, 21.0, 16.25, 31.0]
    out, _ = orc.query("price * discount", fixtures["extended"])
    assert hexf(out) == ["3f866667", "40800000", "3f433333", "40900000"]
    assert int(out[0]) == 1  # tests/extended_types_test.cpp:9


def test_kat_dense_filter_leaves_slots_untouched(fixtures):
    out, mask = orc.query("price * 0.9 WHERE price > 20", fixtures["test"], fill=-7.0)
    assert mask.tolist() == [0, 0, 0, 1]
    assert out.tolist() == [-7.0, -7.0, -7.0, 27.0]
    out, mask = orc.query("price WHERE price > 15", fixtures["test"], fill=-1.0)
    assert out.tolist() == [-1.0, 20.0, 15.25, 30.0]


def test_jit_arch_kat():
    # tests/jit_arch_test.cpp: expr "price" over a 1-row table gives the price back bit-exactly
    out, _ = orc.query("price", {"price": np.array([2.0], np.float32), "quantity": np.array([0], np.int32)})
    assert out.tolist() == [2.0]


def test_kat_sql(fixtures):
    t = fixtures["test"]
    assert orc.query_sql("SELECT SUM(price) FROM test GROUP BY quantity ORDER BY quantity ASC", t).tolist() == [15.25, 10.5, 20.0, 30.0]
    assert orc.query_sql("SELECT SUM(price) FROM test GROUP BY quantity", t).tolist() == [15.25, 10.5, 20.0, 30.0]
    assert orc.query_sql("SELECT price FROM test ORDER BY price DESC LIMIT 2", t).tolist() == [30.0, 20.0]
    assert orc.query_sql("SELECT price FROM test ORDER BY price DESC LIMIT 5", t).tolist() == [30.0, 20.0, 15.25, 10.5]
    assert orc.query_sql("SELECT price FROM test ORDER BY price DESC LIMIT 2 OFFSET 1", t).tolist() == [20.0, 15.25]
    assert len(orc.query_sql("SELECT SUM(price) FROM test GROUP BY quantity HAVING SUM(price) > 15 ORDER BY quantity ASC", t)) == 3
    assert len(orc.query_sql("SELECT SUM(price) FROM test GROUP BY quantity HAVING COUNT(price) > 1 ORDER BY quantity ASC", t)) == 0
    r = orc.query_sql("SELECT DISTINCT quantity FROM test ORDER BY quantity DESC", t)
    assert r.tolist() == [5.0, 4.0, 3.0, 2.0]
    assert orc.query_sql("SELECT price * 0.9 FROM t WHERE price > 20", t).tolist() == [27.0]


def test_group_first_appearance_order(fixtures):
    g = orc.group_agg("price", "quantity", None, fixtures["test"], order=orc.ORDER_FIRST)
    assert g["keys"].tolist() == [3, 4, 2, 5] and g["vals"].tolist() == [10.5, 20.0, 15.25, 30.0]


def test_error_messages(fixtures):
    t = fixtures["test"]
    with pytest.raises(orc.OracleError, match="Empty query expression"):
        orc.query("", t)
    with pytest.raises(orc.OracleError, match="Failed to parse expression: Unexpected tokens remaining: 2"):
        orc.query("1 2", t)
    with pytest.raises(orc.OracleError, match="Unknown column: foo"):
        orc.query("foo + 1", t)
    with pytest.raises(orc.OracleError, match="Failed to parse WHERE clause: Unknown column: bar"):
        orc.query("price WHERE bar > 1", t)
    with pytest.raises(orc.OracleError, match="SELECT clause: Unknown column: foo"):
        orc.query_sql("SELECT foo FROM test", t)
    with pytest.raises(orc.OracleError, match="Only aggregation queries supported with GROUP BY"):
        orc.query_sql("SELECT price FROM test GROUP BY quantity ORDER BY quantity ASC", t)
    with pytest.raises(orc.OracleError, match="Failed to parse SQL: Expected keyword 'FROM'"):
        orc.query_sql("SELECT price", t)


def test_typed_semantics_and_contraction():
    t = {"a": np.array([7, -7, 2000000000], np.int32), "b": np.array([2, 2, 2], np.int32),
         "x": np.array([0.1, 1e10, 3.3], np.float32), "d": np.array([0.1, 0.2, 0.3], np.float64),
         "l": np.array([1 << 40, 5, -3], np.int64)}
    out, _ = orc.query("a / b", t)                 # int / int truncates toward zero (C++ typing, jit.cpp:75-79)
    assert out.tolist() == [3.0, -3.0, 1e9]
    out, _ = orc.query("a * b", t)                 # int32 wraparound, then -> float
    assert out[2] == np.float32(np.int32(np.int64(4000000000) - (1 << 32)))
    out, _ = orc.query("a / 2", t)                 # literal is float -> float division
    assert out.tolist() == [3.5, -3.5, 1e9]
    out, _ = orc.query("d * 2", t)                 # double arithmetic, one rounding to float at the store
    assert out.tolist() == [np.float32(0.1 * 2), np.float32(0.2 * 2), np.float32(0.3 * 2)]
    out, _ = orc.query("l + a", t)
    assert out[0] == np.float32((1 << 40) + 7)
    # fma contraction (NVRTC default --fmad=true): x*x+1 differs from the two-rounding result for some x
    x = orc.synth_f32(4096, 1, 0.0, 10.0)
    c1, _ = orc.query("x * x + 1", {"x": x}, contract=True)
    c0, _ = orc.query("x * x + 1", {"x": x}, contract=False)
    ref = (x.astype(np.float64) * x.astype(np.float64) + 1.0).astype(np.float32)
    assert np.array_equal(c1, ref)  # fmaf == exact product, one rounding (no double rounding issue at these magnitudes)
    assert np.array_equal(c0, (x * x + np.float32(1)).astype(np.float32))
    assert not np.array_equal(c0, c1)


def test_synth_properties_and_shards():
    p = orc.synth_f32(100000, 0xC0FFEE, 0.0, 100.0)
    q = orc.synth_i32(100000, 0xC0FFEF, 1, 101)
    assert p.min() >= 0.0 and p.max() < 100.0 and abs(p.mean() - 50) < 1
    assert q.min() == 1 and q.max() == 100
    # counter based: any window regenerates identically
    assert np.array_equal(orc.synth_f32(1000, 0xC0FFEE, 0.0, 100.0, row0=5000), p[5000:6000])
    assert np.array_equal(orc.synth_i32(1000, 0xC0FFEF, 1, 101, row0=99000), q[99000:])
    # shards: multi_gpu_utils.cpp:24-31
    assert [orc.shard_range(10, 4, d) for d in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [orc.shard_range(4, 8, d) for d in range(8)] == [(0, 1), (1, 2), (2, 3), (3, 4)] + [(4, 4)] * 4
    cover = [orc.shard_range(1000003, 8, d) for d in range(8)]
    assert cover[0][0] == 0 and cover[-1][1] == 1000003 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))


def test_compact_group_topk_against_numpy():
    n = 200000
    t = {"price": orc.synth_f32(n, 11, 0.0, 40.0), "quantity": orc.synth_i32(n, 12, 0, 1000)}
    m = t["price"] > np.float32(20)
    c = orc.filter_compact("price * 0.9", "price > 20", t)
    assert np.array_equal(c, (t["price"][m] * np.float32(0.9)).astype(np.float32))
    g = orc.group_agg("price", "quantity", None, t)
    sums = np.bincount(t["quantity"], weights=t["price"].astype(np.float64), minlength=1000)
    assert g["keys"].tolist() == list(range(1000))
    np.testing.assert_allclose(g["sums"], sums, rtol=1e-12)
    assert np.array_equal(g["counts"], np.bincount(t["quantity"], minlength=1000))
    top = orc.topk("price", None, t, descending=True, k=5)
    assert np.array_equal(top, np.sort(t["price"])[::-1][:5])
    low = orc.topk("price", "quantity < 10", t, descending=False, k=7, offset=2)
    assert np.array_equal(low, np.sort(t["price"][t["quantity"] < 10])[2:9])


def test_stable_sorts_match_bubble_sort():
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 8, 64).astype(np.int32)
    vals = np.arange(64, dtype=np.float32)
    for asc in (True, False):
        k, v = keys.copy(), vals.copy()
        for i in range(len(k) - 1):               # jit.cpp:254-262 verbatim semantics
            for j in range(len(k) - i - 1):
                if (k[j] > k[j + 1]) if asc else (k[j] < k[j + 1]):
                    k[j], k[j + 1] = k[j + 1], k[j]
                    v[j], v[j + 1] = v[j + 1], v[j]
        ok, ov = orc.sort_pairs(keys, vals, asc)
        assert np.array_equal(ok, k) and np.array_equal(ov, v)
    f = rng.standard_normal(50).astype(np.float32)
    assert np.array_equal(orc.sort_float(f, True), np.sort(f))
    assert np.array_equal(orc.sort_float(f, False), np.sort(f)[::-1])


def test_join_pairs_is_the_nested_loop():
    # JOIN is parsed by the reference (src/expression.cpp:375-401) and never executed, so nothing of the
    # reference pins it: the oracle states SQL's inner equi-join as the probe-major nested loop, and its
    # indexed variant (what the larger GPU tests compare with) must enumerate the same pairs in the same order
    rng = np.random.default_rng(5)
    for n, m, span in [(0, 5, 3), (5, 0, 3), (1, 1, 1), (400, 300, 40), (300, 400, 100000), (64, 64, 1)]:
        probe, build = rng.integers(-span, span, n), rng.integers(-span, span, m)
        pi, pj = orc.join_pairs(probe, build, indexed=False)
        qi, qj = orc.join_pairs(probe, build, indexed=True)
        ii, jj = np.nonzero(probe[:, None] == build[None, :]) if n and m else (np.empty(0, np.int64), np.empty(0, np.int64))
        assert np.array_equal(pi, ii) and np.array_equal(pj, jj)
        assert np.array_equal(qi, ii) and np.array_equal(qj, jj)
    lim = np.array([np.iinfo(np.int64).min, -1, 0, np.iinfo(np.int64).max])
    i, j = orc.join_pairs(lim, lim[::-1].copy())
    assert i.tolist() == [0, 1, 2, 3] and j.tolist() == [3, 2, 1, 0]
