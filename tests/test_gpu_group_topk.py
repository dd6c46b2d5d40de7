"""GPU parity: hash GROUP BY, ORDER BY ... LIMIT, device sorts through the C ABI vs the CPU oracle.
Keys, group order, counts, compaction and top-k order are bit-exact; fp64-accumulated sums agree to
1e-6 relative (BASELINE.json north_star) -- in fact far tighter, asserted at 1e-12 on the raw sums."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import pyoracle as orc
from warpdb_b200 import _core as wc
from warpdb_b200 import ops

UDF = "__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n"
SUM_RTOL = 1e-6


@pytest.fixture(scope="module", autouse=True)
def gpu():
    assert torch.cuda.is_available()
    wc.check(wc.lib().wdb_init(0))
    wc.set_udf_source(UDF)
    yield
    wc.set_udf_source("")


def dev(table):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in table.items()}


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def cu(text):
    return orc.Expr(text).cuda() if text else None


def test_fixture_group_by_kats(fixtures):
    t = fixtures["test"]
    k, v = ops.group_agg(dev(t), "price[idx]", "quantity[idx]", order=wc.ORDER_KEY_ASC)
    assert k.cpu().tolist() == [2, 3, 4, 5] and v.cpu().tolist() == [15.25, 10.5, 20.0, 30.0]
    k, v = ops.group_agg(dev(t), "price[idx]", "quantity[idx]", order=wc.ORDER_FIRST)   # src/jit.cpp:196-213
    assert k.cpu().tolist() == [3, 4, 2, 5] and v.cpu().tolist() == [10.5, 20.0, 15.25, 30.0]
    k, v = ops.group_agg(dev(t), "price[idx]", "quantity[idx]", order=wc.ORDER_KEY_DESC)
    assert k.cpu().tolist() == [5, 4, 3, 2]


@pytest.mark.parametrize("groups,n", [(1, 10007), (7, 100003), (2, 1_100_003), (40, 1_500_007), (110, 1_200_011), (1000, 1_000_003), (100_000, 2_000_003),
                                      (3_000_000, 4_000_001)])
@pytest.mark.parametrize("agg", [wc.SUM, wc.AVG, wc.COUNT, wc.MIN, wc.MAX])
def test_group_by_matches_oracle(groups, n, agg):
    # 2 / 40 / 110 keys over more than 2^20 rows: the measured key range selects the lane-private accumulators (wdb_group_wp, mode 2)
    t = {"price": orc.synth_f32(n, 21, 0.0, 100.0), "quantity": orc.synth_i32(n, 22, -groups // 2, groups - groups // 2)}
    ref = orc.group_agg("price", "quantity", None, t, agg=agg)
    k, v = ops.group_agg(dev(t), "price[idx]", "quantity[idx]", agg=agg, expected_groups=groups)
    assert np.array_equal(k.cpu().numpy(), ref["keys"])
    got = v.cpu().numpy()
    if agg in (wc.COUNT, wc.MIN, wc.MAX):
        assert np.array_equal(bits(got), bits(ref["vals"]))
    else:
        np.testing.assert_allclose(got, ref["vals"], rtol=SUM_RTOL, atol=0)


@pytest.mark.parametrize("groups", [3, 60])
def test_tiny_key_ranges_first_appearance_order_and_where(groups):
    """Lane-private accumulators (every lane owns a copy of every accumulator): first-appearance order, WHERE, AVG / MIN."""
    n = 1_300_021
    t = {"price": orc.synth_f32(n, 41, -50.0, 50.0), "quantity": orc.synth_i32(n, 42, 5, 5 + groups)}
    d = dev(t)
    for agg in (wc.SUM, wc.AVG, wc.MIN, wc.COUNT):
        ref = orc.group_agg("price", "quantity", "price > 0 - 20", t, agg=agg, order=orc.ORDER_FIRST)
        k, v = ops.group_agg(d, "price[idx]", "quantity[idx]", cu("price > 0 - 20"), agg=agg, order=wc.ORDER_FIRST, expected_groups=groups)
        assert np.array_equal(k.cpu().numpy(), ref["keys"]), (groups, agg)
        if agg in (wc.COUNT, wc.MIN):
            assert np.array_equal(bits(v.cpu().numpy()), bits(ref["vals"]))
        else:
            np.testing.assert_allclose(v.cpu().numpy(), ref["vals"], rtol=SUM_RTOL)


def test_group_by_where_expressions_and_unknown_cardinality():
    n = 500_009
    t = {"price": orc.synth_f32(n, 31, 0.0, 40.0), "quantity": orc.synth_i32(n, 32, 0, 5000)}
    ref = orc.group_agg("price * 2 + 1", "quantity / 3", "price > 20 AND quantity < 4000", t)
    # expected_groups unknown (0): the one-shot call sizes the table and grows it on overflow
    k, v = ops.group_agg(dev(t), cu("price * 2 + 1"), cu("quantity / 3"), cu("price > 20 AND quantity < 4000"))
    assert np.array_equal(k.cpu().numpy(), ref["keys"])
    np.testing.assert_allclose(v.cpu().numpy(), ref["vals"], rtol=SUM_RTOL)
    # float-valued key expression truncates toward zero, saturates like CUDA's cvt.rzi
    t2 = {"price": np.array([1.9, -1.9, 3e9, -3e9, 2.0, 1.1], np.float32)}
    ref = orc.group_agg("price", "price", None, t2, order=orc.ORDER_FIRST)
    k, v = ops.group_agg(dev(t2), "price[idx]", "price[idx]", order=wc.ORDER_FIRST)
    assert k.cpu().tolist() == ref["keys"].tolist() == [1, -1, 2147483647, -2147483648, 2]




@pytest.mark.parametrize("opts", [{}, {"group.wp_ilp": 2}, {"group.wp_ilp": 4, "group.wp_unroll": 1}, {"group.wp_warps": 3, "group.wp_vec": 4},
                                  {"group.wp_max_span": 0}])
def test_warp_private_accumulators_any_key_distribution(opts):
    """Small-cardinality kernel (warp-private direct-indexed accumulators, tag arbitration; selected
    from the key column's min/max, gathered automatically): heavy same-key conflicts inside a warp,
    negative keys, ranges touching INT_MIN (the table's empty sentinel) and INT_MAX, a WHERE clause --
    all identical to the oracle and to the shared-atomic kernel (wp_max_span 0); sparse keys fall
    back to the atomic kernel."""
    wc.set_option("group.auto_stats_min_rows", 0)
    for k, v in opts.items():
        wc.set_option(k, v)
    try:
        n = 700_003
        rng = np.random.default_rng(7)
        pool = np.concatenate([rng.integers(-2**31, 2**31, 890, dtype=np.int64), [-2**31, 2**31 - 1, 0, -1, 1024, 2048, 4096, 1 << 20]])
        for keys in (pool[rng.integers(0, len(pool), n)], rng.integers(-1, 2, n), rng.integers(-1000, 1000, n),
                     -2**31 + rng.integers(0, 900, n), 2**31 - 1 - rng.integers(0, 1500, n)):
            t = {"price": orc.synth_f32(n, 91, -10.0, 100.0), "quantity": np.ascontiguousarray(keys, dtype=np.int32)}
            for agg, cond in ((wc.SUM, None), (wc.AVG, "price > 20"), (wc.COUNT, "price > 95"), (wc.MIN, "price > 20"), (wc.MAX, None)):
                ref = orc.group_agg("price", "quantity", cond, t, agg=agg)
                k, v = ops.group_agg(dev(t), "price[idx]", "quantity[idx]", cu(cond), agg=agg, expected_groups=2000)
                assert np.array_equal(k.cpu().numpy(), ref["keys"])
                if agg in (wc.COUNT, wc.MIN, wc.MAX):
                    assert np.array_equal(bits(v.cpu().numpy()), bits(ref["vals"]))
                else:
                    np.testing.assert_allclose(v.cpu().numpy(), ref["vals"], rtol=SUM_RTOL, atol=0)
        # first-appearance order (jit_group_sum's contract, src/jit.cpp:196-213): smallest row id per key, kept per warp
        t = {"price": orc.synth_f32(n, 91, -10.0, 100.0), "quantity": np.ascontiguousarray(rng.integers(-700, 700, n), dtype=np.int32)}
        ref = orc.group_agg("price", "quantity", "price > 50", t, agg=orc.SUM, order=orc.ORDER_FIRST)
        k, v = ops.group_agg(dev(t), "price[idx]", "quantity[idx]", cu("price > 50"), agg=wc.SUM, order=wc.ORDER_FIRST, expected_groups=2000)
        assert np.array_equal(k.cpu().numpy(), ref["keys"])
        np.testing.assert_allclose(v.cpu().numpy(), ref["vals"], rtol=SUM_RTOL, atol=0)
        # key EXPRESSIONS: the core learns their range by evaluating the expression itself (keyrange.cuh)
        t = {"price": orc.synth_f32(n, 93, -10.0, 100.0), "quantity": orc.synth_i32(n, 94, -3000, 3000)}
        for key_text in ("quantity / 7", "price / 2", "quantity * 0 + 5"):
            ref = orc.group_agg("price", key_text, "price > 5", t, agg=orc.AVG)
            k, v = ops.group_agg(dev(t), "price[idx]", cu(key_text), cu("price > 5"), agg=wc.AVG, expected_groups=2000)
            assert np.array_equal(k.cpu().numpy(), ref["keys"]), key_text
            np.testing.assert_allclose(v.cpu().numpy(), ref["vals"], rtol=SUM_RTOL, atol=0)
    finally:
        wc.set_option("group.auto_stats_min_rows", None)
        for k in opts:
            wc.set_option(k, None)


@pytest.mark.parametrize("lo,hi", [(0, 999), (-500, 1200), (100, 300), (-2**31, -2**31 + 1500), (2**31 - 900, 2**31 - 1), (0, 5000)])
def test_key_range_statistics_select_direct_indexing_and_stay_correct_when_wrong(lo, hi):
    """wdb_agg_set_key_range: a small [lo, hi] makes the kernel index its accumulators with key - lo;
    keys outside the promised range (stale statistics) must still be aggregated correctly."""
    n = 400_003
    base = {"price": orc.synth_f32(n, 95, 0.0, 100.0)}
    for klo, khi in ((0, 1000), (lo, min(hi + 1, 2**31 - 1) if hi < 2**31 - 1 else hi)):
        q = (orc.synth_i32(n, 96, 0, 1000).astype(np.int64) * max((khi - klo) // 1000, 1) + klo).clip(-2**31, 2**31 - 1).astype(np.int32)
        t = dict(base, quantity=q)
        ref = orc.group_agg("price", "quantity", None, t, agg=orc.AVG)
        tab = ops.AggTable(0, 1000, wc.NEED_SUM | wc.NEED_COUNT)
        tab.set_key_range(lo, hi)
        d = dev(t)
        half = n // 2 // 8 * 8
        tab.consume({k: v[:half] for k, v in d.items()}, "price[idx]", "quantity[idx]")
        tab.set_key_range(None, None)                                   # second chunk: statistics gathered by the core, or the atomic kernel
        tab.consume({k: v[half:] for k, v in d.items()}, "price[idx]", "quantity[idx]", row_base=half)
        out = tab.export(wc.AVG, wc.ORDER_KEY_ASC)
        tab.close()
        assert np.array_equal(out["keys"].cpu().numpy(), ref["keys"])
        assert np.array_equal(out["counts"].cpu().numpy(), ref["counts"])
        np.testing.assert_allclose(out["sums"].cpu().numpy(), ref["sums"], rtol=1e-12)


@pytest.mark.parametrize("agg", [wc.SUM, wc.AVG, wc.MAX])
def test_sparse_keys_in_the_big_shared_table_tier(agg):
    """2 K - 4 K sparse keys (no usable range): one 16 K-slot shared-memory table per SM where the
    accumulators fit (SUM), the global hash table otherwise -- same groups either way."""
    n = 600_017
    rng = np.random.default_rng(13)
    pool = rng.integers(-2**31, 2**31, 3000, dtype=np.int64).astype(np.int32)
    t = {"price": orc.synth_f32(n, 98, -10.0, 100.0), "quantity": pool[rng.integers(0, len(pool), n)]}
    ref = orc.group_agg("price", "quantity", "price > 0", t, agg=agg)
    k, v = ops.group_agg(dev(t), "price[idx]", "quantity[idx]", cu("price > 0"), agg=agg, expected_groups=3000)
    assert np.array_equal(k.cpu().numpy(), ref["keys"])
    if agg == wc.MAX:
        assert np.array_equal(bits(v.cpu().numpy()), bits(ref["vals"]))
    else:
        np.testing.assert_allclose(v.cpu().numpy(), ref["vals"], rtol=SUM_RTOL, atol=0)


def test_direct_addressed_table_for_wide_integer_ranges():
    """Key ranges too wide for shared memory but known to the optimizer use a direct-addressed table
    (no probe, ordered export without a sort).  Covers: WHERE, negative keys, a range starting at
    INT_MIN, stale statistics (keys outside the promised range), chunks with different ranges (flush +
    re-prepare), merging partials afterwards (flush into the hash table), DESC export, -0.0 values."""
    wc.set_option("group.auto_stats_min_rows", 0)
    try:
        n = 600_011
        rng = np.random.default_rng(11)
        price = orc.synth_f32(n, 97, -5.0, 100.0)
        for keys in (rng.integers(-40_000, 60_000, n), -2**31 + rng.integers(0, 70_000, n), 2**31 - 1 - rng.integers(0, 70_000, n)):
            t = {"price": price, "quantity": np.ascontiguousarray(keys, dtype=np.int32)}
            d = dev(t)
            for agg, cond, order in ((wc.SUM, None, wc.ORDER_KEY_ASC), (wc.AVG, "price > 20", wc.ORDER_KEY_DESC), (wc.COUNT, "price > 95", wc.ORDER_KEY_ASC)):
                ref = orc.group_agg("price", "quantity", cond, t, agg=agg, order=order)
                k, v = ops.group_agg(d, "price[idx]", "quantity[idx]", cu(cond), agg=agg, order=order, expected_groups=100_000)
                assert np.array_equal(k.cpu().numpy(), ref["keys"])
                if agg == wc.COUNT:
                    assert np.array_equal(bits(v.cpu().numpy()), bits(ref["vals"]))
                else:
                    np.testing.assert_allclose(v.cpu().numpy(), ref["vals"], rtol=SUM_RTOL, atol=0)
        # stale statistics + chunks with different ranges + partial merge
        keys = rng.integers(0, 100_000, n).astype(np.int32)
        t = {"price": price, "quantity": keys}
        d = dev(t)
        ref = orc.group_agg("price", "quantity", None, t, agg=orc.AVG)
        tab = ops.AggTable(0, 100_000, wc.NEED_SUM | wc.NEED_COUNT)
        third = n // 3 // 8 * 8
        tab.set_key_range(0, 49_999)                                  # wrong: keys reach 99 999
        tab.consume({k: v[:third] for k, v in d.items()}, "price[idx]", "quantity[idx]")
        tab.set_key_range(10_000, 99_999)                             # not covered by the live range: flush, new side table
        tab.consume({k: v[third:2 * third] for k, v in d.items()}, "price[idx]", "quantity[idx]", row_base=third)
        part = ops.AggTable(0, 100_000, wc.NEED_SUM | wc.NEED_COUNT)
        part.consume({k: v[2 * third:] for k, v in d.items()}, "price[idx]", "quantity[idx]", row_base=2 * third)   # statistics gathered by the core
        tab.merge(part.export(wc.AVG, wc.ORDER_KEY_ASC))              # partials land in the hash table: the side table is flushed first
        part.close()
        out = tab.export(wc.AVG, wc.ORDER_KEY_ASC)
        tab.close()
        assert np.array_equal(out["keys"].cpu().numpy(), ref["keys"])
        assert np.array_equal(out["counts"].cpu().numpy(), ref["counts"])
        np.testing.assert_allclose(out["sums"].cpu().numpy(), ref["sums"], rtol=1e-12)
        # merging partials into a fresh table whose key range is known: direct-addressed merge
        halves = []
        for a, b in ((0, n // 2 // 8 * 8), (n // 2 // 8 * 8, n)):
            ptab = ops.AggTable(0, 100_000, wc.NEED_SUM | wc.NEED_COUNT)
            ptab.consume({k: v[a:b] for k, v in d.items()}, "price[idx]", "quantity[idx]", row_base=a)
            halves.append(ptab.export(wc.AVG, wc.ORDER_KEY_ASC))
            ptab.close()
        final = ops.AggTable(0, 100_000, wc.NEED_SUM | wc.NEED_COUNT)
        final.set_key_range(0, 99_999)
        for h in halves:
            final.merge(h)
        out = final.export(wc.AVG, wc.ORDER_KEY_DESC)
        final.close()
        assert np.array_equal(out["keys"].cpu().numpy(), ref["keys"][::-1])
        assert np.array_equal(out["counts"].cpu().numpy(), ref["counts"][::-1])
        np.testing.assert_allclose(out["sums"].cpu().numpy(), ref["sums"][::-1], rtol=1e-12)
        # a group whose values are all -0.0 must still exist (untouched slots are recognised by the -0.0 bit pattern)
        t2 = {"price": np.array([-0.0, -0.0, 1.5, -0.0], np.float32), "quantity": np.array([7, 7, 90_000, 5], np.int32)}
        tab = ops.AggTable(0, 100_000, wc.NEED_SUM)
        tab.set_key_range(0, 90_000)
        tab.consume(dev(t2), "price[idx]", "quantity[idx]")
        out = tab.export(wc.SUM, wc.ORDER_KEY_ASC)
        tab.close()
        assert out["keys"].cpu().tolist() == [5, 7, 90_000] and out["vals"].cpu().tolist() == [0.0, 0.0, 1.5]
    finally:
        wc.set_option("group.auto_stats_min_rows", None)


def test_table_overflow_is_reported_and_retried():
    n = 300_000
    t = {"price": orc.synth_f32(n, 41, 0.0, 1.0), "quantity": np.arange(n, dtype=np.int32)}
    tab = ops.AggTable(0, expected_groups=1000, needs=wc.NEED_SUM)
    tab.consume(dev(t), "price[idx]", "quantity[idx]")
    with pytest.raises(wc.WarpcoreError, match="table overflow"):
        tab.size()
    tab.close()
    k, v = ops.group_agg(dev(t), "price[idx]", "quantity[idx]", expected_groups=1000, cap=n)   # grows and reruns
    assert np.array_equal(k.cpu().numpy(), t["quantity"]) and np.array_equal(bits(v.cpu().numpy()), bits(t["price"]))


def test_chunked_consume_and_partial_merge_equal_one_shot():
    """query_multi_gpu_csv-style chunks and the multi-GPU partial-aggregate merge."""
    n = 1_200_007
    t = {"price": orc.synth_f32(n, 51, 0.0, 100.0), "quantity": orc.synth_i32(n, 52, 0, 5000)}
    ref = orc.group_agg("price", "quantity", None, t, agg=orc.AVG)
    d = dev(t)
    needs = wc.NEED_SUM | wc.NEED_COUNT | wc.NEED_MINMAX | wc.NEED_FIRST_ROW
    parts = []
    for s in range(4):                                     # four "ranks", each folding two chunks
        lo, hi = orc.shard_range(n, 4, s)
        mid = (lo + hi) // 2 // 8 * 8
        tab = ops.AggTable(0, 5000, needs)
        for a, b in ((lo, mid), (mid, hi)):
            tab.consume({k: v[a:b] for k, v in d.items()}, "price[idx]", "quantity[idx]", row_base=a)
        parts.append(tab.export(wc.AVG, wc.ORDER_KEY_ASC))
        tab.close()
    final = ops.AggTable(0, 5000, needs)
    for p in parts:
        final.merge(p)
    out = final.export(wc.AVG, wc.ORDER_KEY_ASC)
    assert np.array_equal(out["keys"].cpu().numpy(), ref["keys"])
    assert np.array_equal(out["counts"].cpu().numpy(), ref["counts"])
    np.testing.assert_allclose(out["sums"].cpu().numpy(), ref["sums"], rtol=1e-12)
    np.testing.assert_allclose(out["vals"].cpu().numpy(), ref["vals"], rtol=SUM_RTOL)
    first = orc.group_agg("price", "quantity", None, t, order=orc.ORDER_FIRST)
    outf = final.export(wc.SUM, wc.ORDER_FIRST)
    assert np.array_equal(outf["keys"].cpu().numpy(), first["keys"])
    mins = orc.group_agg("price", "quantity", None, t, agg=orc.MIN)
    assert np.array_equal(bits(final.export(wc.MIN)["vals"].cpu().numpy()), bits(mins["vals"]))


@pytest.mark.parametrize("n", [1, 4, 5, 100, 8191, 8192, 100003, 3_000_001])
@pytest.mark.parametrize("desc", [True, False])
def test_topk_small_matches_oracle(n, desc):
    t = {"price": orc.synth_f32(n, 61, 0.0, 1e6), "quantity": orc.synth_i32(n, 62, 0, 50)}
    d = dev(t)
    for k, off in [(5, 0), (1, 0), (16, 0), (2, 1), (5, 11)]:
        ref = orc.topk("price", None, t, descending=desc, k=k, offset=off)
        got = ops.topk(d, "price[idx]", None, None, desc, k, off)
        assert np.array_equal(bits(got.cpu().numpy()), bits(ref)), (n, desc, k, off)
    ref = orc.topk("discount(price, 0.9)", "quantity < 10", t, descending=desc, k=5)
    got = ops.topk(d, cu("discount(price, 0.9)"), None, cu("quantity < 10"), desc, 5)
    assert np.array_equal(bits(got.cpu().numpy()), bits(ref))


@pytest.mark.parametrize("desc", [True, False])
def test_topk_ties_are_stable_and_payload_follows_key(desc):
    """ORDER BY quantity (many ties) returning price: equal keys keep row order (bubble sort is stable)."""
    n = 200_003
    t = {"price": orc.synth_f32(n, 71, 0.0, 100.0), "quantity": orc.synth_i32(n, 72, 0, 7)}
    sql = f"SELECT price FROM t ORDER BY quantity {'DESC' if desc else 'ASC'} LIMIT 9"
    ref = orc.query_sql(sql, t)
    got = ops.topk(dev(t), "quantity[idx]", "price[idx]", None, desc, 9)
    assert np.array_equal(bits(got.cpu().numpy()), bits(ref))
    big = orc.query_sql(sql.replace("LIMIT 9", "LIMIT 5000"), t)           # large-k path (threshold + compaction + sort)
    got = ops.topk(dev(t), "quantity[idx]", "price[idx]", None, desc, 5000)
    assert np.array_equal(bits(got.cpu().numpy()), bits(big))


@pytest.mark.parametrize("k,off", [(17, 0), (100, 3), (5000, 0), (-1, 0), (-1, 10)])
def test_topk_large_and_full_sort(k, off):
    n = 1_000_003
    t = {"price": orc.synth_f32(n, 81, -50.0, 50.0), "quantity": orc.synth_i32(n, 82, 0, 50)}
    for desc in (True, False):
        kk = n if k < 0 else k
        ref = orc.topk("price * 2", "quantity > 4", t, descending=desc, k=kk, offset=off)
        got = ops.topk(dev(t), cu("price * 2"), None, cu("quantity > 4"), desc, k, off)
        assert np.array_equal(bits(got.cpu().numpy()), bits(ref)), (k, off, desc)


def test_device_sorts_match_reference_bubble_sort_semantics():
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 255, 256, 257, 4097, 300_001):
        f = rng.standard_normal(n).astype(np.float32)
        for asc in (True, False):
            got = ops.sort_float(torch.from_numpy(f.copy()).cuda(), asc).cpu().numpy()
            assert np.array_equal(bits(got), bits(orc.sort_float(f, asc)))
        keys = rng.integers(-5, 5, n).astype(np.int32)
        vals = np.arange(n, dtype=np.float32)
        for asc in (True, False):
            rk, rv = orc.sort_pairs(keys, vals, asc)
            gk, gv = ops.sort_pairs(torch.from_numpy(keys.copy()).cuda(), torch.from_numpy(vals.copy()).cuda(), asc)
            assert np.array_equal(gk.cpu().numpy(), rk) and np.array_equal(gv.cpu().numpy(), rv)
