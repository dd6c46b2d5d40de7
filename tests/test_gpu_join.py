"""GPU: inner equi-join (wdb_join_build / wdb_join_probe / wdb_gather) against the oracle's
nested-loop definition (orc_join_pairs), and `JOIN ... ON a = b` through WarpDB.query_sql against
the oracle's query_sql run on the oracle-joined columns.  The reference parses JOIN and never executes
it (src/expression.cpp:375-401, include/warpdb.hpp:22), so the oracle states SQL's definition."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import pyoracle as orc
from warpdb_b200 import _core as wc
from warpdb_b200 import ops

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "data")


@pytest.fixture(scope="module", autouse=True)
def gpu():
    assert torch.cuda.is_available()
    wc.check(wc.lib().wdb_init(0))


def check_pairs(probe, build):
    want_p, want_b = orc.join_pairs(probe, build, indexed=True)
    idx = ops.JoinIndex(torch.from_numpy(build).cuda())
    dp = torch.from_numpy(probe).cuda()
    assert idx.count(dp) == len(want_p)
    got_p, got_b = idx.probe(dp)
    assert np.array_equal(got_p.cpu().numpy(), want_p)
    assert np.array_equal(got_b.cpu().numpy(), want_b)
    if idx.direct_span:                      # the same index probed by binary search instead of the key-range table
        wc.set_option("join.direct", 0)
        try:
            alt_p, alt_b = idx.probe(dp)
        finally:
            wc.set_option("join.direct", None)
        assert torch.equal(alt_p, got_p) and torch.equal(alt_b, got_b)
    idx.close()
    return len(want_p)


@pytest.mark.parametrize("pdt,bdt", [(np.int32, np.int32), (np.int64, np.int64), (np.int32, np.int64), (np.int64, np.int32)])
@pytest.mark.parametrize("n,m,span", [(1, 1, 1), (1000, 300, 50), (5000, 4097, 100000), (1_000_003, 100_000, 100_000), (300_000, 5, 3)])
def test_pairs_equal_nested_loop(pdt, bdt, n, m, span):
    rng = np.random.default_rng(n * 31 + m)
    probe = rng.integers(-span // 2, span - span // 2, n).astype(pdt)
    build = rng.integers(-span // 2, span - span // 2, m).astype(bdt)
    check_pairs(probe, build)


def test_definition_is_the_nested_loop():
    # the indexed oracle used above against the O(n*m) loop it abbreviates
    rng = np.random.default_rng(7)
    probe, build = rng.integers(-9, 30, 700), rng.integers(-9, 30, 450)
    a, b = orc.join_pairs(probe, build, indexed=False), orc.join_pairs(probe, build, indexed=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert check_pairs(probe.astype(np.int32), build.astype(np.int32)) == len(a[0])


def test_edges():
    e32 = np.empty(0, np.int32)
    some = np.arange(10, dtype=np.int32)
    assert check_pairs(e32, some) == 0                       # empty probe side
    assert check_pairs(some, e32) == 0                       # empty build side
    assert check_pairs(some, some + 100) == 0                # no key in common
    assert check_pairs(np.full(3000, 7, np.int32), np.full(2000, 7, np.int32)) == 6_000_000   # one key: full cross product
    lim32 = np.array([np.iinfo(np.int32).min, -1, 0, 1, np.iinfo(np.int32).max], np.int32)
    assert check_pairs(lim32, lim32[::-1].copy()) == 5       # sign handling of the sort keys
    lim64 = np.array([np.iinfo(np.int64).min, -(1 << 40), -1, 0, 1 << 33, np.iinfo(np.int64).max], np.int64)
    assert check_pairs(lim64, np.concatenate([lim64, lim64])) == 12
    wide = np.array([1 << 32, (1 << 32) + 5, 5], np.int64)   # an int64 probe value must not alias an int32 build key
    assert check_pairs(wide, np.array([0, 5], np.int32)) == 1
    dup = np.array([3, 1, 3, 3, 2, 1], np.int32)             # equal build keys come back in build-row order
    p, b = orc.join_pairs(np.array([3, 1], np.int32), dup)
    assert b.tolist() == [0, 2, 3, 1, 5]
    check_pairs(np.array([3, 1], np.int32), dup)
    import ctypes as C
    f = torch.zeros(4, device="cuda")
    h = C.c_void_p()
    with pytest.raises(wc.WarpcoreError, match="JOIN needs integer key columns"):
        wc.check(wc.lib().wdb_join_build(0, None, wc.make_cols([("f", wc.FLOAT32, f.data_ptr(), 4)])[0], C.byref(h)))


def test_dense_key_ranges_are_tabulated():
    ids = torch.randperm(100_000, dtype=torch.int32).cuda()
    assert ops.JoinIndex(ids).direct_span == 100_000                       # ids of a dimension table
    assert ops.JoinIndex(ids.long() * 1000).direct_span == 0               # sparse: 1000 key values per row
    assert ops.JoinIndex(torch.tensor([5, 40_000], dtype=torch.int32).cuda()).direct_span == 39_996   # small tables: up to 2^16 values
    assert ops.JoinIndex(torch.tensor([-(1 << 62), 1 << 62], dtype=torch.int64).cuda()).direct_span == 0


def test_capacity_is_checked():
    import ctypes as C
    build = torch.arange(100, dtype=torch.int32, device="cuda")
    probe = torch.arange(100, dtype=torch.int32, device="cuda")
    idx = ops.JoinIndex(build)
    cols, _ = wc.make_cols([("probe", wc.INT32, probe.data_ptr(), 100)])
    out = torch.full((200,), -1, dtype=torch.int64, device="cuda")
    n = C.c_int64(0)
    rc = wc.lib().wdb_join_probe(idx.handle, None, cols, out.data_ptr(), out[100:].data_ptr(), 50, C.byref(n))
    assert rc != 0 and b"exceed the output capacity" in wc.lib().wdb_last_error()
    assert n.value == 100 and bool((out == -1).all())
    idx.close()


@pytest.mark.parametrize("dtype", [torch.float32, torch.int32, torch.float64, torch.int64])
def test_gather(dtype):
    g = torch.Generator().manual_seed(3)
    src = (torch.rand(100_003, generator=g) * 1000).to(dtype).cuda()
    rows = torch.randint(0, src.shape[0], (250_001,), generator=g).cuda()
    assert torch.equal(ops.gather(src, rows), src[rows])
    assert torch.equal(ops.gather(src, None), src)
    assert ops.gather(src, rows[:0]).shape[0] == 0


# ---- JOIN through the SQL front end ------------------------------------------------------------
@pytest.fixture(scope="module")
def pw():
    from warpdb_b200 import build as wbuild
    wbuild.build_host()
    from warpdb_b200 import pywarpdb
    cwd = os.getcwd()
    os.chdir(DATA)
    yield pywarpdb
    os.chdir(cwd)


def write_csv(path, cols):
    names = list(cols)
    with open(path, "w") as f:
        f.write(",".join(names) + "\n")
        for i in range(len(cols[names[0]])):
            f.write(",".join(repr(float(cols[c][i])) if cols[c].dtype.kind == "f" else str(int(cols[c][i])) for c in names) + "\n")


@pytest.fixture(scope="module")
def star(pw, tmp_path_factory):
    """sales(item, price, quantity) x items(id, rate, cat) x cats(cid, boost): fact table with dangling and repeated keys"""
    d = tmp_path_factory.mktemp("join")
    rng = np.random.default_rng(11)
    n, m = 20_000, 300
    sales = {"item": rng.integers(0, m + 40, n).astype(np.int32),             # ids >= m have no match
             "price": (rng.random(n) * 100).astype(np.float32),
             "quantity": rng.integers(1, 9, n).astype(np.int32)}
    ids = rng.permutation(m).astype(np.int32)
    ids[:10] = ids[10:20]                                                      # ten repeated ids: rows match twice
    items = {"id": ids, "rate": (rng.random(m)).astype(np.float32), "cat": rng.integers(0, 12, m).astype(np.int32)}
    cats = {"cid": np.arange(10, dtype=np.int32), "boost": (1 + rng.random(10)).astype(np.float32)}   # cats 10, 11 dangle
    for name, t in (("sales", sales), ("items", items), ("cats", cats)):
        write_csv(str(d / f"{name}.csv"), t)
    F, I = pw.DataType.Float32, pw.DataType.Int32
    db = pw.WarpDB(str(d / "sales.csv"), [I, F, I])
    db.attach("items", str(d / "items.csv"), [I, F, I])
    db.attach("cats", str(d / "cats.csv"), [I, F])
    return db, sales, items, cats


def test_sql_join_matches_oracle(star):
    db, sales, items, cats = star
    p, b = orc.join_pairs(sales["item"], items["id"])
    j = {"item": sales["item"][p], "price": sales["price"][p], "quantity": sales["quantity"][p],
         "id": items["id"][b], "rate": items["rate"][b], "cat": items["cat"][b]}
    cases = [
        ("SELECT price * rate FROM sales JOIN items ON sales.item = items.id", "SELECT price * rate FROM j"),
        ("SELECT sales.price * items.rate FROM sales JOIN items ON item = id WHERE items.rate > 0.5 AND sales.quantity < 5",
         "SELECT price * rate FROM j WHERE rate > 0.5 AND quantity < 5"),
        ("SELECT price FROM sales JOIN items ON items.id == sales.item WHERE cat == 3", "SELECT price FROM j WHERE cat == 3"),
        ("SELECT SUM(price * rate) FROM sales JOIN items ON sales.item = items.id GROUP BY cat", "SELECT SUM(price * rate) FROM j GROUP BY cat"),
        ("SELECT AVG(price) FROM sales JOIN items ON sales.item = items.id WHERE quantity > 2 GROUP BY cat ORDER BY cat DESC",
         "SELECT AVG(price) FROM j WHERE quantity > 2 GROUP BY cat ORDER BY cat DESC"),
        ("SELECT COUNT(price) FROM sales JOIN items ON sales.item = items.id GROUP BY cat, quantity", None),
        ("SELECT price * rate FROM sales JOIN items ON sales.item = items.id ORDER BY price * rate DESC LIMIT 7",
         "SELECT price * rate FROM j ORDER BY price * rate DESC LIMIT 7"),
        ("SELECT DISTINCT cat FROM sales JOIN items ON sales.item = items.id ORDER BY cat ASC", "SELECT DISTINCT cat FROM j ORDER BY cat ASC"),
        ("SELECT price FROM sales JOIN items ON sales.item = items.id WHERE price > 99 LIMIT 5 OFFSET 2", "SELECT price FROM j WHERE price > 99 LIMIT 5 OFFSET 2"),
    ]
    for sql, flat in cases:
        got = np.array(db.query_sql(sql), np.float32)
        assert db.last_join_rows() == len(p)
        if flat is None:                              # two GROUP BY keys: lexicographic (cat, quantity) order, counted with numpy
            key = j["cat"].astype(np.int64) * 100 + j["quantity"]
            _, cnt = np.unique(key, return_counts=True)
            assert np.array_equal(got, cnt.astype(np.float32)), sql
        else:
            want = orc.query_sql(flat, j)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), sql
    assert db.query_sql("SELECT price FROM sales ORDER BY price DESC LIMIT 1") and db.last_join_rows() == -1


def test_sql_three_way_join_and_self_join(star):
    db, sales, items, cats = star
    p, b = orc.join_pairs(sales["item"], items["id"])
    p2, c = orc.join_pairs(items["cat"][b], cats["cid"])
    j = {"price": sales["price"][p][p2], "rate": items["rate"][b][p2], "boost": cats["boost"][c], "cid": cats["cid"][c],
         "quantity": sales["quantity"][p][p2]}
    sql = "SELECT price * rate * boost FROM sales JOIN items ON sales.item = items.id JOIN cats ON items.cat = cats.cid WHERE quantity > 1"
    got = np.array(db.query_sql(sql), np.float32)
    want = orc.query_sql("SELECT price * rate * boost FROM j WHERE quantity > 1", j)
    assert db.last_join_rows() == len(p2) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    got = np.array(db.query_sql("SELECT MAX(price * boost) FROM sales JOIN items ON sales.item = items.id JOIN cats ON items.cat = cats.cid GROUP BY cid"), np.float32)
    assert np.array_equal(got, orc.query_sql("SELECT MAX(price * boost) FROM j GROUP BY cid", j))
    # a table that was not attached: the reference's "JOIN loads the same table" (include/warpdb.hpp:22) -- a self-join
    ps, bs = orc.join_pairs(sales["item"], sales["item"])
    got = np.array(db.query_sql("SELECT sales.price - other.price FROM sales JOIN other ON sales.item = other.item WHERE sales.quantity > other.quantity"), np.float32)
    jj = {"lp": sales["price"][ps], "rp": sales["price"][bs], "lq": sales["quantity"][ps], "rq": sales["quantity"][bs]}
    want = orc.query_sql("SELECT lp - rp FROM j WHERE lq > rq", jj)
    assert db.last_join_rows() == len(ps) and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_sql_join_errors(star):
    db = star[0]
    with pytest.raises(RuntimeError, match="JOIN condition: Unknown column: items.nope"):
        db.query_sql("SELECT price FROM sales JOIN items ON sales.item = items.nope")
    with pytest.raises(RuntimeError, match="SELECT clause: Unknown column: nope"):
        db.query_sql("SELECT nope FROM sales JOIN items ON sales.item = items.id")
    with pytest.raises(RuntimeError, match="only `<column> = <column>` is supported"):
        db.query_sql("SELECT price FROM sales JOIN items ON sales.item > items.id")
    with pytest.raises(RuntimeError, match="key columns must be Int32 or Int64"):
        db.query_sql("SELECT price FROM sales JOIN items ON sales.price = items.id")
    with pytest.raises(RuntimeError, match="must compare a column of items with a column of the tables before it"):
        db.query_sql("SELECT price FROM sales JOIN items ON items.id = items.cat")
