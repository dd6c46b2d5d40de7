"""GPU: zone-map pruning returns exactly what the unpruned filter returns (oracle-checked), and
actually skips zones on clustered data."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import pyoracle as orc
from warpdb_b200 import _core as wc
from warpdb_b200 import ops


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module", autouse=True)
def gpu():
    assert torch.cuda.is_available()
    wc.check(wc.lib().wdb_init(0))


@pytest.mark.parametrize("layout", ["sorted", "clustered", "random"])
@pytest.mark.parametrize("n", [5000, 1_000_003])
def test_pruned_filter_equals_plain_filter(layout, n):
    price = orc.synth_f32(n, 91, 0.0, 100.0)
    qty = orc.synth_i32(n, 92, 0, 50)
    if layout == "sorted":
        price = np.sort(price)
    elif layout == "clustered":                      # runs of 10k rows drawn from narrow, shuffled bands
        band = (np.arange(n) // 10000).astype(np.int64)
        band = (band * 7919) % 97
        price = (band + price / 100.0).astype(np.float32)
    t = {"price": np.ascontiguousarray(price), "quantity": qty}
    d = {k: torch.from_numpy(v).cuda() for k, v in t.items()}
    zp, zq = ops.ZoneMap(d["price"], "price"), ops.ZoneMap(d["quantity"], "quantity")
    cases = [("price * 0.9", "price > 90", [(zp, ">", 90.0)]),
             ("price * quantity", "price >= 10 AND price < 12.5 AND quantity != 7", [(zp, ">=", 10.0), (zp, "<", 12.5), (zq, "!=", 7.0)]),
             ("price", "price == 50", [(zp, "==", 50.0)]),
             ("price", "price > 1000", [(zp, ">", 1000.0)]),
             ("price", "price <= 100 AND quantity >= 0", [(zp, "<=", 100.0), (zq, ">=", 0.0)])]
    for text, where, preds in cases:
        e, c = orc.Expr(text).cuda(), orc.Expr(where).cuda()
        ref, mask = orc.project_filter(text, where, t, fill=0.0)
        out, cnt, live = ops.project_filter_pruned(d, e, c, preds, wc.DENSE_ZERO)
        assert np.array_equal(bits(out.cpu().numpy()), bits(ref)), (layout, where)
        refc = orc.filter_compact(text, where, t)
        outc, cnt, livec = ops.project_filter_pruned(d, e, c, preds, wc.COMPACT)
        assert cnt == len(refc) and np.array_equal(bits(outc[:cnt].cpu().numpy()), bits(refc)), (layout, where)
        out_u = torch.full((n,), -3.0, device="cuda")
        ops.project_filter_pruned(d, e, c, preds, wc.DENSE, out=out_u)
        refu, _ = orc.project_filter(text, where, t, fill=-3.0)
        assert np.array_equal(bits(out_u.cpu().numpy()), bits(refu)), (layout, where)
        assert 0 <= live <= zp.nzones and live == livec
        if where == "price > 1000":
            assert live == 0 and cnt == 0
        if layout == "sorted" and where == "price > 90" and n > 100000:
            assert live < 0.15 * zp.nzones            # ~10 % of the rows qualify and they are contiguous
        if layout == "random" and where == "price > 90" and n > 100000:
            assert live == zp.nzones                  # nothing to prune on uniformly random data


def test_int_column_compared_as_float_is_never_wrongly_pruned():
    # 16777217 converts to the float 16777216: `q == 16777216` is TRUE in the kernel for that row
    q = np.full(8192, 16777217, np.int32)
    d = {"q": torch.from_numpy(q).cuda()}
    z = ops.ZoneMap(d["q"], "q")
    out, cnt, live = ops.project_filter_pruned(d, "q[idx]", "(q[idx] == 16777216.0f)", [(z, "==", 16777216.0)], wc.COMPACT)
    plain, cnt_plain = ops.project_filter(d, "q[idx]", "(q[idx] == 16777216.0f)", wc.COMPACT)
    assert cnt == cnt_plain == 8192 and live == z.nzones


@pytest.mark.parametrize("layout", ["sorted", "clustered", "random"])
def test_pruned_group_by_and_topk_equal_the_plain_calls(layout):
    """GROUP BY / ORDER BY ... LIMIT with a WHERE clause over a zone-mapped table run on the live runs only."""
    n = 1_000_003
    price = orc.synth_f32(n, 91, 0.0, 100.0)
    qty = orc.synth_i32(n, 92, 0, 50)
    if layout == "sorted":
        price = np.sort(price)
    elif layout == "clustered":
        band = ((np.arange(n) // 10000).astype(np.int64) * 7919) % 97
        price = (band + price / 100.0).astype(np.float32)
    t = {"price": np.ascontiguousarray(price), "quantity": qty}
    d = {k: torch.from_numpy(v).cuda() for k, v in t.items()}
    zp = ops.ZoneMap(d["price"], "price")
    for where, preds in [("price > 90", [(zp, ">", 90.0)]), ("price >= 10 AND price < 12.5", [(zp, ">=", 10.0), (zp, "<", 12.5)]),
                         ("price > 1000", [(zp, ">", 1000.0)])]:
        c = orc.Expr(where).cuda()
        for agg, needs in ((wc.SUM, wc.NEED_SUM), (wc.MIN, wc.NEED_MINMAX), (wc.COUNT, wc.NEED_COUNT)):
            r = orc.group_agg("price", "quantity", where, t, agg=agg)
            tab = ops.AggTable(0, 64, needs)
            live = tab.consume(d, "price[idx]", "quantity[idx]", c, preds=preds)
            out = tab.export(agg, wc.ORDER_KEY_ASC)
            tab.close()
            assert np.array_equal(out["keys"].cpu().numpy(), r["keys"]), (layout, where, agg)
            if agg == wc.SUM:
                np.testing.assert_allclose(out["vals"].cpu().numpy(), r["vals"], rtol=1e-6)
            else:
                assert np.array_equal(bits(out["vals"].cpu().numpy()), bits(r["vals"]))
            if layout == "sorted" and where == "price > 90":
                assert live < 0.15 * zp.nzones
        # first-appearance order needs global row ids: every run is consumed with its own row_base
        r = orc.group_agg("price", "quantity", where, t, agg=orc.SUM, order=orc.ORDER_FIRST)
        tab = ops.AggTable(0, 64, wc.NEED_SUM | wc.NEED_FIRST_ROW)
        tab.consume(d, "price[idx]", "quantity[idx]", c, preds=preds)
        out = tab.export(wc.SUM, wc.ORDER_FIRST)
        tab.close()
        assert np.array_equal(out["keys"].cpu().numpy(), r["keys"]), (layout, where)
        for desc in (True, False):
            for k, off in ((5, 0), (7, 3)):
                want = orc.query_sql(f"SELECT price * 2 FROM t WHERE {where} ORDER BY quantity {'DESC' if desc else 'ASC'} LIMIT {k} OFFSET {off}", t)
                got = ops.topk(d, "quantity[idx]", "(price[idx] * 2.0f)", c, desc, k, off, preds=preds).cpu().numpy()
                assert np.array_equal(bits(got), bits(want)), (layout, where, desc, k, off)


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("dtype", [np.float32, np.int32, np.float64])
def test_upload_column_builds_the_zone_map_behind_the_copies(pinned, dtype):
    """Resident ingest (wdb_upload_column, the replacement of upload_to_gpu's synchronous cudaMemcpy): chunked asynchronous
    copies from pinned memory or through the pinned staging ring; the zone map built chunk by chunk prunes exactly like one
    built over the resident column, and the exact min / max comes back."""
    n = 40_000_003 if dtype != np.float64 else 9_000_001          # several 64 MB chunks, ragged tail
    rng = np.random.default_rng(3)
    host = np.sort(rng.uniform(-1000, 1000, n)).astype(dtype) if dtype != np.int32 else np.sort(rng.integers(-10**6, 10**6, n)).astype(np.int32)
    src = torch.from_numpy(host)
    if pinned:
        src = src.pin_memory()
    col, zm, (lo, hi) = ops.upload_column(src, "price")
    torch.cuda.synchronize()
    assert torch.equal(col.cpu(), torch.from_numpy(host))
    assert lo == float(host.min()) and hi == float(host.max())
    ref = ops.ZoneMap(col, "price")
    assert (zm.zone_rows, zm.nzones) == (ref.zone_rows, ref.nzones)
    if dtype == np.float32:
        d = {"price": col}
        for zmap in (zm, ref):
            out, cnt, live = ops.project_filter_pruned(d, "price[idx]", "(price[idx] > 990.0f)", [(zmap, ">", 990.0)], wc.COMPACT)
            assert cnt == int((host > 990.0).sum()) and np.array_equal(out[:cnt].cpu().numpy(), host[host > 990.0])
            assert live < 0.02 * zmap.nzones
