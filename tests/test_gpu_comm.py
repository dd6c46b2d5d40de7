"""GPU: the sharded entry points of the core (wdb_comm_* / wdb_multi_*) on ONE rank -- the same code
path the N-GPU runs take, minus the collectives -- plus what they lean on: extrema in the
direct-addressed table, the export without a host synchronisation and the NaN policy of ORDER BY.
The >= 2-GPU runs of the same entry points live in tests/_nccl_worker.py (tests/test_gpu_multi.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import pyoracle as orc
from warpdb_b200 import _core as wc
from warpdb_b200 import ops


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def comm():
    assert torch.cuda.is_available(), "these tests need the B200"
    wc.check(wc.lib().wdb_init(0))
    c = ops.Comm(0, 0, 1)
    yield c
    c.close()


def make(n, lo=-50, hi=1950):
    t = {"price": orc.synth_f32(n, 77, 0.0, 40.0), "quantity": orc.synth_i32(n, 78, lo, hi)}
    return t, {k: torch.from_numpy(v).cuda() for k, v in t.items()}


@pytest.mark.parametrize("agg", [wc.SUM, wc.AVG, wc.COUNT, wc.MIN, wc.MAX])
@pytest.mark.parametrize("span", [7, 1000, 20000, 300000])
def test_group_agg_through_the_comm_entry_point(comm, agg, span):
    t, d = make(400_003, -span // 2, span - span // 2)
    for order in (wc.ORDER_KEY_ASC, wc.ORDER_KEY_DESC):
        r = orc.group_agg("price", "quantity", "price > 5", t, agg=agg, order=order)
        keys, vals = comm.group_agg(d, "price[idx]", "quantity[idx]", "(price[idx] > 5.0f)", agg, order)
        assert np.array_equal(keys.cpu().numpy(), r["keys"]), (agg, span, order)
        if agg in (wc.COUNT, wc.MIN, wc.MAX):
            assert np.array_equal(bits(vals.cpu().numpy()), bits(r["vals"]))
        else:
            np.testing.assert_allclose(vals.cpu().numpy(), r["vals"], rtol=1e-6)


def test_group_agg_async_leaves_the_count_on_the_device(comm):
    t, d = make(200_001)
    r = orc.group_agg("price", "quantity", None, t, agg=orc.SUM)
    keys, vals, groups = comm.group_agg(d, "price[idx]", "quantity[idx]", None, wc.SUM, wc.ORDER_KEY_ASC, key_range=(-50, 1949), sync=False)
    torch.cuda.synchronize()
    g = int(groups.item())
    assert g == len(r["keys"]) and np.array_equal(keys[:g].cpu().numpy(), r["keys"])
    np.testing.assert_allclose(vals[:g].cpu().numpy(), r["vals"], rtol=1e-6)


def test_stale_key_statistics_are_reported(comm):
    _, d = make(100_000)
    with pytest.raises(wc.WarpcoreError, match="stale"):
        comm.group_agg(d, "price[idx]", "quantity[idx]", None, wc.SUM, wc.ORDER_KEY_ASC, key_range=(0, 100))


def test_first_appearance_order_and_empty_result(comm):
    t, d = make(50_000, 0, 300)
    r = orc.group_agg("price", "quantity", None, t, agg=orc.SUM, order=orc.ORDER_FIRST)
    keys, vals = comm.group_agg(d, "price[idx]", "quantity[idx]", None, wc.SUM, wc.ORDER_FIRST, expected_groups=300)
    assert np.array_equal(keys.cpu().numpy(), r["keys"])
    np.testing.assert_allclose(vals.cpu().numpy(), r["vals"], rtol=1e-6)
    keys, vals = comm.group_agg(d, "price[idx]", "quantity[idx]", "(price[idx] > 1000.0f)", wc.SUM, wc.ORDER_KEY_ASC)
    assert keys.numel() == 0


@pytest.mark.parametrize("k,offset", [(5, 0), (7, 2), (16, 0), (40, 3), (300, 11)])
def test_topk_through_the_comm_entry_point(comm, k, offset):
    t, d = make(300_007)
    for desc in (True, False):
        want = orc.query_sql(f"SELECT price FROM t WHERE price > 20 ORDER BY quantity {'DESC' if desc else 'ASC'} LIMIT {k} OFFSET {offset}", t)
        got = comm.topk(d, "quantity[idx]", "price[idx]", "(price[idx] > 20.0f)", desc, k, offset).cpu().numpy()
        assert np.array_equal(bits(got), bits(want)), (k, offset, desc)


def test_compaction_counts_through_the_comm_entry_point(comm):
    t, d = make(123_457)
    refc = orc.filter_compact("price * 0.9", "price > 20", t)
    out, (cnt, off, total) = comm.project_filter(d, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT)
    assert (cnt, off, total) == (len(refc), 0, len(refc)) and np.array_equal(bits(out[:cnt].cpu().numpy()), bits(refc))


def test_dense_table_holds_extrema(comm):
    """MIN / MAX over a wide key range used to fall to the hash table; they now live in the direct-addressed table."""
    n = 1_000_003
    t = {"price": orc.synth_f32(n, 5, -100.0, 100.0), "quantity": orc.synth_i32(n, 6, 0, 200_000)}
    d = {k: torch.from_numpy(v).cuda() for k, v in t.items()}
    for agg in (wc.MIN, wc.MAX):
        r = orc.group_agg("price", "quantity", None, t, agg=agg)
        tab = ops.AggTable(0, 200_000, wc.NEED_MINMAX)
        tab.set_key_range(0, 199_999)
        tab.consume(d, "price[idx]", "quantity[idx]")
        out = tab.export(agg, wc.ORDER_KEY_ASC)
        assert np.array_equal(out["keys"].cpu().numpy(), r["keys"]) and np.array_equal(bits(out["vals"].cpu().numpy()), bits(r["vals"]))
        tab.close()


def test_order_by_nan_keys_sort_last_in_every_path():
    """One NaN policy for LIMIT <= 16 (register top-k) and beyond (threshold compaction + radix sort): rows
    whose key is NaN come after every number in both directions and are not dropped."""
    n = 100_000
    key = orc.synth_f32(n, 9, -10.0, 10.0)
    key[::1000] = np.nan
    val = np.arange(n, dtype=np.float32)
    d = {"k": torch.from_numpy(key).cuda(), "v": torch.from_numpy(val).cuda()}
    for desc in (True, False):
        num = np.flatnonzero(~np.isnan(key))
        order = num[np.argsort(-key[num] if desc else key[num], kind="stable")]
        full = np.concatenate([order, np.flatnonzero(np.isnan(key))])
        for k in (16, 17, 2000):
            got = ops.topk(d, "k[idx]", "v[idx]", None, desc, k).cpu().numpy()
            assert np.array_equal(got, val[full[:k]]), (desc, k)
        # all survivors come back, NaN keys at the end
        got = ops.topk(d, "k[idx]", "v[idx]", None, desc, -1).cpu().numpy()
        assert np.array_equal(got, val[full])
    few = {"k": torch.tensor([float("nan"), 1.0, float("nan"), 3.0], device="cuda"), "v": torch.tensor([0.0, 1.0, 2.0, 3.0], device="cuda")}
    assert ops.topk(few, "k[idx]", "v[idx]", None, True, 4).cpu().tolist() == [3.0, 1.0, 0.0, 2.0]
    assert ops.topk(few, "k[idx]", "v[idx]", None, False, 4).cpu().tolist() == [1.0, 3.0, 0.0, 2.0]
