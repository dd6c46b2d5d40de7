"""GPU parity against the REFERENCE'S OWN GPU functions (not the oracle): oracle/_ref/libref_jit.so is
/root/reference/src/jit.cpp:48-307 compiled from the reference's sources (oracle/Makefile) behind a C
shim (oracle/ref_jit_shim.cpp).  The same device arrays go through the reference's NVRTC path and
through the product's C ABI:

  jit_compile_and_launch (include/jit.hpp:7-10)   vs wdb_project_filter(WDB_DENSE)    bit-exact, untouched slots included
  jit_sort_float / jit_sort_pairs (:22-27)         vs wdb_sort_float / wdb_sort_pairs  bit-exact (stable, ties included)
  jit_group_sum (:15-18)                           vs wdb_group_agg(WDB_ORDER_FIRST)   keys and their order bit-exact; sums within the
                                                                                       bound of fp32 sequential summation (the reference adds
                                                                                       float32 one row at a time, the product accumulates in fp64)
The oracle is checked against the same reference outputs, so the three agree pairwise.
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import pyoracle as orc
from warpdb_b200 import _core as wc
from warpdb_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_jit.so")
UDF = "__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n"
DT = {np.dtype(np.int32): wc.INT32, np.dtype(np.int64): wc.INT64, np.dtype(np.float32): wc.FLOAT32, np.dtype(np.float64): wc.FLOAT64}


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libref_jit.so was not built (needs /root/reference at build time)")
    assert torch.cuda.is_available(), "these tests need the B200"
    wc.check(wc.lib().wdb_init(0))
    wc.set_udf_source(UDF)
    lib = C.CDLL(REF_LIB)
    lib.ref_set_primary_ctx(1)   # the reference's per-call cuCtxCreate cannot see runtime allocations cheaply
    # the reference reads ./custom.cu on every call (src/jit.cpp:65-73)
    d = tmp_path_factory.mktemp("refcwd")
    (d / "custom.cu").write_text(UDF)
    old = os.getcwd()
    os.chdir(d)
    yield lib
    os.chdir(old)
    wc.set_udf_source("")


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def ref_project(lib, table_dev, expr, cond, out):
    names = list(table_dev)
    n = out.numel()
    c_names = (C.c_char_p * len(names))(*[s.encode() for s in names])
    c_dts = (C.c_int * len(names))(*[{torch.int32: 0, torch.int64: 1, torch.float32: 2, torch.float64: 3}[table_dev[k].dtype] for k in names])
    c_ptrs = (C.c_void_p * len(names))(*[table_dev[k].data_ptr() for k in names])
    err = C.create_string_buffer(1024)
    rc = lib.ref_jit_compile_and_launch(expr.encode(), (cond or "").encode(), c_names, c_dts, c_ptrs, len(names), n, C.c_void_p(out.data_ptr()), 0, err, 1024)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()


MATRIX = [
    ("price * quantity * 1.08", None),
    ("price * 0.9", "price > 20"),
    ("discount(price, 0.9)", "price > 20 AND quantity < 50"),
    ("price + quantity * 2", "price > 10 OR quantity < 5"),
    ("(price + quantity) * 2", "quantity <= 5"),
    ("price * price + 1", None),                 # fma contraction
    ("1 - price * quantity", "price != quantity"),
    ("price * quantity - quantity * 3", None),   # both operands are products
    ("price / quantity - 2", "quantity >= 2"),
    ("quantity / quantity2", None),              # int / int
    ("quantity * quantity2 + quantity", None),   # int arithmetic
    ("d * 2 + price", "d > 0.5"),                # double column
    ("l + quantity", "l > 100"),                 # int64 column
    ("sqrtf(price) + fminf(price, 3)", "fmaxf(price, 20) > 20"),
    ("price", "price > 100"),                    # nothing survives
    ("price", "price > 0 OR price == 0"),        # everything survives
    ("7", None),
    ("price > 20", None),                        # boolean valued expression
]


@pytest.mark.parametrize("text,where", MATRIX)
def test_project_filter_matches_the_reference_kernel(ref, text, where):
    n = 70001
    t = {"price": orc.synth_f32(n, 3, 0.0, 40.0), "quantity": orc.synth_i32(n, 4, 1, 101),
         "quantity2": orc.synth_i32(n, 5, 1, 7), "d": orc.synth_f32(n, 6, 0.0, 1.0).astype(np.float64) / 3.0,
         "l": orc.synth_i32(n, 7, 0, 1000).astype(np.int64) * 1000003}
    d = {k: torch.from_numpy(v).cuda() for k, v in t.items()}
    e, c = orc.Expr(text).cuda(), (orc.Expr(where).cuda() if where else None)
    fill = -123.0
    want = torch.full((n,), fill, dtype=torch.float32, device="cuda")
    ref_project(ref, d, e, c, want)
    got = torch.full((n,), fill, dtype=torch.float32, device="cuda")
    ops.project_filter(d, e, c, wc.DENSE, out=got)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32)), (text, where)
    # ... and the oracle agrees with the reference's GPU path too (what every other parity test leans on)
    o, _ = orc.project_filter(text, where, t, fill=fill)
    assert np.array_equal(bits(o), bits(want.cpu().numpy())), (text, where)


@pytest.mark.parametrize("cfg", ["projection", "filter1", "filter50", "filter99"])
def test_baseline_configs_at_2p28_rows_match_the_reference_kernel(ref, cfg):
    n = 1 << 28
    if cfg == "projection":
        d = {"price": ops.synth_f32(n, 0xC0FFEE + 2, 0.0, 100.0), "quantity": ops.synth_i32(n, 0xC0FFEE + 102, 1, 101)}
        e, c = "((price[idx] * quantity[idx]) * 1.08f)", None
    else:
        sel = {"filter1": 0.01, "filter50": 0.5, "filter99": 0.99}[cfg]
        d = {"price": ops.synth_f32(n, 0xC0FFEE + 3, 0.0, 20.0 / (1.0 - sel))}
        e, c = "(price[idx] * 0.9f)", "(price[idx] > 20.0f)"
    want = torch.full((n,), -1.0, dtype=torch.float32, device="cuda")
    ref_project(ref, d, e, c, want)
    got = torch.full((n,), -1.0, dtype=torch.float32, device="cuda")
    ops.project_filter(d, e, c, wc.DENSE, out=got)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))
    if c:   # the compacted output is the reference's dense output with the untouched slots squeezed out
        mask = d["price"] > 20.0
        outc, cnt = ops.project_filter(d, e, c, wc.COMPACT)
        assert cnt == int(mask.sum().item())
        assert torch.equal(outc[:cnt].view(torch.int32), want[mask].view(torch.int32))


@pytest.mark.parametrize("n", [1, 2, 33, 1000, 4097])
@pytest.mark.parametrize("asc", [1, 0])
def test_sorts_match_the_reference_bubble_sorts(ref, n, asc):
    err = C.create_string_buffer(512)
    vals = orc.synth_f32(n, 11, -50.0, 50.0)
    vals[::7] = vals[0]                     # ties
    a = torch.from_numpy(vals.copy()).cuda()
    b = a.clone()
    assert ref.ref_jit_sort_float(C.c_void_p(a.data_ptr()), n, asc, 0, err, 512) == 0, err.value.decode()
    torch.cuda.synchronize()
    ops.sort_float(b, ascending=bool(asc))
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    # pairs: many equal keys; payloads must keep their relative order (the bubble sort swaps only strictly out-of-order neighbours)
    keys = orc.synth_i32(n, 12, -5, 6)
    pay = np.arange(n, dtype=np.float32)
    k1, p1 = torch.from_numpy(keys.copy()).cuda(), torch.from_numpy(pay.copy()).cuda()
    k2, p2 = k1.clone(), p1.clone()
    assert ref.ref_jit_sort_pairs(C.c_void_p(k1.data_ptr()), C.c_void_p(p1.data_ptr()), n, asc, 0, err, 512) == 0, err.value.decode()
    torch.cuda.synchronize()
    ops.sort_pairs(k2, p2, ascending=bool(asc))
    assert torch.equal(k1, k2) and torch.equal(p1, p2)


@pytest.mark.parametrize("n,groups", [(4, 3), (1000, 7), (100_000, 64), (30_000, 1500)])
def test_group_sum_matches_the_reference_group_kernel(ref, n, groups):
    price = orc.synth_f32(n, 21, 0.0, 100.0)
    qty = orc.synth_i32(n, 22, -3, groups - 3)
    dp, dq = torch.from_numpy(price).cuda(), torch.from_numpy(qty).cuda()
    out_v = torch.zeros(n, dtype=torch.float32, device="cuda")
    out_k = torch.zeros(n, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    err = C.create_string_buffer(512)
    rc = ref.ref_jit_group_sum(b"(price[idx] * 0.5f)", b"quantity[idx]", C.c_void_p(dp.data_ptr()), C.c_void_p(dq.data_ptr()), C.c_void_p(out_v.data_ptr()),
                               C.c_void_p(out_k.data_ptr()), C.c_void_p(cnt.data_ptr()), n, 0, err, 512)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    g = int(cnt.item())
    rk, rv = out_k[:g].cpu().numpy(), out_v[:g].cpu().numpy()
    keys, vals = ops.group_agg({"price": dp, "quantity": dq}, "(price[idx] * 0.5f)", "quantity[idx]", None, wc.SUM, wc.ORDER_FIRST, expected_groups=groups)
    keys, vals = keys.cpu().numpy(), vals.cpu().numpy()
    assert np.array_equal(keys, rk)                       # same groups in the same (first-appearance) order
    # fp32 recursive summation of m terms: |error| <= (m - 1) * 2^-24 * sum|v_i| (Higham, Accuracy and Stability, eq. 4.4, first order)
    half = (price * np.float32(0.5)).astype(np.float64)
    for i, k in enumerate(rk):
        sel = qty == k
        m, tot = int(sel.sum()), float(np.abs(half[sel]).sum())
        assert abs(float(vals[i]) - half[sel].sum()) <= 1e-6 * tot + 1e-30          # the product: fp64 accumulation, 1e-6 relative (north-star)
        assert abs(float(rv[i]) - float(vals[i])) <= (m + 1) * 2.0 ** -24 * tot + 2.0 ** -24 * tot, (k, m)
    # the oracle's first-appearance GROUP BY agrees on keys and order as well
    o = orc.group_agg("price * 0.5", "quantity", None, {"price": price, "quantity": qty}, agg=orc.SUM, order=orc.ORDER_FIRST)
    assert np.array_equal(o["keys"], rk)
