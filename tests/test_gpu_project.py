"""GPU parity: fused filter/project/compaction through the C ABI vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import pyoracle as orc
from warpdb_b200 import _core as wc
from warpdb_b200 import ops

UDF = "__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n"


@pytest.fixture(scope="module", autouse=True)
def gpu():
    assert torch.cuda.is_available(), "these tests need the B200"
    wc.check(wc.lib().wdb_init(0))
    wc.set_udf_source(UDF)
    yield
    wc.set_udf_source("")


def dev(table):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in table.items()}


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def check_dense(table, text, where=None, fill=-123.0):
    e = orc.Expr(text).cuda()
    c = orc.Expr(where).cuda() if where else None
    ref, mask = orc.project_filter(text, where, table, fill=fill)
    d = dev(table)
    n = len(ref)
    out = torch.full((n,), fill, dtype=torch.float32, device="cuda")
    ops.project_filter(d, e, c, wc.DENSE, out=out)
    got = out.cpu().numpy()
    assert np.array_equal(bits(got), bits(ref)), (text, where)
    # zero-fill mode
    out0, _ = ops.project_filter(d, e, c, wc.DENSE_ZERO)
    ref0 = np.where(mask.astype(bool), ref, np.float32(0))
    assert np.array_equal(bits(out0.cpu().numpy()), bits(ref0)), (text, where)
    # stable compaction
    outc, cnt = ops.project_filter(d, e, c, wc.COMPACT)
    refc = orc.filter_compact(text, where, table)
    assert cnt == len(refc), (text, where, cnt, len(refc))
    assert np.array_equal(bits(outc[:cnt].cpu().numpy()), bits(refc)), (text, where)


def test_reference_fixture_kats(fixtures):
    # SURVEY Appendix B / tests/extended_types_test.cpp / tests/jit_arch_test.cpp
    t = fixtures["test"]
    out, _ = ops.project_filter(dev(t), "(price[idx] * quantity[idx])", "(price[idx] > 10.0f)")
    assert [format(int(x), "08x") for x in bits(out.cpu().numpy())] == ["41fc0000", "42a00000", "41f40000", "43160000"]
    out, _ = ops.project_filter(dev(t), "((price[idx] * quantity[idx]) * 1.08f)")
    assert [format(int(x), "08x") for x in bits(out.cpu().numpy())] == ["4208147b", "42accccd", "4203c290", "43220000"]
    out, _ = ops.project_filter(dev(fixtures["extended"]), "(price[idx] * discount[idx])")
    assert [format(int(x), "08x") for x in bits(out.cpu().numpy())] == ["3f866667", "40800000", "3f433333", "40900000"]
    one = {"price": np.array([2.0], np.float32), "quantity": np.array([0], np.int32)}
    out, _ = ops.project_filter(dev(one), "price[idx]")
    assert out.cpu().tolist() == [2.0]


def test_synth_generators_match_oracle():
    for n, row0 in [(1, 0), (1000, 0), (100003, 12345), (1 << 20, (1 << 33) + 7)]:
        a = ops.synth_f32(n, 0xC0FFEE, 0.0, 100.0, row0).cpu().numpy()
        assert np.array_equal(bits(a), bits(orc.synth_f32(n, 0xC0FFEE, 0.0, 100.0, row0)))
        b = ops.synth_i32(n, 0xC0FFEF, 1, 101, row0).cpu().numpy()
        assert np.array_equal(b, orc.synth_i32(n, 0xC0FFEF, 1, 101, row0))


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 31, 32, 33, 255, 1023, 4095, 4096, 4097, 8191, 8192, 8193, 100003, 1 << 20])
def test_ragged_sizes(n):
    t = {"price": orc.synth_f32(n, 1, 0.0, 40.0), "quantity": orc.synth_i32(n, 2, 1, 101)}
    if n == 0:
        d = {k: torch.empty(0, dtype=torch.float32 if k == "price" else torch.int32, device="cuda") for k in t}
        out, cnt = ops.project_filter(d, "(price[idx] * quantity[idx])", None)
        assert out.numel() == 0 and cnt == 0
        out, cnt = ops.project_filter(d, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT)
        assert cnt == 0
        return
    check_dense(t, "price * quantity * 1.08")
    check_dense(t, "price * 0.9", "price > 20")


@pytest.mark.parametrize("text,where", [
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
])
def test_expression_matrix(text, where):
    n = 70001
    t = {"price": orc.synth_f32(n, 3, 0.0, 40.0), "quantity": orc.synth_i32(n, 4, 1, 101),
         "quantity2": orc.synth_i32(n, 5, 1, 7), "d": orc.synth_f32(n, 6, 0.0, 1.0).astype(np.float64) / 3.0,
         "l": orc.synth_i32(n, 7, 0, 1000).astype(np.int64) * 1000003}
    check_dense(t, text, where)


@pytest.mark.parametrize("sel", [0.0, 0.01, 0.5, 0.99, 1.0])
def test_compaction_selectivities(sel):
    n = 3_000_017
    hi = 20.0 / (1.0 - sel) if sel < 1.0 else 10.0
    lo = 0.0 if sel > 0 else 0.0
    if sel == 0.0:
        hi = 20.0
    if sel == 1.0:
        lo, hi = 21.0, 40.0
    t = {"price": orc.synth_f32(n, 9, lo, hi)}
    ref = orc.filter_compact("price * 0.9", "price > 20", t)
    out, cnt = ops.project_filter(dev(t), "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT)
    assert cnt == len(ref)
    assert abs(cnt / n - sel) < 0.01
    assert np.array_equal(bits(out[:cnt].cpu().numpy()), bits(ref))


def test_misaligned_columns_take_the_scalar_path():
    n = 50001
    p = orc.synth_f32(n + 3, 10, 0.0, 40.0)
    dp = torch.from_numpy(p).cuda()
    for off in (1, 2, 3):
        t = {"price": p[off:off + n]}
        d = {"price": dp[off:off + n]}
        ref, _ = orc.project_filter("price * 0.9", "price > 20", t, fill=0.0)
        out, _ = ops.project_filter(d, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.DENSE_ZERO)
        assert np.array_equal(bits(out.cpu().numpy()), bits(ref))
        outc, cnt = ops.project_filter(d, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT)
        assert np.array_equal(bits(outc[:cnt].cpu().numpy()), bits(orc.filter_compact("price * 0.9", "price > 20", t)))


@pytest.mark.parametrize("variant,vec,unroll,block", [(0, 4, 1, 128), (0, 8, 4, 256), (1, 8, 2, 512), (1, 4, 8, 256), (2, 4, 4, 256)])
def test_kernel_variants_agree(variant, vec, unroll, block):
    n = 1_000_003
    t = {"price": orc.synth_f32(n, 11, 0.0, 100.0), "quantity": orc.synth_i32(n, 12, 1, 101)}
    ref, _ = orc.project_filter("price * quantity * 1.08", None, t)
    ref0, m0 = orc.project_filter("price * 0.9", "price > 20", t, fill=0.0)
    d = dev(t)
    opts = {"project.variant": variant, "project.vec": vec, "project.unroll": unroll, "project.block": block}
    try:
        for k, v in opts.items():
            wc.set_option(k, v)
        out, _ = ops.project_filter(d, "((price[idx] * quantity[idx]) * 1.08f)")
        assert np.array_equal(bits(out.cpu().numpy()), bits(ref))
        out, _ = ops.project_filter(d, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.DENSE_ZERO)
        assert np.array_equal(bits(out.cpu().numpy()), bits(ref0))
    finally:
        for k in opts:
            wc.set_option(k, None)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_compaction_variants_agree(variant):
    """ticket + register loads / TMA bulk ring / two-pass count-scan-scatter / L2-parked slabs: identical packed output."""
    n = 2_000_003
    t = {"price": orc.synth_f32(n, 13, 0.0, 40.0), "quantity": orc.synth_i32(n, 14, 1, 101)}
    d = dev(t)
    try:
        wc.set_option("compact.variant", variant)
        for text, where in [("price * 0.9", "price > 20"), ("price * quantity", "quantity < 3"), ("price", "price > 39.9"), ("price", "price >= 0")]:
            ref = orc.filter_compact(text, where, t)
            out, cnt = ops.project_filter(d, orc.Expr(text).cuda(), orc.Expr(where).cuda(), wc.COMPACT)
            assert cnt == len(ref) and np.array_equal(bits(out[:cnt].cpu().numpy()), bits(ref)), (variant, text, where)
    finally:
        wc.set_option("compact.variant", None)


@pytest.mark.parametrize("sel", [0.0, 0.004, 0.05, 0.11, 0.13, 0.5, 1.0])
def test_selective_filters_take_the_staged_kernels_and_stay_exact(sel):
    """Optimizer: on large tables every call samples the selectivity; the next call of the same query shape reads it and
    selective filters then run the staged two-pass kernels (variant 5).  (compact.auto = 2 decides on the device
    instead: both pipelines are launched and the one that is not needed returns at once.)  Their
    per-chunk slots overflow -- and are recomputed from the input -- when a chunk holds more than 128 survivors
    (sel ~ 0.11 .. 0.13 straddles that); forced on, they stay exact at any selectivity."""
    n = 5_000_011
    hi = 20.0 / (1.0 - sel) if sel < 1.0 else 20.0
    lo = 0.0 if sel < 1.0 else 21.0
    t = {"price": orc.synth_f32(n, 31, lo, max(hi, lo + 1.0))}
    d = dev(t)
    ref = orc.filter_compact("price * 0.9", "price > 20", t)
    try:
        wc.set_option("compact.auto_min_rows", 1 << 20)     # the optimizer's own choice (default: from 2^27 rows on)
        for mode in (1, 1, 1, 2):     # 1: feedback from the previous call of this query shape (default); 2: selected on the device
            wc.set_option("compact.auto", mode)
            out, cnt = ops.project_filter(d, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT)
            assert cnt == len(ref) and np.array_equal(bits(out[:cnt].cpu().numpy()), bits(ref)), (sel, mode)
        wc.set_option("compact.auto", None)
        dcnt = torch.zeros(1, dtype=torch.int64, device="cuda")   # the device-side count of the asynchronous form
        cols, nc = wc.make_cols(ops.schema_of(d))
        import ctypes as C
        wc.check(wc.lib().wdb_project_filter(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), cols, nc, b"(price[idx] * 0.9f)", b"(price[idx] > 20.0f)",
                                             out.data_ptr(), n, wc.COMPACT, dcnt.data_ptr(), None))
        torch.cuda.synchronize()
        assert int(dcnt.item()) == len(ref)
        wc.set_option("compact.variant", 5)
        out, cnt = ops.project_filter(d, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT)
        assert cnt == len(ref) and np.array_equal(bits(out[:cnt].cpu().numpy()), bits(ref)), sel
    finally:
        wc.set_option("compact.variant", None)
        wc.set_option("compact.auto_min_rows", None)
        wc.set_option("compact.auto", None)


def test_compile_error_and_recovery():
    # tests/jit_error_test.cpp:19-33
    d = dev({"price": np.array([1.0], np.float32), "quantity": np.array([1], np.int32)})
    with pytest.raises(wc.WarpcoreError, match=r"Kernel compilation failed\."):
        ops.project_filter(d, "invalid@")
    out, _ = ops.project_filter(d, "(price[idx] + 1.0f)")
    assert out.cpu().tolist() == [2.0]


def test_kernel_cache_hits():
    d = dev({"price": orc.synth_f32(1000, 1, 0.0, 1.0)})
    ops.project_filter(d, "(price[idx] * 3.25f)")
    s0 = wc.stats()
    ops.project_filter(d, "(price[idx] * 3.25f)")
    s1 = wc.stats()
    assert s1["kernels_compiled"] == s0["kernels_compiled"] and s1["cache_hits"] == s0["cache_hits"] + 1
    assert s1["launches"] > s0["launches"]


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("n", [1, 1023, 8192, 8193, 40_001, 1_000_001])
def test_no_writes_outside_the_output(variant, n):
    """compute-sanitizer is closed on this pool: guard regions around the output buffer instead.  The
    output slice starts 32-byte aligned inside a larger tensor filled with a sentinel; nothing outside
    [0, n) (dense) or [0, count) (compact) may change."""
    GUARD, S = 4096, -12345.0
    t = {"price": orc.synth_f32(n, 15, 0.0, 40.0), "quantity": orc.synth_i32(n, 16, 1, 101)}
    d = dev(t)
    try:
        wc.set_option("compact.variant", variant)
        for mode in (wc.DENSE, wc.DENSE_ZERO, wc.COMPACT):
            buf = torch.full((n + 2 * GUARD,), S, dtype=torch.float32, device="cuda")
            out = buf[GUARD:GUARD + n]
            _, cnt = ops.project_filter(d, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", mode, out=out)
            h = buf.cpu().numpy()
            assert (h[:GUARD] == S).all() and (h[GUARD + n:] == S).all(), (variant, n, mode)
            if mode == wc.COMPACT:
                assert (h[GUARD + cnt:GUARD + n] == S).all(), (variant, n, "compact tail")
            elif mode == wc.DENSE:
                keep = t["price"] > np.float32(20)
                assert (h[GUARD:GUARD + n][~keep] == S).all(), (variant, n, "untouched slots")
    finally:
        wc.set_option("compact.variant", None)
