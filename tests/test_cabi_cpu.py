"""CPU: the C-ABI library loads, exports every symbol include/warpcore.h declares, generates and
compiles its kernel templates for sm_100a offline (NVRTC needs no GPU), and fails loudly when asked
to compute without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

from warpdb_b200 import _core as wc
from warpdb_b200 import build as wbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCHEMA = [("price", wc.FLOAT32, 0, 0), ("quantity", wc.INT32, 0, 0)]


@pytest.fixture(scope="module", autouse=True)
def built():
    wbuild.build()


def sass_of(cubin, tmp_path, name="k.cubin"):
    p = tmp_path / name
    p.write_bytes(cubin)
    return subprocess.run(["cuobjdump", "-sass", str(p)], capture_output=True, text=True, check=True).stdout


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "warpcore.h")).read()
    declared = set(re.findall(r"\b(wdb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(wc.SYMBOLS), declared ^ set(wc.SYMBOLS)
    L = wc.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.wdb_abi_version() == 2   # 2: wdb_comm_* / wdb_multi_* (merges inside the core)


def test_shard_range_matches_reference_formula():
    # src/multi_gpu_utils.cpp:24-31
    assert [wc.shard_range(10, 4, d) for d in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [wc.shard_range(4, 8, d) for d in range(8)] == [(0, 1), (1, 2), (2, 3), (3, 4)] + [(4, 4)] * 4


def test_project_kernel_is_vectorised_sm100(tmp_path):
    src, cubin = wc.debug_compile("project", SCHEMA, "((price[idx] * quantity[idx]) * 1.08f)")
    assert "wdb_cell<wdb_t0> price, wdb_cell<wdb_t1> quantity" in src
    sass = sass_of(cubin, tmp_path)
    assert "sm_100a" in sass or "SM100" in sass.upper()
    # 256-bit global accesses (new on sm_100), output stored with the L2 evict-first hint
    assert re.search(r"LDG\.E\.NA\.\w+\.256\.CONSTANT", sass) and re.search(r"STG\.E\.NA\.EFL2\.256", sass)
    assert "FMUL" in sass


def test_only_referenced_columns_are_loaded(tmp_path):
    src, _ = wc.debug_compile("project", SCHEMA, "(price[idx] * 0.9f)", want_cubin=False)
    assert "#define WDB_NUSED 1" in src and "quantity" not in src.split("---- end UDF ----")[1]


def test_compact_kernels(tmp_path):
    def sass_for(variant, name):
        wc.set_option("compact.variant", variant)
        _, cubin = wc.debug_compile("compact", SCHEMA, "(price[idx] * 0.9f)", None, "(price[idx] > 20.0f)", wc.COMPACT)
        return sass_of(cubin, tmp_path, name)
    try:
        # default: one HBM pass, slab parked in L2 (count -> one look-back per slab -> re-read, rank, write)
        sass = sass_for(3, "k3.cubin")
        assert "wdb_compact_l2" in sass and "VOTE" in sass and "POPC" in sass and "LDG.E.64.STRONG.GPU" in sass
        assert re.search(r"LDG\.E\.NA\.\w+\.256", sass)
        assert "LDG.E.NA.ELL2.256" in sass and "LDG.E.NA.EFL2.256" in sass   # phase 1 parks the slab in the L2, phase 2 releases it
        assert "SHFL.UP" in sass                                               # survivors ranked with a warp scan of per-lane counts
        # two streaming passes (count -> scan -> scatter): no atomics, no block barriers
        sass = sass_for(2, "k2.cubin")
        assert "wdb_count" in sass and "wdb_scatter" in sass and "SHFL.UP" in sass and "POPC" in sass
        assert "ATOMG" not in sass and "BAR.SYNC" not in sass
        # single pass, TMA bulk-copy ring (UBLKCP + mbarrier), status-word look-back
        sass = sass_for(1, "k1.cubin")
        assert "wdb_compact_bulk" in sass and "UBLKCP" in sass and "SYNCS" in sass and "LDG.E.64.STRONG.GPU" in sass
        # single pass, register-staged vector loads, atomic tile ticket
        sass = sass_for(0, "k0.cubin")
        assert "VOTE" in sass and "POPC" in sass and "ATOMG" in sass and "LDG.E.64.STRONG.GPU" in sass
    finally:
        wc.set_option("compact.variant", None)


def test_group_kernels_by_accumulator_layout(tmp_path):
    """GROUP BY layouts: the narrow-range kernel updates warp-private accumulators with
    plain LDS/STS -- no shared-memory atomic anywhere in it; the wide-range layout is one native RED
    per row into the direct-addressed table; the general kernel uses shared CAS + global RED."""
    def fn_sass(sass, name):
        part = sass.split("Function : " + name + "\n")[1]
        return part.split("Function : ")[0]
    try:
        wc.set_option("group.debug_span", 1000)          # key range [0, 1000) known
        src, cubin = wc.debug_compile("group", SCHEMA, "price[idx]", "quantity[idx]", None, wc.SUM)
        wp = fn_sass(sass_of(cubin, tmp_path, "g_wp.cubin"), "wdb_group_wp")
        # the only ATOMS left are the shared-memory branch of the generic fp64 atomicAdd on the GLOBAL tables
        # (fold / spill path, once per key and CTA); the per-row path is LDS / DADD / STS
        assert wp.count("ATOMS") <= 4 and "ATOMS.ADD" not in wp and wp.count("LDS.64") > 16 and wp.count("STS.64") > 16 and "DADD" in wp
        assert re.search(r"LDG\.E\.NA\.\w+\.256", wp)
        assert re.search(r"(ATOMG?|REDG?)\.E\.ADD\.F64", wp)   # the per-CTA totals go to the direct-addressed side table
        wc.set_option("group.debug_span", 10_000_000)    # wide range: direct-addressed table, no wdb_group_wp
        src, cubin = wc.debug_compile("group", SCHEMA, "price[idx]", "quantity[idx]", None, wc.SUM)
        sass = sass_of(cubin, tmp_path, "g_dense.cubin")
        assert "wdb_group_wp" not in sass and "#define WDB_DENSE 1" in src
        g = fn_sass(sass, "wdb_group")
        # one native RED per row; the columns arrive through 128-bit non-allocating loads (small CTAs and
        # narrower loads keep fewer REDs in flight per SM: profiles/r02_sweep_group10m.jsonl)
        assert "REDG.E.ADD.F64" in g and re.search(r"LDG\.E\.NA\.128", g)
        wc.set_option("group.debug_span", 0)             # unknown range: shared-memory table with CAS, global RED behind it
        src, cubin = wc.debug_compile("group", SCHEMA, "price[idx]", "quantity[idx]", None, wc.SUM)
        g = fn_sass(sass_of(cubin, tmp_path, "g_hash.cubin"), "wdb_group")
        assert "ATOMS.CAS" in g and "#define WDB_DENSE 0" in src
    finally:
        wc.set_option("group.debug_span", None)


def test_keyrange_kernel_evaluates_the_key_expression(tmp_path):
    """Optimizer statistics on demand: min/max of the GROUP BY key expression (kernels/keyrange.cuh)."""
    src, cubin = wc.debug_compile("keyrange", SCHEMA, "(quantity[idx] / 3.0f)")
    assert "wdb_fn_key" in src and "(int)((quantity[idx] / 3.0f))" in src.replace("  ", " ")
    sass = sass_of(cubin, tmp_path, "kr.cubin")
    assert "wdb_keyrange" in sass and "REDUX.MIN.S32" in sass and "REDG.E.MAX.S32" in sass and re.search(r"LDG\.E\.NA\.\w+\.256", sass)


def test_bulk_variant_emits_tma_bulk_copies(tmp_path):
    wc.set_option("project.variant", 2)
    try:
        _, cubin = wc.debug_compile("project", SCHEMA, "((price[idx] * quantity[idx]) * 1.08f)")
    finally:
        wc.set_option("project.variant", None)
    sass = sass_of(cubin, tmp_path)
    assert "UBLKCP" in sass and "SYNCS" in sass


def test_udf_source_is_prepended(tmp_path):
    wc.set_udf_source("__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n")
    try:
        src, cubin = wc.debug_compile("project", SCHEMA, "discount(price[idx], 0.9f)")
    finally:
        wc.set_udf_source("")
    assert "__device__ float discount" in src and len(cubin) > 1000


def test_compile_error_contract(capfd):
    # src/jit.cpp:119-129 and tests/jit_error_test.cpp: log on stderr, "Kernel compilation failed."
    with pytest.raises(wc.WarpcoreError, match=r"Kernel compilation failed\."):
        wc.debug_compile("project", SCHEMA, "invalid@")
    assert "NVRTC Compile Log" in capfd.readouterr().err
    # and the library still works afterwards
    _, cubin = wc.debug_compile("project", SCHEMA, "(price[idx] + 1.0f)")
    assert len(cubin) > 1000


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(wc.WarpcoreError, match="no CUDA device|CUDA error"):
        wc.check(wc.lib().wdb_init(0))
    cols, n = wc.make_cols(SCHEMA)
    cnt = C.c_int64(0)
    rc = wc.lib().wdb_project_filter(0, None, cols, n, b"(price[idx] + 1.0f)", b"", None, 4, 0, None, C.byref(cnt))
    assert rc != 0
