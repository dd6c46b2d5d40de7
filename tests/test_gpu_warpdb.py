"""GPU: the drop-in surface -- pywarpdb (bindings/python/pywarpdb.cpp:7-38 of the reference, plus
query_sql) and the C++ seam with the reference's signatures, on the reference's bundled fixtures."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DATA = os.path.join(HERE, "data")


@pytest.fixture(scope="module")
def pw():
    from warpdb_b200 import build as wbuild
    wbuild.build_host()
    from warpdb_b200 import pywarpdb
    cwd = os.getcwd()
    os.chdir(DATA)          # ./custom.cu is read from the working directory (src/jit.cpp:65-73)
    yield pywarpdb
    os.chdir(cwd)


def test_python_smoke_like_reference(pw):
    # tests/test_python.py
    db = pw.WarpDB("test.csv")
    assert len(db.query("price + 1")) == 4


def test_pywarpdb_surface(pw):
    db = pw.WarpDB("test.csv", [pw.DataType.Float32, pw.DataType.Int32])
    assert db.query("price * quantity WHERE price > 10") == [31.5, 80.0, 30.5, 150.0]
    assert db.query("price * quantity * 1.08") == pytest.approx([34.02, 86.4, 32.94, 162.0], rel=1e-6)
    assert db.query_sql("SELECT SUM(price) FROM test GROUP BY quantity ORDER BY quantity ASC") == [15.25, 10.5, 20.0, 30.0]
    assert db.query_sql("SELECT price FROM test ORDER BY price DESC LIMIT 2") == [30.0, 20.0]
    assert db.query_sql("SELECT AVG(price) FROM test GROUP BY quantity ORDER BY quantity DESC") == [30.0, 20.0, 10.5, 15.25]
    assert db.query_sql("SELECT COUNT(price) FROM test WHERE price > 11 GROUP BY quantity") == [1.0, 1.0, 1.0]
    assert db.query_multi_gpu("price * quantity") == [31.5, 80.0, 30.5, 150.0]
    assert pw.WarpDB.query_multi_gpu_csv("test.csv", "price * 2 WHERE price > 15", 2) == [0.0, 40.0, 30.5, 60.0]
    arr, schema = db.query_arrow("price + 1")
    pa = pytest.importorskip("pyarrow")
    import ctypes
    ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
    ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    a = pa.Array._import_from_c(ctypes.pythonapi.PyCapsule_GetPointer(arr, b"arrow_array"),
                                ctypes.pythonapi.PyCapsule_GetPointer(schema, b"arrow_schema"))
    assert a.to_pylist() == [11.5, 21.0, 16.25, 31.0]
    with pytest.raises(RuntimeError, match="Unknown column: nope"):
        db.query("nope * 2")
    with pytest.raises(RuntimeError, match="Failed to parse SQL"):
        db.query_sql("SELECT price")
    with pytest.raises(RuntimeError, match="Only aggregation queries supported with GROUP BY"):
        db.query_sql("SELECT price FROM test GROUP BY quantity")
    with pytest.raises(RuntimeError, match=r"Kernel compilation failed\."):
        db.query("nosuchfunction(price)")


def test_query_sql_matches_oracle_on_fixture(pw, fixtures):
    from oracle import pyoracle as orc
    db = pw.WarpDB("test.csv", [pw.DataType.Float32, pw.DataType.Int32])
    t = fixtures["test"]
    for sql in [
        "SELECT SUM(price) FROM t GROUP BY quantity",
        "SELECT MAX(price * 2) FROM t GROUP BY quantity ORDER BY quantity DESC",
        "SELECT MIN(price) FROM t WHERE quantity > 2 GROUP BY quantity",
        "SELECT price FROM t ORDER BY price DESC LIMIT 5",
        "SELECT price FROM t ORDER BY quantity ASC LIMIT 3",
        "SELECT price * 0.9 FROM t WHERE price > 20",
        "SELECT price FROM t WHERE price > 11 LIMIT 2 OFFSET 1",
        "SELECT DISTINCT quantity FROM t ORDER BY quantity DESC",
        "SELECT discount(price, 0.9) FROM t ORDER BY discount(price, 0.9) DESC LIMIT 5",
        "SELECT SUM(price) FROM t GROUP BY quantity HAVING SUM(price) > 15 ORDER BY quantity ASC",
    ]:
        assert np.array_equal(np.array(db.query_sql(sql), np.float32), orc.query_sql(sql, t)), sql


def test_cpp_seam_with_reference_signatures(pw):
    """Compiles tests/cpp/jit_shim_test.cpp against the host headers and runs it in tests/data."""
    from warpdb_b200 import build as wbuild
    exe = os.path.join(ROOT, "warpdb_b200", "csrc", "build", "jit_shim_test")
    pkg = os.path.join(ROOT, "warpdb_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(HERE, "cpp", "jit_shim_test.cpp"),
           "-I", os.path.join(pkg, "csrc", "host"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(wbuild.CUDA, "include"),
           "-L", pkg, "-lwarpdb_host", "-lwarpcore", "-L", os.path.join(wbuild.CUDA, "lib64"), "-lcudart_static", "-ldl", "-lrt",
           "-lpthread", "-Wl,-rpath," + pkg]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    r = subprocess.run([exe], cwd=DATA, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL HOST-MIRROR TESTS PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "[Optimizer] Filter eliminates all rows." in r.stdout


def test_cli(pw):
    exe = os.path.join(ROOT, "warpdb_b200", "warpdb")
    r = subprocess.run([exe, "price * quantity WHERE price > 10", "test.csv"], cwd=DATA, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "JIT Result[3] = 150" in r.stdout, r.stdout + r.stderr


def test_group_by_several_keys(pw, tmp_path):
    """GROUP BY a, b: the reference parses the key list (include/expression.hpp:128-130) but reads keys[0] only
    (src/warpdb.cpp:362,374); here integer key columns fold into one composite key and the groups come back in
    lexicographic (a, b) order.  Checked against a NumPy group-by (this is beyond what the reference, and hence
    the oracle, computes)."""
    rng = np.random.default_rng(5)
    n = 20_000
    price = rng.uniform(0, 100, n).astype(np.float32)
    a = rng.integers(-3, 9, n).astype(np.int32)
    b = rng.integers(100, 140, n).astype(np.int32)
    path = tmp_path / "multi.csv"
    with open(path, "w") as f:
        f.write("price,a,b\n")
        for p, x, y in zip(price, a, b):
            f.write(f"{float(p)!r},{x},{y}\n")
    db = pw.WarpDB(str(path), [pw.DataType.Float32, pw.DataType.Int32, pw.DataType.Int32])
    comp = (a.astype(np.int64) + 3) * 40 + (b - 100)
    for where, mask in (("", np.ones(n, bool)), (" WHERE price > 50", price > 50)):
        keys = np.unique(comp[mask])
        sums = np.array([price[mask & (comp == k)].astype(np.float64).sum() for k in keys])
        cnts = np.array([(mask & (comp == k)).sum() for k in keys], np.float64)
        mx = np.array([price[mask & (comp == k)].max() for k in keys], np.float32)
        got = np.array(db.query_sql(f"SELECT SUM(price) FROM t{where} GROUP BY a, b"), np.float32)
        assert got.shape == sums.shape
        np.testing.assert_allclose(got, sums.astype(np.float32), rtol=1e-6)
        assert np.array_equal(np.array(db.query_sql(f"SELECT COUNT(price) FROM t{where} GROUP BY a, b"), np.float32), cnts.astype(np.float32))
        assert np.array_equal(np.array(db.query_sql(f"SELECT MAX(price) FROM t{where} GROUP BY a, b ORDER BY a DESC"), np.float32), mx[::-1])
    with pytest.raises(RuntimeError, match="integer columns"):
        db.query_sql("SELECT SUM(price) FROM t GROUP BY a, price")


def test_multi_gpu_sql_surface(pw):
    """query_sql_multi_gpu: aggregates / ORDER BY ... LIMIT over row-range shards of the host table, merged inside the core."""
    db = pw.WarpDB("test.csv", [pw.DataType.Float32, pw.DataType.Int32])
    assert db.query_sql_multi_gpu("SELECT SUM(price) FROM test GROUP BY quantity ORDER BY quantity ASC") == [15.25, 10.5, 20.0, 30.0]
    assert db.query_sql_multi_gpu("SELECT AVG(price) FROM test GROUP BY quantity ORDER BY quantity DESC") == [30.0, 20.0, 10.5, 15.25]
    assert db.query_sql_multi_gpu("SELECT price FROM test ORDER BY price DESC LIMIT 2") == [30.0, 20.0]
    assert db.query_sql_multi_gpu("SELECT price * 2 FROM test WHERE price > 11") == [40.0, 30.5, 60.0]
    with pytest.raises(RuntimeError, match="needs a LIMIT"):
        db.query_sql_multi_gpu("SELECT price FROM test ORDER BY price DESC")
