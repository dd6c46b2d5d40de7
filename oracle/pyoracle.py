"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under warpdb_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

INT32, INT64, FLOAT32, FLOAT64, STRING = 0, 1, 2, 3, 4
SUM, AVG, COUNT, MIN, MAX = 0, 1, 2, 3, 4
ORDER_FIRST, ORDER_KEY_ASC, ORDER_KEY_DESC = 0, 1, 2

_NP2DT = {np.dtype(np.int32): INT32, np.dtype(np.int64): INT64,
          np.dtype(np.float32): FLOAT32, np.dtype(np.float64): FLOAT64}


class OracleError(RuntimeError):
    pass


class _Col(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int), ("data", C.c_void_p), ("len", C.c_int64)]


def build(force=False):
    src = os.path.join(HERE, "wdb_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_parse_expression.restype = C.c_void_p
        L.orc_parse_expression.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_to_cuda_expr.restype = C.c_size_t
        L.orc_to_cuda_expr.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.orc_parse_query.restype = C.c_void_p
        L.orc_parse_query.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.orc_free_query.argtypes = [C.c_void_p]
        L.orc_query_summary.restype = C.c_size_t
        L.orc_query_summary.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.orc_tokenize_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_validate.argtypes = [C.c_void_p, C.POINTER(_Col), C.c_int, C.c_char_p, C.c_size_t]
        L.orc_project_filter.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_Col), C.c_int, C.c_int64, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
        L.orc_filter_compact.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_Col), C.c_int, C.c_int64, C.c_int64,
                                         C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_char_p, C.c_size_t]
        L.orc_group_agg.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(_Col), C.c_int,
                                    C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_int64), C.c_int, C.c_char_p, C.c_size_t]
        L.orc_topk.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_Col), C.c_int, C.c_int64, C.c_int64, C.c_int,
                               C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_char_p, C.c_size_t]
        L.orc_sort_float.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.orc_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.orc_shard_range.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_query.argtypes = [C.c_char_p, C.POINTER(_Col), C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_int, C.c_char_p, C.c_size_t]
        L.orc_query_sql.argtypes = [C.c_char_p, C.POINTER(_Col), C.c_int, C.c_int64, C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_int64), C.c_int, C.c_char_p, C.c_size_t]
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_synth_f32.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_float, C.c_float, C.c_int64]
        L.orc_synth_i32.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_int32, C.c_int64]
        _lib = L
    return _lib


def _err():
    return C.create_string_buffer(1024)


def _cols(table):
    """table: dict name -> 1-D contiguous numpy array (int32/int64/float32/float64)."""
    arr = (_Col * max(len(table), 1))()
    keep = []
    for i, (name, a) in enumerate(table.items()):
        a = np.ascontiguousarray(a)
        keep.append(a)
        arr[i].name = name.encode()
        arr[i].dtype = _NP2DT[a.dtype]
        arr[i].data = a.ctypes.data
        arr[i].len = a.shape[0]
    return arr, len(table), keep


class Expr:
    """A parsed expression (reference grammar)."""

    def __init__(self, text):
        e = _err()
        self.ptr = lib().orc_parse_expression(text.encode(), e, len(e))
        if not self.ptr:
            raise OracleError(e.value.decode())

    def cuda(self):
        n = lib().orc_to_cuda_expr(self.ptr, None, 0)
        b = C.create_string_buffer(n + 1)
        lib().orc_to_cuda_expr(self.ptr, b, n + 1)
        return b.value.decode()

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.orc_free(self.ptr)
            self.ptr = None


def _p(e):
    return e.ptr if e is not None else None


def _as_expr(x):
    if x is None or isinstance(x, Expr):
        return x
    if x == "":
        return None
    return Expr(x)


def tokenize_dump(text):
    e = _err()
    out = C.create_string_buffer(1 << 16)
    if lib().orc_tokenize_dump(text.encode(), out, len(out), e, len(e)):
        raise OracleError(e.value.decode())
    return out.value.decode()


def query_summary(sql, ext=False):
    e = _err()
    q = lib().orc_parse_query(sql.encode(), int(ext), e, len(e))
    if not q:
        raise OracleError(e.value.decode())
    try:
        n = lib().orc_query_summary(q, None, 0)
        b = C.create_string_buffer(n + 1)
        lib().orc_query_summary(q, b, n + 1)
        return b.value.decode()
    finally:
        lib().orc_free_query(q)


def project_filter(expr, cond, table, row0=0, row1=None, fill=0.0, contract=True, nthreads=1):
    """Dense filter/project.  Returns (out float32[n], mask uint8[n]); slots failing cond keep `fill`."""
    expr, cond = _as_expr(expr), _as_expr(cond)
    cols, nc, keep = _cols(table)
    n = len(next(iter(table.values()))) if row1 is None else row1
    out = np.full(n - row0, fill, dtype=np.float32)
    mask = np.zeros(n - row0, dtype=np.uint8)
    e = _err()
    if lib().orc_project_filter(_p(expr), _p(cond), cols, nc, row0, n, out.ctypes.data, mask.ctypes.data,
                                int(contract), nthreads, e, len(e)):
        raise OracleError(e.value.decode())
    return out, mask


def filter_compact(expr, cond, table, row0=0, row1=None, contract=True):
    expr, cond = _as_expr(expr), _as_expr(cond)
    cols, nc, keep = _cols(table)
    n = len(next(iter(table.values()))) if row1 is None else row1
    out = np.empty(max(n - row0, 1), dtype=np.float32)
    cnt = C.c_int64(0)
    e = _err()
    if lib().orc_filter_compact(_p(expr), _p(cond), cols, nc, row0, n, out.ctypes.data, C.byref(cnt), int(contract),
                                e, len(e)):
        raise OracleError(e.value.decode())
    return out[:cnt.value].copy()


def group_agg(val, key, cond, table, agg=SUM, order=ORDER_KEY_ASC, row0=0, row1=None, contract=True):
    """Returns dict(keys int32[G], vals float32[G], sums float64[G], counts int64[G])."""
    val, key, cond = _as_expr(val), _as_expr(key), _as_expr(cond)
    cols, nc, keep = _cols(table)
    n = len(next(iter(table.values()))) if row1 is None else row1
    cap = max(n - row0, 1)
    keys = np.empty(cap, np.int32)
    vals = np.empty(cap, np.float32)
    sums = np.empty(cap, np.float64)
    cnts = np.empty(cap, np.int64)
    g = C.c_int64(0)
    e = _err()
    if lib().orc_group_agg(_p(val), _p(key), _p(cond), agg, order, cols, nc, row0, n, keys.ctypes.data,
                           vals.ctypes.data, sums.ctypes.data, cnts.ctypes.data, cap, C.byref(g), int(contract),
                           e, len(e)):
        raise OracleError(e.value.decode())
    G = g.value
    return dict(keys=keys[:G].copy(), vals=vals[:G].copy(), sums=sums[:G].copy(), counts=cnts[:G].copy())


def topk(expr, cond, table, descending=True, k=5, offset=0, row0=0, row1=None, contract=True):
    expr, cond = _as_expr(expr), _as_expr(cond)
    cols, nc, keep = _cols(table)
    n = len(next(iter(table.values()))) if row1 is None else row1
    out = np.empty(max(k, 1), np.float32)
    m = C.c_int64(0)
    e = _err()
    if lib().orc_topk(_p(expr), _p(cond), cols, nc, row0, n, int(descending), k, offset, out.ctypes.data,
                      C.byref(m), int(contract), e, len(e)):
        raise OracleError(e.value.decode())
    return out[:m.value].copy()


def sort_float(vals, ascending=True):
    a = np.ascontiguousarray(vals, dtype=np.float32).copy()
    lib().orc_sort_float(a.ctypes.data, a.shape[0], int(ascending))
    return a


def sort_pairs(keys, vals, ascending=True):
    k = np.ascontiguousarray(keys, dtype=np.int32).copy()
    v = np.ascontiguousarray(vals, dtype=np.float32).copy()
    lib().orc_sort_pairs(k.ctypes.data, v.ctypes.data, k.shape[0], int(ascending))
    return k, v


def join_pairs(probe, build, indexed=True):
    """every (i, j) with probe[i] == build[j], ordered by i then j -> (probe_rows, build_rows) int64"""
    p = np.ascontiguousarray(probe, dtype=np.int64)
    b = np.ascontiguousarray(build, dtype=np.int64)
    L = lib()
    L.orc_join_pairs.restype = C.c_int64
    L.orc_join_pairs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    total = L.orc_join_pairs(p.ctypes.data, p.shape[0], b.ctypes.data, b.shape[0], int(indexed), None, None, 0)
    pr = np.empty(total, dtype=np.int64)
    br = np.empty(total, dtype=np.int64)
    got = L.orc_join_pairs(p.ctypes.data, p.shape[0], b.ctypes.data, b.shape[0], int(indexed), pr.ctypes.data, br.ctypes.data, total)
    assert got == total
    return pr, br


def shard_range(n, ndev, dev):
    s, e = C.c_int64(0), C.c_int64(0)
    lib().orc_shard_range(n, ndev, dev, C.byref(s), C.byref(e))
    return s.value, e.value


def query(q, table, fill=0.0, contract=True, nthreads=1):
    """WarpDB::query: returns (out float32[num_rows], mask)."""
    cols, nc, keep = _cols(table)
    n = len(next(iter(table.values())))
    out = np.full(n, fill, dtype=np.float32)
    mask = np.zeros(n, dtype=np.uint8)
    e = _err()
    if lib().orc_query(q.encode(), cols, nc, n, out.ctypes.data, mask.ctypes.data, int(contract), nthreads, e, len(e)):
        raise OracleError(e.value.decode())
    return out, mask


def query_sql(sql, table, contract=True):
    cols, nc, keep = _cols(table)
    n = len(next(iter(table.values())))
    out = np.empty(max(n, 1), dtype=np.float32)
    m = C.c_int64(0)
    e = _err()
    if lib().orc_query_sql(sql.encode(), cols, nc, n, out.ctypes.data, out.shape[0], C.byref(m), int(contract),
                           e, len(e)):
        raise OracleError(e.value.decode())
    return out[:m.value].copy()


def synth_f32(n, seed, lo, hi, row0=0):
    out = np.empty(n, np.float32)
    lib().orc_synth_f32(out.ctypes.data, n, seed, lo, hi, row0)
    return out


def synth_i32(n, seed, lo, hi_excl, row0=0):
    out = np.empty(n, np.int32)
    lib().orc_synth_i32(out.ctypes.data, n, seed, lo, hi_excl, row0)
    return out
