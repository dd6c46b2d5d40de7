"""ctypes binding of libwarpcore.so (C ABI: include/warpcore.h).

This is plumbing: device memory and streams come from PyTorch (tensor.data_ptr(),
torch.cuda.current_stream().cuda_stream), the work is done by the CUDA kernels of the library.
There is no CPU fallback: if the library is missing or no CUDA device is present every compute
call raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwarpcore.so")

INT32, INT64, FLOAT32, FLOAT64, STRING = 0, 1, 2, 3, 4
SUM, AVG, COUNT, MIN, MAX = 0, 1, 2, 3, 4
DENSE, COMPACT, DENSE_ZERO = 0, 1, 2
ORDER_FIRST, ORDER_KEY_ASC, ORDER_KEY_DESC = 0, 1, 2
NEED_SUM, NEED_COUNT, NEED_MINMAX, NEED_FIRST_ROW = 1, 2, 4, 8

# every symbol include/warpcore.h declares
SYMBOLS = [
    "wdb_abi_version", "wdb_last_error", "wdb_init", "wdb_shutdown", "wdb_device_count", "wdb_set_udf_source",
    "wdb_set_option", "wdb_get_option", "wdb_get_stats", "wdb_project_filter", "wdb_agg_create", "wdb_agg_destroy",
    "wdb_agg_reset", "wdb_agg_set_key_range", "wdb_agg_consume", "wdb_agg_merge", "wdb_agg_size", "wdb_agg_spilled", "wdb_agg_export", "wdb_group_agg",
    "wdb_topk", "wdb_sort_float", "wdb_sort_pairs", "wdb_column_minmax", "wdb_multi_project_filter_host",
    "wdb_zonemap_build", "wdb_upload_column", "wdb_zonemap_destroy", "wdb_zonemap_info", "wdb_project_filter_pruned", "wdb_agg_consume_pruned", "wdb_topk_pruned",
    "wdb_shard_range", "wdb_synth_f32", "wdb_synth_i32", "wdb_debug_compile", "wdb_free",
    "wdb_comm_unique_id", "wdb_comm_init_rank", "wdb_comm_init_all", "wdb_comm_destroy", "wdb_comm_info",
    "wdb_multi_project_filter", "wdb_multi_group_agg", "wdb_multi_topk", "wdb_multi_group_agg_host", "wdb_multi_topk_host",
    "wdb_join_build", "wdb_join_probe", "wdb_join_info", "wdb_join_destroy", "wdb_gather",
]


class WarpcoreError(RuntimeError):
    pass


class Col(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int), ("dptr", C.c_void_p), ("len", C.c_int64)]


class Prune(C.Structure):
    _fields_ = [("zonemap", C.c_void_p), ("op", C.c_int), ("value", C.c_double)]


PRUNE_OPS = {">": 0, ">=": 1, "<": 2, "<=": 3, "==": 4, "!=": 5}


class Stats(C.Structure):
    _fields_ = [("kernels_compiled", C.c_int64), ("cache_hits", C.c_int64), ("launches", C.c_int64),
                ("last_compile_ms", C.c_double)]


_lib = None


def lib():
    """Load libwarpcore.so (built in-tree by warpdb_b200.build).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WarpcoreError(f"{LIB_PATH} is missing: run `python -m warpdb_b200.build` (there is no fallback path)")
    L = C.CDLL(LIB_PATH)
    vp, i64, ci, cp = C.c_void_p, C.c_int64, C.c_int, C.c_char_p
    PC = C.POINTER(Col)
    P64 = C.POINTER(C.c_int64)
    L.wdb_last_error.restype = cp
    L.wdb_init.argtypes = [ci]
    L.wdb_device_count.argtypes = [C.POINTER(ci)]
    L.wdb_set_udf_source.argtypes = [cp]
    L.wdb_set_option.argtypes = [cp, i64]
    L.wdb_get_option.argtypes = [cp, P64]
    L.wdb_get_stats.argtypes = [C.POINTER(Stats)]
    L.wdb_project_filter.argtypes = [ci, vp, PC, ci, cp, cp, vp, i64, ci, vp, P64]
    L.wdb_agg_create.argtypes = [ci, i64, ci, C.POINTER(vp)]
    L.wdb_agg_destroy.argtypes = [vp]
    L.wdb_agg_reset.argtypes = [vp, vp]
    L.wdb_agg_set_key_range.argtypes = [vp, ci, i64, i64]
    L.wdb_agg_consume.argtypes = [vp, vp, PC, ci, cp, cp, cp, i64, i64]
    L.wdb_agg_merge.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64]
    L.wdb_agg_size.argtypes = [vp, vp, P64]
    L.wdb_agg_spilled.argtypes = [vp, vp, P64]
    L.wdb_agg_export.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp, vp, vp, i64, P64]
    L.wdb_group_agg.argtypes = [ci, vp, PC, ci, cp, cp, cp, ci, ci, i64, i64, vp, vp, i64, P64]
    L.wdb_topk.argtypes = [ci, vp, PC, ci, cp, cp, cp, ci, i64, i64, i64, vp, vp, P64]
    L.wdb_sort_float.argtypes = [ci, vp, vp, i64, ci]
    L.wdb_sort_pairs.argtypes = [ci, vp, vp, vp, i64, ci]
    L.wdb_column_minmax.argtypes = [ci, vp, PC, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.wdb_multi_project_filter_host.argtypes = [ci, C.POINTER(ci), PC, ci, cp, cp, vp, i64, ci, P64]
    L.wdb_zonemap_build.argtypes = [ci, vp, PC, i64, C.POINTER(vp)]
    L.wdb_upload_column.argtypes = [ci, vp, PC, vp, i64, C.POINTER(vp), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.wdb_zonemap_destroy.argtypes = [vp]
    L.wdb_zonemap_info.argtypes = [vp, P64, P64]
    L.wdb_project_filter_pruned.argtypes = [ci, vp, PC, ci, cp, cp, vp, i64, ci, vp, P64, C.POINTER(Prune), ci, P64]
    L.wdb_agg_consume_pruned.argtypes = [vp, vp, PC, ci, cp, cp, cp, i64, i64, C.POINTER(Prune), ci, P64]
    L.wdb_topk_pruned.argtypes = [ci, vp, PC, ci, cp, cp, cp, ci, i64, i64, i64, vp, vp, P64, C.POINTER(Prune), ci, P64]
    L.wdb_shard_range.argtypes = [i64, ci, ci, P64, P64]
    L.wdb_synth_f32.argtypes = [ci, vp, vp, i64, C.c_uint64, C.c_float, C.c_float, i64]
    L.wdb_synth_i32.argtypes = [ci, vp, vp, i64, C.c_uint64, C.c_int32, C.c_int32, i64]
    L.wdb_debug_compile.argtypes = [cp, PC, ci, cp, cp, cp, ci, cp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.wdb_free.argtypes = [vp]
    L.wdb_comm_unique_id.argtypes = [vp]
    L.wdb_comm_init_rank.argtypes = [ci, ci, ci, vp, C.POINTER(vp)]
    L.wdb_comm_init_all.argtypes = [ci, C.POINTER(ci), C.POINTER(vp)]
    L.wdb_comm_destroy.argtypes = [vp]
    L.wdb_comm_info.argtypes = [vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
    L.wdb_multi_project_filter.argtypes = [vp, vp, PC, ci, cp, cp, vp, i64, ci, vp, P64]
    L.wdb_multi_group_agg.argtypes = [vp, vp, PC, ci, cp, cp, cp, ci, ci, i64, i64, i64, ci, i64, i64, vp, vp, i64, vp, P64]
    L.wdb_multi_topk.argtypes = [vp, vp, PC, ci, cp, cp, cp, ci, i64, i64, i64, i64, vp, vp, vp, P64]
    L.wdb_multi_group_agg_host.argtypes = [ci, C.POINTER(ci), PC, ci, cp, cp, cp, ci, ci, i64, i64, vp, vp, i64, P64]
    L.wdb_multi_topk_host.argtypes = [ci, C.POINTER(ci), PC, ci, cp, cp, cp, ci, i64, i64, i64, vp, P64]
    L.wdb_join_build.argtypes = [ci, vp, PC, C.POINTER(vp)]
    L.wdb_join_probe.argtypes = [vp, vp, PC, vp, vp, i64, P64]
    L.wdb_join_info.argtypes = [vp, P64, C.POINTER(ci), P64]
    L.wdb_join_destroy.argtypes = [vp]
    L.wdb_gather.argtypes = [ci, vp, PC, vp, i64, vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise WarpcoreError(lib().wdb_last_error().decode(errors="replace"))


def enc(s):
    return None if s is None else s.encode()


def make_cols(schema):
    """schema: list of (name, dtype, device_ptr, length) -> (ctypes array, count)."""
    arr = (Col * max(len(schema), 1))()
    for i, (name, dtype, ptr, length) in enumerate(schema):
        arr[i].name = name.encode()
        arr[i].dtype = dtype
        arr[i].dptr = ptr
        arr[i].len = length
    return arr, len(schema)


def set_option(key, value):
    """value=None removes the override (built-in default)."""
    check(lib().wdb_set_option(key.encode(), -(1 << 63) if value is None else int(value)))


def set_udf_source(src):
    check(lib().wdb_set_udf_source(enc(src)))


def stats():
    s = Stats()
    check(lib().wdb_get_stats(C.byref(s)))
    return dict(kernels_compiled=s.kernels_compiled, cache_hits=s.cache_hits, launches=s.launches,
                last_compile_ms=s.last_compile_ms)


def shard_range(n, ndev, dev):
    s, e = C.c_int64(0), C.c_int64(0)
    check(lib().wdb_shard_range(n, ndev, dev, C.byref(s), C.byref(e)))
    return s.value, e.value


def debug_compile(kind, schema, expr_a, expr_b=None, cond=None, mode=DENSE, arch="sm_100a", want_cubin=True):
    """Generate (and compile) the kernel a call would use; needs no GPU.  Returns (source, cubin bytes)."""
    cols, n = make_cols(schema)
    src, cub, size = C.c_void_p(), C.c_void_p(), C.c_size_t(0)
    check(lib().wdb_debug_compile(kind.encode(), cols, n, enc(expr_a), enc(expr_b), enc(cond), mode, arch.encode(),
                                  C.byref(src), C.byref(cub) if want_cubin else None, C.byref(size)))
    source = C.string_at(src).decode()
    lib().wdb_free(src)
    cubin = b""
    if want_cubin:
        cubin = C.string_at(cub, size.value)
        lib().wdb_free(cub)
    return source, cubin
