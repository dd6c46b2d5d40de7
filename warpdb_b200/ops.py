"""Device-level operators: thin Python calls into the C ABI (include/warpcore.h) on torch tensors.

PyTorch supplies device memory and streams only; every operator below runs the hand-written CUDA
kernels of libwarpcore.so and raises WarpcoreError if that is not possible.
"""
import ctypes as C

import torch

from . import _core as wc

_TORCH2DT = {torch.int32: wc.INT32, torch.int64: wc.INT64, torch.float32: wc.FLOAT32, torch.float64: wc.FLOAT64}


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev_index(table, device=None):
    if device is not None:
        return torch.device(device).index or 0
    for t in table.values():
        return t.device.index or 0
    return torch.cuda.current_device()


def schema_of(table):
    """table: dict name -> 1-D contiguous CUDA tensor (int32/int64/float32/float64)."""
    out = []
    for name, t in table.items():
        if not t.is_cuda or not t.is_contiguous() or t.dim() != 1:
            raise wc.WarpcoreError(f"column {name} must be a contiguous 1-D CUDA tensor")
        if t.dtype not in _TORCH2DT:
            raise wc.WarpcoreError(f"column {name} has unsupported dtype {t.dtype}")
        out.append((name, _TORCH2DT[t.dtype], t.data_ptr(), t.shape[0]))
    return out


def num_rows(table):
    for t in table.values():
        return t.shape[0]
    return 0


def project_filter(table, expr, cond=None, mode=wc.DENSE, out=None, n=None, sync_count=True):
    """Fused filter+project.  expr/cond are to_cuda_expr() strings.  Returns (out, count):
    dense modes: out float32[n], count == n; COMPACT: out[:count] holds the survivors in row order."""
    dev = _dev_index(table)
    n = num_rows(table) if n is None else n
    if out is None:
        out = torch.empty(max(n, 1), dtype=torch.float32, device=f"cuda:{dev}")[:n]
    cols, nc = wc.make_cols(schema_of(table))
    cnt = C.c_int64(0)
    wc.check(wc.lib().wdb_project_filter(dev, _stream(dev), cols, nc, wc.enc(expr), wc.enc(cond or ""), out.data_ptr(), n,
                                         mode, None, C.byref(cnt) if sync_count else None))
    return out, (cnt.value if sync_count else None)


def synth_f32(n, seed, lo, hi, row0=0, device=0):
    out = torch.empty(n, dtype=torch.float32, device=f"cuda:{device}")
    wc.check(wc.lib().wdb_synth_f32(device, _stream(device), out.data_ptr(), n, seed, lo, hi, row0))
    return out


def synth_i32(n, seed, lo, hi_excl, row0=0, device=0):
    out = torch.empty(n, dtype=torch.int32, device=f"cuda:{device}")
    wc.check(wc.lib().wdb_synth_i32(device, _stream(device), out.data_ptr(), n, seed, lo, hi_excl, row0))
    return out


def column_minmax(column, name="col"):
    """(min, max) of one device column in one streaming pass (wdb_column_minmax): the statistics the
    reference's TableStats were meant to hold (include/csv_loader.hpp:22-37)."""
    dev = column.device.index or 0
    cols, _ = wc.make_cols(schema_of({name: column}))
    lo, hi = C.c_double(0), C.c_double(0)
    wc.check(wc.lib().wdb_column_minmax(dev, _stream(dev), cols, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


class AggTable:
    """Device-resident hash aggregation table (wdb_agg_*): fold row chunks and partial aggregates of
    other GPUs into it, then export ordered groups."""

    def __init__(self, device=0, expected_groups=1 << 16, needs=wc.NEED_SUM | wc.NEED_COUNT):
        self.device = device
        self.needs = needs
        self.handle = C.c_void_p()
        wc.check(wc.lib().wdb_agg_create(device, expected_groups, needs, C.byref(self.handle)))

    def close(self):
        if self.handle:
            wc.lib().wdb_agg_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def reset(self):
        wc.check(wc.lib().wdb_agg_reset(self.handle, _stream(self.device)))

    def set_key_range(self, lo=None, hi=None):
        """Optimizer statistics: every key of the following consume calls lies in [lo, hi] (None: unknown)."""
        known = lo is not None and hi is not None
        wc.check(wc.lib().wdb_agg_set_key_range(self.handle, int(known), int(lo) if known else 0, int(hi) if known else -1))

    def consume(self, table, val_expr, key_expr, cond=None, n=None, row_base=0, preds=None):
        """preds: optional list of (ZoneMap, op, constant) implied by `cond` (zone-map pruning); returns the
        number of live zones when given."""
        n = num_rows(table) if n is None else n
        cols, nc = wc.make_cols(schema_of(table))
        if preds:
            arr = _prune_array(preds)
            live = C.c_int64(0)
            wc.check(wc.lib().wdb_agg_consume_pruned(self.handle, _stream(self.device), cols, nc, wc.enc(val_expr), wc.enc(key_expr), wc.enc(cond or ""), n,
                                                     row_base, arr, len(preds), C.byref(live)))
            return live.value
        wc.check(wc.lib().wdb_agg_consume(self.handle, _stream(self.device), cols, nc, wc.enc(val_expr), wc.enc(key_expr),
                                          wc.enc(cond or ""), n, row_base))

    def merge(self, part):
        """part: dict with keys (int32) and whichever of sums/counts/mins/maxs/first the table tracks."""
        def p(name):
            t = part.get(name)
            return t.data_ptr() if t is not None and t.numel() else None
        m = part["keys"].numel()
        wc.check(wc.lib().wdb_agg_merge(self.handle, _stream(self.device), p("keys"), p("sums"), p("counts"), p("mins"),
                                        p("maxs"), p("first"), m))

    def spilled(self):
        """Rows since the last reset that bypassed the shared-memory accumulators (tuning statistic)."""
        g = C.c_int64(0)
        wc.check(wc.lib().wdb_agg_spilled(self.handle, _stream(self.device), C.byref(g)))
        return g.value

    def size(self):
        g = C.c_int64(0)
        wc.check(wc.lib().wdb_agg_size(self.handle, _stream(self.device), C.byref(g)))
        return g.value

    def export(self, agg=wc.SUM, order=wc.ORDER_KEY_ASC, raw=True):
        """Returns dict(keys, vals[, sums, counts, mins, maxs, first]) of CUDA tensors."""
        g = self.size()
        dev = f"cuda:{self.device}"
        out = {"keys": torch.empty(g, dtype=torch.int32, device=dev)}
        have_val = {wc.SUM: wc.NEED_SUM, wc.AVG: wc.NEED_SUM | wc.NEED_COUNT, wc.COUNT: wc.NEED_COUNT,
                    wc.MIN: wc.NEED_MINMAX, wc.MAX: wc.NEED_MINMAX}[agg]
        if (have_val & ~self.needs) == 0:
            out["vals"] = torch.empty(g, dtype=torch.float32, device=dev)
        if raw:
            if self.needs & wc.NEED_SUM:
                out["sums"] = torch.empty(g, dtype=torch.float64, device=dev)
            if self.needs & wc.NEED_COUNT:
                out["counts"] = torch.empty(g, dtype=torch.int64, device=dev)
            if self.needs & wc.NEED_MINMAX:
                out["mins"] = torch.empty(g, dtype=torch.float64, device=dev)
                out["maxs"] = torch.empty(g, dtype=torch.float64, device=dev)
            if self.needs & wc.NEED_FIRST_ROW:
                out["first"] = torch.empty(g, dtype=torch.int64, device=dev)

        def p(name):
            t = out.get(name)
            return t.data_ptr() if t is not None and t.numel() else None
        gg = C.c_int64(0)
        wc.check(wc.lib().wdb_agg_export(self.handle, _stream(self.device), agg, order, p("keys"), p("vals"), p("sums"),
                                         p("counts"), p("mins"), p("maxs"), p("first"), g, C.byref(gg)))
        assert gg.value == g
        return out


def needs_for(agg, order=wc.ORDER_KEY_ASC):
    n = {wc.SUM: wc.NEED_SUM, wc.AVG: wc.NEED_SUM | wc.NEED_COUNT, wc.COUNT: wc.NEED_COUNT, wc.MIN: wc.NEED_MINMAX,
         wc.MAX: wc.NEED_MINMAX}[agg]
    return n | (wc.NEED_FIRST_ROW if order == wc.ORDER_FIRST else 0)


def group_agg(table, val_expr, key_expr, cond=None, agg=wc.SUM, order=wc.ORDER_KEY_ASC, expected_groups=0, cap=None):
    """One-shot GROUP BY: returns (keys int32[G], vals float32[G])."""
    dev = _dev_index(table)
    n = num_rows(table)
    if cap is None:
        cap = max(min(n, 1 << 26), 1)
        if expected_groups:
            cap = max(min(cap, 2 * expected_groups + 16), 1)
    keys = torch.empty(cap, dtype=torch.int32, device=f"cuda:{dev}")
    vals = torch.empty(cap, dtype=torch.float32, device=f"cuda:{dev}")
    cols, nc = wc.make_cols(schema_of(table))
    g = C.c_int64(0)
    wc.check(wc.lib().wdb_group_agg(dev, _stream(dev), cols, nc, wc.enc(val_expr), wc.enc(key_expr), wc.enc(cond or ""), agg,
                                    order, n, expected_groups, keys.data_ptr(), vals.data_ptr(), cap, C.byref(g)))
    return keys[:g.value], vals[:g.value]


def _prune_array(preds):
    arr = (wc.Prune * max(len(preds), 1))()
    for i, (zm, op, value) in enumerate(preds):
        arr[i].zonemap = zm.handle
        arr[i].op = wc.PRUNE_OPS[op]
        arr[i].value = float(value)
    return arr


def topk(table, key_expr, val_expr=None, cond=None, descending=True, k=5, offset=0, want_keys=False, preds=None):
    """ORDER BY key_expr [LIMIT k [OFFSET offset]] (k < 0: no limit).  Returns vals (and keys).
    preds: optional zone-map pruning terms implied by `cond` (list of (ZoneMap, op, constant))."""
    dev = _dev_index(table)
    n = num_rows(table)
    m = n if k < 0 else min(k, n)
    vals = torch.empty(max(m, 1), dtype=torch.float32, device=f"cuda:{dev}")
    keys = torch.empty(max(m, 1), dtype=torch.float32, device=f"cuda:{dev}")
    cols, nc = wc.make_cols(schema_of(table))
    cnt = C.c_int64(0)
    if preds:
        wc.check(wc.lib().wdb_topk_pruned(dev, _stream(dev), cols, nc, wc.enc(key_expr), wc.enc(val_expr), wc.enc(cond or ""), int(descending), k, offset, n,
                                          vals.data_ptr(), keys.data_ptr(), C.byref(cnt), _prune_array(preds), len(preds), None))
    else:
        wc.check(wc.lib().wdb_topk(dev, _stream(dev), cols, nc, wc.enc(key_expr), wc.enc(val_expr), wc.enc(cond or ""),
                                   int(descending), k, offset, n, vals.data_ptr(), keys.data_ptr(), C.byref(cnt)))
    if want_keys:
        return vals[:cnt.value], keys[:cnt.value]
    return vals[:cnt.value]


def sort_float(vals, ascending=True):
    """In-place stable device sort (jit_sort_float)."""
    dev = vals.device.index or 0
    wc.check(wc.lib().wdb_sort_float(dev, _stream(dev), vals.data_ptr(), vals.numel(), int(ascending)))
    return vals


def sort_pairs(keys, vals, ascending=True):
    """In-place stable device sort of (int32 key, float32 value) pairs by key (jit_sort_pairs)."""
    dev = keys.device.index or 0
    wc.check(wc.lib().wdb_sort_pairs(dev, _stream(dev), keys.data_ptr(), vals.data_ptr(), keys.numel(), int(ascending)))
    return keys, vals


class ZoneMap:
    """Per-zone min/max of one column (wdb_zonemap_*), built in one streaming pass."""

    def __init__(self, column, name="col", zone_rows=0):
        self.device = column.device.index or 0
        self.rows = column.shape[0]
        self.handle = C.c_void_p()
        cols, _ = wc.make_cols(schema_of({name: column}))
        wc.check(wc.lib().wdb_zonemap_build(self.device, _stream(self.device), cols, zone_rows, C.byref(self.handle)))
        zr, nz = C.c_int64(0), C.c_int64(0)
        wc.check(wc.lib().wdb_zonemap_info(self.handle, C.byref(zr), C.byref(nz)))
        self.zone_rows, self.nzones = zr.value, nz.value

    def close(self):
        if self.handle:
            wc.lib().wdb_zonemap_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def upload_column(host, name="col", device=0, zone_rows=0):
    """Resident ingest of one host column (wdb_upload_column): a NumPy array or CPU tensor (pinned or pageable) goes to
    the device in asynchronous chunks while its zone map is built chunk by chunk behind the copies.
    Returns (device tensor, ZoneMap, (min, max))."""
    t = host if isinstance(host, torch.Tensor) else torch.from_numpy(host)
    if t.is_cuda or not t.is_contiguous() or t.dim() != 1 or t.dtype not in _TORCH2DT:
        raise wc.WarpcoreError("upload_column needs a contiguous 1-D CPU tensor of int32/int64/float32/float64")
    out = torch.empty(t.shape[0], dtype=t.dtype, device=f"cuda:{device}")
    cols, _ = wc.make_cols([(name, _TORCH2DT[t.dtype], t.data_ptr(), t.shape[0])])
    zm = ZoneMap.__new__(ZoneMap)
    zm.device, zm.rows, zm.handle = device, t.shape[0], C.c_void_p()
    lo, hi = C.c_double(0), C.c_double(0)
    wc.check(wc.lib().wdb_upload_column(device, _stream(device), cols, out.data_ptr(), zone_rows, C.byref(zm.handle), C.byref(lo), C.byref(hi)))
    zr, nz = C.c_int64(0), C.c_int64(0)
    wc.check(wc.lib().wdb_zonemap_info(zm.handle, C.byref(zr), C.byref(nz)))
    zm.zone_rows, zm.nzones = zr.value, nz.value
    return out, zm, (lo.value, hi.value)


def project_filter_pruned(table, expr, cond, preds, mode=wc.DENSE_ZERO, out=None, sync=True):
    """wdb_project_filter_pruned.  preds: list of (ZoneMap, op, constant) with op in > >= < <= == != and
    every term implied by `cond`.  Returns (out, count, zones_live)."""
    dev = _dev_index(table)
    n = num_rows(table)
    if out is None:
        out = torch.empty(max(n, 1), dtype=torch.float32, device=f"cuda:{dev}")[:n]
    cols, nc = wc.make_cols(schema_of(table))
    arr = (wc.Prune * max(len(preds), 1))()
    for i, (zm, op, value) in enumerate(preds):
        arr[i].zonemap = zm.handle
        arr[i].op = wc.PRUNE_OPS[op]
        arr[i].value = float(value)
    cnt, live = C.c_int64(0), C.c_int64(0)
    wc.check(wc.lib().wdb_project_filter_pruned(dev, _stream(dev), cols, nc, wc.enc(expr), wc.enc(cond), out.data_ptr(), n, mode, None,
                                                C.byref(cnt) if sync else None, arr, len(preds), C.byref(live) if sync else None))
    return out, (cnt.value if sync else None), (live.value if sync else None)


class JoinIndex:
    """The build side of an inner equi-join (wdb_join_build): the (key, row) pairs of an int32 / int64
    device column, sorted once, probed any number of times."""

    def __init__(self, build_key, name="key"):
        if not build_key.is_cuda or build_key.dim() != 1 or build_key.dtype not in (torch.int32, torch.int64):
            raise wc.WarpcoreError("JoinIndex needs a 1-D int32/int64 CUDA tensor")
        self.device = build_key.device.index or 0
        self.rows = build_key.shape[0]
        self.handle = C.c_void_p()
        cols, _ = wc.make_cols([(name, _TORCH2DT[build_key.dtype], build_key.data_ptr(), self.rows)])
        wc.check(wc.lib().wdb_join_build(self.device, _stream(self.device), cols, C.byref(self.handle)))
        span = C.c_int64(0)
        wc.check(wc.lib().wdb_join_info(self.handle, None, None, C.byref(span)))
        self.direct_span = span.value          # > 0: probes index a table of the key range instead of searching

    def count(self, probe_key):
        """number of (probe row, build row) pairs with equal keys"""
        cols, _ = wc.make_cols([("probe", _TORCH2DT[probe_key.dtype], probe_key.data_ptr(), probe_key.shape[0])])
        n = C.c_int64(0)
        wc.check(wc.lib().wdb_join_probe(self.handle, _stream(self.device), cols, None, None, 0, C.byref(n)))
        return n.value

    def probe(self, probe_key):
        """(probe_rows, build_rows): int64 tensors, every pair with equal keys, ordered by probe row then build row"""
        if not probe_key.is_cuda or probe_key.dim() != 1 or probe_key.dtype not in (torch.int32, torch.int64):
            raise wc.WarpcoreError("probe needs a 1-D int32/int64 CUDA tensor")
        pairs = self.count(probe_key)
        pr = torch.empty(pairs, dtype=torch.int64, device=f"cuda:{self.device}")
        br = torch.empty(pairs, dtype=torch.int64, device=f"cuda:{self.device}")
        if pairs:
            cols, _ = wc.make_cols([("probe", _TORCH2DT[probe_key.dtype], probe_key.data_ptr(), probe_key.shape[0])])
            n = C.c_int64(0)
            wc.check(wc.lib().wdb_join_probe(self.handle, _stream(self.device), cols, pr.data_ptr(), br.data_ptr(), pairs, C.byref(n)))
            assert n.value == pairs
        return pr, br

    def close(self):
        if self.handle:
            wc.lib().wdb_join_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def gather(src, rows, out=None):
    """out[i] = src[rows[i]] (wdb_gather); rows: int64 CUDA tensor, or None for a plain copy"""
    dev = src.device.index or 0
    count = src.shape[0] if rows is None else rows.shape[0]
    if out is None:
        out = torch.empty(count, dtype=src.dtype, device=src.device)
    cols, _ = wc.make_cols([("src", _TORCH2DT[src.dtype], src.data_ptr(), src.shape[0])])
    wc.check(wc.lib().wdb_gather(dev, _stream(dev), cols, None if rows is None else rows.data_ptr(), count, out.data_ptr()))
    return out


class Comm:
    """One GPU's membership in a group of GPUs (wdb_comm_*): the handle the sharded operators of the
    core (wdb_multi_*) take.  The cross-GPU merges run inside libwarpcore over NCCL/NVLink; Python only
    carries the 128-byte NCCL id from rank 0 to the other ranks (torch.distributed, any backend)."""

    def __init__(self, device=0, rank=0, world=1, unique_id=None):
        self.device, self.rank, self.world = device, rank, world
        self.handle = C.c_void_p()
        buf = None
        if world > 1:
            if unique_id is None or len(unique_id) != 128:
                raise wc.WarpcoreError("a communicator of several ranks needs the 128-byte id of Comm.unique_id()")
            buf = C.create_string_buffer(bytes(unique_id), 128)
        wc.check(wc.lib().wdb_comm_init_rank(device, world, rank, buf, C.byref(self.handle)))

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        wc.check(wc.lib().wdb_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch(cls, device, group=None):
        """One process per GPU under torch.distributed: rank 0 creates the id, broadcast carries it."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return cls(device, 0, 1)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world == 1:
            return cls(device, 0, 1)
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{device}" if on_gpu else "cpu")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(device, rank, world, bytes(t.cpu().numpy().tobytes()))

    def close(self):
        if self.handle:
            wc.lib().wdb_comm_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---- sharded operators: `table` is this rank's shard -------------------------------------------
    def project_filter(self, table, expr, cond=None, mode=wc.COMPACT, out=None, sync=True):
        """Returns (out, (rows written, global offset, global total)) -- the triple is None when sync=False."""
        n = num_rows(table)
        if out is None:
            out = torch.empty(max(n, 1), dtype=torch.float32, device=f"cuda:{self.device}")[:n]
        cols, nc = wc.make_cols(schema_of(table))
        h3 = (C.c_int64 * 3)()
        wc.check(wc.lib().wdb_multi_project_filter(self.handle, _stream(self.device), cols, nc, wc.enc(expr), wc.enc(cond or ""), out.data_ptr(), n, mode,
                                                   None, h3 if sync else None))
        return out, (tuple(h3) if sync else None)

    def group_agg(self, table, val_expr, key_expr, cond=None, agg=wc.SUM, order=wc.ORDER_KEY_ASC, row_base=0, expected_groups=0,
                  key_range=None, cap=None, out=None, sync=True):
        """Final groups of the whole sharded table on every rank: (keys int32[G], vals float32[G]).
        out=(keys, vals, groups) reuses caller buffers; sync=False returns them unsliced with the group
        count left on the device in groups[0]."""
        n = num_rows(table)
        dev = f"cuda:{self.device}"
        if out is None:
            if cap is None:
                cap = max(2 * expected_groups + 16, 1 << 16) if expected_groups else 1 << 24
                if key_range is not None:
                    cap = min(cap, key_range[1] - key_range[0] + 1)
            out = (torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.float32, device=dev),
                   torch.zeros(1, dtype=torch.int64, device=dev))
        keys, vals, groups = out
        cols, nc = wc.make_cols(schema_of(table))
        g = C.c_int64(0)
        known = key_range is not None
        wc.check(wc.lib().wdb_multi_group_agg(self.handle, _stream(self.device), cols, nc, wc.enc(val_expr), wc.enc(key_expr), wc.enc(cond or ""), agg, order,
                                              n, row_base, expected_groups, int(known), int(key_range[0]) if known else 0, int(key_range[1]) if known else -1,
                                              keys.data_ptr(), vals.data_ptr(), keys.numel(), groups.data_ptr(), C.byref(g) if sync else None))
        if sync:
            return keys[:g.value], vals[:g.value]
        return keys, vals, groups

    def topk(self, table, key_expr, val_expr=None, cond=None, descending=True, k=5, offset=0, row_base=0, out=None, sync=True):
        """ORDER BY key [DESC] LIMIT k OFFSET offset over the whole sharded table; same values on every rank."""
        n = num_rows(table)
        dev = f"cuda:{self.device}"
        if out is None:
            out = (torch.empty(max(k, 1), dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.int64, device=dev))
        vals, cnt = out
        cols, nc = wc.make_cols(schema_of(table))
        h = C.c_int64(0)
        wc.check(wc.lib().wdb_multi_topk(self.handle, _stream(self.device), cols, nc, wc.enc(key_expr), wc.enc(val_expr), wc.enc(cond or ""), int(descending),
                                         k, offset, n, row_base, vals.data_ptr(), None, cnt.data_ptr(), C.byref(h) if sync else None))
        return vals[:h.value] if sync else vals
