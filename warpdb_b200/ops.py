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
