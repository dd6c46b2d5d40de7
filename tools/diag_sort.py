#!/usr/bin/env python
"""Device sorts (jit_sort_float / jit_sort_pairs replacements, full ORDER BY): throughput vs torch.sort."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops
wc.check(wc.lib().wdb_init(0))
def t(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for n in (1 << 24, 1 << 28):
    src = ops.synth_f32(n, 7, 0.0, 1e6)
    buf = torch.empty_like(src)
    def ours():
        buf.copy_(src); ops.sort_float(buf, False)
    def copy_only():
        buf.copy_(src)
    ms_copy = t(copy_only)
    ms = t(ours) - ms_copy
    ms_torch = t(lambda: torch.sort(src, descending=True))
    keys = ops.synth_i32(n, 8, 0, 1 << 30); kb = torch.empty_like(keys)
    def pairs():
        kb.copy_(keys); buf.copy_(src); ops.sort_pairs(kb, buf, True)
    ms_pairs = t(pairs) - 2 * ms_copy
    print(json.dumps({"n": n, "sort_float_ms": round(ms, 3), "gkeys_s": round(n / ms / 1e6, 2), "torch_sort_ms": round(ms_torch, 3),
                      "sort_pairs_ms": round(ms_pairs, 3), "pairs_gkeys_s": round(n / ms_pairs / 1e6, 2)}), flush=True)
    del src, buf, keys, kb
