#!/usr/bin/env python
"""ncu target, round 2: ONE launch of each hot kernel at the size bench.py reports it on (BASELINE sizes
for one GPU): projection 1e9, compaction 4e9 at 1/50/99 %, GROUP BY 2e9 at 1 K / 10 M keys, top-5 8e9."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from warpdb_b200 import _core as wc, ops

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["project", "compact1", "compact50", "compact99", "group1k", "group10m", "topk"]
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
wc.check(wc.lib().wdb_init(0))
wc.set_udf_source("__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n")
S = 0xC0FFEE
for w in which:
    if w == "project":
        n = int(1e9 * scale)
        t = {"price": ops.synth_f32(n, S + 2, 0.0, 100.0), "quantity": ops.synth_i32(n, S + 102, 1, 101)}
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        for _ in range(2):
            ops.project_filter(t, "((price[idx] * quantity[idx]) * 1.08f)", None, wc.DENSE, out=out, sync_count=False)
    elif w.startswith("compact"):
        n = int(4e9 * scale)
        sel = {"compact1": 0.01, "compact50": 0.5, "compact99": 0.99}[w]
        t = {"price": ops.synth_f32(n, S + 3, 0.0, 20.0 / (1.0 - sel))}
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        for _ in range(3 if sel < 0.04 else 2):   # selective filters: the first call runs the L2-parked slabs, the next ones the staged kernels
            ops.project_filter(t, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT, out=out, sync_count=False)
            torch.cuda.synchronize()
    elif w.startswith("group"):
        n = int(2e9 * scale)
        G = 1000 if w == "group1k" else 10_000_000
        t = {"price": ops.synth_f32(n, S + 4, 0.0, 100.0), "quantity": ops.synth_i32(n, S + 104, 0, G)}
        tab = ops.AggTable(0, 1024, wc.NEED_SUM)
        tab.set_key_range(0, G - 1)
        for _ in range(2):
            tab.reset()
            tab.consume(t, "price[idx]", "quantity[idx]")
        tab.close()
    elif w == "topk":
        n = int(8e9 * scale)
        t = {"price": ops.synth_f32(n, S + 5, 0.0, 1e6)}
        for _ in range(2):
            ops.topk(t, "discount(price[idx], 0.9f)", None, None, True, 5)
    torch.cuda.synchronize()
    del t
    torch.cuda.empty_cache()
print("ok", wc.stats())
