#!/usr/bin/env python
"""Small driver for ncu: launches each hot kernel a few times on 2^28-row columns (larger than L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["project", "compact", "group", "group_atomic", "group_dense", "topk"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
wc.check(wc.lib().wdb_init(0))
wc.set_udf_source("__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n")
price = ops.synth_f32(n, 0xC0FFEE + 2, 0.0, 100.0)
qty = ops.synth_i32(n, 0xC0FFEE + 102, 0, 1000)
out = torch.empty(n, dtype=torch.float32, device="cuda")
t = {"price": price, "quantity": qty}
qty10m = ops.synth_i32(n, 0xC0FFEE + 104, 0, 10_000_000) if "group_dense" in which else None
for _ in range(reps):
    if "project" in which:
        ops.project_filter(t, "((price[idx] * quantity[idx]) * 1.08f)", None, wc.DENSE, out=out, sync_count=False)
    if "compact" in which:
        ops.project_filter({"price": price}, "(price[idx] * 0.9f)", "(price[idx] > 50.0f)", wc.COMPACT, out=out, sync_count=False)
    if "group" in which:
        ops.group_agg(t, "price[idx]", "quantity[idx]", expected_groups=1000)
    if "group_atomic" in which:   # keys of unknown range: shared-memory atomic table
        wc.set_option("group.wp_max_span", 0)
        ops.group_agg(t, "price[idx]", "quantity[idx]", expected_groups=1000)
        wc.set_option("group.wp_max_span", None)
    if "group_dense" in which:    # 10 M keys: direct-addressed table, two L2-sized slices
        ops.group_agg({"price": price, "quantity": qty10m}, "price[idx]", "quantity[idx]", expected_groups=10_000_000)
    if "topk" in which:
        ops.topk({"price": price}, "discount(price[idx], 0.9f)", None, None, True, 5)
torch.cuda.synchronize()
print("ok", wc.stats())
