#!/usr/bin/env python
"""Exercise every kernel once on small inputs (for compute-sanitizer memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300_007
wc.check(wc.lib().wdb_init(0))
wc.set_udf_source("__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n")
price = ops.synth_f32(n, 1, 0.0, 100.0)
qty = ops.synth_i32(n, 2, 0, 1000)
t = {"price": price, "quantity": qty}
e, c = "(price[idx] * 0.9f)", "(price[idx] > 50.0f)"
for mode in (wc.DENSE, wc.DENSE_ZERO):
    ops.project_filter(t, "((price[idx] * quantity[idx]) * 1.08f)", None, mode)
    ops.project_filter(t, e, c, mode)
for v in (0, 1, 2, 3):
    wc.set_option("compact.variant", v)
    _, cnt = ops.project_filter(t, e, c, wc.COMPACT)
wc.set_option("compact.variant", None)
wc.set_option("project.variant", 2)
ops.project_filter(t, "((price[idx] * quantity[idx]) * 1.08f)", None, wc.DENSE)
wc.set_option("project.variant", None)
for G in (1000, 200000):
    q = ops.synth_i32(n, 3, 0, G)
    for agg in (wc.SUM, wc.AVG, wc.MAX):
        ops.group_agg({"price": price, "quantity": q}, "price[idx]", "quantity[idx]", agg=agg, expected_groups=G)
ops.group_agg(t, "price[idx]", "quantity[idx]", order=wc.ORDER_FIRST, expected_groups=1000)
wc.set_option("group.wp_slots", 2048)
ops.group_agg(t, "price[idx]", "quantity[idx]", expected_groups=1000)
wc.set_option("group.wp_slots", None)
for k, off in ((5, 0), (16, 0), (100, 3), (-1, 0)):
    ops.topk(t, "discount(price[idx], 0.9f)", "quantity[idx]", c, True, k, off)
ops.sort_float(price.clone(), False)
ops.sort_pairs(qty.clone(), price.clone(), True)
zm = ops.ZoneMap(price, "price")
for mode in (wc.DENSE, wc.DENSE_ZERO, wc.COMPACT):
    ops.project_filter_pruned(t, e, c, [(zm, ">", 50.0)], mode)
zm.close()
torch.cuda.synchronize()
print("sanitize target ok", cnt, wc.stats())
