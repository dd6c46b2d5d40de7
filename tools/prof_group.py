#!/usr/bin/env python
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops
n = 1 << 28
wc.check(wc.lib().wdb_init(0))
price = ops.synth_f32(n, 0xC0FFEE + 4, 0.0, 100.0)
qty = ops.synth_i32(n, 0xC0FFEE + 104, 0, 1000)
t = {"price": price, "quantity": qty}
for wp in (0, -1):
    wc.set_option("group.wp_slots", wp)
    for _ in range(2):
        tab = ops.AggTable(0, 1000, wc.NEED_SUM)
        tab.consume(t, "price[idx]", "quantity[idx]")
        torch.cuda.synchronize()
        tab.close()
print("ok")
