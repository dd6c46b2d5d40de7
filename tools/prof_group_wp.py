#!/usr/bin/env python
"""ncu target: wdb_group_wp in index mode (G=500, slow regime), index mode G=200 (fast), dense mode G=1000."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 27
wc.check(wc.lib().wdb_init(0))
price = ops.synth_f32(n, 0xC0FFEE + 4, 0.0, 100.0)
for G, rng in ((500, None), (200, None), (1000, (0, 999))):
    qty = ops.synth_i32(n, 0xC0FFEE + 104, 0, G)
    tab = ops.AggTable(0, G, wc.NEED_SUM)
    if rng:
        tab.set_key_range(*rng)
    for rep in range(2):
        tab.reset(); tab.consume({"price": price, "quantity": qty}, "price[idx]", "quantity[idx]")
    torch.cuda.synchronize()
    print(G, tab.size()); tab.close()
