#!/usr/bin/env python
"""ncu target: the staged two-pass compaction (variant 5) at 1 % selectivity, 1e9 rows, a few calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
wc.check(wc.lib().wdb_init(0))
wc.set_option("compact.variant", 5)
price = ops.synth_f32(n, 0xC0FFEE + 3, 0.0, 20.0 / 0.99)
out = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(3):
    ops.project_filter({"price": price}, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT, out=out, sync_count=False)
torch.cuda.synchronize()
print("ok")
