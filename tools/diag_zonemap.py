#!/usr/bin/env python
"""Zone-map pruning benefit: sorted / clustered / random price column, `price * 0.9 WHERE price > X` at ~1 % selectivity."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
wc.check(wc.lib().wdb_init(0))
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / iters
out = torch.empty(n, dtype=torch.float32, device="cuda")
base = ops.synth_f32(n, 0xC0FFEE + 9, 0.0, 100.0)
for layout in ("random", "clustered", "sorted"):
    if layout == "sorted":
        price = torch.sort(base).values
    elif layout == "clustered":   # 1M-row runs, each from a 1-unit band; bands shuffled
        band = ((torch.arange(n, device="cuda") // (1 << 20)) * 7919) % 100
        price = (band.float() + base / 100.0).contiguous(); del band
    else:
        price = base
    t = {"price": price}
    zm = ops.ZoneMap(price, "price")
    tb = timeit(lambda: ops.ZoneMap(price, "price").close(), iters=3)
    e, c = "(price[idx] * 0.9f)", "(price[idx] > 99.0f)"
    for mode, name in ((wc.COMPACT, "compact"), (wc.DENSE_ZERO, "dense_zero")):
        plain = timeit(lambda: ops.project_filter(t, e, c, mode, out=out, sync_count=False))
        pruned = timeit(lambda: ops.project_filter_pruned(t, e, c, [(zm, ">", 99.0)], mode, out=out, sync=False))
        _, cnt, live = ops.project_filter_pruned(t, e, c, [(zm, ">", 99.0)], mode, out=out)
        print(json.dumps({"layout": layout, "mode": name, "rows": n, "plain_ms": plain, "pruned_ms": pruned, "speedup": plain / pruned,
                          "zones_live": live, "zones": zm.nzones, "survivors": cnt, "zonemap_build_ms": tb}), flush=True)
    zm.close()
    if layout != "random": del price
