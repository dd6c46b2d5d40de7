#!/usr/bin/env python
"""v3 compaction (slab parked in L2): does the slab stay in the L2?  Sweep of the residency hints,
slab size and CTAs per SM at 1/50/99 % selectivity (BASELINE config 3, 1e9 rows)."""
import sys, os, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
wc.check(wc.lib().wdb_init(0))
out = torch.empty(n, dtype=torch.float32, device="cuda")
def time_op(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
KEYS = ( "l2_hints", "slab_m", "min_ctas", "ctas_per_sm", "unroll", "block", "variant")
for sel in (0.01, 0.5, 0.99):
    price = ops.synth_f32(n, 0xC0FFEE + 3, 0.0, 20.0 / (1.0 - sel))
    table = {"price": price}
    _, cnt = ops.project_filter(table, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT, out=out)
    gb = (4.0 * n + 4.0 * cnt) / 1e9
    cfgs = [{"l2_hints": 0}, {"l2_hints": 1}, {"variant": 2}]
    cfgs += [{"l2_hints": h, "slab_m": m, "min_ctas": c} for h, m, c in itertools.product((0, 1), (2, 4, 8), (2, 3, 4)) if not (m == 4 and c == 4)]
    cfgs += [{"l2_hints": 1, "slab_m": m, "min_ctas": c, "unroll": 2} for m, c in itertools.product((2, 4, 8), (4, 6, 8))]
    if os.environ.get("QUICK"):
        cfgs = cfgs[:3]
    if os.environ.get("SMALLBLOCK"):
        cfgs = [{}] + [{"block": b, "min_ctas": c, "slab_m": m, "ctas_per_sm": 16} for b, c, m in ((128, 8, 4), (128, 8, 8), (128, 6, 8), (192, 5, 4), (192, 5, 6), (384, 2, 4), (384, 2, 3), (320, 3, 4), (256, 4, 3), (256, 4, 5), (256, 4, 6))]
    for cfg in cfgs:
        for k in KEYS:
            wc.set_option("compact." + k, None)
        for k, v in cfg.items():
            wc.set_option("compact." + k, v)
        try:
            ms = time_op(lambda: ops.project_filter(table, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT, out=out, sync_count=False))
        except Exception as e:  # noqa: BLE001
            print("FAILED", cfg, e, flush=True); continue
        print(json.dumps({"sel": round(cnt / n, 4), "cfg": cfg, "ms": round(ms, 4), "gbs": round(gb / ms * 1e3, 1)}), flush=True)
    del price, table
