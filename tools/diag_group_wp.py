#!/usr/bin/env python
"""Small-cardinality GROUP BY: warp-private direct-indexed kernel (wdb_group_wp, chosen from the key
column's min/max) vs the shared-atomic kernel (wdb_group), geometry sweep at 1 K keys."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
wc.check(wc.lib().wdb_init(0))
price = ops.synth_f32(n, 0xC0FFEE + 4, 0.0, 100.0)
ref_total = price.double().sum().item()
KEYS = ("group.wp_max_span", "group.wp_ilp", "group.wp_unroll", "group.wp_warps", "group.wp_vec", "group.ctas_per_sm", "group.smem_slots", "group.ld_hint")

def run(G, cfg, agg=wc.SUM, needs=wc.NEED_SUM, note=""):
    for k in KEYS:
        wc.set_option(k, None)
    for k, v in cfg.items():
        wc.set_option(k, v)
    tab = ops.AggTable(0, G, needs)
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tab.reset(); e0.record(); tab.consume(table, "price[idx]", "quantity[idx]"); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    spilled = tab.spilled()
    out = tab.export(agg, wc.ORDER_KEY_ASC)
    ok = abs(out["sums"].sum().item() / ref_total - 1) < 1e-9
    print(json.dumps({"G": G, "note": note, "rows": n, "cfg": cfg, "consume_ms": round(best, 4), "grows_s": round(n / best / 1e6, 1),
                      "groups": int(out["keys"].numel()), "spilled": spilled, "ok": ok}), flush=True)
    tab.close()

for G in (1000, 500, 100, 8, 1500, 2000, 3000):
    qty = ops.synth_i32(n, 0xC0FFEE + 104, 0, G)
    table = {"price": price, "quantity": qty}
    run(G, {"group.wp_max_span": 0}, note="atomic")
    run(G, {"group.wp_max_span": 4096}, note="wp")
    if G == 1000:
        for cfg in [{"group.wp_ilp": i, "group.wp_unroll": u} for i in (1, 2) for u in (1, 2, 4)] + [{"group.wp_warps": w} for w in (8, 12, 14)] + \
                   [{"group.wp_vec": 4, "group.wp_unroll": 4}, {"group.ld_hint": 2}]:
            run(G, cfg, note="wp")
        run(G, {}, agg=wc.AVG, needs=wc.NEED_SUM | wc.NEED_COUNT, note="wp avg")
        run(G, {"group.wp_max_span": 0}, agg=wc.AVG, needs=wc.NEED_SUM | wc.NEED_COUNT, note="atomic avg")
    del qty, table
