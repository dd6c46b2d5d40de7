#!/usr/bin/env python
"""Round-2 kernel sweeps on one B200 (run under gpurun); JSON lines to gpurun_out/r02_sweep_<what>.jsonl.

    python tools/sweep_r2.py compact [rows]   # variant 3 (L2-parked slabs) vs variant 4 (single pass, pipelined look-back)
    python tools/sweep_r2.py group1k [rows]   # warp-private accumulators: tag arbitration vs MATCH.ANY leader aggregation
    python tools/sweep_r2.py group10m [rows]  # direct-addressed table: slices, launch geometry
Every result is checked (survivor count / group sums) so a fast wrong variant cannot win.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

import torch  # noqa: E402

from warpdb_b200 import _core as wc, ops  # noqa: E402

PEAK = 6534.8
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def time_op(fn, iters=10, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


class Opts:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        for k, v in self.kv.items():
            wc.set_option(k, v)

    def __exit__(self, *a):
        for k in self.kv:
            wc.set_option(k, None)


def emit(f, rec):
    f.write(json.dumps(rec) + "\n")
    f.flush()
    print(json.dumps(rec), flush=True)


def sweep_compact(n):
    f = open(os.path.join(OUT, "r02_sweep_compact.jsonl"), "a")
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    expr, cond = "(price[idx] * 0.9f)", "(price[idx] > 20.0f)"
    cfgs = [{"compact.variant": 3}, {}, {"compact.auto": 2}, {"compact.variant": 5}]   # {} = the optimizer's own choice (feedback); auto 2 = chosen on the device
    best = {}
    for sel in (0.001, 0.01, 0.03, 0.06, 0.1, 0.5, 0.99):
        price = ops.synth_f32(n, 0xC0FFEE + 3, 0.0, 20.0 / (1.0 - sel))
        table = {"price": price}
        want = int((price > 20.0).sum().item())
        ref_head = (price[:1 << 22][price[:1 << 22] > 20.0] * 0.9)
        for cfg in cfgs:
            try:
                with Opts(**cfg):
                    _, cnt = ops.project_filter(table, expr, cond, wc.COMPACT, out=out)
                    ok = cnt == want and bool(torch.equal(out[:ref_head.numel()], ref_head))
                    ms = time_op(lambda: ops.project_filter(table, expr, cond, wc.COMPACT, out=out, sync_count=False))
                gbs = (4.0 + 4.0 * want / n) * n / (ms * 1e-3) / 1e9
                emit(f, {"sel": sel, "cfg": cfg, "ms": ms, "gbs": gbs, "frac": gbs / PEAK, "ok": ok, "rows": n})
                key = (sel, cfg.get("compact.variant", -1))
                if ok and (key not in best or ms < best[key][0]):
                    best[key] = (ms, cfg)
            except Exception as e:  # noqa: BLE001
                emit(f, {"sel": sel, "cfg": cfg, "error": str(e)[:200]})
        del price
    for k, v in sorted(best.items()):
        print("BEST", k, round(v[0], 3), v[1], flush=True)


def group_check(keys, vals, price, qty, G):
    ref = torch.zeros(G, dtype=torch.float64, device="cuda")
    for s in range(0, price.numel(), 1 << 27):
        ref.index_add_(0, qty[s:s + (1 << 27)].long(), price[s:s + (1 << 27)].double())
    return keys.numel() == G and bool(((vals.double() - ref.float().double()).abs() <= 1e-6 * ref.abs()).all().item())


def sweep_group1k(n):
    f = open(os.path.join(OUT, "r02_sweep_group1k.jsonl"), "a")
    price = ops.synth_f32(n, 0xC0FFEE + 4, 0.0, 100.0)
    for G in (2, 10, 50, 100, 200, 1000):
        qty = ops.synth_i32(n, 0xC0FFEE + 104, 0, G)
        table = {"price": price, "quantity": qty}
        cfgs = [{"group.lane_private": 0}, {"group.lane_private": 1}, {"group.lane_private": 1, "group.lane_min_warps": 4},
                {"group.lane_private": 1, "group.wp_unroll": 1}]
        for agg, needs in ((wc.SUM, wc.NEED_SUM), (wc.AVG, wc.NEED_SUM | wc.NEED_COUNT)):
            for cfg in cfgs:
                try:
                    with Opts(**cfg):
                        tab = ops.AggTable(0, 1024, needs)
                        tab.set_key_range(0, G - 1)

                        def run():
                            tab.reset()
                            tab.consume(table, "price[idx]", "quantity[idx]")
                        run()
                        o = tab.export(agg, wc.ORDER_KEY_ASC)
                        ok = group_check(o["keys"], o["sums"].float(), price, qty, G)
                        ms = time_op(run)
                        tab.close()
                    emit(f, {"G": G, "agg": agg, "cfg": cfg, "ms_reset_plus_consume": ms, "grows": n / ms / 1e6, "frac": 8.0 * n / (ms * 1e-3) / 1e9 / PEAK, "ok": ok, "rows": n})
                except Exception as e:  # noqa: BLE001
                    emit(f, {"G": G, "agg": agg, "cfg": cfg, "error": str(e)[:200]})
        del qty


def sweep_group10m(n):
    f = open(os.path.join(OUT, "r02_sweep_group10m.jsonl"), "a")
    price = ops.synth_f32(n, 0xC0FFEE + 4, 0.0, 100.0)
    G = 10_000_000
    qty = ops.synth_i32(n, 0xC0FFEE + 104, 0, G)
    table = {"price": price, "quantity": qty}
    cfgs = []
    for passes in (1, 2, 3, 4):
        for block, unroll, vec in ((512, 1, 8), (256, 2, 8), (1024, 1, 8), (512, 2, 8), (256, 1, 4)):
            for ctas in (8, 2):
                for hint in (2, 0):
                    cfgs.append({"group.dense_passes": passes, "group.dense_block": block, "group.dense_unroll": unroll, "group.dense_vec": vec,
                                 "group.ctas_per_sm": ctas, "group.dense_ld_hint": hint})
    for cfg in cfgs:
        if cfg["group.dense_passes"] != 2 and (cfg["group.dense_block"], cfg["group.dense_unroll"]) != (512, 1):
            continue
        try:
            with Opts(**cfg):
                tab = ops.AggTable(0, 1024, wc.NEED_SUM)
                tab.set_key_range(0, G - 1)

                def run():
                    tab.reset()
                    tab.consume(table, "price[idx]", "quantity[idx]")
                run()
                o = tab.export(wc.SUM, wc.ORDER_KEY_ASC)
                ok = group_check(o["keys"], o["sums"].float(), price, qty, G)
                ms = time_op(run, iters=5)
                tab.close()
            emit(f, {"G": G, "cfg": cfg, "ms_reset_plus_consume": ms, "grows": n / ms / 1e6, "frac": 8.0 * n / (ms * 1e-3) / 1e9 / PEAK, "ok": ok, "rows": n})
        except Exception as e:  # noqa: BLE001
            emit(f, {"G": G, "cfg": cfg, "error": str(e)[:200]})


def sweep_topk(n):
    """fused tail (one launch) vs scan + final + emit (three launches), at shard sizes of 1, 2, 4, 8 GPUs"""
    f = open(os.path.join(OUT, "r02_sweep_topk.jsonl"), "a")
    wc.set_udf_source("__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n")
    for rows in (n // 8, n // 4, n // 2, n):
        price = ops.synth_f32(rows, 0xC0FFEE + 5, 0.0, 1e6)
        table = {"price": price}
        want = torch.topk(price[:1 << 28] * 0.9, 5).values if rows <= 1 << 28 else None
        for fused, waves in ((1, 1), (1, 2)):
            with Opts(**{"topk.fused": fused, "topk.waves": waves}):
                got = ops.topk(table, "discount(price[idx], 0.9f)", None, None, True, 5)
                ms = time_op(lambda: ops.topk(table, "discount(price[idx], 0.9f)", None, None, True, 5), iters=20)
            ok = bool(torch.equal(got, torch.topk(torch.cat([torch.topk(price[s:s + (1 << 28)] * 0.9, 5).values for s in range(0, rows, 1 << 28)]), 5).values))
            emit(f, {"rows": rows, "fused": fused, "waves": waves, "ms_incl_host_sync": ms, "gbs": 4.0 * rows / (ms * 1e-3) / 1e9, "frac": 4.0 * rows / (ms * 1e-3) / 1e9 / PEAK, "ok": ok})
        del price


if __name__ == "__main__":
    what = sys.argv[1]
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
    wc.check(wc.lib().wdb_init(0))
    {"compact": sweep_compact, "group1k": sweep_group1k, "group10m": sweep_group10m, "topk": sweep_topk}[what](n)
