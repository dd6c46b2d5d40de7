#!/usr/bin/env python
"""Where does a GROUP BY step spend its time?  (reset / consume / export, per phase, CUDA events)"""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
wc.check(wc.lib().wdb_init(0))
price = ops.synth_f32(n, 0xC0FFEE + 4, 0.0, 100.0)

def ev():
    return torch.cuda.Event(enable_timing=True)

for G in ((1000, 10_000_000) if len(sys.argv) < 3 else tuple(int(float(x)) for x in sys.argv[2].split(","))):
    qty = ops.synth_i32(n, 0xC0FFEE + 104, 0, G)
    table = {"price": price, "quantity": qty}
    cfgs = [{}]
    if G <= 2000:
        cfgs += [{"group.smem_slots": sl, "group.block": b, "group.unroll": u, "group.vec": v}
                 for sl in (4096, 8192) for b in (512, 1024) for u, v in ((1, 4), (2, 4))]
    else:
        cfgs += [{"group.pass_bits": b} for b in (0, 1, 2, 3, 4, 5)]
        cfgs += [{"group.pass_bits": b, "group.vec": 8, "group.unroll": 1, "group.ld_hint": 2} for b in (2, 3, 4)]
        cfgs += [{"group.pass_bits": b, "group.block": 256} for b in (2, 3)]
    for cfg in cfgs:
        for k, v in {"group.smem_slots": -1, "group.vec": 4, "group.unroll": 2, "group.block": 512, "group.pass_bits": -1, "group.ld_hint": 0,
                     "group.wp_max_span": 0, "group.dense_max_span": 0}.items():   # this tool times the hash-table paths only
            wc.set_option(k, v)
        for k, v in cfg.items():
            wc.set_option(k, v)
        tab = ops.AggTable(0, G, wc.NEED_SUM)
        keys = torch.empty(G, dtype=torch.int32, device="cuda"); vals = torch.empty(G, dtype=torch.float32, device="cuda")
        res = {}
        for rep in range(3):
            e = [ev() for _ in range(4)]
            e[0].record(); tab.reset(); e[1].record(); tab.consume(table, "price[idx]", "quantity[idx]"); e[2].record()
            g = C.c_int64(0)
            wc.check(wc.lib().wdb_agg_export(tab.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream), wc.SUM, wc.ORDER_KEY_ASC,
                                             keys.data_ptr(), vals.data_ptr(), None, None, None, None, None, G, C.byref(g)))
            e[3].record(); torch.cuda.synchronize()
            res = {"reset_ms": e[0].elapsed_time(e[1]), "consume_ms": e[1].elapsed_time(e[2]), "export_ms": e[2].elapsed_time(e[3])}
        ok = abs(vals.double().sum().item() / price.double().sum().item() - 1) < 1e-6
        rec = {"G": G, "rows": n, "cfg": cfg, **res, "consume_grows_s": n / res["consume_ms"] / 1e6, "ok": ok}
        print(json.dumps(rec), flush=True)
        tab.close()
    del qty
