#!/usr/bin/env python
"""Kernel-variant sweeps on the B200 (run under gpurun).  Writes JSON lines to gpurun_out/.

    python tools/sweep.py project [rows]     # BASELINE config 2: price*quantity*1.08
    python tools/sweep.py compact [rows]     # BASELINE config 3: price*0.9 WHERE price>20 at 1/50/99 %
    python tools/sweep.py refjit             # does the reference's own NVRTC path run here?
"""
import ctypes as C
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def time_op(fn, iters=20, warmup=3):
    import torch
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


def sweep_project(n):
    import torch
    from warpdb_b200 import _core as wc, ops
    wc.check(wc.lib().wdb_init(0))
    price = ops.synth_f32(n, 0xC0FFEE + 2, 0.0, 100.0)
    qty = ops.synth_i32(n, 0xC0FFEE + 102, 1, 101)
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    table = {"price": price, "quantity": qty}
    expr = "((price[idx] * quantity[idx]) * 1.08f)"
    gb = 12.0 * n / 1e9
    f = open(os.path.join(OUT, "sweep_project.jsonl"), "a")
    pk = peak()

    def emit(rec):
        rec["gbs"] = gb / (rec["ms"] * 1e-3)
        rec["frac_measured_peak"] = rec["gbs"] / pk
        rec["rows"] = n
        f.write(json.dumps(rec) + "\n")
        f.flush()
        print(json.dumps(rec), flush=True)

    # yardsticks: torch elementwise on the same bytes and a plain copy
    ms = time_op(lambda: torch.mul(price, qty, out=out))
    emit({"name": "torch.mul(price,qty) [12 B/row]", "ms": ms})
    big_a = torch.empty(n * 3 // 2, dtype=torch.float32, device="cuda")
    big_b = torch.empty_like(big_a)
    ms = time_op(lambda: big_b.copy_(big_a))
    emit({"name": "torch copy_ 6B+6B/row", "ms": ms})
    del big_a, big_b

    def run(cfg):
        for k, v in cfg.items():
            wc.set_option("project." + k, v)
        try:
            ms = time_op(lambda: ops.project_filter(table, expr, None, wc.DENSE, out=out, sync_count=False))
            emit({"name": "wdb_project", "cfg": cfg, "ms": ms})
            return ms
        except Exception as e:  # noqa: BLE001
            print("FAILED", cfg, e, flush=True)
            torch.cuda.synchronize()
            return 1e9

    base = {"variant": 0, "vec": 8, "unroll": 4, "block": 256, "ld_hint": 0, "st_hint": 0, "ctas_per_sm": 8, "tile": 4096, "stages": 3}
    results = []
    for vec, unroll, block in itertools.product([4, 8], [1, 2, 4, 8], [128, 256, 512]):
        cfg = dict(base, vec=vec, unroll=unroll, block=block)
        results.append((run(cfg), cfg))
    best = min(results, key=lambda r: r[0])[1]
    for ld, st in itertools.product([0, 1, 2], [0, 1, 2, 3]):
        if best["vec"] == 4 and (ld == 2 or st == 3):
            continue
        cfg = dict(best, ld_hint=ld, st_hint=st)
        results.append((run(cfg), cfg))
    best = min(results, key=lambda r: r[0])[1]
    for ctas, unroll in itertools.product([1, 2, 4, 8], [2, 4]):
        cfg = dict(best, variant=1, ctas_per_sm=ctas, unroll=unroll)
        results.append((run(cfg), cfg))
    for tile, stages, block, ctas in itertools.product([2048, 4096, 8192], [2, 3, 4], [256, 512], [1, 2]):
        if tile % (4 * block) or (128 + stages * tile * 12) * ctas > 227 * 1024:
            continue
        cfg = dict(base, variant=2, vec=4, tile=tile, stages=stages, block=block, ctas_per_sm=ctas)
        results.append((run(cfg), cfg))
    best = min(results, key=lambda r: r[0])
    emit({"name": "BEST", "cfg": best[1], "ms": best[0]})


def sweep_compact(n):
    import torch
    from warpdb_b200 import _core as wc, ops
    wc.check(wc.lib().wdb_init(0))
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    f = open(os.path.join(OUT, "sweep_compact.jsonl"), "a")
    pk = peak()
    for sel in (0.01, 0.5, 0.99):
        hi = 20.0 / (1.0 - sel)
        price = ops.synth_f32(n, 0xC0FFEE + 3, 0.0, hi)
        table = {"price": price}
        _, cnt = ops.project_filter(table, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT, out=out)
        gb = (4.0 * n + 4.0 * cnt) / 1e9
        ms = time_op(lambda: torch.masked_select(price, price > 20.0), iters=5, warmup=1)
        rec = {"name": "torch masked_select", "sel": cnt / n, "ms": ms, "gbs": gb / (ms * 1e-3)}
        print(json.dumps(rec), flush=True)
        f.write(json.dumps(rec) + "\n")
        cfgs = []
        for unroll, block, lb in itertools.product([2, 4], [256, 512], [1, 4]):
            if block * 8 * unroll * 4 <= 46 * 1024:
                cfgs.append({"variant": 0, "vec": 8, "unroll": unroll, "block": block, "min_ctas": 1, "lookback": lb, "ctas_per_sm": 8})
        for vec, unroll, block, minc, lb in itertools.product([4, 8], [1, 2, 4], [256, 512], [1, 2, 3, 4], [1, 4]):
            rows = block * vec * unroll
            if rows < 2048 or 128 + rows * 12 > 200 * 1024 or minc * block > 2048:
                continue
            cfgs.append({"variant": 1, "vec": vec, "unroll": unroll, "block": block, "min_ctas": minc, "lookback": lb, "ctas_per_sm": 8})
        for vec, unroll, block in itertools.product([4, 8], [1, 2, 4], [128, 256, 512]):
            if block * vec * unroll * 4 <= 46 * 1024:
                cfgs.append({"variant": 2, "vec": vec, "unroll": unroll, "block": block, "min_ctas": 1, "lookback": 1, "ctas_per_sm": 8})
        for unroll, block, slab_m, minc in itertools.product([2, 4], [256, 512], [2, 4, 8], [1, 2, 3, 4]):
            if block * 8 * unroll * 4 <= 46 * 1024 and minc * block <= 2048:
                cfgs.append({"variant": 3, "vec": 8, "unroll": unroll, "block": block, "min_ctas": minc, "lookback": 1, "ctas_per_sm": 8, "slab_m": slab_m})
        if os.environ.get("SWEEP_ONLY_VARIANT"):
            cfgs = [c for c in cfgs if c["variant"] == int(os.environ["SWEEP_ONLY_VARIANT"])]
        for cfg in cfgs:
            for k, v in cfg.items():
                wc.set_option("compact." + k, v)
            try:
                ms = time_op(lambda: ops.project_filter(table, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT, out=out, sync_count=False), iters=10)
            except Exception as e:  # noqa: BLE001
                print("FAILED", cfg, e, flush=True)
                continue
            rec = {"name": "wdb_compact", "sel": cnt / n, "cfg": cfg, "ms": ms, "gbs": gb / (ms * 1e-3), "frac_measured_peak": gb / (ms * 1e-3) / pk, "rows": n}
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
            f.flush()
        # dense filter (reference semantics) on the same data
        for k, v in {"variant": 0, "vec": 8, "unroll": 4, "block": 256}.items():
            wc.set_option("project." + k, v)
        for mode, name in ((wc.DENSE, "dense-untouched"), (wc.DENSE_ZERO, "dense-zero")):
            ms = time_op(lambda: ops.project_filter(table, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", mode, out=out, sync_count=False), iters=10)
            gbm = (4.0 * n + 4.0 * (cnt if mode == wc.DENSE else n)) / 1e9
            rec = {"name": "wdb_filter " + name, "sel": cnt / n, "ms": ms, "gbs": gbm / (ms * 1e-3), "frac_measured_peak": gbm / (ms * 1e-3) / pk}
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
        del price


def probe_refjit(n=1 << 24):
    """Run the reference's jit_compile_and_launch (oracle/_ref/libref_jit.so) as shipped, then with
    the primary-context interposition, each in this process; print what happened."""
    import numpy as np
    import torch
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_jit.so"))
    lib.ref_last_kernel_ms.restype = C.c_float
    price = torch.rand(n, device="cuda") * 100
    qty = torch.randint(1, 101, (n,), device="cuda", dtype=torch.int32)
    out = torch.zeros(n, device="cuda")
    names = (C.c_char_p * 2)(b"price", b"quantity")
    dts = (C.c_int * 2)(2, 0)
    ptrs = (C.c_void_p * 2)(price.data_ptr(), qty.data_ptr())
    err = C.create_string_buffer(512)
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lib.ref_set_primary_ctx(mode)
    t0 = time.time()
    rc = lib.ref_jit_compile_and_launch(b"((price[idx] * quantity[idx]) * 1.08f)", b"", names, dts, ptrs, 2, n, C.c_void_p(out.data_ptr()), 0, err, 512)
    dt = time.time() - t0
    ok = False
    if rc == 0:
        torch.cuda.synchronize()
        ok = bool(torch.equal(out, (price * qty.float()) * 1.08))
    rec = {"name": "ref_jit_compile_and_launch", "primary_ctx_mode": mode, "rc": rc, "err": err.value.decode(), "call_s": dt,
           "kernel_ms": float(lib.ref_last_kernel_ms()), "result_matches": ok, "rows": n}
    print(json.dumps(rec), flush=True)
    open(os.path.join(OUT, "probe_refjit.jsonl"), "a").write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "project":
        sweep_project(int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000)
    elif what == "compact":
        sweep_compact(int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000)
    elif what == "refjit":
        probe_refjit()


def sweep_topk(n):
    import torch
    from warpdb_b200 import _core as wc, ops
    wc.check(wc.lib().wdb_init(0))
    wc.set_udf_source("__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n")
    price = ops.synth_f32(n, 0xC0FFEE + 5, 0.0, 1e6)
    table = {"price": price}
    f = open(os.path.join(OUT, "sweep_topk.jsonl"), "a")
    pk = peak()
    gb = 4.0 * n / 1e9
    ms = time_op(lambda: torch.topk(price, 5), iters=5, warmup=1)
    print(json.dumps({"name": "torch.topk", "ms": ms, "gbs": gb / (ms * 1e-3)}), flush=True)
    ms = time_op(lambda: torch.max(price), iters=5, warmup=1)
    print(json.dumps({"name": "torch.max (4 B/row streaming reduce)", "ms": ms, "gbs": gb / (ms * 1e-3)}), flush=True)
    for vec, unroll, block, ctas in itertools.product([4, 8], [1, 2, 4], [256, 512], [2, 4, 8]):
        cfg = {"vec": vec, "unroll": unroll, "block": block, "ctas_per_sm": ctas}
        for k, v in cfg.items():
            wc.set_option("topk." + k, v)
        try:
            ms = time_op(lambda: ops.topk(table, "discount(price[idx], 0.9f)", None, None, True, 5), iters=10)
        except Exception as e:  # noqa: BLE001
            print("FAILED", cfg, e, flush=True)
            continue
        rec = {"name": "wdb_topk k=5", "cfg": cfg, "ms": ms, "gbs": gb / (ms * 1e-3), "frac_measured_peak": gb / (ms * 1e-3) / pk, "rows": n}
        print(json.dumps(rec), flush=True)
        f.write(json.dumps(rec) + "\n")
        f.flush()


if __name__ == "__main__" and sys.argv[1] == "topk":
    sweep_topk(int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000_000)
