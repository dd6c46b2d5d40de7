#!/usr/bin/env python
"""bench.py -- headline benchmark of the WarpDB hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of synthetic, HBM-resident columns.
Default workload = BASELINE.json configs[1]: pure projection "price * quantity * 1.08" over
1e9 rows (price float32, quantity int32; 12 algorithmic bytes per row) per GPU (weak scaling:
row-range shards, no data-path collective).  Other workloads (--workload) are the remaining
BASELINE configs; they are parity-test cases with a measurement, not the headline line.

Prints ONE JSON line (rank 0).  `value` is rows/s over all GPUs with inputs resident in HBM;
`e2e` is the same metric through the host-buffer entry point (H2D and D2H inside the timed region);
`roofline` compares the dominant kernel with the measured HBM copy peak (MEASURED_PEAKS.json);
`cpu_baseline` is the CPU oracle (scalar AST evaluation, all host threads) on a bounded sample;
`ref_nvrtc` is the reference's own NVRTC kernel (oracle/_ref/libref_jit.so) on the same arrays.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rows per GPU, description)
    "projection": (1_000_000_000, "price * quantity * 1.08 over 1e9 rows/GPU (f32 price, i32 quantity): BASELINE configs[1]"),
    "filter1": (1_000_000_000, "price * 0.9 WHERE price > 20, 1% selectivity, stable compaction: BASELINE configs[2] (per-GPU shard of 4e9/4)"),
    "filter50": (1_000_000_000, "price * 0.9 WHERE price > 20, 50% selectivity, stable compaction"),
    "filter99": (1_000_000_000, "price * 0.9 WHERE price > 20, 99% selectivity, stable compaction"),
    "group1k": (1_000_000_000, "SELECT SUM(price) FROM t GROUP BY quantity, 1K keys: BASELINE configs[3] (per-GPU shard)"),
    "group10m": (1_000_000_000, "SELECT SUM(price) FROM t GROUP BY quantity, 10M keys"),
    "topk5": (2_000_000_000, "SELECT discount(price, 0.9) FROM t ORDER BY discount(price, 0.9) DESC LIMIT 5: BASELINE configs[4] (per-GPU shard)"),
}
UDF = "__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n"


def metric_name():
    """BASELINE.json's metric string (value is its rows/s part; GB/s is reported under roofline.achieved)."""
    try:
        return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    except Exception:  # noqa: BLE001
        return "rows/sec & achieved HBM GB/s (filter/project/agg) at 1/2/4/8 B200"


def measured_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    return rank, world, local


def barrier(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------
# workloads: setup(rank, world, local) -> dict(step=callable, rows, bytes_per_row, kernel, ...)
# ------------------------------------------------------------------------------------------------
def make_workload(name, rows, rank, world, local):
    import torch
    from warpdb_b200 import _core as wc, ops
    wc.check(wc.lib().wdb_init(local))
    wc.set_udf_source(UDF)
    row0 = rank * rows                       # contiguous row-range shard of the global table (multi_gpu_utils.cpp:24-31)
    seed = 0xC0FFEE
    if name == "projection":
        price = ops.synth_f32(rows, seed + 2, 0.0, 100.0, row0, local)
        qty = ops.synth_i32(rows, seed + 102, 1, 101, row0, local)
        out = torch.empty(rows, dtype=torch.float32, device=f"cuda:{local}")
        table = {"price": price, "quantity": qty}
        expr = "((price[idx] * quantity[idx]) * 1.08f)"
        return dict(step=lambda: ops.project_filter(table, expr, None, wc.DENSE, out=out, sync_count=False),
                    bytes_per_row=12.0, kernel="wdb_project", table=table, expr=expr, cond=None, query="price * quantity * 1.08",
                    check=lambda: bool(torch.equal(out[:1 << 20], ((price[:1 << 20] * qty[:1 << 20].float()) * 1.08))))
    if name.startswith("filter"):
        sel = {"filter1": 0.01, "filter50": 0.5, "filter99": 0.99}[name]
        price = ops.synth_f32(rows, seed + 3, 0.0, 20.0 / (1.0 - sel), row0, local)
        out = torch.empty(rows, dtype=torch.float32, device=f"cuda:{local}")
        table = {"price": price}
        expr, cond = "(price[idx] * 0.9f)", "(price[idx] > 20.0f)"
        _, cnt = ops.project_filter(table, expr, cond, wc.COMPACT, out=out)
        return dict(step=lambda: ops.project_filter(table, expr, cond, wc.COMPACT, out=out, sync_count=False),
                    bytes_per_row=4.0 + 4.0 * cnt / rows, kernel="wdb_compact_l2", table=table, expr=expr, cond=cond,
                    query="price * 0.9 WHERE price > 20", selectivity=cnt / rows,
                    check=lambda: bool(torch.equal(out[:cnt][:1 << 20], (price[price > 20.0][:1 << 20] * 0.9))))
    if name.startswith("group"):
        from warpdb_b200.sharded import ShardedDB
        G = 1000 if name == "group1k" else 10_000_000
        price = ops.synth_f32(rows, seed + 4, 0.0, 100.0, row0, local)
        qty = ops.synth_i32(rows, seed + 104, 0, G, row0, local)
        table = {"price": price, "quantity": qty}
        db = ShardedDB(table, rows * world, rank, world)
        res = {}

        def step():   # per-GPU partial aggregation + (N > 1) NCCL merge of the partials; every rank ends with the final groups
            res["g"] = db.group_agg("price[idx]", "quantity[idx]", None, wc.SUM, wc.ORDER_KEY_ASC, expected_groups=G)

        def check():
            tot = price.double().sum()
            if world > 1:
                import torch.distributed as dist
                dist.all_reduce(tot)
            g = res["g"]
            return bool(g["keys"].numel() == G and abs(g["vals"].double().sum().item() / tot.item() - 1.0) < 1e-6)
        def kernel_ms(reps=5):   # the dominant kernel alone (consume), CUDA events on the launching stream
            tab = ops.AggTable(local, G, wc.NEED_SUM)
            tab.set_key_range(0, G - 1)
            best = []
            for _ in range(reps + 1):
                tab.reset()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); tab.consume(table, "price[idx]", "quantity[idx]"); e1.record(); torch.cuda.synchronize()
                best.append(e0.elapsed_time(e1))
            tab.close()
            return sum(best[1:]) / reps
        return dict(step=step, bytes_per_row=8.0, kernel="wdb_group_wp" if G <= 4096 else "wdb_group", table=table, kernel_ms=kernel_ms,
                    query="SELECT SUM(price) FROM t GROUP BY quantity",
                    groups=G, check=check)
    if name == "topk5":
        from warpdb_b200.sharded import ShardedDB
        price = ops.synth_f32(rows, seed + 5, 0.0, 1e6, row0, local)
        table = {"price": price}
        db = ShardedDB(table, rows * world, rank, world)
        res = {}

        def step():   # per-GPU top-k + (N > 1) all_gather of k candidates per GPU and a final merge
            res["top"] = db.topk("discount(price[idx], 0.9f)", None, None, True, 5)

        def check():
            loc = torch.topk(price * 0.9, 5).values
            if world > 1:
                import torch.distributed as dist
                allc = [torch.empty_like(loc) for _ in range(world)]
                dist.all_gather(allc, loc)
                loc = torch.topk(torch.cat(allc), 5).values
            return bool(torch.equal(res["top"], loc))
        return dict(step=step, bytes_per_row=4.0, kernel="wdb_topk_scan", table=table,
                    query="SELECT discount(price, 0.9) FROM t ORDER BY discount(price, 0.9) DESC LIMIT 5", check=check)
    raise SystemExit(f"unknown workload {name}")


def time_steps(step, steps, warmup, world):
    import torch
    for _ in range(warmup):
        step()
    barrier(world)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    barrier(world)
    return max_over_ranks(a.elapsed_time(b), world)


def e2e_projection(rows, steps, warmup, world, local):
    """Same metric through the host-buffer entry point (wdb_multi_project_filter_host: the
    run_multi_gpu_jit_host replacement): pinned host columns in, host floats out, every step."""
    import torch
    from warpdb_b200 import _core as wc
    from warpdb_b200 import ops
    n = rows
    # synthesise on the device once, keep pinned host copies as "the user's data"
    hp = torch.empty(n, dtype=torch.float32, pin_memory=True)
    hq = torch.empty(n, dtype=torch.int32, pin_memory=True)
    ho = torch.empty(n, dtype=torch.float32, pin_memory=True)
    chunk = 1 << 27
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        hp[s:s + m].copy_(ops.synth_f32(m, 0xC0FFEE + 2, 0.0, 100.0, s, local))
        hq[s:s + m].copy_(ops.synth_i32(m, 0xC0FFEE + 102, 1, 101, s, local))
    torch.cuda.synchronize()
    cols, nc = wc.make_cols([("price", wc.FLOAT32, hp.data_ptr(), n), ("quantity", wc.INT32, hq.data_ptr(), n)])
    cnt = C.c_int64(0)
    devs = (C.c_int * 1)(local)

    def step():
        wc.check(wc.lib().wdb_multi_project_filter_host(1, devs, cols, nc, b"((price[idx] * quantity[idx]) * 1.08f)", b"",
                                                        ho.data_ptr(), n, wc.DENSE_ZERO, C.byref(cnt)))
    for _ in range(warmup):
        step()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    barrier(world)
    dt = max_over_ranks(time.perf_counter() - t0, world)
    ok = bool(torch.equal(ho[:1 << 20], (hp[:1 << 20] * hq[:1 << 20].float()) * 1.08))
    return dict(value=world * n * steps / dt, unit="rows/s", h2d_bytes_per_step=8 * n, d2h_bytes_per_step=4 * n,
                steps=steps, ms_per_step=dt / steps * 1e3, result_checked=ok,
                api="wdb_multi_project_filter_host (run_multi_gpu_jit_host replacement), pinned host buffers")


def cpu_baseline_projection(sample_rows, steps=1):
    """Scalar AST evaluation on the host cores (oracle port of src/warpdb.cpp:128-151 semantics)."""
    from oracle import pyoracle as orc
    cores = os.cpu_count() or 1
    table = {"price": orc.synth_f32(sample_rows, 0xC0FFEE + 2, 0.0, 100.0), "quantity": orc.synth_i32(sample_rows, 0xC0FFEE + 102, 1, 101)}
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.project_filter("price * quantity * 1.08", None, table, nthreads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return dict(value=sample_rows / best, unit="rows/s", cores=cores, kind="port",
                sample=f"{sample_rows} rows of the same synthetic columns, oracle/wdb_oracle.c orc_project_filter with {cores} pthreads")


def ref_nvrtc_projection(table, rows, local):
    """The reference's own NVRTC kernel (src/jit.cpp user_kernel, built unmodified into
    oracle/_ref/libref_jit.so) on the same device arrays."""
    import torch
    path = os.path.join(ROOT, "oracle", "_ref", "libref_jit.so")
    if not os.path.exists(path):
        return {"unavailable": "oracle/_ref/libref_jit.so not built"}
    try:
        lib = C.CDLL(path)
    except OSError as e:
        return {"unavailable": str(e)}
    lib.ref_last_kernel_ms.restype = C.c_float
    lib.ref_set_primary_ctx(1)
    out = torch.empty(rows, dtype=torch.float32, device=f"cuda:{local}")
    names = (C.c_char_p * 2)(b"price", b"quantity")
    dts = (C.c_int * 2)(2, 0)
    ptrs = (C.c_void_p * 2)(table["price"].data_ptr(), table["quantity"].data_ptr())
    err = C.create_string_buffer(512)
    kms, calls = [], []
    for _ in range(4):
        t0 = time.perf_counter()
        rc = lib.ref_jit_compile_and_launch(b"((price[idx] * quantity[idx]) * 1.08f)", b"", names, dts, ptrs, 2, rows,
                                            C.c_void_p(out.data_ptr()), local, err, 512)
        calls.append(time.perf_counter() - t0)
        if rc != 0:
            return {"unavailable": err.value.decode()}
        kms.append(float(lib.ref_last_kernel_ms()))
    torch.cuda.synchronize()
    ok = bool(torch.equal(out[:1 << 20], (table["price"][:1 << 20] * table["quantity"][:1 << 20].float()) * 1.08))
    k = min(kms[1:])
    return dict(kernel_ms=k, kernel_rows_per_s=rows / (k * 1e-3), kernel_gbs=12.0 * rows / (k * 1e-3) / 1e9,
                call_ms=min(calls[1:]) * 1e3, call_rows_per_s=rows / min(calls[1:]), result_checked=ok,
                note="jit_compile_and_launch re-compiles with NVRTC on every call (src/jit.cpp:98-136); kernel_ms is the "
                     "user_kernel launch alone (events around cuLaunchKernel), call_ms the whole call; primary context interposed")


def run_ours(args):
    import torch
    rank, world, local = dist_setup(args.gpus)
    from warpdb_b200 import _core as wc
    rows = args.rows or WORKLOADS[args.workload][0]
    w = make_workload(args.workload, rows, rank, world, local)
    peak, peak_src = measured_peak()
    sampler = ClockSampler(local)
    # untimed: compile + warm
    s0 = wc.stats()
    w["step"]()
    torch.cuda.synchronize()
    compile_ms = wc.stats()["last_compile_ms"]
    l0 = wc.stats()["launches"]
    if rank == 0:
        sampler.start()
    ms = time_steps(w["step"], args.steps, args.warmup, world)
    clocks = sampler.stop() if rank == 0 else None
    launches = wc.stats()["launches"] - l0 - 0
    launches_timed = launches * args.steps // (args.steps + args.warmup) if (args.steps + args.warmup) else 0
    ok = w["check"]()
    value = world * rows * args.steps / (ms * 1e-3)
    ms_per_step = ms / args.steps
    # roofline: the dominant kernel's own launch time where a step launches more than one kernel
    # (GROUP BY: consume vs reset / export / merge), else the step time
    kernel_ms = w["kernel_ms"]() if "kernel_ms" in w else ms_per_step
    achieved = w["bytes_per_row"] * rows / (kernel_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
    # (profiles/traffic.json holds bytes per row measured on 2^28-row columns; scaled to this launch)
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["kernels"].get(w["kernel"])
        if t and args.workload in ("projection", "topk5", "group1k", "filter50"):   # the captures were taken on these configurations
            traffic = t["dram_bytes_per_row"] * rows
    except Exception:  # noqa: BLE001
        pass
    line = {
        "metric": metric_name(), "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (counter-based generator, identical on the CPU oracle)",
        "config": {"workload": args.workload, "query": w["query"], "description": WORKLOADS[args.workload][1], "rows_per_gpu": rows,
                   "sharding": "contiguous row ranges, one process per GPU, no data-path collective" if args.workload in ("projection",) or args.workload.startswith("filter") else "contiguous row ranges, one process per GPU; partial aggregates / top-k candidates merged with NCCL inside the step",
                   "l2": "inputs (>= 4 GB per column) are far larger than the 126 MB L2; no flush needed",
                   "algorithmic_bytes_per_row": w["bytes_per_row"], "result_checked": ok},
        "gbs_per_gpu": achieved,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": w["kernel"], "peak_source": peak_src, "frac_of_8TBs_spec": achieved / 8000.0,
                     "algorithmic_bytes_per_launch": w["bytes_per_row"] * rows, "kernel_ms": kernel_ms},
        "gpu_launches": launches_timed, "clocks": clocks, "nvrtc_compile_ms_untimed": compile_ms,
    }
    for k in ("selectivity", "groups"):
        if k in w:
            line["config"][k] = w[k]
    if args.workload == "projection":
        if not args.no_e2e:
            del w["step"]
            line["e2e"] = e2e_projection(rows, max(1, min(args.steps, args.e2e_steps)), 1, world, local)
        if rank == 0 and world == 1:
            if not args.no_ref:
                line["ref_nvrtc"] = ref_nvrtc_projection(w["table"], rows, local)
            if not args.no_cpu:
                line["cpu_baseline"] = cpu_baseline_projection(args.cpu_rows)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the
    reference's GPU-less evaluator eval_node, src/warpdb.cpp:128-151, is part of the un-compilable
    query_sql), all host threads, bounded sample per step.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as orc
    cores = os.cpu_count() or 1
    sample = args.cpu_rows
    table = {"price": orc.synth_f32(sample, 0xC0FFEE + 2, 0.0, 100.0), "quantity": orc.synth_i32(sample, 0xC0FFEE + 102, 1, 101)}
    for _ in range(args.warmup if args.warmup < 2 else 1):
        orc.project_filter("price * quantity * 1.08", None, table, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.project_filter("price * quantity * 1.08", None, table, nthreads=cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    cb = dict(value=value, unit="rows/s", cores=cores, kind="port",
              sample=f"{sample} rows per step of the same synthetic columns (bounded sample of the 1e9-row workload), {cores} pthreads")
    print(json.dumps({
        "impl": "reference", "metric": metric_name(), "value": value, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (counter-based generator)",
        "config": {"workload": "projection", "query": "price * quantity * 1.08", "description": WORKLOADS["projection"][1],
                   "rows_per_step": sample},
        "cpu_baseline": cb, "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="projection", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: the workload's)")
    ap.add_argument("--cpu-rows", type=int, default=1 << 28, help="bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
