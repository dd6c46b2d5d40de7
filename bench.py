#!/usr/bin/env python
"""bench.py -- headline benchmark of the WarpDB hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workloads LIST]

One "step" = one pass of the hot path over one batch of synthetic, HBM-resident columns.

Headline (`value`, `roofline`, `e2e`): BASELINE.json configs[1], pure projection
"price * quantity * 1.08" over 1e9 rows PER GPU (price float32, quantity int32; 12 algorithmic bytes
per row); row-range shards, no data-path collective, weak scaling.

`workloads`: the remaining BASELINE configs in the SAME JSON line, each at its BASELINE size for the
whole job (shards of it for N > 1, i.e. strong scaling) with the cross-GPU merge INSIDE the step:
    filter1 / filter50 / filter99   "price * 0.9 WHERE price > 20", stable compaction, 4e9 rows   (configs[2])
    group1k / group10m              "SELECT SUM(price) FROM t GROUP BY quantity", 2e9 rows       (configs[3])
    topk5                           "SELECT discount(price, 0.9) ... ORDER BY ... DESC LIMIT 5", 8e9 rows (configs[4])
Every sub-record carries ms_per_step (CUDA events, max over ranks), rows/s of the whole job, the
roofline of its dominant kernel, merge_ms (step minus the same step without collectives), the
collectives used, clocks sampled during its timed region and a result check.

Prints ONE JSON line (rank 0).  `e2e` is the projection through the host-buffer entry point
(wdb_multi_project_filter_host: H2D and D2H inside the timed region); `cpu_baseline` is the CPU
oracle on a bounded sample; `ref_nvrtc` is the reference's own NVRTC kernel
(oracle/_ref/libref_jit.so) on the same arrays.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALL_WORKLOADS = ["filter1", "filter50", "filter99", "group1k", "group10m", "topk5"]
TOTAL_ROWS = {"projection": None, "filter1": 4_000_000_000, "filter50": 4_000_000_000, "filter99": 4_000_000_000,
              "group1k": 2_000_000_000, "group10m": 2_000_000_000, "topk5": 8_000_000_000}
PROJECTION_ROWS_PER_GPU = 1_000_000_000
UDF = "__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n"
SEED = 0xC0FFEE


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


def traffic_per_row(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per row and LAUNCH of the committed `ncu --set full` capture
    of this kernel on this workload at this size (profiles/traffic.json <- profiles/r02_ncu_full_summary.csv);
    scaled by the rows of the launch here -- it is NOT measured in this run."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        k = t["kernels"].get(f"{kernel}:{workload}")
        return (k["dram_bytes_per_row"], t.get("source", "profiles/traffic.json")) if k else (None, None)
    except Exception:  # noqa: BLE001
        return None, None


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled in-process through NVML every ~2 ms while a
    timed region is open (`with sampler.region(): ...`).  nvidia-smi -lms needs ~100 ms to produce
    its first line, longer than a 30 ms timed region; NVML does not."""
    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, torch_device_index):
        self.ok = False
        self.samples = []
        self.on = False
        self.stop_flag = False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_device_index)   # no CUDA_VISIBLE_DEVICES remapping: same order
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            if self.on:
                try:
                    mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    try:
                        reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:  # noqa: BLE001
                        reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.samples.append((float(mhz), int(reasons)))
                except Exception:  # noqa: BLE001
                    pass
                time.sleep(0.001)
            else:
                time.sleep(0.0005)

    class _Region:
        def __init__(self, s):
            self.s = s

        def __enter__(self):
            self.s.samples = []
            self.s.on = True

        def __exit__(self, *a):
            self.s.on = False

    def region(self):
        return ClockSampler._Region(self)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")], "samples": 0}
        sm = sorted(m for m, _ in self.samples)
        reasons = set()
        for _, r in self.samples:
            for name, bit in self.BAD.items():
                if r & bit:
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(sm),
                "source": "NVML in-process, ~2 ms period, timed region only"}

    def close(self):
        self.stop_flag = True


def dist_setup():
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


def time_steps(step, steps, warmup, world, sampler=None):
    """W untimed steps, then exactly K steps between barrier + synchronize on both sides, CUDA events on
    the launching stream, max over ranks.  Returns (total ms, launches of our kernels inside the timed region, clocks)."""
    import torch
    from warpdb_b200 import _core as wc
    for _ in range(warmup):
        step()
    barrier(world)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = wc.stats()["launches"]
    if sampler:
        sampler.on = True
        sampler.samples = []
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    if sampler:
        sampler.on = False
    launches = wc.stats()["launches"] - l0
    clocks = sampler.summary() if sampler else None
    barrier(world)
    return max_over_ranks(a.elapsed_time(b), world), launches, clocks


def roofline(kernel, workload, bytes_per_row, rows, kernel_ms, peak, peak_src, note=None):
    achieved = bytes_per_row * rows / (kernel_ms * 1e-3) / 1e9
    tpr, tsrc = traffic_per_row(kernel, workload)
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": tpr * rows if tpr else None,
         "traffic_source": (f"{tsrc}: {tpr:.3f} DRAM bytes per row and launch x rows of this launch (ncu capture of this workload, not measured in this run)" if tpr else None),
         "kernel": kernel, "kernel_ms": kernel_ms, "peak_source": peak_src, "frac_of_8TBs_spec": achieved / 8000.0,
         "algorithmic_bytes_per_row": bytes_per_row, "algorithmic_bytes_per_launch": bytes_per_row * rows}
    if note:
        r["note"] = note
    return r


def shard_rows(total, world, rank):
    chunk = (total + world - 1) // world
    s = min(rank * chunk, total)
    return s, min(s + chunk, total) - s


# ------------------------------------------------------------------------------------------------
# sub-workloads (BASELINE configs[2..4]); every function returns the record of its workload
# ------------------------------------------------------------------------------------------------
def run_filter(name, args, rank, world, local, comm, comm1, sampler, peak, peak_src):
    import torch
    from warpdb_b200 import _core as wc, ops
    sel = {"filter1": 0.01, "filter50": 0.5, "filter99": 0.99}[name]
    total = args.rows_total or TOTAL_ROWS[name]
    row0, rows = shard_rows(total, world, rank)
    price = ops.synth_f32(rows, SEED + 3, 0.0, 20.0 / (1.0 - sel), row0, local)
    out = torch.empty(rows, dtype=torch.float32, device=f"cuda:{local}")
    table = {"price": price}
    expr, cond = "(price[idx] * 0.9f)", "(price[idx] > 20.0f)"
    _, (cnt, off, tot) = comm.project_filter(table, expr, cond, wc.COMPACT, out=out)
    ms, launches, clocks = time_steps(lambda: comm.project_filter(table, expr, cond, wc.COMPACT, out=out, sync=False), args.steps, args.warmup, world, sampler)
    ms_local = ms
    if world > 1:
        ms_local, _, _ = time_steps(lambda: ops.project_filter(table, expr, cond, wc.COMPACT, out=out, sync_count=False), args.steps, args.warmup, world)
    # result: survivor count and the first 2^20 survivors of this shard against torch; global offsets add up
    m = min(rows, 1 << 24)
    head = price[:m]
    want = (head[head > 20.0] * 0.9)[:1 << 20]
    ok = cnt == int((price > 20.0).sum().item()) and bool(torch.equal(out[:want.numel()], want))
    if world > 1:
        import torch.distributed as dist
        c = torch.tensor([cnt, off], dtype=torch.int64, device="cuda")
        allc = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(allc, c)
        cs = [int(x[0]) for x in allc]
        ok = ok and tot == sum(cs) and off == sum(cs[:rank])
        okt = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(okt.item())
    s = tot / total
    bpr = 4.0 + 4.0 * cnt / max(rows, 1)
    per = ms / args.steps
    # which kernels the optimizer ran (selectivity feedback from the previous call of this query shape): 5 = staged two-pass
    # (wdb_count_stage + scan + wdb_gather_stage) for selective filters, 3 = L2-parked slabs (wdb_compact_l2)
    lv = C.c_int64(0)
    variant = int(lv.value) if wc.lib().wdb_get_option(b"compact.last_variant", C.byref(lv)) == 0 else 3
    kernel = "wdb_count_stage" if variant == 5 else "wdb_compact_l2"
    rec = {"query": "SELECT price * 0.9 FROM t WHERE price > 20 (stable compaction)", "baseline_config": "configs[2]", "selectivity": s,
           "rows_total": total, "rows_per_gpu": rows, "scaling": "strong" if world > 1 else "n/a", "ms_per_step": per,
           "value": total / (per * 1e-3), "unit": "rows/s",
           "roofline": roofline(kernel, name, bpr, rows, ms_local / args.steps, peak, peak_src,
                                "kernel_ms = the local compaction call timed with CUDA events: " +
                                ("wdb_count_stage (one streaming pass that parks each chunk's survivors) + group scan + wdb_gather_stage"
                                 if variant == 5 else "wdb_compact_l2 (one HBM pass, slab re-read from the L2)")),
           "optimizer": {"compact_variant": variant, "chosen_by": "survivor count of the previous call of this query shape (pinned-slot feedback, no host sync)"},
           "merge_ms": max(per - ms_local / args.steps, 0.0), "collectives": ["ncclAllGather(1 x int64 per rank: survivor counts)"] if world > 1 else [],
           "gpu_launches": launches, "clocks": clocks, "result_checked": ok}
    del price, out
    return rec


def run_group(name, args, rank, world, local, comm, comm1, sampler, peak, peak_src):
    import torch
    from warpdb_b200 import _core as wc, ops
    G = 1000 if name == "group1k" else 10_000_000
    total = args.rows_total or TOTAL_ROWS[name]
    row0, rows = shard_rows(total, world, rank)
    dev = f"cuda:{local}"
    price = ops.synth_f32(rows, SEED + 4, 0.0, 100.0, row0, local)
    qty = ops.synth_i32(rows, SEED + 104, 0, G, row0, local)
    table = {"price": price, "quantity": qty}
    # optimizer statistics of the key column (gathered once at load time, like WarpDB::query_sql caches them)
    lo, hi = ops.column_minmax(qty, "quantity")
    ends = torch.tensor([lo, -hi], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ends, op=dist.ReduceOp.MIN)
    rng = (int(ends[0].item()), int(-ends[1].item()))
    bufs = (torch.empty(G + 16, dtype=torch.int32, device=dev), torch.empty(G + 16, dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.int64, device=dev))

    def step(c=comm):   # reset + local aggregation + NCCL all-reduce of the partial tables + ordered export; no host synchronisation
        c.group_agg(table, "price[idx]", "quantity[idx]", None, wc.SUM, wc.ORDER_KEY_ASC, row_base=row0, expected_groups=G, key_range=rng, out=bufs, sync=False)
    ms, launches, clocks = time_steps(step, args.steps, args.warmup, world, sampler)
    keys, vals, groups = bufs
    torch.cuda.synchronize()
    g = int(groups.item())
    got_vals = vals[:g].double().clone()
    got_keys = keys[:g].clone()
    ms_local = ms
    if world > 1:
        ms_local, _, _ = time_steps(lambda: step(comm1), args.steps, args.warmup, world)

    # the dominant kernel alone (consume), CUDA events on the launching stream
    tab = ops.AggTable(local, 1024, wc.NEED_SUM)
    tab.set_key_range(*rng)
    ts = []
    for _ in range(6):
        tab.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tab.consume(table, "price[idx]", "quantity[idx]"); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    tab.close()
    kernel_ms = sum(ts[1:]) / (len(ts) - 1)
    # result: every group's fp64 sum recomputed with torch (index_add_ in chunks), all-reduced; 1e-6 relative
    ref = torch.zeros(G, dtype=torch.float64, device=dev)
    step_rows = 1 << 27
    for s in range(0, rows, step_rows):
        ref.index_add_(0, qty[s:s + step_rows].long(), price[s:s + step_rows].double())
    if world > 1:
        dist.all_reduce(ref)
    present = ref != 0
    ok = g == int(present.sum().item()) and bool(torch.equal(got_keys.long(), torch.nonzero(present).flatten()))
    if ok:
        want = ref[present].float().double()
        ok = bool(((got_vals - want).abs() <= 1e-6 * want.abs()).all().item())
    per = ms / args.steps
    kernel = "wdb_group_wp" if G <= 4096 else "wdb_group"
    rec = {"query": "SELECT SUM(price) FROM t GROUP BY quantity", "baseline_config": "configs[3]", "groups": G,
           "rows_total": total, "rows_per_gpu": rows, "scaling": "strong" if world > 1 else "n/a", "ms_per_step": per,
           "value": total / (per * 1e-3), "unit": "rows/s",
           "roofline": roofline(kernel, name, 8.0, rows, kernel_ms, peak, peak_src, "kernel_ms = the consume launch(es) alone; the step adds table reset, all-reduce and ordered export"),
           "local_ms": ms_local / args.steps, "merge_ms": max(per - ms_local / args.steps, 0.0),
           "collectives": [f"ncclAllReduce(sum, float64 x {rng[1] - rng[0] + 1}: direct-addressed partial sums" +
                           ("; per index slice of the table on a side stream, overlapped with the next slice's kernel)" if G > 4096 else ")")] if world > 1 else [],
           "merge_bytes_per_gpu": 8 * (rng[1] - rng[0] + 1) if world > 1 else 0,
           "gpu_launches": launches, "clocks": clocks, "result_checked": ok}
    del price, qty, ref
    return rec


def run_topk(name, args, rank, world, local, comm, comm1, sampler, peak, peak_src):
    import torch
    from warpdb_b200 import ops
    total = args.rows_total or TOTAL_ROWS[name]
    row0, rows = shard_rows(total, world, rank)
    dev = f"cuda:{local}"
    price = ops.synth_f32(rows, SEED + 5, 0.0, 1e6, row0, local)
    table = {"price": price}
    bufs = (torch.empty(5, dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.int64, device=dev))

    def step(c=comm):   # per-GPU register top-k + one all-gather of 5 candidates per GPU + one-warp selection; no host synchronisation
        c.topk(table, "discount(price[idx], 0.9f)", None, None, True, 5, 0, row_base=row0, out=bufs, sync=False)
    ms, launches, clocks = time_steps(step, args.steps, args.warmup, world, sampler)
    torch.cuda.synchronize()
    got = bufs[0].clone()
    ms_local = ms
    if world > 1:
        ms_local, _, _ = time_steps(lambda: step(comm1), args.steps, args.warmup, world)
    loc = torch.zeros(5, dtype=torch.float32, device=dev)
    step_rows = 1 << 28
    cands = [torch.topk(price[s:s + step_rows] * 0.9, min(5, price[s:s + step_rows].numel())).values for s in range(0, rows, step_rows)]
    loc = torch.topk(torch.cat(cands), 5).values
    if world > 1:
        import torch.distributed as dist
        allc = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(allc, loc)
        loc = torch.topk(torch.cat(allc), 5).values
    ok = int(bufs[1].item()) == 5 and bool(torch.equal(got, loc))
    per = ms / args.steps
    rec = {"query": "SELECT discount(price, 0.9) FROM t ORDER BY discount(price, 0.9) DESC LIMIT 5 (custom.cu UDF)", "baseline_config": "configs[4]",
           "rows_total": total, "rows_per_gpu": rows, "scaling": "strong" if world > 1 else "n/a", "ms_per_step": per,
           "value": total / (per * 1e-3), "unit": "rows/s",
           "roofline": roofline("wdb_topk_scan", name, 4.0, rows, ms_local / args.steps, peak, peak_src,
                                "kernel_ms = the local step timed with CUDA events: wdb_topk_scan over the first 2^20 rows (threshold pre-pass) + "
                                "wdb_topk_scan over the shard with the final selection and SELECT evaluation fused into its last CTA"),
           "merge_ms": max(per - ms_local / args.steps, 0.0),
           "collectives": ["ncclAllGather(80 B per rank: 5 x (key f32, value f32, global row i64))"] if world > 1 else [],
           "gpu_launches": launches, "clocks": clocks, "result_checked": ok}
    del price
    return rec


RUNNERS = {"filter1": run_filter, "filter50": run_filter, "filter99": run_filter, "group1k": run_group, "group10m": run_group, "topk5": run_topk}


# ------------------------------------------------------------------------------------------------
# headline: projection, e2e, baselines
# ------------------------------------------------------------------------------------------------
def e2e_projection(rows, steps, warmup, world, rank, local):
    """Same metric through the host-buffer entry point (wdb_multi_project_filter_host: the
    run_multi_gpu_jit_host replacement): pinned host columns in, host floats out, every step.  ONE
    process (rank 0) drives all N GPUs with one call, the shape of the reference's function; the
    other ranks have released their GPU memory and wait."""
    import torch
    from warpdb_b200 import _core as wc
    from warpdb_b200 import ops
    torch.cuda.empty_cache()
    barrier(world)
    res = None
    if rank == 0:
        n = rows * world
        hp = torch.empty(n, dtype=torch.float32, pin_memory=True)
        hq = torch.empty(n, dtype=torch.int32, pin_memory=True)
        ho = torch.empty(n, dtype=torch.float32, pin_memory=True)
        chunk = 1 << 27
        for s in range(0, n, chunk):   # synthesise on the device once, keep pinned host copies as "the user's data"
            m = min(chunk, n - s)
            hp[s:s + m].copy_(ops.synth_f32(m, SEED + 2, 0.0, 100.0, s, local))
            hq[s:s + m].copy_(ops.synth_i32(m, SEED + 102, 1, 101, s, local))
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        cols, nc = wc.make_cols([("price", wc.FLOAT32, hp.data_ptr(), n), ("quantity", wc.INT32, hq.data_ptr(), n)])
        cnt = C.c_int64(0)

        def step():
            wc.check(wc.lib().wdb_multi_project_filter_host(world, None, cols, nc, b"((price[idx] * quantity[idx]) * 1.08f)", b"",
                                                            ho.data_ptr(), n, wc.DENSE_ZERO, C.byref(cnt)))
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
        idx = torch.randint(0, n, (1 << 20,))
        ok = bool(torch.equal(ho[idx], (hp[idx] * hq[idx].float()) * 1.08)) and bool(torch.equal(ho[-(1 << 16):], (hp[-(1 << 16):] * hq[-(1 << 16):].float()) * 1.08))
        yard = pcie_yardstick(hp, ho, world)   # after the check: it overwrites the head of ho
        res = dict(value=n * steps / dt, unit="rows/s", h2d_bytes_per_step=8 * n, d2h_bytes_per_step=4 * n,
                   steps=steps, ms_per_step=dt / steps * 1e3, result_checked=ok, pcie_gbs=12.0 * n * steps / dt / 1e9,
                   pcie_yardstick=yard,
                   pcie_bound_rows_per_s=(min(yard["duplex_h2d_gbs"] / 8.0, yard["duplex_d2h_gbs"] / 4.0) * 1e9 if "duplex_h2d_gbs" in yard else None),
                   api=f"one call of wdb_multi_project_filter_host(ndev={world}) per step from ONE process (run_multi_gpu_jit_host replacement), pinned host buffers, "
                       "wall clock around the calls (the call returns when the results are on the host)")
        del hp, hq, ho
    barrier(world)
    return res


def pcie_yardstick(h_src, h_dst, ndev, mb=1024):
    """What the box's PCIe links deliver with plain cudaMemcpyAsync from / to the same pinned buffers on all
    ndev GPUs at once (one process, one stream per direction and device): H2D alone, D2H alone, both together.
    The projection moves 8 B/row up and 4 B/row down, so its end-to-end ceiling is
    min(duplex_h2d / 8, duplex_d2h / 4) rows/s (`pcie_bound_rows_per_s`)."""
    import torch
    try:
        m = mb * (1 << 20) // 4
        devs = list(range(ndev))
        dbuf = [(torch.empty(m, dtype=torch.float32, device=f"cuda:{d}"), torch.empty(m, dtype=torch.float32, device=f"cuda:{d}")) for d in devs]
        up = [torch.cuda.Stream(device=d) for d in devs]
        down = [torch.cuda.Stream(device=d) for d in devs]

        def run(do_up, do_down, reps=3):
            for d in devs:
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            for _ in range(reps):
                for d in devs:
                    if do_up:
                        with torch.cuda.stream(up[d]):
                            dbuf[d][0].copy_(h_src[d * m:(d + 1) * m], non_blocking=True)
                    if do_down:
                        with torch.cuda.stream(down[d]):
                            h_dst[d * m:(d + 1) * m].copy_(dbuf[d][1], non_blocking=True)
            for d in devs:
                torch.cuda.synchronize(d)
            return reps * ndev * m * 4 / (time.perf_counter() - t0) / 1e9
        run(True, True, 1)
        h2d, d2h, both = run(True, False), run(False, True), run(True, True)
        return {"h2d_gbs": h2d, "d2h_gbs": d2h, "duplex_h2d_gbs": both, "duplex_d2h_gbs": both, "n_gpus": ndev, "mb_per_copy": mb,
                "how": "torch copy_(non_blocking) of pinned slices, all GPUs concurrently from one process, 3 repetitions, aggregate GB/s per direction"}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}


def cpu_baseline_projection(sample_rows, steps=1):
    """Scalar AST evaluation on the host cores (oracle port of src/warpdb.cpp:128-151 semantics)."""
    from oracle import pyoracle as orc
    cores = os.cpu_count() or 1
    table = {"price": orc.synth_f32(sample_rows, SEED + 2, 0.0, 100.0), "quantity": orc.synth_i32(sample_rows, SEED + 102, 1, 101)}
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
    # stdout carries exactly ONE line, the JSON: anything libraries print on fd 1 meanwhile (NCCL's version banner
    # under NCCL_DEBUG, for one) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = dist_setup()
    from warpdb_b200 import _core as wc, ops
    wc.check(wc.lib().wdb_init(local))
    wc.set_udf_source(UDF)
    for k, v in os.environ.items():   # tuning experiments: WARPDB_OPT_group__dense_waves=1 -> option group.dense_waves
        if k.startswith("WARPDB_OPT_"):
            wc.set_option(k[len("WARPDB_OPT_"):].replace("__", "."), int(v))
    peak, peak_src = measured_peak()
    sampler = ClockSampler(local) if rank == 0 else None
    rows = args.rows or PROJECTION_ROWS_PER_GPU
    row0 = rank * rows                       # contiguous row-range shard of the global table (multi_gpu_utils.cpp:24-31)
    price = ops.synth_f32(rows, SEED + 2, 0.0, 100.0, row0, local)
    qty = ops.synth_i32(rows, SEED + 102, 1, 101, row0, local)
    out = torch.empty(rows, dtype=torch.float32, device=f"cuda:{local}")
    table = {"price": price, "quantity": qty}
    expr = "((price[idx] * quantity[idx]) * 1.08f)"
    step = lambda: ops.project_filter(table, expr, None, wc.DENSE, out=out, sync_count=False)  # noqa: E731
    step()
    torch.cuda.synchronize()
    compile_ms = wc.stats()["last_compile_ms"]
    ms, launches, clocks = time_steps(step, args.steps, args.warmup, world, sampler)
    ok = bool(torch.equal(out[:1 << 20], ((price[:1 << 20] * qty[:1 << 20].float()) * 1.08))) and \
        bool(torch.equal(out[-(1 << 20):], ((price[-(1 << 20):] * qty[-(1 << 20):].float()) * 1.08)))
    ms_per_step = ms / args.steps
    line = {
        "metric": metric_name(), "value": world * rows / (ms_per_step * 1e-3), "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (counter-based generator, identical on the CPU oracle)",
        "config": {"workload": "projection", "query": "price * quantity * 1.08",
                   "description": "price * quantity * 1.08 over 1e9 rows/GPU (f32 price, i32 quantity): BASELINE configs[1]; configs[2..4] are under `workloads`",
                   "rows_per_gpu": rows, "sharding": "contiguous row ranges, one process per GPU, no data-path collective",
                   "l2": "inputs (>= 4 GB per column) are far larger than the 126 MB L2; no flush needed",
                   "algorithmic_bytes_per_row": 12.0, "result_checked": ok},
        "roofline": roofline("wdb_project", "projection", 12.0, rows, ms_per_step, peak, peak_src),
        "gpu_launches": launches, "clocks": clocks, "nvrtc_compile_ms_untimed": compile_ms,
    }
    line["gbs_per_gpu"] = line["roofline"]["achieved"]
    if rank == 0 and world == 1 and not args.no_ref:
        line["ref_nvrtc"] = ref_nvrtc_projection(table, rows, local)
    del price, qty, out, table, step
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, merges inside the step --------------------------------------
    names = [w for w in (args.workloads.split(",") if args.workloads else []) if w]
    if names:
        comm = ops.Comm.from_torch(local)       # libwarpcore's own NCCL communicator (one rank per GPU)
        comm1 = ops.Comm(local, 0, 1)           # the same entry points without collectives: local_ms / merge_ms
        recs = {}
        for w in names:
            recs[w] = RUNNERS[w](w, args, rank, world, local, comm, comm1, sampler, peak, peak_src)
            torch.cuda.empty_cache()
        line["workloads"] = recs
        comm.close()
        comm1.close()
    if not args.no_e2e:
        e = e2e_projection(rows, max(1, min(args.steps, args.e2e_steps)), 1, world, rank, local)
        if e:
            line["e2e"] = e
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_projection(args.cpu_rows)
    if sampler:
        sampler.close()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
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
    table = {"price": orc.synth_f32(sample, SEED + 2, 0.0, 100.0), "quantity": orc.synth_i32(sample, SEED + 102, 1, 101)}
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
        "config": {"workload": "projection", "query": "price * quantity * 1.08",
                   "description": "price * quantity * 1.08 over 1e9 rows/GPU (f32 price, i32 quantity): BASELINE configs[1]",
                   "rows_per_step": sample},
        "cpu_baseline": cb, "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workloads", default=",".join(ALL_WORKLOADS), help="comma separated subset of " + ",".join(ALL_WORKLOADS) + " ('' = headline only)")
    ap.add_argument("--rows", type=int, default=0, help="projection rows per GPU (default 1e9)")
    ap.add_argument("--rows-total", type=int, default=0, help="override the total rows of every sub-workload (tests)")
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
