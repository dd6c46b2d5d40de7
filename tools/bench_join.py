#!/usr/bin/env python
"""Inner equi-join micro-benchmark on one B200 (run under gpurun); JSON lines to gpurun_out/r02_join.jsonl.

Fact-to-dimension shape: n probe keys drawn from [0, 1.1 m) against a build side of m distinct ids
(10 % of the probe rows find no partner).  Algorithmic bytes: count = 4 n (one read of the probe keys);
probe = 8 n + 16 pairs (the keys twice, two int64 row numbers per pair); gather = 12 per row (row number
in, one 4-byte value out).  Every timed result is checked (pairs found, keys equal, order).
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


def time_op(fn, iters=5, warmup=1):
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


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28
    f = open(os.path.join(OUT, "r02_join.jsonl"), "a")

    def emit(rec):
        rec["peak_gbs"] = PEAK
        f.write(json.dumps(rec) + "\n")
        f.flush()
        print(json.dumps(rec), flush=True)

    wc.check(wc.lib().wdb_init(0))
    for m in ([int(float(a)) for a in sys.argv[2:]] or [1 << 20, 1 << 24]):
        ids = torch.randperm(m, dtype=torch.int32, device="cuda")
        rate = torch.rand(m, device="cuda")
        probe = ops.synth_i32(n, 0xC0FFEE + 7, 0, m + m // 10)
        want_pairs = int((probe < m).sum().item())
        box = {}

        def build():
            if "ix" in box:
                box["ix"].close()
            box["ix"] = ops.JoinIndex(ids)
        ms = time_op(build, iters=3)
        ix = box["ix"]
        emit({"op": "build", "m": m, "ms": ms, "direct_span": ix.direct_span})
        ref = None
        for direct in (1, 0):
            wc.set_option("join.direct", direct)
            cnt = ix.count(probe)
            ms_count = time_op(lambda: ix.count(probe))
            pr, br = ix.probe(probe)
            ok = cnt == want_pairs and pr.numel() == want_pairs
            ok = ok and bool(torch.equal(ids[br].long(), probe[pr].long())) and bool((pr[1:] > pr[:-1]).all().item())
            if ref is None:
                ref = (pr, br)
            else:
                ok = ok and bool(torch.equal(ref[0], pr)) and bool(torch.equal(ref[1], br))
            out_p, out_b = torch.empty_like(pr), torch.empty_like(br)
            cols, _ = wc.make_cols([("probe", wc.INT32, probe.data_ptr(), n)])
            import ctypes as C
            got = C.c_int64(0)

            def full():
                wc.check(wc.lib().wdb_join_probe(ix.handle, None, cols, out_p.data_ptr(), out_b.data_ptr(), want_pairs, C.byref(got)))
            ms_full = time_op(full)          # the warm-up call consumes the counts left by count(); the timed calls count themselves

            def count_then_emit():           # what a caller that does not know the result size does: the emit reuses the count
                ix.count(probe)
                full()
            ms_both = time_op(count_then_emit)
            ok = ok and bool(torch.equal(out_p, pr)) and bool(torch.equal(out_b, br))
            emit({"op": "probe", "n": n, "m": m, "direct": direct, "pairs": cnt, "ok": ok,
                  "count_ms": ms_count, "count_gbs": 4.0 * n / ms_count / 1e6, "count_frac": 4.0 * n / ms_count / 1e6 / PEAK,
                  "probe_ms": ms_full, "probe_gbs": (8.0 * n + 16.0 * cnt) / ms_full / 1e6, "probe_frac": (8.0 * n + 16.0 * cnt) / ms_full / 1e6 / PEAK,
                  "probe_rows_per_s": n / ms_full * 1e3, "count_then_emit_ms": ms_both})
            del out_p, out_b
            if direct == 0:
                del pr, br
        wc.set_option("join.direct", None)
        pr, br = ref
        dst = torch.empty(br.numel(), dtype=torch.float32, device="cuda")
        ms = time_op(lambda: ops.gather(rate, br, out=dst))
        emit({"op": "gather", "rows": br.numel(), "src_rows": m, "ms": ms, "gbs": 12.0 * br.numel() / ms / 1e6, "frac": 12.0 * br.numel() / ms / 1e6 / PEAK,
              "ok": bool(torch.equal(dst, rate[br]))})
        ix.close()
        del ref, pr, br, dst, probe, ids, rate
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
