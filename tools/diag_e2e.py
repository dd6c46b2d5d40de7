#!/usr/bin/env python
"""End-to-end (host buffers) path: sweep ring depth / chunk size of wdb_multi_project_filter_host."""
import sys, os, json, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from warpdb_b200 import _core as wc, ops
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
wc.check(wc.lib().wdb_init(0))
hp = torch.empty(n, dtype=torch.float32, pin_memory=True); hq = torch.empty(n, dtype=torch.int32, pin_memory=True)
ho = torch.empty(n, dtype=torch.float32, pin_memory=True)
chunk = 1 << 27
for s in range(0, n, chunk):
    m = min(chunk, n - s)
    hp[s:s + m].copy_(ops.synth_f32(m, 0xC0FFEE + 2, 0.0, 100.0, s)); hq[s:s + m].copy_(ops.synth_i32(m, 0xC0FFEE + 102, 1, 101, s))
torch.cuda.synchronize()
cols, nc = wc.make_cols([("price", wc.FLOAT32, hp.data_ptr(), n), ("quantity", wc.INT32, hq.data_ptr(), n)])
cnt = C.c_int64(0); devs = (C.c_int * 1)(0)
d = torch.empty(n, dtype=torch.float32, device="cuda")
for name, fn in (("H2D 4GB", lambda: d.copy_(hp, non_blocking=True)), ("D2H 4GB", lambda: ho.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(json.dumps({"name": name, "gbs": 4 * n / dt / 1e9}), flush=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(hp, non_blocking=True)
with torch.cuda.stream(s2): ho.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(json.dumps({"name": "H2D 4GB || D2H 4GB concurrently", "ms": dt * 1e3, "gbs_each": 4 * n / dt / 1e9}), flush=True)
del d
for rows_log2, slots in ((22, 3), (24, 2), (24, 3), (24, 4), (26, 3), (26, 4), (27, 2)):
    wc.set_option("multi.chunk_rows", 1 << rows_log2); wc.set_option("multi.slots", slots)
    def step():
        wc.check(wc.lib().wdb_multi_project_filter_host(1, devs, cols, nc, b"((price[idx] * quantity[idx]) * 1.08f)", b"", ho.data_ptr(), n, wc.DENSE_ZERO, C.byref(cnt)))
    step(); t = time.perf_counter(); step(); step(); dt = (time.perf_counter() - t) / 2
    print(json.dumps({"name": "wdb_multi_project_filter_host", "chunk_rows": 1 << rows_log2, "slots": slots, "ms": dt * 1e3, "rows_per_s": n / dt, "pcie_gbs": 12 * n / dt / 1e9}), flush=True)
