#!/usr/bin/env python
"""query_multi_gpu_csv on a generated CSV: how much of the GPU work hides behind the parse of the next chunk
(SURVEY 8(f3)); checks the result against NumPy."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from warpdb_b200 import build as wbuild  # noqa: E402

wbuild.build_host()
from warpdb_b200 import pywarpdb  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
path = "/tmp/wdb_stream.csv"
rng = np.random.default_rng(1)
price = rng.uniform(0, 100, n).astype(np.float32)
qty = rng.integers(1, 100, n).astype(np.int32)
with open(path, "w") as f:
    f.write("price,quantity\n")
    f.write("\n".join(f"{p:.4f},{q}" for p, q in zip(price.tolist(), qty.tolist())))
    f.write("\n")
price = np.array([float(f"{p:.4f}") for p in price.tolist()], np.float32)   # what the parser will read
for chunk in (n, n // 4, n // 16, n // 64):
    t0 = time.perf_counter()
    out = pywarpdb.WarpDB.query_multi_gpu_csv(path, "price * quantity WHERE price > 50", chunk)
    dt = (time.perf_counter() - t0) * 1e3
    st = pywarpdb.WarpDB.last_csv_stream_stats()
    want = np.where(price > 50, price * qty.astype(np.float32), np.float32(0))
    ok = bool(np.array_equal(np.array(out, np.float32), want))
    print(json.dumps({"rows": n, "rows_per_chunk": chunk, "wall_ms": dt, "parse_ms": st["parse_ms"], "gpu_ms": st["gpu_ms"], "chunks": st["chunks"],
                      "hidden_ms": st["parse_ms"] + st["gpu_ms"] - st["wall_ms"], "ok": ok}), flush=True)
