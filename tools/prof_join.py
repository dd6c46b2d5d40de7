#!/usr/bin/env python
"""Target for `ncu -k regex:join_(count|emit)`: one count + one emit of the benchmarked join shape
(2^28 probe keys against 2^20 distinct ids, 10 % of the rows without a partner)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from warpdb_b200 import _core as wc, ops  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28
m = 1 << 20
wc.check(wc.lib().wdb_init(0))
ids = torch.randperm(m, dtype=torch.int32, device="cuda")
probe = ops.synth_i32(n, 0xC0FFEE + 7, 0, m + m // 10)
ix = ops.JoinIndex(ids)
pr, br = ix.probe(probe)          # count (pass 1), then the emit that reuses it
torch.cuda.synchronize()
print("pairs", pr.numel(), "ok", bool(torch.equal(ids[br].long(), probe[pr].long())))
