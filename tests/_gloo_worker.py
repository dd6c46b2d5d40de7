"""Worker of tests/test_sharded_gloo.py: one rank of a world_size-N gloo group on CPU.  The per-rank
work is done by a NumPy/oracle backend (test infrastructure) so that warpdb_b200/sharded.py's
sharding and collective logic runs without a GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pyoracle as orc  # noqa: E402
from warpdb_b200 import _core as wc  # noqa: E402
from warpdb_b200.sharded import ShardedDB, shard_range  # noqa: E402

CUDA2TEXT = {"(price[idx] * 0.9f)": "price * 0.9", "(price[idx] > 20.0f)": "price > 20", "price[idx]": "price",
             "quantity[idx]": "quantity", "((price[idx] * quantity[idx]) * 1.08f)": "price * quantity * 1.08",
             "k[idx]": "k", "v[idx]": "v", None: None, "(price[idx] > 1000000.0f)": "price > 1000000",
             "(quantity[idx] < (0.0f - 40.0f))": "quantity < 0 - 40", "(quantity[idx] * 0)": "quantity * 0"}


def npy(table):
    return {k: v.numpy() for k, v in table.items()}


class OracleBackend:
    def project_filter(self, table, expr, cond, mode):
        t = npy(table)
        if mode == wc.COMPACT:
            out = orc.filter_compact(CUDA2TEXT[expr], CUDA2TEXT[cond], t)
            return torch.from_numpy(out), len(out)
        out, _ = orc.project_filter(CUDA2TEXT[expr], CUDA2TEXT[cond], t, fill=0.0)
        return torch.from_numpy(out), len(out)

    def group_partials(self, table, val, key, cond, needs, expected, row_base):
        g = orc.group_agg(CUDA2TEXT[val], CUDA2TEXT[key], CUDA2TEXT[cond], npy(table), agg=orc.SUM)
        mn = orc.group_agg(CUDA2TEXT[val], CUDA2TEXT[key], CUDA2TEXT[cond], npy(table), agg=orc.MIN)["vals"].astype(np.float64)
        mx = orc.group_agg(CUDA2TEXT[val], CUDA2TEXT[key], CUDA2TEXT[cond], npy(table), agg=orc.MAX)["vals"].astype(np.float64)
        return {"keys": torch.from_numpy(g["keys"]), "sums": torch.from_numpy(g["sums"]), "counts": torch.from_numpy(g["counts"]),
                "mins": torch.from_numpy(mn), "maxs": torch.from_numpy(mx)}

    def merge_partials(self, parts, needs, expected, agg, order):
        acc = {}
        for p in parts:
            for i, k in enumerate(p["keys"].tolist()):
                s, c, mn, mx = acc.get(k, (0.0, 0, np.inf, -np.inf))
                acc[k] = (s + float(p["sums"][i]), c + int(p["counts"][i]), min(mn, float(p["mins"][i])), max(mx, float(p["maxs"][i])))
        keys = sorted(acc, reverse=(order == wc.ORDER_KEY_DESC))
        sums = np.array([acc[k][0] for k in keys]); cnts = np.array([acc[k][1] for k in keys], np.int64)
        mins = np.array([acc[k][2] for k in keys]); maxs = np.array([acc[k][3] for k in keys])
        vals = {wc.SUM: sums, wc.AVG: sums / np.maximum(cnts, 1), wc.COUNT: cnts.astype(np.float64), wc.MIN: mins, wc.MAX: maxs}[agg]
        return {"keys": torch.tensor(keys, dtype=torch.int32), "vals": torch.from_numpy(vals.astype(np.float32)),
                "sums": torch.from_numpy(sums), "counts": torch.from_numpy(cnts), "mins": torch.from_numpy(mins), "maxs": torch.from_numpy(maxs)}

    def topk_local(self, table, key, val, cond, descending, k):
        t = npy(table)
        keys = orc.filter_compact(CUDA2TEXT[key], CUDA2TEXT[cond], t)
        vals = orc.filter_compact(CUDA2TEXT[val], CUDA2TEXT[cond], t)
        order = np.argsort(-keys if descending else keys, kind="stable")[:k]
        return torch.from_numpy(vals[order]), torch.from_numpy(keys[order])

    def topk_merge(self, vals, keys, descending, k, offset, valid=None):
        if valid is not None:
            keep = valid.numpy() > 0.5
            vals, keys = vals[torch.from_numpy(keep)], keys[torch.from_numpy(keep)]
        kk = keys.numpy()
        order = np.argsort(-kk if descending else kk, kind="stable")[offset:offset + k]
        return vals[torch.from_numpy(order)]


def main():
    rank, world, port, n = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port
    dist.init_process_group("gloo", rank=rank, world_size=world)
    price = orc.synth_f32(n, 77, 0.0, 40.0)
    qty = orc.synth_i32(n, 78, -50, 1950)
    s, e = shard_range(n, world, rank)
    assert (s, e) == orc.shard_range(n, world, rank)
    table = {"price": torch.from_numpy(price[s:e].copy()), "quantity": torch.from_numpy(qty[s:e].copy())}
    full = {"price": price, "quantity": qty}
    db = ShardedDB(table, n, rank, world, backend=OracleBackend())

    # dense projection: no collective; gather=True concatenates in rank order == row order
    ref, _ = orc.project_filter("price * quantity * 1.08", None, full)
    got = db.query("((price[idx] * quantity[idx]) * 1.08f)", None, gather=True).numpy()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # compaction: local survivors + scanned global offset
    refc = orc.filter_compact("price * 0.9", "price > 20", full)
    loc, off, total = db.query_compact("(price[idx] * 0.9f)", "(price[idx] > 20.0f)")
    assert total == len(refc) and np.array_equal(loc.numpy(), refc[off:off + len(loc)])
    # GROUP BY with both merge strategies
    for agg in (wc.SUM, wc.AVG, wc.COUNT, wc.MIN, wc.MAX):
        r = orc.group_agg("price", "quantity", "price > 20", full, agg=agg)
        for strat in ("allgather", "exchange"):
            g = db.group_agg("price[idx]", "quantity[idx]", "(price[idx] > 20.0f)", agg=agg, strategy=strat)
            assert np.array_equal(g["keys"].numpy(), r["keys"]), (agg, strat)
            np.testing.assert_allclose(g["vals"].numpy(), r["vals"], rtol=1e-6)
        gd = db.group_agg("price[idx]", "quantity[idx]", None, agg=agg, order=wc.ORDER_KEY_DESC, strategy="exchange")
        assert np.array_equal(gd["keys"].numpy(), orc.group_agg("price", "quantity", None, full, agg=agg, order=orc.ORDER_KEY_DESC)["keys"])
    # range-partitioned exchange on awkward inputs: no group anywhere; groups on one rank only (the other
    # ranks send empty pieces); one group; a skewed key range (most keys in the first owner's slice)
    g = db.group_agg("price[idx]", "quantity[idx]", "(price[idx] > 1000000.0f)", agg=wc.SUM, strategy="exchange")
    assert g["keys"].numel() == 0
    r = orc.group_agg("price", "quantity", "quantity < 0 - 40", full, agg=wc.AVG)
    g = db.group_agg("price[idx]", "quantity[idx]", "(quantity[idx] < (0.0f - 40.0f))", agg=wc.AVG, strategy="exchange")
    assert np.array_equal(g["keys"].numpy(), r["keys"]) and np.array_equal(g["counts"].numpy(), r["counts"])
    r = orc.group_agg("price", "quantity * 0", None, full, agg=wc.COUNT)
    g = db.group_agg("price[idx]", "(quantity[idx] * 0)", None, agg=wc.COUNT, strategy="exchange")
    assert g["keys"].tolist() == [0] and np.array_equal(g["vals"].numpy(), r["vals"])
    # ORDER BY ... LIMIT: ties (quantity has many) keep global row order
    for desc in (True, False):
        want = orc.query_sql(f"SELECT price FROM t ORDER BY quantity {'DESC' if desc else 'ASC'} LIMIT 7 OFFSET 2", full)
        got = db.topk("quantity[idx]", "price[idx]", None, desc, 7, 2).numpy()
        assert np.array_equal(got, want), (desc, got, want)
        want = orc.topk("price", "price > 20", full, descending=desc, k=5)
        assert np.array_equal(db.topk("price[idx]", None, "(price[idx] > 20.0f)", desc, 5).numpy(), want)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
